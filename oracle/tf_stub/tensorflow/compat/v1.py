"""Fake ``tensorflow.compat.v1``: only what ``single/*.py`` touches at import
and constructor time (``rec.py:15``, ``bpr.py:27-28``, type annotations such as
``mlp.py:24``).  Any attempt to *use* TensorFlow raises."""


class _GpuOptions:
    allow_growth = False


class ConfigProto:
    def __init__(self):
        self.gpu_options = _GpuOptions()


def disable_eager_execution():
    return None


class _Missing:
    def __init__(self, name):
        self._name = name

    def __call__(self, *a, **k):
        raise RuntimeError("tensorflow stub: tf.%s needs TensorFlow, which is not installed" % self._name)

    def __getattr__(self, name):
        return _Missing(self._name + "." + name)


def __getattr__(name):
    return _Missing(name)
