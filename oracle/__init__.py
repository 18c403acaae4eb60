"""CPU oracle for the two hot paths of domainxz/top-k-rec.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and there only as the
checker / the timed CPU baseline.  The product path (``top-k-rec_b200/``)
never imports this package and fails loudly when its CUDA library is missing.

Pinning status
--------------
* path 2 (score + filtered top-k, ``evaluate.py``): PINNED.  ``evaluate.py`` is
  numpy-only and runs in the build container; ``tests/golden/make_golden.py``
  ran the unmodified script and committed its printed numbers, and
  ``oracle.evaluate_ref`` must reproduce them.
* sampler / loaders / ``.dat`` codec: PINNED against the reference's own
  functions imported through ``oracle/tf_stub`` (a 20-line fake of
  ``tensorflow.compat.v1``, import-time only).
* ALS for WMF/CER (``single/cer.py:24-73``): PINNED.  ``CER.train`` is numpy-only once imported through
  ``oracle/tf_stub``; ``tests/golden/make_golden_als.py`` ran it unmodified and ``oracle.als_ref`` reproduces its
  factors bit for bit.
* path 1 step arithmetic (``single/bpr.py:71-101`` executed by TensorFlow 1.15,
  pinned ``tensorflow-gpu == 1.15.*`` in ``requirements.txt:3``, not vendored,
  not installable offline): **PARITY UNPINNED**.  ``oracle.bpr_ref`` restates
  the published TF-1.15 semantics and is cross-checked by
  fp64 finite differences of the objective, by hand-computed cases, and by
  ``oracle.tf_literal`` -- an independent derivation: the reference's graph
  lines under torch autograd + a statement-by-statement transcription of
  TF-1.15's duplicate-index plumbing and RMSProp kernels (momentum slot and
  all).  Transcribing ``vbpr.py:61`` literally exposed defect D-14 (the
  ``[n,1]`` bias variables broadcast x to ``[B,B]``); both readings are kept.
"""
