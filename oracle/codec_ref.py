"""``.dat`` text codec of the reference, restated.  TEST INFRASTRUCTURE.

Writer ``utils.py:47-55``: every element as ``'%f '`` (6 decimals + one space),
one matrix row per line.  Reader ``utils.py:28-44`` / ``evaluate.py:19-28``:
``line.strip().split(' ')`` -> ``np.float32`` per token, row r of the file is
the id on line r of the id file.
"""
from __future__ import annotations

import numpy as np


def dat_bytes(embed):
    embed = np.asarray(embed)
    assert embed.ndim == 2
    return "".join("".join("%f " % v for v in row) + "\n" for row in embed).encode()


def dat_parse(data):
    rows = [line.strip().split(" ") for line in data.decode().splitlines()]
    return np.array([[np.float32(t) for t in r] for r in rows], np.float32)
