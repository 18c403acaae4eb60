#!/usr/bin/env python3
"""Build-container half of the fold-0 BPR/VBPR acceptance: runs the UNMODIFIED reference evaluate.py (subprocess, its own
numpy code) on the models the B200 run trained (gpurun_out/fold0/fold0_models.tar.gz) and compares its printed lines with
the lines the GPU evaluator printed for the same files.  final-B.dat is left out for the reference script: its bias line
(evaluate.py:79-80) raises whenever n_te != n_items (SURVEY D-4); the with-bias lists are checked against the oracle on
the GPU box instead.  usage: python profiles/fold0_bpr_check.py gpurun_out/fold0 profiles/r02_fold0_bpr.json"""
import json
import os
import subprocess
import sys
import tarfile
import tempfile
import time

REF = "/root/reference"


def main():
    src, dst = sys.argv[1], sys.argv[2]
    rec = json.load(open(os.path.join(src, "fold0_bpr.json")))
    with tempfile.TemporaryDirectory() as td:
        with tarfile.open(os.path.join(src, "fold0_models.tar.gz")) as tf:
            tf.extractall(td)
        for name in ("bpr", "vbpr"):
            os.remove(os.path.join(td, name, "final-B.dat"))
            t0 = time.time()
            p = subprocess.run([sys.executable, os.path.join(REF, "evaluate.py"), "-d", os.path.join(REF, "data"), "-m",
                                os.path.join(td, name), "-f", "0", "-sl", "im", "om"], capture_output=True, text=True, cwd=REF)
            lines = [ln for ln in p.stdout.splitlines() if ln.startswith(("im,", "om,"))]
            ours = rec[name]["evaluate"]["no_bias"]["gpu_lines"]
            diff = [max(abs(float(a) - float(b)) for a, b in zip(x.split(",")[1:], y.split(",")[1:])) for x, y in zip(lines, ours)]
            rec[name]["reference_evaluate_py"] = {"stdout_lines": lines, "seconds": time.time() - t0, "returncode": p.returncode,
                                                  "equal_to_gpu_lines": lines == ours, "max_abs_diff_per_scenario": diff,
                                                  "note": "unmodified /root/reference/evaluate.py on the same .dat files without final-B.dat"}
            print(name, lines, ours, lines == ours)
    json.dump(rec, open(dst, "w"), indent=1)


if __name__ == "__main__":
    main()
