#!/usr/bin/env python3
"""Aggregate the warp-stall samples of an `ncu --page source --csv` export by code section (sections = the SASS between two
barrier instructions), with the barrier-stall samples (attributed to the instruction after a BAR) listed separately.
usage: ncu -i X.ncu-rep --page source --csv > src.csv; python profiles/ncu_sections.py src.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
h, data = rows[1], rows[2:]
S, I, SRC, BARR = h.index("# Samples"), h.index("Instructions Executed"), h.index("Source"), h.index("stall_barrier")
keys = [h.index(k) for k in ("stall_long_sb", "stall_short_sb", "stall_wait", "stall_math", "stall_dispatch", "stall_no_inst", "stall_not_selected", "stall_selected", "stall_mio")]
tot = sum(int(r[S]) for r in data)
print(rows[0][1], "| samples", tot, "| warp instructions", sum(int(r[I]) for r in data), "| barrier samples", sum(int(r[BARR]) for r in data))
marks = [n for n, r in enumerate(data) if "BAR." in r[SRC]]
prev = 0
for m in marks + [len(data) - 1]:
    seg = data[prev:m + 1]
    nb = sum(int(r[S]) - int(r[BARR]) for r in seg); b = sum(int(r[BARR]) for r in seg)
    ins = sum(int(r[I]) for r in seg); ff = sum(int(r[I]) for r in seg if "FFMA" in r[SRC])
    br = {h[k][6:]: sum(int(r[k]) for r in seg) for k in keys}
    if nb + b > 0.002 * tot:
        print("lines %5d-%5d  work %5.1f%%  barrier-wait %5.1f%%  instr %7.1fM (ffma %7.1fM)  %s | ends: %s" % (
            prev, m, 100.0 * nb / tot, 100.0 * b / tot, ins / 1e6, ff / 1e6, {k: round(100.0 * v / tot, 1) for k, v in br.items() if v > 0.01 * tot},
            data[m][SRC].strip()[:36]))
    prev = m + 1
