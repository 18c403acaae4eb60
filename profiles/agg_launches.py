import csv, sys
from collections import defaultdict
rows = list(csv.reader(open(sys.argv[1])))
st = next(k for k, r in enumerate(rows) if "Kernel Name" in r)
h = rows[st]; ki, vi = h.index("Kernel Name"), h.index("Metric Value")
agg = defaultdict(list)
for r in rows[st + 1:]:
    if len(r) > vi:
        try: agg[r[ki].split("(")[0][:60]].append(float(r[vi].replace(",", "")) / 1e3)
        except ValueError: pass
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])): print("%-62s n=%3d mean %9.1f us  last %9.1f" % (k, len(v), sum(v) / len(v), v[-1]))
