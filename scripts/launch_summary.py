#!/usr/bin/env python
"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals of the LAST binned-SAH build."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, seq = None, []
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        try:
            v = float(d["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        v *= {"us": 1e-3, "ns": 1e-6, "ms": 1.0, "s": 1e3}.get(d["Metric Unit"], 1.0)
        seq.append((re.sub(r"\(.*", "", d["Kernel Name"])[:64], v))
starts = [i for i, (n, _) in enumerate(seq) if "world_init" in n]
s = starts[-1] if starts else 0
e = next((i for i in range(s, len(seq)) if "parents_kernel" in seq[i][0]), len(seq))
agg = collections.defaultdict(lambda: [0, 0.0])
for n, v in seq[s:e]:
    agg[n][0] += 1
    agg[n][1] += v
for k, v in sorted(agg.items(), key=lambda x: -x[1][1])[:14]:
    print(f"{v[1]:9.3f} ms {v[0]:5d}  {k}")
print(f"{sum(v[1] for v in agg.values()):9.3f} ms total")
for key in sys.argv[2:]:
    print(key, [round(v * 1e3) for n, v in seq[s:e] if key in n])
