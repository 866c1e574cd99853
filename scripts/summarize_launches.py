"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) by kernel: launches, total us, share."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[start]
ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[start + 2:]:
    if len(r) <= iv:
        continue
    name = re.sub(r"<.*", "", re.sub(r"\(.*", "", r[ik])).replace("void ", "")
    try:
        v = float(r[iv].replace(",", ""))
    except ValueError:
        continue
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:72]:72s} {v[0]:5d} {v[1] / 1e3:10.1f} us {100 * v[1] / tot:5.1f}%")
print(f"total {tot / 1e6:.3f} ms in {sum(v[0] for v in agg.values())} launches")
