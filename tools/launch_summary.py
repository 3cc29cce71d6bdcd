#!/usr/bin/env python
"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: launches, total time, share."""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
hdr = rows[hi]
kn, mv, mn = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= mv or r[mn] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r[kn])
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", name)[:110]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += float(r[mv].replace(",", ""))
tot = sum(a[1] for a in agg.values())
print(f"{'total ms':>10} {'launches':>8} {'share':>6}  kernel   (ncu per-launch times are serialised and cold-cache: shares, not absolutes)")
for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{t / 1e6:10.3f} {c:8d} {100 * t / tot:5.1f}%  {n}")
print(f"{tot / 1e6:10.3f} {sum(a[0] for a in agg.values()):8d} 100.0%  all")
