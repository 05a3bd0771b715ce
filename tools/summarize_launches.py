"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
hdr, data = rows[hi], rows[hi + 1:]
kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, agg = 0.0, collections.defaultdict(lambda: [0, 0.0])
for r in data:
    if len(r) <= mv:
        continue
    v = float(r[mv].replace(",", ""))
    v = v / 1000 if r[mu] == "ns" else v * 1000 if r[mu] == "ms" else v
    short = re.sub(r"hd::(tc::|simt::)?", "", re.sub(r"\(.*", "", r[kn])).replace("void ", "")
    agg[short][0] += 1
    agg[short][1] += v
    tot += v
print(f"# {sys.argv[1]}: {sum(a[0] for a in agg.values())} launches, {tot:.1f} us total (cold-cache, serialised: compare shares)")
print(f"{'us':>10} {'share':>6} {'n':>4} {'avg us':>8}  kernel")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t:10.1f} {100 * t / tot:5.1f}% {n:4d} {t / n:8.1f}  {k}")
