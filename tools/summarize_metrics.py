"""Summarise a multi-metric ncu launch list (--csv, one row per launch x metric) by kernel family.

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,\
gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,\
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,\
sm__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none --profile-from-start off \
        --csv --log-file launches.csv python tools/profile_sampler_step.py 256
    python tools/summarize_metrics.py launches.csv > launches.summary.txt
"""
import collections
import csv
import json
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
hdr, data = rows[hi], rows[hi + 1:]
ci = {k: hdr.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value")}
launches = collections.OrderedDict()
for r in data:
    if len(r) <= ci["Metric Value"]:
        continue
    d = launches.setdefault(r[ci["ID"]], {"k": r[ci["Kernel Name"]]})
    v = float(r[ci["Metric Value"]].replace(",", ""))
    u = r[ci["Metric Unit"]]
    if r[ci["Metric Name"]] == "gpu__time_duration.sum":
        v = v / 1000 if u in ("ns", "nsecond") else v * 1000 if u in ("ms", "msecond") else v
    if u in ("Kbyte", "Mbyte", "Gbyte"):
        v *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    d[r[ci["Metric Name"]]] = v

T, RD, WR, L2 = "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum"
TP = "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
DP = "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"
SP = "sm__throughput.avg.pct_of_peak_sustained_elapsed"
agg = collections.defaultdict(lambda: collections.defaultdict(float))
tot = 0.0
for d in launches.values():
    short = re.sub(r"hd::", "", re.sub(r"\(.*", "", d["k"])).replace("void ", "")
    a = agg[short]
    t = d.get(T, 0.0)
    a["n"] += 1
    a["t"] += t
    for m in (RD, WR, L2):
        a[m] += d.get(m, 0.0)
    for m in (TP, DP, SP):
        a[m] += d.get(m, 0.0) * t
    tot += t
n = len(launches)
rd = sum(a[RD] for a in agg.values())
wr = sum(a[WR] for a in agg.values())
print(f"# {sys.argv[1]}: {n} launches, {tot:.1f} us (cold caches, serialised: compare SHARES); "
      f"DRAM read {rd / 1e6:.1f} MB, write {wr / 1e6:.1f} MB")
print("# time-weighted averages per kernel family; dram% / tensor% / sm% are pct of peak sustained (ncu)")
print(f"{'kernel':72s} {'n':>4} {'us':>8} {'share':>6} {'tensor%':>8} {'dram%':>6} {'sm%':>6} {'DRAM rd MB':>10} {'wr MB':>7} {'L2 MB':>8}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["t"]):
    t = a["t"] or 1e-9
    print(f"{k[:72]:72s} {int(a['n']):4d} {a['t']:8.1f} {100 * a['t'] / tot:5.1f}% {a[TP] / t:8.1f} {a[DP] / t:6.1f} {a[SP] / t:6.1f} "
          f"{a[RD] / 1e6:10.1f} {a[WR] / 1e6:7.1f} {a[L2] / 1e6:8.1f}")
if len(sys.argv) > 2:
    json.dump({"per_step_bytes": int(rd + wr), "read": int(rd), "write": int(wr), "launches": n, "source": sys.argv[1]},
              open(sys.argv[2], "w"))
