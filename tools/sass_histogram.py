"""Per-kernel SASS instruction histogram of the built library (what proves the Blackwell-native paths):

    python tools/sass_histogram.py hifidiff_b200/csrc/libhifidiff_b200.so > profiles/r2_sass_histogram.txt

Counts, per kernel, the mnemonics that matter (guide: tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA ->
UTMALDG / UTMASTG / UBLKCP / UBLKPF, mma.sync -> HMMA, ldmatrix -> LDSM, clusters -> UCGABAR, mbarrier -> SYNCS) plus
the total instruction count (code size = 16 bytes each)."""
import collections
import re
import subprocess
import sys

KEYS = ["UTCHMMA", "UTCBAR", "UTCATOM", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "UBLKPF", "HMMA", "LDSM", "UCGABAR",
        "SYNCS", "ACQBULK", "LDGSTS", "FFMA", "STL", "LDL"]
out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
kern, counts, totals = None, collections.OrderedDict(), {}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern).replace("void ", "").replace("hd::", "")
        counts[kern] = collections.Counter()
        totals[kern] = 0
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        op = m.group(1)
        totals[kern] += 1
        for k in KEYS:
            if op.startswith(k):
                counts[kern][k + (".2CTA" if ".2CTA" in op else "") + (".MULTICAST" if "MULTICAST" in op else "")] += 1
print(f"# cuobjdump -sass {sys.argv[1]}: {len(counts)} kernels, arch sm_100a")
tot = collections.Counter()
for k, c in counts.items():
    tot.update(c)
print("# whole library: " + ", ".join(f"{k} x{v}" for k, v in sorted(tot.items())))
for k, c in sorted(counts.items(), key=lambda kv: -totals[kv[0]]):
    keys = ", ".join(f"{kk} x{v}" for kk, v in sorted(c.items()) if kk not in ("FFMA",))
    print(f"{totals[k]:6d} instr  {k[:110]:110s} {keys}")
