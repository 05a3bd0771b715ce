"""Key counters of every launch in an .ncu-rep (--set full) as text: python tools/ncu_summary.py file.ncu-rep > summary.txt"""
import csv
import re
import subprocess
import sys

PAT = (r"^(Kernel Name|Grid Size|Block Size|dram__bytes_(read|write)\.sum|gpu__time_duration\.sum|gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|"
       r"sm__pipe_tensor_cycles_active\.avg\.pct_of_peak_sustained_(active|elapsed)|sm__warps_active\.avg\.pct_of_peak_sustained_active|"
       r"launch__registers_per_thread|launch__grid_size|launch__block_size|launch__cluster_size|launch__shared_mem_per_block_dynamic|sm__throughput\.avg\.pct_of_peak_sustained_elapsed|"
       r"smsp__inst_executed\.sum|lts__t_bytes\.sum|lts__t_sector_hit_rate\.pct|sm__cycles_active\.avg|smsp__issue_active\.avg\.pct_of_peak_sustained_active|"
       r"lts__throughput\.avg\.pct_of_peak_sustained_elapsed|l1tex__throughput\.avg\.pct_of_peak_sustained_elapsed|"
       r"smsp__average_warps_issue_stalled_(long_scoreboard|barrier|wait|short_scoreboard|sleeping|membar|no_instruction)_per_issue_active\.ratio)$")
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, u = rows[0], rows[1]
for r in rows[2:]:
    print("-" * 100)
    for k, uu, v in zip(h, u, r):
        if re.search(PAT, k):
            print(f"{k} [{uu}] = {v}")
