#!/bin/bash
# One GPU call: bench line, warm per-launch costs, GEMM phase timeline, ncu launch list with metrics, one --set full capture.
# Usage: bash tools/gpu_profile.sh <tag> [kernel-regex for the --set full capture]
tag=${1:-cur}; kre=${2:-gemm_tc_kernel}
mkdir -p gpurun_out
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"; cat gpurun_out/${tag}_bench.json
python tools/profile_ops.py 256 -3 > gpurun_out/${tag}_ops_graph_B256.txt 2>&1; echo "ops rc=$?"; head -40 gpurun_out/${tag}_ops_graph_B256.txt
python tools/gemm_trace.py > gpurun_out/${tag}_gemm_trace.txt 2>&1; echo "trace rc=$?"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed
python tools/profile_sampler_step.py 256 > gpurun_out/plain.log 2>&1 &&
timeout 600 ncu --metrics $M --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/${tag}_launches_metrics_B256.csv python tools/profile_sampler_step.py 256 > gpurun_out/ncu1.log 2>&1
echo "ncu launches rc=$?"
python tools/summarize_metrics.py gpurun_out/${tag}_launches_metrics_B256.csv gpurun_out/${tag}_traffic.json > gpurun_out/${tag}_launches_metrics_B256.summary.txt; head -30 gpurun_out/${tag}_launches_metrics_B256.summary.txt
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:$kre" -s 40 -c 4 \
    -f -o gpurun_out/${tag}_prof_full python tools/profile_sampler_step.py 256 > gpurun_out/ncu2.log 2>&1
echo "ncu full rc=$?"
