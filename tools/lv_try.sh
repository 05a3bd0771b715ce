#!/bin/bash
# Level-chain bring-up: parity test, trace, short bench with and without the chain.
mkdir -p gpurun_out
export HD_LV=${HD_LV:-16}
timeout 300 python -m pytest tests/test_gpu_denoiser.py -m gpu -q --tb=short -s -p no:cacheprovider -k "level_chain or face_kernel" > gpurun_out/lv_test.log 2>&1; echo "test rc=$?"; tail -15 gpurun_out/lv_test.log
HD_LV_TRACE=1 timeout 120 python tools/lv_trace.py 256 > gpurun_out/lv_trace.txt 2>&1; echo "trace rc=$?"; tail -30 gpurun_out/lv_trace.txt
for e in "HD_LV=0" "HD_LV=16" "HD_LV=16 HD_LV_COOP=0"; do
  env $e timeout 300 python bench.py --steps 1 --warmup 2 --no-cpu-baseline 2>gpurun_out/lv_bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$e', 'ms/step', round(d['ms_per_denoise_step'],4), 'faces/s', round(d['value'],2), 'launches', d['launches_per_denoise_step'], 'finite', d['finite'])" || tail -5 gpurun_out/lv_bench.err
done
