#!/bin/bash
# Runs every GPU test file in its own process (a CUDA fault in one must not poison the rest)
# and keeps the logs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
rc=0
for f in ${@:-tests/test_gpu_gemm.py tests/test_gpu_sampler.py tests/test_gpu_denoiser.py tests/test_gpu_refiner.py tests/test_gpu_edge_cases.py tests/test_gpu_eager_bar.py tests/test_gpu_cr.py tests/test_gpu_latent32.py}; do
  name=$(basename $f .py)
  echo "=== $f"
  timeout 900 python -m pytest $f -m gpu -q --tb=short -s -p no:cacheprovider --timeout 600 > gpurun_out/$name.log 2>&1
  r=$?
  [ $r -ne 0 ] && rc=$r
  tail -25 gpurun_out/$name.log
done
exit $rc
