#!/bin/bash
# Usage: bash tools/batch_sweep.sh BATCH "HD_X=1" "HD_Y=0" ...  -> ms per denoise step of a 50-step DDIM run at that batch
B=$1; shift
for e in "$@"; do
  env $e python bench.py --batch $B --sampler ddim --sampler-steps 50 --steps 2 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('B=$B $e', 'ms/step', round(d['ms_per_denoise_step'],4), 'faces/s', round(d['value'],1), 'launches', d['launches_per_denoise_step'], 'roofline', round(d['roofline']['frac'],4), 'finite', d['finite'])"
done
