#!/bin/bash
# Usage: bash tools/bench_env.sh "HD_X=1 HD_Y=2" "HD_Z=3" ...   -> one short bench line per environment
for e in "$@"; do
  env $e python bench.py --steps 1 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$e', 'ms/step', round(d['ms_per_denoise_step'],4), 'faces/s', round(d['value'],2), 'launches', d['launches_per_denoise_step'], 'finite', d['finite'])"
done
