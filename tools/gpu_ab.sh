#!/bin/bash
# A/B bench lines under environment switches: tools/gpu_ab.sh <tag> "ENV1=a ENV2=b" "ENV1=c" ...
TAG=$1; shift
mkdir -p gpurun_out
i=0
for envs in "$@"; do
  i=$((i+1))
  echo "== $envs" | tee -a gpurun_out/${TAG}_ab.log
  env $envs timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('pairs/s %.1f  ms %.3f  e2e %.1f  serial_ms %.3f  clk %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['serial_ms_per_step'], d['clocks']['sm_mhz']), {k: round(v['ms_per_step'], 3) for k, v in d['roofline']['kernels'].items()})
    else: print(l.rstrip())
" | tee -a gpurun_out/${TAG}_ab.log
done
