#!/bin/bash
# One GPU-box visit: the full GPU test suite (all failures, not -x), the bench line, and optionally the per-launch
# timing list.  Outputs under gpurun_out/.   usage: tools/gpu_round.sh <tag> [nobench] [ncu]
TAG=${1:-x}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader > gpurun_out/${TAG}_gpu.txt 2>&1
timeout 2400 python -m pytest tests -q -m gpu --durations=15 -s > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/${TAG}_pytest_gpu.log
grep -E "passed|failed|FAILED|Error" gpurun_out/${TAG}_pytest_gpu.log | tail -40
if [[ "$*" != *nobench* ]]; then
  timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
  cat gpurun_out/${TAG}_bench.json
  tail -3 gpurun_out/${TAG}_bench.err
fi
if [[ "$*" == *ncu* ]]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
  echo "ncu exit $?"
fi
