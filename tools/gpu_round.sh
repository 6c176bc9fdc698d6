#!/bin/bash
# One GPU-box visit: kernel unit tests, parity tests, bench line, per-launch timing list.  Outputs under gpurun_out/.
# usage: tools/gpu_round.sh <tag> [skip_parity]
TAG=${1:-x}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader > gpurun_out/${TAG}_gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_tc.py -x -q > gpurun_out/${TAG}_test_tc.log 2>&1
echo "tc tests exit $?" | tee -a gpurun_out/${TAG}_test_tc.log
tail -5 gpurun_out/${TAG}_test_tc.log
if [ -z "$2" ]; then
  timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/${TAG}_test_parity.log 2>&1
  echo "parity tests exit $?" | tee -a gpurun_out/${TAG}_test_parity.log
  tail -5 gpurun_out/${TAG}_test_parity.log
fi
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_bench.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
echo "ncu exit $?"
