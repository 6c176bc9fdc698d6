#!/bin/bash
# The round-end sequence the driver runs (GPU tests, smoke, both bench arms) plus the captures profiles/ is built from.
#   tools/gpu_final.sh <tag> [reps]      reps = how many times the GPU suite runs back to back (default 1)
TAG=${1:-final}; REPS=${2:-1}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader > gpurun_out/${TAG}_gpu.txt 2>&1
for r in $(seq 1 $REPS); do
  timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_pytest_gpu_$r.log 2>&1; echo "pytest gpu #$r exit $?"; tail -2 gpurun_out/${TAG}_pytest_gpu_$r.log
done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"; cut -c1-300 gpurun_out/${TAG}_bench.json
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err; echo "reference exit $?"; cut -c1-400 gpurun_out/${TAG}_bench_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/${TAG}_ncu_bench.log 2>&1; echo "launch list exit $?"
timeout 900 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,launch__grid_size \
  --clock-control none -k regex:"k_conv3x3_tc3|k_wgrad3x3_tc2|k_first_conv_tc|k_first_wgrad_tc|k_bias_grad" --launch-skip 47 --launch-count 47 \
  -f -o gpurun_out/${TAG}_conv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-configs > gpurun_out/${TAG}_conv_ncu.log 2>&1; echo "conv capture exit $?"
