#!/bin/bash
# The round-end sequence the driver runs, plus the profile captures committed under profiles/.
TAG=${1:-final}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest gpu exit $?"; tail -3 gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1; echo "launch list exit $?"
timeout 900 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,launch__grid_size \
  --clock-control none -k regex:"k_conv3x3_tc3|k_wgrad3x3_tc2|k_first_conv_tc|k_first_wgrad_tc|k_bias_grad" --launch-skip 47 --launch-count 47 \
  -f -o gpurun_out/${TAG}_conv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_conv_ncu.log 2>&1; echo "conv capture exit $?"
timeout 600 python tools/bench_embed.py > gpurun_out/${TAG}_embed.log 2>&1; tail -1 gpurun_out/${TAG}_embed.log
