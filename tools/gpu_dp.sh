#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): the data-parallel tests and the N-GPU bench line.  usage: tools/gpu_dp.sh <tag> <N>
TAG=${1:-x}; N=${2:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/${TAG}_gpus.txt 2>&1
timeout 1200 python -m pytest tests/test_gpu_g_dp.py tests/test_gpu_d_bf16.py -q -m gpu -s -k "dp or frozen or library or fit_generator" > gpurun_out/${TAG}_pytest_dp.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/${TAG}_pytest_dp.log
grep -E "passed|failed|FAILED|Error|frozen-routing" gpurun_out/${TAG}_pytest_dp.log | cut -c1-600 | tail -20
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
  bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_n${N}.json 2> gpurun_out/${TAG}_bench_n${N}.err
echo "bench N=$N exit $?"; cat gpurun_out/${TAG}_bench_n${N}.json | cut -c1-1500; tail -5 gpurun_out/${TAG}_bench_n${N}.err
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-configs > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
cat gpurun_out/${TAG}_bench_n1.json | cut -c1-700
