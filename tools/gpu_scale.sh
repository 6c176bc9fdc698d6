#!/bin/bash
# N-GPU bench line (the driver's launch): tools/gpu_scale.sh <tag> <N>
TAG=${1:-x}; N=${2:-4}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 \
  bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_n${N}.json 2> gpurun_out/${TAG}_bench_n${N}.err
echo "bench N=$N exit $?"; cut -c1-600 gpurun_out/${TAG}_bench_n${N}.json; tail -3 gpurun_out/${TAG}_bench_n${N}.err
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-configs > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
cut -c1-400 gpurun_out/${TAG}_bench_n1.json
