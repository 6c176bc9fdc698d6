#!/bin/bash
# diagnostic at N GPUs: where does the N-GPU step lose against N=1?  (exchange on / off, NCCL channel count)
N=${1:-8}
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29566 \
  bench.py --gpus $N --steps 20 --warmup 5 --no-configs $EXTRA > gpurun_out/ab_${tag}.json 2> gpurun_out/ab_${tag}.err
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/ab_${tag}.json') if l.startswith('{')][-1])
print('${tag}', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['clocks']['sm_mhz'])
PY
}
EXTRA="--no-exchange" run noexchange X=1
EXTRA="" run default X=1
EXTRA="" run ch4 NCCL_MAX_NCHANNELS=4
EXTRA="" run ch8 NCCL_MAX_NCHANNELS=8
EXTRA="" run nvls NCCL_ALGO=NVLS
EXTRA="--no-exchange" run noexchange2 X=1
