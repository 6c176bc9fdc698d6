#!/bin/bash
# N-GPU diagnostic: which all-reduce algorithm NCCL picks for the gradient buckets, and the step with NVLS forced.
#   tools/gpu_nccl_algo.sh <N>
N=${1:-8}
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577 \
  bench.py --gpus $N --steps 20 --warmup 5 --no-configs > gpurun_out/al_${tag}.json 2> gpurun_out/al_${tag}.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/al_${tag}.json') if l.startswith('{')][-1])
    print('${tag}', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['clocks']['sm_mhz'])
except Exception as e:
    print('${tag} failed', e)
PY
}
run info NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=TUNING
grep -E "AllReduce: " gpurun_out/al_info.err | sed -E 's/.*(AllReduce: [0-9]+ Bytes -> Algo [A-Z_a-z]+ proto [A-Za-z0-9]+).*/\1/' | sort | uniq -c | sort -rn | head -20 > gpurun_out/al_info_summary.txt; cat gpurun_out/al_info_summary.txt
rm gpurun_out/al_info.err
run noring NCCL_ALGO="allreduce:nvls,nvlstree,tree"
grep -i -E "Last error|Error :" gpurun_out/al_noring.err | head -3
