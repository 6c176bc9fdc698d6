"""Kernel timeline of a few training steps (all streams) through torch.profiler / CUPTI: which kernels overlap, where the
GPU idles.  Writes gpurun_out/<tag>_timeline.json (list of [name, stream, start_us, dur_us]) for offline analysis.

    python tools/timeline.py <tag> [batch] [dtype]
"""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from torch.profiler import ProfilerActivity, profile

from l3embedding_b200.engine import Engine
from l3embedding_b200.synthetic import synthetic_batch

tag = sys.argv[1] if len(sys.argv) > 1 else "x"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
dtype = sys.argv[3] if len(sys.argv) > 3 else "bf16"
world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
eng = Engine("cnn_L3_melspec2", B, dtype, training=True)
if world > 1:      # under torchrun: the library's own data-parallel exchange (NCCL kernels show up in the timeline)
    import torch.distributed as dist
    from l3embedding_b200 import dp
    dist.init_process_group("gloo")
    dp.LibraryReplicas().attach(eng)
v, a, l = (torch.from_numpy(x).cuda() for x in synthetic_batch(B, seed=1 + rank))
for _ in range(5):
    eng.forward_backward(v, a, l, global_batch=B * world)
    eng.adam_step(1e-5)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        eng.forward_backward(v, a, l, global_batch=B * world)
        eng.adam_step(1e-5)
    torch.cuda.synchronize()
if rank != 0:
    sys.exit(0)
path = "gpurun_out/%s_trace.json" % tag
prof.export_chrome_trace(path)
ev = json.load(open(path))["traceEvents"]
rows = [[e["name"], e["args"].get("stream"), e["ts"], e["dur"]] for e in ev
        if e.get("cat") in ("kernel", "gpu_memset", "gpu_memcpy") and "dur" in e]
rows.sort(key=lambda r: r[2])
json.dump(rows, open("gpurun_out/%s_timeline.json" % tag, "w"))
os.remove(path)
print("kernels recorded:", len(rows), "span ms:", (rows[-1][2] + rows[-1][3] - rows[0][2]) / 1e3)
