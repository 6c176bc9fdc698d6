"""Kernel timeline of the audio-embedding inference path (one batch of 500 clips), via torch.profiler / CUPTI."""
import json, os, re, sys, collections
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from torch.profiler import ProfilerActivity, profile
from l3embedding_b200.engine import Engine
from l3embedding_b200.synthetic import synthetic_batch
B = 500
eng = Engine("cnn_L3_melspec2", B, "bf16", training=False, towers=("audio",), host_staging=False)
a = torch.from_numpy(synthetic_batch(B, seed=7)[1]).cuda()
out = torch.empty(B, 6144, device="cuda")
for _ in range(3):
    eng.embed_audio(a, "original", out=out)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(2):
        eng.embed_audio(a, "original", out=out)
    torch.cuda.synchronize()
prof.export_chrome_trace("gpurun_out/embed_trace.json")
ev = json.load(open("gpurun_out/embed_trace.json"))["traceEvents"]
rows = sorted([[e["name"], e["ts"], e["dur"]] for e in ev if e.get("cat") in ("kernel", "gpu_memset", "gpu_memcpy") and "dur" in e], key=lambda r: r[1])
os.remove("gpurun_out/embed_trace.json")
half = rows[len(rows) // 2:]
t0 = half[0][1]
for n, ts, d in half:
    print("%8.3f %8.1f %s" % ((ts - t0) / 1e3, d, re.sub(r"\(.*", "", n).replace("void ", "").replace("l3::", "")[:60]))
print("wall ms", (half[-1][1] + half[-1][2] - t0) / 1e3, "sum ms", sum(r[2] for r in half) / 1e3)
