"""BASELINE.json configs[2]: audio-tower embedding inference (05_generate_embedding_samples path), 10k x 1 s clips on
one B200.  Prints clips/s for the bf16 tcgen05 path and the fp32 parity path, device-resident and from host int16."""
import json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
from l3embedding_b200.engine import Engine
from l3embedding_b200.synthetic import synthetic_batch

N, BATCH = 10000, 500
_, audio, _ = synthetic_batch(BATCH, seed=7)
host = torch.from_numpy(audio).pin_memory()
out = {}
for dtype in ("bf16", "f32"):
    eng = Engine("cnn_L3_melspec2", BATCH, dtype, training=False, towers=("audio",), host_staging=False)
    dev = host.cuda()
    res = torch.empty(BATCH, 6144, device="cuda")
    n_iter = N // BATCH if dtype == "bf16" else 4
    for _ in range(2):
        eng.embed_audio(dev, "original", out=res)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n_iter):
        eng.embed_audio(dev, "original", out=res)
    e1.record(); torch.cuda.synchronize()
    resident = n_iter * BATCH / (e0.elapsed_time(e1) / 1e3)
    hres = torch.empty(BATCH, 6144).pin_memory()
    t0 = time.perf_counter()
    for _ in range(n_iter):
        d = host.cuda(non_blocking=True)
        eng.embed_audio(d, "original", out=res)
        hres.copy_(res, non_blocking=True)
    torch.cuda.synchronize()
    e2e = n_iter * BATCH / (time.perf_counter() - t0)
    out[dtype] = {"clips_per_s_resident": resident, "clips_per_s_host_to_host": e2e, "clips": n_iter * BATCH,
                  "tflops_conv": resident * 20.405 / 1e3}
    eng.close()
print(json.dumps({"workload": "cnn_L3_melspec2 audio embedding (6144-d), batch %d" % BATCH, **out}))
