"""Run ONE convolution op of the C ABI at a real layer size (for `ncu -k regex:...` captures and quick timing).

    python tools/run_op.py --op fwd_stats --shape 64,224,224,3,64 --iters 5
    ops: fwd | fwd_stats | dgrad | dgrad_stats | wgrad       shape: B,H,W,Cin,Cout
"""
import argparse
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from l3embedding_b200 import _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--op", default="fwd_stats")
    ap.add_argument("--shape", default="64,224,224,3,64")
    ap.add_argument("--iters", type=int, default=5)
    a = ap.parse_args()
    B, H, W, Ci, Co = (int(v) for v in a.shape.split(","))
    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(1)
    xp = torch.zeros(B, H + 2, W + 2, Ci, dtype=torch.bfloat16, device="cuda")
    xp[:, 1:-1, 1:-1] = torch.randn(B, H, W, Ci, generator=g, device="cuda").bfloat16()
    dzp = torch.zeros(B, H + 2, W + 2, Co, dtype=torch.bfloat16, device="cuda")
    dzp[:, 1:-1, 1:-1] = torch.randn(B, H, W, Co, generator=g, device="cuda").bfloat16()
    w = torch.randn(3, 3, Ci, Co, generator=g, device="cuda") * (2.0 / (9 * Ci)) ** 0.5
    b = torch.randn(Co, generator=g, device="cuda")
    out = torch.empty(B, H, W, Co, dtype=torch.bfloat16, device="cuda")
    da = torch.empty(B, H, W, Ci, dtype=torch.bfloat16, device="cuda")
    dw = torch.empty(3, 3, Ci, Co, device="cuda")
    db = torch.empty(Co, device="cuda")
    stats = torch.empty(2 * Co, dtype=torch.float64, device="cuda")
    scratch = torch.empty(2 * 9 * max(Ci, 64) * Co, dtype=torch.float32, device="cuda")
    p = lambda t: C.c_void_p(t.data_ptr())
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    zb = torch.randn(B, H, W, Ci, generator=g, device="cuda").bfloat16()      # z of the layer below (dgrad_stats)
    sc = torch.rand(Ci, generator=g, device="cuda") + 0.5
    sh = torch.randn(Ci, generator=g, device="cuda") * 0.1
    sums = torch.empty(2 * Ci, dtype=torch.float64, device="cuda")

    def run():
        if a.op == "fwd":
            rc = lib.l3_conv3x3_fwd(p(xp), p(w), p(b), p(out), B, H, W, Ci, Co, 1, 1, p(scratch), st)
        elif a.op == "fwd_stats":
            rc = lib.l3_conv3x3_fwd_stats(p(xp), p(w), p(b), p(out), B, H, W, Ci, Co, p(scratch), p(stats), 0, st)
        elif a.op == "dgrad":
            rc = lib.l3_conv3x3_dgrad(p(dzp), p(w), p(da), B, H, W, Ci, Co, 1, 1, p(scratch), st)
        elif a.op == "dgrad_stats":   # dgrad with pass 1 of the BN/ReLU backward of the layer below in its epilogue
            rc = lib.l3_conv3x3_dgrad_stats(p(dzp), p(w), p(da), B, H, W, Ci, Co, p(scratch), p(zb), p(sc), p(sh), p(sums), st)
        else:
            rc = lib.l3_conv3x3_wgrad(p(xp), p(dzp), p(dw), p(db), B, H, W, Ci, Co, 1, 1, st)
        _lib.check(rc, a.op)

    ts = []
    for _ in range(a.iters):
        flush.zero_()   # evict L2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    flop = 2.0 * 9 * Ci * Co * B * H * W
    best = min(ts[1:]) if len(ts) > 1 else ts[0]
    print("%s %s: %s us  (best %.1f us incl. pack/memset launches, %.1f TFLOP/s)" %
          (a.op, a.shape, ["%.1f" % t for t in ts], best, flop / best / 1e6))


if __name__ == "__main__":
    main()
