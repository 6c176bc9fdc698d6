"""Structured-input diagnostics for the tcgen05 conv kernels (developer tool, GPU only)."""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from l3embedding_b200 import _lib
lib = _lib.load()
torch.set_printoptions(linewidth=200, precision=3, sci_mode=False)
p = lambda t: C.c_void_p(t.data_ptr())
pad = lambda x: torch.nn.functional.pad(x, (0, 0, 1, 1, 1, 1))


def fwd(x, w, b=None):
    B, H, W, Ci = x.shape
    Co = w.shape[-1]
    xp = pad(x.bfloat16()).contiguous().cuda()
    out = torch.full((B, H, W, Co), float("nan"), dtype=torch.bfloat16, device="cuda")
    scratch = torch.empty(9 * Ci * Co, dtype=torch.bfloat16, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    bb = b.cuda() if b is not None else None
    _lib.check(lib.l3_conv3x3_fwd(p(xp), p(w.cuda()), p(bb) if bb is not None else None, p(out), B, H, W, Ci, Co, 1, 1, p(scratch), st), "fwd")
    torch.cuda.synchronize()
    return out.float().cpu()


def ref(x, w, b=None):
    return torch.nn.functional.conv2d(x.bfloat16().float().permute(0, 3, 1, 2), w.bfloat16().float().permute(3, 2, 0, 1), b, padding=1).permute(0, 2, 3, 1)


Ci = Co = 64
B, H, W = 1, 4, 5
# A: center-tap identity, pixel id in channel 0
x = torch.zeros(B, H, W, Ci)
x[..., 0] = torch.arange(1, H * W + 1).reshape(1, H, W).float()
x[..., 5] = 100 + torch.arange(1, H * W + 1).reshape(1, H, W).float()
w = torch.zeros(3, 3, Ci, Co)
w[1, 1] = torch.eye(Ci)
g = fwd(x, w)
print("A ch0:\n", g[0, :, :, 0], "\nA ch5:\n", g[0, :, :, 5], "\nA nonzero channels:", (g.abs().sum(dim=(0, 1, 2)) > 0).nonzero().flatten().tolist())
# B: random x, identity: which input channel does each output channel carry?
x = torch.randn(B, H, W, Ci)
g = fwd(x, w)
xr = x.bfloat16().float()
print("B max err identity:", (g - xr).abs().max().item())
corr = torch.einsum("bhwi,bhwo->io", xr, g)
print("B argmax in-channel per out-channel:", corr.argmax(dim=0).tolist())
# C: single tap at (0,0) identity -> out[y,x] = in[y-1,x-1]
w2 = torch.zeros(3, 3, Ci, Co); w2[0, 0] = torch.eye(Ci)
x = torch.zeros(B, H, W, Ci); x[..., 0] = torch.arange(1, H * W + 1).reshape(1, H, W).float()
print("C tap(0,0):\n", fwd(x, w2)[0, :, :, 0])
w2 = torch.zeros(3, 3, Ci, Co); w2[2, 1] = torch.eye(Ci)
print("C tap(2,1):\n", fwd(x, w2)[0, :, :, 0])
# D: channel mixing: w center = random matrix
wm = torch.zeros(3, 3, Ci, Co); wm[1, 1] = torch.randn(Ci, Co) * 0.1
x = torch.randn(B, H, W, Ci)
print("D center random matrix err:", (fwd(x, wm) - ref(x, wm)).abs().max().item())
# E: full random
wf = torch.randn(3, 3, Ci, Co) * 0.05
print("E full random err:", (fwd(x, wf) - ref(x, wf)).abs().max().item())
bias = torch.randn(Co)
print("E with bias err:", (fwd(x, wf, bias) - ref(x, wf, bias)).abs().max().item())
for (B, H, W, Ci, Co) in [(2, 16, 13, 64, 64), (1, 8, 24, 128, 128), (1, 8, 24, 128, 256), (2, 6, 5, 256, 512)]:
    x = torch.randn(B, H, W, Ci); wf = torch.randn(3, 3, Ci, Co) * 0.05
    print((B, H, W, Ci, Co), "err:", (fwd(x, wf) - ref(x, wf)).abs().max().item())
