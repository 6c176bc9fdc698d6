"""GPU tests of the tcgen05 convolution kernels (forward, dgrad, wgrad) against an fp32 PyTorch reference of the
same op evaluated on the bf16-rounded operands.  Tolerances: outputs stored in bf16 -> 2^-8 relative rounding plus
fp32 accumulation-order noise; wgrad is fp32 -> 2e-3 of the tensor's max."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [(2, 16, 13, 64, 64), (2, 20, 11, 64, 128), (1, 8, 24, 128, 128), (1, 8, 24, 128, 256), (2, 6, 5, 256, 256),
          (2, 6, 5, 256, 512), (1, 7, 9, 512, 512), (3, 33, 31, 64, 64)]


def _pad(x):
    return torch.nn.functional.pad(x, (0, 0, 1, 1, 1, 1))


def _setup(shape, seed=3):
    B, H, W, Ci, Co = shape
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, H, W, Ci, generator=g).bfloat16()
    w = (torch.randn(3, 3, Ci, Co, generator=g) * (2.0 / (9 * Ci)) ** 0.5)
    b = torch.randn(Co, generator=g)
    dz = torch.randn(B, H, W, Co, generator=g).bfloat16()
    return x, w, b, dz


@pytest.fixture(scope="module")
def lib():
    from l3embedding_b200 import _lib
    return _lib.load()


def _p(t):
    return C.c_void_p(t.data_ptr())


@pytest.mark.parametrize("shape", SHAPES)
def test_tc_forward(lib, shape):
    from l3embedding_b200 import _lib
    B, H, W, Ci, Co = shape
    x, w, b, _ = _setup(shape)
    wq = w.bfloat16().float()
    ref = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), wq.permute(3, 2, 0, 1), b, padding=1).permute(0, 2, 3, 1)
    xp = _pad(x).contiguous().cuda()
    out = torch.full((B, H, W, Co), float("nan"), dtype=torch.bfloat16, device="cuda")
    scratch = torch.empty(9 * Ci * Co, dtype=torch.bfloat16, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    wd, bd = w.cuda(), b.cuda()   # keep the device copies alive across the call
    _lib.check(lib.l3_conv3x3_fwd(_p(xp), _p(wd), _p(bd), _p(out), B, H, W, Ci, Co, 1, 1, _p(scratch), st), "fwd")
    torch.cuda.synchronize()
    got = out.float().cpu()
    assert torch.isfinite(got).all()
    err = (got - ref).abs().max().item()
    assert err <= 2e-2 * max(1.0, ref.abs().max().item()), err


FIRST_SHAPES = [(2, 16, 13, 3, 64), (2, 9, 20, 1, 64), (1, 33, 31, 3, 64), (3, 40, 37, 1, 64), (5, 64, 50, 3, 64)]


@pytest.mark.parametrize("shape", FIRST_SHAPES)
def test_tc_first_layer_forward(lib, shape):
    """Cin = 1 / 3 forward: the K-major im2col operand is gathered into shared memory by the kernel's builder warps."""
    test_tc_forward(lib, shape)


@pytest.mark.parametrize("relu", [0, 1])
@pytest.mark.parametrize("shape", SHAPES + FIRST_SHAPES + [(2, 70, 66, 64, 64), (2, 40, 36, 64, 128)])
def test_tc_forward_fused_bn_statistics(lib, shape, relu):
    """The epilogue's per-channel sum / sum of squares (BN batch statistics) equal those of the stored bf16 output;
    the output itself is bit-identical to the statistics-free launch."""
    from l3embedding_b200 import _lib
    B, H, W, Ci, Co = shape
    if relu and Ci < 64:
        pytest.skip("no relu statistics on the first layer")
    x, w, b, _ = _setup(shape, seed=11)
    xp = _pad(x).contiguous().cuda()
    out = torch.full((B, H, W, Co), float("nan"), dtype=torch.bfloat16, device="cuda")
    out0 = torch.full((B, H, W, Co), float("nan"), dtype=torch.bfloat16, device="cuda")
    stats = torch.full((2 * Co,), float("nan"), dtype=torch.float64, device="cuda")
    scratch = torch.empty(9 * max(Ci, 64) * Co, dtype=torch.bfloat16, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    wd, bd = w.cuda(), b.cuda()
    _lib.check(lib.l3_conv3x3_fwd(_p(xp), _p(wd), _p(bd), _p(out0), B, H, W, Ci, Co, 1, 1, _p(scratch), st), "fwd")
    _lib.check(lib.l3_conv3x3_fwd_stats(_p(xp), _p(wd), _p(bd), _p(out), B, H, W, Ci, Co, _p(scratch), _p(stats), relu, st),
               "fwd_stats")
    torch.cuda.synchronize()
    assert torch.equal(out.view(torch.int16), out0.view(torch.int16))
    z = out.double().cpu().reshape(-1, Co)
    if relu:
        z = z.clamp_min(0)
    s1, s2 = z.sum(0), (z * z).sum(0)
    got = stats.cpu()
    # fp32 partial sums per lane, double across warps: relative error ~1e-6 of the absolute sums
    assert (got[:Co] - s1).abs().max().item() <= 1e-5 * z.abs().sum(0).max().item() + 1e-6
    assert (got[Co:] - s2).abs().max().item() <= 1e-5 * s2.max().item() + 1e-6


@pytest.mark.parametrize("shape", SHAPES)
def test_tc_dgrad(lib, shape):
    from l3embedding_b200 import _lib
    B, H, W, Ci, Co = shape
    x, w, b, dz = _setup(shape)
    wq = w.bfloat16().float()
    xr = x.float().clone().requires_grad_(True)
    y = torch.nn.functional.conv2d(xr.permute(0, 3, 1, 2), wq.permute(3, 2, 0, 1), None, padding=1).permute(0, 2, 3, 1)
    y.backward(dz.float())
    dzp = _pad(dz).contiguous().cuda()
    da = torch.full((B, H, W, Ci), float("nan"), dtype=torch.bfloat16, device="cuda")
    scratch = torch.empty(9 * Ci * Co * 2, dtype=torch.bfloat16, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    wd = w.cuda()
    _lib.check(lib.l3_conv3x3_dgrad(_p(dzp), _p(wd), _p(da), B, H, W, Ci, Co, 1, 1, _p(scratch), st), "dgrad")
    torch.cuda.synchronize()
    got = da.float().cpu()
    assert torch.isfinite(got).all()
    err = (got - xr.grad).abs().max().item()
    assert err <= 2e-2 * max(1.0, xr.grad.abs().max().item()), err


@pytest.mark.parametrize("shape", SHAPES + [(2, 70, 66, 64, 64), (2, 40, 36, 128, 128), (3, 20, 17, 256, 256)])
def test_tc_dgrad_fused_bn_backward_statistics(lib, shape):
    """dgrad whose epilogue also reduces sum(dy), sum(dy*z) of the layer below (dy = da where scale*z + shift > 0): da is
    bit-identical to the plain dgrad, the sums match the stored da."""
    from l3embedding_b200 import _lib
    B, H, W, Ci, Co = shape
    x, w, b, dz = _setup(shape, seed=17)
    g = torch.Generator().manual_seed(5)
    z = torch.randn(B, H, W, Ci, generator=g).bfloat16()
    scale = torch.randn(Ci, generator=g)            # both signs: the mask is on scale*z + shift, not on z
    shift = 0.3 * torch.randn(Ci, generator=g)
    dzp = _pad(dz).contiguous().cuda()
    da = torch.full((B, H, W, Ci), float("nan"), dtype=torch.bfloat16, device="cuda")
    da0 = torch.full((B, H, W, Ci), float("nan"), dtype=torch.bfloat16, device="cuda")
    sums = torch.full((2 * Ci,), float("nan"), dtype=torch.float64, device="cuda")
    scratch = torch.empty(9 * Ci * Co * 2, dtype=torch.bfloat16, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    wd, zd, scd, shd = w.cuda(), z.cuda(), scale.cuda(), shift.cuda()
    _lib.check(lib.l3_conv3x3_dgrad(_p(dzp), _p(wd), _p(da0), B, H, W, Ci, Co, 1, 1, _p(scratch), st), "dgrad")
    _lib.check(lib.l3_conv3x3_dgrad_stats(_p(dzp), _p(wd), _p(da), B, H, W, Ci, Co, _p(scratch), _p(zd), _p(scd), _p(shd),
                                          _p(sums), st), "dgrad_stats")
    torch.cuda.synchronize()
    assert torch.equal(da.view(torch.int16), da0.view(torch.int16))
    dav, zv = da.float().cpu().reshape(-1, Ci), z.float().reshape(-1, Ci)
    mask = (zv.double() * scale.double() + shift.double()) > 0   # the kernel's fp32 fma rounds the exact value: same sign
    dy = torch.where(mask, dav, torch.zeros_like(dav)).double()
    s1, s2 = dy.sum(0), (dy * zv.double()).sum(0)
    got = sums.cpu()
    tol1 = 1e-5 * dy.abs().sum(0).max().item() + 1e-6
    tol2 = 1e-5 * (dy * zv.double()).abs().sum(0).max().item() + 1e-6
    assert (got[:Ci] - s1).abs().max().item() <= tol1, ((got[:Ci] - s1).abs().max().item(), tol1)
    assert (got[Ci:] - s2).abs().max().item() <= tol2, ((got[Ci:] - s2).abs().max().item(), tol2)


@pytest.mark.parametrize("shape", SHAPES)
def test_tc_wgrad(lib, shape):
    from l3embedding_b200 import _lib
    B, H, W, Ci, Co = shape
    x, w, b, dz = _setup(shape)
    wr = w.clone().requires_grad_(True)
    y = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), wr.permute(3, 2, 0, 1), None, padding=1).permute(0, 2, 3, 1)
    y.backward(dz.float())
    xp, dzp = _pad(x).contiguous().cuda(), _pad(dz).contiguous().cuda()
    dw = torch.full((3, 3, Ci, Co), float("nan"), device="cuda")
    db = torch.full((Co,), float("nan"), device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.l3_conv3x3_wgrad(_p(xp), _p(dzp), _p(dw), _p(db), B, H, W, Ci, Co, 1, 1, st), "wgrad")
    torch.cuda.synchronize()
    got = dw.cpu()
    assert torch.isfinite(got).all()
    scale = wr.grad.abs().max().item()
    assert (got - wr.grad).abs().max().item() <= 2e-3 * scale, ((got - wr.grad).abs().max().item(), scale)
    ref_db = dz.float().sum(dim=(0, 1, 2))
    assert (db.cpu() - ref_db).abs().max().item() <= 2e-3 * max(1.0, ref_db.abs().max().item())


@pytest.mark.parametrize("shape", [(2, 16, 13, 3, 64), (2, 9, 20, 1, 64), (1, 33, 31, 3, 64), (3, 40, 37, 1, 64)])
def test_tc_first_layer_wgrad(lib, shape):
    """Cin = 1 / 3: the im2col operand is built in shared memory by the kernel itself."""
    test_tc_wgrad(lib, shape)


def test_tc_path_is_active_in_bf16_engine():
    from l3embedding_b200.engine import Engine
    eng = Engine("cnn_L3_melspec2", 2, "bf16", training=True)
    assert eng.uses_tensor_cores


_VARIANT_SCRIPT = """
import sys, numpy as np
sys.path.insert(0, %r)
from oracle import l3_oracle as O
from l3embedding_b200.engine import Engine
w = O.init_weights('cnn_L3_melspec2', seed=3, randomize_bn=True)
v, a, l = O.synthetic_batch(3, seed=21)
e = Engine('cnn_L3_melspec2', 3, 'bf16', training=True, weights=w)
e.forward_backward(v, a, l)
g = e.get_grads()
g['__loss__'] = np.array(e.metrics()['loss'])
np.savez(sys.argv[1], **{k.replace('/', '.'): x for k, x in g.items()})
"""


def test_kernel_variants_agree_on_a_training_step(tmp_path):
    """The same bf16 training step through (a) per-CTA kernels with per-chunk butterfly statistics, per-tap wgrad tiles
    and the SIMT first-layer wgrad (L3_CONV_TC_VARIANT=2, L3_WGRAD_TC_VARIANT=1, L3_FIRST_WGRAD_TC=0), (b) the
    defaults (CTA-pair kernels, per-configuration statistics epilogues, shared-halo wgrad, tensor-core first layer) and
    (c) the store-phase statistics everywhere plus BN-backward pass 1 fused into the dgrad epilogues
    (L3_CONV_EPI=store, L3_DGRAD_FUSE_STATS=1), (d) max-pool routing re-derived in the backward kernels instead of read
    from the forward pass's record, weight gradients on the towers' own streams (L3_POOL_RECORD=0, L3_WGRAD_STREAMS=0):
    identical math, different reduction orders."""
    import os
    import subprocess
    import sys
    import numpy as np
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = {}
    for name, env in (("a", {"L3_CONV_TC_VARIANT": "2", "L3_FIRST_WGRAD_TC": "0", "L3_WGRAD_TC_VARIANT": "1"}), ("b", {}),
                      ("c", {"L3_CONV_EPI": "store", "L3_DGRAD_FUSE_STATS": "1"}),
                      ("d", {"L3_POOL_RECORD": "0", "L3_WGRAD_STREAMS": "0"})):
        path = str(tmp_path / (name + ".npz"))
        e = dict(os.environ)
        e.update(env)
        subprocess.run([sys.executable, "-c", _VARIANT_SCRIPT % root, path], check=True, env=e, timeout=300)
        outs[name] = dict(np.load(path))
    gb = outs["b"]
    for other in ("a", "c", "d"):
        ga = outs[other]
        assert abs(float(ga["__loss__"]) - float(gb["__loss__"])) <= 2e-3
        worst = []
        for k in ga:
            if k == "__loss__":
                continue
            a, b = ga[k].ravel().astype(np.float64), gb[k].ravel().astype(np.float64)
            # conv biases before a training-mode BN have analytically zero gradients, and the input-BN gamma/beta are
            # small residuals of huge cancelling sums: all noise-dominated in bf16 storage (the fp64 and bf16-emulating
            # oracles themselves differ by 9x on audio/bn0/gamma), so only the well-conditioned tensors are compared
            if k.endswith(".bias") or ".bn0." in k or a.size < 8:
                continue
            cos = float(a @ b / max(np.linalg.norm(a) * np.linalg.norm(b), 1e-30))
            worst.append((cos, k))
        worst.sort()
        print(other, "vs b, lowest cosines:", worst[:5])
        # bf16 storage makes the reduction orders diverge by 1-ulp flips that re-route ReLU / max-pool gradient paths:
        # the same bar as the comparison with the bf16-emulating oracle
        assert worst[0][0] >= 0.9, (other, worst[:5])
