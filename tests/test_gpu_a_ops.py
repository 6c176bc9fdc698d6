"""GPU unit tests of the element-wise layer kernels of the training step, one op at a time through the C ABI
(l3_act_fwd, l3_bn_act_bwd, l3_gmaxpool_fwd, l3_gmaxpool_bwd), in bf16 (throughput mode) and fp32 (parity mode)
storage, against float64 PyTorch autograd of the same keras block evaluated on IDENTICAL inputs (the bf16 tensors are
generated as bf16 and widened, so both sides see the same values).

Reference block (audio_model.py:376-437, vision_model.py:130-190):
    normal     : a = MaxPool2x2?( relu( BN_train(z) ) )
    relu_first : a = MaxPool2x2?( BN_train( relu(z) ) )          (vision_model.py:135-139)
BN_train = keras BatchNormalization in training mode: batch mean, biased variance, eps 1e-3; its backward includes the
gradient through the statistics.

Bars: stored bf16 tensors within 2 bf16 ulp of the float64 value (+1e-5 of the tensor's largest magnitude for sums of
cancelling fp32 terms); fp32 tensors 1e-5 relative to the largest magnitude; per-channel sums (d_gamma, d_beta) 1e-5
relative to the sum of absolute terms.  Collected first (file name): these are the kernels every later test builds on.
"""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

EPS = 1e-3
# (B, H, W, C): odd sizes exercise 'valid' pooling (dropped row / column), C covers 1..32 channel groups per pixel
SHAPES = [(2, 12, 10, 64), (3, 9, 7, 64), (2, 8, 6, 128), (1, 7, 9, 256), (2, 6, 4, 512), (4, 33, 27, 64)]
DTYPES = ["bf16", "f32"]


@pytest.fixture(scope="module")
def lib():
    from l3embedding_b200 import _lib
    return _lib.load()


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _tdt(dtype):
    return torch.bfloat16 if dtype == "bf16" else torch.float32


def _did(dtype):
    return 1 if dtype == "bf16" else 0


def _ulp_bf16(x):
    """spacing of bfloat16 (8 significant bits) at |x|"""
    ax = np.maximum(np.abs(x), 2.0 ** -126)
    return 2.0 ** (np.floor(np.log2(ax)) - 7)


def _close_stored(got, ref, dtype, what):
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    scale = max(float(np.abs(ref).max()), 1e-30)
    if dtype == "bf16":
        tol = 2 * _ulp_bf16(ref) + 1e-5 * scale
    else:
        tol = 1e-5 * scale
    bad = np.abs(got - ref) > tol
    assert not bad.any(), (what, int(bad.sum()), float(np.abs(got - ref).max()), scale)


def _unpad(t):
    return t[:, 1:-1, 1:-1, :]


def _make(shape, dtype, seed, relu_first):
    """z, gamma, beta and the batch statistics (float64) of the BN input.  z is drawn on the bfloat16 grid in BOTH
    storage modes: distinct values are then >= 2^-8 apart and equal values tie exactly, so the ReLU / max-pool routing
    decisions are the same on the device (fp32 fma) and in the float64 reference -- the test measures arithmetic, not
    coin flips at a decision boundary."""
    B, H, W, Cc = shape
    g = torch.Generator().manual_seed(seed)
    z = (torch.randn(B, H, W, Cc, generator=g) * 1.3 + 0.3).bfloat16().to(_tdt(dtype))
    gamma = (torch.rand(Cc, generator=g) + 0.5).double()
    beta = (torch.randn(Cc, generator=g) * 0.2).double()
    x = z.double()
    if relu_first:
        x = x.clamp_min(0)
    mean = x.mean(dim=(0, 1, 2))
    var = x.var(dim=(0, 1, 2), unbiased=False)
    invstd = 1.0 / torch.sqrt(var + EPS)
    scale = gamma * invstd
    shift = beta - mean * scale
    bn4 = torch.cat([scale, shift, mean, invstd]).float()
    return z, gamma, beta, bn4


def _windows(t, OH, OW):
    """(B,H,W,C) -> (B,OH,OW,4,C) 2x2 windows in the order (0,0),(0,1),(1,0),(1,1); 'valid': odd tails dropped"""
    B, _, _, Cc = t.shape
    return t[:, :2 * OH, :2 * OW, :].reshape(B, OH, 2, OW, 2, Cc).permute(0, 1, 3, 2, 4, 5).reshape(B, OH, OW, 4, Cc)


def _ref_block(z64, gamma, beta, pool, relu_first, bn4, batch_stats=True):
    """float64 forward of the block (NHWC in, NHWC out).  VALUES come from BN_train(z) (batch statistics in the
    autograd graph) or, with batch_stats=False, from the float coefficients the device is given; DECISIONS (ReLU mask,
    first-maximum pool routing) are always taken on scale*x+shift with those float coefficients -- a single-rounding
    fma on the device has the exact sign and is monotone in x, so both sides decide identically."""
    Cc = z64.shape[-1]
    sc, sh = bn4[:Cc].double(), bn4[Cc:2 * Cc].double()
    x = z64.clamp_min(0) if relu_first else z64
    if batch_stats:
        mean = x.mean(dim=(0, 1, 2))
        var = x.var(dim=(0, 1, 2), unbiased=False)
        y = (x - mean) / torch.sqrt(var + EPS) * gamma + beta
    else:
        y = x * sc + sh
    dec = x.detach() * sc + sh
    if not relu_first:
        y = y * (dec > 0)
        dec = dec.clamp_min(0)
    if pool:
        OH, OW = z64.shape[1] // 2, z64.shape[2] // 2
        dw = _windows(dec, OH, OW)
        idx = (dw == dw.amax(dim=3, keepdim=True)).to(torch.uint8).argmax(dim=3, keepdim=True)   # first maximum
        y = torch.gather(_windows(y, OH, OW), 3, idx).squeeze(3)
    return y


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("relu_first", [0, 1])
@pytest.mark.parametrize("pool", [0, 1])
@pytest.mark.parametrize("shape", SHAPES)
def test_act_fwd(lib, shape, pool, relu_first, dtype):
    """k_act_fwd<T, POOL, REC>: activation (+2x2 max-pool) into the zero-haloed padded layout, and -- pooled -- the
    recorded routing (winning pre-activation, window position, sign)."""
    from l3embedding_b200 import _lib
    B, H, W, Cc = shape
    z, gamma, beta, bn4 = _make(shape, dtype, 7, relu_first)
    OH, OW = (H // 2, W // 2) if pool else (H, W)
    zd = z.cuda()
    bd = bn4.cuda()
    a = torch.zeros(B, OH + 2, OW + 2, Cc, dtype=_tdt(dtype), device="cuda")
    zsel = torch.zeros(B, OH, OW, Cc, dtype=_tdt(dtype), device="cuda") if pool else None
    sel = torch.zeros(B, OH, OW, Cc, dtype=torch.uint8, device="cuda") if pool else None
    sc, sh = bd[:Cc].contiguous(), bd[Cc:2 * Cc].contiguous()
    _lib.check(lib.l3_act_fwd(_p(zd), _p(a), B, H, W, Cc, _p(sc), _p(sh), pool, relu_first, _did(dtype), _p(zsel), _p(sel),
                              _stream()), "l3_act_fwd")
    torch.cuda.synchronize()
    ref = _ref_block(z.double(), gamma, beta, pool, relu_first, bn4, batch_stats=False).numpy()
    got = _unpad(a).float().cpu().numpy()
    _close_stored(got, ref, dtype, "a")
    halo = a.clone()
    halo[:, 1:-1, 1:-1, :] = 0
    assert float(halo.float().abs().max()) == 0.0            # the halo is never written
    if pool:
        # the record: y(zsel) is the window maximum, the recorded position holds that pre-activation, the sign bit is
        # (max > 0).  Compared through values, not positions: two window entries within one fp32 ulp may swap.
        zs = zsel.double().cpu()
        ys = _ref_block(zs, gamma, beta, 0, relu_first, bn4, batch_stats=False).numpy()
        _close_stored(ys, ref, dtype, "y(zsel)")
        s = sel.cpu().numpy()
        pos = s & 3
        zc = _windows(z.double(), OH, OW)
        at_pos = torch.gather(zc, 3, torch.from_numpy(pos.astype(np.int64)).unsqueeze(3)).squeeze(3).numpy()
        assert np.array_equal(at_pos, zs.numpy())
        sign = (s >> 2) & 1
        clear = np.abs(ref) > 1e-6 * max(float(np.abs(ref).max()), 1e-30)
        assert np.array_equal(sign[clear], (ref > 0)[clear].astype(sign.dtype))
        assert (s >> 3).max() == 0


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("use_record", [0, 1])
@pytest.mark.parametrize("relu_first", [0, 1])
@pytest.mark.parametrize("pool", [0, 1])
@pytest.mark.parametrize("shape", SHAPES)
def test_bn_act_bwd(lib, shape, pool, relu_first, use_record, dtype):
    """k_bwd_stats / k_bwd_stats_sel -> k_bn_bwd_finalize -> k_bwd_apply<T, POOL, SEL>: dz, d_gamma, d_beta of the
    block against float64 autograd (gradient through the batch statistics included)."""
    from l3embedding_b200 import _lib
    if use_record and not pool:
        pytest.skip("the routing record exists for pooled layers only")
    B, H, W, Cc = shape
    z, gamma, beta, bn4 = _make(shape, dtype, 11, relu_first)
    OH, OW = (H // 2, W // 2) if pool else (H, W)
    g = torch.Generator().manual_seed(13)
    da = torch.randn(B, OH, OW, Cc, generator=g).to(_tdt(dtype))
    # float64 autograd reference
    zr = z.double().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    y = _ref_block(zr, gr, br, pool, relu_first, bn4)
    y.backward(da.double())
    zd, dad, bd = z.cuda(), da.cuda(), bn4.cuda()
    zsel = sel = None
    if use_record:
        a = torch.zeros(B, OH + 2, OW + 2, Cc, dtype=_tdt(dtype), device="cuda")
        zsel = torch.zeros(B, OH, OW, Cc, dtype=_tdt(dtype), device="cuda")
        sel = torch.zeros(B, OH, OW, Cc, dtype=torch.uint8, device="cuda")
        sc, sh = bd[:Cc].contiguous(), bd[Cc:2 * Cc].contiguous()
        _lib.check(lib.l3_act_fwd(_p(zd), _p(a), B, H, W, Cc, _p(sc), _p(sh), pool, relu_first, _did(dtype), _p(zsel),
                                  _p(sel), _stream()), "l3_act_fwd")
    dz = torch.full((B, H + 2, W + 2, Cc), float("nan"), dtype=_tdt(dtype), device="cuda")
    dg = torch.full((Cc,), float("nan"), device="cuda")
    db = torch.full((Cc,), float("nan"), device="cuda")
    _lib.check(lib.l3_bn_act_bwd(_p(dad), _p(zd), _p(dz), B, H, W, Cc, _p(bd), pool, relu_first, _did(dtype), _p(zsel),
                                 _p(sel), _p(dg), _p(db), _stream()), "l3_bn_act_bwd")
    torch.cuda.synchronize()
    got = dz.float().cpu()
    assert torch.isfinite(got).all()
    halo = got.clone()
    halo[:, 1:-1, 1:-1, :] = 0
    assert float(halo.abs().max()) == 0.0                      # the halo is (re)zeroed by the op
    _close_stored(_unpad(got).numpy(), zr.grad.numpy(), dtype, "dz")
    # per-channel sums: relative to the sum of absolute terms.  dy = gradient at the BN output, routed as above.
    Cn = Cc
    with torch.no_grad():
        x = zr.clamp_min(0) if relu_first else zr
        xhat = (x - bn4[2 * Cn:3 * Cn].double()) * bn4[3 * Cn:].double()
    probe = torch.zeros(B, H, W, Cc, dtype=torch.float64, requires_grad=True)
    sc64, sh64 = bn4[:Cn].double(), bn4[Cn:2 * Cn].double()
    dec = x * sc64 + sh64
    yo = probe
    if not relu_first:
        yo = yo * (dec > 0)
        dec = dec.clamp_min(0)
    if pool:
        dwin = _windows(dec, OH, OW)
        idx = (dwin == dwin.amax(dim=3, keepdim=True)).to(torch.uint8).argmax(dim=3, keepdim=True)
        yo = torch.gather(_windows(yo, OH, OW), 3, idx).squeeze(3)
    yo.backward(da.double())
    dy = probe.grad
    abs_beta = dy.abs().sum(dim=(0, 1, 2)).numpy()
    abs_gamma = (dy * xhat).abs().sum(dim=(0, 1, 2)).numpy() + \
        (bn4[2 * Cn:3 * Cn].double().abs() * bn4[3 * Cn:].double()).numpy() * abs_beta
    assert np.all(np.abs(db.cpu().numpy() - br.grad.numpy()) <= 1e-5 * abs_beta + 1e-7)
    assert np.all(np.abs(dg.cpu().numpy() - gr.grad.numpy()) <= 1e-5 * abs_gamma + 1e-7)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", [(2, 8, 6, 512), (3, 7, 5, 512), (64, 28, 28, 512), (1, 32, 24, 512), (5, 4, 4, 64)])
def test_gmaxpool_fwd_bwd(lib, shape, dtype):
    """k_gmaxpool_fwd (+finish) and k_gmaxpool_bwd + k_bn_bwd_apply: the global max-pool over relu(BN_train(z)) of the
    last layer (audio_model.py:436, vision_model.py:189) and its backward through the batch statistics."""
    from l3embedding_b200 import _lib
    B, H, W, Cc = shape
    z, gamma, beta, bn4 = _make(shape, dtype, 17, 0)
    zd, bd = z.cuda(), bn4.cuda()
    sc, sh = bd[:Cc].contiguous(), bd[Cc:2 * Cc].contiguous()
    out = torch.full((B, Cc), float("nan"), device="cuda")
    arg = torch.full((B, Cc), -1, dtype=torch.int32, device="cuda")
    _lib.check(lib.l3_gmaxpool_fwd(_p(zd), B, H, W, Cc, _p(sc), _p(sh), _did(dtype), _p(out), _p(arg), _stream()),
               "l3_gmaxpool_fwd")
    torch.cuda.synchronize()
    zr = z.double().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    y = _ref_block(zr, gr, br, 0, 0, bn4)                          # (B,H,W,C)
    yy = y.reshape(B, H * W, Cc)
    # first maximum in row-major order (torch.argmax documents first-occurrence on ties; bf16 inputs DO tie)
    idx = (yy.detach() == yy.detach().amax(dim=1, keepdim=True)).to(torch.uint8).argmax(dim=1)
    pooled = torch.gather(yy, 1, idx.unsqueeze(1)).squeeze(1)
    ref = pooled.detach().numpy()
    scale = max(float(np.abs(ref).max()), 1e-30)
    assert np.abs(out.cpu().numpy() - ref).max() <= 1e-5 * scale    # float output in both modes
    # inputs on the bf16 grid: distinct values are >= 2^-8 apart, equal ones tie on both sides -> identical routing
    assert torch.equal(arg.cpu().long(), idx)
    g = torch.Generator().manual_seed(19)
    dpool = torch.randn(B, Cc, generator=g).bfloat16().float()     # bf16-representable: the scatter stores it exactly
    pooled.backward(dpool.double())
    dz = torch.full((B, H + 2, W + 2, Cc), float("nan"), dtype=_tdt(dtype), device="cuda")
    dg = torch.full((Cc,), float("nan"), device="cuda")
    db = torch.full((Cc,), float("nan"), device="cuda")
    dpd = dpool.cuda()
    _lib.check(lib.l3_gmaxpool_bwd(_p(dpd), _p(arg), _p(zd), _p(dz), B, H, W, Cc, _p(bd), _did(dtype), _p(dg), _p(db),
                                   _stream()), "l3_gmaxpool_bwd")
    torch.cuda.synchronize()
    got = dz.float().cpu()
    halo = got.clone()
    halo[:, 1:-1, 1:-1, :] = 0
    assert float(halo.abs().max()) == 0.0
    _close_stored(_unpad(got).numpy(), zr.grad.numpy(), dtype, "dz")
    assert np.all(np.abs(db.cpu().numpy() - br.grad.numpy()) <= 1e-5 * dpool.abs().sum(0).numpy() + 1e-6)
    xhat = ((zr.detach() - bn4[2 * Cc:3 * Cc].double()) * bn4[3 * Cc:].double()).reshape(B, H * W, Cc)
    xh_at = torch.gather(xhat, 1, idx.unsqueeze(1)).squeeze(1)
    assert np.all(np.abs(dg.cpu().numpy() - gr.grad.numpy()) <= 1e-5 * (dpool.double().abs() * xh_at.abs()).sum(0).numpy() + 1e-6)
