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
