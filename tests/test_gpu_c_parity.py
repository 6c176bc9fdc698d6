"""GPU parity tests, fp32 parity mode: front-end, SIMT convolutions and the inference path (keras predict /
load_embedding) against the fp64 CPU oracle on the same seeded inputs, through the C ABI.

Tolerances (SURVEY 7.2 / north_star):
  front-end dB map            <= 1e-3 dB max-abs vs the fp64 oracle (the reference's own fp32 DFT-as-matmul is 5e-4 off)
  embeddings / logits (f32)   <= 1e-3 max-abs vs the fp64 oracle       (north_star: "within 1e-3 max-abs")
"""
import json
import os

import numpy as np
import pytest
import torch

from oracle import l3_oracle as O
from _gpu_common import MODEL_TYPES, GOLDEN, F64, engine as _engine, rel_l2, pad as _pad, oracle_inputs as _oracle_inputs

pytestmark = pytest.mark.gpu

@pytest.mark.parametrize("model_type", MODEL_TYPES)
def test_frontend_matches_oracle(model_type):
    _, audio, _ = O.synthetic_batch(3, seed=101)
    audio[2] = 0                                 # a silent clip: every cell at the amin floor
    audio[1, 0, :24000] = 0                      # half-silent
    eng = _engine(model_type, 3, "f32", training=False, towers=("audio",))
    got = eng.frontend(audio).cpu().numpy()
    ref = O.frontend(torch.from_numpy(O.pcm2float(audio, "float64")), model_type, F64)[..., 0].numpy()
    assert got.shape == ref.shape
    tol = 1e-3   # dB for the decibel models; log(x)/5 units for cnn_L3_orig (audio_model.py:43)
    assert np.abs(got - ref).max() <= tol, np.abs(got - ref).max()
    # float32 input path == int16 input path (pcm2float is exact in fp32)
    got_f = eng.frontend(O.pcm2float(audio, "float32")).cpu().numpy()
    assert np.array_equal(got, got_f)


def test_frontend_matches_golden_fixture():
    with np.load(GOLDEN) as z:
        meta = json.loads(str(z["meta"]))
        _, audio, _ = O.synthetic_batch(meta["batch"], seed=meta["data_seed"])
        for mt in meta["model_types"]:
            eng = _engine(mt, meta["batch"], "f32", training=False, towers=("audio",))
            got = eng.frontend(audio).cpu().numpy()[:, ::meta["stride_f"], ::meta["stride_t"]]
            assert np.abs(got - z[mt + "/frontend"]).max() <= 1e-3


@pytest.mark.parametrize("shape", [(2, 9, 7, 3, 64), (1, 16, 13, 64, 64), (2, 8, 24, 128, 256), (3, 5, 5, 1, 64)])
def test_conv_simt_fwd_dgrad_wgrad_f32(shape):
    import ctypes as C
    from l3embedding_b200 import _lib
    lib = _lib.load()
    B, H, W, Ci, Co = shape
    g = torch.Generator().manual_seed(5)
    x = torch.randn(B, H, W, Ci, generator=g)
    w = torch.randn(3, 3, Ci, Co, generator=g) * 0.1
    b = torch.randn(Co, generator=g)
    dz = torch.randn(B, H, W, Co, generator=g)
    xr = x.clone().requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    y = torch.nn.functional.conv2d(xr.permute(0, 3, 1, 2), wr.permute(3, 2, 0, 1), b, padding=1).permute(0, 2, 3, 1)
    y.backward(dz)
    dev = "cuda"
    p = lambda t: C.c_void_p(t.data_ptr())
    xp, dzp = _pad(x).contiguous().to(dev), _pad(dz).contiguous().to(dev)
    wd, bd = w.contiguous().to(dev), b.to(dev)
    out = torch.empty(B, H, W, Co, device=dev)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.l3_conv3x3_fwd(p(xp), p(wd), p(bd), p(out), B, H, W, Ci, Co, 0, 0, None, st), "fwd")
    assert torch.allclose(out.cpu(), y.detach(), atol=2e-4, rtol=1e-4)
    da = torch.empty(B, H, W, Ci, device=dev)
    scratch = torch.empty(9 * Ci * Co, device=dev)
    _lib.check(lib.l3_conv3x3_dgrad(p(dzp), p(wd), p(da), B, H, W, Ci, Co, 0, 0, p(scratch), st), "dgrad")
    assert torch.allclose(da.cpu(), xr.grad, atol=2e-4, rtol=1e-4)
    dw = torch.empty(3, 3, Ci, Co, device=dev)
    db = torch.empty(Co, device=dev)
    _lib.check(lib.l3_conv3x3_wgrad(p(xp), p(dzp), p(dw), p(db), B, H, W, Ci, Co, 0, 0, st), "wgrad")
    assert torch.allclose(dw.cpu(), wr.grad, atol=1e-3, rtol=1e-4)
    assert torch.allclose(db.cpu(), dz.sum(dim=(0, 1, 2)), atol=1e-3, rtol=1e-4)


PARITY_MODES = ["f32", "f32tc"]      # SIMT fp32 / split 16-bit operands on tcgen05 (include/l3b200.h): the same 1e-3 bar


@pytest.mark.parametrize("mode", PARITY_MODES)
@pytest.mark.parametrize("model_type", MODEL_TYPES)
def test_inference_logits_and_embeddings_f32(model_type, mode):
    """keras predict path (BN moving statistics): AVC logits/probabilities and both audio embeddings + the vision
    embedding within 1e-3 max-abs of the fp64 oracle -- in both parity modes."""
    B = 2
    w_np = O.init_weights(model_type, seed=20180123, randomize_bn=True)
    video, audio, _ = O.synthetic_batch(B, seed=202)
    vf, af = _oracle_inputs(video, audio)
    w = O.to_torch(w_np, dtype=torch.float64)
    ref_logits = O.avc_forward(vf, af, w, model_type, False, F64).numpy()
    eng = _engine(model_type, B, mode, training=False, weights=w_np)
    assert eng.uses_tensor_cores == (mode == "f32tc")
    probs, logits = eng.predict(video, audio)
    assert np.abs(logits - ref_logits).max() <= 1e-3, np.abs(logits - ref_logits).max()
    ref_p = torch.softmax(torch.from_numpy(ref_logits), dim=1).numpy()
    assert np.abs(probs - ref_p).max() <= 1e-3
    for pooling in ("original", "short"):
        ref = O.audio_embedding(af, w, model_type, pooling, F64).numpy()
        got = eng.embed_audio(audio, pooling).cpu().numpy()
        assert got.shape == ref.shape and np.abs(got - ref).max() <= 1e-3, (pooling, np.abs(got - ref).max())
    ref_v = O.vision_embedding(vf, w, model_type, F64).numpy()
    got_v = eng.embed_vision(video).cpu().numpy()
    assert got_v.shape == (B, 8192) and np.abs(got_v - ref_v).max() <= 1e-3
    # float inputs as the reference generator yields them (train.py:186,189) give the same result as raw u8/i16
    p2, _ = eng.predict(O.scale_video(video), O.pcm2float(audio, "float32"))
    assert np.abs(p2 - probs).max() <= 1e-6


@pytest.mark.parametrize("mode", PARITY_MODES)
def test_embedding_matches_golden_fixture(mode):
    with np.load(GOLDEN) as z:
        meta = json.loads(str(z["meta"]))
        _, audio, _ = O.synthetic_batch(meta["batch"], seed=meta["data_seed"])
        for mt in meta["embedding_types"]:
            w_np = O.init_weights(mt, seed=meta["weight_seed"], randomize_bn=True)
            eng = _engine(mt, meta["batch"], mode, training=False, towers=("audio",), weights=w_np)
            assert np.abs(eng.embed_audio(audio, "short").cpu().numpy() - z[mt + "/embedding_short"]).max() <= 1e-3
            assert np.abs(eng.embed_audio(audio, "original").cpu().numpy() - z[mt + "/embedding_original"]).max() <= 1e-3


@pytest.mark.parametrize("model_type", ["cnn_L3_melspec2", "cnn_L3_kapredbinputbn", "cnn_L3_orig"])
def test_fused_input_stage_equals_the_standalone_front_end(model_type):
    """Inside a tower the dB reference (max(x - clip max, -80)) is applied by the fused input pass (k_input_stage mode 2,
    together with the input BatchNorm), the stand-alone front-end op applies it with its own finish kernel: the tower's
    x0 must equal the front-end output bit for bit -- in inference (one pass) and in training (statistics pass first)."""
    B = 3
    video, audio, label = O.synthetic_batch(B, seed=77)
    audio[1, 0, 20000:] = 0
    for training in (False, True):
        eng = _engine(model_type, B, "f32", training=training)
        want = eng.frontend(audio).cpu().numpy()
        if training:
            eng.train_step_host(video, audio, label, 0.0)
        else:
            eng.predict(video, audio)
        got = eng.debug_read("audio/x0", B).reshape(want.shape)
        assert np.array_equal(got, want), (training, np.abs(got - want).max())


def test_video_scaling_is_bit_exact():
    """train.py:186: `2 * img_as_float(u8).astype('float32') - 1` -- the device's u8 -> float conversion (k_input_stage mode 1)
    against the numpy expression, bit for bit (cnn_L3_orig has no input BN, so x0 is the scaled frame itself)."""
    B = 2
    video, audio, _ = O.synthetic_batch(B, seed=31)
    video[0, 0, 0] = (0, 255, 128)                      # the range ends and the middle
    eng = _engine("cnn_L3_orig", B, "f32", training=False)
    eng.predict(video, audio)
    got = eng.debug_read("vision/x0", B).reshape(B, 224, 224, 3)
    want = O.scale_video(video)
    assert got.dtype == want.dtype == np.float32 and np.array_equal(got, want)
    assert got.min() == -1.0 and got.max() == 1.0
