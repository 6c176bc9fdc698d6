"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Tolerances (stated per SURVEY 7.2 / north_star):
  front-end dB map            <= 1e-3 dB max-abs vs the fp64 oracle (the reference's own fp32 DFT-as-matmul is 5e-4 off)
  embeddings / logits (f32)   <= 1e-3 max-abs vs the fp64 oracle       (north_star: "within 1e-3 max-abs")
  gradients (f32)             <= 1e-2 relative L2 per tensor vs fp64 autograd (measured fp32 noise floor 5.8e-3)
  bf16 throughput mode        reported, bounded loosely (bf16 cannot meet 1e-3)
"""
import json
import os

import numpy as np
import pytest
import torch

from oracle import l3_oracle as O

pytestmark = pytest.mark.gpu

MODEL_TYPES = ["cnn_L3_orig", "cnn_L3_kapredbinputbn", "cnn_L3_melspec1", "cnn_L3_melspec2"]
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "oracle_golden.npz")
F64 = O.OracleConfig(dtype=torch.float64)


def _engine(*a, **k):
    from l3embedding_b200.engine import Engine
    return Engine(*a, **k)


def rel_l2(a, b):
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


# ---------------------------------------------------------------------------------------------- front-end
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


# ---------------------------------------------------------------------------------------------- conv ops
def _pad(x):  # (B,H,W,C) -> zero-haloed (B,H+2,W+2,C)
    return torch.nn.functional.pad(x, (0, 0, 1, 1, 1, 1))


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


# ---------------------------------------------------------------------------------------------- whole model
def _oracle_inputs(video, audio):
    return (torch.from_numpy(O.scale_video(video)).double(), torch.from_numpy(O.pcm2float(audio, "float64")))


@pytest.mark.parametrize("model_type", MODEL_TYPES)
def test_inference_logits_and_embeddings_f32(model_type):
    """keras predict path (BN moving statistics): AVC logits/probabilities and both audio embeddings + the vision
    embedding within 1e-3 max-abs of the fp64 oracle."""
    B = 2
    w_np = O.init_weights(model_type, seed=20180123, randomize_bn=True)
    video, audio, _ = O.synthetic_batch(B, seed=202)
    vf, af = _oracle_inputs(video, audio)
    w = O.to_torch(w_np, dtype=torch.float64)
    ref_logits = O.avc_forward(vf, af, w, model_type, False, F64).numpy()
    eng = _engine(model_type, B, "f32", training=False, weights=w_np)
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


def test_embedding_matches_golden_fixture():
    with np.load(GOLDEN) as z:
        meta = json.loads(str(z["meta"]))
        _, audio, _ = O.synthetic_batch(meta["batch"], seed=meta["data_seed"])
        for mt in meta["embedding_types"]:
            w_np = O.init_weights(mt, seed=meta["weight_seed"], randomize_bn=True)
            eng = _engine(mt, meta["batch"], "f32", training=False, towers=("audio",), weights=w_np)
            assert np.abs(eng.embed_audio(audio, "short").cpu().numpy() - z[mt + "/embedding_short"]).max() <= 1e-3
            assert np.abs(eng.embed_audio(audio, "original").cpu().numpy() - z[mt + "/embedding_original"]).max() <= 1e-3


@pytest.mark.parametrize("model_type", ["cnn_L3_melspec2", "cnn_L3_orig"])
def test_training_step_gradients_f32(model_type):
    """train_on_batch: loss, accuracy, every gradient tensor, BN moving statistics and the Adam update against
    fp64 autograd of the oracle."""
    B = 2
    w_np = O.init_weights(model_type, seed=7, randomize_bn=True)
    video, audio, label = O.synthetic_batch(B, seed=303)
    vf, af = _oracle_inputs(video, audio)
    w = O.to_torch(w_np, dtype=torch.float64, requires_grad=True)
    grads, out, stats = O.compute_grads(vf, af, torch.from_numpy(label), w, model_type, F64)
    eng = _engine(model_type, B, "f32", training=True, weights=w_np)
    eng.forward_backward(video, audio, label)
    m = eng.metrics()
    assert abs(m["loss"] - float(out["loss"])) <= 1e-4 * max(1.0, abs(float(out["loss"])))
    assert abs(m["acc"] - float(out["acc"])) < 1e-6
    got = eng.get_grads()
    bad = []
    for name, g_ref in grads.items():
        g_ref = g_ref.numpy()
        if name.endswith("/kernel"):
            g_ref = g_ref - 2e-5 * w_np[name]          # the device applies the l2 term inside Adam
        err = rel_l2(got[name], g_ref)
        max_abs = float(np.abs(got[name] - g_ref).max())
        # fp32 noise floor, measured: the SAME graph in PyTorch-CPU fp32 deviates from its fp64 self by up to 5.8e-3
        # relative L2 on these gradients at B=2 (BN backward cancels large terms), and by up to 6e-5 absolute on the
        # analytically-zero ones (bias of a conv feeding training-mode BN).  Hence 1e-2 relative / 1e-4 absolute.
        tol = 1e-2
        if not (err <= tol or max_abs <= 1e-4):
            bad.append((name, err, max_abs, float(np.abs(g_ref).max())))
    assert not bad, bad
    # BN moving statistics (momentum 0.99, Bessel-corrected variance)
    O.update_moving_stats(w, stats, F64)
    w_after = eng.get_weights()
    for name in w_np:
        if name.endswith(("moving_mean", "moving_variance")):
            assert np.abs(w_after[name] - w[name].detach().numpy()).max() <= 1e-5, name
    # keras Adam step, isolated from gradient noise: apply the oracle's update rule to the DEVICE gradients (+ the l2
    # term the device folds into Adam) and compare the resulting weights.  (Feeding each side its own gradients is
    # meaningless for the analytically-zero ones: Adam normalises pure round-off noise to a +-lr step.)
    g_dev = {k: torch.from_numpy(v.astype(np.float64) + (2e-5 * w_np[k] if k.endswith("/kernel") else 0.0))
             for k, v in got.items()}
    w_ref = O.to_torch(w_np, dtype=torch.float64)
    O.adam_update(w_ref, g_dev, O.AdamState(), 1e-3, F64)
    eng.adam_step(1e-3)
    w_after = eng.get_weights()
    for name in g_dev:
        assert np.abs(w_after[name] - w_ref[name].numpy()).max() <= 2e-6, name


def test_data_parallel_replicas_match_oracle_replica_by_replica():
    """Row (e): N replicas x B/N samples with PER-REPLICA BatchNorm statistics (training_utils.py:121-170 semantics, no
    sync-BN).  Two replicas are emulated on one GPU: each slice runs forward/backward with global_batch = 4, the two
    gradient arenas are summed (what the NCCL all-reduce does) and compared with the oracle evaluated replica by
    replica and gradient-averaged over the global batch."""
    mt, G, R = "cnn_L3_melspec2", 4, 2
    w_np = O.init_weights(mt, seed=11, randomize_bn=True)
    video, audio, label = O.synthetic_batch(G, seed=909)
    eng = _engine(mt, G // R, "f32", training=True, weights=w_np)
    summed, ref = {}, {}
    loss_dev = loss_ref = 0.0
    for r in range(R):
        sl = slice(r * (G // R), (r + 1) * (G // R))          # contiguous slices, training_utils.py:121-133
        eng.forward_backward(video[sl], audio[sl], label[sl], global_batch=G)
        loss_dev += eng.metrics()["loss"] * (G // R) / G
        for k, v in eng.get_grads().items():
            summed[k] = summed.get(k, 0.0) + v.astype(np.float64)
        vf, af = _oracle_inputs(video[sl], audio[sl])
        w = O.to_torch(w_np, dtype=torch.float64, requires_grad=True)
        grads, out, _ = O.compute_grads(vf, af, torch.from_numpy(label[sl]), w, mt, F64)
        loss_ref += float(out["loss"]) * (G // R) / G
        for k, g in grads.items():
            g = g.numpy()
            if k.endswith("/kernel"):
                g = g - 2e-5 * w_np[k]                          # the l2 term is applied once, inside Adam
            ref[k] = ref.get(k, 0.0) + g * ((G // R) / G)       # mean over the slice -> share of the global mean
    assert abs(loss_dev - loss_ref) <= 1e-4 * max(1.0, abs(loss_ref)), (loss_dev, loss_ref)
    bad = []
    for k in ref:
        err, max_abs = rel_l2(summed[k], ref[k]), float(np.abs(summed[k] - ref[k]).max())
        # the input-BN gradients are residuals of large cancelling sums (see the gamma = 0 test below): 2e-2
        tol = 2e-2 if "/bn0/" in k else 1e-2
        if not (err <= tol or max_abs <= 1e-4):
            bad.append((k, round(err, 5), max_abs))
    print("replica test: loss (device, oracle)", loss_dev, loss_ref)
    assert not bad, bad


def test_input_bn_gradient_fallback_when_gamma_is_zero():
    """The input-BN gradients normally come from the first layer's weight gradient divided by gamma; a zero gamma
    must take the direct path and still match the oracle."""
    mt, B = "cnn_L3_kapredbinputbn", 2
    w_np = O.init_weights(mt, seed=9, randomize_bn=True)
    # only one of the three vision channels: zeroing the single audio channel would make the whole audio tower's
    # input constant and every later BatchNorm amplify pure round-off (an ill-conditioned comparison)
    w_np["vision/bn0/gamma"][1] = 0.0
    video, audio, label = O.synthetic_batch(B, seed=404)
    vf, af = _oracle_inputs(video, audio)
    w = O.to_torch(w_np, dtype=torch.float64, requires_grad=True)
    grads, _, _ = O.compute_grads(vf, af, torch.from_numpy(label), w, mt, F64)
    eng = _engine(mt, B, "f32", training=True, weights=w_np)
    eng.forward_backward(video, audio, label)
    got = eng.get_grads()
    for name in ("audio/bn0/gamma", "audio/bn0/beta", "vision/bn0/gamma", "vision/bn0/beta"):
        ref = grads[name].numpy()
        # the input-BN gradients are residuals of large cancelling sums accumulated with fp32 atomics in a
        # run-dependent order: measured 0.5e-2 .. 1.3e-2 from the fp64 oracle across runs of identical code
        assert rel_l2(got[name], ref) <= 2e-2 or np.abs(got[name] - ref).max() <= 1e-4, (name, got[name], ref)


def test_train_function_writes_reference_files_and_resumes(tmp_path):
    """train() (train.py:218-421): same output files, CSV history, and resume from model_latest.h5 + the CSV."""
    from l3embedding_b200 import train as T
    from l3embedding_b200.synthetic import synthetic_batch
    for name, seed in (("demo_train", 0), ("demo_valid", 50)):
        d = tmp_path / name
        d.mkdir()
        for i in range(2):
            v, a, l = synthetic_batch(4, seed=seed + i)
            np.savez(d / ("b%d.npz" % i), video=v, audio=a, label=l)
    kw = dict(train_epoch_size=2, validation_epoch_size=1, train_batch_size=2, validation_batch_size=2,
              model_type="cnn_L3_melspec2", learning_rate=1e-4, checkpoint_interval=1, disable_logging=True, gpus=1,
              dtype="bf16")
    model_dir, hist = T.train(str(tmp_path / "demo_train"), str(tmp_path / "demo_valid"), str(tmp_path / "out"),
                              num_epochs=2, **kw)
    files = set(os.listdir(model_dir))
    assert {"config.json", "model.json", "model_spec.pkl", "model_latest.h5", "model_best_valid_accuracy.h5",
            "model_best_valid_loss.h5", "model_checkpoint.01.h5", "model_checkpoint.02.h5", "history_checkpoint.pkl",
            "history_csvlog.csv", "history.pkl"} <= files
    assert "embedding/demo/cnn_L3_melspec2" in model_dir.replace(os.sep, "/")
    assert len(hist.history["loss"]) == 2 and set(hist.history) == {"loss", "acc", "val_loss", "val_acc"}
    assert T.get_restart_info(os.path.join(model_dir, "history_csvlog.csv"))[0] == 1
    _, hist2 = T.train(str(tmp_path / "demo_train"), str(tmp_path / "demo_valid"), str(tmp_path / "out"), num_epochs=3,
                       continue_model_dir=model_dir, **kw)
    assert len(hist2.history["loss"]) == 1                      # only epoch index 2 ran
    assert T.get_restart_info(os.path.join(model_dir, "history_csvlog.csv"))[0] == 2


def test_device_side_framing_equals_host_framing():
    """get_l3_frames_uniform: embeddings of overlapping 1 s windows read in place on the device == the reference's
    framed-copy route, for int16 and float32 signals, across a max_batch boundary."""
    from l3embedding_b200 import model as M
    from l3embedding_b200.features import get_l3_frames_uniform, frame_signal
    m, _, _ = M.MODELS["cnn_L3_melspec2"]()
    m.configure(dtype="f32")
    m.set_named_weights(O.init_weights("cnn_L3_melspec2", seed=5, randomize_bn=True))
    e, _, _ = M.convert_audio_model_to_embedding(m.get_layer("audio_model"), m.inputs[1], "cnn_L3_melspec2", "short")
    rng = np.random.default_rng(0)
    sig = (0.1 * rng.standard_normal(48000 + 6 * 4800 + 100)).astype(np.float32)
    got = get_l3_frames_uniform(sig, e, hop_size=0.1)
    a, hop, n = frame_signal(sig)
    assert got.shape == (7, 512) and n == 7
    idx = np.arange(48000)[None, :] + hop * np.arange(n)[:, None]
    ref = e.predict(a[idx].reshape(n, 1, 48000), batch_size=4)
    assert np.abs(got - ref).max() <= 1e-5
    w = O.to_torch(m.named_weights(), dtype=torch.float64)
    orc = O.audio_embedding(torch.from_numpy(a[idx].reshape(n, 1, 48000)).double(), w, "cnn_L3_melspec2", "short", F64).numpy()
    assert np.abs(got - orc).max() <= 1e-3
    eng = e._get_engine(4)       # smaller than n: exercises the chunk loop
    i16 = np.clip(np.round(sig * 32767), -32768, 32767).astype(np.int16)
    g16 = eng.embed_audio_frames(i16, hop, "short").cpu().numpy()
    r16 = eng.embed_audio(i16[idx].reshape(n, 1, 48000)[:4], "short").cpu().numpy()
    assert np.abs(g16[:4] - r16).max() <= 1e-5


def test_train_steps_from_host_decrease_loss():
    """BASELINE config 1 on the device path: cnn_L3_orig, batch 4, 8 steps from host buffers."""
    mt = "cnn_L3_orig"
    video, audio, label = O.synthetic_batch(4, seed=11)
    eng = _engine(mt, 4, "f32", training=True)
    losses = [eng.train_step_host(video, audio, label, 1e-4)["loss"] for _ in range(8)]
    assert np.isfinite(losses).all() and losses[-1] < losses[0]


def test_keras_style_api_end_to_end(tmp_path):
    from l3embedding_b200 import model as M
    m, _, _ = M.MODELS["cnn_L3_melspec2"]()
    m.configure(dtype="f32")
    m.compile(M.Adam(lr=1e-4), loss="categorical_crossentropy", metrics=["accuracy"])
    video, audio, label = O.synthetic_batch(2, seed=17)

    def gen():
        while True:
            yield [O.scale_video(video), O.pcm2float(audio, "float32")], label
    h = m.fit_generator(gen(), steps_per_epoch=2, epochs=2, validation_data=gen(), validation_steps=1, verbose=0)
    assert set(h.history) == {"loss", "acc", "val_loss", "val_acc"} and len(h.history["loss"]) == 2
    p = str(tmp_path / "model_latest.h5")
    m.save_weights(p)
    e = M.load_embedding(p, "cnn_L3_melspec2", "audio", "original")
    e.parent.configure(dtype="f32")
    x = O.pcm2float(audio, "float32")
    emb = e.predict(x)
    assert emb.shape == (2, 6144)
    w = O.to_torch(m.named_weights(), dtype=torch.float64)
    ref = O.audio_embedding(torch.from_numpy(x).double(), w, "cnn_L3_melspec2", "original", F64).numpy()
    assert np.abs(emb - ref).max() <= 1e-3


@pytest.mark.parametrize("model_type", MODEL_TYPES)
@pytest.mark.parametrize("batch", [1, 3])
def test_bf16_tensor_core_path_tracks_f32_path(model_type, batch):
    """Every model type and odd batch sizes through the tcgen05 path (odd widths 197/199/99/49, 'valid' pooling that
    drops columns, ReLU-before-BN layer): logits, gradients and a training step against the fp32 SIMT path of the same
    library on identical inputs.  Bars are bf16-sized: 5 % of the largest logit / of the loss (gradients: see the bf16-emulating-oracle test)."""
    w_np = O.init_weights(model_type, seed=3, randomize_bn=True)
    video, audio, label = O.synthetic_batch(batch, seed=505)
    res = {}
    for dt in ("f32", "bf16"):
        eng = _engine(model_type, batch, dt, training=True, weights=w_np)
        _, logits = eng.predict(video, audio)
        eng.forward_backward(video, audio, label)
        m = eng.metrics()
        res[dt] = (logits, eng.get_grads(), m)
        if dt == "bf16":
            assert eng.uses_tensor_cores
        eng.close()
    lf, gf, mf = res["f32"]
    lb, gb, mb = res["bf16"]
    assert np.abs(lb - lf).max() <= 0.05 * max(1.0, np.abs(lf).max()), (lb, lf)
    assert abs(mb["loss"] - mf["loss"]) <= 0.05 * max(1.0, abs(mf["loss"]))
    assert all(np.isfinite(v).all() for v in gb.values())


@pytest.mark.parametrize("model_type", MODEL_TYPES)
def test_bf16_path_matches_bf16_emulating_oracle(model_type):
    """Throughput-mode parity proper: the oracle rounds to bfloat16 at exactly the points where the device stores bf16
    (OracleConfig.emulate_bf16), so the two differ only by fp32 accumulation order.  Layer by layer (training-mode
    forward, all real shapes incl. the odd widths 197/199/99/49 and the ReLU-before-BN layer):
      * every conv output z agrees within 2 (first two layers) / 4 bf16 ulps of the layer's largest value,
      * the first two layers are bit-identical in >= 99 % / 97 % of their elements (an accumulation-order difference only
        shows when it straddles a bf16 rounding boundary; deeper layers inherit and multiply those 1-ulp flips --
        measured 1e-4 -> 6e-4 -> 1.5e-2 -> 0.12 -> ... of the elements, always by one ulp),
      * logits within 0.06, gradient tensors point the same way (cosine >= 0.85; a 1-ulp flip that changes a ReLU /
        max-pool decision re-routes a whole gradient path, so magnitudes are not comparable on a random network)."""
    import torch.nn.functional as F
    B = 3
    w_np = O.init_weights(model_type, seed=3, randomize_bn=True)
    video, audio, label = O.synthetic_batch(B, seed=505)
    cfg = O.OracleConfig(dtype=torch.float32, emulate_bf16=True)
    w = O.to_torch(w_np, dtype=torch.float32, requires_grad=True)
    vf = torch.from_numpy(O.scale_video(video))
    af = torch.from_numpy(O.pcm2float(audio, "float32"))
    grads, out, _ = O.compute_grads(vf, af, torch.from_numpy(label), w, model_type, cfg)
    eng = _engine(model_type, B, "bf16", training=True, weights=w_np)
    assert eng.uses_tensor_cores
    eng.forward_backward(video, audio, label)
    got = eng.get_grads()
    logits = eng.debug_read("logits", B).reshape(B, 2)
    with torch.no_grad():
        wd = O.to_torch(w_np)
        for tower, x in (("vision", vf), ("audio", O.frontend(af, model_type, cfg))):
            spec = (O.AUDIO_SPECS if tower == "audio" else O.VISION_SPECS)[model_type]
            x = x.permute(0, 3, 1, 2)
            if spec["input_bn"]:
                x = O._bn(x, wd, f"{tower}/bn0", True, cfg, {})
            for i, nm in enumerate(O.CONV_NAMES):
                z = O._conv(x, wd, f"{tower}/{nm}", cfg)
                zo = z.permute(0, 2, 3, 1).numpy()
                zd = eng.debug_read(f"{tower}/z{i}", B).reshape(zo.shape)
                d = np.abs(zd - zo)
                ulp = 2.0 ** (np.floor(np.log2(np.abs(zo).max())) - 7)      # bf16: 8 significant bits
                assert d.max() <= (2 if i < 2 else 4) * ulp, (tower, i, d.max(), ulp)   # deeper: several 1-ulp inputs add up
                if i < 2:
                    assert (d > 0).mean() <= (1e-2 if i == 0 else 3e-2), (tower, i, (d > 0).mean())
                bnn = f"{tower}/bn{nm[4:]}"
                if tower == "vision" and nm == "conv1b":
                    x = O._bn(F.relu(z), wd, bnn, True, cfg, {})
                else:
                    x = F.relu(O._bn(z, wd, bnn, True, cfg, {}))
                if nm in ("conv1b", "conv2b", "conv3b"):
                    x = O._pool_same(x, 2, 2) if tower == "vision" else F.max_pool2d(x, 2, 2)
    d_logit = float(np.abs(logits - out["logits"].numpy()).max())
    rows = []
    for name in ("dense_1/kernel", "vision/conv4b/kernel", "audio/conv4b/kernel", "vision/conv3a/kernel",
                 "audio/conv2b/kernel", "vision/conv1b/kernel", "audio/conv1b/kernel", "audio/conv1a/kernel",
                 "vision/conv1a/kernel", "vision/bn2a/gamma", "audio/bn3b/beta"):
        a = got[name].ravel().astype(np.float64)
        g = grads[name].detach().numpy().astype(np.float64)
        if name.endswith("/kernel"):
            g = g - 2e-5 * w_np[name]
        g = g.ravel()
        rows.append((name, round(float(a @ g / max(np.linalg.norm(a) * np.linalg.norm(g), 1e-30)), 4), round(rel_l2(a, g), 4)))
    print(model_type, "logits max|d| %.3g" % d_logit, rows)
    assert d_logit <= 0.06
    # measured over the kernel configurations of this round: 0.89 .. 0.96 for every conv kernel / BN parameter gradient
    # (the same tensors agree to 0.975 .. 0.99 between two device configurations that only differ in summation order, and
    # to 0.93 between the device and itself with other kernel variants: tests/test_gpu_tc.py) -- the spread is the 1-ulp
    # re-routing noise of bf16 storage on a random network at B = 3, not a property of any one kernel
    assert all(r[1] >= 0.85 for r in rows), rows


@pytest.mark.parametrize("model_type", ["cnn_L3_melspec2"])
def test_bf16_throughput_mode_reports_error(model_type):
    """bf16 storage/operands cannot meet 1e-3; bound it loosely and print the measured error."""
    B = 2
    w_np = O.init_weights(model_type, seed=20180123, randomize_bn=True)
    video, audio, label = O.synthetic_batch(B, seed=202)
    vf, af = _oracle_inputs(video, audio)
    w = O.to_torch(w_np, dtype=torch.float64)
    ref = O.audio_embedding(af, w, model_type, "original", F64).numpy()
    eng = _engine(model_type, B, "bf16", training=True, weights=w_np)
    got = eng.embed_audio(audio, "original").cpu().numpy()
    err = np.abs(got - ref).max()
    print("bf16 embedding max|d| = %.4g (|e|max %.3g), tensor cores: %s" % (err, np.abs(ref).max(), eng.uses_tensor_cores))
    assert err <= 0.02 * np.abs(ref).max()      # bf16 operands + bf16 activations: ~1 % of the largest value
    eng.forward_backward(video, audio, label)
    assert np.isfinite(eng.metrics()["loss"])
    g = eng.get_grads()
    assert all(np.isfinite(v).all() for v in g.values())
