"""CPU tests of the oracle (oracle/l3_oracle.py) against every pin the reference offers for this path
(param counts, layer shapes, pooling table, empty mel rows, front-end algebra) and against the
committed golden vectors.  The reference ships no numerical tests: parity is UNPINNED beyond these."""
import json
import math
import os

import numpy as np
import pytest
import torch

from oracle import l3_oracle as O

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
MODEL_TYPES = ["cnn_L3_orig", "cnn_L3_kapredbinputbn", "cnn_L3_melspec1", "cnn_L3_melspec2"]


def test_param_counts_match_notebook_pin():
    # notebooks/test_load_converted_model.ipynb:110-123 (older commit, no input BN): trainable 9 508 738,
    # vision 4 693 056, audio 9 021 504 (melspec1: 128 mels), dense_1 131 200, dense_2 258.  HEAD adds the input
    # BNs: +2 trainable +2 moving (audio), +6 +6 (vision).
    c = O.count_params("cnn_L3_melspec1")
    assert c["trainable"] == 9508738 + 8
    lay = O.model_layout("cnn_L3_melspec1")
    vision = sum(int(np.prod(s)) for n, s, _ in lay if n.startswith("vision/"))
    audio = sum(int(np.prod(s)) for n, s, _ in lay if n.startswith("audio/"))
    assert vision == 4693056 + 12
    assert audio + c["kapre_constants"] == 9021504 + 4
    assert c["kapre_constants"] == 2 * 2048 * 1025 + 1025 * 128      # melspectrogram_1: 4 329 600
    c2 = O.count_params("cnn_L3_melspec2")
    assert c2["trainable"] == 9508746 and c2["bn_moving"] == 7688 and c2["kapre_constants"] == 4460800


def test_frame_geometry():
    # SAME: ceil(48000/242) = 199 frames, pad 982/982 ; VALID n_dft 512: 197 frames
    assert O.frame_geometry(48000, 2048, 242, "same") == (199, 982)
    assert O.frame_geometry(48000, 512, 242, "valid") == (197, 0)


def test_mel_filterbank_empty_rows_and_sparsity():
    # notebooks/extract_embedding_models_from_avc_models.ipynb:66-67: librosa warns about empty filters
    fb = O.mel_filterbank(48000, 2048, 256)
    empty = np.where(fb.sum(axis=1) == 0)[0].tolist()
    assert empty == [0, 7]
    assert (np.count_nonzero(fb, axis=0) <= 2).all()
    assert np.where(O.mel_filterbank(48000, 2048, 128).sum(axis=1) == 0)[0].size == 0


def test_mel_filterbank_matches_torchaudio():
    ta = pytest.importorskip("torchaudio")
    ref = ta.functional.melscale_fbanks(n_freqs=1025, f_min=0.0, f_max=24000.0, n_mels=256, sample_rate=48000,
                                        norm="slaney", mel_scale="htk").numpy().T
    assert np.abs(O.mel_filterbank(48000, 2048, 256) - ref).max() < 1e-5


@pytest.mark.parametrize("model_type", MODEL_TYPES)
def test_frontend_fft_equals_kapre_dft_convolution(model_type):
    """kapre evaluates the STFT as strided convolutions with cos/sin kernels (Appendix B items 1-3); the oracle
    uses rfft of the same zero-padded windowed frames.  Check the algebraic identity on real frames (fp64)."""
    a = O.AUDIO_SPECS[model_type]
    rng = np.random.default_rng(3)
    x = rng.standard_normal(48000) * 0.1
    n, hop = a["n_dft"], a["n_hop"]
    n_frames, left = O.frame_geometry(48000, n, hop, a["padding"])
    xp = np.zeros(max((n_frames - 1) * hop + n, left + 48000))
    xp[left:left + 48000] = x
    win = O.hann_periodic(n)
    t = np.arange(n)[:, None]
    k = np.arange(n // 2 + 1)[None, :]
    real_k = np.cos(2 * np.pi * k * t / n) * win[:, None]
    imag_k = -np.sin(2 * np.pi * k * t / n) * win[:, None]
    for f in (0, 1, n_frames // 2, n_frames - 1):
        fr = xp[f * hop:f * hop + n]
        power_conv = (fr @ real_k) ** 2 + (fr @ imag_k) ** 2
        spec = np.fft.rfft(fr * win)
        assert np.allclose(power_conv, spec.real ** 2 + spec.imag ** 2, rtol=1e-9, atol=1e-12)
    out = O.frontend(torch.from_numpy(x).reshape(1, 1, -1), model_type, O.OracleConfig(dtype=torch.float64))
    n_out = a.get("n_mels") if a["kind"] == "mel" else n // 2 + 1
    assert tuple(out.shape) == (1, n_out, n_frames, 1)
    if a["decibel"]:
        assert float(out.max()) == 0.0 and float(out.min()) >= -80.0


def test_pcm2float_and_video_scaling():
    s = np.array([-32768, -1, 0, 1, 32767], dtype=np.int16)
    assert np.array_equal(O.pcm2float(s, "float32"), s.astype(np.float32) / 32768)
    with pytest.raises(TypeError):
        O.pcm2float(np.zeros(3, np.float32))
    v = np.array([0, 127, 255], dtype=np.uint8)
    assert np.allclose(O.scale_video(v), [-1.0, 2 * 127 / 255 - 1, 1.0], atol=1e-7)


@pytest.mark.parametrize("model_type", MODEL_TYPES)
def test_layer_shapes_and_pool_table(model_type):
    """Per-layer shapes of notebooks/test_load_converted_model.ipynb:156-214 and the pooling table
    audio_model.py:461-478: every embedding is 6144-d ('original') or 512-d ('short'); vision 8192-d."""
    w = O.to_torch(O.init_weights(model_type, seed=1))
    a = O.AUDIO_SPECS[model_type]
    n_out = a.get("n_mels") if a["kind"] == "mel" else a["n_dft"] // 2 + 1
    n_frames, _ = O.frame_geometry(48000, a["n_dft"], a["n_hop"], a["padding"])
    x = torch.zeros(1, n_out, n_frames, 1)
    z = O.tower_forward(x, w, "audio", model_type, False, return_embedding_map=True)
    eh, ew = {"cnn_L3_melspec1": (16, 24)}.get(model_type, (32, 24))
    assert tuple(z.shape) == (1, eh, ew, 512)
    for pooling, dim in (("original", 6144), ("short", 512)):
        ph, pw = O.EMBED_POOL[model_type][pooling]
        assert (eh // ph) * (ew // pw) * 512 == dim
    assert tuple(O.tower_forward(x, w, "audio", model_type, False).shape) == (1, 512)


def test_train_step_decreases_loss_small():
    """BASELINE config 1 in miniature: cnn_L3_orig plumbing on CPU (2 pairs, 2 steps)."""
    mt = "cnn_L3_orig"
    w = O.to_torch(O.init_weights(mt, seed=5), requires_grad=True)
    video, audio, label = O.synthetic_batch(2, seed=11)
    st = O.AdamState()
    l0 = O.train_step(video, audio, label, w, st, mt, lr=1e-4)["loss"]
    l1 = O.train_step(video, audio, label, w, st, mt, lr=1e-4)["loss"]
    assert math.isfinite(float(l0)) and float(l1) < float(l0)


def test_dp_emulation_matches_manual_replicas():
    """compute_grads(n_replicas=2) == mean-loss gradient with BN statistics per contiguous slice
    (training_utils.py:121-162)."""
    mt = "cnn_L3_orig"
    w = O.to_torch(O.init_weights(mt, seed=5), requires_grad=True)
    video, audio, label = O.synthetic_batch(2, seed=13)
    vf = torch.from_numpy(O.scale_video(video))
    af = torch.from_numpy(O.pcm2float(audio, "float32"))
    lab = torch.from_numpy(label)
    g2, out2, _ = O.compute_grads(vf, af, lab, w, mt, n_replicas=2)
    logits = torch.cat([O.avc_forward(vf[i:i + 1], af[i:i + 1], w, mt, True) for i in range(2)])
    assert torch.allclose(logits, out2["logits"], atol=1e-5)
    g1, _, _ = O.compute_grads(vf, af, lab, w, mt, n_replicas=1)
    # per-replica BN changes the result: the two must differ (guards against silently syncing statistics)
    assert not torch.allclose(g1["dense_1/kernel"], g2["dense_1/kernel"], atol=1e-7)


def test_adam_matches_keras_formula():
    w = {"p": torch.tensor([1.0, -2.0])}
    st = O.AdamState()
    g = {"p": torch.tensor([0.5, -0.25])}
    O.adam_update(w, g, st, lr=0.1)
    lr_t = 0.1 * math.sqrt(1 - 0.999) / (1 - 0.9)
    m = 0.1 * np.array([0.5, -0.25])
    v = 0.001 * np.array([0.25, 0.0625])
    exp = np.array([1.0, -2.0]) - lr_t * m / (np.sqrt(v) + 1e-8)
    assert np.allclose(w["p"].numpy(), exp, rtol=1e-6)


def test_golden_vectors():
    """Oracle outputs committed by tests/golden/make_golden.py (fp64 oracle, seeded inputs)."""
    path = os.path.join(GOLDEN, "oracle_golden.npz")
    with np.load(path) as z:
        meta = json.loads(str(z["meta"]))
        for mt in meta["model_types"]:
            video, audio, label = O.synthetic_batch(meta["batch"], seed=meta["data_seed"])
            cfg = O.OracleConfig(dtype=torch.float64)
            af = torch.from_numpy(O.pcm2float(audio, "float64"))
            fe = O.frontend(af, mt, cfg)[..., 0].numpy()
            assert np.abs(fe[:, ::meta["stride_f"], ::meta["stride_t"]] - z[mt + "/frontend"]).max() < 1e-9
            if mt in meta["embedding_types"]:
                w = O.to_torch(O.init_weights(mt, seed=meta["weight_seed"], randomize_bn=True), dtype=torch.float64)
                emb = O.audio_embedding(af, w, mt, "short", cfg).numpy()
                assert np.abs(emb - z[mt + "/embedding_short"]).max() < 1e-8


def test_emulate_bf16_first_layer_weight_switch():
    """OracleConfig.first_layer_bf16_weights: the bf16-emulating oracle rounds the Cin = 1 / 3 kernels like the tensor-core
    first layer does (default) or keeps them fp32 like the SIMT first layer (L3_FIRST_CONV_TC=0)."""
    torch.manual_seed(0)
    x = torch.randn(1, 3, 9, 11)
    w = {"t/kernel": torch.randn(3, 3, 3, 64) * 0.2 + 1e-3, "t/bias": torch.zeros(64)}
    on = O._conv(x, w, "t", O.OracleConfig(emulate_bf16=True))
    off = O._conv(x, w, "t", O.OracleConfig(emulate_bf16=True, first_layer_bf16_weights=False))
    wq = {"t/kernel": w["t/kernel"].bfloat16().float(), "t/bias": w["t/bias"]}
    pre = O._conv(x, wq, "t", O.OracleConfig(emulate_bf16=True, first_layer_bf16_weights=False))
    assert torch.equal(on, pre) and not torch.equal(on, off)
    # Cin >= 64 layers are rounded either way
    x64 = torch.randn(1, 64, 5, 5)
    w64 = {"t/kernel": torch.randn(3, 3, 64, 64) * 0.05, "t/bias": torch.zeros(64)}
    assert torch.equal(O._conv(x64, w64, "t", O.OracleConfig(emulate_bf16=True)),
                       O._conv(x64, w64, "t", O.OracleConfig(emulate_bf16=True, first_layer_bf16_weights=False)))


def _pimodel_frontend(frame, model_type):
    """The authors' numpy front-end, restated from notebooks/pimodel.ipynb (cell 12, lines 303-319: librosa-style
    `stft(frame, n_fft, hop_length=242, window='hann', center=True, pad_mode='constant')`, magnitude, for the mel models
    `melspectrogram(sr=48000, S=S, n_mels, power=1.0, htk=True)` = mel_basis . S; cell 4, lines 139-148:
    `amplitude_to_db` which squares the magnitude IN PLACE and then takes 10*log10(max(amin, .)) of that same array,
    subtracts the per-frame-set maximum and floors at -80).  Plain numpy, float64, one 1 s frame."""
    n_fft = 2048 if model_type in ("cnn_L3_melspec1", "cnn_L3_melspec2") else 512
    hop = 242
    y = np.pad(np.asarray(frame, np.float64), n_fft // 2, mode="constant")           # center=True, pad_mode='constant'
    win = 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n_fft) / n_fft)                 # get_window('hann', fftbins=True)
    n_frames = 1 + (len(y) - n_fft) // hop
    S = np.empty((n_fft // 2 + 1, n_frames))
    for t in range(n_frames):
        S[:, t] = np.abs(np.fft.rfft(y[t * hop:t * hop + n_fft] * win))
    if n_fft == 2048:
        n_mels = 256 if model_type == "cnn_L3_melspec2" else 128
        S = O.mel_filterbank(48000, n_fft, n_mels) @ S
    magnitude = np.abs(S)
    np.square(magnitude, out=magnitude)                                                # the in-place square
    log_spec = 10.0 * np.log10(np.maximum(1e-10, magnitude))
    log_spec -= log_spec.max()
    return np.maximum(log_spec, -80.0)


@pytest.mark.parametrize("model_type", ["cnn_L3_melspec2", "cnn_L3_melspec1", "cnn_L3_kapredbinputbn"])
def test_frontend_reproduces_the_authors_numpy_restatement(model_type):
    """notebooks/pimodel.ipynb is the only front-end ARITHMETIC held by the reference tree.  It is not the
    kapre graph: it centres the STFT (1024 zeros each side instead of TF-SAME's 982 / 'valid'), applies the mel basis to
    the magnitude (kapre: sqrt of the mel of the power) and -- through an in-place square -- scales by 20*log10 with amin
    on the squared value (kapre: 10*log10 of the amplitude).  With exactly those three switches (+ db_multiplier 20) the
    oracle's front-end reproduces it (to float64 round-off; 1e-5 dB where the float32 mel basis enters), which pins everything else the two share: n_fft 2048 / 512,
    hop 242, periodic hann, the htk / Slaney-normalised mel basis incl. its two empty rows, amin 1e-10, the 80 dB range
    and the per-window maximum.  The defaults of OracleConfig remain the kapre semantics of SURVEY App. B."""
    _, audio, _ = O.synthetic_batch(2, seed=5)
    x = O.pcm2float(audio, "float64")
    cfg = O.OracleConfig(dtype=torch.float64, stft_center=True, mel_on_magnitude=True, db_amin_on_square=True,
                         db_multiplier=20.0)
    got = O.frontend(torch.from_numpy(x), model_type, cfg)[..., 0].numpy()
    for b in range(2):
        want = _pimodel_frontend(x[b, 0], model_type)
        assert got[b].shape == want.shape == ((256 if model_type == "cnn_L3_melspec2" else 128 if model_type == "cnn_L3_melspec1" else 257), 199)
        # 1e-5 dB: the oracle keeps the mel basis in float32, as the kapre layer stores it (measured 6.5e-7 dB; the
        # spectrogram model, which has no basis, agrees to 1e-9)
        assert np.abs(got[b] - want).max() <= (1e-5 if "melspec" in model_type else 1e-9)
    # the default (kapre) semantics are a different map: the switches matter (a 2x dB scale changes what the trained
    # input BN sees), which is why each stays an explicit OracleConfig field
    kapre = O.frontend(torch.from_numpy(x), model_type, O.OracleConfig(dtype=torch.float64))[..., 0].numpy()
    if "melspec" in model_type:
        assert kapre.shape == got.shape and np.abs(kapre - got).max() > 1.0
