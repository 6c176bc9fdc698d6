"""CPU oracle for the L3-Net AVC hot path -- TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: the reference (marl/l3embedding @ 8a0d31b) is pure Python over
keras==2.0.9 / tensorflow==1.4.0 / kapre==0.1.3.1|0.1.4 / librosa==0.5.1, none of which
is vendored under /root/reference or installable here, and the reference ships no tests
or golden vectors for this path.  This file restates the published algorithms of those
pinned versions and anchors on the reference's own call sites.  Structural pins that DO
exist (param counts, layer shapes, pooling table, empty mel rows) are checked in
tests/test_oracle.py.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference leg may import
this module.  The product package (l3embedding_b200) never does.

Reference sites followed (relative to /root/reference):
  l3embedding/audio.py:21-31            pcm2float
  l3embedding/train.py:186,189          video / audio input scaling
  l3embedding/audio_model.py:8-115      cnn_L3_orig audio tower (Spectrogram n_dft 512 valid, log/5)
  l3embedding/audio_model.py:118-223    kapredbinputbn audio tower (Spectrogram dB + input BN)
  l3embedding/audio_model.py:225-332    melspec1 audio tower (128 mels)
  l3embedding/audio_model.py:335-442    melspec2 audio tower (256 mels)
  l3embedding/audio_model.py:445-487    convert_audio_model_to_embedding (pool table)
  l3embedding/vision_model.py:7-99      orig vision tower
  l3embedding/vision_model.py:102-195   orig_inputbn vision tower (block-1b ReLU->BN swap :135-139)
  l3embedding/vision_model.py:198-218   vision embedding head
  l3embedding/model.py:7-35             concat + Dense128 relu + Dense2 softmax, L2 1e-5
  l3embedding/train.py:270-284          categorical_crossentropy + accuracy, Adam(lr)
  l3embedding/training_utils.py:121-170 multi_gpu_model batch slicing, per-replica BN
Third-party semantics are explicit switches in OracleConfig.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# Model specifications (restated from the builder functions cited above)
# --------------------------------------------------------------------------------------

AUDIO_SPECS = {
    # audio_model.py:26-43
    "cnn_L3_orig": dict(kind="spec", n_dft=512, n_hop=242, padding="valid", decibel=False,
                        log_lambda=True, input_bn=False, final_pool=(32, 24)),
    # audio_model.py:138-151
    "cnn_L3_kapredbinputbn": dict(kind="spec", n_dft=512, n_hop=242, padding="valid", decibel=True,
                                  log_lambda=False, input_bn=True, final_pool=(32, 24)),
    # audio_model.py:245-260
    "cnn_L3_melspec1": dict(kind="mel", n_dft=2048, n_hop=242, n_mels=128, padding="same", decibel=True,
                            log_lambda=False, input_bn=True, final_pool=(16, 24)),
    # audio_model.py:355-370
    "cnn_L3_melspec2": dict(kind="mel", n_dft=2048, n_hop=242, n_mels=256, padding="same", decibel=True,
                            log_lambda=False, input_bn=True, final_pool=(32, 24)),
}
VISION_SPECS = {
    "cnn_L3_orig": dict(input_bn=False),            # model.py:212 -> vision_model.py:7
    "cnn_L3_kapredbinputbn": dict(input_bn=True),   # model.py:234 -> vision_model.py:102
    "cnn_L3_melspec1": dict(input_bn=True),
    "cnn_L3_melspec2": dict(input_bn=True),
}
# audio_model.py:461-478
EMBED_POOL = {
    "cnn_L3_orig": {"original": (8, 8), "short": (32, 24)},
    "cnn_L3_kapredbinputbn": {"original": (8, 8), "short": (32, 24)},
    "cnn_L3_melspec1": {"original": (4, 8), "short": (16, 24)},
    "cnn_L3_melspec2": {"original": (8, 8), "short": (32, 24)},
}
VISION_EMBED_POOL = (7, 7)  # vision_model.py:212
CONV_CHANNELS = [(None, 64), (64, 64), (64, 128), (128, 128), (128, 256), (256, 256), (256, 512), (512, 512)]
CONV_NAMES = ["conv1a", "conv1b", "conv2a", "conv2b", "conv3a", "conv3b", "conv4a", "conv4b"]
SR = 48000
WEIGHT_DECAY = 1e-5


@dataclass
class OracleConfig:
    """Switches for the third-party semantics that could not be pinned."""
    db_ref: str = "per_sample"            # kapre amplitude_to_decibel max axis: per_sample | per_batch
    db_multiplier: float = 10.0           # 10*log10(amplitude) (kapre) ; 20 = notebooks/pimodel.ipynb variant
    # The authors' own numpy restatement of the front-end (notebooks/pimodel.ipynb cells 4 and 12) differs from the
    # kapre graph in three places; with all three switched (and db_multiplier=20) this oracle reproduces the notebook's
    # arithmetic (tests/test_oracle.py::test_frontend_reproduces_the_authors_numpy_restatement) -- the one piece of
    # reference-held front-end arithmetic there is, and the pin for n_fft / hop / window / mel / amin / range / max axis:
    stft_center: bool = False             # librosa-style centring: pad n_fft//2 zeros on both sides, 1 + L//hop frames
    mel_on_magnitude: bool = False        # mel(|STFT|)  instead of kapre's sqrt(mel(|STFT|^2))
    db_amin_on_square: bool = False       # amin floors the squared value (10*log10(max(amin, x^2))), not x
    bn_eps: float = 1e-3                  # keras BatchNormalization default
    bn_momentum: float = 0.99
    bn_moving_var_unbiased: bool = True   # TF fused BN feeds Bessel-corrected var to the moving average
    adam_beta1: float = 0.9
    adam_beta2: float = 0.999
    adam_eps: float = 1e-8                # keras 2.0.9 K.epsilon()=1e-7? -> see note below
    dtype: torch.dtype = torch.float32
    # Emulate the bf16 throughput mode of the CUDA path: round to bfloat16 exactly where the device stores bf16
    # (conv inputs, conv outputs z, conv weights of the tensor-core layers, and in backward dz / da); all arithmetic
    # stays in `dtype`.  With this on, the oracle differs from the device only by accumulation order.
    emulate_bf16: bool = False
    # emulate_bf16 only: the first conv layer (Cin 1/3) also runs on the tensor cores with bf16-rounded weights
    # (k_first_conv_tc, the default); False = the SIMT first layer with fp32 weights (L3_FIRST_CONV_TC=0)
    first_layer_bf16_weights: bool = True


# Note on adam_eps: keras 2.0.9 Adam(epsilon=1e-8) is the constructor default; the update is
# p - lr_t * m / (sqrt(v) + eps) with lr_t = lr*sqrt(1-b2^t)/(1-b1^t).

# --------------------------------------------------------------------------------------
# Weight inventory (Keras array shapes: conv HWIO, dense (in,out), BN 4x(C))
# --------------------------------------------------------------------------------------

def tower_layout(tower: str, model_type: str) -> List[Tuple[str, Tuple[int, ...], bool]]:
    """Ordered (name, shape, trainable) for one tower, in Keras *layer* order
    (each layer: trainable arrays then non-trainable), kapre constants excluded."""
    cin0 = 1 if tower == "audio" else 3
    input_bn = (AUDIO_SPECS if tower == "audio" else VISION_SPECS)[model_type]["input_bn"]
    out: List[Tuple[str, Tuple[int, ...], bool]] = []

    def bn(name, c):
        out.extend([(f"{tower}/{name}/gamma", (c,), True), (f"{tower}/{name}/beta", (c,), True),
                    (f"{tower}/{name}/moving_mean", (c,), False), (f"{tower}/{name}/moving_variance", (c,), False)])

    if input_bn:
        bn("bn0", cin0)
    for nm, (ci, co) in zip(CONV_NAMES, CONV_CHANNELS):
        ci = cin0 if ci is None else ci
        out.append((f"{tower}/{nm}/kernel", (3, 3, ci, co), True))
        out.append((f"{tower}/{nm}/bias", (co,), True))
        bn("bn" + nm[4:], co)
    return out


def model_layout(model_type: str) -> List[Tuple[str, Tuple[int, ...], bool]]:
    lay = tower_layout("vision", model_type) + tower_layout("audio", model_type)
    lay += [("dense_1/kernel", (1024, 128), True), ("dense_1/bias", (128,), True),
            ("dense_2/kernel", (128, 2), True), ("dense_2/bias", (2,), True)]
    return lay


def count_params(model_type: str) -> Dict[str, int]:
    lay = model_layout(model_type)
    tr = sum(int(np.prod(s)) for _, s, t in lay if t)
    nt = sum(int(np.prod(s)) for _, s, t in lay if not t)
    a = AUDIO_SPECS[model_type]
    nfreq = a["n_dft"] // 2 + 1
    kapre = 2 * a["n_dft"] * nfreq + (nfreq * a["n_mels"] if a["kind"] == "mel" else 0)
    return dict(trainable=tr, bn_moving=nt, kapre_constants=kapre, total=tr + nt + kapre)


def truncated_normal(rng: np.random.Generator, shape, std):
    """Keras/TF truncated_normal: resample outside +-2 sigma."""
    x = rng.standard_normal(shape)
    bad = np.abs(x) > 2.0
    while bad.any():
        x[bad] = rng.standard_normal(int(bad.sum()))
        bad = np.abs(x) > 2.0
    return (x * std).astype(np.float32)


def init_weights(model_type: str, seed: int = 20180123, randomize_bn: bool = False) -> Dict[str, np.ndarray]:
    """he_normal kernels (fan_in = kh*kw*cin), zero biases, BN gamma 1 / beta 0 / mean 0 / var 1.
    randomize_bn=True perturbs every BN array so inference parity exercises all of them."""
    rng = np.random.default_rng(seed)
    w: Dict[str, np.ndarray] = {}
    for name, shape, _ in model_layout(model_type):
        leaf = name.rsplit("/", 1)[1]
        if leaf == "kernel":
            fan_in = int(np.prod(shape[:-1]))
            w[name] = truncated_normal(rng, shape, math.sqrt(2.0 / fan_in))
        elif leaf == "bias":
            w[name] = (rng.standard_normal(shape) * 0.05).astype(np.float32) if randomize_bn else np.zeros(shape, np.float32)
        elif leaf == "gamma":
            w[name] = rng.uniform(0.5, 1.5, shape).astype(np.float32) if randomize_bn else np.ones(shape, np.float32)
        elif leaf == "beta":
            w[name] = (rng.standard_normal(shape) * 0.1).astype(np.float32) if randomize_bn else np.zeros(shape, np.float32)
        elif leaf == "moving_mean":
            w[name] = (rng.standard_normal(shape) * 0.1).astype(np.float32) if randomize_bn else np.zeros(shape, np.float32)
        elif leaf == "moving_variance":
            w[name] = rng.uniform(0.5, 1.5, shape).astype(np.float32) if randomize_bn else np.ones(shape, np.float32)
    return w


def synthetic_batch(batch: int, seed: int = 20180123):
    """SURVEY 8(d) synthetic AVC pairs: video u8 U{0..255}; audio i16 = noise + one sine per clip;
    label rows [l, 1-l] (data/avc/sample.py:376)."""
    rng = np.random.default_rng(seed)
    video = rng.integers(0, 256, size=(batch, 224, 224, 3), dtype=np.uint8)
    t = np.arange(SR, dtype=np.float64) / SR
    f = rng.uniform(100.0, 8000.0, size=(batch, 1))
    sig = 0.1 * rng.standard_normal((batch, SR)) + 0.2 * np.sin(2 * np.pi * f * t[None, :])
    audio = np.clip(np.round(32767.0 * sig), -32768, 32767).astype(np.int16).reshape(batch, 1, SR)
    lab = rng.integers(0, 2, size=(batch,))
    label = np.stack([lab, 1 - lab], axis=1).astype(np.float32)
    return video, audio, label


# --------------------------------------------------------------------------------------
# Input scaling  (train.py:186,189 ; audio.py:21-31)
# --------------------------------------------------------------------------------------

def pcm2float(sig: np.ndarray, dtype="float32") -> np.ndarray:
    sig = np.asarray(sig)
    if sig.dtype.kind not in "iu":
        raise TypeError("'sig' must be an array of integers")
    dtype = np.dtype(dtype)
    if dtype.kind != "f":
        raise TypeError("'dtype' must be a floating point type")
    i = np.iinfo(sig.dtype)
    abs_max = 2 ** (i.bits - 1)
    offset = i.min + abs_max
    return (sig.astype(dtype) - offset) / abs_max


def scale_video(video_u8: np.ndarray) -> np.ndarray:
    # skimage.img_as_float(uint8) = x / 255 in float64 ; then astype(float32) ; 2*x - 1 in float32
    return 2 * (video_u8.astype(np.float64) / 255.0).astype(np.float32) - 1


# --------------------------------------------------------------------------------------
# kapre front-end (Appendix B)
# --------------------------------------------------------------------------------------

def hann_periodic(n: int) -> np.ndarray:
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)


def mel_filterbank(sr: int, n_fft: int, n_mels: int) -> np.ndarray:
    """librosa 0.5.1 filters.mel(sr, n_fft, n_mels, fmin=0, fmax=sr/2, htk=True, norm=1) -> (n_mels, 1+n_fft/2) float64."""
    fftfreqs = np.linspace(0.0, sr / 2.0, 1 + n_fft // 2)
    mmin, mmax = 0.0, 2595.0 * np.log10(1.0 + (sr / 2.0) / 700.0)
    mels = np.linspace(mmin, mmax, n_mels + 2)
    mel_f = 700.0 * (10.0 ** (mels / 2595.0) - 1.0)
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    w = np.zeros((n_mels, 1 + n_fft // 2))
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        w[i] = np.maximum(0.0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    return w * enorm[:, None]


def frame_geometry(n_samples: int, n_dft: int, n_hop: int, padding: str) -> Tuple[int, int]:
    """TF conv2d SAME/VALID along time -> (n_frames, left_pad)."""
    if padding == "same":
        n_frames = -(-n_samples // n_hop)
        total = max((n_frames - 1) * n_hop + n_dft - n_samples, 0)
        return n_frames, total // 2
    n_frames = (n_samples - n_dft) // n_hop + 1
    return n_frames, 0


def frontend(audio: torch.Tensor, model_type: str, cfg: OracleConfig = OracleConfig()) -> torch.Tensor:
    """audio (B,1,48000) float in [-1,1) -> (B, n_freq|n_mels, n_frames, 1) in cfg.dtype.
    STFT-as-strided-conv == rfft of windowed frames (zero padded, not centred/reflect)."""
    a = AUDIO_SPECS[model_type]
    dt = cfg.dtype
    x = audio.reshape(audio.shape[0], -1).to(dt)
    B, L = x.shape
    n_dft, n_hop = a["n_dft"], a["n_hop"]
    n_frames, left = frame_geometry(L, n_dft, n_hop, a["padding"])
    if cfg.stft_center:
        n_frames, left = 1 + L // n_hop, n_dft // 2
    total = (n_frames - 1) * n_hop + n_dft
    xp = torch.zeros(B, max(total, left + L), dtype=dt)
    xp[:, left:left + L] = x
    frames = xp.unfold(1, n_dft, n_hop)[:, :n_frames]                      # (B, T, n_dft)
    win = torch.from_numpy(hann_periodic(n_dft)).to(dt)
    spec = torch.fft.rfft(frames * win, dim=-1)                            # (B, T, n_freq)
    power = spec.real ** 2 + spec.imag ** 2
    if a["kind"] == "mel":
        fb = torch.from_numpy(mel_filterbank(SR, n_dft, a["n_mels"]).astype(np.float32)).to(dt)  # kapre stores float32
        if cfg.mel_on_magnitude:
            out = torch.sqrt(power) @ fb.t()
        else:
            out = torch.sqrt(power @ fb.t())                               # power_melgram=1.0
    else:
        out = torch.sqrt(power)                                            # power_spectrogram=1.0
    out = out.transpose(1, 2)                                              # (B, F, T)
    if a["decibel"]:
        if cfg.db_amin_on_square:
            log_spec = 0.5 * cfg.db_multiplier * torch.log(torch.clamp(out * out, min=1e-10)) / math.log(10.0)
        else:
            log_spec = cfg.db_multiplier * torch.log(torch.clamp(out, min=1e-10)) / math.log(10.0)
        if cfg.db_ref == "per_sample":
            mx = log_spec.amax(dim=(1, 2), keepdim=True)
        else:
            mx = log_spec.max()
        out = torch.clamp(log_spec - mx, min=-80.0)
    if a["log_lambda"]:
        out = torch.log(torch.clamp(out, min=1e-12)) / 5.0                 # audio_model.py:43
    return out.unsqueeze(-1)


# --------------------------------------------------------------------------------------
# Towers, head, loss
# --------------------------------------------------------------------------------------

def _q(x):
    """round to bf16 in forward, straight-through gradient"""
    return x + (x.detach().to(torch.bfloat16).to(x.dtype) - x.detach())


class _RoundGrad(torch.autograd.Function):
    """identity in forward; rounds the incoming gradient to bf16 in backward (the device stores dz / da as bf16)"""

    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(g.dtype)


def _bn(x, w, prefix, training, cfg, stats):
    """x NCHW.  Keras BatchNormalization(axis=-1 in NHWC): batch mean / biased var when training."""
    g, b = w[prefix + "/gamma"], w[prefix + "/beta"]
    if training:
        mean = x.mean(dim=(0, 2, 3))
        var = x.var(dim=(0, 2, 3), unbiased=False)
        n = x.numel() // x.shape[1]
        stats[prefix] = (mean.detach(), var.detach(), n)
    else:
        mean, var = w[prefix + "/moving_mean"], w[prefix + "/moving_variance"]
    inv = torch.rsqrt(var + cfg.bn_eps)
    return (x - mean[None, :, None, None]) * (inv * g)[None, :, None, None] + b[None, :, None, None]


class _InputBnThroughputMode(torch.autograd.Function):
    """Input BatchNormalization of the bf16 throughput mode, forward AND the device's backward arithmetic.
    forward: y = bf16(scale*x + shift) -- the stored conv input.  backward: the device never materialises the data
    gradient da of the first convolution (csrc/conv_simt.cu k_bn0_from_dw): with xin = the STORED bf16 input it forms
        sum(da)      = sum_{tap,co} w * d1                     (exact, da is not rounded to bf16)
        sum(da*xhat) = (sum_{tap,co} w * dW - beta*sum(da)) / gamma = sum(da * (xin - beta) / gamma)
    i.e. the normalised input is re-derived from the bf16-rounded xin, not from the fp32 x."""

    @staticmethod
    def forward(ctx, x, gamma, beta, mean, inv):
        sh = (1, -1, 1, 1)
        y = (x - mean.view(sh)) * (inv * gamma).view(sh) + beta.view(sh)
        yq = y.to(torch.bfloat16).to(y.dtype)
        ctx.save_for_backward(yq, gamma, beta)
        return yq

    @staticmethod
    def backward(ctx, g):
        yq, gamma, beta = ctx.saved_tensors
        s1 = g.sum(dim=(0, 2, 3))
        s2 = ((g * yq).sum(dim=(0, 2, 3)) - beta * s1) / gamma
        return None, s2, s1, None, None


def _conv(x, w, prefix, cfg=None, round_input_grad=True):
    k = w[prefix + "/kernel"].permute(3, 2, 0, 1)     # HWIO -> OIHW
    if cfg is not None and cfg.emulate_bf16:
        # device: bf16 conv input (gradient da stored as bf16), bf16 weight operands on every tcgen05 layer (the first
        # layer included unless first_layer_bf16_weights is off), fp32 accumulate, output z and its gradient dz
        # stored as bf16
        x = _RoundGrad.apply(_q(x)) if round_input_grad else _q(x)
        if k.shape[1] >= 64 or cfg.first_layer_bf16_weights:
            k = _q(k)
        return _RoundGrad.apply(_q(F.conv2d(x, k, w[prefix + "/bias"], padding=1)))
    return F.conv2d(x, k, w[prefix + "/bias"], padding=1)


def _pool_same(x, ph, pw):
    H, W = x.shape[2], x.shape[3]
    oh, ow = -(-H // ph), -(-W // pw)
    pad_h, pad_w = oh * ph - H, ow * pw - W
    if pad_h or pad_w:
        x = F.pad(x, (pad_w // 2, pad_w - pad_w // 2, pad_h // 2, pad_h - pad_h // 2), value=float("-inf"))
    return F.max_pool2d(x, (ph, pw))


def tower_forward(x_nhwc: torch.Tensor, w: Dict[str, torch.Tensor], tower: str, model_type: str,
                  training: bool, cfg: OracleConfig = OracleConfig(), stats: Optional[dict] = None,
                  return_embedding_map: bool = False, record: Optional[dict] = None):
    """x_nhwc: vision (B,224,224,3) in [-1,1] / audio front-end output (B,F,T,1).
    Returns (B,512) tower output, or the raw conv4b map (B,H,W,512) (embedding tap,
    audio_model.py:482 / vision_model.py:213) when return_embedding_map."""
    if stats is None:
        stats = {}
    spec = (AUDIO_SPECS if tower == "audio" else VISION_SPECS)[model_type]
    x = x_nhwc.permute(0, 3, 1, 2).to(cfg.dtype)
    if spec["input_bn"]:
        x = _bn(x, w, f"{tower}/bn0", training, cfg, stats)
    same_pool = tower == "vision"
    for i, nm in enumerate(CONV_NAMES):
        z = _conv(x, w, f"{tower}/{nm}", cfg)
        if nm == "conv4b" and return_embedding_map:
            return z.permute(0, 2, 3, 1)
        bnn = f"{tower}/bn{nm[4:]}"
        if tower == "vision" and nm == "conv1b":          # vision_model.py:39-43 / :135-139
            x = _bn(F.relu(z), w, bnn, training, cfg, stats)
            if record is not None:
                record.setdefault("relu", {})[i] = (z.detach() > 0).permute(0, 2, 3, 1)
        else:
            x = F.relu(_bn(z, w, bnn, training, cfg, stats))
        if record is not None:                            # the decisions tower_forward_frozen replays
            xa = x.detach().permute(0, 2, 3, 1)
            if nm == "conv4b":
                B_, H_, W_, C_ = xa.shape
                flat = xa.reshape(B_, H_ * W_, C_)
                record["argmax"] = (flat == flat.amax(dim=1, keepdim=True)).to(torch.uint8).argmax(dim=1)
                record["gmask"] = flat.amax(dim=1) > 0
            elif nm in ("conv1b", "conv2b", "conv3b"):
                win = _windows(xa, xa.shape[1] // 2, xa.shape[2] // 2)
                record.setdefault("pos", {})[i] = (win == win.amax(dim=3, keepdim=True)).to(torch.uint8).argmax(dim=3)
                record.setdefault("sign", {})[i] = win.amax(dim=3) > 0
            else:
                record.setdefault("mask", {})[i] = xa > 0
        if nm in ("conv1b", "conv2b", "conv3b"):
            x = _pool_same(x, 2, 2) if same_pool else F.max_pool2d(x, 2, 2)
    if tower == "audio":
        ph, pw = spec["final_pool"]
        x = F.max_pool2d(x, (ph, pw))                      # padding='valid' (keras default)
    else:
        x = _pool_same(x, 28, 28)
    return x.permute(0, 2, 3, 1).reshape(x.shape[0], -1)  # Flatten over (H,W,C)


def _windows(t_nhwc, OH, OW):
    """(B,H,W,C) -> (B,OH,OW,4,C): 2x2 windows in the order (0,0),(0,1),(1,0),(1,1); odd tails dropped ('valid')"""
    B, _, _, C = t_nhwc.shape
    return (t_nhwc[:, :2 * OH, :2 * OW, :].reshape(B, OH, 2, OW, 2, C).permute(0, 1, 3, 2, 4, 5)
            .reshape(B, OH, OW, 4, C))


def tower_forward_frozen(x_nhwc, w, tower, model_type, cfg, routing, stats=None, taps=None):
    """Training-mode tower forward whose DECISIONS (ReLU masks, max-pool routing) are not taken from its own values
    but from `routing`, recorded on the device: arithmetic differences of one bf16 ulp then stay one-ulp differences
    instead of re-routing whole gradient paths, so gradients become comparable tensor by tensor.
      routing["mask"][i]   bool (B,H,W,C)    un-pooled Conv->BN->ReLU layer i: activation > 0
      routing["relu"][i]   bool (B,H,W,C)    Conv->ReLU->BN layer (vision conv1b): z > 0
      routing["pos"][i]    int64 (B,OH,OW,C) pooled layer i: winning window position 0..3 (order of _windows)
      routing["sign"][i]   bool (B,OH,OW,C)  pooled layer i: window maximum > 0
      routing["argmax"]    int64 (B,512)     global max-pool: winning pixel (row-major) ; routing["gmask"] bool: > 0
    Same layer sequence as tower_forward (audio_model.py:370-437, vision_model.py:124-190).  With emulate_bf16 the input
    BN follows the device's throughput-mode arithmetic (_InputBnThroughputMode).  `taps` (optional dict) receives the
    conv outputs z ("<tower>/convNx") and BN outputs ("<tower>/bnNx") so that a caller can ask autograd for the
    gradients AT those tensors (the terms the bias / beta / gamma gradients are sums of)."""
    if stats is None:
        stats = {}
    spec = (AUDIO_SPECS if tower == "audio" else VISION_SPECS)[model_type]
    x = x_nhwc.permute(0, 3, 1, 2).to(cfg.dtype)
    device_bn0 = False
    if spec["input_bn"]:
        if cfg.emulate_bf16:
            mean = x.mean(dim=(0, 2, 3))
            inv = torch.rsqrt(x.var(dim=(0, 2, 3), unbiased=False) + cfg.bn_eps)
            stats[f"{tower}/bn0"] = (mean.detach(), x.var(dim=(0, 2, 3), unbiased=False).detach(), x.numel() // x.shape[1])
            x = _InputBnThroughputMode.apply(x, w[f"{tower}/bn0/gamma"], w[f"{tower}/bn0/beta"], mean, inv)
            device_bn0 = True
        else:
            x = _bn(x, w, f"{tower}/bn0", True, cfg, stats)
        if taps is not None:
            taps[f"{tower}/bn0"] = x
    for i, nm in enumerate(CONV_NAMES):
        z = _conv(x, w, f"{tower}/{nm}", cfg, round_input_grad=not (i == 0 and device_bn0))
        if taps is not None:
            taps[f"{tower}/{nm}"] = z
        bnn = f"{tower}/bn{nm[4:]}"
        relu_first = tower == "vision" and nm == "conv1b"
        if relu_first:
            z = z * routing["relu"][i].permute(0, 3, 1, 2).contiguous().to(z.dtype)
        yb = _bn(z, w, bnn, True, cfg, stats)
        if taps is not None:
            taps[bnn] = yb
        y = yb.permute(0, 2, 3, 1)                                           # NHWC
        if nm == "conv4b":
            B, H, W, C = y.shape
            v = torch.gather(y.reshape(B, H * W, C), 1, routing["argmax"].unsqueeze(1)).squeeze(1)
            return v * routing["gmask"].to(v.dtype)
        if nm in ("conv1b", "conv2b", "conv3b"):
            OH, OW = y.shape[1] // 2, y.shape[2] // 2                        # even sizes for the vision tower ('same')
            y = torch.gather(_windows(y, OH, OW), 3, routing["pos"][i].unsqueeze(3)).squeeze(3)
            if not relu_first:
                y = y * routing["sign"][i].to(y.dtype)
        else:
            y = y * routing["mask"][i].to(y.dtype)
        x = y.permute(0, 3, 1, 2).contiguous()    # same memory layout (hence conv kernel) as tower_forward
    raise AssertionError("unreachable")


def compute_grads_frozen(video_f, audio_f, label, w, model_type, cfg, routing, with_noise_scales=False):
    """compute_grads with the device's routing decisions (routing = {"vision": ..., "audio": ...}).
    with_noise_scales: also returns, for every bias / BN beta / BN gamma gradient, the per-channel root-sum-square of the
    terms it is a sum of (sqrt(sum dz^2), sqrt(sum dy^2), sqrt(sum (dy*xhat)^2)): a sum of N bf16-stored terms cannot be
    reproduced more closely than a fraction of one bf16 ulp of each term, added in quadrature."""
    st: dict = {}
    taps: dict = {}
    v = tower_forward_frozen(video_f, w, "vision", model_type, cfg, routing["vision"], st, taps)
    a = tower_forward_frozen(frontend(audio_f, model_type, cfg), w, "audio", model_type, cfg, routing["audio"], st, taps)
    logits = head_forward(v, a, w)
    loss, ce, acc = avc_loss(logits, label, w, cfg)
    names = [k for k, t in w.items() if t.requires_grad]
    tap_names = [k for k, t in taps.items() if t.requires_grad] if with_noise_scales else []
    grads = torch.autograd.grad(loss, [w[k] for k in names] + [taps[k] for k in tap_names])
    out = dict(loss=loss.detach(), ce=ce.detach(), acc=acc.detach(), logits=logits.detach())
    gd = dict(zip(names, grads[:len(names)]))
    if not with_noise_scales:
        return gd, out
    scales = {}
    for k, g in zip(tap_names, grads[len(names):]):
        g = g.detach()
        rss = torch.sqrt((g * g).sum(dim=(0, 2, 3)))
        if "/conv" in k:
            scales[k + "/bias"] = rss
        else:
            y = taps[k].detach()
            xhat = (y - w[k + "/beta"].detach().view(1, -1, 1, 1)) / w[k + "/gamma"].detach().view(1, -1, 1, 1)
            scales[k + "/beta"] = rss
            scales[k + "/gamma"] = torch.sqrt(((g * xhat) ** 2).sum(dim=(0, 2, 3)))
    return gd, out, scales


def head_forward(v, a, w):
    y = torch.cat([v, a], dim=1)                          # model.py:25 concatenate([vision, audio])
    h = F.relu(y @ w["dense_1/kernel"] + w["dense_1/bias"])
    logits = h @ w["dense_2/kernel"] + w["dense_2/bias"]
    return logits


def to_torch(w_np: Dict[str, np.ndarray], dtype=torch.float32, requires_grad=False) -> Dict[str, torch.Tensor]:
    out = {}
    for k, v in w_np.items():
        t = torch.from_numpy(np.ascontiguousarray(v)).to(dtype)
        if requires_grad and not k.endswith(("moving_mean", "moving_variance")):
            t.requires_grad_(True)
        out[k] = t
    return out


def avc_forward(video_f: torch.Tensor, audio_f: torch.Tensor, w, model_type: str, training: bool,
                cfg: OracleConfig = OracleConfig(), stats: Optional[dict] = None):
    """video_f (B,224,224,3) in [-1,1]; audio_f (B,1,48000) in [-1,1). -> logits (B,2)."""
    spec_a = frontend(audio_f, model_type, cfg)
    v = tower_forward(video_f, w, "vision", model_type, training, cfg, stats)
    a = tower_forward(spec_a, w, "audio", model_type, training, cfg, stats)
    return head_forward(v, a, w)


def avc_loss(logits, label, w, cfg: OracleConfig = OracleConfig()):
    """keras categorical_crossentropy on softmax output (clip 1e-7) + sum of l2(1e-5) kernel penalties."""
    p = torch.softmax(logits, dim=1)
    p = p / p.sum(dim=1, keepdim=True)
    p = torch.clamp(p, 1e-7, 1.0 - 1e-7)
    ce = -(label.to(p.dtype) * torch.log(p)).sum(dim=1).mean()
    reg = sum((w[k] ** 2).sum() for k in w if k.endswith("/kernel")) * WEIGHT_DECAY
    acc = (p.argmax(dim=1) == label.argmax(dim=1)).to(p.dtype).mean()
    return ce + reg, ce, acc


def audio_embedding(audio_f, w, model_type, pooling_type="original", cfg: OracleConfig = OracleConfig()):
    """load_embedding(..., 'audio', pooling_type) path: inference-mode BN, raw conv4b output,
    MaxPooling2D(pool,'same'), Flatten (h,w,c)."""
    spec_a = frontend(audio_f, model_type, cfg)
    z = tower_forward(spec_a, w, "audio", model_type, False, cfg, return_embedding_map=True)
    ph, pw = EMBED_POOL[model_type][pooling_type]
    y = _pool_same(z.permute(0, 3, 1, 2), ph, pw)
    return y.permute(0, 2, 3, 1).reshape(y.shape[0], -1)


def vision_embedding(video_f, w, model_type, cfg: OracleConfig = OracleConfig()):
    z = tower_forward(video_f, w, "vision", model_type, False, cfg, return_embedding_map=True)
    y = _pool_same(z.permute(0, 3, 1, 2), *VISION_EMBED_POOL)
    return y.permute(0, 2, 3, 1).reshape(y.shape[0], -1)


# --------------------------------------------------------------------------------------
# Training step (autograd) with Keras-Adam and per-replica BN data parallelism
# --------------------------------------------------------------------------------------

@dataclass
class AdamState:
    t: int = 0
    m: Dict[str, torch.Tensor] = field(default_factory=dict)
    v: Dict[str, torch.Tensor] = field(default_factory=dict)


def compute_grads(video_f, audio_f, label, w, model_type, cfg=OracleConfig(), n_replicas: int = 1):
    """Forward+backward.  n_replicas>1 emulates multi_gpu_model (training_utils.py:121-170): contiguous
    batch slices, BN statistics per replica, loss = mean over the global batch, L2 term once."""
    B = video_f.shape[0]
    assert B % max(n_replicas, 1) == 0
    step = B // max(n_replicas, 1)
    logits_all, stats_all = [], []
    for r in range(max(n_replicas, 1)):
        sl = slice(r * step, (r + 1) * step)
        st: dict = {}
        logits_all.append(avc_forward(video_f[sl], audio_f[sl], w, model_type, True, cfg, st))
        stats_all.append(st)
    logits = torch.cat(logits_all, dim=0)
    loss, ce, acc = avc_loss(logits, label, w, cfg)
    names = [k for k, t in w.items() if t.requires_grad]
    grads = torch.autograd.grad(loss, [w[k] for k in names])
    return dict(zip(names, grads)), dict(loss=loss.detach(), ce=ce.detach(), acc=acc.detach(), logits=logits.detach()), stats_all


def update_moving_stats(w, stats_all, cfg=OracleConfig()):
    """moving <- moving*momentum + batch*(1-momentum); replicas applied in order (reference order undefined)."""
    with torch.no_grad():
        for st in stats_all:
            for prefix, (mean, var, n) in st.items():
                if cfg.bn_moving_var_unbiased and n > 1:
                    var = var * (n / (n - 1.0))
                mm, mv = w[prefix + "/moving_mean"], w[prefix + "/moving_variance"]
                mm.mul_(cfg.bn_momentum).add_(mean * (1 - cfg.bn_momentum))
                mv.mul_(cfg.bn_momentum).add_(var * (1 - cfg.bn_momentum))


def adam_update(w, grads, state: AdamState, lr: float, cfg=OracleConfig()):
    """Keras 2.0.9 Adam.get_updates."""
    state.t += 1
    t = state.t
    lr_t = lr * math.sqrt(1.0 - cfg.adam_beta2 ** t) / (1.0 - cfg.adam_beta1 ** t)
    with torch.no_grad():
        for k, g in grads.items():
            if k not in state.m:
                state.m[k] = torch.zeros_like(g)
                state.v[k] = torch.zeros_like(g)
            m, v = state.m[k], state.v[k]
            m.mul_(cfg.adam_beta1).add_(g * (1 - cfg.adam_beta1))
            v.mul_(cfg.adam_beta2).add_(g * g * (1 - cfg.adam_beta2))
            w[k].sub_(lr_t * m / (torch.sqrt(v) + cfg.adam_eps))


def train_step(video_u8, audio_i16, label, w, state, model_type, lr, cfg=OracleConfig(), n_replicas=1):
    """One train_on_batch from raw u8/i16 inputs (train.py:186,189 scaling included)."""
    video_f = torch.from_numpy(scale_video(np.asarray(video_u8))).to(cfg.dtype)
    audio_f = torch.from_numpy(pcm2float(np.asarray(audio_i16), "float32")).to(cfg.dtype)
    lab = torch.from_numpy(np.asarray(label))
    grads, out, stats_all = compute_grads(video_f, audio_f, lab, w, model_type, cfg, n_replicas)
    update_moving_stats(w, stats_all, cfg)
    adam_update(w, grads, state, lr, cfg)
    out["grads"] = grads
    return out


# FLOP accounting (2*MACs of conv + dense)
def conv_flops_per_pair(model_type: str) -> Dict[str, float]:
    a = AUDIO_SPECS[model_type]
    nfreq = a["n_mels"] if a["kind"] == "mel" else a["n_dft"] // 2 + 1
    nfr, _ = frame_geometry(SR, a["n_dft"], a["n_hop"], a["padding"])

    def tower(h, w_, cin0, same):
        tot = 0.0
        for nm, (ci, co) in zip(CONV_NAMES, CONV_CHANNELS):
            ci = cin0 if ci is None else ci
            tot += 2.0 * h * w_ * 9 * ci * co
            if nm in ("conv1b", "conv2b", "conv3b"):
                h, w_ = ((h + 1) // 2, (w_ + 1) // 2) if same else (h // 2, w_ // 2)
        return tot
    va, au = tower(224, 224, 3, True), tower(nfreq, nfr, 1, False)
    head = 2.0 * (1024 * 128 + 128 * 2)
    return dict(vision=va, audio=au, head=head, fwd=va + au + head, train=3 * (va + au + head))
