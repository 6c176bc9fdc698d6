"""Synthetic AVC pairs of the shapes the reference trains on (data/avc/sample.py:373-377: video uint8
(n,224,224,3), audio int16 (n,1,48000), label (n,2) = [l, 1-l]); used by bench.py and smoke()."""
import numpy as np

SR = 48000


def synthetic_batch(batch: int, seed: int = 20180123):
    """video ~ U{0..255}; audio = 0.1*N(0,1) + 0.2*sin(2*pi*f*t), f ~ U(100, 8000) Hz per clip, as int16."""
    rng = np.random.default_rng(seed)
    video = rng.integers(0, 256, size=(batch, 224, 224, 3), dtype=np.uint8)
    t = np.arange(SR, dtype=np.float64) / SR
    f = rng.uniform(100.0, 8000.0, size=(batch, 1))
    sig = 0.1 * rng.standard_normal((batch, SR)) + 0.2 * np.sin(2 * np.pi * f * t[None, :])
    audio = np.clip(np.round(32767.0 * sig), -32768, 32767).astype(np.int16).reshape(batch, 1, SR)
    lab = rng.integers(0, 2, size=(batch,))
    label = np.stack([lab, 1 - lab], axis=1).astype(np.float32)
    return video, audio, label
