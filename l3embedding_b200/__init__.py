"""l3embedding_b200: B200-native (sm_100a) L3-Net audio-visual-correspondence training / embedding path.

Drop-in for the hot path of marl/l3embedding (model builders, train_on_batch, predict, embedding extraction).
All arithmetic runs in libl3b200.so (hand-written CUDA); there is no CPU fallback.
"""
from ._lib import L3Error  # noqa: F401

__all__ = ["L3Error"]
