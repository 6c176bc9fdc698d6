"""Shared helpers of the GPU parity tests (tests/test_gpu_*.py): the CUDA path is always driven through the C ABI
(l3embedding_b200.engine / _lib -> libl3b200.so) and compared with the CPU oracle on the same seeded inputs.

Collection order is alphabetical by file name and deliberate: single kernels first (a_ops, b_tc), then the fp32
parity mode (c_parity), the bf16 throughput mode (d_bf16), training steps (e_training) and last the keras-style API /
file-level tests (f_api) -- with `pytest -x` a failure stops at the most specific test that can explain it.
"""
import os

import numpy as np
import torch

from oracle import l3_oracle as O

MODEL_TYPES = ["cnn_L3_orig", "cnn_L3_kapredbinputbn", "cnn_L3_melspec1", "cnn_L3_melspec2"]
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "oracle_golden.npz")
F64 = O.OracleConfig(dtype=torch.float64)


def engine(*a, **k):
    from l3embedding_b200.engine import Engine
    return Engine(*a, **k)


def rel_l2(a, b):
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def pad(x):  # (B,H,W,C) -> zero-haloed (B,H+2,W+2,C)
    return torch.nn.functional.pad(x, (0, 0, 1, 1, 1, 1))


def oracle_inputs(video, audio):
    return (torch.from_numpy(O.scale_video(video)).double(), torch.from_numpy(O.pcm2float(audio, "float64")))
