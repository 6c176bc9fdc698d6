"""Generates tests/golden/oracle_golden.npz from the fp64 oracle on seeded synthetic inputs.

    python tests/golden/make_golden.py

The reference (keras 2.0.9 / TF 1.4 / kapre) cannot be imported in this environment, so these vectors pin the
ORACLE (regression guard + GPU-box fixture), not the reference: parity with keras itself stays unpinned.
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from oracle import l3_oracle as O  # noqa: E402

META = dict(model_types=["cnn_L3_orig", "cnn_L3_kapredbinputbn", "cnn_L3_melspec1", "cnn_L3_melspec2"],
            embedding_types=["cnn_L3_melspec2", "cnn_L3_orig"], batch=2, data_seed=77, weight_seed=20180123,
            stride_f=5, stride_t=7)


def main():
    out = {"meta": np.array(json.dumps(META))}
    cfg = O.OracleConfig(dtype=torch.float64)
    video, audio, label = O.synthetic_batch(META["batch"], seed=META["data_seed"])
    af = torch.from_numpy(O.pcm2float(audio, "float64"))
    for mt in META["model_types"]:
        fe = O.frontend(af, mt, cfg)[..., 0].numpy()
        out[mt + "/frontend"] = fe[:, ::META["stride_f"], ::META["stride_t"]]
        if mt in META["embedding_types"]:
            w = O.to_torch(O.init_weights(mt, seed=META["weight_seed"], randomize_bn=True), dtype=torch.float64)
            out[mt + "/embedding_short"] = O.audio_embedding(af, w, mt, "short", cfg).numpy()
            out[mt + "/embedding_original"] = O.audio_embedding(af, w, mt, "original", cfg).numpy().astype(np.float32)
    np.savez_compressed(os.path.join(os.path.dirname(__file__), "oracle_golden.npz"), **out)


if __name__ == "__main__":
    main()
