"""Front-end only: all four model types, int16 and float input, against the fp64 oracle (max |dB| error) -- small enough
to run under compute-sanitizer.   python tools/fe_check.py [batch]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
from l3embedding_b200.engine import Engine
from oracle import l3_oracle as O   # checker

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
_, audio, _ = O.synthetic_batch(B, seed=5)
audio[0, :, 30000:] = 0            # half-silent clip
F64 = O.OracleConfig(dtype=torch.float64)
for mt in ("cnn_L3_orig", "cnn_L3_kapredbinputbn", "cnn_L3_melspec1", "cnn_L3_melspec2"):
    eng = Engine(mt, B, "f32", training=False, towers=("audio",), host_staging=False)
    ref = O.frontend(torch.from_numpy(O.pcm2float(audio, "float64")), mt, F64)[..., 0].numpy()
    got_i = eng.frontend(audio).cpu().numpy()
    got_f = eng.frontend(O.pcm2float(audio, "float32")).cpu().numpy()
    print(mt, "max|d| i16 %.3g f32 %.3g  i16==f32 %s" % (np.abs(got_i - ref).max(), np.abs(got_f - ref).max(),
                                                        np.array_equal(got_i, got_f)))
    eng.close()
