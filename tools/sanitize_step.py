"""One small training step + one prediction per model type and numeric mode -- sized for `compute-sanitizer --tool memcheck`.
    compute-sanitizer --tool memcheck python tools/sanitize_step.py [batch]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from l3embedding_b200.engine import Engine
from l3embedding_b200.synthetic import synthetic_batch

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
video, audio, label = synthetic_batch(B, seed=3)
for mt in ("cnn_L3_melspec2", "cnn_L3_kapredbinputbn", "cnn_L3_orig"):
    for dtype in ("bf16", "f32tc"):
        eng = Engine(mt, B, dtype, training=True)
        m = eng.train_step_host(video, audio, label, 1e-5)
        p, _ = eng.predict(video, audio)
        e = eng.embed_audio(audio, "short").cpu().numpy()
        print(mt, dtype, "loss %.4f" % m["loss"], "probs", np.round(p[0], 3), "emb", e.shape, flush=True)
        eng.close()
