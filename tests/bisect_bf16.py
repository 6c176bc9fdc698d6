"""Layer-by-layer comparison of the bf16 device path with the bf16-emulating oracle (developer tool, GPU only)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
import torch.nn.functional as F
from oracle import l3_oracle as O
from l3embedding_b200.engine import Engine

mt = sys.argv[1] if len(sys.argv) > 1 else "cnn_L3_melspec2"
training = (sys.argv[2] if len(sys.argv) > 2 else "train") == "train"
B = 2
w_np = O.init_weights(mt, seed=3, randomize_bn=True)
video, audio, label = O.synthetic_batch(B, seed=505)
cfg = O.OracleConfig(dtype=torch.float32, emulate_bf16=True)
w = O.to_torch(w_np)
eng = Engine(mt, B, "bf16", training=True, weights=w_np)
if training:
    eng.forward_backward(video, audio, label)
else:
    eng.predict(video, audio)
vf = torch.from_numpy(O.scale_video(video)); af = torch.from_numpy(O.pcm2float(audio, "float32"))
for tower, x in (("vision", vf), ("audio", O.frontend(af, mt, cfg))):
    spec = (O.AUDIO_SPECS if tower == "audio" else O.VISION_SPECS)[mt]
    stats = {}
    x0 = eng.debug_read(tower + "/x0", B).reshape(x.shape)
    print(tower, "x0 max|d|", np.abs(x0 - x.numpy()).max())
    x = x.permute(0, 3, 1, 2)
    if spec["input_bn"]:
        x = O._bn(x, w, f"{tower}/bn0", training, cfg, stats)
    xin = eng.debug_read(tower + "/xin", B).reshape(x.permute(0, 2, 3, 1).shape)
    print(tower, "xin max|d| vs bf16(oracle)", np.abs(xin - x.permute(0, 2, 3, 1).to(torch.bfloat16).float().numpy()).max())
    for i, nm in enumerate(O.CONV_NAMES):
        z = O._conv(x, w, f"{tower}/{nm}", cfg)
        zd = eng.debug_read(f"{tower}/z{i}", B).reshape(z.permute(0, 2, 3, 1).shape)
        zo = z.permute(0, 2, 3, 1).detach().numpy()
        diff = np.abs(zd - zo)
        print("%s z%d  max|z| %.3f  max|d| %.4f  mean|d| %.2e  frac(d>0) %.4f" % (tower, i, np.abs(zo).max(), diff.max(), diff.mean(), (diff > 0).mean()))
        bnn = f"{tower}/bn{nm[4:]}"
        if tower == "vision" and nm == "conv1b":
            x = O._bn(F.relu(z), w, bnn, training, cfg, stats)
        else:
            x = F.relu(O._bn(z, w, bnn, training, cfg, stats))
        if nm in ("conv1b", "conv2b", "conv3b"):
            x = O._pool_same(x, 2, 2) if tower == "vision" else F.max_pool2d(x, 2, 2)
        if i < 7:
            ad = eng.debug_read(f"{tower}/a{i}", B).reshape(x.permute(0, 2, 3, 1).shape)
            ao = x.permute(0, 2, 3, 1).to(torch.bfloat16).float().detach().numpy()
            d2 = np.abs(ad - ao)
            print("%s a%d  max|a| %.3f  max|d| %.4f  mean|d| %.2e  frac(d>0) %.4f" % (tower, i, np.abs(ao).max(), d2.max(), d2.mean(), (d2 > 0).mean()))
