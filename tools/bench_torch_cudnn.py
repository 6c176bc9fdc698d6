"""Context number ("secondary"): the same cnn_L3_melspec2 training step in stock PyTorch + cuDNN on the same
GPU -- bf16 autocast, channels_last, fused Adam -- i.e. what the off-the-shelf stack gives on a B200.  The audio
front-end is NOT included (the tower starts from a random (B,1,256,199) mel map), which favours this baseline by
~0.2 ms per step.  Independent of oracle/ and of the library: plain torch.nn.

    python tools/bench_torch_cudnn.py [--batch 64] [--steps 20]
"""
import argparse
import json

import torch
import torch.nn as nn
import torch.nn.functional as F


def tower(c_in, same_pool):
    layers, c = [nn.BatchNorm2d(c_in)], c_in
    for i, co in enumerate([64, 64, 128, 128, 256, 256, 512, 512]):
        layers += [nn.Conv2d(c, co, 3, padding=1), nn.BatchNorm2d(co), nn.ReLU(inplace=True)]
        if i in (1, 3, 5):
            layers.append(nn.MaxPool2d(2, 2, ceil_mode=same_pool))
        c = co
    layers.append(nn.AdaptiveMaxPool2d(1))
    return nn.Sequential(*layers)


class L3(nn.Module):
    def __init__(self):
        super().__init__()
        self.vision, self.audio = tower(3, True), tower(1, False)
        self.fc1, self.fc2 = nn.Linear(1024, 128), nn.Linear(128, 2)

    def forward(self, v, a):
        x = torch.cat([self.vision(v).flatten(1), self.audio(a).flatten(1)], 1)
        return self.fc2(F.relu(self.fc1(x)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    a = ap.parse_args()
    torch.backends.cudnn.benchmark = True
    dev = "cuda"
    m = L3().to(dev).to(memory_format=torch.channels_last)
    opt = torch.optim.Adam(m.parameters(), lr=1e-5, fused=True)
    B = a.batch
    pool = [(torch.randn(B, 3, 224, 224, device=dev).contiguous(memory_format=torch.channels_last),
             torch.randn(B, 1, 256, 199, device=dev).contiguous(memory_format=torch.channels_last),
             torch.randint(0, 2, (B,), device=dev)) for _ in range(4)]

    def step(i):
        v, au, y = pool[i % 4]
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = F.cross_entropy(m(v, au).float(), y)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()

    for i in range(a.warmup):
        step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.steps):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    print(json.dumps({"impl": "torch %s + cuDNN %s, bf16 autocast, channels_last, eager" % (torch.__version__, torch.backends.cudnn.version()),
                      "workload": "cnn_L3_melspec2 train step without the audio front-end", "batch": B,
                      "ms_per_step": ms, "pairs_per_s": B * 1e3 / ms}))


if __name__ == "__main__":
    main()
