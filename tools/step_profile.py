"""Kernel-time breakdown of one training step via torch.profiler (CUPTI), far cheaper than an ncu launch list.
    python tools/step_profile.py [--config swinv2_b] [--batch 256] --steps 3"""
import argparse
import os
import re
import sys
from collections import defaultdict

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from hierarchical_vision_b200 import train as T  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="swinv2_t", choices=sorted(bench.CONFIGS))
ap.add_argument("--batch", type=int, default=0)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--top", type=int, default=45)
a = ap.parse_args()
dev = torch.device("cuda", 0)
cfg = bench.CONFIGS[a.config]
a.batch = a.batch or cfg["batch"]
model = bench.build_model(dev, config=a.config)
opt = T.build_optimizer(model, lr=0.05)
img = torch.randint(0, 256, (a.batch, 3, 256, 256), dtype=torch.uint8, device=dev)
lab = bench.synth_labels(cfg["classes"], a.batch, torch.Generator().manual_seed(0)).to(dev)
step = lambda: T.train_step(model, opt, (img, lab), autocast_dtype=torch.bfloat16, clip_norm=2.0)
for _ in range(3):
    step()
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile

with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(a.steps):
        step()
    torch.cuda.synchronize()
agg = defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        n = ev.name.replace("void ", "").replace("(anonymous namespace)::", "").replace("at::native::", "")
        n = re.sub(r"\(.*", "", n)[:100]
        agg[n][0] += 1
        agg[n][1] += ev.device_time
tot = sum(v[1] for v in agg.values())
print(f"total CUDA time per step: {tot / a.steps / 1e3:.3f} ms over {sum(v[0] for v in agg.values()) // a.steps} kernels")
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[: a.top]:
    print(f"{t / a.steps / 1e3:8.3f} ms {100 * t / tot:5.1f}%  x{c // a.steps:4d}  {n}")
