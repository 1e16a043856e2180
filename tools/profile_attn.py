"""Tiny driver for ncu: a few launches of the window-attention forward/backward at one stage shape."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hierarchical_vision_b200 import functional as hvf  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--res", type=int, default=64)
ap.add_argument("--C", type=int, default=96)
ap.add_argument("--heads", type=int, default=3)
ap.add_argument("--ws", type=int, default=8)
ap.add_argument("--shift", type=int, default=4)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--dtype", default="bf16")
ap.add_argument("--tau", type=float, default=10.0)
a = ap.parse_args()
dev = "cuda"
dt = torch.bfloat16 if a.dtype == "bf16" else torch.float32
B, L, C, h, ws = a.batch, a.res * a.res, a.C, a.heads, a.ws
nW = (a.res // ws) ** 2
qkv = torch.randn(B, L, 3 * C, device=dev).to(dt)
out = torch.empty(B, L, C, device=dev, dtype=dt)
lse = hvf.window_attention_stats(qkv, B, a.res, a.res, C, h, ws)
dout = torch.randn(B, L, C, device=dev).to(dt)
dqkv = torch.empty_like(qkv)
tab = 16 * torch.rand((2 * ws - 1) ** 2, h, device=dev)
tau = torch.full((h,), a.tau, device=dev)
dbias, dtau = torch.empty_like(tab), torch.empty_like(tau)
wsp = hvf.window_attention_bwd_workspace(qkv, B, a.res, a.res, C, h, ws)
geom = (B, a.res, a.res, C, h, ws, a.shift)
for _ in range(a.iters):
    hvf.window_attention_fwd_raw(qkv, tab, tau, None, out, lse, *geom)
    hvf.window_attention_bwd_raw(qkv, out, dout, lse, tab, tau, None, dqkv, dbias, dtau, wsp, *geom)
torch.cuda.synchronize()
print("ok", float(out.float().abs().mean()), float(dqkv.float().abs().mean()))
