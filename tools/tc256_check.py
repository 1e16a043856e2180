"""GPU: parity of the 16 x 16-window tcgen05 attention kernels (wattn_tc256_*) against the fp64 oracle, forward only or
forward + backward -- a development report, not a test.

    python tools/tc256_check.py [--bwd]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hierarchical_vision_b200 import functional as hvf  # noqa: E402
from oracle import swin_oracle as O  # noqa: E402
from tests._util import rel_l2  # noqa: E402

DEV = "cuda:0"
CASES = [
    (1, 16, 16, 32, 1, 16, 0),
    (1, 32, 32, 64, 2, 16, 8),
    (2, 32, 48, 128, 4, 16, 0),
    (3, 16, 32, 64, 2, 16, 8),
    (2, 48, 32, 96, 3, 16, 8),
    (1, 16, 16, 64, 2, 16, 8),
]


def main():
    bwd = "--bwd" in sys.argv
    for case in CASES:
        B, H, W, C, h, ws, s = case
        g = O.Geometry(B, H, W, C, h, ws, s)
        for taus in ("rand", 100.0):
            gen = torch.Generator().manual_seed(hash(case) % 1000)
            qkv = torch.randn(B, H * W, 3 * C, generator=gen).to(DEV, torch.bfloat16).requires_grad_(bwd)
            tab = (16 * torch.rand((2 * ws - 1) ** 2, h, generator=gen)).to(DEV).requires_grad_(bwd)
            tau = ((5 + 40 * torch.rand(h, generator=gen)) if taus == "rand" else torch.full((h,), taus)).to(DEV).requires_grad_(bwd)
            do = torch.randn(B, H * W, C, generator=gen).to(DEV, torch.bfloat16)
            print(case, taus, hvf.window_attention_kernel_name(B, H, W, C, h, ws, s, torch.bfloat16, False), flush=True)
            out = hvf.window_attention(qkv, tab, tau, B=B, H=H, W=W, C=C, heads=h, ws=ws, shift=s)
            torch.cuda.synchronize()
            bias = O.expand_bias(tab.detach().double().cpu(), ws)
            o, lse = O.attention_core_forward(qkv.detach().double().cpu(), bias, tau.detach().double().cpu(), g, use_shift_mask=True)
            msg = f"   out {rel_l2(out, o):.2e}"
            if bwd:
                out.backward(do)
                torch.cuda.synchronize()
                dqkv, dbias, dtau = O.attention_core_backward(qkv.detach().double().cpu(), bias, tau.detach().double().cpu(), g,
                                                              do.double().cpu(), use_shift_mask=True)
                rpi = torch.from_numpy(O.relative_position_index(ws)).reshape(-1)
                dtab = torch.zeros_like(tab.detach().double().cpu())
                dtab.index_add_(0, rpi, dbias.permute(1, 2, 0).reshape(-1, h))
                Cc = C
                gq = qkv.grad.float().cpu()
                msg += (f" dq {rel_l2(gq[..., :Cc], dqkv[..., :Cc]):.2e} dk {rel_l2(gq[..., Cc:2*Cc], dqkv[..., Cc:2*Cc]):.2e}"
                        f" dv {rel_l2(gq[..., 2*Cc:], dqkv[..., 2*Cc:]):.2e} dtab {rel_l2(tab.grad, dtab):.2e} dtau {rel_l2(tau.grad, dtau):.2e}")
            print(msg, flush=True)


if __name__ == "__main__":
    main()
