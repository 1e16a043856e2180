"""GPU: print the parity errors (rel-L2 vs the fp64 oracle) of every attention kernel variant on the core test
cases -- a quick table for tolerance decisions, not a test.

    python tools/attn_err_report.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hierarchical_vision_b200 import functional as hvf  # noqa: E402
from oracle import swin_oracle as O  # noqa: E402
from tests.test_gpu_parity import CORE_CASES, _oracle_core  # noqa: E402
from tests._util import rel_l2  # noqa: E402

DEV = "cuda:0"


def main():
    for case in CORE_CASES:
        B, H, W, C, h, ws, s = case
        if ws != 8 or C // h != 32:
            continue
        g = O.Geometry(B, H, W, C, h, ws, s)
        for taus in ("rand", 100.0):
            gen = torch.Generator().manual_seed(hash(case) % 1000)
            qkv0 = torch.randn(B, H * W, 3 * C, generator=gen).to(DEV, torch.bfloat16)
            tab0 = (16 * torch.rand((2 * ws - 1) ** 2, h, generator=gen)).to(DEV)
            tau0 = (5 + 40 * torch.rand(h, generator=gen)).to(DEV) if taus == "rand" else torch.full((h,), taus, device=DEV)
            do = torch.randn(B, H * W, C, generator=gen).to(DEV, torch.bfloat16)
            ref = _oracle_core(qkv0, tab0, tau0, g, do)
            for fv in (0, 1):
                for bv in (0, 1):
                    hvf.set_attention_forward_variant(fv)
                    hvf.set_attention_backward_variant(bv)
                    qkv, tab, tau = (t.clone().requires_grad_(True) for t in (qkv0, tab0, tau0))
                    out = hvf.window_attention(qkv, tab, tau, B=B, H=H, W=W, C=C, heads=h, ws=ws, shift=s)
                    out.backward(do)
                    torch.cuda.synchronize()
                    o, lse, dqkv, dtab, dtau = ref
                    print(f"{case} tau={taus} fwd={fv} bwd={bv}: out {rel_l2(out, o):.2e} dqkv {rel_l2(qkv.grad, dqkv):.2e} "
                          f"dtab {rel_l2(tab.grad, dtab):.2e} dtau {rel_l2(tau.grad, dtau):.2e}", flush=True)
    hvf.set_attention_forward_variant(-1)
    hvf.set_attention_backward_variant(-1)


if __name__ == "__main__":
    main()
