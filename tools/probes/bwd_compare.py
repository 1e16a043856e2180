"""Compare the tcgen05 backward against the mma.sync backward on the same inputs (GPU). Prints rel-L2 per output."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from hierarchical_vision_b200 import functional as hvf

def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))

def run(B, H, W, C, h, shift, seed=0):
    ws = 8
    torch.manual_seed(seed)
    dev = "cuda"
    qkv = torch.randn(B, H * W, 3 * C, device=dev).to(torch.bfloat16)
    tab = 16 * torch.rand(225, h, device=dev)
    tau = 5 + 40 * torch.rand(h, device=dev)
    do = torch.randn(B, H * W, C, device=dev).to(torch.bfloat16)
    nW = (H // ws) * (W // ws)
    out = torch.empty(B, H * W, C, device=dev, dtype=torch.bfloat16)
    lse = torch.empty(B * nW, h, 64, device=dev)
    hvf.set_attention_forward_variant(0)
    hvf.window_attention_fwd_raw(qkv, tab, tau, None, out, lse, B, H, W, C, h, ws, shift)
    wsp = hvf.window_attention_bwd_workspace(qkv, B, H, W, C, h, ws)
    res = []
    for v in (0, 1):
        hvf.set_attention_backward_variant(v)
        dqkv = torch.full_like(qkv, float("nan"))
        dbias, dtau, col = torch.empty_like(tab), torch.empty_like(tau), torch.empty(C, device=dev)
        hvf.window_attention_bwd_raw(qkv, out, do, lse, tab, tau, None, dqkv, dbias, dtau, wsp, B, H, W, C, h, ws, shift, dq_colsum=col)
        torch.cuda.synchronize()
        res.append((dqkv, dbias, dtau, col))
    a, b = res[1], res[0]
    print(f"B{B} {H}x{W} C{C} h{h} shift{shift}: dq {rel(a[0][..., :C], b[0][..., :C]):.3e} dk {rel(a[0][..., C:2*C], b[0][..., C:2*C]):.3e} "
          f"dv {rel(a[0][..., 2*C:], b[0][..., 2*C:]):.3e} dbias {rel(a[1], b[1]):.3e} dtau {rel(a[2], b[2]):.3e} col {rel(a[3], b[3]):.3e} "
          f"nan {int(torch.isnan(a[0].float()).sum())}", flush=True)
    if os.environ.get("HV_DUMP"):
        for hh in range(h):
            sl = slice(hh * 32, hh * 32 + 32)
            print("  head", hh, "dq", f"{rel(a[0][..., :C][..., sl], b[0][..., :C][..., sl]):.2e}",
                  "dk", f"{rel(a[0][..., C:2*C][..., sl], b[0][..., C:2*C][..., sl]):.2e}",
                  "dv", f"{rel(a[0][..., 2*C:][..., sl], b[0][..., 2*C:][..., sl]):.2e}",
                  "dtau", float(a[2][hh]), float(b[2][hh]))

if __name__ == "__main__":
    run(1, 8, 8, 64, 2, 0)
    run(2, 16, 16, 64, 2, 0)
    run(2, 16, 16, 96, 3, 0)
    run(2, 16, 16, 96, 3, 4)
    run(3, 16, 24, 192, 6, 4)
    run(4, 8, 8, 32, 1, 0)
    run(8, 64, 64, 96, 3, 4)
