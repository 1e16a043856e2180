"""GPU: the fused dGELU GEMM (hv_mlp_dgelu_gemm) against torch on the same bf16 inputs, and its time next to the two-kernel
path it replaces (cuBLAS dgrad GEMM + hv_bias_gelu_bwd) -- a development report, not a test.

    python tools/mlp_gemm_check.py [--batch 256]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hierarchical_vision_b200 import _lib  # noqa: E402
from hierarchical_vision_b200 import functional as hvf  # noqa: E402
from hierarchical_vision_b200.functional import _ptr, _stream, check  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--shapes", default="64:96,32:192,16:384,8:768", help="token-grid side : C per stage (SwinV2-B: 64:128,32:256,16:512)")
a = ap.parse_args()
dev = "cuda"
lib = _lib.load()


def fused(dy, w2, h, b1):
    M, C = dy.shape
    N = w2.shape[1]
    dh = torch.empty((M, N), dtype=torch.bfloat16, device=dev)
    db1 = torch.empty((N,), dtype=torch.float32, device=dev)
    nb = int(lib.hv_mlp_dgelu_gemm_workspace_bytes(M, N, C))
    assert nb > 0
    ws = torch.empty((nb,), dtype=torch.uint8, device=dev)
    check(lib.hv_mlp_dgelu_gemm(_ptr(dy), _ptr(w2), _ptr(h), _ptr(b1), _ptr(dh), _ptr(db1), _ptr(ws), nb, M, N, C, 1,
                                _stream(dy.device)), "hv_mlp_dgelu_gemm")
    return dh, db1


def fused_fwd(x, w1, b1):
    M, C = x.shape
    N = w1.shape[0]
    hh = torch.empty((M, N), dtype=torch.bfloat16, device=dev)
    act = torch.empty((M, N), dtype=torch.bfloat16, device=dev)
    check(lib.hv_mlp_fc1_gelu_gemm(_ptr(x), _ptr(w1), _ptr(b1), _ptr(hh), _ptr(act), M, N, C, 1, _stream(x.device)),
          "hv_mlp_fc1_gelu_gemm")
    return hh, act


def timeit(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


for res, C in [tuple(int(v) for v in item.split(":")) for item in a.shapes.split(",")]:
    M, N = a.batch * res * res, 4 * C
    g = torch.Generator(device=dev).manual_seed(res)
    dy = torch.randn(M, C, device=dev, generator=g).to(torch.bfloat16)
    w2 = (torch.randn(C, N, device=dev, generator=g) / C ** 0.5).to(torch.bfloat16)
    h = (2 * torch.randn(M, N, device=dev, generator=g)).to(torch.bfloat16)
    b1 = 0.5 * torch.randn(N, device=dev, generator=g)
    dh, db1 = fused(dy, w2, h, b1)
    torch.cuda.synchronize()
    # reference on a slice of rows (fp32 math on the same bf16 inputs)
    rows = slice(0, 4096)
    x = (h[rows].float() + b1).requires_grad_(True)
    torch.nn.functional.gelu(x).backward(dy[rows].float() @ w2.float())
    ref = x.grad
    err = ((dh[rows].float() - ref).norm() / ref.norm()).item()
    # column sums against the kernel's own dh (fp32 sums of the unrounded values vs bf16-rounded rows: ~1e-3)
    cs = dh.float().sum(0)
    err_b = ((db1 - cs).norm() / cs.norm()).item()
    t_f = timeit(lambda: fused(dy, w2, h, b1), a.iters)

    def two_kernel():
        da = dy @ w2
        out = torch.empty_like(h)
        dbias = torch.empty_like(b1)
        nb = lib.hv_bias_gelu_bwd_workspace_bytes(M, N)
        ws = torch.empty((int(nb),), dtype=torch.uint8, device=dev)
        check(lib.hv_bias_gelu_bwd(_ptr(da), _ptr(h), _ptr(b1), _ptr(out), _ptr(dbias), _ptr(ws), ws.numel(), M, N, 1,
                                   _stream(h.device)), "hv_bias_gelu_bwd")
        return out, dbias

    o2, d2 = two_kernel()
    err2 = ((dh.float() - o2.float()).norm() / o2.float().norm()).item()
    t_2 = timeit(two_kernel, a.iters)
    # ---- forward: h = x W1^T, a = GELU(h + b1)
    x = torch.randn(M, C, device=dev, generator=g).to(torch.bfloat16)
    w1 = (torch.randn(N, C, device=dev, generator=g) / C ** 0.5).to(torch.bfloat16)
    hh, act = fused_fwd(x, w1, b1)
    torch.cuda.synchronize()
    h_ref = x[rows].float() @ w1.float().t()
    a_ref = torch.nn.functional.gelu(h_ref + b1)
    eh = ((hh[rows].float() - h_ref).norm() / h_ref.norm()).item()
    ea = ((act[rows].float() - a_ref).norm() / a_ref.norm()).item()
    t_ff = timeit(lambda: fused_fwd(x, w1, b1), a.iters)

    def two_kernel_fwd():
        h2 = x @ w1.t()
        return h2, hvf.bias_gelu(h2, b1)

    t_f2 = timeit(two_kernel_fwd, a.iters)
    print(f"C {C:4d} forward: h {eh:.2e} a {ea:.2e} | fused {t_ff:.3f} ms ({(M * N * 4 + M * C * 2) / t_ff / 1e6:.0f} GB/s) vs GEMM + bias_gelu "
          f"{t_f2:.3f} ms", flush=True)
    nbytes = M * N * 2 * 2 + M * C * 2
    print(f"C {C:4d} M {M:8d}: dh rel-L2 vs torch {err:.2e}, vs two-kernel path {err2:.2e}, db1 {err_b:.2e} | fused {t_f:.3f} ms "
          f"({nbytes / t_f / 1e6:.0f} GB/s) vs GEMM + bias_gelu_bwd {t_2:.3f} ms", flush=True)
