"""Kernel-level microbenchmark (GPU): achieved HBM GB/s of each hand-written kernel at the SwinV2-T stage
shapes, CUDA-event timed on the launching stream, inputs rotated through buffers larger than L2.

    python tools/bench_kernels.py [--batch 128] [--iters 30] [--json out.json]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hierarchical_vision_b200 import functional as hvf  # noqa: E402

PEAK = 6531.6
if os.path.exists("MEASURED_PEAKS.json"):
    PEAK = json.load(open("MEASURED_PEAKS.json")).get("hbm_gbs", PEAK)
L2_BYTES = 126 * 2 ** 20


def timeit(fns, iters, warmup=5):
    """fns: list of closures (one per rotated buffer set). Returns mean ms per call."""
    for i in range(warmup):
        fns[i % len(fns)]()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fns[i % len(fns)]()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def bench_attn(B, res, C, heads, ws, shift, dtype, iters, tau_mode="init", colsum=False):
    dev = "cuda"
    L = res * res
    nW = (res // ws) ** 2
    esz = 2 if dtype == torch.bfloat16 else 4
    per_set = B * L * C * esz * 12  # q,k,v,o,do + dq,dk,dv
    nsets = max(2, int(2 * L2_BYTES / per_set) + 1)
    nsets = min(nsets, 6)
    M = (2 * ws - 1) ** 2
    tab = (16 * torch.rand(M, heads, device=dev)).float()
    # init: exp(logit_scale) of a freshly constructed model (log 10, swinv2.py:135-137); clamp: every head at the
    # clamp of 100 (swinv2.py:230), the other softmax path of the tensor-core kernel
    tau = {"init": torch.full((heads,), 10.0, device=dev), "clamp": torch.full((heads,), 100.0, device=dev),
           "mixed": (5 + 20 * torch.rand(heads, device=dev)).float()}[tau_mode]
    sets = []
    for _ in range(nsets):
        qkv = torch.randn(B, L, 3 * C, device=dev).to(dtype)
        out = torch.empty(B, L, C, device=dev, dtype=dtype)
        lse = hvf.window_attention_stats(qkv, B, res, res, C, heads, ws)
        dout = torch.randn(B, L, C, device=dev).to(dtype)
        dqkv = torch.empty_like(qkv)
        sets.append((qkv, out, lse, dout, dqkv))
    dbias = torch.empty_like(tab)
    dtau = torch.empty_like(tau)
    wsp = hvf.window_attention_bwd_workspace(sets[0][0], B, res, res, C, heads, ws)
    geom = (B, res, res, C, heads, ws, shift)
    fw = [lambda s=s: hvf.window_attention_fwd_raw(s[0], tab, tau, None, s[1], s[2], *geom) for s in sets]
    for f in fw:
        f()
    cs = torch.empty(C, device=dev) if colsum else None  # the training step asks for d(q_bias) = column sums of dq
    bw = [lambda s=s: hvf.window_attention_bwd_raw(s[0], s[1], s[3], s[2], tab, tau, None, s[4], dbias, dtau, wsp, *geom, dq_colsum=cs)
          for s in sets]
    t_f = timeit(fw, iters)
    t_b = timeit(bw, iters)
    windows = B * nW
    bytes_f = windows * 4 * ws * ws * C * esz
    bytes_b = windows * 8 * ws * ws * C * esz
    return dict(kernel="window_attn", B=B, res=res, C=C, heads=heads, ws=ws, shift=shift, dtype=str(dtype).split(".")[-1], tau=tau_mode,
                windows=windows, fwd_ms=t_f, bwd_ms=t_b, fwd_gbs=bytes_f / t_f / 1e6, bwd_gbs=bytes_b / t_b / 1e6,
                fwdbwd_gbs=(bytes_f + bytes_b) / (t_f + t_b) / 1e6, windows_per_s=windows / (t_f + t_b) * 1e3,
                frac_fwd=bytes_f / t_f / 1e6 / PEAK, frac_bwd=bytes_b / t_b / 1e6 / PEAK,
                frac_fwdbwd=(bytes_f + bytes_b) / (t_f + t_b) / 1e6 / PEAK)


def bench_gelu(rows, cols, dtype, iters):
    """bias + GELU streaming kernels (Mlp activation): fwd reads h, writes a; bwd reads dout, h, writes dh."""
    dev = "cuda"
    esz = 2 if dtype == torch.bfloat16 else 4
    per_set = rows * cols * esz * 3
    nsets = min(6, max(2, int(2 * L2_BYTES / per_set) + 1))
    lib = hvf._lib.load()
    P = hvf._ptr
    st = hvf._stream(torch.device("cuda", torch.cuda.current_device()))
    bias = torch.randn(cols, device=dev)
    wbytes = lib.hv_bias_gelu_bwd_workspace_bytes(rows, cols)
    wsp = torch.empty(wbytes, dtype=torch.uint8, device=dev)
    db = torch.empty(cols, device=dev)
    sets = [(torch.randn(rows, cols, device=dev).to(dtype), torch.empty(rows, cols, device=dev, dtype=dtype),
             torch.randn(rows, cols, device=dev).to(dtype)) for _ in range(nsets)]
    code = hvf._code(sets[0][0])

    def f(s):
        rc = lib.hv_bias_gelu_fwd(P(s[0]), P(bias), P(s[1]), rows, cols, code, st)
        assert rc == 0, lib.hv_last_error()

    def b(s):
        rc = lib.hv_bias_gelu_bwd(P(s[2]), P(s[0]), P(bias), P(s[1]), P(db), P(wsp), wbytes, rows, cols, code, st)
        assert rc == 0, lib.hv_last_error()

    t_f = timeit([lambda s=s: f(s) for s in sets], iters)
    t_b = timeit([lambda s=s: b(s) for s in sets], iters)
    bf, bb = rows * cols * esz * 2, rows * cols * esz * 3
    return dict(kernel="bias_gelu", rows=rows, cols=cols, dtype=str(dtype).split(".")[-1], fwd_ms=t_f, bwd_ms=t_b,
                fwd_gbs=bf / t_f / 1e6, bwd_gbs=bb / t_b / 1e6, frac_fwd=bf / t_f / 1e6 / PEAK, frac_bwd=bb / t_b / 1e6 / PEAK)


def bench_ln(B, L, C, ydt, rdt, iters):
    dev = "cuda"
    ey = 2 if ydt == torch.bfloat16 else 4
    er = 2 if rdt == torch.bfloat16 else 4
    per_set = B * L * C * (ey + 2 * er)
    nsets = min(6, max(2, int(2 * L2_BYTES / per_set) + 1))
    gam = torch.ones(C, device=dev)
    bet = torch.zeros(C, device=dev)
    lib = hvf._lib.load()
    sets = []
    for _ in range(nsets):
        y = torch.randn(B, L, C, device=dev).to(ydt)
        sc = torch.randn(B, L, C, device=dev).to(rdt)
        out = torch.empty_like(sc)
        mean = torch.empty(B * L, device=dev)
        rstd = torch.empty(B * L, device=dev)
        dy = torch.empty_like(y)
        sets.append((y, sc, out, mean, rstd, dy))
    rows = B * L
    P = hvf._ptr
    st = hvf._stream(torch.device("cuda", torch.cuda.current_device()))
    wbytes = lib.hv_ln_residual_bwd_workspace_bytes(rows, C)
    wsp = torch.empty(wbytes, dtype=torch.uint8, device=dev)
    dg, db, dbi = torch.empty(C, device=dev), torch.empty(C, device=dev), torch.empty(C, device=dev)
    cy, cr = hvf._code(sets[0][0]), hvf._code(sets[0][1])

    def f(s):
        rc = lib.hv_ln_residual_fwd(P(s[0]), P(s[1]), P(gam), P(bet), P(gam), P(None), P(s[2]), P(s[3]), P(s[4]), rows, C, L, 1e-5, cy, cr, st)
        assert rc == 0, lib.hv_last_error()

    def b(s):
        rc = lib.hv_ln_residual_bwd(P(s[2]), P(s[0]), P(gam), P(gam), P(s[3]), P(s[4]), P(None), P(s[5]), P(dg), P(db), P(dbi), P(wsp), wbytes,
                                    rows, C, L, cy, cr, st)
        assert rc == 0, lib.hv_last_error()

    fw = [lambda s=s: f(s) for s in sets]
    bw = [lambda s=s: b(s) for s in sets]
    t_f, t_b = timeit(fw, iters), timeit(bw, iters)
    bf = rows * C * (ey + 2 * er)
    bb = rows * C * (er + 2 * ey)
    return dict(kernel="ln_residual", rows=rows, C=C, y=str(ydt).split(".")[-1], res=str(rdt).split(".")[-1], fwd_ms=t_f,
                bwd_ms=t_b, fwd_gbs=bf / t_f / 1e6, bwd_gbs=bb / t_b / 1e6, frac_fwd=bf / t_f / 1e6 / PEAK,
                frac_bwd=bb / t_b / 1e6 / PEAK)


def bench_merge(B, res, C, dtype, iters):
    dev = "cuda"
    esz = 2 if dtype == torch.bfloat16 else 4
    per_set = B * res * res * C * esz * 2
    nsets = min(6, max(2, int(2 * L2_BYTES / per_set) + 1))
    lib = hvf._lib.load()
    P = hvf._ptr
    st = hvf._stream(torch.device("cuda", torch.cuda.current_device()))
    sets = [(torch.randn(B, res * res, C, device=dev).to(dtype), torch.empty(B, res * res // 4, 4 * C, device=dev, dtype=dtype))
            for _ in range(nsets)]
    code = hvf._code(sets[0][0])
    fw = [lambda s=s: lib.hv_patch_merge_gather_fwd(P(s[0]), P(s[1]), B, res, res, C, code, st) for s in sets]
    bw = [lambda s=s: lib.hv_patch_merge_gather_bwd(P(s[1]), P(s[0]), B, res, res, C, code, st) for s in sets]
    t_f, t_b = timeit(fw, iters), timeit(bw, iters)
    return dict(kernel="patch_merge_gather", B=B, res=res, C=C, dtype=str(dtype).split(".")[-1], fwd_ms=t_f, bwd_ms=t_b,
                fwd_gbs=per_set / t_f / 1e6, bwd_gbs=per_set / t_b / 1e6, frac_fwd=per_set / t_f / 1e6 / PEAK,
                frac_bwd=per_set / t_b / 1e6 / PEAK)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--json", default="")
    ap.add_argument("--only", default="")
    ap.add_argument("--tau", default="init", choices=["init", "clamp", "mixed"])
    ap.add_argument("--sweep", action="store_true", help="BASELINE.json configs[4]: shifted window attention at the stage-0 and "
                    "stage-1 shapes + PatchMerging's gather over batch 32 ... 512")
    ap.add_argument("--generic16", action="store_true", help="attn16: also time the generic CUDA-core kernels (batch / 8)")
    ap.add_argument("--colsum", action="store_true", help="attention backward also produces d(q_bias) (as in the training step)")
    a = ap.parse_args()
    rows = []
    B = a.batch
    stages = [(64, 96, 3), (32, 192, 6), (16, 384, 12), (8, 768, 24)]
    if a.sweep:
        for b in (32, 64, 128, 256, 512):
            for res, C, h in stages[:2]:
                for shift in (0, 4):
                    rows.append(bench_attn(b, res, C, h, 8, shift, torch.bfloat16, a.iters, a.tau))
                    print(json.dumps(rows[-1]), flush=True)
                rows.append(dict(bench_merge(b, res, C, torch.bfloat16, a.iters), B=b))
                print(json.dumps(rows[-1]), flush=True)
        if a.json:
            with open(a.json, "w") as f:
                json.dump(rows, f, indent=1)
        return
    if a.only in ("", "attn", "attn0"):
        for res, C, h in stages:
            for shift in ((0, 4) if res > 8 and a.only != "attn0" else (0,)):
                rows.append(bench_attn(B, res, C, h, 8, shift, torch.bfloat16, a.iters, a.tau, a.colsum))
                print(json.dumps(rows[-1]), flush=True)
        if a.only != "attn0":
            rows.append(bench_attn(max(B // 8, 8), 64, 96, 3, 8, 4, torch.float32, max(a.iters // 6, 3)))
            print(json.dumps(rows[-1]), flush=True)
    if a.only in ("", "attn16"):
        # SwinV2-B at window 16 (BASELINE configs[3]): stages 0-2 run 16 x 16 windows (N = 256), stage 3 is 8 x 8 tokens
        # (window clamped to 8: the N = 64 kernels).  tcgen05 kernels, then the generic CUDA-core kernels for comparison.
        for res, C, h in ((64, 128, 4), (32, 256, 8), (16, 512, 16)):
            for shift in ((0, 8) if res > 16 else (0,)):
                for variant in ((1, 0) if a.generic16 else (1,)):
                    hvf.set_attention_tc256_variant(variant)
                    r = bench_attn(B if variant else max(B // 8, 4), res, C, h, 16, shift, torch.bfloat16,
                                   a.iters if variant else max(a.iters // 6, 3), a.tau)
                    r["kernel"] = "window_attn16_tcgen05" if variant else "window_attn16_generic"
                    rows.append(r)
                    print(json.dumps(rows[-1]), flush=True)
        hvf.set_attention_tc256_variant(-1)
    if a.only in ("", "gelu"):
        for res, C, h in stages:
            rows.append(bench_gelu(B * res * res, 4 * C, torch.bfloat16, a.iters))
            print(json.dumps(rows[-1]), flush=True)
    if a.only in ("", "ln"):
        for res, C, h in stages:
            for ydt, rdt in ((torch.bfloat16, torch.bfloat16), (torch.bfloat16, torch.float32), (torch.float32, torch.float32)):
                rows.append(bench_ln(B, res * res, C, ydt, rdt, a.iters))
                print(json.dumps(rows[-1]), flush=True)
    if a.only in ("", "merge"):
        for res, C, h in stages[:3]:
            rows.append(bench_merge(B, res, C, torch.bfloat16, a.iters))
            print(json.dumps(rows[-1]), flush=True)
    if a.json:
        with open(a.json, "w") as f:
            json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
