"""torch.autograd.Function wrappers over the C ABI (include/hv_swin.h).

PyTorch is plumbing here: it owns device memory, streams and the autograd graph; all the
arithmetic of these ops happens in libhv_swin.so.  Tensors are passed as raw device pointers
plus sizes; nothing in the signatures of the library is a torch type.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib
from ._lib import HV_BF16, HV_F32, HV_U8, check

_P = _lib.c_void_p


def _code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return HV_F32
    if t.dtype == torch.bfloat16:
        return HV_BF16
    raise RuntimeError(f"hierarchical_vision_b200 kernels take float32 or bfloat16 activations, got {t.dtype}")


def _ptr(t: Optional[torch.Tensor]):
    return _P(t.data_ptr()) if t is not None else _P(0)


def _need_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{what}: tensor is on {t.device}; this package only runs on CUDA (sm_100a), "
                           "there is no CPU path")


def _stream(device) -> _P:
    return _P(torch.cuda.current_stream(device).cuda_stream)


def _f32c(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.float32).contiguous()


# The residual shortcut's gradient arrives in the node that also produces the branch's dx (see _QkvWindowAttention /
# _LinearShortcut): the dx GEMM then accumulates straight into that incoming buffer (beta = 1, in place) instead of
# copying it first -- but ONLY when the buffer is one this package allocated itself and nothing else can see: the
# pass-through gradient of _LnResidual.backward is tagged with `_hv_owned` and the tag is checked here.  A gradient
# handed in by the caller (y.backward(g), torch.autograd.grad) or by any other autograd node is never modified.
INPLACE_SHORTCUT_GRAD = True


def _own(t: torch.Tensor) -> torch.Tensor:
    """Mark a gradient buffer allocated by this package as safe to accumulate into."""
    t._hv_owned = True
    return t


def _dx_with_shortcut(dx_shortcut, d2, weight, shape):
    C = shape[-1]
    if dx_shortcut is None:
        return torch.matmul(d2, weight).view(shape)
    if (INPLACE_SHORTCUT_GRAD and getattr(dx_shortcut, "_hv_owned", False) and dx_shortcut.dtype == d2.dtype
            and dx_shortcut.is_contiguous() and not torch.is_grad_enabled() and not dx_shortcut.requires_grad):
        return _own(dx_shortcut.view(-1, C).addmm_(d2, weight).view(shape))
    return _own(torch.addmm(dx_shortcut.reshape(-1, C).to(d2.dtype), d2, weight).view(shape))


# ---- low-precision shadows of fp32 master weights -------------------------------------------------------------------
# Under autocast every Linear casts its fp32 weight to bf16 in the forward and its bf16 weight gradient back to fp32 in
# the backward: ~126 tiny launches per SwinV2-T step.  Here a weight keeps one persistent bf16 shadow, rewritten only
# when the parameter's version counter has moved (optimizer step, load_state_dict) -- by ONE multi-tensor copy for all
# weights (`refresh_weight_shadows`, called at the start of a training step) -- and the weight-gradient GEMMs write
# fp32 directly (`torch.mm(..., out_dtype=torch.float32)`).
import os
import weakref

_SHADOWS = {}  # id(parameter) -> [weakref to the parameter, shadow tensor, parameter version it holds]
# (keyed by id: tensors compare element-wise, which rules out WeakKeyDictionary; the weakref callback drops dead entries)


def weight_shadow(w: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """``w`` in ``dtype``: ``w`` itself, or its persistent shadow (refreshed here if the parameter changed since)."""
    if w.dtype == dtype:
        return w
    key = id(w)
    ent = _SHADOWS.get(key)
    if ent is None or ent[0]() is not w or ent[1].dtype != dtype or ent[1].device != w.device or ent[1].shape != w.shape:
        ent = [weakref.ref(w, lambda _r, k=key: _SHADOWS.pop(k, None)), torch.empty_like(w, dtype=dtype), -1]
        _SHADOWS[key] = ent
    if ent[2] != w._version:
        with torch.no_grad():
            ent[1].copy_(w)
        ent[2] = w._version
    return ent[1]


def refresh_weight_shadows(force: bool = False) -> int:
    """Bring every registered shadow up to date with one multi-tensor copy; ``force`` rewrites all of them (what a
    captured CUDA graph needs: the copy must be part of the graph whether or not anything is stale at capture time).
    Returns the number of shadows rewritten."""
    src, dst, ents = [], [], []
    for ent in list(_SHADOWS.values()):
        w = ent[0]()
        if w is not None and w.is_cuda and (force or ent[2] != w._version):
            src.append(w.detach())
            dst.append(ent[1])
            ents.append((ent, w))
    if src:
        with torch.no_grad():
            torch._foreach_copy_(dst, src)
        for ent, w in ents:
            ent[2] = w._version
    return len(src)


def mark_weight_shadows_stale() -> None:
    """Host-only: every shadow is refreshed at its next use.  For parameter updates that do not move the autograd version
    counters: a replayed CUDA graph (no Python runs during a replay) and the one-pass optimizer kernel (raw pointers)."""
    for ent in _SHADOWS.values():
        ent[2] = -1


def _weight_grad(d2t: torch.Tensor, x2: torch.Tensor, wdtype: torch.dtype) -> torch.Tensor:
    """d2t (N, tokens) @ x2 (tokens, K) in the master weight's dtype, without a separate cast kernel."""
    if wdtype == torch.float32 and d2t.dtype == torch.bfloat16 and x2.dtype == torch.bfloat16 and d2t.is_cuda:
        return torch.mm(d2t, x2, out_dtype=torch.float32)
    g = torch.matmul(d2t, x2)
    return g if g.dtype == wdtype else g.to(wdtype)


class _ShadowLinear(torch.autograd.Function):
    """F.linear(x, weight, bias) for an fp32 master weight and low-precision activations: the GEMMs read the weight's
    persistent shadow, the weight gradient comes out of its GEMM in fp32."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        w = weight_shadow(weight, x.dtype)
        ctx.save_for_backward(x, w)
        ctx.wdtype = weight.dtype
        ctx.bdtype = bias.dtype if bias is not None else None
        return torch.nn.functional.linear(x, w, None if bias is None else bias.to(x.dtype))

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        d2 = dy.reshape(-1, dy.shape[-1])
        if d2.dtype != x.dtype:
            d2 = d2.to(x.dtype)
        dx = torch.matmul(d2, w).view(x.shape) if ctx.needs_input_grad[0] else None
        dw = _weight_grad(d2.t(), x.reshape(-1, x.shape[-1]), ctx.wdtype) if ctx.needs_input_grad[1] else None
        db = d2.sum(dim=0, dtype=torch.float32).to(ctx.bdtype) if (ctx.bdtype is not None and ctx.needs_input_grad[2]) else None
        return dx, dw, db


def shadow_linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``F.linear`` for activations in a lower precision than the (fp32) master ``weight``; see ``weight_shadow``."""
    return _ShadowLinear.apply(x, weight, bias)


# Optional per-launch instrumentation used by bench.py: when set to a list, every attention
# kernel launch appends (tag, start_event, end_event, windows).
PROFILE_EVENTS = None
LAUNCH_COUNT = 0


def _timed(tag, windows, stream_device, fn, kernels=1):
    global LAUNCH_COUNT
    LAUNCH_COUNT += kernels
    if PROFILE_EVENTS is None:
        return fn()
    s = torch.cuda.current_stream(stream_device)
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record(s)
    r = fn()
    e1.record(s)
    PROFILE_EVENTS.append((tag, e0, e1, windows))
    return r


def set_attention_forward_variant(variant: int) -> None:
    """0: mma.sync forward kernel, 1: tcgen05 / TMEM / TMA forward kernels (default), 2: the first-generation tcgen05
    kernel wherever valid, -1: HV_ATTN_TCGEN05 environment (unset: 1)."""
    check(_lib.load().hv_window_attn_fwd_variant(int(variant)), "hv_window_attn_fwd_variant")


def set_attention_backward_variant(variant: int) -> None:
    """0: mma.sync backward kernel, 1: tcgen05/TMEM/TMA backward kernel, -1: HV_ATTN_TCGEN05_BWD environment (unset: automatic)."""
    check(_lib.load().hv_window_attn_bwd_variant(int(variant)), "hv_window_attn_bwd_variant")


def set_attention_tc256_variant(variant: int) -> None:
    """16 x 16 windows, head dim 32, bf16: 1 tcgen05 / TMEM / TMA kernels (default), 0 generic CUDA-core kernels,
    -1: HV_ATTN_TC256 environment (unset: 1).  Set it before a forward and keep it until that forward's backward ran."""
    check(_lib.load().hv_window_attn_tc256_variant(int(variant)), "hv_window_attn_tc256_variant")


def window_attention_fwd_raw(qkv, bias_table, tau, mask, out, lse, B, H, W, C, heads, ws, shift):
    """Enqueue hv_window_attn_fwd on the current stream; every tensor is caller-allocated."""
    lib = _lib.load()
    with torch.cuda.device(qkv.device):
        rc = lib.hv_window_attn_fwd(_ptr(qkv), _ptr(bias_table), _ptr(tau), _ptr(mask),
                                    mask.shape[0] if mask is not None else 0, _ptr(out), _ptr(lse),
                                    B, H, W, C, heads, ws, shift, _code(qkv), _stream(qkv.device))
    check(rc, "hv_window_attn_fwd")


def window_attention_stats(qkv, B, H, W, C, heads, ws):
    """Uninitialised statistics buffer (`lse` of hv_window_attn_fwd / _bwd) for a geometry; its size depends on the
    kernel kind (hv_window_attn_stats_floats)."""
    n = int(_lib.load().hv_window_attn_stats_floats(B, H, W, C, heads, ws, _code(qkv)))
    if n <= 0:
        raise RuntimeError(f"window_attention: invalid geometry B={B} H={H} W={W} C={C} heads={heads} ws={ws}")
    return torch.empty((n,), dtype=torch.float32, device=qkv.device)


def window_attention_bwd_workspace(qkv, B, H, W, C, heads, ws):
    lib = _lib.load()
    with torch.cuda.device(qkv.device):
        nbytes = lib.hv_window_attn_bwd_workspace_bytes(B, H, W, C, heads, ws, _code(qkv))
    return torch.empty((max(int(nbytes), 16),), dtype=torch.uint8, device=qkv.device)


def window_attention_bwd_raw(qkv, out, dout, lse, bias_table, tau, mask, dqkv, dbias, dtau, workspace,
                             B, H, W, C, heads, ws, shift, dq_colsum=None):
    """Enqueue hv_window_attn_bwd on the current stream; every tensor is caller-allocated."""
    lib = _lib.load()
    with torch.cuda.device(qkv.device):
        rc = lib.hv_window_attn_bwd(_ptr(qkv), _ptr(out), _ptr(dout), _ptr(lse), _ptr(bias_table), _ptr(tau), _ptr(mask),
                                    mask.shape[0] if mask is not None else 0, _ptr(dqkv), _ptr(dbias), _ptr(dtau),
                                    _ptr(dq_colsum), _ptr(workspace), workspace.numel(), B, H, W, C, heads, ws, shift,
                                    _code(qkv), _stream(qkv.device))
    check(rc, "hv_window_attn_bwd")


class _WindowAttention(torch.autograd.Function):
    """out = fused shifted-window scaled-cosine attention(qkv); see hv_window_attn_fwd."""

    @staticmethod
    def forward(ctx, qkv, bias_table, tau, mask, B, H, W, C, heads, ws, shift):
        _need_cuda(qkv, "window_attention")
        qkv = qkv.contiguous()
        bias_table = _f32c(bias_table)
        tau = _f32c(tau)
        if mask is not None:
            mask = _f32c(mask)
        N = ws * ws
        nW = (H // ws) * (W // ws)
        out = torch.empty((B, H * W, C), dtype=qkv.dtype, device=qkv.device)
        lse = window_attention_stats(qkv, B, H, W, C, heads, ws)
        _timed(f"attn_fwd/C{C}/s{shift}", B * nW, qkv.device, lambda: window_attention_fwd_raw(
            qkv, bias_table, tau, mask, out, lse, B, H, W, C, heads, ws, shift))
        ctx.save_for_backward(qkv, out, lse, bias_table, tau, mask)
        ctx.geom = (B, H, W, C, heads, ws, shift)
        return out

    @staticmethod
    def backward(ctx, dout):
        qkv, out, lse, bias_table, tau, mask = ctx.saved_tensors
        B, H, W, C, heads, ws, shift = ctx.geom
        dout = dout.contiguous()
        if dout.dtype != qkv.dtype:
            dout = dout.to(qkv.dtype)
        dqkv = torch.empty_like(qkv)
        dbias = torch.empty_like(bias_table)
        dtau = torch.empty_like(tau)
        workspace = window_attention_bwd_workspace(qkv, B, H, W, C, heads, ws)
        _timed(f"attn_bwd/C{C}/s{shift}", B * (H // ws) * (W // ws), qkv.device, lambda: window_attention_bwd_raw(
            qkv, out, dout, lse, bias_table, tau, mask, dqkv, dbias, dtau, workspace, B, H, W, C, heads, ws, shift),
            kernels=2)  # backward kernel + partial-sum reduction kernel
        return dqkv, dbias, dtau, None, None, None, None, None, None, None, None


def window_attention(qkv: torch.Tensor, bias_table: torch.Tensor, tau: torch.Tensor, *, B: int, H: int, W: int,
                     C: int, heads: int, ws: int, shift: int, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """qkv (B, H*W, 3C) in image token order -> (B, H*W, C).  ``bias_table`` is
    16*sigmoid(cpb_mlp(coords)) of shape ((2ws-1)^2, heads); ``tau`` = exp(clamped logit_scale), (heads,).
    ``mask`` None: shifted-window mask generated in-kernel from the geometry; else (nW, N, N) added as given."""
    return _WindowAttention.apply(qkv, bias_table, tau.reshape(-1), mask, B, H, W, C, heads, ws, shift)


def window_attention_kernel_name(B: int, H: int, W: int, C: int, heads: int, ws: int, shift: int, dtype: torch.dtype,
                                 backward: bool) -> str:
    """Name of the kernel the fused attention launches for this geometry (hv_window_attn_kernel_name)."""
    import ctypes
    buf = ctypes.create_string_buffer(96)
    check(_lib.load().hv_window_attn_kernel_name(B, H, W, C, heads, ws, shift, HV_BF16 if dtype == torch.bfloat16 else HV_F32,
                                                 1 if backward else 0, buf, len(buf)), "hv_window_attn_kernel_name")
    return buf.value.decode()


def window_attention_kind(C: int, heads: int, ws: int, dtype: torch.dtype) -> int:
    """1 when (C, heads, ws, dtype) runs on the tensor-core kernel (N = 64, head dim 32, bf16), else 0."""
    if dtype not in (torch.float32, torch.bfloat16):
        return 0
    return int(_lib.load().hv_window_attn_kernel_kind(C, heads, ws, HV_BF16 if dtype == torch.bfloat16 else HV_F32))


class _QkvWindowAttention(torch.autograd.Function):
    """out = window_attention(x @ W^T + [q_bias, 0, 0]) for the tensor-core kernel.  One autograd node for the
    qkv Linear (cuBLAS) and the fused attention, so that the gradient of q_bias is the kernel's dq column sum
    (hv_window_attn_bwd's dq_colsum) instead of a separate reduction over the (tokens, 3C) gradient.  v_bias is
    not an input: softmax rows sum to one, so it is added to the output by the caller (swinv2.py:211-220, 261)."""

    @staticmethod
    def forward(ctx, x, weight, q_bias, bias_table, tau, B, H, W, C, heads, ws, shift):
        _need_cuda(x, "qkv_window_attention")
        x = x.contiguous()
        bias = None
        if q_bias is not None:
            bias = torch.zeros((3 * C,), dtype=x.dtype, device=x.device)
            bias[:C] = q_bias
        wdtype = weight.dtype
        weight = weight_shadow(weight, x.dtype)  # fp32 master weight under autocast: its persistent bf16 shadow
        qkv = torch.nn.functional.linear(x, weight, bias)
        bias_table = _f32c(bias_table)
        tau = _f32c(tau)
        nW = (H // ws) * (W // ws)
        out = torch.empty((B, H * W, C), dtype=qkv.dtype, device=qkv.device)
        lse = window_attention_stats(qkv, B, H, W, C, heads, ws)
        _timed(f"attn_fwd/C{C}/s{shift}", B * nW, qkv.device, lambda: window_attention_fwd_raw(
            qkv, bias_table, tau, None, out, lse, B, H, W, C, heads, ws, shift))
        ctx.save_for_backward(x, weight, qkv, out, lse, bias_table, tau)
        ctx.geom = (B, H, W, C, heads, ws, shift)
        ctx.q_bias_dtype = q_bias.dtype if q_bias is not None else None
        ctx.wdtype = wdtype
        # second output: x itself, for the residual shortcut.  Its gradient comes back into this node, where it is
        # the accumulator (beta = 1) of the dx GEMM instead of a separate elementwise add kernel.
        return out, x.view_as(x)

    @staticmethod
    def backward(ctx, dout, dx_shortcut):
        x, weight, qkv, out, lse, bias_table, tau = ctx.saved_tensors
        B, H, W, C, heads, ws, shift = ctx.geom
        dout = dout.contiguous()
        if dout.dtype != qkv.dtype:
            dout = dout.to(qkv.dtype)
        dqkv = torch.empty_like(qkv)
        dbias = torch.empty_like(bias_table)
        dtau = torch.empty_like(tau)
        workspace = window_attention_bwd_workspace(qkv, B, H, W, C, heads, ws)
        _timed(f"attn_bwd/C{C}/s{shift}", B * (H // ws) * (W // ws), qkv.device, lambda: window_attention_bwd_raw(
            qkv, out, dout, lse, bias_table, tau, None, dqkv, dbias, dtau, workspace, B, H, W, C, heads, ws, shift),
            kernels=2)  # backward + partial-sum fold
        dq_colsum = None
        if ctx.q_bias_dtype is not None and ctx.needs_input_grad[2]:
            # gradient of q_bias: column sums of dq, a streaming pass of its own (hv_dq_colsum)
            dq_colsum = torch.empty((C,), dtype=torch.float32, device=qkv.device)
            lib = _lib.load()
            with torch.cuda.device(qkv.device):
                nb = int(lib.hv_dq_colsum_workspace_bytes(C))
                cws = torch.empty((nb,), dtype=torch.uint8, device=qkv.device)
                check(lib.hv_dq_colsum(_ptr(dqkv), _ptr(dq_colsum), _ptr(cws), nb, B * H * W, C, _code(dqkv), _stream(qkv.device)),
                      "hv_dq_colsum")
            global LAUNCH_COUNT
            LAUNCH_COUNT += 2
        d2 = dqkv.view(-1, 3 * C)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = _dx_with_shortcut(dx_shortcut, d2, weight, x.shape)
        elif dx_shortcut is not None:
            dx = dx_shortcut
        dw = _weight_grad(d2.t(), x.view(-1, C), ctx.wdtype) if ctx.needs_input_grad[1] else None
        dqb = dq_colsum.to(ctx.q_bias_dtype) if dq_colsum is not None and ctx.needs_input_grad[2] else None
        return dx, dw, dqb, dbias, dtau, None, None, None, None, None, None, None


def qkv_window_attention(x: torch.Tensor, weight: torch.Tensor, q_bias: Optional[torch.Tensor],
                         bias_table: torch.Tensor, tau: torch.Tensor, *, B: int, H: int, W: int, C: int, heads: int,
                         ws: int, shift: int) -> torch.Tensor:
    """x (B, H*W, C) bf16, weight (3C, C) bf16 -> (attention output (B, H*W, C) WITHOUT the v_bias term (add it to the
    result, or fold W_proj v_bias into the proj bias), x again for the residual shortcut).  Use the returned x for the
    shortcut: its gradient is then accumulated by the dx GEMM.  Tensor-core kernel geometries only."""
    return _QkvWindowAttention.apply(x, weight, q_bias, bias_table, tau.reshape(-1), B, H, W, C, heads, ws, shift)


class _LinearShortcut(torch.autograd.Function):
    """(x W^T, x): a bias-free Linear whose input also feeds a residual shortcut.  The shortcut's gradient is the
    accumulator (beta = 1) of the dx GEMM, which removes the elementwise add autograd would otherwise launch at
    the fork (swinv2.py:434: x = x + drop_path(norm2(mlp(x))))."""

    @staticmethod
    def forward(ctx, x, weight):
        ctx.wdtype = weight.dtype
        weight = weight_shadow(weight, x.dtype)
        ctx.save_for_backward(x, weight)
        return torch.nn.functional.linear(x, weight), x.view_as(x)

    @staticmethod
    def backward(ctx, dy, dx_shortcut):
        x, weight = ctx.saved_tensors
        d2 = dy.reshape(-1, dy.shape[-1])
        dx = None
        if ctx.needs_input_grad[0]:
            dx = _dx_with_shortcut(dx_shortcut, d2, weight, x.shape)
        elif dx_shortcut is not None:
            dx = dx_shortcut
        dw = _weight_grad(d2.t(), x.reshape(-1, x.shape[-1]), ctx.wdtype) if ctx.needs_input_grad[1] else None
        return dx, dw


def linear_shortcut(x: torch.Tensor, weight: torch.Tensor):
    """Returns (F.linear(x, weight), x); use the returned x for the residual shortcut."""
    return _LinearShortcut.apply(x, weight)


class _LnResidual(torch.autograd.Function):
    """out = shortcut + keep_scale[sample] * LayerNorm(y + bias) (shortcut / bias / keep_scale optional)."""

    @staticmethod
    def forward(ctx, y, shortcut, gamma, beta, bias, keep_scale, rows_per_sample, eps, out_dtype=None):
        _need_cuda(y, "ln_residual")
        lib = _lib.load()
        y = y.contiguous()
        C = y.shape[-1]
        rows = y.numel() // C
        gamma32, beta32 = _f32c(gamma), _f32c(beta)
        bias32 = _f32c(bias) if bias is not None else None
        if shortcut is not None:
            shortcut = shortcut.contiguous()
            res_dtype = shortcut.dtype
        else:
            res_dtype = out_dtype or y.dtype
        if keep_scale is not None:
            keep_scale = _f32c(keep_scale)
        out = torch.empty(y.shape, dtype=res_dtype, device=y.device)
        mean = torch.empty((rows,), dtype=torch.float32, device=y.device)
        rstd = torch.empty((rows,), dtype=torch.float32, device=y.device)
        with torch.cuda.device(y.device):
            rc = lib.hv_ln_residual_fwd(_ptr(y), _ptr(shortcut), _ptr(gamma32), _ptr(beta32), _ptr(bias32), _ptr(keep_scale),
                                        _ptr(out), _ptr(mean), _ptr(rstd), rows, C, rows_per_sample, float(eps), _code(y),
                                        _code(out), _stream(y.device))
        check(rc, "hv_ln_residual_fwd")
        global LAUNCH_COUNT
        LAUNCH_COUNT += 1
        ctx.save_for_backward(y, gamma32, bias32, mean, rstd, keep_scale)
        ctx.meta = (rows, C, rows_per_sample, shortcut is not None, gamma.dtype, beta.dtype,
                    bias.dtype if bias is not None else None)
        return out

    @staticmethod
    def backward(ctx, dout):
        y, gamma32, bias32, mean, rstd, keep_scale = ctx.saved_tensors
        rows, C, rows_per_sample, has_shortcut, gdt, bdt, biasdt = ctx.meta
        lib = _lib.load()
        dout = dout.contiguous()
        dy = torch.empty_like(y)
        dgamma = torch.empty_like(gamma32)
        dbeta = torch.empty_like(gamma32)
        dbias = torch.empty_like(gamma32) if bias32 is not None else None
        with torch.cuda.device(y.device):
            nbytes = lib.hv_ln_residual_bwd_workspace_bytes(rows, C)
            workspace = torch.empty((int(nbytes),), dtype=torch.uint8, device=y.device)
            rc = lib.hv_ln_residual_bwd(_ptr(dout), _ptr(y), _ptr(gamma32), _ptr(bias32), _ptr(mean), _ptr(rstd),
                                        _ptr(keep_scale), _ptr(dy), _ptr(dgamma), _ptr(dbeta), _ptr(dbias), _ptr(workspace),
                                        workspace.numel(), rows, C, rows_per_sample, _code(y), _code(dout), _stream(y.device))
        check(rc, "hv_ln_residual_bwd")
        global LAUNCH_COUNT
        LAUNCH_COUNT += 2
        return (dy, (dout if has_shortcut else None), dgamma.to(gdt), dbeta.to(bdt),
                (dbias.to(biasdt) if dbias is not None else None), None, None, None, None)


def ln_residual(y: torch.Tensor, shortcut: Optional[torch.Tensor], gamma: torch.Tensor, beta: torch.Tensor,
                keep_scale: Optional[torch.Tensor] = None, eps: float = 1e-5,
                bias: Optional[torch.Tensor] = None, out_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """shortcut + keep_scale[b] * LayerNorm(y + bias) over the last dim; y is (B, L, C) (or (rows, C)).
    ``bias``: the bias of the Linear that produced ``y`` when that Linear was run without it.
    The result has the shortcut's dtype; without a shortcut ``out_dtype`` (float32 / bfloat16, default: y's)."""
    rows_per_sample = (y.numel() // y.shape[-1]) // y.shape[0] if y.dim() >= 2 else 1
    return _LnResidual.apply(y, shortcut, gamma, beta, bias, keep_scale, rows_per_sample, eps, out_dtype)


class _BiasGelu(torch.autograd.Function):
    """out = GELU_erf(h + bias); backward returns (dh, dbias) with dbias = column sums of dh."""

    @staticmethod
    def forward(ctx, h, bias):
        _need_cuda(h, "bias_gelu")
        lib = _lib.load()
        h = h.contiguous()
        cols = h.shape[-1]
        rows = h.numel() // cols
        bias32 = _f32c(bias)
        out = torch.empty_like(h)
        with torch.cuda.device(h.device):
            rc = lib.hv_bias_gelu_fwd(_ptr(h), _ptr(bias32), _ptr(out), rows, cols, _code(h), _stream(h.device))
        check(rc, "hv_bias_gelu_fwd")
        global LAUNCH_COUNT
        LAUNCH_COUNT += 1
        ctx.save_for_backward(h, bias32)
        ctx.meta = (rows, cols, bias.dtype)
        return out

    @staticmethod
    def backward(ctx, dout):
        h, bias32 = ctx.saved_tensors
        rows, cols, bdt = ctx.meta
        lib = _lib.load()
        dout = dout.contiguous()
        if dout.dtype != h.dtype:
            dout = dout.to(h.dtype)
        dh = torch.empty_like(h)
        dbias = torch.empty_like(bias32)
        with torch.cuda.device(h.device):
            nbytes = lib.hv_bias_gelu_bwd_workspace_bytes(rows, cols)
            workspace = torch.empty((int(nbytes),), dtype=torch.uint8, device=h.device)
            rc = lib.hv_bias_gelu_bwd(_ptr(dout), _ptr(h), _ptr(bias32), _ptr(dh), _ptr(dbias), _ptr(workspace),
                                      workspace.numel(), rows, cols, _code(h), _stream(h.device))
        check(rc, "hv_bias_gelu_bwd")
        global LAUNCH_COUNT
        LAUNCH_COUNT += 2
        return dh, dbias.to(bdt)


def bias_gelu_supported(h: torch.Tensor) -> bool:
    cols = h.shape[-1]
    return h.is_cuda and ((h.dtype == torch.bfloat16 and cols % 128 == 0) or (h.dtype == torch.float32 and cols % 64 == 0))


def bias_gelu(h: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    """GELU (exact erf form) of ``h + bias`` where ``h`` is the bias-free fc1 output (..., cols)."""
    return _BiasGelu.apply(h, bias)


# the fused backward GEMM (hv_mlp_dgelu_gemm) replaces the fc2 dgrad GEMM + hv_bias_gelu_bwd where it is faster than that
# pair; measured on B200 (tools/mlp_gemm_check.py): faster up to C = 512 (0.123 vs 0.132 ms at the SwinV2-B stage-2 shape),
# slower at C = 768 (0.123 vs 0.117).  A test / benchmarking switch.
MLP_DGELU_GEMM_MAX_C = int(os.environ.get("HV_MLP_DGELU_GEMM_MAX_C", "512"))


class _GeluFc2(torch.autograd.Function):
    """m = GELU_erf(h + b1) @ W2^T (the fc2 bias is added by the caller or folded into its LayerNorm kernel): reference
    swinv2.py:61-64.  One autograd node for the activation and the fc2 GEMM so that the backward can form
    dh = (dm W2) * GELU'(h + b1) and db1 in ONE tcgen05 GEMM (hv_mlp_dgelu_gemm): the gradient of the activation output,
    the largest tensor of the block, is never written."""

    @staticmethod
    def forward(ctx, h, b1, w2):
        _need_cuda(h, "gelu_fc2")
        lib = _lib.load()
        h = h.contiguous()
        cols = h.shape[-1]
        rows = h.numel() // cols
        b32 = _f32c(b1)
        a = torch.empty_like(h)
        with torch.cuda.device(h.device):
            check(lib.hv_bias_gelu_fwd(_ptr(h), _ptr(b32), _ptr(a), rows, cols, _code(h), _stream(h.device)), "hv_bias_gelu_fwd")
        global LAUNCH_COUNT
        LAUNCH_COUNT += 1
        w = weight_shadow(w2, h.dtype)
        ctx.save_for_backward(h, b32, a, w)
        ctx.meta = (rows, cols, b1.dtype, w2.dtype)
        return torch.nn.functional.linear(a, w)

    @staticmethod
    def backward(ctx, dm):
        h, b32, a, w = ctx.saved_tensors
        rows, cols, bdt, wdt = ctx.meta
        lib = _lib.load()
        C = w.shape[0]
        dm2 = dm.reshape(rows, C)
        if dm2.dtype != h.dtype:
            dm2 = dm2.to(h.dtype)
        dm2 = dm2.contiguous()
        global LAUNCH_COUNT
        dh = db1 = None
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
            dh = torch.empty_like(h)
            db1 = torch.empty_like(b32)
            with torch.cuda.device(h.device):
                nb = int(lib.hv_mlp_dgelu_gemm_workspace_bytes(rows, cols, C)) if (h.dtype == torch.bfloat16 and C <= MLP_DGELU_GEMM_MAX_C) else 0
                if nb > 0:
                    ws = torch.empty((nb,), dtype=torch.uint8, device=h.device)
                    check(lib.hv_mlp_dgelu_gemm(_ptr(dm2), _ptr(w), _ptr(h), _ptr(b32), _ptr(dh), _ptr(db1), _ptr(ws), nb, rows,
                                                cols, C, _code(h), _stream(h.device)), "hv_mlp_dgelu_gemm")
                else:  # the two-kernel path: dgrad GEMM, then the bias + GELU backward
                    da = torch.mm(dm2, w)
                    nb = int(lib.hv_bias_gelu_bwd_workspace_bytes(rows, cols))
                    ws = torch.empty((nb,), dtype=torch.uint8, device=h.device)
                    check(lib.hv_bias_gelu_bwd(_ptr(da), _ptr(h), _ptr(b32), _ptr(dh), _ptr(db1), _ptr(ws), ws.numel(), rows, cols,
                                               _code(h), _stream(h.device)), "hv_bias_gelu_bwd")
            LAUNCH_COUNT += 2
        dw2 = _weight_grad(dm2.t(), a.reshape(rows, cols), wdt) if ctx.needs_input_grad[2] else None
        return dh, (db1.to(bdt) if db1 is not None else None), dw2


# the fused forward GEMM (hv_mlp_fc1_gelu_gemm: h and GELU(h + b1) from one kernel) where it beats cuBLAS + hv_bias_gelu_fwd:
# measured faster up to C = 384 (0.306 / 0.214 / 0.133 ms against 0.474 / 0.239 / 0.143), slower at 768 (a 128 x 128 single-CTA
# tile is not a cuBLAS-class mainloop)
MLP_FC1_GELU_GEMM_MAX_C = int(os.environ.get("HV_MLP_FC1_GELU_GEMM_MAX_C", "384"))


class _MlpFused(torch.autograd.Function):
    """(m, x) with m = GELU_erf(x W1^T + b1) W2^T (the fc2 bias is left to the caller) for bf16 activations: the whole Mlp of
    reference swinv2.py:43-66 as ONE autograd node.  Forward: h = x W1^T and a = GELU(h + b1) from one tcgen05 GEMM where
    that is faster than cuBLAS + the activation kernel, then the fc2 GEMM.  Backward: dh = (dm W2) * GELU'(h + b1) and
    db1 from one tcgen05 GEMM (hv_mlp_dgelu_gemm), the weight gradients in fp32, and dx = dh W1 accumulated onto the
    gradient of the residual shortcut (second output) inside the GEMM."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2):
        _need_cuda(x, "mlp_fused")
        lib = _lib.load()
        x = x.contiguous()
        C = x.shape[-1]
        rows = x.numel() // C
        hidden = w1.shape[0]
        w1s, w2s = weight_shadow(w1, x.dtype), weight_shadow(w2, x.dtype)
        b32 = _f32c(b1)
        x2 = x.view(rows, C)
        global LAUNCH_COUNT
        with torch.cuda.device(x.device):
            fused = C <= MLP_FC1_GELU_GEMM_MAX_C and int(lib.hv_mlp_dgelu_gemm_workspace_bytes(rows, hidden, C)) > 0
            if fused:
                h = torch.empty((rows, hidden), dtype=x.dtype, device=x.device)
                a = torch.empty_like(h)
                check(lib.hv_mlp_fc1_gelu_gemm(_ptr(x2), _ptr(w1s), _ptr(b32), _ptr(h), _ptr(a), rows, hidden, C, _code(x),
                                               _stream(x.device)), "hv_mlp_fc1_gelu_gemm")
            else:
                h = torch.mm(x2, w1s.t())
                a = torch.empty_like(h)
                check(lib.hv_bias_gelu_fwd(_ptr(h), _ptr(b32), _ptr(a), rows, hidden, _code(h), _stream(h.device)), "hv_bias_gelu_fwd")
        LAUNCH_COUNT += 1
        ctx.save_for_backward(x2, w1s, h, b32, a, w2s)
        ctx.meta = (x.shape, b1.dtype, w1.dtype, w2.dtype)
        m = torch.mm(a, w2s.t()).view(*x.shape[:-1], w2.shape[0])
        return m, x.view_as(x)

    @staticmethod
    def backward(ctx, dm, dx_shortcut):
        x2, w1s, h, b32, a, w2s = ctx.saved_tensors
        xshape, bdt, w1dt, w2dt = ctx.meta
        lib = _lib.load()
        rows, C = x2.shape
        hidden = h.shape[1]
        Co = w2s.shape[0]
        dm2 = dm.reshape(rows, Co)
        if dm2.dtype != h.dtype:
            dm2 = dm2.to(h.dtype)
        dm2 = dm2.contiguous()
        dh = torch.empty_like(h)
        db1 = torch.empty_like(b32)
        global LAUNCH_COUNT
        with torch.cuda.device(h.device):
            nb = int(lib.hv_mlp_dgelu_gemm_workspace_bytes(rows, hidden, Co)) if Co <= MLP_DGELU_GEMM_MAX_C else 0
            if nb > 0:
                ws = torch.empty((nb,), dtype=torch.uint8, device=h.device)
                check(lib.hv_mlp_dgelu_gemm(_ptr(dm2), _ptr(w2s), _ptr(h), _ptr(b32), _ptr(dh), _ptr(db1), _ptr(ws), nb, rows, hidden,
                                            Co, _code(h), _stream(h.device)), "hv_mlp_dgelu_gemm")
            else:  # the two-kernel path: dgrad GEMM, then the bias + GELU backward
                da = torch.mm(dm2, w2s)
                nb = int(lib.hv_bias_gelu_bwd_workspace_bytes(rows, hidden))
                ws = torch.empty((nb,), dtype=torch.uint8, device=h.device)
                check(lib.hv_bias_gelu_bwd(_ptr(da), _ptr(h), _ptr(b32), _ptr(dh), _ptr(db1), _ptr(ws), ws.numel(), rows, hidden,
                                           _code(h), _stream(h.device)), "hv_bias_gelu_bwd")
        LAUNCH_COUNT += 2
        dw2 = _weight_grad(dm2.t(), a, w2dt) if ctx.needs_input_grad[3] else None
        dx = None
        if ctx.needs_input_grad[0]:
            dx = _dx_with_shortcut(dx_shortcut, dh, w1s, xshape)
        elif dx_shortcut is not None:
            dx = dx_shortcut
        dw1 = _weight_grad(dh.t(), x2, w1dt) if ctx.needs_input_grad[1] else None
        return dx, dw1, (db1.to(bdt) if ctx.needs_input_grad[2] else None), dw2


def mlp_fused(x: torch.Tensor, w1: torch.Tensor, b1: torch.Tensor, w2: torch.Tensor):
    """Returns (GELU(x w1^T + b1) w2^T, x): bf16 ``x`` (..., C); use the returned x for the residual shortcut."""
    return _MlpFused.apply(x, w1, b1, w2)


def gelu_fc2(h: torch.Tensor, b1: torch.Tensor, w2: torch.Tensor) -> torch.Tensor:
    """GELU(h + b1) @ w2^T for the bias-free fc1 output ``h`` (..., hidden) and ``w2`` = fc2.weight (C, hidden)."""
    return _GeluFc2.apply(h, b1, w2)


class _PatchMergeGather(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, H, W):
        _need_cuda(x, "patch_merge_gather")
        lib = _lib.load()
        x = x.contiguous()
        B, L, C = x.shape
        out = torch.empty((B, L // 4, 4 * C), dtype=x.dtype, device=x.device)
        with torch.cuda.device(x.device):
            rc = lib.hv_patch_merge_gather_fwd(_ptr(x), _ptr(out), B, H, W, C, _code(x), _stream(x.device))
        check(rc, "hv_patch_merge_gather_fwd")
        global LAUNCH_COUNT
        LAUNCH_COUNT += 1
        ctx.meta = (B, H, W, C)
        return out

    @staticmethod
    def backward(ctx, dout):
        B, H, W, C = ctx.meta
        lib = _lib.load()
        dout = dout.contiguous()
        dx = torch.empty((B, H * W, C), dtype=dout.dtype, device=dout.device)
        with torch.cuda.device(dout.device):
            rc = lib.hv_patch_merge_gather_bwd(_ptr(dout), _ptr(dx), B, H, W, C, _code(dout), _stream(dout.device))
        check(rc, "hv_patch_merge_gather_bwd")
        global LAUNCH_COUNT
        LAUNCH_COUNT += 1
        return dx, None, None


def patch_merge_gather(x: torch.Tensor, H: int, W: int) -> torch.Tensor:
    """(B, H*W, C) -> (B, H/2*W/2, 4C), channel blocks ordered (0,0),(1,0),(0,1),(1,1)."""
    return _PatchMergeGather.apply(x, H, W)


def patch_rows(img: torch.Tensor, scale: Optional[torch.Tensor], shift: Optional[torch.Tensor],
               out_dtype: torch.dtype, patch: int = 4) -> torch.Tensor:
    """(B, 3, H, W) uint8 / float32 / bfloat16 image -> (B * H/4 * W/4, 48) rows in the conv weight's (c, dy, dx)
    order, each pixel mapped to ``x * scale[c] + shift[c]`` (both None: identity).  No gradient flows to the image."""
    _need_cuda(img, "patch_rows")
    if img.requires_grad:
        raise RuntimeError("patch_rows does not differentiate with respect to the image")
    lib = _lib.load()
    img = img.contiguous()
    B, C, H, W = img.shape
    code = HV_U8 if img.dtype == torch.uint8 else _code(img)
    out = torch.empty((B * (H // patch) * (W // patch), C * patch * patch), dtype=out_dtype, device=img.device)
    if scale is not None:
        scale, shift = _f32c(scale), _f32c(shift)
    with torch.cuda.device(img.device):
        rc = lib.hv_patch_rows(_ptr(img), code, _ptr(scale), _ptr(shift), _ptr(out), _code(out), B, C, H, W, patch,
                               _stream(img.device))
    check(rc, "hv_patch_rows")
    global LAUNCH_COUNT
    LAUNCH_COUNT += 1
    return out


class _CpbBias(torch.autograd.Function):
    """table = 16 * sigmoid(Linear(hidden, heads, no bias)(relu(Linear(2, hidden)(coords)))), fp32."""

    @staticmethod
    def forward(ctx, coords, w1, b1, w2):
        _need_cuda(coords, "cpb_bias")
        lib = _lib.load()
        coords, w1, b1, w2 = _f32c(coords.reshape(-1, 2)), _f32c(w1), _f32c(b1), _f32c(w2)
        M, hid, heads = coords.shape[0], w1.shape[0], w2.shape[0]
        table = torch.empty((M, heads), dtype=torch.float32, device=coords.device)
        with torch.cuda.device(coords.device):
            rc = lib.hv_cpb_bias_fwd(_ptr(coords), _ptr(w1), _ptr(b1), _ptr(w2), _ptr(table), M, hid, heads,
                                     _stream(coords.device))
        check(rc, "hv_cpb_bias_fwd")
        global LAUNCH_COUNT
        LAUNCH_COUNT += 1
        ctx.save_for_backward(coords, w1, b1, w2)
        return table

    @staticmethod
    def backward(ctx, dtable):
        coords, w1, b1, w2 = ctx.saved_tensors
        lib = _lib.load()
        M, hid, heads = coords.shape[0], w1.shape[0], w2.shape[0]
        dtable = _f32c(dtable)
        dw1, db1, dw2 = torch.empty_like(w1), torch.empty_like(b1), torch.empty_like(w2)
        workspace = torch.empty((M, heads), dtype=torch.float32, device=coords.device)
        with torch.cuda.device(coords.device):
            rc = lib.hv_cpb_bias_bwd(_ptr(coords), _ptr(w1), _ptr(b1), _ptr(w2), _ptr(dtable), _ptr(dw1), _ptr(db1),
                                     _ptr(dw2), _ptr(workspace), M, hid, heads, _stream(coords.device))
        check(rc, "hv_cpb_bias_bwd")
        global LAUNCH_COUNT
        LAUNCH_COUNT += 2
        return None, dw1, db1, dw2


def cpb_bias_supported(w1: torch.Tensor, w2: torch.Tensor) -> bool:
    return w1.is_cuda and w1.shape[0] <= 512 and w1.shape[0] % 32 == 0 and w2.shape[0] <= 32


def cpb_bias(coords: torch.Tensor, w1: torch.Tensor, b1: torch.Tensor, w2: torch.Tensor) -> torch.Tensor:
    """((2ws-1)^2, heads) continuous position bias table, 16 * sigmoid(cpb_mlp(coords)) (swinv2.py:233-246)."""
    return _CpbBias.apply(coords, w1, b1, w2)


class _CrossEntropy(torch.autograd.Function):
    """mean over rows of CE(logits, target) with label smoothing; loss rows and d logits come out of one kernel
    (hv_cross_entropy_fwd_grad), the backward only multiplies by the upstream scalar."""

    @staticmethod
    def forward(ctx, logits, target, smoothing, weight):
        _need_cuda(logits, "cross_entropy")
        lib = _lib.load()
        logits = logits.contiguous()
        rows, classes = logits.shape
        target = target.contiguous()
        if target.dtype != torch.int64 or target.shape != (rows,):
            raise RuntimeError("cross_entropy: target must be int64 class indices of shape (rows,)")
        loss_rows = torch.empty((rows,), dtype=torch.float32, device=logits.device)
        dlogits = torch.empty_like(logits)
        with torch.cuda.device(logits.device):
            rc = lib.hv_cross_entropy_fwd_grad(_ptr(logits), _ptr(target), _ptr(loss_rows), _ptr(dlogits), rows, classes,
                                               float(smoothing), float(weight) / rows, _code(logits), _stream(logits.device))
        check(rc, "hv_cross_entropy_fwd_grad")
        global LAUNCH_COUNT
        LAUNCH_COUNT += 1
        ctx.save_for_backward(dlogits)
        return loss_rows.mean() * weight

    @staticmethod
    def backward(ctx, g):
        (dlogits,) = ctx.saved_tensors
        return dlogits * g.to(dlogits.dtype), None, None, None


def cross_entropy(logits: torch.Tensor, target: torch.Tensor, label_smoothing: float = 0.0, weight: float = 1.0) -> torch.Tensor:
    """``weight * F.cross_entropy(logits, target, label_smoothing=...)`` (mean reduction) for (rows, classes) float32 /
    bfloat16 logits on the device; one kernel forward, one elementwise multiply backward."""
    return _CrossEntropy.apply(logits, target, label_smoothing, weight)
