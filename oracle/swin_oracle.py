"""TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the SwinV2 windowed-attention
hot path of samuelstevens/hierarchical-vision (``swinv2.py``).

Who may import this: ``tests/``, ``__graft_entry__.smoke()`` and the CPU legs of
``bench.py`` (``cpu_baseline`` / ``--impl reference``).  The product package
``hierarchical_vision_b200`` never imports, links or executes anything in ``oracle/``.

Parity status: **pinned against the reference executed in the build container** -- the
reference ships no tests or golden vectors of its own, so
``oracle/make_goldens.py`` imports the unmodified ``/root/reference/swinv2.py``, runs it
on seeded inputs and commits inputs + outputs + gradients as fixtures under
``tests/golden/``; ``tests/test_oracle_golden.py`` checks this file against those fixtures
(and, when ``/root/reference`` is present, against the live reference).

The file is a *functional* restatement: closed-form integer maps (numpy) and the
floating-point algebra (torch CPU tensors, dtype-generic so it can also run in fp64).
Every function names the reference lines it follows.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

MASK_VALUE = -100.0  # swinv2.py:382-384 -- finite, *not* -inf
NORM_EPS = 1e-12  # F.normalize default eps, swinv2.py:229
LN_EPS = 1e-5  # nn.LayerNorm default, swinv2.py:336/348/473
# swinv2.py:138 -- a float32 buffer: log(float32(100)) = 4.605170249938965 (not the fp64 log)
LOGIT_CLAMP_MAX = float(torch.log(torch.tensor(1.0 / 0.01)))


# --------------------------------------------------------------------------------------
# Integer maps (bit-exact contracts)
# --------------------------------------------------------------------------------------
def relative_position_index(ws: int) -> np.ndarray:
    """(N,N) int64, N = ws*ws.  swinv2.py:175-190.

    Token i sits at (ih, iw) = divmod(i, ws) (``ij`` meshgrid order, swinv2.py:178).
    index[i, j] = (ih - jh + ws - 1) * (2*ws - 1) + (iw - jw + ws - 1).
    """
    t = np.arange(ws * ws, dtype=np.int64)
    th, tw = t // ws, t % ws
    dh = th[:, None] - th[None, :] + (ws - 1)
    dw = tw[:, None] - tw[None, :] + (ws - 1)
    return dh * (2 * ws - 1) + dw


def relative_coords_table(ws: int, pretrained_ws: int = 0) -> torch.Tensor:
    """(1, 2ws-1, 2ws-1, 2) float32 log-spaced coordinates.  swinv2.py:147-173.

    Same float32 operation order as the reference so the result is bit-identical:
    delta / (ws-1) * 8  ->  sign * log2(|.| + 1) / log2(8).
    """
    span = torch.arange(-(ws - 1), ws, dtype=torch.float32)
    denom = (pretrained_ws - 1) if pretrained_ws > 0 else (ws - 1)
    gh = span.view(-1, 1).expand(2 * ws - 1, 2 * ws - 1)
    gw = span.view(1, -1).expand(2 * ws - 1, 2 * ws - 1)
    tab = torch.stack([gh, gw], dim=-1).contiguous().unsqueeze(0)
    tab = tab / denom
    tab = tab * 8
    return torch.sign(tab) * torch.log2(torch.abs(tab) + 1.0) / np.log2(8)


def effective_window(resolution: Tuple[int, int], ws: int, shift: int) -> Tuple[int, int]:
    """Window / shift clamp when the grid is not larger than the window.  swinv2.py:328-334."""
    if min(resolution) <= ws:
        return min(resolution), 0
    assert 0 <= shift < ws, "shift_size must in 0-window_size"
    return ws, shift


def _region(p: np.ndarray, length: int, ws: int, shift: int) -> np.ndarray:
    # three bands of the *shifted* axis: [0, L-ws), [L-ws, L-shift), [L-shift, L)   swinv2.py:361-370
    return np.where(p < length - ws, 0, np.where(p < length - shift, 1, 2))


def shift_window_mask(H: int, W: int, ws: int, shift: int) -> Optional[np.ndarray]:
    """(nW, N, N) float32 with values {0, -100}; None when shift == 0.  swinv2.py:357-388."""
    if shift == 0:
        return None
    hh, ww = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    ident = 3 * _region(hh, H, ws, shift) + _region(ww, W, ws, shift)  # (H, W) in shifted coords
    # partition the id image exactly like the activations (swinv2.py:377-380)
    ident = ident.reshape(H // ws, ws, W // ws, ws).transpose(0, 2, 1, 3).reshape(-1, ws * ws)
    differ = ident[:, None, :] != ident[:, :, None]
    return np.where(differ, np.float32(MASK_VALUE), np.float32(0.0)).astype(np.float32)


def window_token_index(B: int, H: int, W: int, ws: int, shift: int) -> np.ndarray:
    """(B*nW, N) int64: flat index into the (B*H*W) token axis feeding window row r, slot i.

    Composition of the cyclic shift (roll by -shift, swinv2.py:399-404) with
    ``window_partition`` (swinv2.py:69-83): window row r = b*nW + wh*(W/ws) + ww,
    slot i = ih*ws + iw reads token (b, (wh*ws+ih+shift) % H, (ww*ws+iw+shift) % W).
    The reverse path (swinv2.py:420-429) is the inverse permutation, so the same table
    is the scatter map for the output.
    """
    nwh, nww = H // ws, W // ws
    b = np.arange(B).reshape(B, 1, 1, 1, 1)
    wh = np.arange(nwh).reshape(1, nwh, 1, 1, 1)
    wwi = np.arange(nww).reshape(1, 1, nww, 1, 1)
    ih = np.arange(ws).reshape(1, 1, 1, ws, 1)
    iw = np.arange(ws).reshape(1, 1, 1, 1, ws)
    row = (wh * ws + ih + shift) % H
    col = (wwi * ws + iw + shift) % W
    flat = (b * H + row) * W + col
    return flat.reshape(B * nwh * nww, ws * ws).astype(np.int64)


def merge_token_index(B: int, H: int, W: int) -> np.ndarray:
    """(B*H/2*W/2, 4) int64 source tokens of PatchMerging's concat.  swinv2.py:486-490.

    Channel block m of output token (b, i, j) comes from m=0:(2i,2j) 1:(2i+1,2j)
    2:(2i,2j+1) 3:(2i+1,2j+1).
    """
    assert H % 2 == 0 and W % 2 == 0, f"x size ({H}*{W}) are not even."
    b = np.arange(B).reshape(B, 1, 1, 1)
    i = np.arange(H // 2).reshape(1, H // 2, 1, 1)
    j = np.arange(W // 2).reshape(1, 1, W // 2, 1)
    m = np.arange(4).reshape(1, 1, 1, 4)
    row = 2 * i + (m % 2)
    col = 2 * j + (m // 2)
    return ((b * H + row) * W + col).reshape(-1, 4).astype(np.int64)


# --------------------------------------------------------------------------------------
# Floating point pieces
# --------------------------------------------------------------------------------------
def position_bias_table(coords: torch.Tensor, w0, b0, w2) -> torch.Tensor:
    """cpb_mlp(relative_coords_table).view(-1, heads) followed by 16*sigmoid.

    swinv2.py:141-145 (Linear(2,512)+ReLU+Linear(512,h,no bias)), 233-235, 246.  The
    sigmoid commutes with the gather at 236-245, so it is applied on the ((2ws-1)^2, h)
    table; ``expand_bias`` does the gather.
    """
    hidden = F.relu(F.linear(coords.to(w0.dtype), w0, b0))
    table = F.linear(hidden, w2).reshape(-1, w2.shape[0])
    return 16.0 * torch.sigmoid(table)


def expand_bias(bias_table: torch.Tensor, ws: int) -> torch.Tensor:
    """((2ws-1)^2, h) -> (h, N, N) through relative_position_index.  swinv2.py:236-245."""
    rpi = torch.from_numpy(relative_position_index(ws)).reshape(-1)
    N = ws * ws
    return bias_table[rpi].reshape(N, N, -1).permute(2, 0, 1).contiguous()


def logit_tau(logit_scale: torch.Tensor) -> torch.Tensor:
    """exp(min(logit_scale, log 100)) per head, flattened to (h,).  swinv2.py:230."""
    return torch.clamp(logit_scale, max=LOGIT_CLAMP_MAX).exp().reshape(-1)


@dataclass(frozen=True)
class Geometry:
    B: int
    H: int
    W: int
    C: int
    heads: int
    ws: int
    shift: int

    @property
    def N(self):
        return self.ws * self.ws

    @property
    def nW(self):
        return (self.H // self.ws) * (self.W // self.ws)

    @property
    def d(self):
        return self.C // self.heads


def attention_core_forward(qkv, bias, tau, g: Geometry, mask: Optional[torch.Tensor] = None,
                           use_shift_mask: bool = True):
    """Fused-kernel boundary: qkv (B, H*W, 3C) in *image token order* -> o (B, H*W, C).

    roll+partition (swinv2.py:399-412), head split (221-226), cosine logits (229-231),
    +bias (247), +mask (249-254), softmax (255-257), P@V and head merge (261), window
    reverse + un-roll (420-429).  The per-token linear layers before/after commute with
    the token permutation, which is why the boundary can sit on un-partitioned tokens.

    Returns (o, lse) with lse (B*nW, heads, N) the natural-log row normaliser.
    """
    B, L, C3 = qkv.shape
    C, h, d, N = g.C, g.heads, g.d, g.N
    assert L == g.H * g.W and C3 == 3 * C
    idx = torch.from_numpy(window_token_index(B, g.H, g.W, g.ws, g.shift)).reshape(-1)
    win = qkv.reshape(B * L, 3 * C)[idx].reshape(-1, N, 3, h, d).permute(2, 0, 3, 1, 4)
    q, k, v = win[0], win[1], win[2]  # (B_, h, N, d)
    qn = q / q.norm(dim=-1, keepdim=True).clamp_min(NORM_EPS)
    kn = k / k.norm(dim=-1, keepdim=True).clamp_min(NORM_EPS)
    s = (qn @ kn.transpose(-2, -1)) * tau.reshape(1, h, 1, 1) + bias.unsqueeze(0)
    if mask is None and use_shift_mask and g.shift > 0:
        m = shift_window_mask(g.H, g.W, g.ws, g.shift)
        mask = torch.from_numpy(m).to(s.dtype)
    if mask is not None:
        nW = mask.shape[0]
        s = (s.reshape(-1, nW, h, N, N) + mask.reshape(1, nW, 1, N, N)).reshape(-1, h, N, N)
    lse = torch.logsumexp(s, dim=-1)
    p = torch.exp(s - lse.unsqueeze(-1))
    o_win = (p @ v).transpose(1, 2).reshape(-1, C)  # (B_*N, C)
    o = torch.empty(B * L, C, dtype=qkv.dtype)
    o[idx] = o_win
    return o.reshape(B, L, C), lse


def attention_core_backward(qkv, bias, tau, g: Geometry, d_o, mask: Optional[torch.Tensor] = None,
                            use_shift_mask: bool = True):
    """Hand-derived gradient of ``attention_core_forward`` (checked
    against autograd in tests).  Returns (dqkv, dbias (h,N,N), dtau (h,))."""
    B, L, _ = qkv.shape
    C, h, d, N = g.C, g.heads, g.d, g.N
    idx = torch.from_numpy(window_token_index(B, g.H, g.W, g.ws, g.shift)).reshape(-1)
    win = qkv.reshape(B * L, 3 * C)[idx].reshape(-1, N, 3, h, d).permute(2, 0, 3, 1, 4)
    q, k, v = win[0], win[1], win[2]
    rq = 1.0 / q.norm(dim=-1, keepdim=True).clamp_min(NORM_EPS)
    rk = 1.0 / k.norm(dim=-1, keepdim=True).clamp_min(NORM_EPS)
    qn, kn = q * rq, k * rk
    cos = qn @ kn.transpose(-2, -1)
    s = cos * tau.reshape(1, h, 1, 1) + bias.unsqueeze(0)
    if mask is None and use_shift_mask and g.shift > 0:
        mask = torch.from_numpy(shift_window_mask(g.H, g.W, g.ws, g.shift)).to(s.dtype)
    if mask is not None:
        nW = mask.shape[0]
        s = (s.reshape(-1, nW, h, N, N) + mask.reshape(1, nW, 1, N, N)).reshape(-1, h, N, N)
    p = torch.softmax(s, dim=-1)
    o = p @ v
    do = d_o.reshape(B * L, C)[idx].reshape(-1, N, h, d).permute(0, 2, 1, 3)
    dv = p.transpose(-2, -1) @ do
    dp = do @ v.transpose(-2, -1)
    delta = (do * o).sum(-1, keepdim=True)
    ds = p * (dp - delta)
    dbias = ds.sum(0)
    dtau = (ds * cos).sum(dim=(0, 2, 3))
    t = tau.reshape(1, h, 1, 1)
    dqn = t * (ds @ kn)
    dkn = t * (ds.transpose(-2, -1) @ qn)
    dq = (dqn - qn * (qn * dqn).sum(-1, keepdim=True)) * rq
    dk = (dkn - kn * (kn * dkn).sum(-1, keepdim=True)) * rk
    dwin = torch.stack([dq, dk, dv], dim=0).permute(1, 3, 0, 2, 4).reshape(-1, 3 * C)
    dqkv = torch.empty(B * L, 3 * C, dtype=qkv.dtype)
    dqkv[idx] = dwin
    return dqkv.reshape(B, L, 3 * C), dbias, dtau


def layer_norm_residual(y, shortcut, gamma, beta, keep_scale: Optional[torch.Tensor] = None,
                        eps: float = LN_EPS):
    """shortcut + drop_path(LayerNorm(y)).  swinv2.py:431 / 434; DropPath = per-sample
    scale (B,) equal to bernoulli(keep)/keep (timm semantics), None = identity."""
    z = F.layer_norm(y, (y.shape[-1],), gamma, beta, eps)
    if keep_scale is not None:
        z = z * keep_scale.reshape(-1, 1, 1).to(z.dtype)
    return shortcut + z


def layer_norm_residual_backward(y, gamma, d_out, keep_scale=None, eps: float = LN_EPS):
    """Returns (dy, dgamma, dbeta); d_shortcut is d_out itself."""
    C = y.shape[-1]
    mu = y.mean(-1, keepdim=True)
    var = y.var(-1, unbiased=False, keepdim=True)
    rstd = torch.rsqrt(var + eps)
    xhat = (y - mu) * rstd
    g = d_out if keep_scale is None else d_out * keep_scale.reshape(-1, 1, 1).to(d_out.dtype)
    dgamma = (g * xhat).reshape(-1, C).sum(0)
    dbeta = g.reshape(-1, C).sum(0)
    gx = g * gamma
    dy = rstd * (gx - gx.mean(-1, keepdim=True) - xhat * (gx * xhat).mean(-1, keepdim=True))
    return dy, dgamma, dbeta


def patch_merge_gather(x, H: int, W: int):
    """(B, H*W, C) -> (B, H/2*W/2, 4C): the 2x2 strided concat of swinv2.py:484-491."""
    B, L, C = x.shape
    assert L == H * W, "input feature has wrong size"
    idx = torch.from_numpy(merge_token_index(B, H, W)).reshape(-1)
    return x.reshape(B * L, C)[idx].reshape(B, (H // 2) * (W // 2), 4 * C)


def patch_merging(x, H: int, W: int, w_red, gamma, beta, eps: float = LN_EPS):
    """gather -> Linear(4C->2C, no bias) -> LayerNorm(2C).  swinv2.py:475-496 (SwinV2 order:
    reduce, then norm)."""
    z = F.linear(patch_merge_gather(x, H, W), w_red)
    return F.layer_norm(z, (z.shape[-1],), gamma, beta, eps)


def gelu_mlp(x, w1, b1, w2, b2):
    """fc2(GELU_erf(fc1(x))), dropout p=0.  swinv2.py:60-66."""
    return F.linear(F.gelu(F.linear(x, w1, b1)), w2, b2)


# --------------------------------------------------------------------------------------
# Block / stage / model composition from a reference-layout state_dict
# --------------------------------------------------------------------------------------
def window_attention(x_windows, p: Dict[str, torch.Tensor], prefix: str, heads: int, ws: int,
                     mask: Optional[torch.Tensor] = None, pretrained_ws: int = 0):
    """``WindowAttention.forward(x:(B_,N,C), mask)`` (swinv2.py:204-264) on pre-partitioned
    windows.  Implemented by viewing each window as a ws x ws image with shift 0."""
    B_, N, C = x_windows.shape
    bias_vec = None
    if prefix + "q_bias" in p:
        qb, vb = p[prefix + "q_bias"], p[prefix + "v_bias"]
        bias_vec = torch.cat([qb, torch.zeros_like(vb), vb])  # k has no bias, swinv2.py:213-219
    qkv = F.linear(x_windows, p[prefix + "qkv.weight"], bias_vec)
    coords = relative_coords_table(ws, pretrained_ws)
    table = position_bias_table(coords, p[prefix + "cpb_mlp.0.weight"], p[prefix + "cpb_mlp.0.bias"],
                                p[prefix + "cpb_mlp.2.weight"])
    bias = expand_bias(table, ws).to(qkv.dtype)
    tau = logit_tau(p[prefix + "logit_scale"]).to(qkv.dtype)
    g = Geometry(B_, ws, ws, C, heads, ws, 0)
    o, _ = attention_core_forward(qkv, bias, tau, g, mask=mask, use_shift_mask=False)
    return F.linear(o, p[prefix + "proj.weight"], p[prefix + "proj.bias"])


def swin_block(x, p: Dict[str, torch.Tensor], prefix: str, resolution: Tuple[int, int], heads: int,
               ws: int, shift: int, keep_scale1=None, keep_scale2=None, pretrained_ws: int = 0):
    """``SwinTransformerBlock.forward`` (swinv2.py:390-436), res-post-norm."""
    H, W = resolution
    B, L, C = x.shape
    assert L == H * W, "input feature has wrong size"
    ws, shift = effective_window(resolution, ws, shift)
    a = prefix + "attn."
    bias_vec = None
    if a + "q_bias" in p:
        bias_vec = torch.cat([p[a + "q_bias"], torch.zeros_like(p[a + "v_bias"]), p[a + "v_bias"]])
    qkv = F.linear(x, p[a + "qkv.weight"], bias_vec)
    table = position_bias_table(relative_coords_table(ws, pretrained_ws), p[a + "cpb_mlp.0.weight"],
                                p[a + "cpb_mlp.0.bias"], p[a + "cpb_mlp.2.weight"])
    bias = expand_bias(table, ws).to(qkv.dtype)
    tau = logit_tau(p[a + "logit_scale"]).to(qkv.dtype)
    g = Geometry(B, H, W, C, heads, ws, shift)
    o, _ = attention_core_forward(qkv, bias, tau, g)
    y = F.linear(o, p[a + "proj.weight"], p[a + "proj.bias"])
    x = layer_norm_residual(y, x, p[prefix + "norm1.weight"], p[prefix + "norm1.bias"], keep_scale1)
    m = gelu_mlp(x, p[prefix + "mlp.fc1.weight"], p[prefix + "mlp.fc1.bias"],
                 p[prefix + "mlp.fc2.weight"], p[prefix + "mlp.fc2.bias"])
    return layer_norm_residual(m, x, p[prefix + "norm2.weight"], p[prefix + "norm2.bias"], keep_scale2)


@dataclass(frozen=True)
class ModelSpec:
    """Hyper-parameters the reference takes as ctor kwargs (swinv2.py:699-720)."""
    img_size: int = 256
    patch_size: int = 4
    in_chans: int = 3
    num_classes: object = 10000  # int, or a tuple of tier sizes (swinv2.py:785-795)
    embed_dim: int = 96
    depths: Sequence[int] = (2, 2, 6, 2)
    num_heads: Sequence[int] = (3, 6, 12, 24)
    window_size: int = 8
    pretrained_window_sizes: Sequence[int] = (0, 0, 0, 0)


SWINV2_T = ModelSpec()
SWINV2_B = ModelSpec(embed_dim=128, depths=(2, 2, 18, 2), num_heads=(4, 8, 16, 32), window_size=16,
                     num_classes=(3, 13, 51, 273, 1103, 4884, 10000))


def swin_model(images, p: Dict[str, torch.Tensor], spec: ModelSpec):
    """``SwinTransformerV2.forward`` (swinv2.py:818-845) in eval-style determinism
    (DropPath identity; ape=False; dropout p=0).  Returns logits, or a list per tier."""
    ps = spec.patch_size
    x = F.conv2d(images, p["patch_embed.proj.weight"], p["patch_embed.proj.bias"], stride=ps)
    x = x.flatten(2).transpose(1, 2)  # swinv2.py:654
    x = F.layer_norm(x, (x.shape[-1],), p["patch_embed.norm.weight"], p["patch_embed.norm.bias"], LN_EPS)
    res = spec.img_size // ps
    for li, (depth, heads) in enumerate(zip(spec.depths, spec.num_heads)):
        r = res // (2 ** li)
        for bi in range(depth):
            shift = 0 if bi % 2 == 0 else spec.window_size // 2  # swinv2.py:559
            x = swin_block(x, p, f"layers.{li}.blocks.{bi}.", (r, r), heads, spec.window_size, shift,
                           pretrained_ws=spec.pretrained_window_sizes[li])
        if li < len(spec.depths) - 1:
            d = f"layers.{li}.downsample."
            x = patch_merging(x, r, r, p[d + "reduction.weight"], p[d + "norm.weight"], p[d + "norm.bias"])
    x = F.layer_norm(x, (x.shape[-1],), p["norm.weight"], p["norm.bias"], LN_EPS)
    feat = x.mean(dim=1)  # AdaptiveAvgPool1d(1) over tokens, swinv2.py:834-835
    if isinstance(spec.num_classes, int):
        return F.linear(feat, p["head.weight"], p["head.bias"])
    return [F.linear(feat, p[f"head.heads.{t}.weight"], p[f"head.heads.{t}.bias"])
            for t in range(len(spec.num_classes))]


MULTITASK_COEFFS = (8.0, 5.65, 4.0, 2.82, 2.0, 1.41, 1.0)


def multitask_cross_entropy(logits: List[torch.Tensor], targets: torch.Tensor,
                            coeffs: Sequence[float] = MULTITASK_COEFFS):
    """dot(coeffs, [CE(logits_t, targets[:, t])]) -- hierarchy.py:65-94 semantics."""
    losses = torch.stack([F.cross_entropy(lg.float(), targets[:, t]) for t, lg in enumerate(logits)])
    return (losses * torch.tensor(coeffs[: len(logits)], dtype=losses.dtype)).sum()


def init_state(spec: ModelSpec, seed: int = 0, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Random-init weights with the reference's shapes and init scheme (trunc_normal 0.02 for
    Linear, swinv2.py:801-808) but with every block's LayerNorm gamma ~ N(1,0.1), beta ~ N(0,0.1):
    the reference zero-inits them (swinv2.py:603-608) which makes every block the identity
    and all attention gradients exactly zero."""
    g = torch.Generator().manual_seed(seed)

    def tn(*shape):
        return torch.nn.init.trunc_normal_(torch.empty(*shape, dtype=dtype), std=0.02, generator=g)

    def ln(prefix, n, out):
        out[prefix + "weight"] = (1.0 + 0.1 * torch.randn(n, generator=g)).to(dtype)
        out[prefix + "bias"] = (0.1 * torch.randn(n, generator=g)).to(dtype)

    p: Dict[str, torch.Tensor] = {}
    E = spec.embed_dim
    fan_in = spec.in_chans * spec.patch_size ** 2
    bound = 1.0 / math.sqrt(fan_in)
    p["patch_embed.proj.weight"] = (torch.rand(E, spec.in_chans, spec.patch_size, spec.patch_size,
                                               generator=g) * 2 - 1).mul(bound).to(dtype)
    p["patch_embed.proj.bias"] = (torch.rand(E, generator=g) * 2 - 1).mul(bound).to(dtype)
    ln("patch_embed.norm.", E, p)
    for li, (depth, heads) in enumerate(zip(spec.depths, spec.num_heads)):
        C = E * 2 ** li
        for bi in range(depth):
            b = f"layers.{li}.blocks.{bi}."
            a = b + "attn."
            p[a + "logit_scale"] = torch.empty(heads, 1, 1, dtype=dtype).uniform_(
                math.log(5.0), math.log(50.0), generator=g)
            p[a + "cpb_mlp.0.weight"] = tn(512, 2)
            p[a + "cpb_mlp.0.bias"] = torch.zeros(512, dtype=dtype)
            p[a + "cpb_mlp.2.weight"] = tn(heads, 512)
            p[a + "qkv.weight"] = tn(3 * C, C)
            p[a + "q_bias"] = (0.02 * torch.randn(C, generator=g)).to(dtype)
            p[a + "v_bias"] = (0.02 * torch.randn(C, generator=g)).to(dtype)
            p[a + "proj.weight"] = tn(C, C)
            p[a + "proj.bias"] = torch.zeros(C, dtype=dtype)
            ln(b + "norm1.", C, p)
            ln(b + "norm2.", C, p)
            p[b + "mlp.fc1.weight"] = tn(4 * C, C)
            p[b + "mlp.fc1.bias"] = torch.zeros(4 * C, dtype=dtype)
            p[b + "mlp.fc2.weight"] = tn(C, 4 * C)
            p[b + "mlp.fc2.bias"] = torch.zeros(C, dtype=dtype)
        if li < len(spec.depths) - 1:
            dn = f"layers.{li}.downsample."
            p[dn + "reduction.weight"] = tn(2 * C, 4 * C)
            ln(dn + "norm.", 2 * C, p)
    F_out = E * 2 ** (len(spec.depths) - 1)
    ln("norm.", F_out, p)
    if isinstance(spec.num_classes, int):
        p["head.weight"] = tn(spec.num_classes, F_out)
        p["head.bias"] = torch.zeros(spec.num_classes, dtype=dtype)
    else:
        for t, n in enumerate(spec.num_classes):
            p[f"head.heads.{t}.weight"] = tn(n, F_out)
            p[f"head.heads.{t}.bias"] = torch.zeros(n, dtype=dtype)
    return p
