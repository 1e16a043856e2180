"""Drop-in module surface for the reference's ``swinv2.py`` (samuelstevens/hierarchical-vision).

Same class names, constructor arguments, forward signatures, parameter / buffer names and
``state_dict`` keys as the reference (citations are ``/root/reference/swinv2.py:line``), so a
model built from these classes loads reference checkpoints with ``strict=True`` and can be
handed to the reference's ``models.py`` / Composer loop.  What differs is
*how* a block runs: the roll / window partition / reverse copies, the whole scaled-cosine
attention, the res-post-norm LayerNorm+residual and PatchMerging's gather+LayerNorm execute
as hand-written sm_100a kernels through the C ABI in ``include/hv_swin.h``.  The per-token
``nn.Linear`` layers (qkv, proj, MLP, reduction, heads) and the patch-embedding conv stay
library GEMMs.  There is no CPU path: calling ``forward`` on CPU tensors raises.
"""
from __future__ import annotations

import collections.abc
import dataclasses
import math
import re
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F
import torch.utils.checkpoint as checkpoint

from . import functional as hvf

__all__ = ["set_amp_residual_dtype", "DropPath", "to_2tuple", "trunc_normal_", "MultitaskHead", "Mlp", "window_partition", "window_reverse",
           "WindowAttention", "SwinTransformerBlock", "PatchMerging", "BasicLayer", "PatchEmbed",
           "SwinTransformerV2", "Checkpoint", "swinv2_tiny", "swinv2_base"]

trunc_normal_ = nn.init.trunc_normal_  # the reference takes it from timm (swinv2.py:9)

# dtype of the residual stream under torch.autocast.  The reference's AMP run keeps it in float32: autocast executes
# layer_norm in fp32, so `x = shortcut + drop_path(norm(...))` (swinv2.py:431, 434, 494, 656) never leaves fp32 and only
# the Linear inputs are cast down.  None (default) = the stream follows the autocast dtype (bf16): the LayerNorm +
# residual kernels then move 6 instead of 10 bytes per element and the casts in front of every Linear disappear, at the
# price of one bf16 rounding per residual add -- a documented deviation from reference AMP, inside the bf16 tolerance of
# the parity tests (tests/test_gpu_parity.py::test_swinv2_tiny_bf16_autocast_vs_oracle runs both settings).
# torch.float32 = reference AMP semantics.  Outside autocast the stream always has the input's dtype.
AMP_RESIDUAL_DTYPE: Optional[torch.dtype] = None


def set_amp_residual_dtype(dtype: Optional[torch.dtype]) -> None:
    """None: bf16 residual stream under autocast (fast); torch.float32: fp32 stream as in the reference's AMP run."""
    global AMP_RESIDUAL_DTYPE
    if dtype not in (None, torch.float32, torch.bfloat16):
        raise ValueError(f"residual stream dtype must be None, float32 or bfloat16, not {dtype}")
    AMP_RESIDUAL_DTYPE = dtype


def _linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """F.linear; under CUDA autocast with an fp32 master weight the GEMMs read the weight's persistent low-precision
    shadow and write the weight gradient in fp32 (functional.shadow_linear) instead of casting per call."""
    if x.is_cuda and torch.is_autocast_enabled("cuda") and weight.dtype == torch.float32 and weight.is_cuda:
        dt = torch.get_autocast_dtype("cuda")
        return hvf.shadow_linear(x if x.dtype == dt else x.to(dt), weight, bias)
    return F.linear(x, weight, bias)


def _stream_dtype(y: torch.Tensor) -> Optional[torch.dtype]:
    """Output dtype of a LayerNorm that starts (a stage of) the residual stream."""
    if y.is_cuda and torch.is_autocast_enabled("cuda") and AMP_RESIDUAL_DTYPE is not None:
        return AMP_RESIDUAL_DTYPE
    return None


def to_2tuple(v):
    if isinstance(v, collections.abc.Iterable) and not isinstance(v, str):
        return tuple(v)
    return (v, v)


def _keep_scale(x: torch.Tensor, drop_prob: float, training: bool) -> Optional[torch.Tensor]:
    """Per-sample stochastic-depth factor (B,) = bernoulli(keep)/keep, or None when inactive."""
    if drop_prob == 0.0 or not training:
        return None
    keep = 1.0 - drop_prob
    s = torch.empty((x.shape[0],), dtype=torch.float32, device=x.device).bernoulli_(keep)
    if keep > 0.0:
        s.div_(keep)
    return s


class DropPath(nn.Module):
    """Stochastic depth per sample (timm semantics, used at swinv2.py:347)."""

    def __init__(self, drop_prob: float = 0.0):
        super().__init__()
        self.drop_prob = float(drop_prob)

    def forward(self, x):
        s = _keep_scale(x, self.drop_prob, self.training)
        if s is None:
            return x
        return x * s.to(x.dtype).view((-1,) + (1,) * (x.dim() - 1))

    def extra_repr(self):
        return f"drop_prob={self.drop_prob:.3f}"


class MultitaskHead(nn.Module):
    """One linear classifier per taxonomy tier; forward returns a list (swinv2.py:12-40)."""

    def __init__(self, num_features, num_classes):
        super().__init__()
        self.num_classes = tuple(num_classes)
        assert all(n > 0 for n in self.num_classes)
        self.heads = nn.ModuleList(nn.Linear(num_features, n) for n in self.num_classes)

    def forward(self, x):
        return [_linear(x, head.weight, head.bias) for head in self.heads]


class Mlp(nn.Module):
    """fc1 -> act -> drop -> fc2 -> drop (swinv2.py:43-66)."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.0):
        super().__init__()
        self.fc1 = nn.Linear(in_features, hidden_features or in_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features or in_features, out_features or in_features)
        self.drop = nn.Dropout(drop)

    def _hidden(self, x, with_shortcut=False):
        """act(fc1(x)): for the reference's default nn.GELU (exact erf form) the fc1 bias add and the
        activation run as one kernel whose backward also yields d fc1.bias.  ``with_shortcut``: also return x for
        the caller's residual connection (its gradient is then accumulated inside the fc1 dx GEMM)."""
        sc = x
        if type(self.act) is nn.GELU and getattr(self.act, "approximate", "none") == "none" and self.fc1.bias is not None:
            if with_shortcut and x.is_cuda and torch.is_autocast_enabled("cuda") and x.dtype == torch.get_autocast_dtype("cuda"):
                h, sc = hvf.linear_shortcut(x, self.fc1.weight)  # the node reads the weight's bf16 shadow
            elif with_shortcut and x.is_cuda and not torch.is_autocast_enabled("cuda") and x.dtype == self.fc1.weight.dtype:
                h, sc = hvf.linear_shortcut(x, self.fc1.weight)
            else:
                h = _linear(x, self.fc1.weight)
            if hvf.bias_gelu_supported(h):
                a = hvf.bias_gelu(h, self.fc1.bias)
            else:
                a = self.act(h + self.fc1.bias)
        else:
            a = self.act(self.fc1(x))
        return (a, sc) if with_shortcut else a

    def _hidden_fc2(self, x, with_shortcut=False):
        """fc2(act(fc1(x))) WITHOUT the fc2 bias (the caller adds it or folds it into its LayerNorm kernel), and x for the
        residual connection.  bf16 activations with the exact GELU take one autograd node for the activation and the fc2
        GEMM, whose backward is the fused dGELU GEMM (functional._GeluFc2); everything else the two separate nodes."""
        sc = x
        fused = (type(self.act) is nn.GELU and getattr(self.act, "approximate", "none") == "none" and self.fc1.bias is not None
                 and self.drop.p == 0.0 and x.is_cuda)
        if fused and x.dtype == torch.bfloat16 and self.fc1.out_features % 128 == 0 \
                and self.fc2.in_features == self.fc1.out_features and (
                    (torch.is_autocast_enabled("cuda") and torch.get_autocast_dtype("cuda") == torch.bfloat16)
                    or self.fc1.weight.dtype == torch.bfloat16):
            # bf16 activations: the whole Mlp as one autograd node (fused fc1 + GELU GEMM, fused dGELU GEMM in the backward)
            return hvf.mlp_fused(x, self.fc1.weight, self.fc1.bias, self.fc2.weight)
        if fused:
            if with_shortcut and torch.is_autocast_enabled("cuda") and x.dtype == torch.get_autocast_dtype("cuda"):
                h, sc = hvf.linear_shortcut(x, self.fc1.weight)
            elif with_shortcut and not torch.is_autocast_enabled("cuda") and x.dtype == self.fc1.weight.dtype:
                h, sc = hvf.linear_shortcut(x, self.fc1.weight)
            else:
                h = _linear(x, self.fc1.weight)
            if h.dtype == torch.bfloat16 and hvf.bias_gelu_supported(h):
                return hvf.gelu_fc2(h, self.fc1.bias, self.fc2.weight), sc
            a = hvf.bias_gelu(h, self.fc1.bias) if hvf.bias_gelu_supported(h) else self.act(h + self.fc1.bias)
            return _linear(a, self.fc2.weight), sc
        a = self._hidden(x)
        return _linear(self.drop(a), self.fc2.weight), sc

    def forward(self, x):
        if self.drop.p == 0.0 and x.is_cuda and self.fc2.bias is not None:
            m, _ = self._hidden_fc2(x)
            return m + self.fc2.bias.to(m.dtype)
        return self.drop(self.fc2(self.drop(self._hidden(x))))


def window_partition(x, window_size):
    """(B, H, W, C) -> (B*nW, ws, ws, C) (swinv2.py:69-83).  Kept for API parity; the blocks
    below never materialise this tensor."""
    B, H, W, C = x.shape
    ws = window_size
    return x.reshape(B, H // ws, ws, W // ws, ws, C).transpose(2, 3).reshape(-1, ws, ws, C)


def window_reverse(windows, window_size, H, W):
    """(B*nW, ws, ws, C) -> (B, H, W, C) (swinv2.py:86-102)."""
    ws = window_size
    B = int(windows.shape[0] / (H * W / ws / ws))
    return windows.reshape(B, H // ws, W // ws, ws, ws, -1).transpose(2, 3).reshape(B, H, W, -1)


def _relative_coords_table(ws_h: int, ws_w: int, pre_h: int, pre_w: int) -> torch.Tensor:
    """(1, 2Wh-1, 2Ww-1, 2) float32 buffer, bit-identical to swinv2.py:147-171."""
    rh = torch.arange(-(ws_h - 1), ws_h, dtype=torch.float32)
    rw = torch.arange(-(ws_w - 1), ws_w, dtype=torch.float32)
    t = torch.stack([rh.view(-1, 1).expand(-1, rw.numel()), rw.view(1, -1).expand(rh.numel(), -1)], dim=-1)
    t = t.contiguous().unsqueeze(0)
    t[..., 0] /= (pre_h - 1) if pre_h > 0 else (ws_h - 1)
    t[..., 1] /= (pre_w - 1) if pre_h > 0 else (ws_w - 1)
    t *= 8
    return torch.sign(t) * torch.log2(torch.abs(t) + 1.0) / math.log2(8)


def _relative_position_index(ws_h: int, ws_w: int) -> torch.Tensor:
    """(N, N) int64 buffer (swinv2.py:175-190): (dh + Wh-1) * (2Ww-1) + (dw + Ww-1)."""
    idx = torch.arange(ws_h * ws_w)
    ih, iw = idx // ws_w, idx % ws_w
    dh = ih.view(-1, 1) - ih.view(1, -1) + ws_h - 1
    dw = iw.view(-1, 1) - iw.view(1, -1) + ws_w - 1
    return (dh * (2 * ws_w - 1) + dw).to(torch.int64)


def _shift_mask(H: int, W: int, ws: int, shift: int) -> Optional[torch.Tensor]:
    """(nW, N, N) float32 {0,-100} buffer of swinv2.py:357-388, or None when shift == 0."""
    if shift <= 0:
        return None

    def band(n):
        p = torch.arange(n)
        return (p >= n - ws).long() + (p >= n - shift).long()

    ident = 3 * band(H).view(H, 1) + band(W).view(1, W)
    ident = ident.reshape(H // ws, ws, W // ws, ws).transpose(1, 2).reshape(-1, ws * ws)
    differ = ident.unsqueeze(1) != ident.unsqueeze(2)
    return torch.where(differ, torch.tensor(-100.0), torch.tensor(0.0))


class WindowAttention(nn.Module):
    r"""Window multi-head self attention with scaled-cosine logits and continuous relative
    position bias (swinv2.py:105-283).

    Args:
        dim (int): channels.  window_size (tuple[int]): (Wh, Ww).  num_heads (int).
        qkv_bias (bool): learnable q/v bias (k has none, swinv2.py:213-219).
        attn_drop, proj_drop (float).  pretrained_window_size (tuple[int]).
    """

    def __init__(self, dim, window_size, num_heads, qkv_bias=True, attn_drop=0.0, proj_drop=0.0,
                 pretrained_window_size=[0, 0]):
        super().__init__()
        self.dim = dim
        self.window_size = window_size
        self.pretrained_window_size = pretrained_window_size
        self.num_heads = num_heads

        self.logit_scale = nn.Parameter(torch.log(10 * torch.ones((num_heads, 1, 1))), requires_grad=True)
        self.register_buffer("logit_clamp_max", torch.log(torch.tensor(1.0 / 0.01)))
        self.cpb_mlp = nn.Sequential(nn.Linear(2, 512, bias=True), nn.ReLU(inplace=True),
                                     nn.Linear(512, num_heads, bias=False))
        self.register_buffer("relative_coords_table", _relative_coords_table(
            window_size[0], window_size[1], pretrained_window_size[0], pretrained_window_size[1]))
        self.register_buffer("relative_position_index", _relative_position_index(window_size[0], window_size[1]))

        self.qkv = nn.Linear(dim, dim * 3, bias=False)
        if qkv_bias:
            self.q_bias = nn.Parameter(torch.zeros(dim))
            self.v_bias = nn.Parameter(torch.zeros(dim))
        else:
            self.q_bias = None
            self.v_bias = None
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)
        self.softmax = nn.Softmax(dim=-1)

    # -- pieces shared with SwinTransformerBlock (which feeds un-partitioned tokens) ----------
    def _qkv(self, x):
        bias = None
        if self.q_bias is not None:
            bias = torch.cat((self.q_bias, torch.zeros_like(self.v_bias, requires_grad=False), self.v_bias))
        return F.linear(x, self.qkv.weight, bias)

    def _bias_table(self):
        """((2Wh-1)(2Ww-1), heads): 16*sigmoid(cpb_mlp(coords)); sigmoid commutes with the gather
        of swinv2.py:236-246, which the kernel performs from the closed-form index."""
        # 225..961 rows x (2 -> 512 -> heads): microscopic, so it is kept out of autocast (fp32 weights
        # stay fp32); the reference lets autocast run it in bf16, which costs it ~5% on these gradients.
        l1, l2 = self.cpb_mlp[0], self.cpb_mlp[2]
        if (self.relative_coords_table.is_cuda and l1.weight.dtype == torch.float32
                and hvf.cpb_bias_supported(l1.weight, l2.weight)):
            return hvf.cpb_bias(self.relative_coords_table, l1.weight, l1.bias, l2.weight)  # one kernel each way
        with torch.autocast(device_type=self.relative_coords_table.device.type, enabled=False):
            table = self.cpb_mlp(self.relative_coords_table.to(self.cpb_mlp[0].weight.dtype))
        return 16 * torch.sigmoid(table.view(-1, self.num_heads).float())

    def _tau(self):
        return torch.clamp(self.logit_scale.float(), max=self.logit_clamp_max.float()).exp().reshape(-1)

    def _fused(self, x_tokens, B, H, W, shift, mask, with_proj_bias=True):
        """Returns (y, bias, shortcut): the attention branch output; with ``with_proj_bias=False`` the proj bias the caller
        folds into its LayerNorm kernel (else None, already added); and the tensor the caller must use for the residual
        shortcut -- ``x_tokens`` routed through the attention node when that node can absorb the shortcut's gradient into
        its dx GEMM, else ``x_tokens`` itself."""
        if self.window_size[0] != self.window_size[1]:
            raise NotImplementedError("fused window attention needs square windows (the reference's "
                                      "SwinTransformerBlock only ever builds square ones, swinv2.py:339)")
        if self.attn_drop.p > 0.0 and self.training:
            raise NotImplementedError("attention dropout is not fused; every reference config uses attn_drop=0")
        ws = self.window_size[0]
        shortcut = x_tokens
        table, tau = self._bias_table(), self._tau()
        dt = torch.get_autocast_dtype("cuda") if (x_tokens.is_cuda and torch.is_autocast_enabled("cuda")) else x_tokens.dtype
        v_bias = None
        if mask is None and x_tokens.is_cuda and hvf.window_attention_kind(self.dim, self.num_heads, ws, dt) == 1:
            # tensor-core kernel: qkv Linear + attention are one autograd node (q_bias gradient from the kernel);
            # v_bias leaves the attention as a plain additive term because softmax rows sum to one
            o, sc = hvf.qkv_window_attention(x_tokens.to(dt), self.qkv.weight, self.q_bias, table, tau, B=B, H=H,
                                             W=W, C=self.dim, heads=self.num_heads, ws=ws, shift=shift)
            if sc.dtype == x_tokens.dtype:
                shortcut = sc  # same values as x_tokens; its gradient is accumulated inside the node's dx GEMM
            v_bias = self.v_bias
        else:
            o = hvf.window_attention(self._qkv(x_tokens), table, tau, B=B, H=H, W=W, C=self.dim,
                                     heads=self.num_heads, ws=ws, shift=shift, mask=mask)
        if with_proj_bias:
            if v_bias is not None:
                o = o + v_bias.to(o.dtype)
            return self.proj_drop(_linear(o, self.proj.weight, self.proj.bias)), None, shortcut
        # the caller folds the bias into its LayerNorm kernel: proj(o + v_bias) = o W^T + (W v_bias + proj.bias)
        bias = self.proj.bias
        if v_bias is not None:
            with torch.autocast(device_type="cuda", enabled=False):
                bias = bias + F.linear(v_bias.float(), self.proj.weight.float())
        return _linear(o, self.proj.weight), bias, shortcut

    def forward(self, x, mask=None):
        """x: (num_windows*B, N, C); mask: (num_windows, N, N) of 0/-100 or None (swinv2.py:204-264)."""
        B_, N, C = x.shape
        ws = self.window_size[0]
        assert N == self.window_size[0] * self.window_size[1], "input feature has wrong size"
        # a pre-partitioned window is a ws x ws image with no shift
        return self._fused(x, B_, ws, ws, 0, mask)[0]

    def extra_repr(self) -> str:
        return (f"dim={self.dim}, window_size={self.window_size}, "
                f"pretrained_window_size={self.pretrained_window_size}, num_heads={self.num_heads}")

    def flops(self, N):
        head_dim = self.dim // self.num_heads
        return N * self.dim * 3 * self.dim + 2 * self.num_heads * N * N * head_dim + N * self.dim * self.dim


class SwinTransformerBlock(nn.Module):
    r"""SwinV2 block, res-post-norm (swinv2.py:286-456).  Same arguments as the reference."""

    def __init__(self, dim, input_resolution, num_heads, window_size=7, shift_size=0, mlp_ratio=4.0, qkv_bias=True,
                 drop=0.0, attn_drop=0.0, drop_path=0.0, act_layer=nn.GELU, norm_layer=nn.LayerNorm,
                 pretrained_window_size=0):
        super().__init__()
        self.dim = dim
        self.input_resolution = input_resolution
        self.num_heads = num_heads
        self.window_size = window_size
        self.shift_size = shift_size
        self.mlp_ratio = mlp_ratio
        if min(self.input_resolution) <= self.window_size:  # swinv2.py:328-331
            self.shift_size = 0
            self.window_size = min(self.input_resolution)
        assert 0 <= self.shift_size < self.window_size, "shift_size must in 0-window_size"

        self.norm1 = norm_layer(dim)
        self.attn = WindowAttention(dim, window_size=to_2tuple(self.window_size), num_heads=num_heads,
                                    qkv_bias=qkv_bias, attn_drop=attn_drop, proj_drop=drop,
                                    pretrained_window_size=to_2tuple(pretrained_window_size))
        self.drop_path = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)
        H, W = self.input_resolution
        # persistent buffer in the reference (swinv2.py:388); kept for state_dict parity.  The kernel
        # regenerates the same pattern from (H, W, ws, shift) and never reads this tensor.
        self.register_buffer("attn_mask", _shift_mask(H, W, self.window_size, self.shift_size))

    @staticmethod
    def _fusable(norm):
        return type(norm) is nn.LayerNorm and norm.elementwise_affine and norm.bias is not None

    def _post_norm(self, norm, branch, shortcut, bias=None):
        p = self.drop_path.drop_prob if isinstance(self.drop_path, DropPath) else 0.0
        if self._fusable(norm):
            return hvf.ln_residual(branch, shortcut, norm.weight, norm.bias, _keep_scale(branch, p, self.training),
                                   norm.eps, bias=bias)
        if bias is not None:
            branch = branch + bias
        return shortcut + self.drop_path(norm(branch))  # non-LayerNorm norm_layer: library ops

    def forward(self, x):
        H, W = self.input_resolution
        B, L, C = x.shape
        assert L == H * W, "input feature has wrong size"
        # The biases of the two Linears that feed a LayerNorm (attn.proj, mlp.fc2) are added inside the
        # LayerNorm kernel (dropout p = 0 in between, as in every reference config), so their gradients
        # come out of its backward instead of two extra full-tensor column reductions.
        fold1 = self._fusable(self.norm1) and self.attn.proj_drop.p == 0.0
        fold2 = self._fusable(self.norm2) and self.mlp.drop.p == 0.0 and isinstance(self.mlp, Mlp)
        # roll + partition + attention + reverse + roll back: one kernel, no rolled / partitioned copy
        y, proj_bias, x = self.attn._fused(x, B, H, W, self.shift_size, None, with_proj_bias=not fold1)
        x = self._post_norm(self.norm1, y, x, proj_bias)                                       # swinv2.py:431
        if fold2:
            m, x = self.mlp._hidden_fc2(x, with_shortcut=True)
            return self._post_norm(self.norm2, m, x, self.mlp.fc2.bias)                         # swinv2.py:434
        return self._post_norm(self.norm2, self.mlp(x), x)

    def extra_repr(self) -> str:
        return (f"dim={self.dim}, input_resolution={self.input_resolution}, num_heads={self.num_heads}, "
                f"window_size={self.window_size}, shift_size={self.shift_size}, mlp_ratio={self.mlp_ratio}")

    def flops(self):
        H, W = self.input_resolution
        nW = H * W / self.window_size / self.window_size
        return (2 * self.dim * H * W + nW * self.attn.flops(self.window_size * self.window_size)
                + 2 * H * W * self.dim * self.dim * self.mlp_ratio)


class PatchMerging(nn.Module):
    r"""2x2 patch merging: gather -> Linear(4C, 2C, no bias) -> LayerNorm(2C) (swinv2.py:459-505)."""

    def __init__(self, input_resolution, dim, norm_layer=nn.LayerNorm):
        super().__init__()
        self.input_resolution = input_resolution
        self.dim = dim
        self.reduction = nn.Linear(4 * dim, 2 * dim, bias=False)
        self.norm = norm_layer(2 * dim)

    def forward(self, x):
        H, W = self.input_resolution
        B, L, C = x.shape
        assert L == H * W, "input feature has wrong size"
        assert H % 2 == 0 and W % 2 == 0, f"x size ({H}*{W}) are not even."
        x = _linear(hvf.patch_merge_gather(x, H, W), self.reduction.weight)
        if SwinTransformerBlock._fusable(self.norm):
            return hvf.ln_residual(x, None, self.norm.weight, self.norm.bias, None, self.norm.eps, out_dtype=_stream_dtype(x))
        return self.norm(x)

    def extra_repr(self) -> str:
        return f"input_resolution={self.input_resolution}, dim={self.dim}"

    def flops(self):
        H, W = self.input_resolution
        return (H // 2) * (W // 2) * 4 * self.dim * 2 * self.dim + H * W * self.dim // 2


class BasicLayer(nn.Module):
    """One stage: ``depth`` blocks alternating shift 0 / ws//2, then optional downsample
    (swinv2.py:508-608)."""

    def __init__(self, dim, input_resolution, depth, num_heads, window_size, mlp_ratio=4.0, qkv_bias=True, drop=0.0,
                 attn_drop=0.0, drop_path=0.0, norm_layer=nn.LayerNorm, downsample=None, use_checkpoint=False,
                 pretrained_window_size=0):
        super().__init__()
        self.dim = dim
        self.input_resolution = input_resolution
        self.depth = depth
        self.use_checkpoint = use_checkpoint
        self.blocks = nn.ModuleList(
            SwinTransformerBlock(dim=dim, input_resolution=input_resolution, num_heads=num_heads,
                                 window_size=window_size, shift_size=0 if i % 2 == 0 else window_size // 2,
                                 mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, drop=drop, attn_drop=attn_drop,
                                 drop_path=drop_path[i] if isinstance(drop_path, list) else drop_path,
                                 norm_layer=norm_layer, pretrained_window_size=pretrained_window_size)
            for i in range(depth))
        self.downsample = downsample(input_resolution, dim=dim, norm_layer=norm_layer) if downsample is not None else None

    def forward(self, x):
        for blk in self.blocks:
            x = checkpoint.checkpoint(blk, x, use_reentrant=False) if self.use_checkpoint else blk(x)
        if self.downsample is not None:
            x = self.downsample(x)
        return x

    def extra_repr(self) -> str:
        return f"dim={self.dim}, input_resolution={self.input_resolution}, depth={self.depth}"

    def flops(self):
        total = sum(blk.flops() for blk in self.blocks)
        return total + (self.downsample.flops() if self.downsample is not None else 0)

    def _init_respostnorm(self):
        for blk in self.blocks:  # swinv2.py:603-608
            for n in (blk.norm1, blk.norm2):
                nn.init.constant_(n.bias, 0)
                nn.init.constant_(n.weight, 0)


class _PatchProj(torch.autograd.Function):
    """rows @ W^T (+ bias) of the patch embedding for an fp32 conv weight whose bf16 shadow ``w2d`` (E, 48) does the
    GEMM; the weight gradient is written in fp32 in the conv weight's (E, 3, 4, 4) shape."""

    @staticmethod
    def forward(ctx, rows, weight, w2d, bias):
        ctx.save_for_backward(rows)
        ctx.wshape, ctx.bdtype = weight.shape, (bias.dtype if bias is not None else None)
        return F.linear(rows, w2d, None if bias is None else bias.to(rows.dtype))

    @staticmethod
    def backward(ctx, dy):
        (rows,) = ctx.saved_tensors
        d2 = dy.reshape(-1, dy.shape[-1]).to(rows.dtype)
        dw = hvf._weight_grad(d2.t(), rows, torch.float32).view(ctx.wshape) if ctx.needs_input_grad[1] else None
        db = d2.sum(dim=0, dtype=torch.float32).to(ctx.bdtype) if (ctx.bdtype is not None and ctx.needs_input_grad[3]) else None
        return None, dw, None, db


class PatchEmbed(nn.Module):
    """Conv2d(k=s=patch) patch embedding + optional norm (swinv2.py:611-670)."""

    def __init__(self, img_size=224, patch_size=4, in_chans=3, embed_dim=96, norm_layer=None):
        super().__init__()
        self.img_size = to_2tuple(img_size)
        self.patch_size = to_2tuple(patch_size)
        self.patches_resolution = [self.img_size[0] // self.patch_size[0], self.img_size[1] // self.patch_size[1]]
        self.num_patches = self.patches_resolution[0] * self.patches_resolution[1]
        self.in_chans = in_chans
        self.embed_dim = embed_dim
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=self.patch_size, stride=self.patch_size)
        self.norm = norm_layer(embed_dim) if norm_layer is not None else None

    def forward(self, x, input_norm=None):
        """x: (B, C, H, W) float image (reference swinv2.py:648-657).  Extension: a uint8 image together with
        ``input_norm = (mean, std)`` per channel in pixel units, in which case the reference's on-device
        normalisation (data.py:130-136) is folded into the patch gather."""
        B, C, H, W = x.shape
        assert H == self.img_size[0] and W == self.img_size[1], \
            f"Input image size ({H}*{W}) doesn't match model ({self.img_size[0]}*{self.img_size[1]})."
        fast = (x.is_cuda and C == 3 and tuple(self.patch_size) == (4, 4) and not x.requires_grad
                and x.dtype in (torch.uint8, torch.float32, torch.bfloat16))
        if not fast:
            if x.dtype == torch.uint8:
                mean, std = input_norm
                x = (x.float() - mean.view(1, -1, 1, 1)) / std.view(1, -1, 1, 1)
            x = self.proj(x).flatten(2).transpose(1, 2)
            bias = None
        else:
            # Conv2d(k = s = 4) == per-patch GEMM: gather kernel (normalisation folded in) + library GEMM; replaces
            # cuDNN's NCHW<->NHWC transposes around the convolution
            scale = shift = None
            if input_norm is not None:
                mean, std = input_norm
                scale, shift = 1.0 / std.float(), -mean.float() / std.float()
            if torch.is_autocast_enabled("cuda"):
                dt = torch.get_autocast_dtype("cuda")
            else:
                dt = x.dtype if x.is_floating_point() else self.proj.weight.dtype
            if dt not in (torch.float32, torch.bfloat16):
                dt = torch.float32
            rows = hvf.patch_rows(x, scale, shift, dt)
            fold = self.norm is not None and SwinTransformerBlock._fusable(self.norm)
            pbias = None if (fold or self.proj.bias is None) else self.proj.bias
            if self.proj.weight.dtype == torch.float32 and dt != torch.float32:
                w = hvf.weight_shadow(self.proj.weight, dt).view(self.embed_dim, -1)  # images carry no gradient: no dx
                x = _PatchProj.apply(rows, self.proj.weight, w, pbias)
            else:
                x = F.linear(rows, self.proj.weight.view(self.embed_dim, -1).to(dt), None if pbias is None else pbias.to(dt))
            x = x.view(B, -1, self.embed_dim)
            bias = self.proj.bias if fold else None
        if self.norm is not None:
            if SwinTransformerBlock._fusable(self.norm) and x.is_cuda and x.dtype in (torch.float32, torch.bfloat16):
                x = hvf.ln_residual(x, None, self.norm.weight, self.norm.bias, None, self.norm.eps, bias=bias,
                                    out_dtype=_stream_dtype(x))
            else:
                x = self.norm(x)
        return x

    def flops(self):
        Ho, Wo = self.patches_resolution
        f = Ho * Wo * self.embed_dim * self.in_chans * (self.patch_size[0] * self.patch_size[1])
        return f + (Ho * Wo * self.embed_dim if self.norm is not None else 0)


class SwinTransformerV2(nn.Module):
    r"""SwinV2 backbone + head, same ctor arguments as swinv2.py:699-720 (``num_classes`` may be a
    tuple of taxonomy tier sizes -> ``MultitaskHead``, swinv2.py:785-795)."""

    def __init__(self, img_size=224, patch_size=4, in_chans=3, num_classes=1000, embed_dim=96, depths=[2, 2, 6, 2],
                 num_heads=[3, 6, 12, 24], window_size=7, mlp_ratio=4.0, qkv_bias=True, drop_rate=0.0,
                 attn_drop_rate=0.0, drop_path_rate=0.1, norm_layer=nn.LayerNorm, ape=False, patch_norm=True,
                 use_checkpoint=False, pretrained_window_sizes=[0, 0, 0, 0], **kwargs):
        super().__init__()
        self.num_classes = num_classes
        self.num_layers = len(depths)
        self.embed_dim = embed_dim
        self.ape = ape
        self.patch_norm = patch_norm
        self.num_features = int(embed_dim * 2 ** (self.num_layers - 1))
        self.mlp_ratio = mlp_ratio

        self.patch_embed = PatchEmbed(img_size=img_size, patch_size=patch_size, in_chans=in_chans, embed_dim=embed_dim,
                                      norm_layer=norm_layer if self.patch_norm else None)
        self.patches_resolution = self.patch_embed.patches_resolution
        if self.ape:
            self.absolute_pos_embed = nn.Parameter(torch.zeros(1, self.patch_embed.num_patches, embed_dim))
            trunc_normal_(self.absolute_pos_embed, std=0.02)
        self.pos_drop = nn.Dropout(p=drop_rate)

        dpr = [v.item() for v in torch.linspace(0, drop_path_rate, sum(depths))]  # swinv2.py:753-755
        self.layers = nn.ModuleList()
        for i in range(self.num_layers):
            self.layers.append(BasicLayer(
                dim=int(embed_dim * 2 ** i),
                input_resolution=(self.patches_resolution[0] // 2 ** i, self.patches_resolution[1] // 2 ** i),
                depth=depths[i], num_heads=num_heads[i], window_size=window_size, mlp_ratio=self.mlp_ratio,
                qkv_bias=qkv_bias, drop=drop_rate, attn_drop=attn_drop_rate,
                drop_path=dpr[sum(depths[:i]):sum(depths[:i + 1])], norm_layer=norm_layer,
                downsample=PatchMerging if i < self.num_layers - 1 else None, use_checkpoint=use_checkpoint,
                pretrained_window_size=pretrained_window_sizes[i]))

        self.norm = norm_layer(self.num_features)
        self.avgpool = nn.AdaptiveAvgPool1d(1)
        if isinstance(num_classes, int):
            self.head = nn.Linear(self.num_features, num_classes) if num_classes > 0 else nn.Identity()
            self.hierarchical = False
        else:
            self.num_classes = tuple(num_classes)
            self.head = MultitaskHead(self.num_features, num_classes)
            self.hierarchical = True

        self.apply(self._init_weights)
        for layer in self.layers:
            layer._init_respostnorm()

    def _init_weights(self, m):
        if isinstance(m, nn.Linear):
            trunc_normal_(m.weight, std=0.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    @torch.jit.ignore
    def no_weight_decay(self):
        return {"absolute_pos_embed"}

    @torch.jit.ignore
    def no_weight_decay_keywords(self):
        return {"cpb_mlp", "logit_scale", "relative_position_bias_table"}

    def set_input_normalization(self, mean, std):
        """Extension for uint8 batches: per-channel mean / std in pixel units (data.py:130-136 applies
        ``(x - mean) / std`` on the device before the model); folded into the patch-embedding gather."""
        self.register_buffer("input_mean", torch.as_tensor(mean, dtype=torch.float32).reshape(-1), persistent=False)
        self.register_buffer("input_std", torch.as_tensor(std, dtype=torch.float32).reshape(-1), persistent=False)
        return self

    def forward_features(self, x, output_activations=False):
        if x.dtype == torch.uint8:
            if getattr(self, "input_mean", None) is None:
                raise RuntimeError("uint8 input needs set_input_normalization(mean, std) first")
            x = self.patch_embed(x, input_norm=(self.input_mean, self.input_std))
        else:
            x = self.patch_embed(x)
        if self.ape:
            x = x + self.absolute_pos_embed
        x = self.pos_drop(x)
        activations = [] if output_activations else None
        for layer in self.layers:
            x = layer(x)
            if output_activations:
                activations.append(x)
        if SwinTransformerBlock._fusable(self.norm) and x.is_cuda and x.dtype in (torch.float32, torch.bfloat16):
            x = hvf.ln_residual(x, None, self.norm.weight, self.norm.bias, None, self.norm.eps)  # swinv2.py:833
        else:
            x = self.norm(x)
        x = torch.flatten(self.avgpool(x.transpose(1, 2)), 1)
        return (x, activations) if output_activations else x

    def forward(self, x):
        x = self.forward_features(x)
        if type(self.head) is nn.Linear:
            return _linear(x, self.head.weight, self.head.bias)
        return self.head(x)

    def flops(self):
        total = self.patch_embed.flops() + sum(layer.flops() for layer in self.layers)
        total += self.num_features * self.patches_resolution[0] * self.patches_resolution[1] // (2 ** self.num_layers)
        if isinstance(self.num_classes, int):
            return total + self.num_features * self.num_classes
        if isinstance(self.num_classes, tuple):
            return total + sum(self.num_features * n for n in self.num_classes)
        raise RuntimeError(f"Internal error: self.num_classes should be int or tuple, not {type(self.num_classes)}")


@dataclasses.dataclass(frozen=True)
class Checkpoint:
    """``swin://path`` checkpoint reference (swinv2.py:870-895): loads ``["model"]`` and drops the
    three non-learned buffers so they are rebuilt by the constructor."""
    source: str
    path: str

    @classmethod
    def parse(cls, uri):
        m = re.match(r"^swin://([\w./-]+)$", uri)
        if m is None:
            raise ValueError(f"uri '{uri}' doesn't match the pattern!")
        return cls("swin", m.group(1))

    def load_model_dict(self, cache):
        return self.filter(torch.load(self.path, map_location="cpu")["model"])

    @staticmethod
    def filter(model_dict):
        skip = ("relative_position_index", "relative_coords_table", "logit_clamp_max")
        return {k: v for k, v in model_dict.items() if not any(s in k for s in skip)}


def swinv2_tiny(num_classes=10000, img_size=256, window_size=8, **kw):
    """SwinV2-T hyper-parameters (not in the reference tree; upstream Swin-V2 values)."""
    cfg = dict(embed_dim=96, depths=[2, 2, 6, 2], num_heads=[3, 6, 12, 24])
    cfg.update(kw)
    return SwinTransformerV2(img_size=img_size, num_classes=num_classes, window_size=window_size, **cfg)


def swinv2_base(num_classes=(3, 13, 51, 273, 1103, 4884, 10000), img_size=256, window_size=16, **kw):
    """SwinV2-B with the iNat21 taxonomy tiers as multitask heads (BASELINE.json configs[3])."""
    cfg = dict(embed_dim=128, depths=[2, 2, 18, 2], num_heads=[4, 8, 16, 32])
    cfg.update(kw)
    return SwinTransformerV2(img_size=img_size, num_classes=num_classes, window_size=window_size, **cfg)
