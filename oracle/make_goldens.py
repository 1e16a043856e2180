"""TEST INFRASTRUCTURE ONLY -- generate ``tests/golden/*`` by running the UNMODIFIED reference.

Run in the build container (the only place ``/root/reference`` exists):

    python -m oracle.make_goldens

The reference has no tests or golden vectors of its own, so parity is
pinned on outputs of the reference *executed here*: seeded inputs, the reference module's
state_dict, its fp32 outputs/gradients and (same inputs, module cast to double) fp64
outputs/gradients.  Fixtures are small on purpose (a few hundred kB each).
"""
from __future__ import annotations

import hashlib
import json
import os

import numpy as np
import torch

from . import ref_loader

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _sha(a) -> str:
    if isinstance(a, torch.Tensor):
        a = a.contiguous().numpy()
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _randomise(block, gen):
    """Reference init zeroes norm gamma/beta (swinv2.py:603-608) and leaves q/v bias at 0:
    re-randomise so that every gradient path carries signal."""
    with torch.no_grad():
        for n in (block.norm1, block.norm2):
            n.weight.copy_(1.0 + 0.1 * torch.randn(n.weight.shape, generator=gen))
            n.bias.copy_(0.1 * torch.randn(n.bias.shape, generator=gen))
        a = block.attn
        a.logit_scale.copy_(torch.empty(a.logit_scale.shape).uniform_(1.6, 3.9, generator=gen))
        a.q_bias.copy_(0.1 * torch.randn(a.q_bias.shape, generator=gen))
        a.v_bias.copy_(0.1 * torch.randn(a.v_bias.shape, generator=gen))
        a.cpb_mlp[0].bias.copy_(0.1 * torch.randn(a.cpb_mlp[0].bias.shape, generator=gen))
        a.proj.bias.copy_(0.1 * torch.randn(a.proj.bias.shape, generator=gen))


def _run(module, x, gy, dtype):
    m = module.to(dtype)
    m.zero_grad(set_to_none=True)
    xin = x.to(dtype).clone().requires_grad_(True)
    y = m(xin) if not isinstance(xin, tuple) else m(*xin)
    y.backward(gy.to(dtype))
    grads = {k: p.grad.detach().clone() for k, p in m.named_parameters()}
    return y.detach(), xin.grad.detach(), grads


def _rel(a, b):
    a, b = a.double().reshape(-1), b.double().reshape(-1)
    n = b.norm().item()
    return ((a - b).norm().item() / n) if n > 0 else (a - b).norm().item()


def _bf16_autocast_errors(module, x, gy, res64):
    """How far the *reference itself* lands from its fp64 result when run the way the reference's
    training loop runs it (autocast, main.py:32 precision="amp"), here with bfloat16 on the CPU.
    Stored in the manifest to calibrate the bf16 tolerance of hard cases (large logit_scale
    amplifies the bf16 rounding of q/k by up to 100x before the softmax)."""
    m = module.to(torch.float32)
    m.zero_grad(set_to_none=True)
    xin = x.clone().requires_grad_(True)
    with torch.autocast("cpu", dtype=torch.bfloat16):
        y = m(xin)
    y.float().backward(gy)
    y64, dx64, g64 = res64
    out = {"y": _rel(y, y64), "dx": _rel(xin.grad, dx64)}
    for k, p_ in m.named_parameters():
        out["grad." + k] = _rel(p_.grad, g64[k])
    return out


def _pack(prefix, state, x, gy, res32, res64):
    out = {f"{prefix}x": x.numpy(), f"{prefix}gy": gy.numpy()}
    for k, v in state.items():
        out[f"{prefix}state.{k}"] = v.numpy()
    # "ref.*" = reference run in float64 on the same inputs, stored as float32 (the rounding is
    # 6e-8 relative, far below every tolerance used); "f32.y" = the reference's own fp32 output.
    y, dx, grads = res64
    out[f"{prefix}ref.y"] = y.numpy().astype(np.float32)
    out[f"{prefix}ref.dx"] = dx.numpy().astype(np.float32)
    for k, v in grads.items():
        out[f"{prefix}ref.grad.{k}"] = v.numpy().astype(np.float32)
    out[f"{prefix}f32.y"] = res32[0].numpy()
    return out


def block_case(ref, name, B, H, W, C, heads, ws, shift, seed, clamp_head=False, mlp_ratio=1.0, init_scale=False):
    gen = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    blk = ref.SwinTransformerBlock(C, (H, W), heads, window_size=ws, shift_size=shift, mlp_ratio=mlp_ratio)
    _randomise(blk, gen)
    if init_scale:  # keep logit_scale at the reference's init value log(10) (swinv2.py:135-137)
        with torch.no_grad():
            blk.attn.logit_scale.fill_(float(torch.log(torch.tensor(10.0))))
    if clamp_head:  # one head above log(100): clamp active, zero logit_scale gradient (swinv2.py:230)
        with torch.no_grad():
            blk.attn.logit_scale[0] = 5.0
    state = {k: v.detach().clone() for k, v in blk.state_dict().items()}
    x = torch.randn(B, H * W, C, generator=gen)
    gy = torch.randn(B, H * W, C, generator=gen)
    r32 = _run(blk, x, gy, torch.float32)
    r64 = _run(blk, x, gy, torch.float64)
    bf16_err = _bf16_autocast_errors(blk, x, gy, r64)
    meta = dict(kind="block", ref_bf16_autocast_rel_l2=bf16_err, init_scale=init_scale, B=B, H=H, W=W, C=C, heads=heads, ws=ws, shift=shift, mlp_ratio=mlp_ratio,
                eff_ws=blk.window_size, eff_shift=blk.shift_size)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **_pack("", state, x, gy, r32, r64))
    return meta


def window_attention_case(ref, name, B_, ws, C, heads, nW, seed):
    gen = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    wa = ref.WindowAttention(C, (ws, ws), heads)
    with torch.no_grad():
        wa.logit_scale.copy_(torch.empty(wa.logit_scale.shape).uniform_(1.6, 3.9, generator=gen))
        wa.q_bias.copy_(0.1 * torch.randn(C, generator=gen))
        wa.v_bias.copy_(0.1 * torch.randn(C, generator=gen))
    N = ws * ws
    # an arbitrary {0,-100} mask that is NOT the shift pattern: exercises the explicit-mask path
    mask = torch.where(torch.rand(nW, N, N, generator=gen) < 0.3, torch.tensor(-100.0), torch.tensor(0.0))
    idx = torch.arange(N)
    mask[:, idx, idx] = 0.0
    state = {k: v.detach().clone() for k, v in wa.state_dict().items()}
    x = torch.randn(B_, N, C, generator=gen)
    gy = torch.randn(B_, N, C, generator=gen)

    class WithMask(torch.nn.Module):
        def __init__(self, inner, m):
            super().__init__()
            self.inner = inner
            self.register_buffer("m", m)

        def forward(self, t):
            return self.inner(t, mask=self.m)

    wrapped = WithMask(wa, mask)
    r32 = _run(wrapped, x, gy, torch.float32)
    r64 = _run(wrapped, x, gy, torch.float64)
    strip = lambda r: (r[0], r[1], {k[len("inner."):]: v for k, v in r[2].items()})
    out = _pack("", state, x, gy, strip(r32), strip(r64))
    out["mask"] = mask.numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    return dict(kind="window_attention", B_=B_, ws=ws, C=C, heads=heads, nW=nW)


def patch_merging_case(ref, name, B, H, W, C, seed):
    gen = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    pm = ref.PatchMerging((H, W), C)
    with torch.no_grad():
        pm.norm.weight.copy_(1.0 + 0.1 * torch.randn(2 * C, generator=gen))
        pm.norm.bias.copy_(0.1 * torch.randn(2 * C, generator=gen))
    state = {k: v.detach().clone() for k, v in pm.state_dict().items()}
    x = torch.randn(B, H * W, C, generator=gen)
    gy = torch.randn(B, (H // 2) * (W // 2), 2 * C, generator=gen)
    r32 = _run(pm, x, gy, torch.float32)
    r64 = _run(pm, x, gy, torch.float64)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **_pack("", state, x, gy, r32, r64))
    return dict(kind="patch_merging", B=B, H=H, W=W, C=C)


def model_case(ref, name, seed, **kw):
    gen = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    model = ref.SwinTransformerV2(drop_path_rate=0.0, **kw)
    for layer in model.layers:
        for blk in layer.blocks:
            _randomise(blk, gen)
    state = {k: v.detach().clone() for k, v in model.state_dict().items()}
    B = 2
    x = torch.randn(B, 3, kw["img_size"], kw["img_size"], generator=gen)
    gy = torch.randn(B, kw["num_classes"], generator=gen)
    r32 = _run(model, x, gy, torch.float32)
    r64 = _run(model, x, gy, torch.float64)
    # keep the fixture small: store only a handful of gradients
    keep = ("layers.0.blocks.1.attn.qkv.weight", "layers.0.blocks.1.attn.logit_scale",
            "layers.0.blocks.1.attn.cpb_mlp.2.weight", "layers.0.downsample.reduction.weight",
            "layers.1.blocks.0.norm1.weight", "patch_embed.proj.weight", "head.bias")
    f = lambda r: (r[0], r[1], {k: v for k, v in r[2].items() if k in keep})
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **_pack("", state, x, gy, f(r32), f(r64)))
    meta = dict(kind="model", B=B)
    meta.update({k: (list(v) if isinstance(v, (list, tuple)) else v) for k, v in kw.items()})
    return meta


def index_digests(ref):
    """sha256 of the reference's integer / constant buffers (swinv2.py:147-190, 357-388) and of
    the composed roll+partition gather map (swinv2.py:69-83, 399-412)."""
    out = {"relative_position_index": {}, "relative_coords_table": {}, "attn_mask": {},
           "window_token_index": {}, "small_arrays": {}}
    for ws in (2, 4, 7, 8, 12, 16):
        wa = ref.WindowAttention(32, (ws, ws), 1)
        rpi = wa.relative_position_index
        out["relative_position_index"][str(ws)] = dict(
            sha256=_sha(rpi), shape=list(rpi.shape), dtype=str(rpi.dtype), sum=int(rpi.sum()),
            max=int(rpi.max()), samples={"0,0": int(rpi[0, 0]), "0,1": int(rpi[0, 1]), "1,0": int(rpi[1, 0]),
                                         f"0,{ws}": int(rpi[0, ws]), f"{ws * ws - 1},0": int(rpi[-1, 0])})
        tab = wa.relative_coords_table
        out["relative_coords_table"][f"{ws}/0"] = dict(sha256=_sha(tab), shape=list(tab.shape))
    for ws, pws in ((8, 12), (16, 8)):
        wa = ref.WindowAttention(32, (ws, ws), 1, pretrained_window_size=(pws, pws))
        out["relative_coords_table"][f"{ws}/{pws}"] = dict(sha256=_sha(wa.relative_coords_table),
                                                           shape=list(wa.relative_coords_table.shape))
    for (H, W, ws, s) in ((64, 64, 8, 4), (64, 64, 16, 8), (32, 32, 8, 4), (16, 16, 8, 4), (16, 32, 8, 4),
                          (16, 16, 8, 3), (24, 16, 8, 1), (14, 21, 7, 3), (32, 32, 16, 8), (16, 16, 8, 7)):
        blk = ref.SwinTransformerBlock(32, (H, W), 1, window_size=ws, shift_size=s)
        m = blk.attn_mask
        counts = (m != 0).reshape(m.shape[0], -1).sum(1).tolist()
        out["attn_mask"][f"{H},{W},{ws},{s}"] = dict(sha256=_sha(m), shape=list(m.shape), nonzero_per_window=counts)
        B = 2
        ids = torch.arange(B * H * W, dtype=torch.int64).view(B, H, W, 1)
        win = ref.window_partition(torch.roll(ids, (-s, -s), (1, 2)), ws).reshape(-1, ws * ws)
        out["window_token_index"][f"{B},{H},{W},{ws},{s}"] = dict(sha256=_sha(win), shape=list(win.shape))
    wa = ref.WindowAttention(32, (2, 2), 1)
    out["small_arrays"]["rpi_ws2"] = wa.relative_position_index.tolist()
    blk = ref.SwinTransformerBlock(32, (4, 4), 1, window_size=2, shift_size=1)
    out["small_arrays"]["mask_4x4_ws2_s1"] = blk.attn_mask.tolist()
    out["constants"] = dict(logit_scale_init=float(wa.logit_scale[0, 0, 0]), logit_clamp_max=float(wa.logit_clamp_max))
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = ref_loader.load()
    manifest = {"generator": "oracle/make_goldens.py", "torch": torch.__version__,
                "reference": "samuelstevens/hierarchical-vision swinv2.py (unmodified, imported from /root/reference)",
                "cases": {}}
    c = manifest["cases"]
    c["block_ws8_shift4"] = block_case(ref, "block_ws8_shift4", 2, 16, 16, 64, 2, 8, 4, seed=1, clamp_head=True, mlp_ratio=4.0)
    c["block_ws8_noshift"] = block_case(ref, "block_ws8_noshift", 1, 16, 16, 96, 3, 8, 0, seed=2)
    c["block_ws8_rect_shift3"] = block_case(ref, "block_ws8_rect_shift3", 1, 16, 24, 96, 3, 8, 3, seed=3)
    c["block_res_le_ws"] = block_case(ref, "block_res_le_ws", 1, 8, 8, 128, 4, 8, 4, seed=4)
    c["block_ws16_shift8"] = block_case(ref, "block_ws16_shift8", 1, 32, 32, 64, 2, 16, 8, seed=5)
    c["block_ws8_shift4_initscale"] = block_case(ref, "block_ws8_shift4_initscale", 1, 16, 16, 96, 3, 8, 4, seed=11,
                                                 init_scale=True)
    c["block_ws4_shift2"] = block_case(ref, "block_ws4_shift2", 2, 8, 8, 32, 1, 4, 2, seed=6)
    c["window_attention_mask"] = window_attention_case(ref, "window_attention_mask", 4, 8, 64, 2, 2, seed=7)
    c["patch_merging"] = patch_merging_case(ref, "patch_merging", 2, 16, 16, 48, seed=8)
    c["patch_merging_rect"] = patch_merging_case(ref, "patch_merging_rect", 1, 8, 12, 64, seed=9)
    c["model_tiny"] = model_case(ref, "model_tiny", seed=10, img_size=64, patch_size=4, in_chans=3, num_classes=10,
                                 embed_dim=32, depths=[2, 2], num_heads=[1, 2], window_size=8)
    with open(os.path.join(OUT, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    with open(os.path.join(OUT, "index_digests.json"), "w") as f:
        json.dump(index_digests(ref), f, indent=1, sort_keys=True)
    for fn in sorted(os.listdir(OUT)):
        print(f"{fn:40s} {os.path.getsize(os.path.join(OUT, fn)) / 1024:8.1f} kB")


if __name__ == "__main__":
    main()
