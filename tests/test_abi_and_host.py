"""CPU: the C-ABI library loads and exports every symbol include/hv_swin.h declares; its host-side
integer maps (the same inline functions the kernels use) are bit-exact against the oracle and the
reference-generated digests; the module surface matches the reference's state_dict contract; and
the product refuses to run without CUDA (no CPU fallback)."""
import ctypes
import hashlib
import os
import re

import numpy as np
import pytest
import torch

import hierarchical_vision_b200 as hv
from hierarchical_vision_b200 import _lib
from oracle import swin_oracle as O
from tests._util import digests, load_case, manifest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_lib.LIB_PATH):
        from hierarchical_vision_b200 import build
        build.build()
    return _lib.load()


def test_every_declared_symbol_is_exported(lib):
    header = open(os.path.join(ROOT, "include", "hv_swin.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(hv_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.hv_abi_version() == 4
    assert lib.hv_compiled_arch() == 100


def test_shared_object_contains_sm100a_code():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout


@pytest.mark.parametrize("ws", [2, 4, 7, 8, 12, 16])
def test_host_relative_position_index_bit_exact(lib, ws):
    N = ws * ws
    out = np.empty((N, N), dtype=np.int64)
    assert lib.hv_relative_position_index(ws, out.ctypes.data_as(ctypes.c_void_p)) == 0
    assert np.array_equal(out, O.relative_position_index(ws))
    assert _sha(out) == digests()["relative_position_index"][str(ws)]["sha256"]


def test_host_shift_mask_bit_exact(lib):
    for key, rec in digests()["attn_mask"].items():
        H, W, ws, s = map(int, key.split(","))
        out = np.empty(rec["shape"], dtype=np.float32)
        assert lib.hv_shift_window_mask(H, W, ws, s, out.ctypes.data_as(ctypes.c_void_p)) == 0
        assert _sha(out) == rec["sha256"], key
    out = np.empty((1,), dtype=np.float32)
    assert lib.hv_shift_window_mask(16, 16, 8, 0, out.ctypes.data_as(ctypes.c_void_p)) != 0
    assert b"shift" in lib.hv_last_error()


def test_host_window_token_index_bit_exact(lib):
    for key, rec in digests()["window_token_index"].items():
        B, H, W, ws, s = map(int, key.split(","))
        out = np.empty(rec["shape"], dtype=np.int64)
        assert lib.hv_window_token_index(B, H, W, ws, s, out.ctypes.data_as(ctypes.c_void_p)) == 0
        assert _sha(out) == rec["sha256"], key


@pytest.mark.parametrize("geom", [(2, 32, 48, 0), (2, 32, 48, 8), (1, 16, 16, 8), (3, 16, 64, 8), (1, 64, 16, 0)])
def test_window16_tile_order_is_the_reference_roll_and_partition(lib, geom):
    """The tile order of the N = 256 kernels (two column parts of 8 x 16 tokens per window) is a fixed permutation of the
    window slots; composed with it, the kernels' token map equals torch.roll(-shift) + window_partition of the reference
    (swinv2.py:69-83, 399-412) bit for bit -- including the windows that wrap along the rows, the columns, or both."""
    B, H, W, shift = geom
    nW = (H // 16) * (W // 16)
    out = np.empty((B, nW, 256), dtype=np.int64)
    assert lib.hv_window16_tile_token_index(B, H, W, shift, out.ctypes.data_as(ctypes.c_void_p)) == 0
    ref = O.window_token_index(B, H, W, 16, shift).reshape(B, nW, 256)   # slot order: ih * 16 + iw
    t = np.arange(256)
    slot = ((t & 127) >> 3) * 16 + 8 * (t >> 7) + (t & 7)
    assert np.array_equal(out, ref[:, :, slot])
    assert lib.hv_window16_tile_token_index(B, H, W, 4, out.ctypes.data_as(ctypes.c_void_p)) != 0


def test_host_merge_token_index(lib):
    out = np.empty((2 * 3 * 4, 4), dtype=np.int64)
    assert lib.hv_merge_token_index(2, 6, 8, out.ctypes.data_as(ctypes.c_void_p)) == 0
    assert np.array_equal(out, O.merge_token_index(2, 6, 8))
    assert lib.hv_merge_token_index(1, 5, 8, out.ctypes.data_as(ctypes.c_void_p)) != 0
    assert b"not even" in lib.hv_last_error()


def test_bad_arguments_return_codes_not_crashes(lib):
    assert lib.hv_window_attn_fwd(None, None, None, None, 0, None, None, 1, 8, 8, 32, 1, 8, 0, 0, None) == 7  # HV_ERR_NULL
    assert lib.hv_relative_position_index(0, None) != 0
    assert lib.hv_window_attn_bwd_workspace_bytes(1, 8, 8, 30, 4, 8, 1) == 0  # C % heads != 0


def test_window16_statistics_and_workspace_sizes(lib):
    """Host-only queries for the N = 256 kernels: three statistics planes (lse | r | c) like the N = 64 kernels, and a
    backward workspace that holds the D = dO . O plane plus the per-CTA partials of d(bias table) / d(tau)."""
    B, H, W, C, heads, ws = 2, 32, 48, 128, 4, 16
    plane = B * (H // ws) * (W // ws) * heads * ws * ws
    assert lib.hv_window_attn_stats_floats(B, H, W, C, heads, ws, _lib.HV_BF16) == 3 * plane
    assert lib.hv_window_attn_stats_floats(B, H, W, C, heads, ws, _lib.HV_F32) == plane       # generic kernel
    assert lib.hv_window_attn_stats_floats(B, H, W, C, 2, ws, _lib.HV_BF16) == plane // 2    # head dim 64: generic kernel
    nbytes = lib.hv_window_attn_bwd_workspace_bytes(B, H, W, C, heads, ws, _lib.HV_BF16)
    assert nbytes >= 4 * plane + 4 * 961 * heads                                              # D plane + at least one partial per head
    assert lib.hv_window_attn_bwd_workspace_bytes(B, H, W, C, heads, ws, _lib.HV_F32) == 16  # generic: atomics, no workspace
    assert lib.hv_window_attn_stats_floats(B, H + 1, W, C, heads, ws, _lib.HV_BF16) == 0      # H % ws != 0


def test_kernel_kind_dispatch(lib):
    # bf16, head dim 32, window 8 (N = 64) or 16 (N = 256) -> tensor-core kernels; everything else -> generic kernel
    assert lib.hv_window_attn_kernel_kind(96, 3, 8, _lib.HV_BF16) == 1
    assert lib.hv_window_attn_kernel_kind(768, 24, 8, _lib.HV_BF16) == 1
    assert lib.hv_window_attn_kernel_kind(96, 3, 7, _lib.HV_BF16) == 0
    assert lib.hv_window_attn_kernel_kind(96, 3, 8, _lib.HV_F32) == 0
    assert lib.hv_window_attn_kernel_kind(128, 4, 16, _lib.HV_BF16) == 1
    assert lib.hv_window_attn_kernel_kind(128, 2, 16, _lib.HV_BF16) == 0  # head dim 64
    assert lib.hv_window_attn_kernel_kind(128, 4, 16, _lib.HV_F32) == 0
    assert lib.hv_window_attn_kernel_kind(128, 2, 8, _lib.HV_BF16) == 0  # head dim 64


# ------------------------------------------------------------------ module surface
BLOCKS = [k for k, v in manifest()["cases"].items() if v["kind"] == "block"]


@pytest.mark.parametrize("name", BLOCKS)
def test_block_state_dict_contract(name):
    meta, state, _ = load_case(name)
    blk = hv.SwinTransformerBlock(meta["C"], (meta["H"], meta["W"]), meta["heads"], window_size=meta["ws"],
                                  shift_size=meta["shift"], mlp_ratio=meta["mlp_ratio"])
    assert blk.window_size == meta["eff_ws"] and blk.shift_size == meta["eff_shift"]
    mine = blk.state_dict()
    assert list(mine.keys()) == list(state.keys())  # same names, same order
    for k in state:
        assert mine[k].shape == state[k].shape and mine[k].dtype == state[k].dtype, k
    # constant buffers are bit-identical to the reference's
    for k in ("attn.relative_position_index", "attn.relative_coords_table", "attn.logit_clamp_max"):
        assert torch.equal(mine[k], state[k]), k
    if "attn_mask" in state:
        assert torch.equal(mine["attn_mask"], state["attn_mask"])
    blk.load_state_dict(state, strict=True)


def test_model_state_dict_contract_and_flops():
    meta, state, _ = load_case("model_tiny")
    model = hv.SwinTransformerV2(img_size=meta["img_size"], patch_size=meta["patch_size"], in_chans=meta["in_chans"],
                                 num_classes=meta["num_classes"], embed_dim=meta["embed_dim"], depths=meta["depths"],
                                 num_heads=meta["num_heads"], window_size=meta["window_size"], drop_path_rate=0.0)
    assert list(model.state_dict().keys()) == list(state.keys())
    model.load_state_dict(state, strict=True)
    # reference init: res-post-norm gammas are zero (swinv2.py:603-608)
    fresh = hv.SwinTransformerV2(img_size=64, embed_dim=32, depths=[2, 2], num_heads=[1, 2], window_size=8, num_classes=5)
    assert float(fresh.layers[0].blocks[0].norm1.weight.abs().sum()) == 0.0
    assert float(fresh.layers[0].blocks[0].attn.logit_scale[0]) == pytest.approx(float(np.log(10.0)), rel=1e-6)
    # reference flops() of SwinV2-T 256/w8, 10k classes = 5.93 G MACs
    t = hv.swinv2_tiny()
    assert abs(t.flops() / 1e9 - 5.93) < 0.01
    assert sum(p.numel() for p in t.parameters()) == pytest.approx(35.27e6, rel=2e-3)
    heads = hv.swinv2_tiny(num_classes=(3, 5)).head
    assert isinstance(heads, hv.MultitaskHead) and len(heads.heads) == 2


def test_checkpoint_filter_and_parse():
    ck = hv.Checkpoint.parse("swin://some/dir/model.pth")
    assert ck.source == "swin" and ck.path == "some/dir/model.pth"
    with pytest.raises(ValueError):
        hv.Checkpoint.parse("http://x")
    d = {"layers.0.blocks.0.attn.relative_position_index": 1, "layers.0.blocks.0.attn.qkv.weight": 2,
         "a.relative_coords_table": 3, "a.logit_clamp_max": 4}
    assert list(hv.Checkpoint.filter(d)) == ["layers.0.blocks.0.attn.qkv.weight"]


def test_window_partition_reverse_roundtrip():
    x = torch.arange(2 * 8 * 12 * 3).view(2, 8, 12, 3)
    w = hv.window_partition(x, 4)
    assert w.shape == (2 * 2 * 3, 4, 4, 3)
    assert torch.equal(hv.window_reverse(w, 4, 8, 12), x)
    idx = O.window_token_index(2, 8, 12, 4, 0)
    assert np.array_equal(w[..., 0].reshape(-1, 16).numpy() // 3, idx)


def test_no_cpu_fallback():
    blk = hv.SwinTransformerBlock(32, (8, 8), 1, window_size=8)
    with pytest.raises(RuntimeError, match="no CPU path|CUDA"):
        blk(torch.randn(1, 64, 32))
    pm = hv.PatchMerging((8, 8), 32)
    with pytest.raises(RuntimeError, match="no CPU path|CUDA"):
        pm(torch.randn(1, 64, 32))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "hierarchical_vision_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, fn)).read()
                assert "import oracle" not in src and "from oracle" not in src and "oracle/" not in src.replace(
                    "routes through ``oracle/``", ""), fn


@pytest.mark.parametrize("setter", ["hv_window_attn_fwd_variant", "hv_window_attn_bwd_variant"])
def test_kernel_variant_setters_validate_their_argument(lib, setter):
    """The forward / backward variant switches are host-only state: -1 (automatic), 0 (mma.sync), 1 (tcgen05) are
    accepted (the forward also takes 2: first-generation tcgen05 kernel), anything else is an error code with a message,
    never a crash."""
    fn = getattr(lib, setter)
    for v in (0, 1, -1):
        assert fn(v) == 0
    top = 2 if setter == "hv_window_attn_fwd_variant" else 1
    assert fn(top) == 0
    assert fn(top + 1) != 0 and b"variant" in lib.hv_last_error()
    assert fn(-2) != 0
    assert fn(-1) == 0


def test_kernel_name_query_follows_the_dispatch(lib):
    """hv_window_attn_kernel_name (host-only): which kernel a geometry launches -- what bench.py prints next to its
    roofline numbers.  SwinV2-T stages run on the tcgen05 kernels in both directions."""
    import ctypes

    def name(B, H, W, C, heads, ws, shift, dtype, bwd):
        buf = ctypes.create_string_buffer(96)
        assert lib.hv_window_attn_kernel_name(B, H, W, C, heads, ws, shift, dtype, bwd, buf, len(buf)) == 0
        return buf.value.decode()

    for C, heads, res in ((96, 3, 64), (192, 6, 32), (384, 12, 16), (768, 24, 8)):
        assert name(4, res, res, C, heads, 8, 0, _lib.HV_BF16, 0) == "wattn_tc64_fwd2_kernel<false>"
        assert name(4, res, res, C, heads, 8, 0, _lib.HV_BF16, 1) == "wattn_tc64_bwd_kernel<false>"
        if res > 8:
            assert name(4, res, res, C, heads, 8, 4, _lib.HV_BF16, 1) == "wattn_tc64_bwd_kernel<true>"
            assert name(4, res, res, C, heads, 8, 4, _lib.HV_BF16, 0) == "wattn_tc64_fwd2_kernel<true>"
    assert name(1, 24, 24, 96, 3, 8, 2, _lib.HV_BF16, 0) == "wattn_tc64_fwd_kernel"         # even shift != ws / 2: first generation
    assert name(1, 24, 24, 96, 3, 8, 3, _lib.HV_BF16, 0) == "wattn_mma64_fwd_kernel<3>"     # odd shift: no TMA split
    # SwinV2-B window 16, head dim 32, bf16: the N = 256 tcgen05 kernels for shift 0 / 8, generic otherwise
    assert name(1, 32, 32, 128, 4, 16, 8, _lib.HV_BF16, 1) == "wattn_tc256_bwd_kernel"
    assert name(2, 16, 16, 512, 16, 16, 0, _lib.HV_BF16, 0) == "wattn_tc256_fwd_kernel"
    assert name(1, 32, 32, 128, 4, 16, 4, _lib.HV_BF16, 0) == "wattn_generic_fwd_kernel<bf16>"
    assert name(1, 32, 32, 128, 2, 16, 8, _lib.HV_BF16, 1) == "wattn_generic_bwd_kernel<bf16>"  # head dim 64
    assert name(1, 32, 32, 128, 4, 16, 8, _lib.HV_F32, 1) == "wattn_generic_bwd_kernel<float>"
    assert lib.hv_window_attn_tc256_variant(0) == 0
    assert name(1, 32, 32, 128, 4, 16, 8, _lib.HV_BF16, 1) == "wattn_generic_bwd_kernel<bf16>"
    assert lib.hv_window_attn_tc256_variant(2) != 0 and b"variant" in lib.hv_last_error()
    assert lib.hv_window_attn_tc256_variant(-1) == 0
    assert name(1, 16, 16, 96, 3, 8, 4, _lib.HV_F32, 0) == "wattn_generic_fwd_kernel<float>"
    buf = ctypes.create_string_buffer(8)
    assert lib.hv_window_attn_kernel_name(1, 16, 16, 96, 5, 8, 0, _lib.HV_BF16, 0, buf, 8) != 0  # C % heads
