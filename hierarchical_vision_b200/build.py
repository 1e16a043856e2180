"""Build libhv_swin.so (the C-ABI library declared in include/hv_swin.h) in-tree with nvcc for sm_100a.

    python -m hierarchical_vision_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU box with the
repository snapshot.  Objects are rebuilt only when a source/header is newer.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
BUILD = os.path.join(PKG, "_build")
LIB = os.path.join(PKG, "libhv_swin.so")
SOURCES = ["api.cu", "wattn_generic.cu", "wattn_mma64.cu", "wattn_tc64.cu", "wattn_tc64_fwd2.cu", "wattn_tc64_bwd.cu", "wattn_tc256_fwd.cu", "wattn_tc256_bwd.cu", "ln_residual.cu", "bias_gelu.cu", "mlp_dgelu_gemm.cu", "patch_merge.cu", "patch_embed.cu", "cpb_bias.cu", "ce_loss.cu", "sgdw_step.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
          "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libhv_swin.so cannot be built (there is no non-CUDA fallback)")


def _newest_dep() -> float:
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(os.path.dirname(PKG), "include", "hv_swin.h")]
    return max(os.path.getmtime(p) for p in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(BUILD, exist_ok=True)
    nvcc = _nvcc()
    # extra flags (HV_NVCC_FLAGS, e.g. -DHV_TC_TRACE) are part of the build identity: a change rebuilds everything
    flags, stamp = os.environ.get("HV_NVCC_FLAGS", ""), os.path.join(BUILD, "flags.txt")
    if not os.path.exists(stamp) or open(stamp).read() != flags:
        force = True
        with open(stamp, "w") as f:
            f.write(flags)
    newest = _newest_dep()
    headers_newest = max(os.path.getmtime(os.path.join(CSRC, f)) for f in os.listdir(CSRC) if not f.endswith(".cu"))
    headers_newest = max(headers_newest, os.path.getmtime(os.path.join(os.path.dirname(PKG), "include", "hv_swin.h")))

    def compile_one(src: str):
        obj = os.path.join(BUILD, src.replace(".cu", ".o"))
        spath = os.path.join(CSRC, src)
        if (not force and os.path.exists(obj)
                and os.path.getmtime(obj) >= max(os.path.getmtime(spath), headers_newest)):
            return obj, ""
        cmd = [nvcc, *ARCH, *CFLAGS, *os.environ.get("HV_NVCC_FLAGS", "").split(), "-c", spath, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        with open(obj + ".log", "w") as f:
            f.write(r.stderr)
        return obj, r.stderr

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        results = list(ex.map(compile_one, SOURCES))
    objs = [o for o, _ in results]
    if verbose:
        for _, log in results:
            if log:
                print(log)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < max(newest, *(os.path.getmtime(o) for o in objs)):
        cmd = [nvcc, *ARCH, "-shared", "-o", LIB, *objs, "-Xcompiler", "-fPIC"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
