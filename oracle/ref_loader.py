"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* reference ``swinv2.py``.

Only ``tests/``, ``oracle/make_goldens.py`` and ``bench.py``'s CPU legs may import this.
The product package (``hierarchical_vision_b200``) never does.

The reference (samuelstevens/hierarchical-vision, ``/root/reference/swinv2.py``) is pure
Python and needs exactly three symbols from ``timm.models.layers`` (swinv2.py:9), and
``timm`` is not installed in this image.  We inject a three-symbol stand-in into
``sys.modules`` and then import the reference file *from where it lies* -- nothing is
copied into this repository's history.  The reference tree exists only in the build container; for the
GPU box ``oracle/build_ref.py`` (run by ``__graft_entry__.build()``) places a byte-identical copy of that one
file under ``oracle/_ref/`` (git-ignored, shipped with the snapshot like a built ``.so``), which this loader
finds when ``/root/reference`` is absent.  It is used ONLY as the CPU reference arm of bench.py and by tests.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import torch

REF_COPY_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
_MODULE = None


def _root() -> str:
    """Where the unmodified reference swinv2.py lies: the reference tree, else the travelling copy."""
    for cand in (os.environ.get("HV_REFERENCE_ROOT"), "/root/reference", REF_COPY_DIR):
        if cand and os.path.isfile(os.path.join(cand, "swinv2.py")):
            return cand
    return os.environ.get("HV_REFERENCE_ROOT", "/root/reference")


REFERENCE_ROOT = _root()


def available() -> bool:
    return os.path.isfile(os.path.join(_root(), "swinv2.py"))


class _DropPath(torch.nn.Module):
    """Per-sample stochastic depth with timm's semantics: keep-mask of shape (B,1,..,1)
    drawn Bernoulli(1-p) and divided by the keep probability; identity in eval."""

    def __init__(self, drop_prob: float = 0.0, scale_by_keep: bool = True):
        super().__init__()
        self.drop_prob = drop_prob
        self.scale_by_keep = scale_by_keep

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1.0 - self.drop_prob
        shape = (x.shape[0],) + (1,) * (x.ndim - 1)
        mask = x.new_empty(shape).bernoulli_(keep)
        if keep > 0.0 and self.scale_by_keep:
            mask.div_(keep)
        return x * mask


def _to_2tuple(v):
    if isinstance(v, (tuple, list)):
        return tuple(v)
    return (v, v)


def _install_timm_stub() -> None:
    if "timm.models.layers" in sys.modules:
        return
    timm = types.ModuleType("timm")
    models = types.ModuleType("timm.models")
    layers = types.ModuleType("timm.models.layers")
    layers.DropPath = _DropPath
    layers.to_2tuple = _to_2tuple
    layers.trunc_normal_ = torch.nn.init.trunc_normal_
    timm.models = models
    models.layers = layers
    sys.modules.setdefault("timm", timm)
    sys.modules.setdefault("timm.models", models)
    sys.modules.setdefault("timm.models.layers", layers)


def load():
    """Return the reference ``swinv2`` module (imported in place), or raise."""
    global _MODULE
    if _MODULE is not None:
        return _MODULE
    if not available():
        raise FileNotFoundError(
            f"reference swinv2.py not found under {_root()} "
            "(expected on the GPU box: use tests/golden fixtures instead)"
        )
    _install_timm_stub()
    spec = importlib.util.spec_from_file_location(
        "hv_reference_swinv2", os.path.join(_root(), "swinv2.py")
    )
    mod = importlib.util.module_from_spec(spec)
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        spec.loader.exec_module(mod)
    _MODULE = mod
    return mod
