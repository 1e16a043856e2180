"""Shared helpers for the test-suite (test infrastructure)."""
from __future__ import annotations

import json
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def manifest():
    with open(os.path.join(GOLDEN, "manifest.json")) as f:
        return json.load(f)


def digests():
    with open(os.path.join(GOLDEN, "index_digests.json")) as f:
        return json.load(f)


def load_case(name):
    """Returns (meta, state: dict[str, Tensor], arrays: dict[str, Tensor])."""
    meta = manifest()["cases"][name]
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    state, other = {}, {}
    for k in z.files:
        t = torch.from_numpy(z[k])
        if k.startswith("state."):
            state[k[len("state."):]] = t
        else:
            other[k] = t
    return meta, state, other


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a = a.detach().double().cpu().reshape(-1)
    b = b.detach().double().cpu().reshape(-1)
    denom = b.norm().item()
    if denom == 0.0:
        return (a - b).norm().item()
    return ((a - b).norm() / denom).item()


def assert_close(name, got, want, tol):
    err = rel_l2(got, want)
    assert err <= tol, f"{name}: rel-L2 {err:.3e} > {tol:.1e} (|want|={want.double().norm().item():.3e})"
    return err
