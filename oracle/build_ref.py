"""TEST INFRASTRUCTURE ONLY -- recipe for ``oracle/_ref``: the unmodified reference ``swinv2.py`` for the GPU box.

    python oracle/build_ref.py

The reference is pure Python: there is nothing to compile.  The one file of the hot path, ``/root/reference/swinv2.py``,
is copied byte for byte into ``oracle/_ref/`` (git-ignored: it never enters this repository's history; not
gpurun-ignored: it travels to the GPU box with the snapshot like a built ``.so``), together with its sha256.  There
it serves as the CPU reference arm of ``bench.py --impl reference`` (``cpu_baseline.kind = "reference"``) and as the
live reference of the parity tests; the product package never imports it (tests/test_abi_and_host.py checks).  Without
``/root/reference`` (on the GPU box) this script does nothing and an existing copy stays in place.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import sys

SRC = os.path.join(os.environ.get("HV_REFERENCE_ROOT", "/root/reference"), "swinv2.py")
DST_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def build() -> str:
    if not os.path.isfile(SRC):
        return ""
    os.makedirs(DST_DIR, exist_ok=True)
    dst = os.path.join(DST_DIR, "swinv2.py")
    shutil.copyfile(SRC, dst)
    digest = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    with open(os.path.join(DST_DIR, "SOURCE.txt"), "w") as f:
        f.write(f"byte-identical copy of {SRC}\nsha256 {digest}\n")
    return dst


if __name__ == "__main__":
    out = build()
    print(out or f"{SRC} not present: nothing to do", file=sys.stderr if not out else sys.stdout)
