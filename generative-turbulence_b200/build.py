"""Build libturbdiff_b200.so in-tree with nvcc for sm_100a (no torch headers: pure C ABI).

    python generative-turbulence_b200/build.py [--force]
"""

from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
LIB = HERE / "turbdiff_b200" / "libturbdiff_b200.so"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo", "-shared", "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr", "-Xcompiler", "-fvisibility=hidden",
]


def sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _digest() -> str:
    h = hashlib.sha256()
    files = [f for f in CSRC.glob("*") if f.is_file()]
    for f in sorted(files + [HERE.parent / "include" / "turbdiff_b200.h", Path(__file__)]):
        h.update(f.name.encode())
        h.update(f.read_bytes())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    stamp = LIB.with_suffix(".so.stamp")
    digest = _digest()
    if not force and LIB.exists() and stamp.exists() and stamp.read_text() == digest:
        return LIB
    cmd = [NVCC, *FLAGS, *(["-Xptxas", "-v"] if verbose else []), "-o", str(LIB), *map(str, sources())]
    print("[build]", " ".join(cmd), file=sys.stderr)
    subprocess.run(cmd, check=True)
    stamp.write_text(digest)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
