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
OBJ = HERE / "build" / "obj"
CFLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
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


def _compile(job):
    src, obj, verbose = job
    cmd = [NVCC, *CFLAGS, *(["-Xptxas", "-v"] if verbose else []), "-c", "-o", str(obj), str(src)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return src.name, r.returncode, r.stdout + r.stderr


def build(force: bool = False, verbose: bool = False) -> Path:
    """One object per source file (compiled in parallel, rebuilt only when the file or a shared header changed),
    then one link step.  Objects live in build/ (git-ignored)."""
    from concurrent.futures import ThreadPoolExecutor

    stamp = LIB.with_suffix(".so.stamp")
    digest = _digest()
    if not force and LIB.exists() and stamp.exists() and stamp.read_text() == digest:
        return LIB
    OBJ.mkdir(parents=True, exist_ok=True)
    hdr = hashlib.sha256()
    for f in sorted([*CSRC.glob("*.cuh"), *CSRC.glob("*.h"), HERE.parent / "include" / "turbdiff_b200.h", Path(__file__)]):
        hdr.update(f.read_bytes())
    jobs, objs = [], []
    for src in sources():
        obj = OBJ / (src.stem + ".o")
        tag = OBJ / (src.stem + ".digest")
        d = hashlib.sha256(hdr.digest() + src.read_bytes()).hexdigest()
        objs.append(obj)
        if force or verbose or not obj.exists() or not tag.exists() or tag.read_text() != d:
            jobs.append((src, obj, verbose, tag, d))
    print(f"[build] nvcc {' '.join(CFLAGS)} -c  ({len(jobs)} of {len(objs)} sources)", file=sys.stderr)
    with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4) or 1) as ex:
        for (name, rc, log), job in zip(ex.map(_compile, [j[:3] for j in jobs]), jobs):
            if log.strip():
                print(f"[build] {name}:\n{log}", file=sys.stderr)
            if rc != 0:
                raise RuntimeError(f"nvcc failed on {name}")
            job[3].write_text(job[4])
    cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", str(LIB), *map(str, objs)]
    print("[build]", " ".join(cmd[:6]), "...", file=sys.stderr)
    subprocess.run(cmd, check=True)
    stamp.write_text(digest)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
