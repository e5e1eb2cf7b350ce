"""Oracle (test infrastructure): install the UNMODIFIED reference package into ``oracle/_ref/`` (git-ignored, travels
to the GPU box with the gpurun snapshot) so that bench.py's reference arms and the drop-in tests can run it there.

    python oracle/install_ref.py          # also called by __graft_entry__.build() when /root/reference exists

The base recipe ``pip install --no-index --no-build-isolation --no-deps --target oracle/_ref /root/reference`` is tried
first; the reference builds with flit (pyproject.toml:1-3) and ``flit_core`` is not in this image, so that fails here.
A flit wheel of a pure-Python project is the package directory verbatim, so the fallback places exactly those files:
``/root/reference/turbdiff`` -> ``oracle/_ref/turbdiff``.  Nothing from it is ever committed (``.gitignore``) and the
product never imports it.
"""

from __future__ import annotations

import shutil
import subprocess
import sys
from pathlib import Path

SRC = Path("/root/reference")
DST = Path(__file__).resolve().parent / "_ref"


def install(force: bool = False) -> str:
    if not (SRC / "turbdiff").is_dir():
        return "skipped: /root/reference is not present (the prebuilt oracle/_ref is used as is)"
    marker = DST / "turbdiff" / "models" / "ddpm.py"
    if marker.is_file() and not force:
        return "present"
    if DST.exists():
        shutil.rmtree(DST)
    DST.mkdir(parents=True)
    r = subprocess.run([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--quiet", "--find-links",
                        "/opt/wheelhouse", "--target", str(DST), str(SRC)], capture_output=True, text=True)
    if r.returncode == 0 and marker.is_file():
        return "pip"
    shutil.copytree(SRC / "turbdiff", DST / "turbdiff", ignore=shutil.ignore_patterns("__pycache__"))
    (DST / "INSTALL.txt").write_text("placed by oracle/install_ref.py: pip could not build the flit project offline (flit_core missing); "
                                     "these are the files its wheel would contain\n")
    return "copied (pip failed: " + (r.stderr.strip().splitlines()[-1][:120] if r.stderr.strip() else "?") + ")"


if __name__ == "__main__":
    print("oracle/_ref:", install(force="--force" in sys.argv))
