"""Per-kernel SASS evidence of the Blackwell paths: counts of the tcgen05 / TMEM / TMA / mbarrier mnemonics in the built
library (cuobjdump -sass).  UTCHMMA = tcgen05.mma (kind::f16), LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA tensor
load / store, UTCBAR = tcgen05.commit, SYNCS = mbarrier, HMMA = legacy mma.sync / wmma.

    python profiles/sass_summary.py > profiles/r02_sass_summary.txt
"""

import re
import subprocess
import sys
from collections import Counter, OrderedDict
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
LIB = ROOT / "generative-turbulence_b200" / "turbdiff_b200" / "libturbdiff_b200.so"
PATS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "UTCATOMSWS", "SYNCS", "HMMA", "FFMA", "DFMA", "MUFU", "ATOM", "RED"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    kernels = OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(anonymous namespace\)::", "", name)
            name = re.sub(r"\(.*", "", name)
            name = re.sub(r"^void ", "", name)
            cur = kernels.setdefault(name, Counter())
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            cur["instr"] += 1
            for p in PATS:
                if op.startswith(p):
                    cur[p] += 1
    cols = ["instr"] + PATS
    print(f"# SASS mnemonic counts per kernel of {LIB.name} (sm_100a); built from the sources at HEAD")
    print(f"{'kernel':70s} " + " ".join(f"{c:>8s}" for c in cols))
    for name, c in kernels.items():
        print(f"{name[:70]:70s} " + " ".join(f"{c.get(k, 0):8d}" for k in cols))
    tc = [n for n, c in kernels.items() if c.get("UTCHMMA")]
    print(f"\n# {len(tc)} kernels issue tcgen05.mma (UTCHMMA); kernels with legacy HMMA: {[n for n, c in kernels.items() if c.get('HMMA')]}")


if __name__ == "__main__":
    sys.exit(main())
