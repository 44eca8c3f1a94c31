"""Opcode histogram of the sweep kernels in libfw25.so (cuobjdump -sass), the evidence behind "all reads go through
TMA, completion on mbarriers, no __syncthreads in the loop" (DESIGN.md 4).  Runs without a GPU.

    python tools/sass_histogram.py > profiles/sass_r02.txt
"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
lib = ROOT / "fullwave25_b200" / "libfw25.so"
out = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True).stdout
kernels, cur = {}, None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kernels[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and cur:
        kernels[cur][m.group(1)] += 1
WATCH = ("UTMALDG", "UTMASTG", "SYNCS", "BAR", "LDS", "STS", "LDG", "STG", "LD.", "ST.", "FFMA", "FADD", "FMUL", "MUFU", "CALL",
         "BRA", "IMAD", "LDSM", "LDGSTS", "ERRBAR", "MEMBAR", "ATOM", "RED")
print(f"cuobjdump -sass {lib.name}: instruction counts per kernel (static, per-opcode-family)\n")
for name in sorted(kernels):
    c = kernels[name]
    short = re.sub(r"\(.*", "", name.replace("(anonymous namespace)::", ""))
    if not any(k in short for k in ("k_sweep", "k_inject", "k_record", "k_mapgen", "k_persist")):
        continue
    fam = collections.Counter()
    for op, n in c.items():
        for w in WATCH:
            if op.startswith(w):
                fam[w.rstrip(".")] += n
                break
    tma = sorted((op, n) for op, n in c.items() if op.startswith("UTMA"))
    print(f"{short}\n   total {sum(c.values()):5d}   " + "  ".join(f"{k} {v}" for k, v in sorted(fam.items())) +
          ("   [" + ", ".join(f"{op} {n}" for op, n in tma) + "]" if tma else ""))
