"""Tuning helper: build libfw25 variants whose warp-specialised sweeps use other tile heights / CTAs per SM.
Only fw25_sweeps_ws.cu is recompiled; the variants land in fullwave25_b200/build/variants/ (git-ignored, travel to the
GPU box) and are selected with FW25_LIB=<path>.   usage: build_variants.py name:TYU,MINBU,TYP,MINBP ..."""
import subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from fullwave25_b200 import build as B

B.build()
nvcc = B._nvcc()
out = B.PKG / "build" / "variants"
out.mkdir(parents=True, exist_ok=True)
objs = [str(o) for o in sorted((B.PKG / "build").glob("*.o")) if o.name not in ("fw25_cli.o", "fw25_sweeps_ws.o")]
for spec in sys.argv[1:]:
    name, vals = spec.split(":")
    tyu, mu, typ, mp, *extra = vals.split(",")
    defs = [f"-DFW25_WS_TY_U={tyu}", f"-DFW25_WS_MINB_U={mu}", f"-DFW25_WS_TY_P={typ}", f"-DFW25_WS_MINB_P={mp}"]
    defs += [f"-D{e}" for e in extra]          # e.g. FW25_EXPERIMENT_FAST_DIV (timing experiment, not bit-exact)
    obj = out / f"ws_{name}.o"
    r = subprocess.run([nvcc, *B.NVCC_FLAGS, *defs, *B._host_cxx(), "-c", "-o", str(obj), str(B.CSRC / "fw25_sweeps_ws.cu")],
                       capture_output=True, text=True)
    if r.returncode:
        print(r.stdout + r.stderr); raise SystemExit(1)
    regs = [l for l in (r.stdout + r.stderr).splitlines() if "registers" in l or "spill" in l]
    lib = out / f"libfw25_{name}.so"
    subprocess.run([nvcc, *B._host_cxx(), "-shared", "-o", str(lib), *objs, str(obj)], check=True)
    print(name, defs, "\n   ", "\n    ".join(regs[:16]))
