"""Informational (GPU box): 2D anisotropic-relaxation sweeps, TMA-tiled (k_sweep_*_2dc<2, .., ANISO>) against the L1/L2-path
kernels, on the convex-transducer example's extended grid.  204 algorithmic bytes per point-update (2D isotropic 164 +
10 per-axis words).  usage: probe_aniso2d.py [XxY] [steps]"""
import json, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
from fullwave25_b200 import engine, synthetic

shape = tuple(int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "1457x2178").split("x"))
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 400
pb = synthetic.make_problem(shape, nT=10 ** 6, modT=50, seed=7, n_pml=36, n_trans=36, block=24, aniso=True, n_air=0,
                            n_sensors=256)
n = shape[0] * shape[1]
out = {"shape": shape, "steps": steps, "algorithmic_bytes_per_point_update": 204}
fields = {}
for name, variant in (("tiled", 0), ("simple", 1)):
    with engine.Engine(pb, variant=variant) as e:
        e.step(64); e.sync()
        r = e.step_timed(steps)
        e.sync()
        fields[name] = e.field("p")
        out[name] = {"us_per_step": r["total_ms"] / steps * 1e3, "gpts": n * steps / r["total_ms"] / 1e6,
                     "GBps": n * steps * 204 / r["total_ms"] / 1e6}
out["fields_identical"] = bool(np.array_equal(fields["tiled"], fields["simple"]))
out["absmax"] = float(np.abs(fields["tiled"]).max())
print(json.dumps(out))
