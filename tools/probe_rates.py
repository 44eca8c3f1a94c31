"""Informational: engine-only rate of the reference's sm_100 binary (by differencing two runs,
SURVEY.md 8(d)) next to this engine's loop rate, on one synthetic 3D/2D grid.  Run on the GPU box."""
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from fullwave25_b200 import engine, synthetic  # noqa: E402
from tools.make_ref_golden import run_reference  # noqa: E402

shape = tuple(int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "280x280x280").split("x"))
nT1, nT2 = 10, 40
tmp = Path("/dev/shm/fw25_probe")
res = {"shape": shape}
pb = synthetic.make_problem(shape, nT=nT2, modT=1000000, n_sensors=1, n_air=64, seed=1234, n_pml=24, n_trans=24)
g, stats = engine.run(pb)
g, stats = engine.run(pb)
res["engine_gpts"] = stats["point_updates"] / stats["loop_ms"] / 1e6
res["engine_stats"] = stats
print(res, flush=True)
if "--no-ref" not in sys.argv:
    walls = []
    for nT in (nT1, nT2):
        pb.nT = nT
        gr, dt, log = run_reference(pb, tmp, os.environ.get("FW25_REF_DEVICES", "0"))
        walls.append(dt)
    res["ref_walls"] = walls
    res["ref_gpts"] = pb.n_points * (nT2 - nT1) / (walls[1] - walls[0]) / 1e9
    pb.nT = nT2
    ge, _ = engine.run(pb)
    res["ref_vs_engine_rel_l2"] = float(np.linalg.norm(ge.astype(float) - gr) / max(np.linalg.norm(gr), 1e-30))
print(json.dumps(res, default=str))
out = ROOT / "gpurun_out"
out.mkdir(exist_ok=True)
(out / f"probe_{'x'.join(map(str, shape))}.json").write_text(json.dumps(res, default=str))
