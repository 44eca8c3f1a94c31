"""Informational (GPU box): device-resident rate of the sweep variants on one synthetic 3D grid."""
import json, sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from fullwave25_b200 import engine, synthetic

shape = tuple(int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "280x280x280").split("x"))
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
t0 = time.time()
pb = synthetic.make_problem(shape, nT=10000, modT=4, n_sensors=1024, n_air=2000, seed=1234, n_pml=24, n_trans=24)
pb.dcmap_full3d = True
print("built problem in", round(time.time() - t0, 1), "s", flush=True)
res = {"shape": shape}
fields = {}
for variant in (1, 2, 3):
    with engine.Engine(pb, variant=variant) as e:
        e.step(5); e.sync()
        r = e.step_timed(steps, detail=True)
        r2 = e.step_timed(steps, detail=False)
        pts = pb.n_points * steps
        res[f"v{variant}"] = {"gpts": pts / r2["total_ms"] / 1e6, "gbps_208": pts * 208 / r2["total_ms"] / 1e6,
                              "u_ms": r["sweep_u_ms"] / steps, "p_ms": r["sweep_p_ms"] / steps,
                              "u_gbps": pb.n_points * 108 / (r["sweep_u_ms"] / steps) / 1e6,
                              "p_gbps": pb.n_points * 100 / (r["sweep_p_ms"] / steps) / 1e6,
                              "total_ms_per_step": r2["total_ms"] / steps, "launches": e.launches}
        fields[variant] = e.field("p")
        print(variant, res[f"v{variant}"], flush=True)
res["variants_bit_identical"] = bool(np.array_equal(fields[1], fields[2]) and np.array_equal(fields[1], fields[3]))
print(json.dumps(res))
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / f"variants_{'x'.join(map(str, shape))}.json").write_text(json.dumps(res))
