"""Engine-only rates on the extended-grid shapes of the reference's four shipped examples (BASELINE.md 2.2
configs 1-4) for BOTH engines on the same box: this engine (device-timed loop, through the
C-ABI) and the reference's sm_100 binary (wall time of the child process, rate by differencing two runs).
Synthetic heterogeneous media of those shapes (the examples' own media need the reference's missing LUT blob).

    gpurun -- python tools/probe_examples.py [--no-ref] [name ...]     -> gpurun_out/examples_rates.json
"""
import json
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from fullwave25_b200 import engine, synthetic  # noqa: E402

# name: (extended grid, steps of the shipped example, modT, steps pair used for differencing)
CONFIGS = {
    "simple_plane_wave_2d": ((861, 628), 7013, 7, (400, 2400)),
    "linear_transducer_2d": ((628, 628), 2805, 2, (400, 2400)),
    "convex_transducer_2d": ((1457, 2178), 3244, 2, (200, 1200)),
    "wave_3d": ((280, 280, 280), 1200, 7, (40, 240)),
    # crossover probes for the 2D kernel choice (not reference examples)
    "sq724_2d": ((724, 724), 0, 2, (400, 2400)),
    "sq1024_2d": ((1024, 1024), 0, 2, (400, 2400)),
    "sq1448_2d": ((1448, 1448), 0, 2, (200, 1200)),
}
BYTES = {2: 164, 3: 208}


def main():
    names = [a for a in sys.argv[1:] if not a.startswith("--")] or [n for n in CONFIGS if not n.startswith("sq")]
    with_ref = "--no-ref" not in sys.argv
    out = {}
    for name in names:
        shape, nT_example, modT, (n1, n2) = CONFIGS[name]
        nd = len(shape)
        pb = synthetic.make_problem(shape, nT=n2, modT=modT, n_sensors=512, n_air=64 if nd == 3 else 16, seed=1234,
                                    n_pml=36, n_trans=36)
        rec = {"shape": shape, "points": pb.n_points, "steps": n2}
        engine.run(pb)                                  # warm-up (module load, allocator)
        g, st = engine.run(pb)
        rec["engine_gpts"] = st["point_updates"] / st["loop_ms"] / 1e6
        rec["engine_us_per_step"] = st["loop_ms"] * 1e3 / n2
        rec["engine_launches_per_step"] = st["kernel_launches"] / n2
        rec["engine_GBps"] = rec["engine_gpts"] * BYTES[nd]
        rec["engine_setup_ms"] = st["setup_ms"]
        if with_ref:
            from tools.make_ref_golden import rel_l2, run_reference
            tmp = Path("/dev/shm/fw25_probe_ex")
            walls = []
            for nT in (n1, n2):
                pb.nT = nT
                gr, dt, _ = run_reference(pb, tmp, os.environ.get("FW25_REF_DEVICES", "0"), timeout=1800)
                walls.append(dt)
            rec["ref_walls_s"] = walls
            rec["ref_gpts"] = pb.n_points * (n2 - n1) / max(walls[1] - walls[0], 1e-9) / 1e9
            rec["ref_us_per_step"] = (walls[1] - walls[0]) * 1e6 / (n2 - n1)
            rec["bit_exact_vs_ref"] = bool(np.array_equal(g, gr))
            rec["rel_l2_vs_ref"] = rel_l2(g, gr)
            rec["speedup"] = rec["engine_gpts"] / rec["ref_gpts"]
        out[name] = rec
        print(name, json.dumps(rec), flush=True)
    d = ROOT / "gpurun_out"
    d.mkdir(exist_ok=True)
    (d / "examples_rates.json").write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
