"""How does the reference's own 2-GPU run differ from its 1-GPU run?  (run on a 2-GPU box)
For each far-source case: run the reference binary on 1 GPU, and twice on 2 GPUs; report determinism, the first
frame that differs, and where the differing sensors sit relative to the x-slab interface."""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from tests import cases  # noqa: E402
from tools.make_ref_golden import rel_l2, run_reference  # noqa: E402

tmp = Path("/dev/shm/fw25_ref_study")
out = {}
for name in sys.argv[1:] or sorted(cases.CASES_2GPU):
    pb = cases.make(name)
    g1, _, _ = run_reference(pb, tmp / "a", "0")
    g2a, _, log = run_reference(pb, tmp / "b", "0,1")
    g2b, _, _ = run_reference(pb, tmp / "c", "0,1")
    diff = g2a != g1
    frames = np.flatnonzero(diff.any(axis=1))
    sens = np.flatnonzero(diff.any(axis=0))
    half = pb.nX // 2
    dist = pb.outc[:, 0] - half
    per_sensor = np.linalg.norm((g2a - g1).astype(np.float64), axis=0) / np.maximum(np.linalg.norm(g1.astype(np.float64), axis=0), 1e-30)
    order = np.argsort(-per_sensor)[:8]
    out[name] = {
        "two_gpu_runs_identical": bool(np.array_equal(g2a, g2b)),
        "rel_l2_2gpu_vs_1gpu": rel_l2(g2a, g1),
        "first_differing_frame": int(frames[0]) if frames.size else None, "frames": int(g1.shape[0]), "modT": pb.modT,
        "n_differing_sensors": int(sens.size), "n_sensors": int(g1.shape[1]),
        "worst_sensors_x_minus_interface": [int(dist[i]) for i in order],
        "worst_sensors_rel_l2": [float(per_sensor[i]) for i in order],
        "partition": [l for l in log.splitlines() if "range" in l][:8],
    }
    print(name, json.dumps(out[name], indent=1), flush=True)
d = ROOT / "gpurun_out"
d.mkdir(exist_ok=True)
(d / "ref_multi_gpu_study.json").write_text(json.dumps(out, indent=1))
