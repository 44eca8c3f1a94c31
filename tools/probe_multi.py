"""In-process multi-GPU (fw25_run with a device list == the reference's cuda_device_id=[0, 1, ...]) next to one
GPU on the same 3D problem: identical frames, loop time, halo bytes; optionally the reference binary on the same
GPUs (it shards in-process too).      gpurun --gpus 2 -- python tools/probe_multi.py [XxYxZ] [nT] [--ref]"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from fullwave25_b200 import engine, synthetic  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
shape = tuple(int(v) for v in (args[0] if args else "384x384x384").split("x"))
nT = int(args[1]) if len(args) > 1 else 60
n_dev = engine.lib().fw25_device_count()
pb = synthetic.make_problem(shape, nT=nT, modT=4, n_sensors=512, n_air=64, seed=1234, n_pml=24, n_trans=24)
res = {"shape": shape, "nT": nT, "devices_visible": n_dev}
engine.run(pb)                                   # warm-up
g1, s1 = engine.run(pb)
res["n1"] = {"gpts": s1["point_updates"] / s1["loop_ms"] / 1e6, "loop_ms": s1["loop_ms"], "setup_ms": s1["setup_ms"]}
for n in (2, 4, 8):
    if n > n_dev:
        break
    engine.run(pb, device_ids=tuple(range(n)))   # warm-up (peer access, module load on the other devices)
    g, s = engine.run(pb, device_ids=tuple(range(n)))
    res[f"n{n}"] = {"gpts": s["point_updates"] / s["loop_ms"] / 1e6, "loop_ms": s["loop_ms"], "setup_ms": s["setup_ms"],
                    "bit_identical_to_1gpu": bool(np.array_equal(g, g1)), "halo_MB_per_step": s["halo_bytes"] / nT / 1e6,
                    "launches_per_step": s["kernel_launches"] / nT, "speedup_vs_1": s1["loop_ms"] / s["loop_ms"]}
if "--ref" in sys.argv:
    from tools.make_ref_golden import run_reference
    tmp = Path("/dev/shm/fw25_probe_multi")
    for n in (1, 2):
        if n > n_dev:
            break
        walls = []
        for steps in (nT // 3, nT):
            pb.nT = steps
            gr, dt, log = run_reference(pb, tmp, ",".join(map(str, range(n))), timeout=1800)
            walls.append(dt)
        pb.nT = nT
        res[f"ref_n{n}"] = {"walls_s": walls, "gpts": pb.n_points * (nT - nT // 3) / max(walls[1] - walls[0], 1e-9) / 1e9,
                            "bit_identical_to_ours_1gpu": bool(np.array_equal(gr, g1)),
                            "partition_log": [l for l in log.splitlines() if "GPU" in l or "region" in l][:12]}
print(json.dumps(res, indent=1))
out = ROOT / "gpurun_out"
out.mkdir(exist_ok=True)
(out / f"probe_multi_{n_dev}gpu.json").write_text(json.dumps(res, indent=1))
