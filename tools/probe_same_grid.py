"""Full-size parity probe: the reference's sm_100 binary and this engine on the SAME synthetic inputs bench.py uses
(tools/bench_reference.write_inputs), sensor frames compared value by value.

    gpurun -- python tools/probe_same_grid.py 96x200x200 560x1240x1240
"""
import json
import os
import shutil
import subprocess
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

from bench import MEDIUM  # noqa: E402
from fullwave25_b200 import engine, synthetic_device  # noqa: E402
from tools.bench_reference import write_inputs  # noqa: E402
from tools.make_ref_golden import REF_BIN  # noqa: E402


def ours(gshape, nT, full3d=False, n_sensors=None):
    dev = torch.device("cuda", 0)
    med = dict(MEDIUM)
    if n_sensors:
        med["n_sensors"] = n_sensors
    pb, maps = synthetic_device.make_slab(gshape, 0, gshape[0], device=dev, nT=nT, **med)
    pb.dcmap_full3d = full3d
    dmaps = {k: (v if k == "pitch" else v.data_ptr()) for k, v in maps.items()}
    eng = engine.Engine(pb, device=0, device_maps=dmaps)
    eng.step(nT)
    eng.sync()
    out = eng.read_frames(0, pb.n_frames)
    ids = eng.local_sensor_ids()
    full = np.zeros((pb.n_frames, pb.ncoordsout), np.float32)
    full[:, ids] = out
    eng.close()
    del maps
    torch.cuda.empty_cache()
    return pb, full


def reference(gshape, nT, real_c=False, n_sensors=None):
    dev = torch.device("cuda", 0)
    work = Path("/dev/shm/fw25_probe_same")
    if work.exists():
        shutil.rmtree(work)
    med = dict(MEDIUM)
    if n_sensors:
        med["n_sensors"] = n_sensors
    pb = write_inputs(work, gshape, nT, med, dev)
    if real_c:
        K = np.fromfile(work / "K.dat", np.float32)
        rho = np.fromfile(work / "rho.dat", np.float32)
        np.sqrt(K / rho).astype(np.float32).tofile(work / "c.dat")
    exe = work / REF_BIN[3].name
    shutil.copy(REF_BIN[3], exe)
    exe.chmod(0o755)
    t0 = time.time()
    with (work / "log.txt").open("w") as log:
        r = subprocess.run([str(exe)], cwd=work, stdout=log, stderr=log, env=dict(os.environ, CUDA_VISIBLE_DEVICES="0"))
    if r.returncode != 0:
        raise RuntimeError((work / "log.txt").read_text()[-1500:])
    g = np.fromfile(work / "genout.dat", np.float32).reshape(-1, pb.ncoordsout)
    shutil.rmtree(work)
    return pb, g, time.time() - t0


def main():
    out = []
    nT = int(os.environ.get("NT", 34))
    ns = int(os.environ.get("NSENS", 0)) or None
    for spec in sys.argv[1:]:
        gshape = tuple(int(v) for v in spec.split("x"))
        pb, mine = ours(gshape, nT, n_sensors=ns)
        _, ref, wall = reference(gshape, nT, n_sensors=ns)
        rec = {"grid": spec, "nT": nT, "frames": int(ref.shape[0]), "sensors": int(ref.shape[1]), "ref_wall_s": round(wall, 1),
               "identical": bool(np.array_equal(mine, ref)), "n_diff": int((mine != ref).sum()),
               "nonzero_ref": int((ref != 0).sum()), "nonzero_ours": int((mine != 0).sum()),
               "absmax_ref": float(np.abs(ref).max()), "absmax_ours": float(np.abs(mine).max()),
               "rel_l2": float(np.linalg.norm(mine.astype(np.float64) - ref) / max(np.linalg.norm(ref.astype(np.float64)), 1e-30))}
        if not rec["identical"]:
            f, s = np.nonzero(mine != ref)
            k = min(8, len(f))
            rec["first_diffs"] = [{"frame": int(f[i]), "sensor": int(s[i]), "coord": pb.outc[s[i]].tolist(),
                                   "ours": float(mine[f[i], s[i]]), "ref": float(ref[f[i], s[i]])} for i in range(k)]
            rec["diff_sensor_x_range"] = [int(pb.outc[s, 0].min()), int(pb.outc[s, 0].max())]
            _, refc, _ = reference(gshape, nT, real_c=True, n_sensors=ns) if gshape[0] * gshape[1] * gshape[2] < 3e8 else (None, None, None)
            if refc is not None:
                rec["identical_with_real_c_dat"] = bool(np.array_equal(mine, refc))
                rec["ref_changes_with_c_dat"] = not bool(np.array_equal(ref, refc))
        print(json.dumps(rec), flush=True)
        out.append(rec)
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "probe_same_grid.json").write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
