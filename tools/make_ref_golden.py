"""Generate golden sensor traces with the REFERENCE's own shipped sm_100 engine (run on a B200 box).

    gpurun -- python tools/make_ref_golden.py            # writes gpurun_out/golden/ref_<case>.npz

For every seeded case of tests/cases.py the script writes the reference's `.dat` simulation directory
(fullwave25_b200.problem.Problem.to_dat_dir == what /root/reference/fullwave/solver/
input_file_writer.py:563-881 writes), copies the reference executable into it and runs it there with
no arguments (launcher.py:196-215), and stores `genout.dat` (float32 [n_frames, ncoordsout]).  The
npz files are then committed under tests/golden/ and pin both the CPU oracle
(tests/test_oracle_golden.py, CPU) and the CUDA engine (tests/test_gpu_parity.py, GPU).

The reference package is read from baseline/_ref (tools/install_reference.sh); /root/reference does
not exist on the GPU box.  The comparison printed at the end (oracle and CUDA engine vs reference)
is informational; the tests are what gate.
"""

from __future__ import annotations

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

from tests import cases  # noqa: E402

BINS = ROOT / "baseline" / "_ref" / "fullwave" / "solver" / "bins" / "gpu"
REF_BIN = {
    2: BINS / "2d" / "num_relax=2" / "fullwave2_2d_2_relax_isotropic_multi_gpu_sm_100_cuda129",
    3: BINS / "3d" / "num_relax=2" / "fullwave2_3d_2_relax_isotropic_multi_gpu_sm_100_cuda129",
}


# the anisotropic-relaxation engine family (upstream use_isotropic_relaxation=False)
REF_BIN_ANISO = {
    2: BINS / "2d" / "num_relax=2" / "fullwave2_2d_2_relax_multi_gpu_sm_100_cuda129",
    3: BINS / "3d" / "num_relax=2" / "fullwave2_3d_2_relax_multi_gpu_sm_100_cuda129",
}


def run_reference(pb, work: Path, devices: str = "0", timeout: float = 600.0):
    """Returns (genout [n_frames, ncoordsout] float32, wall seconds, log text)."""
    if work.exists():
        shutil.rmtree(work)
    pb.to_dat_dir(work)
    src = (REF_BIN_ANISO if getattr(pb, "aniso", None) else REF_BIN)[pb.ndim]
    exe = work / src.name
    shutil.copy(src, exe)
    exe.chmod(0o755)
    env = dict(os.environ, CUDA_VISIBLE_DEVICES=devices)
    t0 = time.time()
    with (work / "fw2_execution.log").open("w") as log:
        r = subprocess.run([str(exe)], cwd=work, stdout=log, stderr=log, env=env, timeout=timeout)
    dt = time.time() - t0
    text = (work / "fw2_execution.log").read_text(errors="replace")
    if r.returncode != 0:
        raise RuntimeError(f"reference engine exited with {r.returncode}:\n{text[-2000:]}")
    g = np.fromfile(work / "genout.dat", dtype=np.float32)
    return g.reshape(-1, max(pb.ncoordsout, 1))[:, : pb.ncoordsout], dt, text


def rel_l2(a, b):
    n = np.linalg.norm(b.astype(np.float64))
    d = np.linalg.norm(a.astype(np.float64) - b.astype(np.float64))
    return float(d / n) if n else float(d)


def main() -> None:
    out = ROOT / "gpurun_out" / "golden"
    out.mkdir(parents=True, exist_ok=True)
    tmp = Path(os.environ.get("FW25_TMP", "/dev/shm" if Path("/dev/shm").exists() else "/tmp")) / "fw25_golden"
    devices = os.environ.get("FW25_REF_DEVICES", "0")
    suffix = "" if devices == "0" else "_g" + str(len(devices.split(",")))
    names = sys.argv[1:] or sorted(cases.CASES)
    report = {}
    try:
        from oracle import oracle
    except Exception as e:  # noqa: BLE001
        oracle = None
        print("oracle unavailable:", e)
    try:
        from fullwave25_b200 import engine
        engine.lib()
    except Exception as e:  # noqa: BLE001
        engine = None
        print("engine unavailable:", e)
    for name in names:
        pb = cases.make(name)
        try:
            g, dt, log = run_reference(pb, tmp / name, devices)
        except Exception as e:  # noqa: BLE001
            print(f"[{name}] reference FAILED: {e}")
            report[name] = {"error": str(e)[:500]}
            continue
        head = "\n".join(l for l in log.splitlines() if "Progress" not in l)[:4000]
        (out / f"ref_{name}{suffix}.log").write_text(head)
        np.savez_compressed(out / f"ref_{name}{suffix}.npz", genout=g,
                            case=json.dumps(cases.CASES.get(name) or cases.CASES_BIG[name]),
                            devices=devices)
        rec = {"frames": int(g.shape[0]), "sensors": int(g.shape[1]), "wall_s": round(dt, 3),
               "finite": bool(np.isfinite(g).all()), "absmax": float(np.abs(g).max()) if g.size else 0.0}
        if oracle is not None:
            o = oracle.run(pb)
            rec["oracle_rel_l2"] = rel_l2(o, g)
            rec["oracle_bit_exact"] = bool(np.array_equal(o, g))
            rec["oracle_n_diff"] = int((o != g).sum())
        if engine is not None:
            try:
                e, _ = engine.run(pb)
                rec["engine_rel_l2"] = rel_l2(e, g)
                rec["engine_bit_exact"] = bool(np.array_equal(e, g))
            except Exception as ex:  # noqa: BLE001
                rec["engine_error"] = str(ex)[:300]
        report[name] = rec
        print(f"[{name}] {rec}")
    (out / f"report{suffix}.json").write_text(json.dumps(report, indent=1))


if __name__ == "__main__":
    main()
