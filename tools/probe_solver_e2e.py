"""`fullwave.Solver.run` end to end on the reference's own Python layer, example-shaped (whole-domain sensor like the
shipped examples): the reference binary vs this engine through every integration level, with the stages timed.

    gpurun -- python tools/probe_solver_e2e.py [NXxNY[xNZ]] [n_steps] [modT]      -> gpurun_out/solver_e2e.json
default: the linear-transducer example's user grid 468 x 468, 2805 steps, whole-domain sensor every 2nd step."""
import json
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from fullwave25_b200 import build, launcher  # noqa: E402
from tools import ref_objects  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
shape = tuple(int(v) for v in (args[0] if args else "468x468").split("x"))
n_steps = int(args[1]) if len(args) > 1 else 2805
modT = int(args[2]) if len(args) > 2 else 2
fw, grid, medium, source, _ = ref_objects.build(shape, n_steps=n_steps, n_sensors=8, n_air=16, modT=modT)
sensor = fw.Sensor(mask=np.ones(shape, dtype=bool), sampling_modulus_time=modT)     # the examples record everything
ndim = len(shape)
res = {"user_grid": shape, "n_steps": n_steps, "modT": modT, "n_sensors": int(np.prod(shape)),
       "genout_GB": float(np.prod(shape)) * -(-n_steps // modT) * 4 / 1e9}
out = {}


def timed(name, fn, reps=1):
    best = None
    for _ in range(reps):                       # in-memory paths: best of `reps` (single shots are noisy at 0.2 s)
        t0 = time.perf_counter()
        r = fn()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    res[name + "_s"] = best
    out[name] = r
    print(name, f"{res[name + '_s']:.2f} s", flush=True)


with tempfile.TemporaryDirectory(dir="/dev/shm") as td:
    def solver(sub, bin_path):
        return fw.Solver(Path(td) / sub, grid, medium, source, sensor, path_fullwave_simulation_bin=bin_path)
    s_ref = solver("ref", ref_objects.ref_bin(ndim))
    timed("reference_solver_run", lambda: s_ref.run())
    s_cli = solver("cli", build.CLI)
    timed("fw25_engine_executable_solver_run", lambda: s_cli.run())
    undo = launcher.install()
    try:
        s_ins = solver("ins", build.CLI)
        timed("launcher_install_solver_run", lambda: s_ins.run())
    finally:
        undo()
    t0 = time.perf_counter()
    s_ref.pml_builder.run(use_pml=s_ref.use_pml)
    res["pml_builder_run_s"] = time.perf_counter() - t0
    timed("run_solver_no_disk", lambda: launcher.run_solver(s_ref, return_stats=True), reps=3)
    got, stats = out.pop("run_solver_no_disk")
    out["run_solver_no_disk"] = got
    res["engine_stats_no_disk"] = {k: (float(v) if isinstance(v, float) else int(v)) for k, v in stats.items()}
    # GPU-built maps (fw25_mapgen) instead of PMLBuilder.run, and the patched Solver.run a user's script would call
    timed("run_solver_device_maps", lambda: launcher.run_solver(s_ref, maps="device", return_stats=True), reps=3)
    got, stats = out.pop("run_solver_device_maps")
    res["rel_l2_device_maps_vs_reference"] = float(
        np.linalg.norm(got.astype(np.float64) - out["reference_solver_run"]) /
        np.linalg.norm(out["reference_solver_run"].astype(np.float64)))
    res["engine_stats_device_maps"] = {k: (v if isinstance(v, str) else float(v)) for k, v in stats.items()}
    undo = launcher.install(in_memory=True, maps="device")
    try:
        def user_script():
            return fw.Solver(Path(td) / "mem", grid, medium, source, sensor, path_fullwave_simulation_bin=build.CLI).run()
        timed("patched_solver_construct_and_run_device_maps", user_script, reps=3)
        out.pop("patched_solver_construct_and_run_device_maps")
    finally:
        undo()
want = out.pop("reference_solver_run")
for k, v in out.items():
    res[k + "_bit_identical"] = bool(np.array_equal(v, want))
print(json.dumps(res, indent=1))
d = ROOT / "gpurun_out"
d.mkdir(exist_ok=True)
(d / "solver_e2e.json").write_text(json.dumps(res, indent=1))
