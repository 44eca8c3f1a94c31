"""Host-built vs GPU-built coefficient maps on example-sized grids, same box, through the reference's own objects.

    gpurun -- python tools/probe_mapgen.py [NXxNY[xNZ] ...]            -> gpurun_out/mapgen_probe.json
default: 120x120x120 (examples/wave_3d) and 1297x2018 (examples/convex_transducer) user grids, ppw 12 -> 36 + 36 + 8
boundary cells per side like `Solver` picks (solver.py:483-486).

Per grid: (host) the reference's `PMLBuilder.run` + `Problem.from_fullwave_objects` (what `run_solver(maps="host")`
does before the engine starts) and the H2D upload the engine then performs; (device) `mapgen.MapSet` = upload of the
user-grid maps + one kernel.  Then `run_solver` end to end both ways and the parity of the maps and the traces."""
import json
import os
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from fullwave25_b200 import build, launcher, mapgen  # noqa: E402
from fullwave25_b200.problem import MAP_NAMES, Problem  # noqa: E402
from oracle import mapgen_oracle as mo  # noqa: E402   (a probe is test infrastructure: it only CHECKS with the oracle)
from tools import ref_objects  # noqa: E402

shapes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]] or [(120, 120, 120), (1297, 2018)]
report = {"cores": os.cpu_count(), "cases": []}
for shape in shapes:
    n_steps = 200 if len(shape) == 3 else 600
    fw, grid, medium, source, sensor = ref_objects.build(shape, n_steps=n_steps, n_sensors=256, n_air=32, modT=4, block=12)
    res = {"user_grid": shape, "n_steps": n_steps}
    with tempfile.TemporaryDirectory(dir="/dev/shm") as td:
        s = fw.Solver(Path(td) / "s", grid, medium, source, sensor, path_fullwave_simulation_bin=build.CLI)
        pmlb = s.pml_builder
        eg = pmlb.extended_grid
        ext = tuple(int(getattr(eg, a)) for a in ("nx", "ny", "nz")[: len(shape)])
        res["extended_grid"] = ext
        res["points"] = int(np.prod(ext))
        t0 = time.perf_counter()
        em = pmlb.run(use_pml=True)
        t1 = time.perf_counter()
        pb = Problem.from_fullwave_objects(eg, em, pmlb.extended_source, pmlb.extended_sensor)
        t2 = time.perf_counter()
        res["host_pml_builder_run_s"] = t1 - t0
        res["host_problem_assembly_s"] = t2 - t1
        spec = mapgen.MediumSpec.from_pml_builder(pmlb, use_pml=True, dcmap_full3d=True)
        mapgen.MapSet(spec).close()                      # warm-up: CUDA context, allocator
        t0 = time.perf_counter()
        spec = mapgen.MediumSpec.from_pml_builder(pmlb, use_pml=True, dcmap_full3d=True)
        ms = mapgen.MapSet(spec)
        res["device_total_s"] = time.perf_counter() - t0
        res["device_upload_ms"], res["device_kernel_ms"] = ms.upload_ms, ms.kernel_ms
        res["device_kernel_GBps_written"] = res["points"] * 56 / (ms.kernel_ms * 1e-3) / 1e9
        worst = 0
        for stem in MAP_NAMES + ("dcmap",):
            got, want = ms.read(stem), getattr(pb, stem)
            if stem[:4] in ("apml", "bpml"):
                d = mo.ulp_distance_f32(got, want)
                worst = max(worst, int(d.max()))
                res.setdefault("ab_elements_off_by_one_ulp", 0)
                res["ab_elements_off_by_one_ulp"] += int((d > 0).sum())
            else:
                assert np.array_equal(got, want), stem
        ms.close()
        res["ab_max_ulp_vs_reference"] = worst
        res["exact_maps_bit_identical"] = True
        for mode in ("host", "device"):
            launcher.run_solver(s, maps=mode)            # warm
            t0 = time.perf_counter()
            out, st = launcher.run_solver(s, maps=mode, return_stats=True)
            res[f"run_solver_{mode}_s"] = time.perf_counter() - t0
            res[f"run_solver_{mode}_loop_ms"] = st["loop_ms"]
            res[f"traces_{mode}"] = out
        a, b = res.pop("traces_host"), res.pop("traces_device")
        res["traces_rel_l2_device_vs_host"] = float(np.linalg.norm(a.astype(np.float64) - b) / np.linalg.norm(a.astype(np.float64)))
    res["speedup_setup"] = (res["host_pml_builder_run_s"] + res["host_problem_assembly_s"]) / res["device_total_s"]
    print(json.dumps(res), flush=True)
    report["cases"].append(res)
d = ROOT / "gpurun_out"
d.mkdir(exist_ok=True)
(d / "mapgen_probe.json").write_text(json.dumps(report, indent=1))
