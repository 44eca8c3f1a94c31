"""Host-side cost of `Solver.run` on this engine, measured WITHOUT a GPU: the reference's example scripts run verbatim
(tools/run_examples.py) with `launcher.install(in_memory=True, maps=...)`, but the two native objects are replaced by
inert stand-ins -- `mapgen.MapSet` marshals the medium and stops before fw25_mapgen, `engine.Engine` marshals the problem
and returns a zero-filled genout -- so everything that remains is Python / numpy work of this repository and of the
reference around the native calls.  A diagnostic for the builder (where does `Solver.run` spend host time?), never a
benchmark value: no engine runs.

    python tools/profile_host_side.py linear_transducer [--maps device|host] [--profile]
"""

from __future__ import annotations

import argparse
import cProfile
import io
import pstats
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def install_stand_ins(marks: dict):
    from fullwave25_b200 import engine, mapgen
    from fullwave25_b200.problem import MAP_NAMES

    class MapSet:
        def __init__(self, spec, device=0, planes=None, background=False):
            t0 = time.perf_counter()
            md, keep, (self.d_table, self.dmap, self.ndmap) = mapgen.marshal_medium(spec)
            marks["marshal_medium_s"] = time.perf_counter() - t0
            self.shape = spec.extended_shape
            self.upload_ms = self.kernel_ms = 0.0
            self.invalid_count = 0

        def device_maps(self):
            out = {name: 0 for name in MAP_NAMES + ("dcmap",)}
            out["pitch"] = (self.shape[-1] + 31) // 32 * 32
            out["owner"] = self
            return out

        def close(self):
            pass

    class Engine:
        def __init__(self, pb, device=0, device_maps=None, **kw):
            t0 = time.perf_counter()
            pb.normalise()
            self.pb = pb
            self._c = engine.marshal(pb, device_maps=device_maps)
            marks["marshal_problem_s"] = time.perf_counter() - t0

        def run(self):
            t0 = time.perf_counter()
            out = np.zeros((self.pb.n_frames, self.pb.ncoordsout), np.float32)
            marks["genout_alloc_s"] = time.perf_counter() - t0
            marks["genout_GB"] = out.nbytes / 1e9
            return out, {"loop_ms": 0.0, "setup_ms": 0.0}

        def close(self):
            pass

    mapgen.MapSet = MapSet
    engine.Engine = Engine
    engine.run = lambda pb, device_ids=(0,): Engine(pb).run()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("name", nargs="?", default="linear_transducer")
    ap.add_argument("--maps", default="device", choices=["device", "host"])
    ap.add_argument("--profile", action="store_true")
    ap.add_argument("--top", type=int, default=25)
    a = ap.parse_args()
    from tools import run_examples
    marks: dict = {}
    install_stand_ins(marks)
    prof = cProfile.Profile() if a.profile else None
    if prof:
        prof.enable()
    rec = run_examples.run_example(a.name, "fw25-device" if a.maps == "device" else "fw25-host")
    if prof:
        prof.disable()
    rec.pop("out", None)
    print({k: rec[k] for k in ("example", "extended_grid", "steps", "sensors", "frames", "solver_init_s", "solver_run_s",
                               "script_total_s")})
    print(marks)
    if prof:
        s = io.StringIO()
        pstats.Stats(prof, stream=s).sort_stats("cumulative").print_stats(a.top)
        print(s.getvalue())


if __name__ == "__main__":
    main()
