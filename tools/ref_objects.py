"""Build the REFERENCE's own API objects (fullwave.Grid / MediumRelaxationMaps / Source / Sensor / Solver) for a
seeded heterogeneous attenuating medium, using the unmodified reference package (baseline/_ref on the GPU box,
/root/reference in the build container).

The relaxation-parameter lookup database the reference's `fullwave.Medium` needs is missing from the checkout
(/root/reference/.MISSING_LARGE_BLOBS), so the medium is built with `fullwave.MediumRelaxationMaps`, which takes
the relaxation maps directly (fullwave/medium.py:33-113) -- the same class `Medium.build()` returns.

Used by: bench.py (host-side setup baseline, reference Solver.run end to end), tests/test_dropin.py.
"""

from __future__ import annotations

import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

from fullwave25_b200 import synthetic  # noqa: E402
from tools.ref_import import import_fullwave  # noqa: E402


def build(user_shape, *, f0=1e6, c0=1540.0, ppw=12, cfl=0.2, n_steps=None, block=6, seed=0, n_sensors=64,
          n_air=16, modT=2, amp=1e5):
    """Returns (fullwave module, grid, medium, source, sensor).  user_shape: USER grid (the solver pads the PML)."""
    fw = import_fullwave()
    ndim = len(user_shape)
    dx = c0 / f0 / ppw
    dt = cfl * dx / c0
    nt = n_steps or 200
    # Grid derives n = round(L/dx), nt = round(T/dt): pick L, T that round back to the requested sizes
    domain = tuple((n + 0.01) * dx for n in user_shape)
    grid = fw.Grid(domain, f0, (nt + 0.01) * dt, c0=c0, ppw=ppw, cfl=cfl)
    shape = tuple(int(getattr(grid, a)) for a in ("nx", "ny", "nz")[:ndim])
    assert shape == tuple(user_shape), (shape, user_shape)
    rng = np.random.default_rng(seed)
    coarse = tuple(-(-s // block) for s in shape)
    lab = rng.integers(0, len(synthetic.TISSUES), size=coarse)
    for ax in range(ndim):
        lab = np.repeat(lab, block, axis=ax)
    lab = lab[tuple(slice(0, s) for s in shape)]
    T = synthetic.TISSUES
    c = T[lab, 0] + rng.uniform(-0.4, 0.4, size=shape)
    tab = synthetic.relaxation_table(f0)
    relax = {
        "kappa_x1": tab[0, lab, 0], "kappa_x2": tab[1, lab, 0],
        "d_x1_nu1": tab[0, lab, 1], "alpha_x1_nu1": tab[0, lab, 2],
        "d_x1_nu2": tab[0, lab, 3], "alpha_x1_nu2": tab[0, lab, 4],
        "d_x2_nu1": tab[1, lab, 1], "alpha_x2_nu1": tab[1, lab, 2],
        "d_x2_nu2": tab[1, lab, 3], "alpha_x2_nu2": tab[1, lab, 4],
    }
    air = np.zeros(shape, dtype=bool)
    if n_air:
        idx = rng.choice(air.size, size=n_air, replace=False)
        air.flat[idx] = True
        air[:4] = False
    medium = fw.MediumRelaxationMaps(grid, c, T[lab, 1], T[lab, 2], relax, air_map=air)
    pmask = np.zeros(shape, dtype=bool)
    pmask[0:3] = True
    nsrc = int(pmask.sum())
    pulse = synthetic.tone_burst(int(grid.nt), dt, f0, amp=amp)
    p0 = np.zeros((nsrc, int(grid.nt)))
    per = nsrc // 3
    for layer in range(3):
        shift = int(round(layer * dx / c0 / dt))
        p0[per * layer: per * (layer + 1), shift:] = pulse[: int(grid.nt) - shift]
    source = fw.Source(p0, pmask)
    smask = np.zeros(shape, dtype=bool)
    cand = np.flatnonzero(~pmask)
    smask.flat[rng.choice(cand, size=min(n_sensors, cand.size), replace=False)] = True
    sensor = fw.Sensor(mask=smask, sampling_modulus_time=modT)
    return fw, grid, medium, source, sensor


def ref_bin(ndim: int, isotropic: bool = True) -> Path:
    for base in (ROOT / "baseline" / "_ref", Path("/root/reference")):
        p = (base / "fullwave" / "solver" / "bins" / "gpu" / f"{ndim}d" / "num_relax=2" /
             f"fullwave2_{ndim}d_2_relax_{'isotropic_' if isotropic else ''}multi_gpu_sm_100_cuda129")
        if p.exists():
            return p
    raise FileNotFoundError("reference sm_100 executable not found (run tools/install_reference.sh)")


def time_host_setup(user_shape, gpu: bool = False, **kw) -> dict:
    """Wall time of the reference's host-side setup for one run on this box's cores: PMLBuilder (pads the maps,
    builds the a/b/kappa PML maps; solver.py:527-536 + :694) and InputFileWriter's stencil tables
    (solver.py:734-743) -- everything `Solver.run` does before it touches the disk or the GPU."""
    import os
    fw, grid, medium, source, sensor = build(user_shape, **kw)
    from fullwave.solver.input_file_writer import InputFileWriter
    from fullwave.solver.pml_builder import PMLBuilder
    t0 = time.perf_counter()
    pml = PMLBuilder(grid=grid, medium=medium, source=source, sensor=sensor, m_spatial_order=8,
                     n_pml_layer=grid.ppw * 3, n_transition_layer=grid.ppw * 3, use_isotropic_relaxation=True)
    ext_medium = pml.run(use_pml=True)
    t1 = time.perf_counter()
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        InputFileWriter(work_dir=Path(td), grid=pml.extended_grid, medium=ext_medium, source=pml.extended_source,
                        sensor=pml.extended_sensor, path_fullwave_simulation_bin=ref_bin(len(user_shape)),
                        use_exponential_attenuation=False, use_isotropic_relaxation=True)
    t2 = time.perf_counter()
    eg = pml.extended_grid
    ext = tuple(int(getattr(eg, a)) for a in ("nx", "ny", "nz")[: len(user_shape)])
    pts = int(np.prod(ext))
    out = {"what": "reference PMLBuilder.__init__+run and InputFileWriter.__init__ (numpy, float64)",
           "user_grid": list(user_shape), "extended_grid": list(ext), "pml_builder_s": t1 - t0,
           "stencil_tables_s": t2 - t1, "Mpoints_per_s": pts / (t2 - t0) / 1e6, "cores": os.cpu_count(),
           "threads_used": 1}
    if gpu:   # the same maps + stencil tables from fw25_mapgen (user-grid upload + one kernel), same box
        from fullwave25_b200 import mapgen
        mapgen.MapSet(mapgen.MediumSpec.from_pml_builder(pml)).close()          # warm-up: context, allocator
        t3 = time.perf_counter()
        ms = mapgen.MapSet(mapgen.MediumSpec.from_pml_builder(pml))
        t4 = time.perf_counter()
        out["gpu_mapgen"] = {"what": "fullwave25_b200.mapgen.MapSet: the 13 maps + dcmap + stencil tables, in HBM",
                             "total_s": t4 - t3, "upload_ms": ms.upload_ms, "kernel_ms": ms.kernel_ms,
                             "Mpoints_per_s": pts / (t4 - t3) / 1e6,
                             "speedup_vs_host": (t2 - t0) / (t4 - t3)}
        ms.close()
    return out


if __name__ == "__main__":
    print(time_host_setup(tuple(int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "40x64x64").split("x"))))
