"""Generate tests/golden/stencil_tables.npz by running the REFERENCE's own InputFileWriter
(/root/reference/fullwave/solver/input_file_writer.py:95-103, :183-559) on small sound-speed maps.

Run in the build container (needs /root/reference):  python tools/make_stencil_golden.py
The committed .npz is what tests/test_stencil.py checks fullwave25_b200.stencil against.
"""

from __future__ import annotations

import sys
from pathlib import Path
from types import SimpleNamespace

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from tools.ref_import import import_fullwave  # noqa: E402

CASES = {
    # name: (is_3d, c0, f0, ppw, cfl, c_lo, c_hi, seed)
    "plane2d": (False, 1540.0, 3e6, 12, 0.2, 1540.0, 1540.0, 0),
    "abdo2d": (False, 1540.0, 1e6, 12, 0.2, 1412.3, 1613.4, 1),
    "convex2d": (False, 1540.0, 3.7e6, 12, 0.4, 1450.0, 1580.9, 2),
    "wave3d": (True, 1540.0, 1e6, 12, 0.2, 1480.5, 1600.2, 3),
    "synth3d": (True, 1540.0, 2e6, 8, 0.3, 1412.0, 1613.0, 4),
}


def main() -> None:
    import_fullwave()
    from fullwave.solver.input_file_writer import InputFileWriter

    out = {}
    for name, (is_3d, c0, f0, ppw, cfl, c_lo, c_hi, seed) in CASES.items():
        rng = np.random.default_rng(seed)
        shape = (5, 6, 7) if is_3d else (9, 11)
        c = rng.uniform(c_lo, c_hi, size=shape) if c_hi > c_lo else np.full(shape, c_lo)
        dx = c0 / f0 / ppw
        dt = cfl * dx / c0
        grid = SimpleNamespace(is_3d=is_3d, cfl=cfl, dt=dt, dx=dx)
        medium = SimpleNamespace(sound_speed=c)
        w = InputFileWriter(Path("/tmp"), grid, medium, None, None, validate_input=False)
        dim = w._dim
        out[f"{name}.c"] = c
        out[f"{name}.params"] = np.array([float(is_3d), dt, dx, cfl])
        out[f"{name}.d"] = w._d.astype(np.float32)
        out[f"{name}.dmap"] = w._d_map.astype(np.float32)
        out[f"{name}.dcmap"] = (w._dc_map - 1).astype(np.int32)
        out[f"{name}.ndmap"] = np.array(1 if dim == 0 else w._d_map.shape[2], np.int32)
    dst = ROOT / "tests" / "golden" / "stencil_tables.npz"
    np.savez_compressed(dst, **out)
    print("wrote", dst, dst.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
