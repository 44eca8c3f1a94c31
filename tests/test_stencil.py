"""fullwave25_b200.stencil against tables produced by the reference's own InputFileWriter
(tools/make_stencil_golden.py; /root/reference/fullwave/solver/input_file_writer.py:183-559)."""

from pathlib import Path

import numpy as np
import pytest

from fullwave25_b200 import stencil

GOLD = np.load(Path(__file__).parent / "golden" / "stencil_tables.npz")
NAMES = sorted({k.split(".")[0] for k in GOLD.files})


@pytest.mark.parametrize("name", NAMES)
def test_tables_bit_exact(name):
    c = GOLD[f"{name}.c"]
    is_3d, dt, dx, cfl = GOLD[f"{name}.params"]
    d, dmap, dcmap, ndmap = stencil.tables(c, dt=dt, dx=dx, cfl=cfl, is_3d=bool(is_3d))
    assert ndmap == int(GOLD[f"{name}.ndmap"])
    np.testing.assert_array_equal(dcmap, GOLD[f"{name}.dcmap"])
    # bit-exact float32 tables
    assert d.astype(np.float32).tobytes() == GOLD[f"{name}.d"].tobytes()
    assert dmap.tobytes() == GOLD[f"{name}.dmap"].tobytes()


def test_weights_are_a_consistent_first_derivative():
    # sum_k (2k-1) D_k = 1 to the scheme's design accuracy (exact-derivative condition on a linear field)
    d = stencil.d_table(0.2, is_3d=True)
    assert abs(sum((2 * k - 1) * d[k, 0] for k in range(1, 9)) - 1.0) < 2e-2
    d2 = stencil.d_table(0.2, is_3d=False)
    assert abs(sum((2 * k - 1) * d2[k, 0] for k in range(1, 9)) - 1.0) < 2e-2


def test_matlab_round_half_up():
    assert stencil.matlab_round(1540.5) == 1541
    assert stencil.matlab_round(1540.4999) == 1540
    assert list(stencil.matlab_round(np.array([0.5, 1.5, 2.5]))) == [1, 2, 3]
