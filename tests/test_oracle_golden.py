"""Pins the CPU oracle to the REFERENCE's own engine: tests/golden/ref_<case>.npz hold the sensor traces
(`genout.dat`) that the reference's shipped sm_100 executable
(fullwave/solver/bins/gpu/{2d,3d}/num_relax=2/fullwave2_*_2_relax_isotropic_multi_gpu_sm_100_cuda129)
produced on a B200 for the seeded cases of tests/cases.py (generator: tools/make_ref_golden.py, run through
gpurun; the reference ships no golden vectors of its own).  The bar is BIT-EXACT."""

import json
from pathlib import Path

import numpy as np
import pytest

from oracle import oracle
from tests import cases

GOLDEN = Path(__file__).resolve().parent / "golden"


def load_golden(name):
    z = np.load(GOLDEN / f"ref_{name}.npz")
    assert json.loads(str(z["case"])) == json.loads(json.dumps(cases.CASES[name])), "case definition drifted"
    return z["genout"]


@pytest.mark.parametrize("name", sorted(cases.CASES))
def test_oracle_reproduces_reference_engine_bit_exactly(name):
    want = load_golden(name)
    got = oracle.run(cases.make(name))
    assert got.shape == want.shape
    assert np.abs(want).max() > 1.0            # the golden is a real wave, not zeros
    np.testing.assert_array_equal(got, want)


def test_3d_dcmap_rule_matters():
    """The reference 3D binary only loads the first nX*nY dcmap entries (oracle/fw25_oracle.c, dcmap_3d);
    honouring the whole map instead moves the traces by ~4e-4 rel-L2, i.e. the golden can tell."""
    pb = cases.make("het3d")
    pb.dcmap_full3d = True
    got = oracle.run(pb)
    want = load_golden("het3d")
    err = np.linalg.norm(got.astype(np.float64) - want) / np.linalg.norm(want.astype(np.float64))
    assert 1e-5 < err < 1e-2
