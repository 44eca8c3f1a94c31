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


def load_golden(name, suffix=""):
    z = np.load(GOLDEN / f"ref_{name}{suffix}.npz")
    assert json.loads(str(z["case"])) == json.loads(json.dumps(cases.CASES.get(name) or cases.CASES_BIG[name])), \
        "case definition drifted"
    return z["genout"]


@pytest.mark.parametrize("name", sorted(cases.CASES))
def test_oracle_reproduces_reference_engine_bit_exactly(name):
    want = load_golden(name)
    got = oracle.run(cases.make(name))
    assert got.shape == want.shape
    assert np.abs(want).max() > 1.0            # the golden is a real wave, not zeros
    np.testing.assert_array_equal(got, want)


G2_CASES = ("far2d", "far3d", "het2d", "het2d_ragged", "het3d", "het3d_long")


@pytest.mark.parametrize("name", G2_CASES)
def test_reference_two_gpu_goldens_are_explained_by_its_two_sharding_deviations(name):
    """tests/golden/ref_<case>_g2.npz: the reference binary run with CUDA_VISIBLE_DEVICES=0,1 (its in-process x-slab
    mode) on a B200 pair.  They are NOT equal to the reference's own 1-GPU traces (2e-5 .. 0.37 rel-L2).  The CPU
    emulation oracle/ref_multigpu.py -- one-domain arithmetic plus the two deviations found in the binary (sensors of
    GPUs >= 1 read plane x-1; sources / air voxels applied in owned planes only) -- reproduces all of them
    BIT-EXACTLY, which pins: global `outc` frame order, a correctly exchanged wave field, and those two deviations.
    The B200 engine reproduces neither: its N-GPU runs equal one domain (tests/test_slab.py, tests/test_multi_gpu.py)."""
    from oracle import ref_multigpu
    g2, g1 = load_golden(name, "_g2"), load_golden(name)
    pb = cases.make(name)
    assert g2.shape == g1.shape and not np.array_equal(g2, g1)
    assert len(set((pb.outc[:, 0] >= (pb.nX + 1) // 2).tolist())) == 2      # sensors on both sides of the interface
    np.testing.assert_array_equal(ref_multigpu.run(pb, 2), g2)
    # the sensor deviation is needed everywhere; without either deviation the emulation is the one-domain oracle
    # (checked where no source sits in a ghost plane: the slab steppers' rim rule is local)
    assert not np.array_equal(ref_multigpu.run(pb, 2, sensor_shift=0), g2)
    if name in cases.CASES_2GPU:
        np.testing.assert_array_equal(ref_multigpu.run(pb, 2, owned_only_injection=False, sensor_shift=0), g1)


def test_3d_dcmap_rule_matters():
    """The reference 3D binary only loads the first nX*nY dcmap entries (oracle/fw25_oracle.c, dcmap_3d);
    honouring the whole map instead moves the traces by ~4e-4 rel-L2, i.e. the golden can tell."""
    pb = cases.make("het3d")
    pb.dcmap_full3d = True
    got = oracle.run(pb)
    want = load_golden("het3d")
    err = np.linalg.norm(got.astype(np.float64) - want) / np.linalg.norm(want.astype(np.float64))
    assert 1e-5 < err < 1e-2


def test_oracle_reproduces_reference_engine_at_a_baseline_size():
    """BASELINE.json configs[2]'s grid (1457 x 2178, the convex-transducer example), 480 steps, 400 sensors in the part
    of the domain the pulse reaches: the oracle equals the reference's 2D sm_100 binary bit for bit at full size too.
    (The 280^3 golden, configs[3], is checked against the CUDA engine in tests/test_gpu_parity.py and against the
    oracle by tools/make_ref_golden.py on the GPU box: 7 G point-updates are too slow for this suite.)"""
    want = load_golden("big2d")
    got = oracle.run(cases.make("big2d"))
    assert got.shape == want.shape == (120, 400) and np.abs(want).max() > 1.0
    assert (want != 0).mean() > 0.5
    np.testing.assert_array_equal(got, want)
