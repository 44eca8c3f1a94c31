"""Randomised parity: seeded random grids (ragged sizes down to the smallest the boundary layer allows), step counts,
recording periods, sensor / source / air lists with entries in the never-updated rim, duplicates and overlaps -- the
CUDA engine through the C-ABI must equal the oracle bit for bit on every one (oracle pinned by the reference's own
binary, tests/test_oracle_golden.py)."""

import numpy as np
import pytest

from fullwave25_b200 import engine, synthetic
from oracle import oracle

pytestmark = pytest.mark.gpu


def random_problem(seed):
    rng = np.random.default_rng(1000 + seed)
    ndim = 2 if seed % 2 == 0 else 3
    n_pml, n_trans = int(rng.integers(0, 4)), int(rng.integers(1, 4))
    nb = 8 + n_pml + n_trans
    span = 70 if ndim == 2 else 26
    shape = tuple(int(2 * nb + 3 + rng.integers(0, span)) for _ in range(ndim))
    nT = int(rng.integers(3, 70))
    pb = synthetic.make_problem(shape, nT=nT, modT=int(rng.integers(1, 10)), seed=seed, n_pml=n_pml, n_trans=n_trans,
                                n_sensors=int(rng.integers(0, 40)), n_air=int(rng.integers(0, 6)),
                                block=int(rng.integers(2, 7)), source_layers=int(rng.integers(1, 4)))
    anywhere = lambda n: np.stack([rng.integers(0, s, n) for s in shape], axis=1).astype(np.int32)  # noqa: E731
    pb.outc = np.vstack([pb.outc, anywhere(6), pb.outc[:3], pb.icc[:2]]).astype(np.int32)       # rim, duplicates, on sources
    if seed % 3 == 0:        # a few sources anywhere (rim included), driven by a copy of the first signals
        extra = anywhere(4)
        pb.icc = np.vstack([pb.icc, extra]).astype(np.int32)
        pb.icmat = np.vstack([pb.icmat, pb.icmat[:4]]).astype(np.float32)
    if seed % 4 == 1:
        pb.icczero = np.vstack([pb.icczero, anywhere(3), pb.icc[:1]]).astype(np.int32)           # air in the rim / on a source
    pb.dcmap_full3d = bool(seed % 5 == 0)
    return pb


@pytest.mark.parametrize("seed", range(24))
def test_random_problem_matches_oracle_bit_exact(built_lib, seed, monkeypatch):
    pb = random_problem(seed)
    if seed % 7 == 3:
        monkeypatch.setenv("FW25_GRAPH", "0")
    if seed % 6 == 2:
        monkeypatch.setenv("FW25_FUSE2D", "1")
    want = oracle.run(pb)
    got, stats = engine.run(pb)
    np.testing.assert_array_equal(got, want)
    assert stats["point_updates"] == pb.n_points * pb.nT


@pytest.mark.parametrize("seed", range(100, 112))
def test_random_anisotropic_and_sharded_problems(built_lib, seed, monkeypatch):
    """The same kind of random problems for the anisotropic-relaxation family (per-axis maps: the ANISO warp-specialised
    3D sweeps and the ANISO tiled 2D sweeps) and for x-sharded runs (2 or 3 slabs through fw25_run's native runner, fused
    push or copies): bit-exact against the oracle."""
    rng = np.random.default_rng(seed)
    ndim = 2 if seed % 2 == 0 else 3
    aniso = seed % 3 != 0
    n_slabs = 1 + seed % 3
    n_pml, n_trans = int(rng.integers(0, 4)), int(rng.integers(1, 4))
    nb = 8 + n_pml + n_trans
    span = 60 if ndim == 2 else 24
    shape = [int(2 * nb + 3 + rng.integers(0, span)) for _ in range(ndim)]
    shape[0] = max(shape[0], 17 * n_slabs + 2 * nb)                     # slabs of >= 16 planes
    pb = synthetic.make_problem(tuple(shape), nT=int(rng.integers(5, 50)), modT=int(rng.integers(1, 6)), seed=seed,
                                n_pml=n_pml, n_trans=n_trans, n_sensors=int(rng.integers(4, 60)),
                                n_air=0 if aniso else int(rng.integers(0, 6)), block=int(rng.integers(2, 7)),
                                source_layers=int(rng.integers(1, 4)), aniso=aniso)
    anywhere = lambda n: np.stack([rng.integers(0, s, n) for s in shape], axis=1).astype(np.int32)  # noqa: E731
    pb.outc = np.vstack([pb.outc, anywhere(8), pb.icc[:2]]).astype(np.int32)
    pb.dcmap_full3d = bool(seed % 4 == 0)
    monkeypatch.setenv("FW25_FUSED_HALO", "1" if seed % 2 else "0")
    want = oracle.run(pb)
    got, stats = engine.run(pb, device_ids=(0,) * n_slabs)
    assert stats["n_devices"] == n_slabs and np.abs(want).max() > 0
    np.testing.assert_array_equal(got, want)
