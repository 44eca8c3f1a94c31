"""Host-side input handling: the reference's .dat protocol round trip and validation
(/root/reference/fullwave/solver/input_file_writer.py:563-881)."""

import numpy as np
import pytest

from fullwave25_b200.problem import MAP_NAMES, Problem
from tests import cases


@pytest.mark.parametrize("name", ["het3d_ragged", "het2d_ragged"])
def test_dat_dir_round_trip(tmp_path, name):
    pb = cases.make(name)
    d = pb.to_dat_dir(tmp_path / "txrx_0")
    # the files the reference binary opens
    expected = {"nX", "nY", "nT", "nTic", "modT", "dX", "dY", "dT", "c0", "ncoords", "ncoordszero",
                "ncoordsout", "ndmap", "d", "dmap", "dcmap", "c", "rho", "K", "beta", "kappax", "kappau",
                "apmlu1", "bpmlu1", "apmlx1", "bpmlx1", "apmlu2", "bpmlu2", "apmlx2", "bpmlx2", "icc",
                "outc", "icmat", "icczero"}
    if pb.ndim == 3:
        expected |= {"nZ", "dZ"}
    assert expected <= {f.stem for f in d.glob("*.dat")}
    assert (d / "icc.dat").stat().st_size == pb.ncoords * pb.ndim * 4
    assert (d / "icmat.dat").stat().st_size == pb.ncoords * pb.nTic * 4
    back = Problem.from_dat_dir(d)
    assert back.shape == pb.shape and back.ndim == pb.ndim
    for k in ("nT", "nTic", "modT", "ndmap"):
        assert getattr(back, k) == getattr(pb, k)
    assert np.float32(back.dX) == np.float32(pb.dX) and np.float32(back.dT) == np.float32(pb.dT)
    for k in MAP_NAMES + ("dmap", "dcmap", "icc", "icmat", "outc", "icczero"):
        np.testing.assert_array_equal(getattr(back, k), getattr(pb, k), err_msg=k)


def test_index_maps_follow_row_major_mask_order():
    """icc / outc rows are (x, y[, z]) in np.where order (fullwave/utils/coordinates.py:41-50) -- a bit-exact contract."""
    pb = cases.make("het3d")
    lin = np.ravel_multi_index(pb.outc.T, pb.shape)
    assert (np.diff(lin) > 0).all()
    lin = np.ravel_multi_index(pb.icc.T, pb.shape)
    assert (np.diff(lin) > 0).all()


def test_validation_errors():
    pb = cases.make("het2d")
    pb.dcmap = pb.dcmap.copy()
    pb.dcmap[10, 10] = pb.ndmap
    with pytest.raises(ValueError):
        pb.normalise()
    pb = cases.make("het2d")
    pb.rho = pb.rho[:-1]
    with pytest.raises(ValueError):
        pb.normalise()
    pb = cases.make("het2d")
    pb.modT = 0
    with pytest.raises(ValueError):
        pb.normalise()


def test_slab_views_share_memory():
    pb = cases.make("het3d")
    s = pb.slab(10, 30)
    assert s.nX == 20 and s.rho.base is not None and np.shares_memory(s.rho, pb.rho)
    assert s.icc is pb.icc


def test_from_fullwave_objects_matches_protocol_fields():
    """Duck-typed reference objects -> Problem, i.e. what InputFileWriter would write."""
    from types import SimpleNamespace
    rng = np.random.default_rng(0)
    shape = (30, 34)
    c = rng.uniform(1450, 1600, shape)
    rho = rng.uniform(950, 1100, shape)
    relax = {k: rng.uniform(0.1, 1.0, shape) for k in
             ("kappa_x", "kappa_u", "a_pml_x1", "b_pml_x1", "a_pml_x2", "b_pml_x2",
              "a_pml_u1", "b_pml_u1", "a_pml_u2", "b_pml_u2")}
    air = np.zeros(shape, bool)
    air[12, 13] = air[20, 5] = True
    smask = np.zeros(shape, bool)
    smask[10, 10:20] = True
    grid = SimpleNamespace(is_3d=False, nx=30, ny=34, nz=0, nt=50, dx=1e-4, dy=1e-4, dz=0, dt=1.2e-8,
                           c0=1540.0, cfl=0.2)
    medium = SimpleNamespace(sound_speed=c, density=rho, beta=np.full(shape, 3.5), bulk_modulus=c**2 * rho,
                             air_map=air, relaxation_param_dict_for_fw2=relax)
    source = SimpleNamespace(incoords=np.stack(np.nonzero(smask), 1), icmat=rng.normal(size=(10, 50)))
    sensor = SimpleNamespace(outcoords=np.array([[15, 15], [16, 20]]), sampling_modulus_time=2)
    pb = Problem.from_fullwave_objects(grid, medium, source, sensor)
    assert pb.ncoordszero == 2 and pb.ncoords == 10 and pb.modT == 2 and pb.nTic == 50
    np.testing.assert_array_equal(pb.kappax, relax["kappa_x"].astype(np.float32))
    np.testing.assert_array_equal(pb.apmlu2, relax["a_pml_u2"].astype(np.float32))
    np.testing.assert_array_equal(pb.icczero, [[12, 13], [20, 5]])
    assert pb.dcmap.min() == 0 and pb.dcmap.max() == pb.ndmap - 1


def test_anisotropic_file_set_round_trip(tmp_path):
    """The anisotropic protocol (upstream use_isotropic_relaxation=False, input_file_writer.py:592-620): per-axis
    kappa / a / b files next to the isotropic-named x-axis members; detected by kappay.dat."""
    import numpy as np

    from fullwave25_b200 import synthetic
    from fullwave25_b200.problem import Problem
    for shape in ((48, 52), (40, 42, 44)):
        pb = synthetic.make_problem(shape, nT=10, aniso=True, n_air=0)
        stems = Problem.aniso_stems(pb.ndim)
        assert len(stems) == (4 + 16 if pb.ndim == 2 else 6 + 24) and set(pb.aniso) == set(stems)
        assert np.array_equal(pb.aniso["kappax"], pb.kappax) and not np.array_equal(pb.aniso["kappay"], pb.kappax)
        d = pb.to_dat_dir(tmp_path / f"an{pb.ndim}")
        assert (d / "apmlw2.dat").exists() and ((d / "kappav.dat").exists() == (pb.ndim == 3))
        back = Problem.from_dat_dir(d)
        assert back.aniso is not None
        for k in stems:
            np.testing.assert_array_equal(back.aniso[k], pb.aniso[k])
        iso = Problem.from_dat_dir(synthetic.make_problem(shape, nT=10).to_dat_dir(tmp_path / f"iso{pb.ndim}"))
        assert iso.aniso is None


def test_problem_for_device_resident_maps():
    """`Problem.for_device_maps`: only step counts, the stencil table and the point lists live on the host; the 13 maps
    and dcmap are None and `normalise` leaves them alone (they sit in HBM, mapgen.MapSet)."""
    from types import SimpleNamespace as NS
    dmap = np.arange(9 * 2 * 3, dtype=np.float32).reshape(9, 2, 3)
    ms = NS(shape=(30, 32, 34), ndmap=3, dmap=dmap, d_table=np.zeros((9, 2)))
    grid = NS(nt=17, dx=1.25e-4, dt=1.5e-8, c0=1540.0)
    src = NS(incoords=np.array([[9, 9, 9], [9, 9, 10]]), icmat=np.ones((2, 5)))
    sen = NS(outcoords=np.array([[12, 13, 14]]), sampling_modulus_time=4)
    pb = Problem.for_device_maps(ms, grid, src, sen, icczero=np.array([[20, 20, 20]]))
    assert pb.rho is None and pb.dcmap is None and pb.dcmap_full3d
    assert (pb.ndim, pb.nX, pb.nY, pb.nZ, pb.nT, pb.nTic, pb.modT, pb.ndmap) == (3, 30, 32, 34, 17, 5, 4, 3)
    assert pb.icc.dtype == np.int32 and pb.icmat.dtype == np.float32 and pb.outc.shape == (1, 3)
    assert pb.icczero.tolist() == [[20, 20, 20]] and pb.n_frames == 5
    boxed = Problem.for_device_maps(ms, grid, src, NS(outcoords=None, sampling_modulus_time=2), out_box=(0, 0, 0, 30, 32, 34))
    assert boxed.ncoordsout == 30 * 32 * 34 and boxed.outc.shape == (0, 3) and boxed.ncoordszero == 0
    air = np.zeros((30, 32, 34), bool)
    air[3, 4, 5] = True
    assert Problem.for_device_maps(ms, grid, src, sen, air_map=air).icczero.tolist() == [[3, 4, 5]]


def test_parallel_cast_equals_numpy_cast():
    """`cast_c` (row blocks converted by a thread pool) returns exactly what `np.ascontiguousarray(a, dtype)` returns."""
    from fullwave25_b200 import problem
    rng = np.random.default_rng(7)
    big = rng.standard_normal((3000, 3001)) * 1e3                 # above the parallel threshold
    assert big.size >= problem._CAST_PARALLEL_MIN
    cases = [big, np.asfortranarray(big), big[::2, 1::3], big.astype(np.float32)[:, ::2],
             rng.integers(-5, 2**31 - 1, (3_000_000, 3)), rng.standard_normal((5, 7)), np.float64(3.5),
             rng.standard_normal((1, 9_000_000))]
    for a in cases:
        for dt in (np.float32, np.int32):
            if dt is np.int32 and np.asarray(a).dtype.kind == "f":
                continue
            got = problem.cast_c(a, dt)
            want = np.ascontiguousarray(a, dtype=dt)
            assert got.dtype == want.dtype and got.shape == want.shape and got.flags["C_CONTIGUOUS"]
            np.testing.assert_array_equal(got, want)
    same = np.zeros((4, 5), np.float32)
    assert problem.cast_c(same, np.float32) is same               # nothing to do: no copy
