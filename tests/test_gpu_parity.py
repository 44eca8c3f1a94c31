"""Parity tests proper: the CUDA engine, called through the C-ABI (libfw25.so), against the CPU oracle
on the same seeded inputs.  Arithmetic is IEEE-identical operation by operation, so the bar is
BIT-EXACT sensor traces and final fields (the north-star tolerance is rel-L2 <= 1e-5; we assert 0)."""

import numpy as np
import pytest

from fullwave25_b200 import engine
from oracle import oracle
from tests import cases

pytestmark = pytest.mark.gpu

ALL = sorted(cases.CASES)


def rel_l2(a, b):
    d = np.linalg.norm(a.astype(np.float64) - b.astype(np.float64))
    n = np.linalg.norm(b.astype(np.float64))
    return d / n if n else d


@pytest.mark.parametrize("name", ALL)
def test_run_matches_oracle_bit_exact(built_lib, name):
    pb = cases.make(name)
    want = oracle.run(pb)
    got, stats = engine.run(pb)
    assert got.shape == want.shape
    assert np.isfinite(got).all()
    assert rel_l2(got, want) <= 1e-5          # the stated FP32 tolerance
    np.testing.assert_array_equal(got, want)  # and in fact identical bits
    assert stats["kernel_launches"] >= 2 * pb.nT
    assert stats["point_updates"] == pb.n_points * pb.nT


@pytest.mark.parametrize("name", ALL)
def test_run_matches_reference_golden_bit_exact(built_lib, name):
    """CUDA engine vs the sensor traces of the reference's own sm_100 binary (tests/golden/ref_*.npz)."""
    from tests.test_oracle_golden import load_golden
    want = load_golden(name)
    got, _ = engine.run(cases.make(name))
    assert rel_l2(got, want) <= 1e-5
    np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("variant", [1, 2, 3])
@pytest.mark.parametrize("name", ["het3d", "het2d", "het3d_ragged", "het3d_long"])
def test_final_fields_match_oracle(built_lib, name, variant):
    """variant 1 = simple sweeps, 2 = TMA-tiled (3D: x-marching; 2D: row tiles), 3 = warp-specialised all-TMA (3D only)."""
    pb = cases.make(name)
    if variant == 3 and pb.ndim == 2:
        pytest.skip("the warp-specialised sweeps are 3D")
    _, want = oracle.run(pb, return_fields=True)
    with engine.Engine(pb, variant=variant) as e:
        e.step(pb.nT)
        e.sync()
        for k in ("p", "u", "v") + (("w",) if pb.ndim == 3 else ()):
            np.testing.assert_array_equal(e.field(k), want[k], err_msg=k)


@pytest.mark.parametrize("name", ["het3d", "het2d"])
def test_piecewise_sweeps_equal_full_sweeps(built_lib, name):
    """fw25_sweep_u / fw25_sweep_p over sub-ranges (the boundary-first schedule) == whole sweeps."""
    pb = cases.make(name)
    nX = pb.nX
    with engine.Engine(pb) as a, engine.Engine(pb) as b:
        a.step(15)
        for t in range(15):
            b.inject(t)
            for lo, hi in ((nX - 16, nX), (0, 16), (16, nX - 16)):
                b.sweep_u(lo, hi)
            for lo, hi in ((nX // 2, nX), (0, nX // 2)):
                b.sweep_p(lo, hi)
            if t % pb.modT == 0:
                b.record(t // pb.modT)
        a.sync(); b.sync()
        for k in ("p", "u", "v"):
            np.testing.assert_array_equal(a.field(k), b.field(k), err_msg=k)
        nf = -(-15 // pb.modT)
        np.testing.assert_array_equal(a.read_frames(0, nf), b.read_frames(0, nf))


def test_edge_cases(built_lib):
    # no sensors, no air, no sources after nTic, nT not a multiple of modT
    pb = cases.make("het2d_ragged")
    pb.outc = pb.outc[:0]
    got, _ = engine.run(pb)
    assert got.shape == (pb.n_frames, 0)
    pb = cases.make("het3d_ragged")
    pb.icczero = pb.icczero[:0]
    np.testing.assert_array_equal(engine.run(pb)[0], oracle.run(pb))
    # sensors in the 8-cell rim read 0; duplicate sensors are allowed
    pb = cases.make("het2d")
    pb.outc = np.vstack([pb.outc, [[3, 40], [40, 2]], pb.outc[:2]]).astype(np.int32)
    got = engine.run(pb)[0]
    np.testing.assert_array_equal(got, oracle.run(pb))
    assert not got[:, -4:-2].any()
    # zero steps
    pb = cases.make("het2d")
    pb.nT = 0
    assert engine.run(pb)[0].shape == (0, pb.ncoordsout)
    # a source that is also an air voxel: zeroing follows injection in the reference's launch order, so it reads 0
    for name in ("het2d", "het3d"):
        pb = cases.make(name)
        pb.icczero = np.vstack([pb.icczero, pb.icc[:3]]).astype(np.int32)
        pb.outc = np.vstack([pb.outc, pb.icc[:4]]).astype(np.int32)
        got = engine.run(pb)[0]
        np.testing.assert_array_equal(got, oracle.run(pb))


@pytest.mark.parametrize("name", ["het2d_long", "het3d_long", "het2d_ragged", "hom3d"])
def test_graph_replayed_steps_equal_launched_steps(built_lib, name, monkeypatch):
    """Whole steps replayed from a CUDA graph (device-side step counter) == the same steps launched one by one:
    sensor frames (incl. a run that stops inside a recording period) and final fields."""
    pb = cases.make(name)
    pb.nT -= 3
    out = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("FW25_GRAPH", mode)
        g, st = engine.run(pb)
        with engine.Engine(pb) as e:
            e.step(7)                      # a partial period first, then the rest
            e.step(pb.nT - 7)
            e.sync()
            out[mode] = (g, st["kernel_launches"], e.read_frames(0, pb.n_frames), {k: e.field(k) for k in "puv"})
    np.testing.assert_array_equal(out["0"][0], out["1"][0])
    np.testing.assert_array_equal(out["0"][2], out["1"][2])
    np.testing.assert_array_equal(out["0"][0], out["0"][2])
    for k in "puv":
        np.testing.assert_array_equal(out["0"][3][k], out["1"][3][k], err_msg=k)
    np.testing.assert_array_equal(out["1"][0], oracle.run(pb))
    assert out["1"][1] >= 2 * pb.nT


def test_errors_are_reported_not_swallowed(built_lib):
    pb = cases.make("het2d")
    pb.outc = pb.outc.copy()
    pb.outc[0, 0] = pb.nX + 5
    with pytest.raises(engine.EngineError, match="outside the grid"):
        engine.run(pb)
    pb = cases.make("het2d")
    with pytest.raises(engine.EngineError):
        engine.run(pb, device_ids=(97,))


def test_mid_size_ragged_grid_all_variants_and_oracle(built_lib):
    """A grid that is not a multiple of any tile size (97 x 203 x 269, 5.3 M points), heterogeneous stencil table
    honoured per voxel (dcmap_full3d): the three sweep implementations and the oracle give identical fields."""
    from fullwave25_b200 import synthetic
    pb = synthetic.make_problem((97, 203, 269), nT=24, modT=3, seed=21, n_pml=10, n_trans=6, block=7, n_sensors=200, n_air=50)
    pb.dcmap_full3d = True
    want_g, want = oracle.run(pb, return_fields=True)
    for variant in (1, 2, 3):
        with engine.Engine(pb, variant=variant) as e:
            e.step(pb.nT)
            e.sync()
            for k in "puvw":
                np.testing.assert_array_equal(e.field(k), want[k], err_msg=f"variant {variant} field {k}")
            np.testing.assert_array_equal(e.read_frames(0, pb.n_frames), want_g, err_msg=f"variant {variant}")


@pytest.mark.parametrize("shape", [(203, 269), (61, 1031), (300, 60)])
def test_2d_ragged_grids_tiled_and_simple_sweeps(built_lib, shape, monkeypatch):
    """2D grids that are not multiples of the 128-column / 4- or 8-row tiles: TMA-tiled 2D sweeps (both tile
    heights) == one-thread-per-cell sweeps == oracle, frames and final fields."""
    from fullwave25_b200 import synthetic
    pb = synthetic.make_problem(shape, nT=90, modT=4, seed=31, n_pml=9, n_trans=5, block=7, n_sensors=120, n_air=20)
    want_g, want = oracle.run(pb, return_fields=True)
    # (variant, rows per thread of the marching tiles, rows per CTA of the one-cell-per-thread tiles)
    for variant, rpt, tr in ((1, "0", "0"), (2, "1", "0"), (2, "2", "0"), (2, "4", "0"), (2, "8", "0"),
                             (2, "0", "2"), (2, "0", "4"), (2, "0", "8"), (0, "0", "0")):
        monkeypatch.setenv("FW25_2D_RPT", rpt)
        monkeypatch.setenv("FW25_2D_TR", tr)
        with engine.Engine(pb, variant=variant) as e:
            e.step(pb.nT)
            e.sync()
            tag = f"variant {variant} rpt {rpt} tr {tr}"
            for k in "puv":
                np.testing.assert_array_equal(e.field(k), want[k], err_msg=f"{tag} field {k}")
            np.testing.assert_array_equal(e.read_frames(0, pb.n_frames), want_g, err_msg=tag)


@pytest.mark.parametrize("name", sorted(cases.CASES_ANISO))
def test_anisotropic_file_set(built_lib, name, tmp_path):
    """Per-axis relaxation maps (the reference's second engine family): final fields == oracle; the executable
    reads the anisotropic .dat set; with identical maps on every axis the result equals the isotropic run (and the
    engine then runs its isotropic kernels); air voxels are ignored like the reference's anisotropic binaries do."""
    import subprocess

    from fullwave25_b200.build import CLI
    pb = cases.make(name)
    want_g, want = oracle.run(pb, return_fields=True)
    with engine.Engine(pb) as e:
        e.step(pb.nT)
        e.sync()
        for k in ("p", "u", "v") + (("w",) if pb.ndim == 3 else ()):
            np.testing.assert_array_equal(e.field(k), want[k], err_msg=k)
        np.testing.assert_array_equal(e.read_frames(0, pb.n_frames), want_g)
    d = pb.to_dat_dir(tmp_path / "sim")
    r = subprocess.run([str(CLI)], cwd=d, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    np.testing.assert_array_equal(np.fromfile(d / "genout.dat", np.float32).reshape(want_g.shape), want_g)
    # identical axes -> the isotropic answer
    kw = dict(cases.CASES[name]); kw.pop("aniso")
    iso = synthetic_iso = __import__("fullwave25_b200.synthetic", fromlist=["x"]).make_problem(**kw)
    same = __import__("fullwave25_b200.synthetic", fromlist=["x"]).make_problem(**kw)
    vel = ("x", "y", "z")[: same.ndim]
    prs = ("u", "w") if same.ndim == 2 else ("u", "v", "w")
    same.aniso = {}
    for letters, fam in ((vel, "x"), (prs, "u")):
        for l in letters:
            same.aniso["kappa" + l] = getattr(same, "kappa" + fam).copy()
            for nu in (1, 2):
                for ab in "ab":
                    same.aniso[f"{ab}pml{l}{nu}"] = getattr(same, f"{ab}pml{fam}{nu}").copy()
    same.icczero = np.array([[20, 21, 22][: same.ndim]], np.int32)      # ignored by the anisotropic family
    got_same, st = engine.run(same.normalise())
    np.testing.assert_array_equal(got_same, engine.run(iso)[0])
    np.testing.assert_array_equal(got_same, oracle.run(same))


@pytest.mark.parametrize("variant", [0, 1, 3])
def test_anisotropic_3d_on_the_warp_specialised_sweeps(built_lib, variant):
    """Truly per-axis kappa / a / b maps run on the TMA warp-specialised sweeps (k_sweep_*_ws<.., ANISO = true>: 28 / 26
    point-wise tiles per plane, 12- / 10-row tiles) by default; variant 1 forces the L1/L2-path kernels.  All equal the
    oracle AND the reference's anisotropic sm_100 binary (tests/golden/ref_aniso3d.npz) bit for bit -- also on a grid
    whose y / z extents are ragged against both tile heights and span several tiles and two x-chunks."""
    from fullwave25_b200 import synthetic
    from tests.test_oracle_golden import load_golden
    pb = cases.make("aniso3d")
    with engine.Engine(pb, variant=variant) as e:      # variant 3 raises if no warp-specialised plan exists
        e.step(pb.nT)
        e.sync()
        np.testing.assert_array_equal(e.read_frames(0, pb.n_frames), load_golden("aniso3d"))
    big = synthetic.make_problem((75, 67, 81), nT=36, modT=3, seed=23, aniso=True, n_air=0, n_sensors=300)
    want_g, want = oracle.run(big, return_fields=True)
    with engine.Engine(big, variant=variant) as e:
        e.step(big.nT)
        e.sync()
        for k in "puvw":
            np.testing.assert_array_equal(e.field(k), want[k], err_msg=k)
        np.testing.assert_array_equal(e.read_frames(0, big.n_frames), want_g)
    assert np.abs(want_g).max() > 0


@pytest.mark.parametrize("variant", [0, 1, 2])
def test_anisotropic_2d_on_the_tiled_sweeps(built_lib, variant):
    """2D per-axis maps: the TMA-tiled one-cell-per-thread sweeps (k_sweep_*_2dc<2, .., ANISO = true>, default and
    variant 2) and the L1/L2-path kernels (variant 1) equal the oracle and the reference's anisotropic 2D binary
    (tests/golden/ref_aniso2d.npz), also on a ragged grid of several column tiles, with graph replay on and off."""
    from fullwave25_b200 import synthetic
    from tests.test_oracle_golden import load_golden
    pb = cases.make("aniso2d")
    with engine.Engine(pb, variant=variant) as e:
        e.step(pb.nT)
        e.sync()
        np.testing.assert_array_equal(e.read_frames(0, pb.n_frames), load_golden("aniso2d"))
    big = synthetic.make_problem((147, 301), nT=90, modT=3, seed=25, aniso=True, n_air=0, n_sensors=300)
    want_g, want = oracle.run(big, return_fields=True)
    with engine.Engine(big, variant=variant) as e:
        e.step(big.nT)
        e.sync()
        for k in "puv":
            np.testing.assert_array_equal(e.field(k), want[k], err_msg=k)
        np.testing.assert_array_equal(e.read_frames(0, big.n_frames), want_g)
    got, _ = engine.run(big)                       # the whole-job path (graph-replayed steps)
    np.testing.assert_array_equal(got, want_g)
    assert np.abs(want_g).max() > 0


@pytest.mark.parametrize("fused", ["1", "0"])
def test_anisotropic_3d_on_two_slabs(built_lib, fused, monkeypatch):
    """The anisotropic family sharded over two x-slabs (fw25_run with a device list; both slabs on device 0 when the
    box has one GPU): boundary sweeps of the ANISO warp-specialised kernels push their halo planes -- identical to one
    domain and to the oracle."""
    from fullwave25_b200 import synthetic
    from tests.test_multi_gpu import _devices
    monkeypatch.setenv("FW25_FUSED_HALO", fused)
    pb = synthetic.make_problem((96, 52, 50), nT=40, modT=2, seed=24, aniso=True, n_air=0, n_sensors=200)
    one, _ = engine.run(pb)
    two, stats = engine.run(pb, device_ids=_devices())
    assert stats["n_devices"] == 2 and stats["halo_bytes"] > 0 and np.abs(one).max() > 0
    np.testing.assert_array_equal(two, one)
    np.testing.assert_array_equal(one, oracle.run(pb))


def test_large_host_maps_take_the_staged_upload(built_lib, monkeypatch):
    """Host maps of >= 32 MB whose rows are not a multiple of 32 floats go up densely into two staging buffers and are
    re-pitched on the device (Engine::upload_dense_rows); with 8 MB chunks every map needs five of them."""
    from fullwave25_b200 import synthetic
    monkeypatch.setenv("FW25_STAGE_MB", "8")
    pb = synthetic.make_problem((216, 200, 203), nT=6, modT=2, seed=41, n_pml=12, n_trans=8, n_sensors=200)
    assert pb.rho.nbytes >= 32 << 20 and pb.nZ % 32
    want = oracle.run(pb)
    got, stats = engine.run(pb)
    np.testing.assert_array_equal(got, want)
    assert np.abs(want).max() > 0 and stats["h2d_bytes"] >= 14 * pb.rho.nbytes


@pytest.mark.parametrize("name", ["het2d", "het2d_long", "het2d_ragged", "hom2d", "far2d"])
@pytest.mark.parametrize("sensors", ["listed", "box"])
def test_fused_2d_step_equals_separate_kernels(built_lib, name, sensors, monkeypatch):
    """Graph-replayed 2D steps fold record(t) and inject(t + 1) into fd_p(t) (k_sweep_p_2dc<2, true>): two kernels per
    step instead of up to four, same bits as the separate kernels and the oracle -- including sensors that sit on
    source cells and air voxels, rim sensors, duplicate sensors and runs that stop inside a recording period."""
    pb = cases.make(name)
    pb.nT -= 3
    if sensors == "box":
        pb.outc = np.stack(np.meshgrid(np.arange(2, pb.nX - 5), np.arange(0, pb.nY), indexing="ij"), -1).reshape(-1, 2)
    else:
        extra = [pb.icc[:5], pb.icc[-3:], pb.outc[:2], [[3, 40], [40, 2]]]
        if pb.ncoordszero:
            extra.append(pb.icczero[:4])
            pb.icczero = np.vstack([pb.icczero, pb.icc[7:9]])          # sources that are air voxels too
        pb.outc = np.vstack([pb.outc] + extra)
    pb.outc = pb.outc.astype(np.int32)
    pb.icczero = pb.icczero.astype(np.int32)
    out = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("FW25_FUSE2D", mode)
        out[mode] = engine.run(pb)
    np.testing.assert_array_equal(out["1"][0], out["0"][0])
    np.testing.assert_array_equal(out["1"][0], oracle.run(pb))
    assert out["1"][1]["kernel_launches"] < 0.75 * out["0"][1]["kernel_launches"]
    assert np.abs(out["1"][0]).max() > 0
    # several transmit events on one engine: the lists follow the new sources
    with engine.Engine(pb) as e:
        first, _ = e.run()
        e.reset(pb.icc[::2], pb.icmat[::2] * 0.5)
        second, _ = e.run()
    np.testing.assert_array_equal(first, out["1"][0])
    half = cases.make(name)
    half.nT, half.outc, half.icczero = pb.nT, pb.outc, pb.icczero
    half.icc, half.icmat = pb.icc[::2], pb.icmat[::2] * np.float32(0.5)
    np.testing.assert_array_equal(second, oracle.run(half))


@pytest.mark.parametrize("name", ["big2d", "big3d"])
def test_baseline_size_grids_match_reference_goldens(built_lib, name):
    """BASELINE.json configs[2] / configs[3] grids (1457 x 2178, 280^3) with the reference Solver's boundary layer:
    the engine's default kernels (TMA-tiled 2D sweeps / warp-specialised 3D sweeps, graph replay as configured) equal
    the reference's sm_100 binaries bit for bit -- tests/golden/ref_big{2d,3d}.npz, tools/make_ref_golden.py."""
    from tests.test_oracle_golden import load_golden
    want = load_golden(name)
    got, stats = engine.run(cases.make(name))
    assert got.shape == want.shape and np.abs(want).max() > 1.0
    np.testing.assert_array_equal(got, want)
