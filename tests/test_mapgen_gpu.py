"""GPU: the CUDA map builder (fw25_mapgen, through the C-ABI) against the reference's golden .dat bytes and the oracle,
and the engine run on device-generated maps against the engine run on host-built maps.

Tolerances (same as tests/test_mapgen_oracle.py): rho, K, beta, kappa*, dcmap bit-exact; a / b within 1 float32 ulp
(CUDA's float64 exp() vs numpy's); sensor traces of the two engine runs within 1e-5 relative L2 (north star)."""

from pathlib import Path

import numpy as np
import pytest

from fullwave25_b200 import engine, mapgen
from fullwave25_b200.problem import MAP_NAMES, Problem
from oracle import mapgen_oracle as mo
from tests import mapgen_cases as mc
from tests.test_mapgen_oracle import AB, EXACT, check_against_golden, oracle_maps, spec_of

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"


def device_maps_of(ms) -> dict:
    return {stem: ms.read(stem) for stem in MAP_NAMES + ("dcmap",)}


@pytest.mark.parametrize("name", list(mc.CASES))
def test_cuda_maps_match_reference_goldens_and_oracle(name):
    _, g, want = oracle_maps(name)
    with mapgen.MapSet(spec_of(name, g)) as ms:
        got = device_maps_of(ms)
        assert ms.shape == tuple(g["ext_shape"])
    check_against_golden(g, got)
    for stem in EXACT:
        assert np.array_equal(got[stem], want[stem]), stem
    for stem in AB:
        d = mo.ulp_distance_f32(got[stem], want[stem])
        assert d.max() <= 1, f"{stem}: {d.max()} ulp vs the oracle"


def test_reference_3d_dcmap_truncation_is_applied_on_request():
    g = np.load(GOLD / "mapgen_m3d.npz")
    spec = spec_of("m3d", g)
    spec.dcmap_full3d = False
    with mapgen.MapSet(spec) as ms:
        dc = ms.read("dcmap")
    full = g["dcmap"].copy().reshape(-1)
    full[g["ext_shape"][0] * g["ext_shape"][1]:] = 0
    assert np.array_equal(dc.reshape(-1), full) and full.any()


@pytest.mark.parametrize("name", ["m3d_big", "m2d_big", "m3d_lut", "m3d_onecell"])
def test_slab_maps_equal_the_planes_of_the_whole_grid(name):
    """fw25_mapgen_slab: an x-slab's planes (ghost planes included), built from host arrays that hold only the user-grid
    planes the slab reads, are byte-identical to the same planes of fw25_mapgen -- in the per-voxel and in the reference
    3D binary's dcmap mode, for first / interior / last slabs and a slab that lies wholly inside the boundary layer."""
    import dataclasses
    g = np.load(GOLD / f"mapgen_{name}.npz")
    for full3d in (True, False):
        spec = spec_of(name, g)
        spec.dcmap_full3d = full3d
        c = np.asarray(spec.sound_speed)
        spec.extra.update(c_min=float(c.min()), c_max=float(c.max()))
        nb, ex, ux = spec.num_boundary_points, spec.extended_shape[0], spec.user_shape[0]
        with mapgen.MapSet(spec) as whole:
            want = device_maps_of(whole)
        cuts = sorted({0, min(3, ex - 1), nb // 2 + 1, ex // 2, max(ex - nb + 2, 1), ex})
        for gx0, gx1 in zip(cuts[:-1], cuts[1:]):
            lo, hi = max(gx0 - 8, 0), min(gx1 + 8, ex)                      # owned planes + 8 ghost planes per side
            u0 = min(max(lo - nb, 0), ux - 1)
            u1 = min(max(hi - 1 - nb, 0), ux - 1) + 1

            def cut(a):
                return np.ascontiguousarray(np.asarray(a)[u0:u1])
            part = dataclasses.replace(
                spec, sound_speed=cut(spec.sound_speed), density=cut(spec.density), beta=cut(spec.beta),
                relax=None if spec.relax is None else {k: cut(v) for k, v in spec.relax.items()},
                alpha_coeff=None if spec.alpha_coeff is None else cut(spec.alpha_coeff),
                alpha_power=None if spec.alpha_power is None else cut(spec.alpha_power), user_planes=(u0, u1 - u0))
            with mapgen.MapSet(part, planes=(lo, hi)) as ms:
                assert ms.shape == (hi - lo,) + tuple(spec.extended_shape[1:])
                got = device_maps_of(ms)
                assert ms.invalid_count >= 0
            for stem in MAP_NAMES + ("dcmap",):
                assert np.array_equal(got[stem], want[stem][lo:hi]), (stem, lo, hi, full3d)
    if ux > 1:           # host arrays that lack planes the slab reads are refused
        one = lambda a: None if a is None else np.ascontiguousarray(np.asarray(a)[:1])   # noqa: E731
        short = dataclasses.replace(
            spec, sound_speed=one(spec.sound_speed), density=one(spec.density), beta=one(spec.beta),
            relax=None if spec.relax is None else {k: one(v) for k, v in spec.relax.items()},
            alpha_coeff=one(spec.alpha_coeff), alpha_power=one(spec.alpha_power), user_planes=(0, 1))
        with pytest.raises(engine.EngineError, match="do not hold the user-grid planes"):
            mapgen.MapSet(short, planes=(0, ex))


@pytest.mark.parametrize("name", ["m3d_big", "m2d_big", "m3d_lut"])
def test_background_generation_equals_the_blocking_call(name):
    """fw25_mapgen_slab_begin / fw25_mapgen_finish: the set's device pointers are final at once (an engine is created on
    them while the planes are still being uploaded, block by block, and generated behind the copies); after wait() the maps
    are byte-identical to the blocking builder's, for the whole grid and for a slab fed from a plane range."""
    import dataclasses
    g = np.load(GOLD / f"mapgen_{name}.npz")
    spec = spec_of(name, g)
    c = np.asarray(spec.sound_speed)
    spec.extra.update(c_min=float(c.min()), c_max=float(c.max()))
    with mapgen.MapSet(spec) as whole:
        want = device_maps_of(whole)
    with mapgen.MapSet(spec, background=True) as ms:
        ptrs = ms.device_maps()
        ms.wait()
        assert ms.device_maps()["rho"] == ptrs["rho"] and ms.shape == spec.extended_shape and ms.upload_ms >= 0
        got = device_maps_of(ms)
    for stem in MAP_NAMES + ("dcmap",):
        assert np.array_equal(got[stem], want[stem]), stem
    nb, ex, ux = spec.num_boundary_points, spec.extended_shape[0], spec.user_shape[0]
    lo, hi = ex // 3, ex - 2
    u0, u1 = min(max(lo - nb, 0), ux - 1), min(max(hi - 1 - nb, 0), ux - 1) + 1
    cut = lambda a: None if a is None else np.ascontiguousarray(np.asarray(a)[u0:u1])   # noqa: E731
    part = dataclasses.replace(spec, sound_speed=cut(spec.sound_speed), density=cut(spec.density), beta=cut(spec.beta),
                               relax=None if spec.relax is None else {k: cut(v) for k, v in spec.relax.items()},
                               alpha_coeff=cut(spec.alpha_coeff), alpha_power=cut(spec.alpha_power),
                               user_planes=(u0, u1 - u0))
    with mapgen.MapSet(part, planes=(lo, hi), background=True) as ms:
        got = device_maps_of(ms) if (ms.wait() or True) else None
    for stem in MAP_NAMES + ("dcmap",):
        assert np.array_equal(got[stem], want[stem][lo:hi]), stem


def test_engine_created_while_the_maps_are_still_arriving(monkeypatch):
    """An engine created on a background map set before wait() steps to the same frames as one created afterwards."""
    from tests.test_pipeline import sequential, spec_and_problem
    spec, pb = spec_and_problem((112, 24, 30), n_pml=5, n_trans=3, nT=20, modT=2, seed=44)
    want, _ = sequential(spec, pb)
    with mapgen.MapSet(spec, background=True) as ms:
        eng = engine.Engine(pb, device=0, device_maps=ms.device_maps())
        try:
            ms.wait()
            got, _ = eng.run()
        finally:
            eng.close()
    np.testing.assert_array_equal(got, want)
    assert np.abs(want).max() > 0


def test_lookup_counts_invalid_entries():
    g = np.load(GOLD / "mapgen_m2d_lut.npz")
    spec = spec_of("m2d_lut", g)
    m = mc.medium_arrays(mc.CASES["m2d_lut"])
    lut = mc.synthetic_lut(mc.CASES["m2d_lut"]["lut"])
    nb = spec.num_boundary_points
    idx = np.ix_(*[np.clip(np.arange(e) - nb, 0, n - 1) for e, n in zip(spec.extended_shape, spec.user_shape)])
    _, ia, ip = mo.lookup(m["alpha_coeff"][idx], m["alpha_power"][idx], lut["database"], lut["alpha_list"], lut["power_list"])
    with mapgen.MapSet(spec) as ms:
        assert ms.invalid_count == int(lut["invalid_matrix"][ia, ip].sum()) > 0


def test_errors_are_reported():
    spec = spec_of("m2d")
    spec.n_transition_layer = 0
    with pytest.raises(engine.EngineError, match="Transition layer is not defined"):
        mapgen.MapSet(spec)
    spec = spec_of("m2d")
    spec.relax = dict(spec.relax, kappa_x1=np.zeros((3, 3)))
    with pytest.raises(ValueError, match="map shape error"):
        mapgen.MapSet(spec)


@pytest.mark.parametrize("name,nT", [("m2d_big", 400), ("m3d_big", 120)])
def test_engine_on_device_maps_matches_engine_on_host_maps(name, nT):
    """Same sources / sensors; maps from fw25_mapgen (adopted in place) vs the oracle's maps uploaded from the host."""
    case, g, want = oracle_maps(name)
    spec = spec_of(name, g)
    ext = spec.extended_shape
    nd = len(ext)
    nb = spec.num_boundary_points
    rng = np.random.default_rng(5)
    # a 2-layer plane source just inside the user domain, sensors scattered over the whole extended grid
    mask = np.zeros(ext, bool)
    mask[nb:nb + 2] = True
    icc = np.stack(np.nonzero(mask), axis=1).astype(np.int32)
    t = np.arange(nT) * spec.dt
    pulse = (1e5 * np.sin(2 * np.pi * mc.F0 * t) * np.exp(-((t - 2.5e-6) / 1e-6) ** 2)).astype(np.float32)
    icmat = np.repeat(pulse[None, :], len(icc), axis=0)
    outc = np.stack([rng.integers(0, n, 300) for n in ext], axis=1).astype(np.int32)
    air = np.stack([rng.integers(nb, n - nb, 6) for n in ext], axis=1).astype(np.int32)
    d_tab, dmap, ndmap, _ = spec.stencil_tables()
    common = dict(ndim=nd, nX=ext[0], nY=ext[1], nZ=ext[2] if nd == 3 else 1, nT=nT, nTic=nT, modT=3, ndmap=ndmap,
                  dX=float(np.float32(spec.dx)), dT=float(np.float32(spec.dt)), dmap=dmap, icc=icc, icmat=icmat,
                  outc=outc, icczero=air, dcmap_full3d=True)
    host = Problem(**{k: want[k] for k in MAP_NAMES}, dcmap=want["dcmap"], **common)
    ref, _ = engine.run(host)
    with mapgen.MapSet(spec) as ms:
        dev = Problem(**{k: None for k in MAP_NAMES}, dcmap=None, **common)
        with engine.Engine(dev, device_maps=ms.device_maps()) as eng:
            got, stats = eng.run()
    assert stats["h2d_bytes"] < host.n_points * 14 * 4     # the maps were not uploaded
    assert np.isfinite(ref).all() and np.abs(ref).max() > 1.0
    err = np.linalg.norm(got.astype(np.float64) - ref) / np.linalg.norm(ref.astype(np.float64))
    assert err <= 1e-5, err
