"""GPU: fw25_run_medium -- the medium uploaded block by block, maps generated as the blocks land, the first time steps
swept block-wise (time-skewed) underneath -- must give the SAME BITS as the sequential path fw25_mapgen -> fw25_create
-> fw25_run_engine, for every split between skewed and whole-grid steps, with sources, air voxels and sensors on
block boundaries, in the never-updated rim and in the block a later step is already working on."""

import numpy as np
import pytest

from fullwave25_b200 import engine, mapgen, synthetic
from fullwave25_b200.problem import MAP_NAMES, Problem
from tests import mapgen_cases as mc

pytestmark = pytest.mark.gpu
M = 8


def spec_and_problem(user_shape, *, n_pml, n_trans, nT, modT, seed, lut=False, f32=False, n_sensors=300, n_air=40):
    case = dict(shape=user_shape, n_pml=n_pml, n_trans=n_trans, seed=seed, lut=77 if lut else None)
    m = mc.medium_arrays(case)
    dx = mc.C0 / mc.F0 / mc.PPW
    dt = mc.CFL * dx / mc.C0
    kw = {}
    if lut:
        t = mc.synthetic_lut(77)
        kw = dict(alpha_coeff=m["alpha_coeff"], alpha_power=m["alpha_power"],
                  lut=mapgen.LookupTable(t["database"], t["alpha_list"], t["power_list"], t["invalid_matrix"]))
    else:
        kw = dict(relax=m["relax"])
    cast = (lambda a: np.asarray(a, np.float32)) if f32 else (lambda a: a)
    if f32:
        kw = {k: ({kk: cast(vv) for kk, vv in v.items()} if k == "relax" else (cast(v) if isinstance(v, np.ndarray) else v))
              for k, v in kw.items()}
    spec = mapgen.MediumSpec(user_shape=tuple(user_shape), dt=dt, dx=dx, c0=mc.C0, cfl=mc.CFL,
                             sound_speed=cast(m["sound_speed"]), density=cast(m["density"]), beta=cast(m["beta"]),
                             n_pml_layer=n_pml, n_transition_layer=n_trans, dcmap_full3d=True, **kw)
    ext = spec.extended_shape
    nb = spec.num_boundary_points
    rng = np.random.default_rng(seed + 1)

    def pts(n, lo=nb, planes=None):
        x = rng.integers(lo, ext[0] - lo, size=n) if planes is None else rng.choice(planes, size=n)
        return np.stack([x, rng.integers(lo, ext[1] - lo, size=n), rng.integers(lo, ext[2] - lo, size=n)], axis=1).astype(np.int32)

    edges = [b for b in range(16, ext[0], 16)]        # fd_u works on [32 m, 32 m + 32), fd_p on ranges shifted by 16
    near = sorted({x for e in edges for x in (e - 9, e - 8, e - 1, e, e + 7, e + 8) if M <= x < ext[0] - M})
    # a plane source near the low-x face, point sources on block edges, a few in the rim; sensors and air everywhere
    ys, zs = np.meshgrid(np.arange(nb, ext[1] - nb), np.arange(nb, ext[2] - nb), indexing="ij")
    plane = np.stack([np.full(ys.size, nb), ys.ravel(), zs.ravel()], axis=1).astype(np.int32)
    # (no point source on the plane source's own plane: two sources on one cell are a write race in any engine)
    icc = np.concatenate([plane, pts(12, planes=[x for x in near if x != nb]), pts(3, lo=0, planes=[1, 5, ext[0] - 3])])
    nTic = min(nT, 30)
    pulse = synthetic.tone_burst(nTic, dt, mc.F0).astype(np.float32)
    icmat = np.concatenate([np.repeat(pulse[None], len(plane), 0), 0.3 * np.repeat(pulse[None], len(icc) - len(plane), 0)])
    outc = np.concatenate([pts(n_sensors), pts(60, planes=near), pts(10, lo=0, planes=[0, 3, ext[0] - 1]), icc[-6:]])
    icczero = np.concatenate([pts(n_air), pts(10, planes=near), icc[len(plane): len(plane) + 2]])   # incl. air ON a source
    none = {name: None for name in MAP_NAMES}
    d_table, dmap, ndmap, _ = spec.stencil_tables()
    pb = Problem(ndim=3, nX=ext[0], nY=ext[1], nZ=ext[2], nT=nT, nTic=nTic, modT=modT, ndmap=ndmap,
                 dX=float(np.float32(dx)), dT=float(np.float32(dt)), **none, dmap=dmap, dcmap=None, icc=icc, icmat=icmat,
                 outc=outc, icczero=icczero, extra={}, dcmap_full3d=True)
    return spec, pb.normalise()


def sequential(spec, pb):
    with mapgen.MapSet(spec) as ms:
        eng = engine.Engine(pb, device=0, device_maps=ms.device_maps())
        try:
            return eng.run()
        finally:
            eng.close()


@pytest.mark.parametrize("skew", ["0", "1", "2", "5", "auto", "all"])
def test_pipelined_run_is_bit_identical_to_sequential(monkeypatch, skew):
    monkeypatch.setenv("FW25_GRAPH", "0")           # small grid: keep it on the launched-steps path the big grids use
    nT = 26
    spec, pb = spec_and_problem((112, 24, 30), n_pml=5, n_trans=3, nT=nT, modT=3, seed=41)   # extended 144 x 56 x 62
    want, _ = sequential(spec, pb)
    assert np.abs(want).max() > 0 and (want[:, -6:] != 0).any()
    if skew not in ("auto",):
        monkeypatch.setenv("FW25_SKEW_STEPS", str(nT) if skew == "all" else skew)
    got, stats = mapgen.run_medium(spec, pb)
    np.testing.assert_array_equal(got, want)
    expect = {"auto": 2, "all": nT}.get(skew, int(skew) if skew.isdigit() else None)      # auto: max(2, 5 blocks // 3)
    assert stats["skewed_steps"] == expect
    assert stats["point_updates"] == pb.n_points * nT


@pytest.mark.parametrize("lut,f32", [(True, False), (False, True), (True, True)])
def test_pipelined_run_lookup_and_float32_inputs(monkeypatch, lut, f32):
    """Look-up media and float32 user maps through the streamed generator: the same bits as the one-shot generator fed
    the same arrays (float32 inputs are widened exactly on the device)."""
    monkeypatch.setenv("FW25_GRAPH", "0")
    spec, pb = spec_and_problem((70, 20, 22), n_pml=4, n_trans=4, nT=18, modT=2, seed=43, lut=lut, f32=f32)   # 102 x 52 x 54
    want, _ = sequential(spec, pb)
    got, stats = mapgen.run_medium(spec, pb)
    np.testing.assert_array_equal(got, want)
    assert stats["skewed_steps"] == 2 and np.abs(want).max() > 0


def test_streamed_maps_equal_one_shot_maps(monkeypatch):
    """A run with zero steps leaves only the generator: frames empty, and a sequential engine over the same medium
    steps identically afterwards (covered above); here the ragged last block and graph-replayed small grids."""
    spec, pb = spec_and_problem((37, 18, 19), n_pml=3, n_trans=2, nT=12, modT=1, seed=47)   # 63 planes: blocks 32 + 31
    want, _ = sequential(spec, pb)
    got, stats = mapgen.run_medium(spec, pb)          # graph replay on: whole-grid steps after the last block
    np.testing.assert_array_equal(got, want)
    assert stats["skewed_steps"] == 0
    monkeypatch.setenv("FW25_GRAPH", "0")
    got2, stats2 = mapgen.run_medium(spec, pb)
    np.testing.assert_array_equal(got2, want)
    assert stats2["skewed_steps"] == 2


@pytest.mark.parametrize("n,fused", [(2, "1"), (2, "0"), (3, "1")])
def test_run_medium_on_several_devices(monkeypatch, n, fused):
    """fw25_run_medium_multi: every device builds its own x-slab of the maps (fw25_mapgen_slab) from the user-grid
    medium and the native multi-device runner steps them -- the same bits as one device (slabs share device 0 when the
    box has a single GPU), with sources / sensors / air voxels near the interfaces and in the rim."""
    from tests.test_multi_gpu import _devices
    monkeypatch.setenv("FW25_FUSED_HALO", fused)
    spec, pb = spec_and_problem((112, 24, 30), n_pml=5, n_trans=3, nT=26, modT=3, seed=41)   # extended 144 x 56 x 62
    want, _ = sequential(spec, pb)
    got, stats = mapgen.run_medium(spec, pb, device_ids=_devices(n))
    assert stats["n_devices"] == n and stats["halo_bytes"] > 0 and np.abs(want).max() > 0
    np.testing.assert_array_equal(got, want)
    spec.dcmap_full3d = False                       # the reference 3D binary's dcmap truncation follows the GLOBAL index
    want2, _ = sequential(spec, pb)
    got2, _ = mapgen.run_medium(spec, pb, device_ids=_devices(n))
    np.testing.assert_array_equal(got2, want2)
    assert not np.array_equal(want2, want)


def test_run_medium_reports_errors():
    spec, pb = spec_and_problem((37, 18, 19), n_pml=3, n_trans=2, nT=4, modT=1, seed=47)
    pb.icc = pb.icc.copy()
    pb.icc[0, 0] = 10_000
    with pytest.raises(engine.EngineError, match="outside the grid"):
        mapgen.run_medium(spec, pb)
