"""CPU: the map-builder oracle (oracle/mapgen_oracle.py) against golden vectors made by the reference itself
(tools/make_mapgen_golden.py: PMLBuilder + InputFileWriter of the unmodified package), and the host-side pieces of
fullwave25_b200/mapgen.py (tables, spec) against the same goldens.

Tolerances: rho, K, beta, kappa*, dcmap and the float64 d / alpha maps after the PML ramps are BIT-EXACT.  The a / b maps
go through exp(): numpy's float64 exp differs in the last bit between CPU families, which moves a float32 result by
at most one ulp -- tolerance 1 float32 ulp, stated here and checked per element."""

import hashlib
from pathlib import Path

import numpy as np
import pytest

from fullwave25_b200 import mapgen
from oracle import mapgen_oracle as mo
from tests import mapgen_cases as mc

GOLD = Path(__file__).resolve().parent / "golden"
EXACT = ("rho", "K", "beta", "kappax", "kappau", "dcmap")
AB = tuple(f"{ab}pml{l}{nu}" for l in "xu" for nu in (1, 2) for ab in "ab")
AB_ULP = 1


def oracle_maps(name):
    case = mc.CASES[name]
    g = np.load(GOLD / f"mapgen_{name}.npz")
    m = mc.medium_arrays(case)
    if case.get("lut"):
        lut = mc.synthetic_lut(case["lut"])
        relax, _, _ = mo.lookup(m["alpha_coeff"], m["alpha_power"], lut["database"], lut["alpha_list"], lut["power_list"])
    else:
        relax = m["relax"]
    use_pml = case.get("use_pml", True)
    out = mo.build_maps(sound_speed=m["sound_speed"], density=m["density"], beta=m["beta"], relax=relax,
                        dt=float(g["dt"]), dx=float(g["dx"]), c0=mc.C0, n_pml=case["n_pml"] if use_pml else 0,
                        n_trans=case["n_trans"] if use_pml else 0, use_pml=use_pml)
    return case, g, out


def check_against_golden(g, got, *, ab_ulp=AB_ULP):
    """got: {stem: array}.  Full maps where the golden stores them, sha256 + sub-sampled a / b otherwise."""
    assert tuple(got["rho"].shape) == tuple(g["ext_shape"])
    sha = dict(zip(g["sha_keys"].tolist(), g["sha_vals"].tolist()))
    worst = 0
    for stem in EXACT:
        assert hashlib.sha256(np.ascontiguousarray(got[stem]).tobytes()).hexdigest() == sha[stem], stem
        if stem in g:
            assert np.array_equal(got[stem], g[stem]), stem
    for stem in AB:
        if stem in g:
            ref = g[stem]
            mine = got[stem]
        else:
            ref = g[f"sub_{stem}"]
            mine = got[stem][(slice(None, None, 4),) * ref.ndim]
        d = mo.ulp_distance_f32(mine, ref)
        assert d.max() <= ab_ulp, f"{stem}: {d.max()} ulp"
        worst = max(worst, int(d.max()))
    return worst


@pytest.mark.parametrize("name", list(mc.CASES))
def test_oracle_matches_reference_goldens(name):
    _, g, out = oracle_maps(name)
    check_against_golden(g, out)


@pytest.mark.parametrize("name", ["m2d", "m2d_lut"])
def test_oracle_ramped_d_alpha_are_bit_exact(name):
    _, g, out = oracle_maps(name)
    for letter in "xu":
        for nu in (1, 2):
            for q in ("d", "alpha"):
                k = f"{q}_{letter}_nu{nu}"
                assert np.array_equal(out[k], g[k]), k


def test_lookup_indices_edges_and_clipping():
    lut = mc.synthetic_lut(77)
    al, pl = lut["alpha_list"], lut["power_list"]
    a = np.array([[al[7].round(10), -1.0, 9.0, al[7].round(10) + 1e-12]])
    p = np.array([[pl[3], 0.0, 9.0, pl[3] - 1e-12]])
    _, ia, ip = mo.lookup(a, p, lut["database"], al, pl)
    # below-range values are clipped to the RAW list minimum, then searched in the list ROUNDED to 10 decimals
    # (relaxation_parameters.py:50-51, :152-155, :216): 0.05 + 5e-13 > 0.05, so they land in bin 1, not 0
    assert al.min() > al.round(10).min()
    assert ia.tolist() == [[7, 1, len(al) - 1, 8]]
    assert ip.tolist() == [[3, 0, len(pl) - 1, 3]]


def spec_of(name, g=None):
    case = mc.CASES[name]
    g = np.load(GOLD / f"mapgen_{name}.npz") if g is None else g
    m = mc.medium_arrays(case)
    use_pml = case.get("use_pml", True)
    kw = dict(user_shape=case["shape"], dt=float(g["dt"]), dx=float(g["dx"]), c0=mc.C0, cfl=mc.CFL,
              sound_speed=m["sound_speed"], density=m["density"], beta=m["beta"],
              n_pml_layer=case["n_pml"] if use_pml else 0, n_transition_layer=case["n_trans"] if use_pml else 0,
              use_pml=use_pml, dcmap_full3d=True)
    if case.get("lut"):
        lut = mc.synthetic_lut(case["lut"])
        return mapgen.MediumSpec(alpha_coeff=m["alpha_coeff"], alpha_power=m["alpha_power"],
                                 lut=mapgen.LookupTable(lut["database"], lut["alpha_list"], lut["power_list"],
                                                        lut["invalid_matrix"]), **kw)
    return mapgen.MediumSpec(relax=m["relax"], **kw)


@pytest.mark.parametrize("name", list(mc.CASES))
def test_host_spec_tables_match_reference(name):
    """MediumSpec's host-side scalars / tables: extended shape, ndmap and the dmap stencil table as the reference
    wrote them (dmap.dat, ndmap.dat)."""
    g = np.load(GOLD / f"mapgen_{name}.npz")
    spec = spec_of(name, g)
    assert spec.extended_shape == tuple(g["ext_shape"])
    _, dmap, ndmap, _ = spec.stencil_tables()
    assert ndmap == int(g["ndmap"])
    assert np.array_equal(dmap.reshape(-1), g["dmap"])


def test_spec_from_reference_like_pml_builder():
    """`MediumSpec.from_pml_builder` reads the attributes `Solver.__init__` leaves on its PMLBuilder."""
    from types import SimpleNamespace as NS
    case = mc.CASES["m2d"]
    m = mc.medium_arrays(case)
    med = NS(sound_speed=m["sound_speed"], density=m["density"], beta=m["beta"], relaxation_param_dict=m["relax"])
    pmlb = NS(medium_org=med, extended_grid=NS(dt=1e-8, dx=1e-4, c0=1540.0, cfl=0.2), m_spatial_order=8,
              n_pml_layer=5, n_transition_layer=4, n_polynomial=2, theoritical_reflection_coefficient=1e-30)
    spec = mapgen.MediumSpec.from_pml_builder(pmlb)
    assert spec.extended_shape == (47, 51) and spec.relax["d_x2_nu2"] is m["relax"]["d_x2_nu2"]
    want = -(2 + 1) * 1540.0 * np.log(1e-30) / (2 * (1e-4 * 5 + 1e-4 * 4))
    assert spec.d_target_pml() == want


def test_spec_from_a_medium_with_lookup_database(tmp_path):
    """A `fullwave.Medium` (alpha_coeff / alpha_power + path to the .mat database, medium.py:766-840): the spec reads
    the database file the medium names, with the reference's schema (relaxation_parameters.py:147-155)."""
    from types import SimpleNamespace as NS
    from scipy.io import savemat
    case = mc.CASES["m2d_lut"]
    m = mc.medium_arrays(case)
    lut = mc.synthetic_lut(case["lut"])
    savemat(tmp_path / "db.mat", {"database": lut["database"], "alpha_0_list": lut["alpha_list"][None, :],
                                  "power_list": lut["power_list"][None, :], "invalid_matrix": lut["invalid_matrix"]})
    med = NS(sound_speed=m["sound_speed"], density=m["density"], beta=m["beta"], alpha_coeff=m["alpha_coeff"],
             alpha_power=m["alpha_power"], path_relaxation_parameters_database=tmp_path / "db.mat",
             relaxation_param_dict={})        # (a built Medium also carries this attribute: alpha_coeff decides)
    pmlb = NS(medium_org=med, extended_grid=NS(dt=1e-8, dx=1e-4, c0=1540.0, cfl=0.2), m_spatial_order=8,
              n_pml_layer=3, n_transition_layer=6)
    spec = mapgen.MediumSpec.from_pml_builder(pmlb)
    assert spec.relax is None and spec.alpha_power is m["alpha_power"]
    np.testing.assert_array_equal(spec.lut.database, lut["database"])
    np.testing.assert_array_equal(spec.lut.alpha_list, lut["alpha_list"])
    np.testing.assert_array_equal(spec.lut.power_list, lut["power_list"])
    assert spec.lut.invalid_matrix.shape == lut["invalid_matrix"].shape and spec.lut.invalid_matrix[-1, -1]


def test_plane_range_specs_marshal_with_the_whole_grid_extent():
    """x-sharded runs: a MediumSpec that holds only a range of user-grid x planes (fw25_mapgen_slab) keeps the WHOLE
    grid's extent in fw25_medium.nx, checks its arrays against the plane count, and needs the global sound-speed range
    (the stencil table must be the same on every rank)."""
    import dataclasses
    spec = spec_of("m3d_big")
    cut = lambda a: np.ascontiguousarray(np.asarray(a)[5:12])   # noqa: E731
    part = dataclasses.replace(spec, sound_speed=cut(spec.sound_speed), density=cut(spec.density), beta=cut(spec.beta),
                               relax={k: cut(v) for k, v in spec.relax.items()}, user_planes=(5, 7))
    with pytest.raises(ValueError, match="c_min"):
        mapgen.marshal_medium(part)
    c = np.asarray(spec.sound_speed)
    part.extra.update(c_min=float(c.min()), c_max=float(c.max()))
    md, keep, tables = mapgen.marshal_medium(part)
    whole_md, _, whole_tables = mapgen.marshal_medium(spec)
    assert (md.nx, md.ny, md.nz) == (whole_md.nx, whole_md.ny, whole_md.nz) == tuple(spec.user_shape)
    assert md.c_round_min == whole_md.c_round_min and tables[2] == whole_tables[2]
    np.testing.assert_array_equal(tables[1], whole_tables[1])
    with pytest.raises(ValueError, match="map shape error"):
        mapgen.marshal_medium(dataclasses.replace(part, user_planes=(5, 8)))


def test_user_medium_generator_plane_ranges_equal_the_whole_medium():
    """bench.py's per-rank user-grid medium (synthetic_device.make_user_medium(x_range=...)) is the same medium as the
    whole one: labels hash GLOBAL block coordinates."""
    from fullwave25_b200 import synthetic_device
    whole, c0, c1, _ = synthetic_device.make_user_medium((40, 12, 10), device="cpu", block=6, seed=9, pin=False)
    part, p0, p1, _ = synthetic_device.make_user_medium((40, 12, 10), device="cpu", block=6, seed=9, pin=False,
                                                        x_range=(13, 31), chunk=7)
    for k in whole:
        np.testing.assert_array_equal(part[k], whole[k][13:31])
    assert c0 <= p0 <= p1 <= c1
