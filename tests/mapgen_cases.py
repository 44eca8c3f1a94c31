"""Seeded media for the map-builder tests (fw25_mapgen): user-grid float64 maps + layer counts.

Shared by tools/make_mapgen_golden.py (which feeds them to the unmodified reference) and the CPU / GPU tests."""

from __future__ import annotations

import numpy as np

from fullwave25_b200 import synthetic

F0, C0, PPW, CFL = 1e6, 1540.0, 12, 0.2

CASES = {
    # small: every output byte is stored
    "m2d": dict(shape=(13, 17), n_pml=5, n_trans=4, seed=31, store_f64=True),
    "m3d": dict(shape=(4, 5, 6), n_pml=2, n_trans=2, seed=32),
    "m2d_nopml": dict(shape=(21, 16), n_pml=5, n_trans=4, seed=33, use_pml=False),
    "m2d_lut": dict(shape=(15, 14), n_pml=3, n_trans=6, seed=34, lut=77, store_f64=True),
    "m3d_onecell": dict(shape=(1, 3, 2), n_pml=1, n_trans=1, seed=35),
    # larger: sha256 of each .dat only
    "m3d_big": dict(shape=(20, 24, 28), n_pml=6, n_trans=7, seed=36, store=False),
    "m2d_big": dict(shape=(150, 131), n_pml=36, n_trans=36, seed=37, store=False),
    "m3d_lut": dict(shape=(12, 10, 9), n_pml=4, n_trans=3, seed=38, lut=78, store=False),
}


def make_grid(fw, case):
    """The reference Grid whose nx, ny[, nz] round back to the case's user shape."""
    dx = C0 / F0 / PPW
    dt = CFL * dx / C0
    domain = tuple((n + 0.01) * dx for n in case["shape"])
    return fw.Grid(domain, F0, (40 + 0.01) * dt, c0=C0, ppw=PPW, cfl=CFL)


def synthetic_lut(seed: int) -> dict:
    """Stand-in for the reference's relaxation-parameter database (missing from the checkout): same
    schema (database [nA, nP, 10], alpha_0_list, power_list, invalid_matrix), seeded values of plausible size."""
    rng = np.random.default_rng(seed)
    alpha_list = np.round(np.linspace(0.05, 2.0, 40), 4) + rng.uniform(0, 1e-12, 40)   # exercises .round(10)
    power_list = np.round(np.linspace(1.0, 2.0, 21), 4)
    nA, nP = len(alpha_list), len(power_list)
    db = np.zeros((nA, nP, 10))
    w1, w2 = 2 * np.pi * F0 * 0.45, 2 * np.pi * F0 * 2.6
    s = 0.012 * alpha_list[:, None] * (1 + 0.25 * (power_list[None, :] - 1))
    db[..., 0] = 1 + 0.01 * rng.standard_normal((nA, nP))
    db[..., 1] = 1 + 0.01 * rng.standard_normal((nA, nP))
    db[..., 2], db[..., 3] = s * w1, w1 * (1 + 0.05 * rng.standard_normal((nA, nP)))
    db[..., 4], db[..., 5] = 0.9 * s * w1, w1 * (1 + 0.05 * rng.standard_normal((nA, nP)))
    db[..., 6], db[..., 7] = 0.6 * s * w2, w2 * (1 + 0.05 * rng.standard_normal((nA, nP)))
    db[..., 8], db[..., 9] = 0.5 * s * w2, w2 * (1 + 0.05 * rng.standard_normal((nA, nP)))
    invalid = np.zeros((nA, nP), dtype=bool)
    invalid[-1, -1] = True
    return dict(database=db, alpha_list=alpha_list, power_list=power_list, invalid_matrix=invalid)


def medium_arrays(case) -> dict:
    """User-grid float64 maps of the case: sound_speed, density, beta and either the ten relaxation maps (`relax`) or
    alpha_coeff / alpha_power (look-up cases; values fall below, inside, on and above the table's bins)."""
    shape = tuple(case["shape"])
    rng = np.random.default_rng(case["seed"])
    T = synthetic.TISSUES
    lab = rng.integers(0, len(T), size=shape)
    out = dict(sound_speed=T[lab, 0] + rng.uniform(-0.6, 0.6, size=shape), density=T[lab, 1] + rng.uniform(-1, 1, size=shape),
               beta=T[lab, 2] + rng.uniform(-0.1, 0.1, size=shape))
    out["sound_speed"].flat[0] = 1500.5          # a tie for round(c + 1e-9)
    if case.get("lut"):
        out["alpha_coeff"] = rng.uniform(0.0, 2.2, size=shape)
        out["alpha_power"] = rng.uniform(0.9, 2.1, size=shape)
        lut = synthetic_lut(case["lut"])
        out["alpha_coeff"].flat[1] = lut["alpha_list"][7].round(10)     # exactly on a bin edge
        out["alpha_power"].flat[1] = lut["power_list"][3]
        if len(shape) == 2:
            out["alpha_coeff"].flat[2] = 5.0                           # clipped into the (invalid) last bin
            out["alpha_power"].flat[2] = 5.0
        else:
            # a 3D medium that hits an invalid entry crashes the reference while it formats the warning
            # (relaxation_parameters.py:58-61 indexes axis 2 of a 4-D tensor): keep 3D cases clear of that bin
            both = (out["alpha_coeff"] > 1.9) & (out["alpha_power"] > 1.9)
            out["alpha_power"][both] = 1.5
    else:
        tab = synthetic.relaxation_table(F0)
        jit = lambda: 1 + 0.03 * rng.standard_normal(shape)  # noqa: E731
        out["relax"] = {
            "kappa_x1": tab[0, lab, 0] * jit(), "kappa_x2": tab[1, lab, 0] * jit(),
            "d_x1_nu1": tab[0, lab, 1] * jit(), "alpha_x1_nu1": tab[0, lab, 2] * jit(),
            "d_x1_nu2": tab[0, lab, 3] * jit(), "alpha_x1_nu2": tab[0, lab, 4] * jit(),
            "d_x2_nu1": tab[1, lab, 1] * jit(), "alpha_x2_nu1": tab[1, lab, 2] * jit(),
            "d_x2_nu2": tab[1, lab, 3] * jit(), "alpha_x2_nu2": tab[1, lab, 4] * jit(),
        }
    return out
