"""Stand-in for the reference's relaxation-parameter database.

`fullwave.Medium.build()` maps (alpha_coeff, alpha_power) to the ten relaxation parameters of the two-mechanism model
through a precomputed table, `fullwave/utils/bins/database/relaxation_params_database_num_relax=2_20251027_1437.mat`
(/root/reference/fullwave/utils/relaxation_parameters.py:115-158).  That blob is missing from the reference checkout
(/root/reference/.MISSING_LARGE_BLOBS), so neither `fullwave.Medium` nor any shipped example can run without a
substitute.  This module generates one with the SAME SCHEMA (tests/utils/test_relaxation_parameters.py:
20-29 upstream):

    database       float64 [nA, nP, 10]   last axis: kappa_x1, kappa_x2, d_x1_nu1, alpha_x1_nu1, d_x2_nu1, alpha_x2_nu1,
                                          d_x1_nu2, alpha_x1_nu2, d_x2_nu2, alpha_x2_nu2   (solver/utils.py:68-86)
    alpha_0_list   float64 [1, nA]        ascending bin values of the attenuation coefficient [dB / MHz^y / cm]
    power_list     float64 [1, nP]        ascending bin values of the power-law exponent y
    invalid_matrix bool    [nA, nP]

The VALUES are a smooth closed form, not the authors' optimisation result: kappa ~ 1, two relaxation peaks bracketing
1 MHz whose strengths grow with alpha_0 and y (b = exp(-(d/kappa + alpha) dt) in (0.5, 1), small negative a), the same
recipe bench.py's synthetic medium uses.  Engine parity is unaffected -- both engines consume the maps this table
yields -- only the physical attenuation law differs from upstream's; every report that uses it says so.
"""

from __future__ import annotations

from pathlib import Path

import numpy as np

DB_NAME = "relaxation_params_database_num_relax=2_20251027_1437.mat"
F_REF = 1e6


def make() -> dict:
    alpha = np.round(np.arange(0.0, 2.5 + 1e-9, 0.025), 4)           # 101 bins
    power = np.round(np.arange(1.0, 2.0 + 1e-9, 0.025), 4)           # 41 bins
    a, y = np.meshgrid(alpha, power, indexing="ij")
    w1, w2 = 2 * np.pi * F_REF * 0.45, 2 * np.pi * F_REF * 2.6
    s = 0.012 * a * (1.0 + 0.25 * (y - 1.0))
    db = np.zeros(a.shape + (10,))
    db[..., 0] = 1.0 + 0.002 * (a - 0.5)          # kappa_x1
    db[..., 1] = 1.0 - 0.002 * (a - 0.5)          # kappa_x2
    db[..., 2], db[..., 3] = s * w1, w1           # d_x1_nu1, alpha_x1_nu1
    db[..., 4], db[..., 5] = 0.9 * s * w1, w1     # d_x2_nu1, alpha_x2_nu1
    db[..., 6], db[..., 7] = 0.6 * s * w2, w2     # d_x1_nu2, alpha_x1_nu2
    db[..., 8], db[..., 9] = 0.54 * s * w2, w2    # d_x2_nu2, alpha_x2_nu2
    invalid = (a > 2.4) & (y > 1.95)
    return {"database": db, "alpha_0_list": alpha[None, :], "power_list": power[None, :], "invalid_matrix": invalid}


def write_mat(path: str | Path) -> Path:
    from scipy.io import savemat
    path = Path(path)
    path.parent.mkdir(parents=True, exist_ok=True)
    savemat(path, make())
    return path


def install_into(package_root: str | Path) -> Path:
    """Write the stand-in where `fullwave.Medium` and the medium-builder domains look for the database
    (fullwave/solver/bins/database/, medium.py:776-780; package_root contains `fullwave/`)."""
    root = Path(package_root) / "fullwave"
    write_mat(root / "utils" / "bins" / "database" / DB_NAME)          # RelaxationParametersGenerator's own default
    return write_mat(root / "solver" / "bins" / "database" / DB_NAME)


def lookup_table():
    """The stand-in as a `mapgen.LookupTable` (for the GPU map builder)."""
    from .mapgen import LookupTable
    d = make()
    return LookupTable(d["database"], d["alpha_0_list"][0], d["power_list"][0], d["invalid_matrix"])
