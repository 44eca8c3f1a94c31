"""Synthetic heterogeneous attenuating media in the engine's input format (BASELINE.json configs[4],
SURVEY.md 8(d) "Concrete inputs").

The reference builds its coefficient maps with `fullwave.Medium` (needs a lookup database that is
missing from the checkout, /root/reference/.MISSING_LARGE_BLOBS) + `PMLBuilder`
(/root/reference/fullwave/solver/pml_builder.py:794-1254).  This module produces maps of the same
*form* -- per-voxel b = exp(-(d/kappa + alpha) dT), a = d / (kappa (d + kappa alpha) + 1e-10) (b - 1)
(pml_builder.py:794-810), two mechanisms per family, nu = 1 ramped to a CPML damping profile and nu = 2
ramped to zero inside the boundary layer -- from a small tissue table, so that tests and benchmarks can
run without the reference package.  It is a workload generator, not a re-implementation of the
reference's medium builder (which stays in Python upstream and is out of scope).
"""

from __future__ import annotations

import numpy as np

from . import stencil
from .problem import Problem

M = 8

# tissue table: c [m/s], rho [kg/m3], beta (= 1 + B/2A), alpha0 [dB/MHz^y/cm], y
# (ranges of /root/reference/fullwave/constants/material_properties.py and SURVEY.md 8(d))
TISSUES = np.array([
    # c       rho     beta  alpha0  y
    [1540.0, 1000.0, 3.50, 0.50, 1.10],   # background
    [1478.0, 937.0, 5.00, 0.40, 1.00],    # fat
    [1547.0, 1050.0, 3.87, 0.15, 1.20],   # muscle
    [1613.0, 1120.0, 4.00, 0.75, 1.30],   # connective
    [1567.0, 1040.0, 3.90, 0.30, 1.05],   # liver
    [1412.0, 960.0, 5.50, 0.60, 1.15],    # oil-like
])


def relaxation_table(f0: float) -> np.ndarray:
    """Per-tissue (kappa, d1, alpha1, d2, alpha2) for the two coefficient families [family, tissue, 5].
    Two relaxation peaks bracketing f0; strengths scale with alpha0 (stand-in for the reference's LUT)."""
    n = len(TISSUES)
    out = np.zeros((2, n, 5))
    w1, w2 = 2 * np.pi * f0 * 0.45, 2 * np.pi * f0 * 2.6
    for fam in range(2):
        for i, (_c, _r, _b, a0, y) in enumerate(TISSUES):
            strength = 0.012 * a0 * (1.0 + 0.25 * (y - 1.0)) * (1.0 if fam == 0 else 0.9)
            out[fam, i] = (1.0 + 0.004 * (i - 2) * (1 if fam == 0 else -1),
                           strength * w1, w1, 0.6 * strength * w2, w2)
    return out


def calc_a_b(d, kappa, alpha, dt):
    """a, b of one mechanism (closed form of pml_builder.py:794-810 / medium.py:273-291), float64."""
    b = np.exp(-(d / kappa + alpha) * dt)
    a = d / (kappa * (d + kappa * alpha) + 1e-10) * (b - 1)
    return a, b


def boundary_depth(n: int, n_pml: int, n_trans: int) -> tuple[np.ndarray, np.ndarray]:
    """1-D profiles along one axis of an extended grid of n cells: (pml depth fraction in [0,1],
    transition blend in [0,1]); both 0 in the user domain, 1 in the outer M ghost cells."""
    xi = np.zeros(n)
    tr = np.zeros(n)
    for i in range(n):
        dist = min(i, n - 1 - i)               # cells from the nearest face
        if dist < M:
            xi[i] = 1.0
            tr[i] = 1.0
        elif dist < M + n_pml:
            xi[i] = (M + n_pml - dist) / max(n_pml, 1)
            tr[i] = 1.0
        elif dist < M + n_pml + n_trans:
            tr[i] = 0.5 * (1 - np.cos(np.pi * (M + n_pml + n_trans - dist) / max(n_trans, 1)))
    return xi, tr


def tone_burst(nt: int, dt: float, f0: float, n_cycles: float = 2.0, amp: float = 1e5) -> np.ndarray:
    t = np.arange(nt) * dt
    dur = n_cycles / f0
    env = np.where(t < dur, np.sin(np.pi * t / dur) ** 2, 0.0)
    return amp * env * np.sin(2 * np.pi * f0 * t)


def make_problem(shape, *, nT: int, f0: float = 1e6, c0: float = 1540.0, ppw: int = 12, cfl: float = 0.2,
                 n_pml: int = 6, n_trans: int = 4, block: int = 5, seed: int = 1234, modT: int = 1,
                 n_sensors: int = 64, n_air: int = 8, homogeneous: bool = False,
                 source_layers: int = 3, amp: float = 1e5, aniso: bool = False) -> Problem:
    """shape: EXTENDED grid (nX, nY[, nZ]) including the boundary layer of M + n_pml + n_trans cells."""
    shape = tuple(int(s) for s in shape)
    ndim = len(shape)
    rng = np.random.default_rng(seed)
    dx = c0 / f0 / ppw
    dt = cfl * dx / c0
    nb = M + n_pml + n_trans
    assert all(s > 2 * nb + 2 for s in shape), "grid too small for the boundary layer"

    # piecewise-constant tissue labels (random boxes); boundary layer replicates the edge tissue
    coarse = tuple(-(-s // block) for s in shape)
    lab = rng.integers(0, 1 if homogeneous else len(TISSUES), size=coarse)
    for ax in range(ndim):
        lab = np.repeat(lab, block, axis=ax)
    lab = lab[tuple(slice(0, s) for s in shape)]
    inner = tuple(slice(nb, s - nb) for s in shape)
    lab = np.pad(lab[inner], nb, mode="edge")

    c = TISSUES[lab, 0] + (0.0 if homogeneous else rng.uniform(-0.4, 0.4, size=shape))
    rho = TISSUES[lab, 1]
    beta = TISSUES[lab, 2]
    K = c**2 * rho

    prof = [boundary_depth(s, n_pml, n_trans) for s in shape]
    xi = np.zeros(shape)
    tr = np.zeros(shape)
    for ax in range(ndim):
        sh = [1] * ndim
        sh[ax] = shape[ax]
        xi = np.maximum(xi, prof[ax][0].reshape(sh))
        tr = np.maximum(tr, prof[ax][1].reshape(sh))

    L = (n_pml + n_trans) * dx
    d_pml = -(2 + 1) * c0 * np.log(1e-30) / (2 * L) if n_pml > 0 else 0.0
    table = relaxation_table(f0)
    maps = {}
    for fam, tag in ((0, "x"), (1, "u")):
        kappa = 1.0 + (table[fam, lab, 0] - 1.0) * (1 - tr)
        d1 = table[fam, lab, 1] * (1 - tr) + d_pml * xi**2
        al1 = table[fam, lab, 2] * (1 - tr)
        d2 = table[fam, lab, 3] * (1 - tr)
        al2 = table[fam, lab, 4] * (1 - tr)
        a1, b1 = calc_a_b(d1, kappa, al1, dt)
        a2, b2 = calc_a_b(d2, kappa, al2, dt)
        maps[f"kappa{tag}"] = kappa
        maps[f"apml{tag}1"], maps[f"bpml{tag}1"] = a1, b1
        maps[f"apml{tag}2"], maps[f"bpml{tag}2"] = a2, b2

    aniso_maps = None
    if aniso:
        # anisotropic protocol: one kappa / a / b set PER AXIS.  Each axis gets its own PML ramp (along that axis
        # only, a split-field layer) and slightly different relaxation strengths, so that every array differs from
        # every other one everywhere -- a permuted axis assignment cannot go unnoticed.
        aniso_maps = {}
        letters = {0: ("x", "y", "z")[:ndim], 1: ("u", "w") if ndim == 2 else ("u", "v", "w")}
        for fam in (0, 1):
            for ax, letter in enumerate(letters[fam]):
                sh = [1] * ndim
                sh[ax] = shape[ax]
                xi_a = np.broadcast_to(prof[ax][0].reshape(sh), shape)
                tr_a = np.broadcast_to(prof[ax][1].reshape(sh), shape)
                kappa = (1.0 + (table[fam, lab, 0] - 1.0) * (1 - tr_a)) * (1.0 + 0.003 * ax)
                d1 = table[fam, lab, 1] * (1 - tr_a) * (1.0 + 0.15 * ax) + d_pml * xi_a**2
                al1 = table[fam, lab, 2] * (1 - tr_a)
                d2 = table[fam, lab, 3] * (1 - tr_a) * (1.0 - 0.1 * ax)
                al2 = table[fam, lab, 4] * (1 - tr_a)
                a1, b1 = calc_a_b(d1, kappa, al1, dt)
                a2, b2 = calc_a_b(d2, kappa, al2, dt)
                aniso_maps[f"kappa{letter}"] = kappa
                aniso_maps[f"apml{letter}1"], aniso_maps[f"bpml{letter}1"] = a1, b1
                aniso_maps[f"apml{letter}2"], aniso_maps[f"bpml{letter}2"] = a2, b2
        for k in list(maps):                       # the isotropic slots hold the x-axis members
            maps[k] = aniso_maps[k]

    d_tab, dmap, dcmap, ndmap = stencil.tables(c, dt=dt, dx=dx, cfl=cfl, is_3d=ndim == 3)

    # plane source: `source_layers` planes at x = nb .. nb+layers-1 over the user cross-section
    smask = np.zeros(shape, bool)
    smask[(slice(nb, nb + source_layers),) + inner[1:]] = True
    icc = np.stack(np.nonzero(smask), axis=1)
    nTic = min(nT, int(np.ceil(2.0 / f0 / dt)) + 1)
    pulse = tone_burst(nTic, dt, f0, amp=amp)
    icmat = np.zeros((icc.shape[0], nTic))
    for layer in range(source_layers):  # delay each layer by one cell's travel time (as examples/wave_3d does)
        sel = icc[:, 0] == nb + layer
        shift = int(round(layer * dx / c0 / dt))
        if shift < nTic:
            icmat[sel, shift:] = pulse[: nTic - shift]

    # point sensors + air voxels scattered in the user domain (row-major order like np.where)
    user = np.zeros(shape, bool)
    user[inner] = True
    user[smask] = False
    cand = np.flatnonzero(user)
    pick = np.sort(rng.choice(cand, size=min(n_sensors, cand.size), replace=False))
    outc = np.stack(np.unravel_index(pick, shape), axis=1)
    air_pick = np.sort(rng.choice(cand, size=min(n_air, cand.size), replace=False)) if n_air else np.zeros(0, int)
    icczero = np.stack(np.unravel_index(air_pick, shape), axis=1) if n_air else np.zeros((0, ndim), int)

    pb = Problem(
        ndim=ndim, nX=shape[0], nY=shape[1], nZ=shape[2] if ndim == 3 else 1, nT=nT, nTic=nTic, modT=modT,
        ndmap=ndmap, dX=float(np.float32(dx)), dT=float(np.float32(dt)), rho=rho, K=K, beta=beta, **maps,
        dmap=dmap, dcmap=dcmap, icc=icc, icmat=icmat, outc=outc, icczero=icczero,
        extra={"c": c, "d": d_tab, "c0": c0, "dY": dx, "dZ": dx}, aniso=aniso_maps,
    )
    return pb.normalise()
