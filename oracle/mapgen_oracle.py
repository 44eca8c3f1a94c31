"""CPU restatement (numpy) of the reference's host-side map builder -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; nothing under
fullwave25_b200/ does.  It checks the CUDA kernel `k_mapgen` (fullwave25_b200/csrc/fw25_mapgen.cu).

What it restates (reference file:line):
  * pad by edge replication                      /root/reference/fullwave/solver/pml_builder.py:321-704
  * per-axis target / ramp passes                pml_builder.py:1256-1498 (`_apply_transition_and_pml`), driven by
                                                 `_apply_pml` :842-1025 and `_apply_pml_3d` :1027-1254
  * a, b                                         pml_builder.py:794-810 (== medium.py:273-291)
  * key renaming (x <- x2, u <- x1)              pml_builder.py:42-91, medium.py:340-357
  * K = c^2 rho                                  medium.py:256-259
  * dcmap and the float32 casts                  input_file_writer.py:95-103, :558-559, :870-881
  * relaxation-parameter look-up                 fullwave/utils/relaxation_parameters.py:18-75, :189-242

Pinned against the reference itself: tools/make_mapgen_golden.py imports the unmodified reference package, runs
`PMLBuilder(...).run()` / `Medium.build()` on seeded media and commits its float32 maps and float64 post-ramp d / alpha
as tests/golden/mapgen_*.npz; tests/test_mapgen_oracle.py compares (bit-exact for everything but a / b, which may
differ by one float32 ulp when numpy's exp() differs between CPUs).

Unlike the reference (sequential in-place passes over whole extended arrays), this is the per-voxel closed form the
kernel uses: value at the clamped user coordinate, then for axis 0, 1[, 2]: target / ramp / keep.
"""

from __future__ import annotations

import numpy as np

RELAX_KEYS = ("kappa_x1", "kappa_x2", "d_x1_nu1", "alpha_x1_nu1", "d_x2_nu1", "alpha_x2_nu1",
              "d_x1_nu2", "alpha_x1_nu2", "d_x2_nu2", "alpha_x2_nu2")

KEEP, TARGET, RAMP = 0, 1, 2


def axis_codes(n: int, m: int, thickness: int, offset: int, tf: np.ndarray):
    """(code[n], weight[n]) of one 1-D pass: pml_builder.py:1340-1375."""
    code = np.full(n, KEEP, np.int8)
    w = np.zeros(n)
    code[: m + offset + thickness] = TARGET
    code[n - m - thickness - offset:] = TARGET
    up_start, up_end = m + offset - 1, m + offset + thickness
    code[up_start:up_end] = RAMP
    w[up_start:up_end] = tf[::-1]
    down_start, down_end = n - m - thickness - offset - 1, n - m - offset
    code[down_start:down_end] = RAMP
    w[down_start:down_end] = tf
    return code, w


def apply_axis(v: np.ndarray, axis: int, code: np.ndarray, w: np.ndarray, target: float) -> np.ndarray:
    shape = [1] * v.ndim
    shape[axis] = -1
    code, w = code.reshape(shape), w.reshape(shape)
    ramped = v - w * (v - target)
    return np.where(code == KEEP, v, np.where(code == TARGET, target, ramped))


def a_and_b(d, kappa, alpha, dt):
    b = np.exp(-(d / kappa + alpha) * dt)
    a = d / (kappa * (d + kappa * alpha) + 1e-10) * (b - 1)
    return a, b


def lookup(alpha_coeff, alpha_power, database, alpha_list, power_list):
    """relaxation_parameters.py:18-75, :189-242 -> dict of RELAX_KEYS, plus the (alpha, power) index arrays."""
    alpha_list, power_list = np.asarray(alpha_list).reshape(-1), np.asarray(power_list).reshape(-1)
    a = np.clip(alpha_coeff, alpha_list.min(), alpha_list.max())
    p = np.clip(alpha_power, power_list.min(), power_list.max().round(4))
    ia = np.clip(np.searchsorted(alpha_list.round(10), a), 0, len(alpha_list) - 1)
    ip = np.clip(np.searchsorted(power_list.round(10), p), 0, len(power_list) - 1)
    out = database[ia, ip, :]
    return {k: out[..., i] for i, k in enumerate(RELAX_KEYS)}, ia, ip


def build_maps(*, sound_speed, density, beta, relax, dt, dx, c0, m=8, n_pml, n_trans, use_pml=True, n_polynomial=2,
               reflection=1e-30, dcmap_full3d=True):
    """User-grid float64 maps -> {".dat" stem: float32 extended map, "dcmap": int32, plus float64 post-ramp d / alpha
    under their fw2 names ("d_x_nu1", "alpha_u_nu2", ...)}."""
    c = np.asarray(sound_speed, np.float64)
    ndim = c.ndim
    nb = m + n_pml + n_trans
    ext = tuple(n + 2 * nb for n in c.shape)
    idx = np.ix_(*[np.clip(np.arange(e) - nb, 0, n - 1) for e, n in zip(ext, c.shape)])

    def pad(a):
        return np.asarray(a, np.float64)[idx]

    r = {k: pad(relax[k]) for k in RELAX_KEYS}
    if use_pml:
        d_target = -(n_polynomial + 1) * c0 * np.log(reflection) / (2 * (dx * n_pml + dx * n_trans))
        x_full = np.linspace(0, 1, n_pml + n_trans + 1)
        x_tr = np.linspace(0, 1, n_trans + 1)
        tf = {"poly": x_full ** n_polynomial, "lin": x_full, "cos": 0.5 * (1 - np.cos(np.pi * x_tr))}
        for axis in range(ndim):
            n = ext[axis]
            cp = axis_codes(n, m, n_pml + n_trans, 0, tf["poly"])
            cl = axis_codes(n, m, n_pml + n_trans, 0, tf["lin"])
            cc = axis_codes(n, m, n_trans, n_pml, tf["cos"])
            for fam in ("x1", "x2"):
                r[f"d_{fam}_nu1"] = apply_axis(r[f"d_{fam}_nu1"], axis, *cp, d_target)
                r[f"alpha_{fam}_nu1"] = apply_axis(r[f"alpha_{fam}_nu1"], axis, *cl, 0.0)
                r[f"d_{fam}_nu2"] = apply_axis(r[f"d_{fam}_nu2"], axis, *cc, 0.0)
                r[f"alpha_{fam}_nu2"] = apply_axis(r[f"alpha_{fam}_nu2"], axis, *cc, 0.0)
    out = {}
    cx = pad(c)
    rho = pad(density)
    out["rho"] = rho.astype(np.float32)
    out["K"] = np.multiply(cx ** 2, rho).astype(np.float32)
    out["beta"] = pad(beta).astype(np.float32)
    out["kappax"] = r["kappa_x2"].astype(np.float32)
    out["kappau"] = r["kappa_x1"].astype(np.float32)
    for letter, fam in (("x", "x2"), ("u", "x1")):
        for nu in (1, 2):
            a, b = a_and_b(r[f"d_{fam}_nu{nu}"], r[f"kappa_{fam}"], r[f"alpha_{fam}_nu{nu}"], dt)
            out[f"apml{letter}{nu}"] = a.astype(np.float32)
            out[f"bpml{letter}{nu}"] = b.astype(np.float32)
            out[f"d_{letter}_nu{nu}"] = r[f"d_{fam}_nu{nu}"]
            out[f"alpha_{letter}_nu{nu}"] = r[f"alpha_{fam}_nu{nu}"]
    rnd = lambda v: np.round(np.asarray(v) + 1e-9).astype(int)  # noqa: E731  (fullwave/utils/numerical.py:22-24)
    dc = (rnd(cx) - rnd(c.min())).astype(np.int32)
    if ndim == 3 and not dcmap_full3d:      # the reference 3D binary honours only the first nX*nY entries
        flat = dc.reshape(-1)
        flat[ext[0] * ext[1]:] = 0
    out["dcmap"] = dc
    return out


def ulp_distance_f32(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Distance in float32 units in the last place (0 for identical bits and for +0 / -0)."""
    ia = np.ascontiguousarray(a, np.float32).view(np.int32).astype(np.int64)
    ib = np.ascontiguousarray(b, np.float32).view(np.int32).astype(np.int64)
    ia = np.where(ia < 0, -(ia & 0x7FFFFFFF), ia)
    ib = np.where(ib < 0, -(ib & 0x7FFFFFFF), ib)
    return np.abs(ia - ib)
