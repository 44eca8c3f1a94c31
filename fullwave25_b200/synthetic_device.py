"""Device-side generator of the synthetic heterogeneous attenuating 3D medium (BASELINE.json configs[4],
SURVEY.md 8(d) item 5) for grids too large to build with numpy on the host.

Same *form* of inputs as synthetic.make_problem (tissue blocks -> c, rho, beta, K; two relaxation
mechanisms per family; nu = 1 ramped into a CPML damping profile and nu = 2 to zero inside the
boundary layer; a, b from the closed form of /root/reference/fullwave/solver/pml_builder.py:794-810),
evaluated in float32 with torch on the GPU, plane chunk by plane chunk, straight into the engine's
padded [nX][nY][pitch] layout, so that the 14 maps of a 150 GB problem never exist on the host.  The
tissue label of a voxel is a hash of its GLOBAL block coordinates, so any x-slab of the grid can be
generated independently by the rank that owns it.

torch is used here for device memory and elementwise setup math only (plumbing); the time-stepping
engine is libfw25.so.  This is a workload generator for bench.py and the full-size tests -- the
reference's medium builder stays in Python upstream and is out of scope.
"""

from __future__ import annotations

import math

import numpy as np

from . import stencil, synthetic
from .problem import MAP_NAMES, Problem

M = 8
C_MIN, C_MAX = 1412, 1613          # integer sound speeds spanned by synthetic.TISSUES


def pitch_of(n_fast: int) -> int:
    return (n_fast + 31) // 32 * 32


def _profiles(n: int, n_pml: int, n_trans: int):
    xi, tr = synthetic.boundary_depth(n, n_pml, n_trans)
    return xi.astype(np.float32), tr.astype(np.float32)


def _lists(nX, nY, nZ, nb, nTic, dt, dx, *, f0, c0, seed, n_sensors, n_air, source_layers, amp):
    ys, zs = np.arange(nb, nY - nb, dtype=np.int32), np.arange(nb, nZ - nb, dtype=np.int32)
    yy, zz = np.meshgrid(ys, zs, indexing="ij")
    icc = np.concatenate([np.stack([np.full(yy.size, nb + l, np.int32), yy.ravel(), zz.ravel()], axis=1)
                          for l in range(source_layers)])
    pulse = synthetic.tone_burst(nTic, dt, f0, amp=amp).astype(np.float32)
    rows = np.zeros((source_layers, nTic), np.float32)
    for l in range(source_layers):
        shift = int(round(l * dx / c0 / dt))
        if shift < nTic:
            rows[l, shift:] = pulse[: nTic - shift]
    icmat = np.repeat(rows, yy.size, axis=0)
    rng = np.random.default_rng(seed)

    def pick(n):
        x = rng.integers(nb + source_layers, nX - nb, size=n)
        y = rng.integers(nb, nY - nb, size=n)
        z = rng.integers(nb, nZ - nb, size=n)
        flat = np.unique((x.astype(np.int64) * nY + y) * nZ + z)
        return np.stack(np.unravel_index(flat, (nX, nY, nZ)), axis=1).astype(np.int32)

    outc = pick(n_sensors)
    icczero = pick(n_air) if n_air else np.zeros((0, 3), np.int32)
    return icc, icmat, outc, icczero


def make_slab(global_shape, gx0: int, gx1: int, *, device, nT: int, f0: float = 1e6, c0: float = 1540.0,
              ppw: int = 12, cfl: float = 0.2, n_pml: int = 36, n_trans: int = 36, block: int = 24,
              seed: int = 1234, modT: int = 4, n_sensors: int = 1024, n_air: int = 2000,
              source_layers: int = 3, amp: float = 1e5, chunk: int = 8, with_lists: bool = True):
    """Returns (pb, maps): pb is a Problem holding the scalars, stencil table and the GLOBAL coordinate
    lists (its map fields are None); maps = {name: torch tensor [gx1-gx0, nY, pitch]} incl. "dcmap"
    (int32) and "pitch".  Planes are global x in [gx0, gx1).  with_lists=False leaves the coordinate lists empty
    (callers that generate a grid slab by slab need them once)."""
    import torch

    nX, nY, nZ = (int(s) for s in global_shape)
    nXl = gx1 - gx0
    pitch = pitch_of(nZ)
    dx = c0 / f0 / ppw
    dt = cfl * dx / c0
    nb = M + n_pml + n_trans
    dev = torch.device(device)
    f32 = torch.float32

    tis = torch.tensor(synthetic.TISSUES, dtype=f32, device=dev)
    tab = torch.tensor(synthetic.relaxation_table(f0), dtype=f32, device=dev)      # [2, ntis, 5]
    ntis = tis.shape[0]
    prof = [_profiles(n, n_pml, n_trans) for n in (nX, nY, nZ)]
    xi_y, tr_y = (torch.tensor(a, device=dev) for a in prof[1])
    xi_z, tr_z = (torch.tensor(a, device=dev) for a in prof[2])
    xi_yz = torch.maximum(xi_y[:, None], xi_z[None, :])
    tr_yz = torch.maximum(tr_y[:, None], tr_z[None, :])
    L = (n_pml + n_trans) * dx
    d_pml = -(2 + 1) * c0 * math.log(1e-30) / (2 * L) if n_pml > 0 else 0.0

    # block coordinates clamp to the user domain so the boundary layer replicates the edge tissue
    def blk(n):
        idx = torch.arange(n, device=dev).clamp(nb, n - nb - 1)
        return (idx // block).to(torch.int64)
    by, bz = blk(nY), blk(nZ)
    bx_all = blk(nX)

    maps = {name: torch.zeros((nXl, nY, pitch), dtype=f32, device=dev) for name in MAP_NAMES}
    maps["dcmap"] = torch.zeros((nXl, nY, pitch), dtype=torch.int32, device=dev)

    for x0 in range(0, nXl, chunk):
        x1 = min(x0 + chunk, nXl)
        gx = torch.arange(gx0 + x0, gx0 + x1, device=dev)
        bx = bx_all[gx]
        h = (bx[:, None, None] * 73856093) ^ (by[None, :, None] * 19349663) ^ (bz[None, None, :] * 83492791) ^ (seed * 2654435761)
        h = (h ^ (h >> 13)) * 1274126177
        lab = ((h ^ (h >> 16)) & 0x7FFFFFFF) % ntis                       # [cx, nY, nZ]
        jit = (gx[:, None, None] * 1103515245 + torch.arange(nY, device=dev)[None, :, None] * 12345 +
               torch.arange(nZ, device=dev)[None, None, :] * 2654435761 + seed) & 0xFFFF
        c = tis[lab, 0] + (jit.to(f32) / 65535.0 - 0.5) * 0.8
        rho = tis[lab, 1]
        sl = (slice(x0, x1), slice(None), slice(0, nZ))
        maps["rho"][sl] = rho
        maps["beta"][sl] = tis[lab, 2]
        maps["K"][sl] = c * c * rho
        maps["dcmap"][sl] = (torch.floor(c + 0.5).to(torch.int32) - C_MIN).clamp(0, C_MAX - C_MIN)
        xi_x = torch.tensor(prof[0][0][gx0 + x0: gx0 + x1], device=dev)
        tr_x = torch.tensor(prof[0][1][gx0 + x0: gx0 + x1], device=dev)
        xi = torch.maximum(xi_x[:, None, None], xi_yz[None])
        tr = torch.maximum(tr_x[:, None, None], tr_yz[None])
        keep = 1.0 - tr
        for fam, tag in ((0, "x"), (1, "u")):
            t = tab[fam][lab]                                            # [cx, nY, nZ, 5]
            kappa = 1.0 + (t[..., 0] - 1.0) * keep
            d1 = t[..., 1] * keep + d_pml * xi * xi
            al1 = t[..., 2] * keep
            d2 = t[..., 3] * keep
            al2 = t[..., 4] * keep
            b1 = torch.exp(-(d1 / kappa + al1) * dt)
            a1 = d1 / (kappa * (d1 + kappa * al1) + 1e-10) * (b1 - 1)
            b2 = torch.exp(-(d2 / kappa + al2) * dt)
            a2 = d2 / (kappa * (d2 + kappa * al2) + 1e-10) * (b2 - 1)
            maps[f"kappa{tag}"][sl] = kappa
            maps[f"apml{tag}1"][sl], maps[f"bpml{tag}1"][sl] = a1, b1
            maps[f"apml{tag}2"][sl], maps[f"bpml{tag}2"][sl] = a2, b2
        del h, lab, jit, c, rho, xi, tr, keep, t, kappa, d1, al1, d2, al2, a1, b1, a2, b2
    maps["pitch"] = pitch

    # stencil table over the whole tissue range (the reference evaluates one column per integer sound speed)
    dim = C_MAX - C_MIN
    dmap = stencil.d_map(float(C_MIN), dim, dt, dx, is_3d=True).astype(np.float32)

    # coordinate lists (GLOBAL, row-major like np.where): plane source, point sensors, air voxels
    nTic = min(nT, int(np.ceil(2.0 / f0 / dt)) + 1)
    if with_lists:
        icc, icmat, outc, icczero = _lists(nX, nY, nZ, nb, nTic, dt, dx, f0=f0, c0=c0, seed=seed, n_sensors=n_sensors,
                                           n_air=n_air, source_layers=source_layers, amp=amp)
    else:
        icc = outc = icczero = np.zeros((0, 3), np.int32)
        icmat = np.zeros((0, nTic), np.float32)

    none = {name: None for name in MAP_NAMES}
    pb = Problem(ndim=3, nX=nXl, nY=nY, nZ=nZ, nT=nT, nTic=nTic, modT=modT, ndmap=dim + 1,
                 dX=float(np.float32(dx)), dT=float(np.float32(dt)), **none, dmap=dmap, dcmap=None,
                 icc=icc, icmat=icmat, outc=outc, icczero=icczero, extra={"c0": c0}, dcmap_full3d=True)
    return pb, maps


def maps_to_host(maps: dict, nZ: int, pin: bool = True) -> dict:
    """Dense (pitch-free) HOST copies of the device maps, as numpy arrays over pinned memory."""
    import torch
    out = {}
    for name, t in maps.items():
        if name == "pitch":
            continue
        h = torch.empty(t.shape[:2] + (nZ,), dtype=t.dtype, pin_memory=pin)
        h.copy_(t[..., :nZ])
        out[name] = h
    return out


def make_user_medium(user_shape, *, device, block: int = 24, seed: int = 1234, pin: bool = True, chunk: int = 16,
                     x_range: tuple | None = None):
    """The synthetic tissue medium as a USER-grid `fullwave.Medium` would hold it -- sound_speed, density, beta,
    alpha_coeff, alpha_power -- as float32 HOST arrays (pinned), generated on the device chunk by chunk.  Input of the
    GPU map builder (`mapgen.MediumSpec` with a look-up table).  x_range = (x0, x1): only those x planes of the grid
    (a rank of an x-sharded run keeps the planes its slab reads; labels hash the GLOBAL block coordinates).
    Returns (maps, c_min, c_max, pinned tensors) -- c_min / c_max over the planes generated."""
    import torch

    nx, ny, nz = (int(s) for s in user_shape)
    x_lo, x_hi = (0, nx) if x_range is None else (int(x_range[0]), int(x_range[1]))
    dev = torch.device(device)
    f32 = torch.float32
    tis = torch.tensor(synthetic.TISSUES, dtype=f32, device=dev)
    ntis = tis.shape[0]
    names = ("sound_speed", "density", "beta", "alpha_coeff", "alpha_power")
    host = {n: torch.empty((x_hi - x_lo, ny, nz), dtype=f32, pin_memory=pin) for n in names}
    by = (torch.arange(ny, device=dev) // block).to(torch.int64)
    bz = (torch.arange(nz, device=dev) // block).to(torch.int64)
    c_min, c_max = float("inf"), float("-inf")
    for x0 in range(x_lo, x_hi, chunk):
        x1 = min(x0 + chunk, x_hi)
        gx = torch.arange(x0, x1, device=dev)
        bx = (gx // block).to(torch.int64)
        h = (bx[:, None, None] * 73856093) ^ (by[None, :, None] * 19349663) ^ (bz[None, None, :] * 83492791) ^ (seed * 2654435761)
        h = (h ^ (h >> 13)) * 1274126177
        lab = ((h ^ (h >> 16)) & 0x7FFFFFFF) % ntis
        jit = (gx[:, None, None] * 1103515245 + torch.arange(ny, device=dev)[None, :, None] * 12345 +
               torch.arange(nz, device=dev)[None, None, :] * 2654435761 + seed) & 0xFFFF
        c = tis[lab, 0] + (jit.to(f32) / 65535.0 - 0.5) * 0.8
        c_min, c_max = min(c_min, float(c.min())), max(c_max, float(c.max()))
        for n, v in zip(names, (c, tis[lab, 1], tis[lab, 2], tis[lab, 3], tis[lab, 4])):
            host[n][x0 - x_lo:x1 - x_lo].copy_(v, non_blocking=True)
        if dev.type == "cuda":
            torch.cuda.synchronize(dev)
        del h, lab, jit, c
    return {n: t.numpy() for n, t in host.items()}, c_min, c_max, host
