"""CPU emulation of how the REFERENCE's binary behaves when it shards a grid over several GPUs.  TEST
INFRASTRUCTURE ONLY (same rules as oracle.py).

The reference's in-process x-slab mode (CUDA_VISIBLE_DEVICES=0,1,...) is not equivalent to its own
single-GPU run.  Two deviations were found by running the sm_100 binary on a B200 pair (tools/make_ref_golden.py with
FW25_REF_DEVICES=0,1, tools/ref_multi_gpu_study.py) and are restated here so that the 2-GPU golden traces
(tests/golden/ref_*_g2.npz) pin them:

  1. sensor gather on GPUs >= 1 reads plane x-1: every sensor owned by a slab other than the first returns the
     pressure of the cell one plane below it (its first recorded frame computes the flat index with the slab's
     x offset off by one; compute_genout_frame_multi, 3D PTX L1477-1570);
  2. sources and air voxels are applied only in a GPU's OWNED planes (inject_source / inject_source_zero test
     x_lo <= x < x_hi with the output range, PTX L1367-1371, L1425-1429), so the neighbour's ghost copy of such a
     cell keeps the post-fd_p value until the next exchange: the neighbour's fd_u reads a stale pressure whenever a
     source or air voxel lies within 8 planes of an interface.

The wave field is otherwise exchanged correctly (8 planes of u, v, w after fd_u and of p after fd_p).  The B200 engine
reproduces NEITHER deviation: its N-GPU runs are bit-identical to one domain.
"""

from __future__ import annotations

import dataclasses

import numpy as np

from . import oracle

HALO = 8


def _partition(nX: int, n: int):
    base, rem = divmod(nX, n)
    out, lo = [], 0
    for r in range(n):
        hi = lo + base + (1 if r < rem else 0)
        out.append((lo, hi, max(lo - HALO, 0) if r > 0 else 0, min(hi + HALO, nX) if r < n - 1 else nX))
        lo = hi
    return out


def run(pb, n_gpus: int = 2, *, owned_only_injection: bool = True, sensor_shift: int = -1) -> np.ndarray:
    """genout [n_frames, ncoordsout] as the reference binary produces it on n_gpus GPUs."""
    pb.normalise()
    parts = _partition(pb.nX, n_gpus)
    steppers, subs = [], []
    for r, (lo, hi, g0, g1) in enumerate(parts):
        sub = pb.slab(g0, g1)
        dc = np.array(sub.dcmap, copy=True)
        if pb.ndim == 3 and not pb.dcmap_full3d:
            flat = np.arange(g0 * pb.nY * pb.nZ, g1 * pb.nY * pb.nZ).reshape(dc.shape)
            dc[flat >= pb.nX * pb.nY] = 0
        a, b = (lo, hi) if owned_only_injection else (g0, g1)

        def local(c, keep_lo=a, keep_hi=b, g0=g0):
            keep = (c[:, 0] >= keep_lo) & (c[:, 0] < keep_hi)
            out = c[keep].copy()
            out[:, 0] -= g0
            return out, keep
        icc, ks = local(pb.icc)
        air, _ = local(pb.icczero)
        sub = dataclasses.replace(sub, dcmap=dc, icc=icc, icmat=pb.icmat[ks], icczero=air,
                                  outc=np.zeros((0, pb.ndim), np.int32), dcmap_full3d=True).normalise()
        subs.append(sub)
        steppers.append(oracle.Stepper(sub))

    def exchange(names, width):
        for r in range(n_gpus - 1):
            (lo0, hi0, g00, _), (lo1, _, g01, _) = parts[r], parts[r + 1]
            for name, w in zip(names, width):
                A, B = steppers[r].field(name), steppers[r + 1].field(name)
                B[hi0 - w - g01: hi0 - g01] = A[hi0 - w - g00: hi0 - g00]       # r's last owned planes -> r+1's ghosts
                A[lo1 - g00: lo1 + w - g00] = B[lo1 - g01: lo1 + w - g01]       # r+1's first owned planes -> r's ghosts

    vel = ("u", "v", "w") if pb.ndim == 3 else ("u", "v")
    M = HALO
    rim_hi = [pb.nX - M, pb.nY - M] + ([pb.nZ - M] if pb.ndim == 3 else [])
    out = np.zeros((oracle.n_frames(pb), pb.ncoordsout), np.float32)
    owner = np.searchsorted([p[1] for p in parts], pb.outc[:, 0], side="right")
    for t in range(pb.nT):
        for r, (lo, hi, g0, g1) in enumerate(parts):
            steppers[r].inject(t)
            steppers[r].sweep_u(lo - g0, hi - g0)
        exchange(vel, (HALO,) * len(vel))
        for r, (lo, hi, g0, g1) in enumerate(parts):
            steppers[r].sweep_p(lo - g0, hi - g0)
        exchange(("p",), (HALO,))
        if t % pb.modT == 0:
            f = t // pb.modT
            for i, c in enumerate(pb.outc):
                if any(v < M or v >= h for v, h in zip(c, rim_hi)):
                    continue                                                    # rim sensors read 0
                r = int(owner[i])
                x = int(c[0]) + (sensor_shift if r > 0 else 0)
                out[f, i] = steppers[r].field("p")[(x - parts[r][2],) + tuple(int(v) for v in c[1:])]
    return out
