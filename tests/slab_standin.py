"""CPU stand-ins that let tests drive fullwave25_b200.slab.SlabDriver without a GPU: the per-rank engine is the
oracle stepping a slab view of the problem, the transport is torch.distributed (gloo)."""

from __future__ import annotations

import dataclasses

import numpy as np

from oracle import oracle


def slab_problem(pb, slab):
    """Planes [gx0, gx1) of pb with coordinate lists shifted to local x and filtered to the local planes
    (sources / air: every local plane, ghosts included; sensors: owned planes only)."""
    sub = pb.slab(slab.gx0, slab.gx1)
    dc = np.array(sub.dcmap, copy=True)
    if pb.ndim == 3 and not pb.dcmap_full3d:       # reference rule on GLOBAL flat indices (fw25.h dcmap_full3d)
        flat = np.arange(slab.gx0 * pb.nY * pb.nZ, slab.gx1 * pb.nY * pb.nZ).reshape(dc.shape)
        dc[flat >= pb.nX * pb.nY] = 0
    def local(c, lo, hi):
        keep = (c[:, 0] >= lo) & (c[:, 0] < hi)
        out = c[keep].copy()
        out[:, 0] -= slab.gx0
        return out, keep
    icc, ks = local(pb.icc, slab.gx0, slab.gx1)
    air, _ = local(pb.icczero, slab.gx0, slab.gx1)
    outc, ko = local(pb.outc, slab.own_lo, slab.own_hi)
    sub = dataclasses.replace(sub, dcmap=dc, icc=icc, icmat=pb.icmat[ks], icczero=air, outc=outc, dcmap_full3d=True)
    return sub.normalise(), np.flatnonzero(ko)


class OracleSlabEngine:
    def __init__(self, pb, slab):
        self.slab = slab
        self.sub, self.sensor_ids = slab_problem(pb, slab)
        self.st = oracle.Stepper(self.sub)
        self.frames = {}

    def inject(self, t, stream=None): self.st.inject(t)
    def sweep_u(self, lo, hi, stream=None): self.st.sweep_u(lo - self.slab.gx0, hi - self.slab.gx0)
    def sweep_p(self, lo, hi, stream=None): self.st.sweep_p(lo - self.slab.gx0, hi - self.slab.gx0)
    def record(self, frame, stream=None): self.frames[frame] = self.st.record()

    def planes(self, name, lo, hi):
        import torch
        return torch.from_numpy(self.st.field(name))[lo - self.slab.gx0: hi - self.slab.gx0]


class GlooComm:
    def __init__(self, dist):
        self.dist = dist
        self.planes_sent = 0

    def record(self, stream): return None
    def wait(self, stream, ev): pass

    def exchange(self, ops, stream=None):
        reqs = []
        for send, recv, peer in ops:
            reqs.append(self.dist.isend(send, peer))
            reqs.append(self.dist.irecv(recv, peer))
            self.planes_sent += send.shape[0]
        for r in reqs:
            r.wait()
