"""Per-rank runtime around the C-ABI engine for x-slab runs: torch owns the device buffers, streams and the
NCCL process group (plumbing); every kernel that touches the wave field is libfw25.so's.

`SlabEngine` adapts `engine.Engine` to what `slab.SlabDriver` needs (global plane ranges, stream handles,
contiguous plane views of the state arrays for the transport); `TorchComm` is the transport: NCCL
send/recv of whole planes between x-neighbours, queued on the boundary stream.
"""

from __future__ import annotations

import numpy as np

from . import engine as _engine
from .slab import Slab, SlabDriver


class SlabEngine:
    """One rank's CUDA engine.  State arrays p,u,v,w are torch tensors [n_local, nY, pitch] handed to the
    engine as caller-owned device arrays (fw25_problem.ext_*), so NCCL reads and writes halo planes in place."""

    def __init__(self, pb, slab: Slab, device, *, device_maps=None, variant: int = 0):
        import torch
        self.torch = torch
        self.slab = slab
        self.pb = pb
        self.device = torch.device(device)
        n_fast = pb.nZ if pb.ndim == 3 else pb.nY
        pitch = int(_engine.lib().fw25_pitch(n_fast))
        rows = pb.nY if pb.ndim == 3 else 1
        names = ("p", "u", "v", "w") if pb.ndim == 3 else ("p", "u", "v")
        self.state = {k: torch.zeros((slab.n_local, rows, pitch), dtype=torch.float32, device=self.device)
                      for k in names}
        ext = {k: t.data_ptr() for k, t in self.state.items()}
        self.eng = _engine.Engine(pb, device=self.device.index or 0, slab=slab.as_tuple(),
                                  device_maps=device_maps, ext_state=ext, variant=variant)

    @staticmethod
    def _h(stream):
        return 0 if stream is None else int(stream.cuda_stream)

    def inject(self, t, stream): self.eng.inject(t, self._h(stream))
    def sweep_u(self, lo, hi, stream): self.eng.sweep_u(lo, hi, self._h(stream))
    def sweep_p(self, lo, hi, stream): self.eng.sweep_p(lo, hi, self._h(stream))
    def record(self, frame, stream): self.eng.record(frame, self._h(stream))

    def planes(self, name, lo, hi):
        g0 = self.slab.gx0
        return self.state[name][lo - g0: hi - g0]

    def close(self):
        self.eng.close()


class TorchComm:
    """Stream ordering with CUDA events + neighbour exchange with torch.distributed (NCCL on GPUs)."""

    def __init__(self, dist=None):
        import torch
        self.torch = torch
        self.dist = dist
        self.bytes_sent = 0

    def record(self, stream):
        ev = self.torch.cuda.Event()
        ev.record(stream)
        return ev

    def wait(self, stream, ev):
        stream.wait_event(ev)

    def exchange(self, ops, stream):
        if not ops:
            return
        d = self.dist
        with self.torch.cuda.stream(stream):
            p2p = []
            for send, recv, peer in ops:
                p2p.append(d.P2POp(d.isend, send, peer))
                p2p.append(d.P2POp(d.irecv, recv, peer))
                self.bytes_sent += send.numel() * 4
            for w in d.batch_isend_irecv(p2p):
                w.wait()


def gather_frames(drv: SlabDriver, eng: SlabEngine, n_frames: int, ncoordsout: int, dist=None) -> np.ndarray | None:
    """Assemble genout [n_frames, ncoordsout] in GLOBAL outc order on rank 0 (the reference writes frames in
    that order whatever the GPU count); other ranks return None."""
    import torch
    drv.finish()
    torch.cuda.synchronize(eng.device)     # the sweeps ran on torch streams, not on the engine's own stream
    eng.eng.sync()
    local = eng.eng.read_frames(0, n_frames) if n_frames else np.zeros((0, eng.eng.n_local_sensors), np.float32)
    ids = eng.eng.local_sensor_ids()
    if dist is None or eng.slab.n_ranks == 1:
        out = np.zeros((n_frames, ncoordsout), np.float32)
        out[:, ids] = local
        return out
    parts = [None] * eng.slab.n_ranks
    dist.all_gather_object(parts, (ids, local))
    if eng.slab.rank != 0:
        return None
    out = np.zeros((n_frames, ncoordsout), np.float32)
    for pid, pl in parts:
        out[:, pid] = pl
    return out


# ------------------------------------------------------------------------------------------------------------
# In-process multi-GPU: one host thread drives one slab per device -- the reference's own model (a single
# process looping over cudaSetDevice) and what `cuda_device_id=[0, 1, ...]` selects through the
# launcher.  Planes move with peer-to-peer copies (torch `copy_` between devices, cudaMemcpyPeerAsync
# underneath) queued on the SENDER's boundary stream.

class _LocalOrder:
    """Event plumbing for SlabDriver when every slab lives in this process."""

    def __init__(self, torch):
        self.torch = torch

    def record(self, stream):
        ev = self.torch.cuda.Event()
        ev.record(stream)
        return ev

    def wait(self, stream, ev):
        stream.wait_event(ev)

    def exchange(self, ops, stream):   # never called: run_local drives the phases in lockstep
        raise RuntimeError("use step_lockstep for in-process slabs")


def run_local(pb, device_ids, *, variant: int = 0, return_stats: bool = False):
    """Whole job on several GPUs of this process.  pb: host Problem (whole grid).  Returns genout
    [n_frames, ncoordsout] in global outc order (and stats)."""
    import time

    import torch

    from .slab import partition, step_lockstep

    n = len(device_ids)
    slabs = partition(pb.nX, n)
    pb.normalise()
    t0 = time.perf_counter()
    engines, drivers, streams = [], [], []
    order = _LocalOrder(torch)
    for slab, dev_id in zip(slabs, device_ids):
        dev = torch.device("cuda", int(dev_id))
        with torch.cuda.device(dev):
            sub = pb.slab(slab.gx0, slab.gx1).normalise()
            eng = SlabEngine(sub, slab, dev, variant=variant)
            main, bnd = torch.cuda.Stream(dev), torch.cuda.Stream(dev, priority=-1)
        engines.append(eng)
        streams.append((main, bnd))
        drivers.append(SlabDriver(slab, eng, order, pb.modT, streams=(main, bnd), ndim=pb.ndim))
    t_setup = time.perf_counter() - t0

    def transfer(batch):
        # batch[r] = [(send, recv, peer), ...]; the k-th op of r towards p pairs with the k-th op of p towards r
        ready = []
        for r, (main, bnd) in enumerate(streams):
            with torch.cuda.device(engines[r].device):
                ev = torch.cuda.Event()
                ev.record(bnd)         # my boundary planes are final AND my ghost planes are no longer read
            ready.append(ev)
        recv_of = {}
        for r, ops in enumerate(batch):
            seen = {}
            for _send, recv, peer in ops:
                k = seen.get(peer, 0)
                seen[peer] = k + 1
                recv_of[(peer, r, k)] = recv          # what `peer` sends to `r` lands here
        done = [[] for _ in range(n)]
        for r, ops in enumerate(batch):
            bnd = streams[r][1]
            seen = {}
            with torch.cuda.device(engines[r].device), torch.cuda.stream(bnd):
                for send, _recv, peer in ops:
                    k = seen.get(peer, 0)
                    seen[peer] = k + 1
                    bnd.wait_event(ready[peer])
                    recv_of[(r, peer, k)].copy_(send, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(bnd)
            for peer in seen:
                done[peer].append(ev)
        for r, evs in enumerate(done):
            for ev in evs:
                streams[r][1].wait_event(ev)

    t1 = time.perf_counter()
    for _ in range(pb.nT):
        step_lockstep(drivers, transfer)
    out = np.zeros((pb.n_frames, pb.ncoordsout), np.float32)
    for drv, eng in zip(drivers, engines):
        with torch.cuda.device(eng.device):
            drv.finish()
            torch.cuda.synchronize(eng.device)
            if pb.n_frames:
                out[:, eng.eng.local_sensor_ids()] = eng.eng.read_frames(0, pb.n_frames)
    t_loop = time.perf_counter() - t1
    stats = {"setup_ms": t_setup * 1e3, "loop_ms": t_loop * 1e3, "d2h_ms": 0.0,
             "kernel_launches": sum(e.eng.launches for e in engines), "h2d_bytes": 0, "d2h_bytes": out.nbytes,
             "point_updates": pb.n_points * pb.nT, "n_devices": n}
    for e in engines:
        e.close()
    return (out, stats) if return_stats else out
