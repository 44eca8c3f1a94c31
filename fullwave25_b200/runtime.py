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
