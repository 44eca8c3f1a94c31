"""x-slab domain decomposition: partition rule, halo schedule and the per-rank stepping driver.

Replaces the reference engine's in-process slab partitioner + `exchange_halos_async`
(binary ASM 0x405f78-0x406204, 0x40e4a0-0x40e6ef).  Differences by design:

  * one process per GPU (torchrun), neighbour exchange with NCCL send/recv over NVLink instead of
    host-driven cudaMemcpyPeerAsync from a single thread;
  * only what the stencils read crosses the interface: after fd_u the 8 boundary planes of u and ONE plane
    each of v and w, after fd_p the 8 boundary planes of p -- 18 planes per direction per step instead of the
    reference's 128 (it also ships the 12 memory-variable arrays, which are only ever read point-wise);
  * boundary-first: the 8 planes next to each interface are swept on a high-priority stream and sent while
    the interior sweep runs on the main stream.

The partition itself follows the reference's rule (base = nX // n, the first nX % n slabs get one more
plane, 8 ghost planes per interior side) so that sensor / source / air ownership matches.

`SlabDriver` is generic over the per-rank engine (the CUDA engine on a GPU; tests drive it with a CPU
stand-in over gloo) and over the transport.
"""

from __future__ import annotations

from dataclasses import dataclass

HALO = 8
# Planes swept by a boundary launch.  Only the outer HALO planes cross the interface, but an 8-plane launch pays
# the x-marching kernels' chunk prologue (15 column loads + pipeline fill) for 8 planes of work; a full 32-plane
# chunk does not (N = 2 bench: 8-plane boundary launches cost 4.5 % of the step with the transfers fully hidden).
BOUNDARY = 32


@dataclass(frozen=True)
class Slab:
    rank: int
    n_ranks: int
    nX_global: int
    own_lo: int      # owned global planes [own_lo, own_hi)
    own_hi: int
    gx0: int         # global x of local plane 0 (owned range widened by HALO on interior sides)
    gx1: int         # one past the last local plane

    @property
    def n_local(self) -> int:
        return self.gx1 - self.gx0

    @property
    def has_lo(self) -> bool:
        return self.rank > 0

    @property
    def has_hi(self) -> bool:
        return self.rank < self.n_ranks - 1

    def as_tuple(self):
        """(nX_global, gx0, own_lo, own_hi): the fw25_slab struct."""
        return (self.nX_global, self.gx0, self.own_lo, self.own_hi)


def partition(nX: int, n_ranks: int) -> list[Slab]:
    """The reference's rule: base = nX // n, remainder spread over the first slabs (ASM 0x405f90-0x40607f)."""
    if n_ranks < 1:
        raise ValueError("n_ranks must be >= 1")
    base, rem = divmod(nX, n_ranks)
    if n_ranks > 1 and base < 2 * HALO:
        raise ValueError(f"slabs of {base} planes are thinner than two halos ({2 * HALO}); use fewer GPUs")
    out, lo = [], 0
    for r in range(n_ranks):
        hi = lo + base + (1 if r < rem else 0)
        out.append(Slab(r, n_ranks, nX, lo, hi, max(lo - HALO, 0) if r > 0 else 0,
                        min(hi + HALO, nX) if r < n_ranks - 1 else nX))
        lo = hi
    return out


class SlabDriver:
    """Steps one slab.  `eng` needs: inject(t, stream), sweep_u(lo, hi, stream), sweep_p(lo, hi, stream),
    record(frame, stream) with GLOBAL plane ranges; `planes(name, lo, hi)` returns the contiguous buffer of
    global planes [lo, hi) of state array `name` ("p","u","v","w") for the transport.
    `comm` needs: exchange(list of (send_buf, recv_buf, peer_rank), stream), record(stream) -> event,
    wait(stream, event).  `streams` = (main, boundary) or None (CPU stand-in: everything is synchronous)."""

    def __init__(self, slab: Slab, eng, comm, modT: int, streams=None, ndim: int = 3):
        self.s, self.eng, self.comm, self.modT = slab, eng, comm, modT
        # after fd_u: 8 planes of u (x-stencil of fd_p) and 1 plane of each transverse velocity (cross terms)
        self.vel_halo = (("u", HALO), ("v", 1), ("w", 1)) if ndim == 3 else (("u", HALO), ("v", 1))
        self.streams = streams
        self.t = 0
        self.exchange_enabled = True      # False: skip the transfers (to expose the halo cost; results invalid)
        self._ev_end = None
        # "concurrent" (default): boundary sweeps on the high-priority stream next to the interior sweep.
        # "serial": boundary planes first ON THE MAIN STREAM, then the interior; only the transfers use the boundary
        # stream (one sweep kernel on the GPU at a time, like a single-GPU run).  Measured equal on a B200 pair
        # (800x1240x1240 per GPU: 45.14 vs 45.38 ms/step, profiles/README.md), so the simpler overlap stays.
        import os
        self.schedule = os.environ.get("FW25_SLAB_SCHEDULE", "concurrent")

    # plane ranges -------------------------------------------------------------------------------
    def _boundary_width(self):
        s = self.s
        n = s.own_hi - s.own_lo
        sides = int(s.has_lo) + int(s.has_hi)
        return max(HALO, min(BOUNDARY, n // max(sides, 1)))

    def _boundary_ranges(self):
        s, w = self.s, self._boundary_width()
        r = []
        if s.has_lo:
            r.append((s.own_lo, min(s.own_lo + w, s.own_hi)))
        if s.has_hi:
            r.append((max(s.own_hi - w, s.own_lo), s.own_hi))
        return r

    def _interior_range(self):
        s, w = self.s, self._boundary_width()
        return (s.own_lo + (w if s.has_lo else 0), s.own_hi - (w if s.has_hi else 0))

    def _halo_ops(self, names_and_widths):
        """[(send, recv, peer)] for each neighbour: my outermost owned planes -> its ghost planes."""
        s, ops = self.s, []
        for name, w in names_and_widths:
            if s.has_lo:
                ops.append((self.eng.planes(name, s.own_lo, s.own_lo + w),
                            self.eng.planes(name, s.own_lo - w, s.own_lo), s.rank - 1))
            if s.has_hi:
                ops.append((self.eng.planes(name, s.own_hi - w, s.own_hi),
                            self.eng.planes(name, s.own_hi, s.own_hi + w), s.rank + 1))
        return ops

    # one time step --------------------------------------------------------------------------------
    def phases(self):
        """Generator form of one step: yields the halo operations [(send, recv, peer)] at the two points
        where planes must cross the interface, so that a caller driving SEVERAL slabs from one thread (the
        in-process multi-GPU path) can advance all of them in lockstep.  `step()` is the one-slab-per-process
        form.  Order per step: inject -> fd_u -> fd_p -> record, boundary planes first on `bnd`, transfers
        overlapped with the interior sweeps on `main`."""
        s, e, c, t = self.s, self.eng, self.comm, self.t
        main, bnd = self.streams if self.streams else (None, None)
        if s.n_ranks == 1:
            e.inject(t, main)
            e.sweep_u(s.own_lo, s.own_hi, main)
            e.sweep_p(s.own_lo, s.own_hi, main)
            if t % self.modT == 0:
                e.record(t // self.modT, main)
            self.t += 1
            return
        if self.schedule == "serial" and self.streams:
            yield from self._phases_serial()
            return
        if self._ev_end is not None:
            c.wait(main, self._ev_end)       # ghost p planes of the previous step have arrived
        e.inject(t, main)                    # sources / air voxels in owned AND ghost planes
        c.wait(bnd, c.record(main))
        for lo, hi in self._boundary_ranges():
            e.sweep_u(lo, hi, bnd)
        ev_bu = c.record(bnd)
        if self.exchange_enabled:
            yield self._halo_ops(self.vel_halo)
        ilo, ihi = self._interior_range()
        e.sweep_u(ilo, ihi, main)
        c.wait(bnd, c.record(main))          # boundary fd_p reads interior u/v/w (+ the ghosts just received)
        for lo, hi in self._boundary_ranges():
            e.sweep_p(lo, hi, bnd)
        ev_bp = c.record(bnd)
        if self.exchange_enabled:
            yield self._halo_ops((("p", HALO),))
        self._ev_end = c.record(bnd)
        c.wait(main, ev_bu)                  # interior fd_p reads the boundary planes' u/v/w
        e.sweep_p(ilo, ihi, main)
        if t % self.modT == 0:
            c.wait(main, ev_bp)
            e.record(t // self.modT, main)
        self.t += 1

    def _phases_serial(self):
        s, e, c, t = self.s, self.eng, self.comm, self.t
        main, bnd = self.streams
        if self._ev_end is not None:
            c.wait(main, self._ev_end)       # ghost p planes of the previous step have arrived
        e.inject(t, main)
        for lo, hi in self._boundary_ranges():
            e.sweep_u(lo, hi, main)
        c.wait(bnd, c.record(main))          # boundary u/v/w final; my ghost velocities are no longer read
        if self.exchange_enabled:
            yield self._halo_ops(self.vel_halo)
        ev_xu = c.record(bnd)
        ilo, ihi = self._interior_range()
        e.sweep_u(ilo, ihi, main)
        c.wait(main, ev_xu)                  # boundary fd_p reads the ghost velocities just received
        for lo, hi in self._boundary_ranges():
            e.sweep_p(lo, hi, main)
        c.wait(bnd, c.record(main))          # boundary p final; my ghost p planes are no longer read
        if self.exchange_enabled:
            yield self._halo_ops((("p", HALO),))
        self._ev_end = c.record(bnd)
        e.sweep_p(ilo, ihi, main)
        if t % self.modT == 0:
            e.record(t // self.modT, main)
        self.t += 1

    def step(self):
        bnd = self.streams[1] if self.streams else None
        for ops in self.phases():
            self.comm.exchange(ops, bnd)

    def finish(self):
        """Order everything still queued on the boundary stream before the caller reads results."""
        if self._ev_end is not None:
            main = self.streams[0] if self.streams else None
            self.comm.wait(main, self._ev_end)


def step_lockstep(drivers, transfer):
    """Advance several slabs (one SlabDriver each, all owned by this thread) by one time step.  At each halo
    point every driver has queued its boundary work; `transfer(all_ops)` then moves the planes, where
    all_ops[r] = [(send, recv, peer), ...] of rank r.  Used by the in-process multi-GPU runner."""
    gens = [d.phases() for d in drivers]
    while True:
        batch, live = [], 0
        for g in gens:
            try:
                batch.append(next(g))
                live += 1
            except StopIteration:
                batch.append(None)
        if live == 0:
            return
        if live != len(gens):
            raise RuntimeError("slabs fell out of lockstep")
        transfer(batch)
