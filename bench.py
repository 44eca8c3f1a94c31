#!/usr/bin/env python
"""bench.py -- 3D FDTD Gpoint-updates/s of the Fullwave 2.5 time-stepping engine on B200 (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--grid XxYxZ]

Workload (BASELINE.json configs[4]): synthetic heterogeneous attenuating 3D medium,
extended grid `--grid` PER GPU (default 800x1240x1240 = 1.23e9 points ~ 148 GB resident), x-slab sharded
over N GPUs (weak scaling: nX = N * 800), 3-layer plane source, 1024 point sensors (modT 4), 2000 air voxels.
A "step" is one time step (inject -> fd_u -> fd_p -> record) over the whole grid.  Every array is far larger
than L2 (126 MB), so no flush is needed between steps.

Our arm: one process per GPU (torchrun for N > 1), libfw25.so kernels, NCCL halo exchange.  `value` is timed
with CUDA events with the maps already resident in HBM; `e2e` runs the same job through the public API from
pinned HOST buffers (upload + steps + sensor read-back in the timed region).
Reference arm (--impl reference): the reference's own shipped sm_100 CUDA executable (the reference has no
CPU engine) driven through its Python launcher (`fullwave.solver.launcher.Launcher`, from baseline/_ref) on a
bounded sample of the same workload (same medium recipe, smaller grid), rate by differencing two runs.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "3D FDTD Gpoint-updates/s"
UNIT = "Gpoint-updates/s"
BYTES_PER_POINT = 208            # 27 (fd_u) + 25 (fd_p) float32 words per point-update
BYTES_FD_U, BYTES_FD_P = 108, 100
MEDIUM = dict(n_pml=36, n_trans=36, block=24, seed=1234, modT=4, n_sensors=1024, n_air=2000)
WORKLOAD = ("synthetic3d_het_atten (BASELINE.json configs[4]: heterogeneous attenuating 3D medium, 3-layer plane source, "
            "1024 point sensors every 4th step, 2000 air voxels; grid in config.grid_per_gpu)")


def parse_grid(s):
    return tuple(int(v) for v in s.lower().split("x"))


def peaks():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        return float(json.loads(f.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self) -> dict:
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [r.split(",") for r in Path(self.f.name).read_text().splitlines() if r.count(",") >= 8]
        os.unlink(self.f.name)
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[1]) for r in rows)
        reasons = set()
        for r in rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower() == "active":
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][2]), "samples": len(rows),
                "power_w_max": max(float(r[3]) for r in rows), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ ours
def cpu_baseline(seconds: float = 12.0) -> dict:
    """The oracle port (oracle/fw25_oracle.c, OpenMP over all host cores) on a bounded sample of the workload:
    same medium recipe on a 96x128x128 extended grid.  A reported baseline, not the optimisation target."""
    from fullwave25_b200 import synthetic
    from oracle import oracle
    shape = (96, 128, 128)
    pb = synthetic.make_problem(shape, nT=4, modT=4, n_sensors=64, n_air=32, seed=1234, n_pml=12, n_trans=12)
    oracle.run(pb)                                       # warm-up (page faults, OpenMP pool)
    st = oracle.Stepper(pb)
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        st.step()
        n += 1
    dt = time.perf_counter() - t0
    cores = int(os.environ.get("OMP_NUM_THREADS", os.cpu_count() or 1))
    return {"value": pb.n_points * n / dt / 1e9, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n} steps of the same medium recipe on {'x'.join(map(str, shape))} ({dt:.1f} s, "
                      "oracle/fw25_oracle.c, OpenMP); the reference itself has no CPU engine"}


def host_setup_baseline() -> dict | None:
    """Host-side medium / relaxation setup of the REFERENCE (PMLBuilder.run + InputFileWriter stencil tables,
    solver.py:693-743) timed on this box's cores on a bounded grid -- BASELINE.json asks for it beside the
    engine numbers -- with the GPU map builder (fw25_mapgen) on the same grid next to it."""
    try:
        from tools import ref_objects
        return ref_objects.time_host_setup((40, 64, 64), gpu=True)
    except Exception as e:  # noqa: BLE001
        return {"unavailable": f"{type(e).__name__}: {e}"[:200]}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from fullwave25_b200 import engine, synthetic_device
    from fullwave25_b200.runtime import SlabEngine, TorchComm, gather_frames
    from fullwave25_b200.slab import SlabDriver, partition

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    engine.lib()                                          # fail loudly if libfw25.so is missing

    nXl, nY, nZ = args.grid
    gshape = (nXl * world, nY, nZ)
    slab = partition(gshape[0], world)[rank]
    K, W = args.steps, args.warmup
    nT = W + K
    t_gen = time.perf_counter()
    pb, maps = synthetic_device.make_slab(gshape, slab.gx0, slab.gx1, device=dev, nT=nT, **MEDIUM)
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t_gen
    dmaps = {k: (v if k == "pitch" else v.data_ptr()) for k, v in maps.items()}

    eng = SlabEngine(pb, slab, dev, device_maps=dmaps)
    comm = TorchComm(dist if world > 1 else None)
    main = torch.cuda.Stream(dev)
    bnd = torch.cuda.Stream(dev, priority=-1)
    drv = SlabDriver(slab, eng, comm, pb.modT, streams=(main, bnd), ndim=3)
    pts_step = gshape[0] * nY * nZ                        # whole-job points per step (extended grid)
    pts_rank = (slab.own_hi - slab.own_lo) * nY * nZ

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(W):
        drv.step()
    drv.finish()
    barrier()

    # ---- timed region: K steps, CUDA events on the launching stream, max over ranks
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = eng.eng.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kev = []
    ev0.record(main)
    if world == 1:
        for _ in range(K):                                # N = 1: bracket each sweep for the roofline numbers
            t = drv.t
            e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            eng.inject(t, main)
            e[0].record(main); eng.sweep_u(slab.own_lo, slab.own_hi, main); e[1].record(main)
            e[2].record(main); eng.sweep_p(slab.own_lo, slab.own_hi, main); e[3].record(main)
            if t % pb.modT == 0:
                eng.record(t // pb.modT, main)
            drv.t += 1
            kev.append(e)
    else:
        for _ in range(K):
            drv.step()
        drv.finish()
    ev1.record(main)
    barrier()
    clocks = sampler.stop() if sampler else None
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    launches = eng.eng.launches - l0
    value = pts_step * K / (ms * 1e-3) / 1e9

    peak, peak_src = peaks()
    roof = None
    if world == 1:
        u_ms = sum(e[0].elapsed_time(e[1]) for e in kev) / K
        p_ms = sum(e[2].elapsed_time(e[3]) for e in kev) / K
        dom, dms, dbytes = ("fd_u (k_sweep_u_ws)", u_ms, BYTES_FD_U) if u_ms >= p_ms else ("fd_p (k_sweep_p_ws)", p_ms, BYTES_FD_P)
        ach = pts_rank * dbytes / (dms * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": None, "peak_source": peak_src, "ms_per_launch": dms,
                "algorithmic_bytes_per_launch": pts_rank * dbytes,
                "fd_u": {"ms": u_ms, "GBps": pts_rank * BYTES_FD_U / (u_ms * 1e-3) / 1e9},
                "fd_p": {"ms": p_ms, "GBps": pts_rank * BYTES_FD_P / (p_ms * 1e-3) / 1e9},
                "step": {"GBps": value * BYTES_PER_POINT / world, "frac": value * BYTES_PER_POINT / world / peak}}
        prof = ROOT / "profiles" / "traffic_r01.json"       # dram bytes per launch from the committed ncu capture
        if prof.exists():
            try:
                tr = json.loads(prof.read_text())
                roof["traffic"] = tr.get("fd_u" if u_ms >= p_ms else "fd_p", {}).get("dram_bytes_per_point", 0) * pts_rank or None
                roof["traffic_source"] = ("dram__bytes_read.sum + dram__bytes_write.sum per point from " + str(tr.get("source"))
                                          + ", scaled to this launch's points")
            except Exception:  # noqa: BLE001
                pass
    else:
        roof = {"bound": "hbm", "kernel": "whole step (fd_u + fd_p + halo exchange), per GPU",
                "achieved": value * BYTES_PER_POINT / world, "peak": peak, "unit": "GB/s",
                "frac": value * BYTES_PER_POINT / world / peak, "traffic": None, "peak_source": peak_src}

    # ---- halo cost exposed: same K' steps with the transfers skipped (results invalid, timing only)
    halo = None
    if world > 1:
        Kh = min(K, 20)
        t_on, t_off = [], []
        for flag, acc in ((True, t_on), (False, t_off)):
            drv.exchange_enabled = flag
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(main)
            for _ in range(Kh):
                drv.step()
            drv.finish()
            b.record(main)
            barrier()
            x = torch.tensor([a.elapsed_time(b)], device=dev, dtype=torch.float64)
            dist.all_reduce(x, op=dist.ReduceOp.MAX)
            acc.append(float(x.item()) / Kh)
        drv.exchange_enabled = True
        planes = 18 * (int(slab.has_lo) + int(slab.has_hi))
        halo = {"exposed_ms_per_step": t_on[0] - t_off[0], "ms_per_step_with": t_on[0], "ms_per_step_without": t_off[0],
                "bytes_per_step_per_interface_direction": 18 * nY * int(maps["pitch"]) * 4, "planes_sent_rank0": planes}

    # ---- end to end through the public API: pinned HOST maps -> upload -> K steps -> sensor frames on the host
    # sensor frames of the resident run (steps 0 .. W+K-1): the end-to-end run below repeats steps 0 .. K-1 from HOST
    # maps through another upload path, so its frames must be the same bits -- a full-size parity property
    frames_chk = gather_frames(drv, eng, pb.n_frames, pb.ncoordsout, dist if world > 1 else None)
    e2e = None
    if not args.no_e2e:
        import psutil
        need = 14 * slab.n_local * nY * nZ * 4
        avail = psutil.virtual_memory().available / max(world, 1)
        e_nXl = nXl
        if need * 1.25 > avail:                           # host RAM bound: shrink the e2e slab, say so
            e_nXl = max(64, int(nXl * avail / (need * 1.25)) // 8 * 8)
        eng.close()
        del eng, drv
        if e_nXl != nXl:
            del maps
            torch.cuda.empty_cache()
            gshape_e = (e_nXl * world, nY, nZ)
            slab_e = partition(gshape_e[0], world)[rank]
            pb_e, maps_e = synthetic_device.make_slab(gshape_e, slab_e.gx0, slab_e.gx1, device=dev, nT=K, **MEDIUM)
        else:
            gshape_e, slab_e, pb_e, maps_e = gshape, slab, pb, maps
            pb_e.nT = K
            pb_e.nTic = min(pb_e.nTic, K)
            pb_e.icmat = np.ascontiguousarray(pb_e.icmat[:, : pb_e.nTic])
        host = synthetic_device.maps_to_host(maps_e, nZ, pin=True)
        del maps_e
        if e_nXl == nXl:
            del maps
        torch.cuda.empty_cache()
        import dataclasses
        pb_h = dataclasses.replace(pb_e, **{k: v.numpy() for k, v in host.items()})
        icm = torch.empty(pb_h.icmat.shape, dtype=torch.float32, pin_memory=True)   # the source signals are inputs too
        icm.numpy()[...] = pb_h.icmat
        pb_h.icmat = icm.numpy()
        h2d = sum(v.numel() * 4 for v in host.values()) + pb_h.icmat.nbytes + pb_h.icc.nbytes
        barrier()
        t0 = time.perf_counter()
        eng2 = SlabEngine(pb_h, slab_e, dev)                # H2D of the 14 maps + coordinate lists happens here
        eng2.eng.sync()
        t_setup = time.perf_counter() - t0                  # allocation + upload (part of the timed region)
        drv2 = SlabDriver(slab_e, eng2, comm, pb_h.modT, streams=(main, bnd))
        for _ in range(K):
            drv2.step()
        out = gather_frames(drv2, eng2, pb_h.n_frames, pb_h.ncoordsout, dist if world > 1 else None)
        barrier()
        dt = time.perf_counter() - t0
        x = torch.tensor([dt], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(x, op=dist.ReduceOp.MAX)
        dt = float(x.item())
        d2h = pb_h.n_frames * eng2.eng.n_local_sensors * 4
        e2e = {"value": gshape_e[0] * nY * nZ * K / dt / 1e9, "unit": UNIT,
               "h2d_bytes_per_step": h2d / K, "d2h_bytes_per_step": d2h / K, "seconds": dt,
               "setup_seconds": t_setup, "upload_GBps": h2d / t_setup / 1e9,
               "grid_per_gpu": f"{e_nXl}x{nY}x{nZ}", "launches": eng2.eng.launches,
               "finite": bool(np.isfinite(out).all()) if out is not None else None,
               "frames_identical_to_resident_run": (bool(np.array_equal(out, frames_chk[: out.shape[0]]))
                                                    if out is not None and frames_chk is not None and e_nXl == nXl
                                                    else None),
               "absmax": float(np.abs(out).max()) if out is not None and out.size else None,
               "api": "fullwave25_b200.runtime.SlabEngine(host maps) + SlabDriver.step + gather_frames"}
        eng2.close()

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"synthetic3d_het_atten {nXl}x{nY}x{nZ} extended points per GPU "
                                   f"(BASELINE.json configs[4]; global {gshape[0]}x{nY}x{nZ})",
                       "parallelism": f"x-slab x{world}", "points_per_step": pts_step,
                       "l2": "every array is >> L2 (126 MB): no flush between steps",
                       "medium_generation_s": round(t_gen, 2), "sensors": int(pb.ncoordsout), "sources": int(pb.ncoords),
                       "air_voxels": int(pb.ncoordszero), "ndmap": int(pb.ndmap), "dcmap": "per-voxel (dcmap_full3d=1)"},
            "roofline": roof, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        }
        if halo:
            line["halo"] = halo
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline()
            line["host_setup_baseline"] = host_setup_baseline()
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------- reference
def run_reference(args):
    """The reference's shipped sm_100 executable through the reference's own launcher, timed per step from its
    progress output (tools/bench_reference.py)."""
    from tools import bench_reference
    bench_reference.run_reference(args, metric=METRIC, unit=UNIT, workload=WORKLOAD, medium=MEDIUM)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--grid", type=parse_grid, default=(800, 1240, 1240), help="extended grid PER GPU")
    ap.add_argument("--ref-planes", type=int, default=None,
                    help="x planes per GPU of the reference arm's grid (default: the largest the reference can hold here)")
    ap.add_argument("--no-ref-crosscheck", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
