#!/usr/bin/env python
"""bench.py -- 3D FDTD Gpoint-updates/s of the Fullwave 2.5 time-stepping engine on B200 (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--grid XxYxZ] [--config synthetic|examples]

Workload (BASELINE.json configs[4]): synthetic heterogeneous attenuating 3D medium,
extended grid `--grid` PER GPU (default 800x1240x1240 = 1.23e9 points ~ 148 GB resident), x-slab sharded
over N GPUs (weak scaling: nX = N * 800), 3-layer plane source, 1024 point sensors (modT 4), 2000 air voxels.
A "step" is one time step (inject -> fd_u -> fd_p -> record) over the whole grid.  Every array is far larger
than L2 (126 MB), so no flush is needed between steps.

Our arm: one process per GPU (torchrun for N > 1), libfw25.so kernels, NCCL halo exchange.
  value        K steps timed with CUDA events, maps resident in HBM, max over ranks
  roofline     the dominant sweep kernel against the measured HBM copy bandwidth (N = 1), traffic from the committed ncu capture
  e2e          N = 1: the whole job through `fw25_run_medium` (C-ABI) from the USER-grid medium in pinned host memory --
               allocation, host->device copies, map generation, K steps, frames back on the host (the device memory is
               given back by a reaper thread after the call has returned) -- checked bit for bit against the sequential
               path; N > 1: every rank uploads the user-grid planes of its slab and builds its slab of the maps on its GPU
               (`fw25_mapgen_slab_begin / _finish`), NCCL halos, frames gathered on rank 0; `hostmaps_variant`: the 14
               engine maps uploaded from pinned host memory (round 1's path)
  same_grid    our rate on the grid the reference arm runs (the largest it can hold), in the reference binary's own dcmap
               mode, with the sha256 of the sensor frames: equal to the reference arm's `genout_sha256` = bit-identical
  parity_n_vs_1 (N > 1) N ranks vs rank 0 alone on a mid-size grid with sources / sensors / air voxels ON the interfaces
  halo         (N > 1) exposed halo-exchange time: the same steps with and without the transfers
  strong       (N > 1) strong scaling: the ONE-GPU grid split over the N GPUs
  cpu_baseline the oracle port on the host cores; host_setup_baseline: the reference's host-side map building at the
               wave_3d example's size, with fw25_mapgen beside it
Reference arm (--impl reference): the reference's own shipped sm_100 CUDA executable (it has no CPU engine) through its
own `fullwave.solver.launcher.Launcher` (baseline/_ref), same workload recipe on the largest grid the reference can
hold (560x1240x1240 per GPU on a 1- or 2-GPU box), timed per step from its progress prints (tools/bench_reference.py).
--config examples: BASELINE.json configs 1-4, the reference's shipped example scripts verbatim on both engines.
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
    solver.py:693-743) timed on this box's cores at BASELINE.json configs[3]'s size (120^3 user grid -> 280^3) --
    BASELINE.json asks for it beside the engine numbers, with the core count -- and the GPU map builder
    (fw25_mapgen) on the same grid next to it."""
    try:
        from tools import ref_objects
        return ref_objects.time_host_setup((120, 120, 120), gpu=True)    # the wave_3d example's user grid (config 4)
    except Exception as e:  # noqa: BLE001
        return {"unavailable": f"{type(e).__name__}: {e}"[:200]}


def parity_problem(world: int):
    """A mid-size problem for the N-vs-1 bit-identity check that precedes the timed runs: >= 3 x-marching chunks per
    slab, source cells, sensors and air voxels ON both sides of every interface (ghost-plane injection, sensor
    ownership) and in the planes the halo carries."""
    from fullwave25_b200 import synthetic
    from fullwave25_b200.slab import partition
    pb = synthetic.make_problem((world * 100, 56, 64), nT=40, modT=2, seed=77, n_pml=6, n_trans=4, n_sensors=256,
                                n_air=64)
    rng = np.random.default_rng(7)
    out, air, src = [], [], []
    for sl in partition(pb.nX, world)[:-1]:
        e = sl.own_hi
        for x in (e - 8, e - 1, e, e + 7):
            yz = rng.integers(20, 34, size=(8, 2))
            out += [(x, y, z) for y, z in yz]
            air.append((x, int(yz[0, 0]) + 1, int(yz[0, 1]) + 2))
            src.append((x, int(yz[1, 0]) - 1, int(yz[1, 1]) + 3))
    pb.outc = np.concatenate([pb.outc, np.asarray(out, np.int32).reshape(-1, 3)])
    pb.icczero = np.concatenate([pb.icczero, np.asarray(air, np.int32).reshape(-1, 3)])
    pb.icc = np.concatenate([pb.icc, np.asarray(src, np.int32).reshape(-1, 3)])
    pb.icmat = np.concatenate([pb.icmat, np.repeat(0.5 * pb.icmat[:1], len(src), axis=0)])
    return pb.normalise(), len(out), len(src)


def run_ours(args):
    import hashlib

    import torch
    import torch.distributed as dist
    from fullwave25_b200 import engine, synthetic_device
    from fullwave25_b200.runtime import SlabEngine, TorchComm, gather_frames
    from fullwave25_b200.slab import SlabDriver, partition
    from tools.bench_reference import choose_planes, ref_nT

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
    D = dist if world > 1 else None
    comm = TorchComm(D)
    main = torch.cuda.Stream(dev)
    bnd = torch.cuda.Stream(dev, priority=-1)
    nXl, nY, nZ = args.grid
    K, W = args.steps, args.warmup
    peak, peak_src = peaks()
    same_planes, _ = choose_planes(world, nY, nZ, args.ref_planes)      # before this process pins any host memory

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return float(x)
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(drv, n):
        """n steps between CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks."""
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(main)
        for _ in range(n):
            drv.step()
        drv.finish()
        b.record(main)
        barrier()
        return max_over_ranks(a.elapsed_time(b))

    def resident(gshape, nT, *, full3d=True):
        """Slab engine of this rank over device-generated maps of `gshape` (maps adopted in place)."""
        slab = partition(gshape[0], world)[rank]
        t0 = time.perf_counter()
        pb, maps = synthetic_device.make_slab(gshape, slab.gx0, slab.gx1, device=dev, nT=nT, **MEDIUM)
        torch.cuda.synchronize()
        t_gen = time.perf_counter() - t0
        pb.dcmap_full3d = full3d
        dmaps = {k: (v if k == "pitch" else v.data_ptr()) for k, v in maps.items()}
        eng = SlabEngine(pb, slab, dev, device_maps=dmaps)
        drv = SlabDriver(slab, eng, comm, pb.modT, streams=(main, bnd), ndim=3)
        return slab, pb, maps, eng, drv, t_gen

    # ---- (0) N ranks == 1 rank, bit for bit, before anything is timed (mid-size grid, everything on the interfaces)
    parity = None
    if world > 1:
        pbp, n_if_sens, n_if_src = parity_problem(world)
        slab = partition(pbp.nX, world)[rank]
        engp = SlabEngine(pbp.slab(slab.gx0, slab.gx1).normalise(), slab, dev)
        drvp = SlabDriver(slab, engp, comm, pbp.modT, streams=(main, bnd), ndim=3)
        for _ in range(pbp.nT):
            drvp.step()
        got = gather_frames(drvp, engp, pbp.n_frames, pbp.ncoordsout, D)
        engp.close()
        if rank == 0:
            one, _ = engine.run(pbp, device_ids=(local,))
            parity = {"grid": "x".join(map(str, pbp.shape)), "steps": pbp.nT, "frames": pbp.n_frames,
                      "sensors": pbp.ncoordsout, "sensors_on_interfaces": n_if_sens, "sources_on_interfaces": n_if_src,
                      "identical": bool(np.array_equal(got, one)), "absmax": float(np.abs(one).max()),
                      "what": f"{world} ranks (NCCL halos) vs rank 0 alone, all sensor frames, same bits"}
        barrier()

    # ---- (1) headline: K steps on the full-size grid, maps resident in HBM
    gshape = (nXl * world, nY, nZ)
    slab, pb, maps, eng, drv, t_gen = resident(gshape, W + K)
    pts_step = gshape[0] * nY * nZ                        # whole-job points per step (extended grid)
    pts_rank = (slab.own_hi - slab.own_lo) * nY * nZ
    for _ in range(W):
        drv.step()
    drv.finish()
    barrier()
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
    ms = max_over_ranks(ev0.elapsed_time(ev1))
    launches = eng.eng.launches - l0
    value = pts_step * K / (ms * 1e-3) / 1e9
    # frames of steps 0 .. W+K-1 BEFORE anything else touches the ring (the halo-exposure loop below runs extra steps)
    frames_resident = gather_frames(drv, eng, pb.n_frames, pb.ncoordsout, D)

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
        roof.update(traffic_from_profile("fd_u" if u_ms >= p_ms else "fd_p", (nXl, nY, nZ)))
    else:
        roof = {"bound": "hbm", "kernel": "whole step (fd_u + fd_p + halo exchange), per GPU",
                "achieved": value * BYTES_PER_POINT / world, "peak": peak, "unit": "GB/s",
                "frac": value * BYTES_PER_POINT / world / peak, "traffic": None, "peak_source": peak_src}

    # ---- (2) halo cost exposed: the same steps with the transfers skipped (results invalid from here on, timing only)
    halo = None
    if world > 1:
        Kh = min(K, 20)
        t_on = timed(drv, Kh) / Kh
        drv.exchange_enabled = False
        t_off = timed(drv, Kh) / Kh
        drv.exchange_enabled = True
        halo = {"exposed_ms_per_step": t_on - t_off, "ms_per_step_with": t_on, "ms_per_step_without": t_off,
                "bytes_per_step_per_interface_direction": 18 * nY * int(maps["pitch"]) * 4,
                "planes_sent_rank0": 18 * (int(slab.has_lo) + int(slab.has_hi))}
    config = {"workload": WORKLOAD, "grid_per_gpu": f"{nXl}x{nY}x{nZ}", "global_grid": "x".join(map(str, gshape)),
              "parallelism": f"x-slab x{world}", "points_per_step": pts_step,
              "l2": "every array is >> L2 (126 MB): no flush between steps",
              "medium_generation_s": round(t_gen, 2), "sensors": int(pb.ncoordsout), "sources": int(pb.ncoords),
              "air_voxels": int(pb.ncoordszero), "ndmap": int(pb.ndmap), "dcmap": "per-voxel (dcmap_full3d=1)"}
    pb_full, maps_full = pb, maps
    eng.close()
    del eng, drv, maps

    def free_device():
        import gc
        gc.collect()
        torch.cuda.empty_cache()

    # ---- (3) strong scaling: the ONE-GPU grid split over the N GPUs (BASELINE.json configs[4] "weak and strong")
    strong = None
    if world > 1 and not args.no_strong:
        del maps_full
        maps_full = None
        free_device()
        sshape = (nXl, nY, nZ)
        s_slab, s_pb, s_maps, s_eng, s_drv, _ = resident(sshape, W + K)
        timed(s_drv, W)
        s_ms = timed(s_drv, K)
        s_drv.exchange_enabled = False
        s_off = timed(s_drv, min(K, 20)) / min(K, 20)
        strong = {"scaling": "strong", "global_grid": "x".join(map(str, sshape)),
                  "planes_per_gpu": s_slab.own_hi - s_slab.own_lo, "boundary_planes_per_gpu":
                  sum(hi - lo for lo, hi in s_drv._boundary_ranges()),
                  "value": nXl * nY * nZ * K / (s_ms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": s_ms / K,
                  "ms_per_step_without_transfers": s_off, "exposed_halo_ms_per_step": s_ms / K - s_off,
                  "roofline_frac_per_gpu": nXl * nY * nZ * K / (s_ms * 1e-3) * BYTES_PER_POINT / world / 1e9 / peak}
        s_eng.close()
        del s_eng, s_drv, s_maps

    # ---- (4) the grid the reference arm runs (the largest it can hold), in the reference binary's own dcmap mode,
    #          same nT: the sha256 of the sensor frames must equal the one the reference arm prints
    same = None
    if same_planes and not args.no_same_grid:
        maps_full = None
        free_device()
        nTs = ref_nT(W, K, MEDIUM["modT"])
        g2 = (same_planes * world, nY, nZ)
        _, g_pb, g_maps, g_eng, g_drv, _ = resident(g2, nTs, full3d=False)
        timed(g_drv, W)
        g_ms = timed(g_drv, K)
        for _ in range(nTs - W - K):
            g_drv.step()
        fr = gather_frames(g_drv, g_eng, g_pb.n_frames, g_pb.ncoordsout, D)
        same = {"grid_per_gpu": f"{same_planes}x{nY}x{nZ}", "global_grid": "x".join(map(str, g2)),
                "value": g2[0] * nY * nZ * K / (g_ms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": g_ms / K,
                "dcmap": "reference 3D binary mode (dcmap_full3d=0)", "nT": nTs,
                "genout_sha256": hashlib.sha256(np.ascontiguousarray(fr, np.float32).tobytes()).hexdigest()
                if fr is not None else None,
                "note": "same inputs, steps and sensor list as `bench.py --impl reference`: its genout_sha256 must match"}
        g_eng.close()
        del g_eng, g_drv, g_maps
        free_device()

    # ---- (5) end to end through the public API from HOST buffers
    def free_host():
        """Give pinned host memory back to the OS (torch caches it): the two end-to-end variants pin up to 70 GB per
        rank one after the other."""
        import gc
        gc.collect()
        try:
            torch._C._host_emptyCache()
        except Exception:  # noqa: BLE001
            pass

    e2e = None
    if not args.no_e2e:
        if world == 1:
            holder = [maps_full]                            # hand the device maps over: run_e2e must be able to free them
            maps_full = None
            hostmaps = run_e2e(args, rank, world, dev, comm, (main, bnd), pb_full, holder, frames_resident, barrier,
                               max_over_ranks)
            free_device()
            free_host()
            e2e = run_e2e_medium(args, dev)
            e2e["hostmaps_variant"] = hostmaps
        else:
            maps_full = None
            free_device()
            e2e = run_e2e_medium_slabs(args, rank, world, dev, comm, (main, bnd), barrier, max_over_ranks)
            free_device()
            free_host()
            e2e["hostmaps_variant"] = run_e2e(args, rank, world, dev, comm, (main, bnd), pb_full, [None],
                                              frames_resident, barrier, max_over_ranks)
        free_host()

    if rank == 0:
        config["same_grid"] = same
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": config,
            "roofline": roof, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        }
        if halo:
            line["halo"] = halo
        if strong:
            line["strong"] = strong
        if parity:
            line["parity_n_vs_1"] = parity
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline()
            line["host_setup_baseline"] = host_setup_baseline()
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def traffic_from_profile(kernel: str, grid) -> dict:
    """roofline.traffic: dram__bytes_read.sum + dram__bytes_write.sum of the committed `ncu --set full` capture, per
    UPDATED point (the 8-cell rim is never touched), scaled to the updated points of this launch."""
    for name in ("traffic_r02.json", "traffic_r01.json"):
        prof = ROOT / "profiles" / name
        if not prof.exists():
            continue
        try:
            tr = json.loads(prof.read_text())
            upd = (grid[0] - 16) * (grid[1] - 16) * (grid[2] - 16)
            per = tr.get(kernel, {}).get("dram_bytes_per_updated_point")
            if per is None:
                continue
            return {"traffic": per * upd, "traffic_per_updated_point": per, "updated_points_per_launch": upd,
                    "algorithmic_bytes_per_updated_point": BYTES_FD_U if kernel == "fd_u" else BYTES_FD_P,
                    "traffic_source": f"profiles/{name}: {tr.get('source')}"}
        except Exception:  # noqa: BLE001
            pass
    return {}


def run_e2e_medium(args, dev, max_mismatch_report: int = 4):
    """N = 1: the whole job through the plugin call a `fullwave.Solver.run` replacement makes -- `fw25_run_medium`
    (include/fw25.h) -- from the USER-grid medium in pinned HOST memory: sound_speed, density, beta, alpha_coeff,
    alpha_power (float32, what a `fullwave.Medium` holds on the user grid) + the relaxation look-up table.  Inside the
    timed region: device allocation, the host->device copies of the medium (block by block), map generation, all K
    steps (the first ones time-skewed under the upload), the sensor frames back on the host, and freeing the device."""
    import torch
    from fullwave25_b200 import engine, lut_standin, mapgen, synthetic_device
    from fullwave25_b200.problem import MAP_NAMES, Problem
    nX, nY, nZ = args.grid
    K = args.steps
    nb = 8 + MEDIUM["n_pml"] + MEDIUM["n_trans"]
    user = (nX - 2 * nb, nY - 2 * nb, nZ - 2 * nb)
    f0, c0, ppw, cfl = 1e6, 1540.0, 12, 0.2
    dx = c0 / f0 / ppw
    dt = cfl * dx / c0
    t_gen = time.perf_counter()
    um, c_min, c_max, pinned = synthetic_device.make_user_medium(user, device=dev, block=MEDIUM["block"], seed=MEDIUM["seed"])
    t_gen = time.perf_counter() - t_gen
    spec = mapgen.MediumSpec(user_shape=user, dt=dt, dx=dx, c0=c0, cfl=cfl, sound_speed=um["sound_speed"],
                             density=um["density"], beta=um["beta"], alpha_coeff=um["alpha_coeff"],
                             alpha_power=um["alpha_power"], lut=lut_standin.lookup_table(),
                             n_pml_layer=MEDIUM["n_pml"], n_transition_layer=MEDIUM["n_trans"], dcmap_full3d=True,
                             extra={"c_min": c_min, "c_max": c_max})
    _, dmap, ndmap, _ = spec.stencil_tables()
    nTic = min(K, int(np.ceil(2.0 / f0 / dt)) + 1)
    icc, icmat, outc, icczero = synthetic_device._lists(nX, nY, nZ, nb, nTic, dt, dx, f0=f0, c0=c0, seed=MEDIUM["seed"],
                                                        n_sensors=MEDIUM["n_sensors"], n_air=MEDIUM["n_air"],
                                                        source_layers=3, amp=1e5)
    icm = torch.empty(icmat.shape, dtype=torch.float32, pin_memory=True)        # the source signals are inputs too
    icm.numpy()[...] = icmat
    none = {name: None for name in MAP_NAMES}
    pb = Problem(ndim=3, nX=nX, nY=nY, nZ=nZ, nT=K, nTic=nTic, modT=MEDIUM["modT"], ndmap=ndmap,
                 dX=float(np.float32(dx)), dT=float(np.float32(dt)), **none, dmap=dmap, dcmap=None, icc=icc,
                 icmat=icm.numpy(), outc=outc, icczero=icczero, extra={}, dcmap_full3d=True)
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    t0 = time.perf_counter()
    out, st = mapgen.run_medium(spec, pb, device=dev.index or 0)
    dt_e2e = time.perf_counter() - t0
    pts = nX * nY * nZ
    e2e = {"value": pts * K / dt_e2e / 1e9, "unit": UNIT, "h2d_bytes_per_step": st["h2d_bytes"] / K,
           "d2h_bytes_per_step": st["d2h_bytes"] / K, "seconds": dt_e2e,
           "api": "fw25_run_medium (C-ABI, include/fw25.h) via fullwave25_b200.mapgen.run_medium: user-grid medium in "
                  "pinned host memory -> sensor frames on the host",
           "grid_per_gpu": f"{nX}x{nY}x{nZ}", "user_grid": "x".join(map(str, user)), "input_dtype": "float32",
           "launches": int(st["kernel_launches"]), "skewed_steps": int(st["skewed_steps"]),
           "engine_setup_ms": st["setup_ms"], "device_loop_ms": st["loop_ms"], "frames_d2h_ms": st["d2h_ms"],
           "host_marshal_ms": st["host_marshal_ms"], "native_call_ms": st["native_call_ms"],
           "medium_generation_s_untimed": round(t_gen, 2), "finite": bool(np.isfinite(out).all()),
           "absmax": float(np.abs(out).max()) if out.size else None,
           "relaxation_table": "stand-in (fullwave25_b200/lut_standin.py): the reference's database blob is missing "
                               "from its checkout",
           "note": "min / max of the sound speed (the stencil-table range) are properties of the synthetic medium and "
                   "are passed in; everything else a Solver.run replacement does is inside the timed region"}
    # parity at full size: the sequential path (one-shot fw25_mapgen, then whole-grid steps) on the same inputs
    if not args.no_e2e_check:
        with mapgen.MapSet(spec, device=dev.index or 0) as ms:
            eng = engine.Engine(pb, device=dev.index or 0, device_maps=ms.device_maps())
            try:
                want, _ = eng.run()
            finally:
                eng.close()
        e2e["frames_identical_to_sequential_run"] = bool(np.array_equal(out, want))
        e2e["nonzero_frame_values"] = int((want != 0).sum())
    del pinned
    return e2e


def run_e2e_medium_slabs(args, rank, world, dev, comm, streams, barrier, max_over_ranks):
    """N > 1: the same job from the USER-grid medium, one process per GPU.  Every rank holds, in pinned host memory, the
    user-grid planes its x-slab reads (sound_speed, density, beta, alpha_coeff, alpha_power, float32); inside the timed
    region it uploads them, builds its own slab of the 14 engine maps on its GPU (`fw25_mapgen_slab`: ghost planes
    included, no 14-map upload), creates the engine on those maps in place, runs the K steps with NCCL halos and the
    sensor frames are gathered on rank 0."""
    import torch
    import torch.distributed as dist
    from fullwave25_b200 import lut_standin, mapgen, synthetic_device
    from fullwave25_b200.problem import MAP_NAMES, Problem
    from fullwave25_b200.runtime import SlabEngine, gather_frames
    from fullwave25_b200.slab import SlabDriver, partition
    main, bnd = streams
    nXl, nY, nZ = args.grid
    K = args.steps
    gX = nXl * world
    nb = 8 + MEDIUM["n_pml"] + MEDIUM["n_trans"]
    user = (gX - 2 * nb, nY - 2 * nb, nZ - 2 * nb)
    slab = partition(gX, world)[rank]
    u0 = min(max(slab.gx0 - nb, 0), user[0] - 1)
    u1 = min(max(slab.gx1 - 1 - nb, 0), user[0] - 1) + 1
    f0, c0, ppw, cfl = 1e6, 1540.0, 12, 0.2
    dx = c0 / f0 / ppw
    dt = cfl * dx / c0
    import psutil
    need = 5 * (u1 - u0) * user[1] * user[2] * 4
    ok = torch.tensor([1.0 if need * 1.5 < psutil.virtual_memory().available / world else 0.0], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if ok.item() == 0.0:
        return {"unavailable": f"host RAM: {need / 1e9:.0f} GB of pinned user-grid planes per rank"}
    um, c_min, c_max, pinned = synthetic_device.make_user_medium(user, device=dev, block=MEDIUM["block"],
                                                                 seed=MEDIUM["seed"], x_range=(u0, u1))
    lim = torch.tensor([-c_min, c_max], device=dev, dtype=torch.float64)     # the stencil table spans the WHOLE medium
    dist.all_reduce(lim, op=dist.ReduceOp.MAX)
    c_min, c_max = -float(lim[0].item()), float(lim[1].item())
    spec = mapgen.MediumSpec(user_shape=user, dt=dt, dx=dx, c0=c0, cfl=cfl, sound_speed=um["sound_speed"],
                             density=um["density"], beta=um["beta"], alpha_coeff=um["alpha_coeff"],
                             alpha_power=um["alpha_power"], lut=lut_standin.lookup_table(),
                             n_pml_layer=MEDIUM["n_pml"], n_transition_layer=MEDIUM["n_trans"], dcmap_full3d=True,
                             extra={"c_min": c_min, "c_max": c_max}, user_planes=(u0, u1 - u0))
    _, dmap, ndmap, _ = spec.stencil_tables()
    nTic = min(K, int(np.ceil(2.0 / f0 / dt)) + 1)
    icc, icmat, outc, icczero = synthetic_device._lists(gX, nY, nZ, nb, nTic, dt, dx, f0=f0, c0=c0, seed=MEDIUM["seed"],
                                                        n_sensors=MEDIUM["n_sensors"], n_air=MEDIUM["n_air"],
                                                        source_layers=3, amp=1e5)
    icm = torch.empty(icmat.shape, dtype=torch.float32, pin_memory=True)
    icm.numpy()[...] = icmat
    none = {name: None for name in MAP_NAMES}
    pb = Problem(ndim=3, nX=slab.n_local, nY=nY, nZ=nZ, nT=K, nTic=nTic, modT=MEDIUM["modT"], ndmap=ndmap,
                 dX=float(np.float32(dx)), dT=float(np.float32(dt)), **none, dmap=dmap, dcmap=None, icc=icc,
                 icmat=icm.numpy(), outc=outc, icczero=icczero, extra={}, dcmap_full3d=True)
    h2d = sum(t.numel() * 4 for t in pinned.values()) + icm.numel() * 4 + icc.nbytes
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    barrier()
    t0 = time.perf_counter()
    # the set's device pointers are final at once: the engine (state arrays, plans, source / sensor lists) is created
    # while the user-grid planes are still going up block by block and the maps are generated behind them
    ms = mapgen.MapSet(spec, device=dev.index or 0, planes=(slab.gx0, slab.gx1), background=True)
    eng = SlabEngine(pb, slab, dev, device_maps=ms.device_maps())
    ms.wait()
    eng.eng.sync()
    t_setup = time.perf_counter() - t0
    drv = SlabDriver(slab, eng, comm, pb.modT, streams=(main, bnd), ndim=3)
    for _ in range(K):
        drv.step()
    out = gather_frames(drv, eng, pb.n_frames, pb.ncoordsout, dist)
    barrier()
    dt_e2e = max_over_ranks(time.perf_counter() - t0)
    e2e = {"value": gX * nY * nZ * K / dt_e2e / 1e9, "unit": UNIT, "h2d_bytes_per_step": h2d / K,
           "d2h_bytes_per_step": pb.n_frames * eng.eng.n_local_sensors * 4 / K, "seconds": dt_e2e,
           "setup_seconds": t_setup, "mapgen_upload_ms": ms.upload_ms, "mapgen_kernel_ms": ms.kernel_ms,
           "api": "per rank: fw25_mapgen_slab_begin / _finish (C-ABI; this rank's user-grid planes in pinned host memory -> "
                  "its slab of the engine maps in HBM, block by block in the background) + fw25_create on those maps "
                  "meanwhile + SlabDriver.step (NCCL halos) + gather_frames",
           "grid_per_gpu": f"{nXl}x{nY}x{nZ}", "user_grid": "x".join(map(str, user)),
           "user_planes_rank0": [int(u0), int(u1)], "input_dtype": "float32", "launches": int(eng.eng.launches),
           "finite": bool(np.isfinite(out).all()) if out is not None else None,
           "absmax": float(np.abs(out).max()) if out is not None and out.size else None,
           "nonzero_frame_values": int((out != 0).sum()) if out is not None else None,
           "relaxation_table": "stand-in (fullwave25_b200/lut_standin.py)",
           "parity": "slab maps == whole-grid maps byte for byte (tests/test_mapgen_gpu.py); N ranks == 1 rank "
                     "(parity_n_vs_1, tests/test_multi_gpu.py)"}
    eng.close()
    ms.close()
    del pinned
    return e2e


def run_e2e(args, rank, world, dev, comm, streams, pb, maps_holder, frames_chk, barrier, max_over_ranks):
    """The same job through the public API with HOST buffers: pinned host maps -> upload -> K steps -> frames."""
    maps = maps_holder.pop()
    import dataclasses

    import psutil
    import torch
    import torch.distributed as dist
    from fullwave25_b200 import synthetic_device
    from fullwave25_b200.runtime import SlabEngine, gather_frames
    from fullwave25_b200.slab import SlabDriver, partition
    D = dist if world > 1 else None
    main, bnd = streams
    nXl, nY, nZ = args.grid
    K = args.steps
    gshape = (nXl * world, nY, nZ)
    slab = partition(gshape[0], world)[rank]
    if maps is None:
        pb, maps = synthetic_device.make_slab(gshape, slab.gx0, slab.gx1, device=dev, nT=K, **MEDIUM)
    need = 14 * slab.n_local * nY * nZ * 4
    avail = psutil.virtual_memory().available / max(world, 1)
    fits = need * 1.25 <= avail
    if world > 1:                                       # one decision for all ranks (the ranks meet at barriers below)
        ok = torch.tensor([1.0 if fits else 0.0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        fits = ok.item() != 0.0
    if not fits:
        return {"unavailable": f"host RAM: {need / 1e9:.0f} GB of pinned maps per rank, {avail / 1e9:.0f} GB available"}
    pb = dataclasses.replace(pb, nT=K, nTic=min(pb.nTic, K))
    pb.icmat = np.ascontiguousarray(pb.icmat[:, : pb.nTic])
    host = synthetic_device.maps_to_host(maps, nZ, pin=True)
    del maps
    torch.cuda.empty_cache()
    pb_h = dataclasses.replace(pb, **{k: v.numpy() for k, v in host.items()})
    icm = torch.empty(pb_h.icmat.shape, dtype=torch.float32, pin_memory=True)   # the source signals are inputs too
    icm.numpy()[...] = pb_h.icmat
    pb_h.icmat = icm.numpy()
    h2d = sum(v.numel() * 4 for v in host.values()) + pb_h.icmat.nbytes + pb_h.icc.nbytes
    barrier()
    t0 = time.perf_counter()
    eng2 = SlabEngine(pb_h, slab, dev)                  # H2D of the 14 maps + coordinate lists happens here
    eng2.eng.sync()
    t_setup = time.perf_counter() - t0                  # allocation + upload (part of the timed region)
    drv2 = SlabDriver(slab, eng2, comm, pb_h.modT, streams=(main, bnd))
    for _ in range(K):
        drv2.step()
    out = gather_frames(drv2, eng2, pb_h.n_frames, pb_h.ncoordsout, D)
    barrier()
    dt = max_over_ranks(time.perf_counter() - t0)
    d2h = pb_h.n_frames * eng2.eng.n_local_sensors * 4
    e2e = {"value": gshape[0] * nY * nZ * K / dt / 1e9, "unit": UNIT,
           "h2d_bytes_per_step": h2d / K, "d2h_bytes_per_step": d2h / K, "seconds": dt,
           "setup_seconds": t_setup, "upload_GBps": h2d / t_setup / 1e9,
           "grid_per_gpu": f"{nXl}x{nY}x{nZ}", "launches": eng2.eng.launches,
           "finite": bool(np.isfinite(out).all()) if out is not None else None,
           "frames_identical_to_resident_run": (bool(np.array_equal(out, frames_chk[: out.shape[0]]))
                                                if out is not None and frames_chk is not None else None),
           "absmax": float(np.abs(out).max()) if out is not None and out.size else None,
           "api": "fullwave25_b200.runtime.SlabEngine(host maps) + SlabDriver.step + gather_frames"}
    eng2.close()
    return e2e


# ------------------------------------------------------------------------------------------- reference
def run_reference(args):
    """The reference's shipped sm_100 executable through the reference's own launcher, timed per step from its
    progress output (tools/bench_reference.py)."""
    from tools import bench_reference
    bench_reference.run_reference(args, metric=METRIC, unit=UNIT, workload=WORKLOAD, medium=MEDIUM)


def run_examples_config(args):
    """--config examples: BASELINE.json configs 1-4, the reference's four shipped example set-ups run verbatim through
    `fullwave.Solver.run` on the reference binary and on this engine (tools/run_examples.py).  One JSON line."""
    from tools import run_examples
    names = args.examples.split(",") if args.examples else list(run_examples.EXAMPLES)
    res = run_examples.run_all(names, tuple(args.example_engines.split(",")), args.duration_scale, log=lambda s: None)
    by = {}
    for r in res:
        by.setdefault(r["example"], {})[r["engine"]] = {k: v for k, v in r.items() if k not in ("example", "engine")}
    table = {}
    for name, e in by.items():
        ref, dev, host = e.get("reference", {}), e.get("fw25-device", {}), e.get("fw25-host", {})
        table[name] = {
            "extended_grid": ref.get("extended_grid") or dev.get("extended_grid"), "steps": ref.get("steps") or dev.get("steps"),
            "sensors": ref.get("sensors") or dev.get("sensors"), "frames": ref.get("frames") or dev.get("frames"),
            "engine_Gpts": {k: v.get("engine_Gpts") for k, v in e.items()},
            "roofline_frac": {k: v.get("roofline_frac") for k, v in e.items() if k != "reference"},
            "solver_run_s": {k: v.get("solver_run_s") for k, v in e.items()},
            "solver_init_s": {k: v.get("solver_init_s") for k, v in e.items()},
            "vs_reference": {k: v.get("vs_reference") for k, v in e.items() if k != "reference"},
            "errors": {k: v["error"] for k, v in e.items() if "error" in v} or None,
        }
    vals = [t["engine_Gpts"].get("fw25-device") or t["engine_Gpts"].get("fw25-host") for t in table.values()]
    vals = [v for v in vals if v]
    print(json.dumps({
        "metric": "FDTD Gpoint-updates/s (engine time loop), BASELINE.json configs 1-4", "unit": UNIT,
        "value": float(np.exp(np.mean(np.log(vals)))) if vals else None, "value_is": "geometric mean over the examples",
        "n_gpus": 1, "higher_is_better": True, "dtype": "f32", "data": "the reference's example scripts, verbatim",
        "config": {"workload": "examples/{simple_plane_wave, linear_transducer, convex_transducer, wave_3d} (BASELINE.json "
                               "configs[0..3])", "duration_scale": args.duration_scale,
                   "relaxation_table": "stand-in (fullwave25_b200/lut_standin.py)", "harness": "tools/run_examples.py"},
        "examples": table}), flush=True)


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
    ap.add_argument("--config", default="synthetic", choices=["synthetic", "examples"],
                    help="synthetic: BASELINE.json configs[4] (default); examples: configs[0..3], the shipped example scripts")
    ap.add_argument("--examples", default="", help="comma-separated subset for --config examples")
    ap.add_argument("--example-engines", default="reference,fw25-host,fw25-device")
    ap.add_argument("--duration-scale", type=float, default=1.0, help="--config examples: scale every example's duration")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-e2e-check", action="store_true")
    ap.add_argument("--no-strong", action="store_true")
    ap.add_argument("--no-same-grid", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.config == "examples":
        run_examples_config(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
