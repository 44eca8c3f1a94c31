"""bench.py's reference arm: the reference's OWN shipped sm_100 CUDA engine, through the reference's own launcher
(`fullwave.solver.launcher.Launcher.run`, /root/reference/fullwave/solver/launcher.py:160-254, from baseline/_ref),
timed PER STEP from its progress output.

Why per step: the engine prints "\\b\\b\\b\\b\\b%0.3f" (10 bytes) + fflush at the TOP of every iteration of its time loop
(ASM 0x407c88-0x407cb7) and the launcher sends stdout to `fw2_execution.log`.  The size of that
file therefore counts the steps started; a thread here polls it and timestamps every increment.  Whole-process wall
times (CUDA context, 35 freads of multi-GB files, 34 allocations) are minutes next to a 20-step timed region, which is
why round 1's run-differencing produced noise; they are reported beside the per-step figure as a cross-check only.

Host/device ordering: the engine synchronises with the GPU after every step that records a frame (`t % modT == 0`:
D2H copy + cudaStreamSynchronize) -- and, on several GPUs, after every sweep.  So the print of
step t is on the device's clock whenever (t - 1) % modT == 0; the timed region starts and ends on such prints.

Nothing of this repo's engine is on this path: inputs are generated with torch (fullwave25_b200.synthetic_device, the
same generator and seeds as our arm) and written as the .dat directory the reference's InputFileWriter would produce.
"""

from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys
import tempfile
import threading
import time
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

PROGRESS_TAG = b"Progress :      "
BYTES_PER_STEP = 10                      # "\b\b\b\b\b" + "%0.3f" of a value in [0, 1]
REF_BYTES_PER_POINT_DEVICE = 184         # two time levels of 16 state arrays + 14 maps
PLANES = (560, 480, 400, 320, 240, 200, 160, 120, 96)   # per-GPU candidates, BASELINE.md 2.2 row "5-ref" first


class ProgressTail(threading.Thread):
    """Timestamps the growth of fw2_execution.log: stamps[k] = perf_counter() when the print of step k appeared."""

    def __init__(self, path: Path, period_s: float = 2e-4):
        super().__init__(daemon=True)
        self.path, self.period = path, period_s
        self.stamps: list[float] = []
        self.header = None
        self._halt = threading.Event()

    def run(self):
        path = str(self.path)
        while not self._halt.is_set():
            try:
                size = os.stat(path).st_size
            except OSError:
                size = 0
            now = time.perf_counter()
            if self.header is None and size:
                try:
                    with open(path, "rb") as f:
                        head = f.read(1 << 16)
                    i = head.find(PROGRESS_TAG)
                    if i >= 0:
                        self.header = i + len(PROGRESS_TAG)
                except OSError:
                    pass
            if self.header is not None:
                n = (size - self.header) // BYTES_PER_STEP
                while len(self.stamps) < n:
                    self.stamps.append(now)
            time.sleep(self.period)

    def stop(self):
        self._halt.set()
        self.join(timeout=5)


def step_times(stamps, W: int, K: int, modT: int):
    """(t0, t1, seconds): the K-step window starting at the first device-synchronised print >= W."""
    t0 = W
    while (t0 - 1) % modT != 0:
        t0 += 1
    t1 = t0 + K
    while (t1 - 1) % modT != 0:
        t1 += 1
    if t1 >= len(stamps):
        return None
    return t0, t1, stamps[t1] - stamps[t0]


def ref_nT(W: int, K: int, modT: int) -> int:
    """Steps the reference arm runs (and our arm's same-grid run, so that the frames can be compared by hash)."""
    return W + K + 2 * modT + 1


def choose_planes(world: int, nY: int, nZ: int, forced: int | None = None):
    """Largest per-GPU slab the reference can run here: int32 element counts in its host code (global points < 2^31),
    184 B/point on the device, and host RAM for the .dat files in tmpfs plus the engine's own host copies."""
    import psutil
    import torch
    if forced:
        return forced, {"forced": True}
    plane = nY * nZ
    ram = psutil.virtual_memory().total         # totals, not what is free right now: both arms must pick the same grid
    shm = shutil.disk_usage("/dev/shm").total if Path("/dev/shm").exists() else 0
    gpu = torch.cuda.get_device_properties(0).total_memory
    why = {"host_ram_GB": round(ram / 1e9, 1), "dev_shm_GB": round(shm / 1e9, 1),
           "gpu_GB": round(gpu / 1e9, 1)}
    for p in PLANES:
        pts = p * world * plane
        files = 14 * pts * 4                                  # 13 maps + dcmap; c.dat is a hole
        if pts >= 2**31 - 2**24:
            continue
        if files * 1.05 > shm or (files + 16 * pts * 4) * 1.1 > ram:
            continue
        if (p + 16) * plane * REF_BYTES_PER_POINT_DEVICE > 0.97 * gpu - (1 << 30):
            continue
        return p, why
    return None, why


def write_inputs(work: Path, gshape, nT: int, medium: dict, device, chunk: int = 40):
    """The reference's .dat directory for the synthetic workload, generated slab by slab with torch (same generator,
    same seeds as our arm) and appended to the 14 map files by a pool of writer threads."""
    import torch
    from fullwave25_b200 import synthetic_device
    from fullwave25_b200.problem import MAP_NAMES
    nX, nY, nZ = gshape
    work.mkdir(parents=True, exist_ok=True)
    names = MAP_NAMES + ("dcmap",)
    fds = {n: os.open(work / f"{n}.dat", os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o644) for n in names}
    plane_bytes = nY * nZ * 4

    def put(fd, arr, offset):
        """Positional write: several chunks of one file may be in flight at once, in any order."""
        buf = memoryview(arr).cast("B")
        done = 0
        while done < len(buf):
            done += os.pwrite(fd, buf[done: done + (1 << 30)], offset + done)
    pool = ThreadPoolExecutor(max_workers=len(names))
    pb0 = None
    stage = [{n: torch.empty((chunk, nY, nZ), dtype=torch.int32 if n == "dcmap" else torch.float32, pin_memory=True)
              for n in names} for _ in range(2)]
    futs: list[list] = []                                     # writer futures per chunk
    for k, x0 in enumerate(range(0, nX, chunk)):
        x1 = min(x0 + chunk, nX)
        pb, maps = synthetic_device.make_slab(gshape, x0, x1, device=device, nT=nT, with_lists=(x0 == 0), **medium)
        if x0 == 0:
            pb0 = pb
        if k >= 2:                                            # the staging set used two chunks ago must be on disk
            for f in futs[k - 2]:
                f.result()
        st = stage[k & 1]
        for n in names:
            st[n][: x1 - x0].copy_(maps[n][..., :nZ], non_blocking=True)
        torch.cuda.synchronize()
        del maps
        futs.append([pool.submit(put, fds[n], st[n].numpy()[: x1 - x0], x0 * plane_bytes) for n in names])
    for fs in futs:
        for f in fs:
            f.result()
    pool.shutdown()
    for fd in fds.values():
        os.close(fd)
    del stage
    torch.cuda.empty_cache()
    pb = pb0
    pts = nX * nY * nZ
    with open(work / "c.dat", "wb") as f:                     # read by the engine, used by no kernel: a hole in tmpfs
        f.truncate(pts * 4)
    np.zeros((9, 2), np.float32).tofile(work / "d.dat")
    pb.dmap.astype(np.float32).tofile(work / "dmap.dat")
    pb.icc.astype(np.int32).tofile(work / "icc.dat")
    pb.outc.astype(np.int32).tofile(work / "outc.dat")
    pb.icczero.astype(np.int32).tofile(work / "icczero.dat")
    pb.icmat.astype(np.float32).tofile(work / "icmat.dat")
    ints = {"nX": nX, "nY": nY, "nZ": nZ, "nT": nT, "ncoords": pb.ncoords, "ncoordsout": pb.ncoordsout,
            "ncoordszero": pb.ncoordszero, "nTic": pb.nTic, "modT": pb.modT, "ndmap": pb.ndmap}
    floats = {"dX": pb.dX, "dY": pb.dX, "dZ": pb.dX, "dT": pb.dT, "c0": pb.extra.get("c0", 1540.0)}
    for k, v in ints.items():
        np.array(v).astype(np.int32).tofile(work / f"{k}.dat")
    for k, v in floats.items():
        np.array(v).astype(np.float32).tofile(work / f"{k}.dat")
    return pb


def run_reference(args, *, metric: str, unit: str, workload: str, medium: dict) -> None:
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    world = args.gpus
    K, W = args.steps, args.warmup
    t_arm = time.perf_counter()

    def unavailable(why: str):
        print(json.dumps({"impl": "reference", "unavailable": why[:400]}), flush=True)

    try:
        import torch
        from tools.ref_import import import_fullwave
        import_fullwave(ROOT / "baseline" / "_ref")
        from fullwave.solver.launcher import Launcher
        from tools.make_ref_golden import REF_BIN
    except Exception as e:  # noqa: BLE001
        return unavailable(f"{type(e).__name__}: {e}")
    if not REF_BIN[3].exists():
        return unavailable(f"reference engine binary not found: {REF_BIN[3]}")
    if not torch.cuda.is_available():
        return unavailable("no CUDA device: the reference has no CPU engine (solver.py:240-272)")

    _, nY, nZ = args.grid
    planes, why = choose_planes(world, nY, nZ, args.ref_planes)
    if planes is None:
        return unavailable(f"no grid of the workload fits this box for the reference engine: {why}")
    modT = medium["modT"]
    nT = ref_nT(W, K, modT)
    dev = torch.device("cuda", 0)
    work = Path("/dev/shm" if Path("/dev/shm").exists() else tempfile.gettempdir()) / "fw25_bench_ref"
    tried = []
    result = None
    while planes is not None:
        gshape = (planes * world, nY, nZ)
        try:
            if work.exists():
                shutil.rmtree(work)
            t0 = time.perf_counter()
            pb = write_inputs(work, gshape, nT, medium, dev)
            t_write = time.perf_counter() - t0
            exe = work / REF_BIN[3].name
            shutil.copy(REF_BIN[3], exe)
            exe.chmod(0o755)
            la = Launcher(exe, is_3d=True, use_gpu=True, cuda_device_id=list(range(world)) if world > 1 else 0)
            tail = ProgressTail(work / "fw2_execution.log")
            tail.start()
            t0 = time.perf_counter()
            try:
                genout = la.run(work, load_results=True)
            finally:
                wall = time.perf_counter() - t0
                tail.stop()
            result = (gshape, pb, genout, wall, t_write, tail)
            break
        except Exception as e:  # noqa: BLE001  (typically: the engine ran out of device or host memory)
            log = ""
            try:
                log = (work / "fw2_execution.log").read_text(errors="replace")[-300:]
            except OSError:
                pass
            tried.append({"planes_per_gpu": planes, "error": f"{type(e).__name__}: {str(e)[:120]}", "log_tail": log})
            smaller = [p for p in PLANES if p < planes]
            planes = smaller[0] if smaller and time.perf_counter() - t_arm < 240 else None
    if result is None:
        shutil.rmtree(work, ignore_errors=True)
        return unavailable(f"the reference engine failed on every grid tried: {tried}")

    gshape, pb, genout, wall, t_write, tail = result
    stamps = tail.stamps
    win = step_times(stamps, W, K, modT)
    if win is None or len(stamps) != nT:
        shutil.rmtree(work, ignore_errors=True)
        return unavailable(f"progress output not understood: {len(stamps)} step prints for nT = {nT} "
                           f"(header at {tail.header})")
    t0s, t1s, secs = win
    n_steps = t1s - t0s
    if secs <= 0:
        shutil.rmtree(work, ignore_errors=True)
        return unavailable(f"non-positive step window: {secs} s over steps {t0s}..{t1s}")
    pts = gshape[0] * nY * nZ
    value = pts * n_steps / secs / 1e9
    gaps = np.diff(np.asarray(stamps[t0s: t1s + 1]))
    per_period = gaps.reshape(-1, modT).sum(axis=1) / modT if n_steps % modT == 0 else gaps
    genout = np.asarray(genout, dtype=np.float32)
    sha = hashlib.sha256(genout.tobytes()).hexdigest()

    # cross-check by differencing whole-process wall times (round 1's method): one more, shorter run
    cross = None
    if not args.no_ref_crosscheck and time.perf_counter() - t_arm + wall < 300:
        try:
            nT2 = t0s + 1
            np.array(nT2).astype(np.int32).tofile(work / "nT.dat")
            (work / "genout.dat").unlink(missing_ok=True)
            t0 = time.perf_counter()
            la.run(work, load_results=True)
            wall2 = time.perf_counter() - t0
            d = wall - wall2
            cross = {"wall_s_nT": {str(nT): round(wall, 3), str(nT2): round(wall2, 3)},
                     "ms_per_step": (d / (nT - nT2) * 1e3) if d > 0 else None,
                     "note": "difference of two whole-process wall times incl. file I/O; informational"}
        except Exception as e:  # noqa: BLE001
            cross = {"error": f"{type(e).__name__}: {e}"[:200]}
    shutil.rmtree(work, ignore_errors=True)

    ms_step = secs * 1e3 / n_steps
    sample = (f"reference sm_100 binary via fullwave.solver.launcher.Launcher, {gshape[0]}x{nY}x{nZ} global "
              f"({gshape[0] // world} planes per GPU), steps {t0s}..{t1s} of nT={nT} timed from its per-step progress prints")
    line = {
        "impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload, "grid_per_gpu": f"{gshape[0] // world}x{nY}x{nZ}",
                   "global_grid": f"{gshape[0]}x{nY}x{nZ}", "points_per_step": pts,
                   "parallelism": f"reference in-process x-slabs x{world}",
                   "grid_note": ("largest slab of the workload the reference holds: int32 element counts (global points "
                                 "< 2^31), 184 B/point on the device, host RAM for its .dat files; our arm prints its rate "
                                 "on this same grid as config.same_grid"),
                   "grid_limits": why, "sensors": int(pb.ncoordsout), "sources": int(pb.ncoords),
                   "air_voxels": int(pb.ncoordszero), "dcmap": "reference 3D binary reads dcmap[:nX*nY] only"},
        "timing": {"method": "timestamps of the engine's per-step progress prints (10 bytes per step in fw2_execution.log), "
                             "window bounded by device-synchronised prints", "window_steps": [t0s, t1s],
                   "window_s": secs, "steps_in_window": n_steps,
                   "median_ms_per_step": float(np.median(per_period) * 1e3),
                   "min_ms_per_step": float(per_period.min() * 1e3), "max_ms_per_step": float(per_period.max() * 1e3),
                   "process_wall_s": wall, "input_generation_and_write_s": t_write, "crosscheck_differencing": cross,
                   "grids_that_failed": tried},
        "genout_sha256": sha, "genout_frames": int(genout.size // max(pb.ncoordsout, 1)),
        "cpu_baseline": {"value": value, "unit": unit, "kind": "reference", "cores": 1, "sample": sample,
                         "note": "the reference has no CPU engine (solver.py:240-272): this is its shipped CUDA "
                                 "engine, one host thread driving the GPU(s)"},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
