"""Run under torchrun on N GPUs: x-slab run of a seeded case through SlabEngine + NCCL halo exchange; rank 0
compares the assembled sensor frames with the single-domain oracle (bit-exact) and prints a JSON verdict.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/slab_check.py het3d
"""
import json
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch
import torch.distributed as dist

from fullwave25_b200.runtime import SlabEngine, TorchComm, gather_frames
from fullwave25_b200.slab import SlabDriver, partition
from tests.test_slab import _problem


def local_main(case, n, mode="native"):
    """Single process, n devices: engine.run(pb, device_ids=[0..n-1]) == the reference's cuda_device_id list.
    mode "native": fw25_run's own multi-device runner (C++); "torch": the Python lockstep driver (runtime.run_local).
    FW25_TEST_DEVICES="0,0" maps the n slabs onto the listed devices (several slabs on one GPU)."""
    from fullwave25_b200 import engine, runtime
    from oracle import oracle
    pb = _problem(case)
    ids = [int(v) for v in os.environ["FW25_TEST_DEVICES"].split(",")] if os.environ.get("FW25_TEST_DEVICES") else list(range(n))
    n = len(ids)
    if mode == "torch":
        got, stats = runtime.run_local(pb, ids, return_stats=True)
    else:
        got, stats = engine.run(pb, device_ids=tuple(ids))
    want = oracle.run(pb)
    print("SLABCHECK " + json.dumps({"case": case, "world": n, "mode": "in-process " + mode,
                                     "bit_exact": bool(np.array_equal(got, want)), "n_devices": int(stats.get("n_devices", 0)),
                                     "halo_bytes": int(stats.get("halo_bytes", 0)),
                                     "absmax": float(np.abs(want).max()), "n_diff": int((got != want).sum())}), flush=True)


def main():
    case = sys.argv[1] if len(sys.argv) > 1 else "het3d"
    if "RANK" not in os.environ:
        return local_main(case, int(sys.argv[2]) if len(sys.argv) > 2 else 2, sys.argv[3] if len(sys.argv) > 3 else "native")
    rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    pb = _problem(case)
    slab = partition(pb.nX, world)[rank]
    sub = pb.slab(slab.gx0, slab.gx1).normalise()
    eng = SlabEngine(sub, slab, dev)
    main_s, bnd = torch.cuda.Stream(dev), torch.cuda.Stream(dev, priority=-1)
    drv = SlabDriver(slab, eng, TorchComm(dist), pb.modT, streams=(main_s, bnd), ndim=pb.ndim)
    for _ in range(pb.nT):
        drv.step()
    got = gather_frames(drv, eng, pb.n_frames, pb.ncoordsout, dist)
    if rank == 0:
        from oracle import oracle
        want = oracle.run(pb)
        print("SLABCHECK " + json.dumps({"case": case, "world": world, "bit_exact": bool(np.array_equal(got, want)),
                                         "absmax": float(np.abs(want).max()), "n_diff": int((got != want).sum())}), flush=True)
    torch.cuda.synchronize()
    dist.barrier()                       # nobody tears NCCL down while a peer is still inside the gather
    eng.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
