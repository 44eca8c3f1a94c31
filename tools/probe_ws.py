"""Informational (GPU box): rate of the default sweeps on a device-generated medium.  usage: probe_ws.py XxYxZ steps [variant]"""
import json, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from fullwave25_b200 import synthetic_device
from fullwave25_b200.runtime import SlabEngine
from fullwave25_b200.slab import partition
shape = tuple(int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "320x632x632").split("x"))
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
variant = int(sys.argv[3]) if len(sys.argv) > 3 else 0
dev = torch.device("cuda", 0)
slab = partition(shape[0], 1)[0]
pb, maps = synthetic_device.make_slab(shape, 0, shape[0], device=dev, nT=10000, n_pml=36, n_trans=36, block=24)
dm = {k: (v if k == "pitch" else v.data_ptr()) for k, v in maps.items()}
eng = SlabEngine(pb, slab, dev, device_maps=dm, variant=variant)
e = eng.eng
e.step(5); e.sync()
r = e.step_timed(steps, detail=True)
r2 = e.step_timed(steps, detail=False)
n = shape[0] * shape[1] * shape[2]
print(json.dumps({"shape": shape, "gpts": n * steps / r2["total_ms"] / 1e6, "frac_6546": n * steps * 208 / r2["total_ms"] / 1e6 / 6546.2,
                  "u_ms": r["sweep_u_ms"] / steps, "p_ms": r["sweep_p_ms"] / steps,
                  "u_gbps": n * 108 / (r["sweep_u_ms"] / steps) / 1e6, "p_gbps": n * 100 / (r["sweep_p_ms"] / steps) / 1e6}))
