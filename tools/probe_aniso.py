"""Informational (GPU box): rate of the anisotropic-relaxation sweeps (per-axis kappa / a / b maps, 288 algorithmic
bytes per point-update) on a device-generated medium, warp-specialised TMA kernels against the L1/L2-path kernels.
usage: probe_aniso.py XxYxZ steps"""
import json, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from fullwave25_b200 import synthetic_device
from fullwave25_b200.runtime import SlabEngine
from fullwave25_b200.slab import partition

BYTES_U, BYTES_P = 148, 140      # fd_u: 28 reads + 9 writes; fd_p: 28 reads + 7 writes (float32)
shape = tuple(int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "320x1240x1240").split("x"))
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = torch.device("cuda", 0)
slab = partition(shape[0], 1)[0]
pb, maps = synthetic_device.make_slab(shape, 0, shape[0], device=dev, nT=10000, n_pml=36, n_trans=36, block=24)
# per-axis members: axis x = the isotropic maps; y, z differ from them everywhere (values stay physical)
an, keep = {}, []
for fam, letters in (("x", "xyz"), ("u", "uvw")):
    for ax, l in enumerate(letters):
        for stem, scale in (("kappa", 1.0 + 0.003 * ax), ("apml%s1", 1.0 + 0.15 * ax), ("bpml%s1", 1.0 - 1e-4 * ax),
                            ("apml%s2", 1.0 - 0.1 * ax), ("bpml%s2", 1.0 - 2e-4 * ax)):
            src = maps[(stem % fam) if "%s" in stem else stem + fam]
            t = src if ax == 0 else src * scale
            keep.append(t)
            an[(stem % l) if "%s" in stem else stem + l] = t.data_ptr()
dm = {k: (v if k == "pitch" else v.data_ptr()) for k, v in maps.items()}
dm["aniso"] = an
n = shape[0] * shape[1] * shape[2]
out = {"shape": shape, "steps": steps, "algorithmic_bytes": {"fd_u": BYTES_U, "fd_p": BYTES_P}}
frames = {}
for name, variant in (("ws", 0), ("simple", 1)):
    eng = SlabEngine(pb, slab, dev, device_maps=dm, variant=variant)
    e = eng.eng
    e.step(5); e.sync()
    r = e.step_timed(steps, detail=True)
    r2 = e.step_timed(steps, detail=False)
    frames[name] = e.read_frames(0, 3).copy()
    u, p = r["sweep_u_ms"] / steps, r["sweep_p_ms"] / steps
    out[name] = {"gpts": n * steps / r2["total_ms"] / 1e6, "ms_per_step": r2["total_ms"] / steps, "u_ms": u, "p_ms": p,
                 "u_GBps": n * BYTES_U / u / 1e6, "p_GBps": n * BYTES_P / p / 1e6,
                 "step_GBps": n * (BYTES_U + BYTES_P) * steps / r2["total_ms"] / 1e6,
                 "frac_of_6546": n * (BYTES_U + BYTES_P) * steps / r2["total_ms"] / 1e6 / 6546.2}
    del e, eng
    torch.cuda.empty_cache()
out["frames_identical"] = bool((frames["ws"] == frames["simple"]).all())
out["frames_absmax"] = float(abs(frames["ws"]).max())
print(json.dumps(out))
