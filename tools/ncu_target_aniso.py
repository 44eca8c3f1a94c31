"""Short run of the anisotropic warp-specialised sweeps on a device-generated medium, for ncu captures (profiles/)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from fullwave25_b200 import synthetic_device
from fullwave25_b200.runtime import SlabEngine
from fullwave25_b200.slab import partition

shape = tuple(int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "192x1240x1240").split("x"))
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
slab = partition(shape[0], 1)[0]
pb, maps = synthetic_device.make_slab(shape, 0, shape[0], device=dev, nT=steps, n_pml=36, n_trans=36, block=24)
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
eng = SlabEngine(pb, slab, dev, device_maps=dm)
eng.eng.step(steps)
eng.eng.sync()
print("done", eng.eng.launches, "launches")
