"""ncu target for the setup / recording kernels: k_mapgen on the wave_3d-sized grid (120^3 user -> 280^3 extended) and
k_record_box on a whole-user-grid box of the linear-transducer-sized 2D grid (profiles/)."""
import os
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
os.environ.setdefault("FW25_GRAPH", "0")          # ncu profiles kernel launches, not graph replays
from fullwave25_b200 import engine, mapgen, synthetic  # noqa: E402
from tests import mapgen_cases as mc  # noqa: E402

case = dict(shape=(120, 120, 120), seed=1, n_pml=36, n_trans=36)
m = mc.medium_arrays(case)
dx = mc.C0 / mc.F0 / mc.PPW
spec = mapgen.MediumSpec(user_shape=case["shape"], dt=mc.CFL * dx / mc.C0, dx=dx, c0=mc.C0, cfl=mc.CFL,
                         sound_speed=m["sound_speed"], density=m["density"], beta=m["beta"], relax=m["relax"],
                         n_pml_layer=36, n_transition_layer=36, dcmap_full3d=True)
for _ in range(2):
    with mapgen.MapSet(spec) as ms:
        print("mapgen", ms.shape, f"{ms.kernel_ms:.3f} ms")

pb = synthetic.make_problem((628, 628), nT=8, modT=2, n_sensors=16, n_air=16, seed=1234, n_pml=36, n_trans=36)
pb.out_box = (80, 80, 548, 548)                   # the user grid of examples/linear_transducer: 468 x 468 sensors
with engine.Engine(pb) as e:
    e.step(8)
    e.sync()
    print("box run done", e.launches, "launches", e.n_local_sensors, "sensors")
