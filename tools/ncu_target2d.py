"""Short 2D run (convex-transducer-sized extended grid by default) for ncu captures of the 2D sweeps (profiles/)."""
import os
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
os.environ.setdefault("FW25_GRAPH", "0")          # ncu profiles kernel launches, not graph replays
from fullwave25_b200 import engine, synthetic

shape = tuple(int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "1457x2178").split("x"))
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
pb = synthetic.make_problem(shape, nT=steps, modT=2, n_sensors=512, n_air=16, seed=1234, n_pml=36, n_trans=36)
with engine.Engine(pb) as e:
    e.step(steps)
    e.sync()
    print("done", e.launches, "launches")
