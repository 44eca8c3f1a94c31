import sys, time, os
sys.path.insert(0, "/root/repo")
import numpy as np
from fullwave25_b200 import engine, synthetic
pb = synthetic.make_problem((628, 628), nT=2805, modT=2, n_sensors=64, n_air=16, seed=1, n_pml=36, n_trans=36)
m = np.zeros(pb.shape, bool); m[80:-80, 80:-80] = True
pb.outc = np.stack(np.nonzero(m), axis=1).astype(np.int32)
pb.normalise()
for pre in ("1", "0", "1", "0"):
    os.environ["FW25_PREFAULT"] = pre
    t0 = time.perf_counter(); g, st = engine.run(pb); t1 = time.perf_counter()
    print("prefault", pre, "wall %.3f s" % (t1 - t0), {k: round(v, 1) if isinstance(v, float) else v for k, v in st.items() if k in ("setup_ms", "loop_ms", "d2h_ms")}, flush=True)
for i in range(2):
    t0 = time.perf_counter(); e = engine.Engine(pb); e.sync(); t1 = time.perf_counter(); e.close(); t2 = time.perf_counter()
    print("create %.3f s destroy %.3f s" % (t1 - t0, t2 - t1))
