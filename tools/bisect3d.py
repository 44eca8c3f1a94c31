"""Informational (GPU box): which ingredient of a heterogeneous 3D problem makes the reference binary and
the oracle differ.  Prints rel-L2 per variant + per-frame onset."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from tests import cases
from oracle import oracle
from tools.make_ref_golden import run_reference, rel_l2

tmp = Path("/dev/shm/fw25_bisect")

def variant(name):
    pb = cases.make("het3d")
    if name == "base": pass
    elif name == "noair": pb.icczero = pb.icczero[:0]
    elif name == "dc0":
        pb.dcmap[:] = 0
    elif name == "dcconst":
        pb.dcmap[:] = int(pb.dcmap.max() // 2)
    elif name == "linear": pb.beta[:] = 0.5
    elif name == "tiny": pb.icmat *= 1e-6
    elif name == "kappa1": pb.kappax[:] = 1; pb.kappau[:] = 1
    elif name == "norelax":
        for k in ("apmlx1","apmlx2","apmlu1","apmlu2"): getattr(pb,k)[:] = 0
    elif name == "rhoK":  # homogeneous rho,K only
        pb.rho[:] = 1000; pb.K[:] = 1540.0**2*1000
    elif name == "onesrc":
        pb.icc = pb.icc[:1]; pb.icmat = pb.icmat[:1]
    return pb.normalise()

for name in sys.argv[1:] or ["base","base","noair","dc0","dcconst","linear","tiny","kappa1","norelax","rhoK","onesrc"]:
    pb = variant(name)
    g, dt, log = run_reference(pb, tmp / name)
    o = oracle.run(pb)
    per_frame = [rel_l2(o[f], g[f]) for f in range(g.shape[0])]
    onset = next((f for f, e in enumerate(per_frame) if e > 0), None)
    print(f"{name:8s} rel_l2={rel_l2(o,g):.3e} n_diff={(o!=g).sum()} onset_frame={onset} first_errs={[f'{e:.1e}' for e in per_frame[:12]]}", flush=True)
    if name == "base":
        np.save(ROOT/"gpurun_out"/"bisect_base_ref.npy", g)
