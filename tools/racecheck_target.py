"""compute-sanitizer target: the fused 2D step with sensors on source and air cells (shared-memory reuse of the u tile
for the tile's p', per-tile lists).    gpurun -- compute-sanitizer --tool racecheck python tools/racecheck_target.py"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import os, copy
os.environ["FW25_FUSE2D"] = "1"
import numpy as np
from fullwave25_b200 import engine
from tests import cases
from oracle import oracle
pb = cases.make("het2d"); pb.nT = 48
pb.outc = np.vstack([pb.outc, pb.icc[:5], pb.icczero[:3]]).astype(np.int32)
got, st = engine.run(pb)
assert np.array_equal(got, oracle.run(pb)); print("fused listed ok", st["kernel_launches"])
