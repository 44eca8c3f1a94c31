"""bench.py's end-to-end arm alone (fw25_run_medium from the user-grid medium), with the setup trace on.

    gpurun -- python tools/probe_e2e.py [--grid 800x1240x1240] [--steps 20] [--repeat 2]
"""
import argparse
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
os.environ.setdefault("FW25_SETUP_TRACE", "1")

import torch  # noqa: E402

import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--grid", type=bench.parse_grid, default=(800, 1240, 1240))
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--repeat", type=int, default=2)
ap.add_argument("--no-e2e-check", action="store_true")
args = ap.parse_args()
torch.cuda.set_device(0)
for i in range(args.repeat):
    r = bench.run_e2e_medium(args, torch.device("cuda", 0))
    r.pop("note", None), r.pop("relaxation_table", None), r.pop("api", None)
    print(json.dumps(r), flush=True)
