import sys, tempfile, time, subprocess, os
from pathlib import Path
sys.path.insert(0, "/root/repo")
import numpy as np
from fullwave25_b200 import build
from tools import ref_objects
shape = (468, 468)
fw, grid, medium, source, _ = ref_objects.build(shape, n_steps=2805, n_sensors=8, n_air=16, modT=2)
sensor = fw.Sensor(mask=np.ones(shape, dtype=bool), sampling_modulus_time=2)
with tempfile.TemporaryDirectory(dir="/dev/shm") as td:
    for name, b in (("ref", ref_objects.ref_bin(2)), ("ours", build.CLI)):
        s = fw.Solver(Path(td) / name, grid, medium, source, sensor, path_fullwave_simulation_bin=b)
        t0 = time.perf_counter(); r = s.run(); t1 = time.perf_counter()
        d = Path(td) / name / "txrx_0"
        # the child alone, again, in the prepared directory
        exe = d / Path(b).name
        (d / "genout.dat").unlink()
        t2 = time.perf_counter(); subprocess.run([str(exe)], cwd=d, stdout=subprocess.DEVNULL); t3 = time.perf_counter()
        t4 = time.perf_counter(); g = np.fromfile(d / "genout.dat", np.float32); t5 = time.perf_counter()
        print(name, "Solver.run %.2f s | child alone %.2f s | np.fromfile(genout) %.2f s" % (t1 - t0, t3 - t2, t5 - t4), flush=True)
        if name == "ours":
            log = (d / "fw2_execution.log").read_text().splitlines()
            print("\n".join(l for l in log if "fw25_engine:" in l))
