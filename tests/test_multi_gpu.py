"""N-slab x-sharded runs must be bit-identical to one domain.

The torchrun / NCCL case needs two GPUs (NCCL refuses two ranks on one device).  Everything that runs in ONE process
-- fw25_run's native multi-device runner (fused NVLink push and copies, both schedules), the Python lockstep driver,
the `fw25_engine` executable, box sensors, the reference goldens -- also runs on a single-GPU box by placing several
slabs on device 0 (`device_ids=(0, 0)`, `MultiRun::init`): same code path, same halo schedule, peer pointers that happen
to be local."""

import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.gpu


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:  # noqa: BLE001
        return 0


def _devices(n=2):
    """n slabs on n GPUs when the box has them, else all on device 0."""
    return tuple(range(n)) if _n_gpus() >= n else (0,) * n


@pytest.mark.parametrize("schedule", ["serial", "concurrent"])
@pytest.mark.parametrize("case", ["het3d", "het2d"])
def test_two_gpus_bit_identical_to_one_domain(built_lib, case, schedule):
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    import os
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29531", str(ROOT / "tools" / "slab_check.py"), case],
                       capture_output=True, text=True, timeout=600, cwd=ROOT,
                       env=dict(os.environ, FW25_SLAB_SCHEDULE=schedule))
    lines = [l for l in r.stdout.splitlines() if l.startswith("SLABCHECK ")]
    assert r.returncode == 0 and lines, r.stdout[-2000:] + r.stderr[-2000:]
    verdict = json.loads(lines[-1][len("SLABCHECK "):])
    assert verdict["bit_exact"] and verdict["absmax"] > 0, verdict


@pytest.mark.parametrize("mode", ["native", "native-copies", "native-concurrent", "native-concurrent-copies", "torch"])
@pytest.mark.parametrize("case", ["het3d", "het2d_long"])
def test_in_process_device_list_bit_identical(built_lib, case, mode):
    """`engine.run(pb, device_ids=(0, 1))` -- what Launcher(cuda_device_id=[0, 1]) calls -- one host thread, two
    GPUs: fw25_run's native multi-device runner (3D: boundary sweeps push their planes into the neighbour's ghost
    planes over NVLink; "native-copies": the same schedule with peer-to-peer copies), and the Python lockstep
    driver over the same C-ABI pieces."""
    import os
    env = dict(os.environ, FW25_FUSED_HALO="0" if mode.endswith("copies") else "1",
               FW25_SLAB_SCHEDULE="concurrent" if "concurrent" in mode else "serial",
               FW25_TEST_DEVICES=",".join(map(str, _devices())))
    r = subprocess.run([sys.executable, str(ROOT / "tools" / "slab_check.py"), case, "2", mode.split("-")[0]],
                       capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    lines = [l for l in r.stdout.splitlines() if l.startswith("SLABCHECK ")]
    assert r.returncode == 0 and lines, r.stdout[-2000:] + r.stderr[-2000:]
    verdict = json.loads(lines[-1][len("SLABCHECK "):])
    assert verdict["bit_exact"] and verdict["absmax"] > 0, verdict
    if mode.startswith("native"):
        assert verdict["n_devices"] == 2 and verdict["halo_bytes"] > 0, verdict


def test_executable_shards_over_visible_devices(built_lib, tmp_path):
    """`fw25_engine` in a .dat directory with CUDA_VISIBLE_DEVICES=0,1 (what the reference launcher sets for
    cuda_device_id=[0, 1], launcher.py:206) uses both GPUs and writes the same genout.dat as with one."""
    import os

    import numpy as np

    from fullwave25_b200.build import CLI
    from tests.test_slab import _problem
    pb = _problem("het3d")
    outs = {}
    for devs in ("0", "0,1"):
        d = tmp_path / f"sim_{len(devs)}"
        pb.to_dat_dir(d)
        env = dict(os.environ, CUDA_VISIBLE_DEVICES=devs)
        if devs == "0,1" and _n_gpus() < 2:       # one GPU: the same two slabs, both on device 0
            env = dict(os.environ, CUDA_VISIBLE_DEVICES="0", FW25_DEVICE_LIST="0,0")
        r = subprocess.run([str(CLI)], cwd=d, capture_output=True, text=True, timeout=600, env=env)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        assert f"{len(devs.split(','))} GPU(s)" in r.stdout
        outs[devs] = np.fromfile(d / "genout.dat", np.float32)
    assert np.abs(outs["0"]).max() > 0
    np.testing.assert_array_equal(outs["0"], outs["0,1"])


@pytest.mark.parametrize("case", ["far3d", "far2d"])
def test_two_gpu_run_against_reference_goldens(built_lib, case):
    """fw25_run on two devices is bit-identical to the reference's ONE-GPU traces; the reference's own 2-GPU traces
    (tests/golden/ref_<case>_g2.npz) deviate from those by 2e-5 .. 6e-3 rel-L2 on these cases (its sensors on the
    second GPU read plane x-1, tests/test_oracle_golden.py), and so from ours by the same."""
    import numpy as np

    from fullwave25_b200 import engine
    from tests import cases
    from tests.test_oracle_golden import load_golden
    got, stats = engine.run(cases.make(case), device_ids=_devices())
    assert stats["n_devices"] == 2
    np.testing.assert_array_equal(got, load_golden(case))
    g2 = load_golden(case, "_g2").astype(np.float64)
    assert np.linalg.norm(got - g2) / np.linalg.norm(g2) < 1e-2


@pytest.mark.parametrize("name,lo,hi", [("het3d", (10, 9, 12), (40, 20, 31)), ("het2d_long", (0, 0), (120, 100))])
def test_box_sensors_on_two_devices(built_lib, name, lo, hi):
    """Box sensors split over two x-slabs (fw25_run with a device list): each slab records its planes of the box
    without an index list and the frames come back in the global row-major order -- identical to one device."""
    import numpy as np

    from fullwave25_b200 import engine
    from tests.test_box_sensors import with_box
    listed, boxed = with_box(name, lo, hi)
    one, _ = engine.run(boxed)
    two, stats = engine.run(boxed, device_ids=_devices())
    assert stats["n_devices"] == 2 and np.abs(one).max() > 0
    np.testing.assert_array_equal(two, one)
    np.testing.assert_array_equal(engine.run(listed, device_ids=_devices())[0], one)


@pytest.mark.parametrize("n", [3, 4])
@pytest.mark.parametrize("fused", ["1", "0"])
def test_many_slabs_with_interior_ranks(built_lib, n, fused, monkeypatch):
    """Three and four slabs: interior slabs exchange with two neighbours, slabs of 3 x-marching chunks, a source
    layer and sensors ON the interfaces, air voxels in ghost planes -- bit-identical to one domain."""
    import numpy as np

    from fullwave25_b200 import engine, synthetic
    from fullwave25_b200.slab import partition
    monkeypatch.setenv("FW25_FUSED_HALO", fused)
    pb = synthetic.make_problem((n * 40, 44, 48), nT=48, modT=3, seed=31, n_pml=5, n_trans=3, n_sensors=96, n_air=24)
    rng = np.random.default_rng(5)
    edges = [s.own_hi for s in partition(pb.nX, n)[:-1]]
    extra_out, extra_air, extra_src = [], [], []
    for e in edges:                                   # both sides of every interface
        for x in (e - 1, e):
            yz = rng.integers(18, 26, size=(6, 2))
            extra_out += [(x, y, z) for y, z in yz]
            extra_air += [(x, int(yz[0, 0]) + 3, int(yz[0, 1]) + 3)]
            extra_src += [(x, int(yz[1, 0]) - 2, int(yz[1, 1]) + 5)]
    pb.outc = np.concatenate([pb.outc, np.asarray(extra_out, np.int32)])
    pb.icczero = np.concatenate([pb.icczero, np.asarray(extra_air, np.int32)])
    pb.icc = np.concatenate([pb.icc, np.asarray(extra_src, np.int32)])
    pb.icmat = np.concatenate([pb.icmat, np.repeat(pb.icmat[:1] * 0.5, len(extra_src), axis=0)])
    pb.normalise()
    one, _ = engine.run(pb)
    many, stats = engine.run(pb, device_ids=_devices(n))
    assert stats["n_devices"] == n and stats["halo_bytes"] > 0 and np.abs(one).max() > 0
    np.testing.assert_array_equal(many, one)
