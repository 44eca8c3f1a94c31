"""N-GPU x-slab run (torchrun, NCCL halo exchange) must be bit-identical to one domain.  Needs >= 2 GPUs."""

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


@pytest.mark.parametrize("case", ["het3d", "het2d"])
def test_two_gpus_bit_identical_to_one_domain(built_lib, case):
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29531", str(ROOT / "tools" / "slab_check.py"), case],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    lines = [l for l in r.stdout.splitlines() if l.startswith("SLABCHECK ")]
    assert r.returncode == 0 and lines, r.stdout[-2000:] + r.stderr[-2000:]
    verdict = json.loads(lines[-1][len("SLABCHECK "):])
    assert verdict["bit_exact"] and verdict["absmax"] > 0, verdict


@pytest.mark.parametrize("case", ["het3d", "het2d_long"])
def test_in_process_device_list_bit_identical(built_lib, case):
    """`engine.run(pb, device_ids=(0, 1))` -- what Launcher(cuda_device_id=[0, 1]) calls -- one host thread, two GPUs."""
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    r = subprocess.run([sys.executable, str(ROOT / "tools" / "slab_check.py"), case, "2"], capture_output=True, text=True,
                       timeout=600, cwd=ROOT)
    lines = [l for l in r.stdout.splitlines() if l.startswith("SLABCHECK ")]
    assert r.returncode == 0 and lines, r.stdout[-2000:] + r.stderr[-2000:]
    verdict = json.loads(lines[-1][len("SLABCHECK "):])
    assert verdict["bit_exact"] and verdict["absmax"] > 0, verdict
