"""Host logic of bench.py's reference arm (tools/bench_reference.py): the per-step clock read off the reference
engine's progress prints.  No GPU, no reference binary: a writer thread plays the engine's part -- a header, then
10 bytes per time step appended to fw2_execution.log (ASM 0x407c88-0x407cb7)."""

import json
import subprocess
import sys
import threading
import time
from pathlib import Path

import pytest

from tools import bench_reference as br

ROOT = Path(__file__).resolve().parent.parent


def test_window_starts_and_ends_on_device_synchronised_prints():
    # the print of step t is on the device's clock when step t - 1 recorded a frame: (t - 1) % modT == 0
    stamps = [0.1 * k for k in range(40)]
    t0, t1, secs = br.step_times(stamps, W=5, K=20, modT=4)
    assert (t0 - 1) % 4 == 0 and (t1 - 1) % 4 == 0
    assert t0 >= 5 and t1 - t0 >= 20
    assert (t0, t1) == (5, 25)
    assert secs == pytest.approx(0.1 * (t1 - t0))
    # a warm-up count that is not on such a print moves forward to the next one, never backward
    t0, t1, _ = br.step_times(stamps, W=6, K=20, modT=4)
    assert (t0, t1) == (9, 29)
    t0, t1, _ = br.step_times(stamps, W=3, K=7, modT=1)
    assert (t0, t1) == (3, 10)


def test_window_that_does_not_fit_is_reported_not_clamped():
    assert br.step_times([0.0] * 10, W=5, K=20, modT=4) is None
    nT = br.ref_nT(5, 20, 4)
    assert br.step_times([0.0] * (nT - 1), W=5, K=20, modT=4) is not None      # the run is long enough by construction
    for W, K, modT in [(5, 20, 4), (3, 3, 7), (20, 200, 4), (0, 1, 1)]:
        n = br.ref_nT(W, K, modT)
        win = br.step_times(list(range(n)), W, K, modT)
        assert win is not None and win[1] < n and win[1] - win[0] >= K


def test_progress_tail_counts_ten_bytes_per_step(tmp_path):
    log = tmp_path / "fw2_execution.log"
    steps, period = 60, 0.004
    written = []

    def engine():
        with open(log, "wb", buffering=0) as f:
            f.write(b"GPU 0: global range [0, 64)\nBody region ...\n" + br.PROGRESS_TAG)
            for k in range(steps):
                f.write(b"\b\b\b\b\b%0.3f" % (k / steps))               # exactly BYTES_PER_STEP bytes
                written.append(time.perf_counter())
                time.sleep(period)
            f.write(b"\nDone\n")                                           # trailing text shorter than one step

    tail = br.ProgressTail(log, period_s=2e-4)
    tail.start()
    th = threading.Thread(target=engine)
    th.start()
    th.join()
    time.sleep(0.05)
    tail.stop()
    assert tail.header is not None
    assert len(tail.stamps) == steps
    assert all(b >= a for a, b in zip(tail.stamps, tail.stamps[1:]))
    # every stamp follows its write closely (polling period + scheduling), and the window length is the writer's
    lag = [s - w for s, w in zip(tail.stamps, written)]
    assert min(lag) >= -1e-3 and max(lag) < 0.5
    t0, t1, secs = br.step_times(tail.stamps, W=5, K=40, modT=4)
    assert secs == pytest.approx(written[t1] - written[t0], abs=0.5)


def test_progress_tail_without_a_header_has_no_stamps(tmp_path):
    log = tmp_path / "fw2_execution.log"
    log.write_bytes(b"error: out of memory\n" * 20)
    tail = br.ProgressTail(log, period_s=1e-3)
    tail.start()
    time.sleep(0.05)
    tail.stop()
    assert tail.header is None and tail.stamps == []


def test_reference_arm_says_unavailable_without_a_gpu():
    """On a box without a CUDA device the arm prints {"impl": "reference", "unavailable": ...} and exits 0 -- it never
    times anything else in the reference's place."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("this box has a GPU: the arm would run the reference engine")
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "20", "--warmup", "5"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-400:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line.get("unavailable")
    assert "value" not in line


def test_our_arm_fails_loudly_without_a_gpu():
    """No CPU fallback: without a CUDA device bench.py's own arm exits non-zero and prints no result line."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("this box has a GPU")
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "2", "--warmup", "3", "--no-cpu-baseline"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode != 0
    assert not [ln for ln in r.stdout.splitlines() if ln.startswith("{") and '"value"' in ln]
