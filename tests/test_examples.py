"""GPU: BASELINE.json configs 1 and 4 -- the reference's own example scripts (baseline/_ref/examples, staged by
tools/install_reference.sh), shortened in time, run verbatim through `fullwave.Solver.run` on the reference's sm_100
binary and on this engine.  With the maps built by the reference's PMLBuilder the sensor output is identical, bit for
bit; with the maps built on the GPU (a, b within one float32 ulp) it agrees to <= 1e-5 relative L2 (north star)."""

from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
HAVE = (ROOT / "baseline" / "_ref" / "examples" / "wave_3d").exists()
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not HAVE, reason="reference examples not staged (tools/install_reference.sh)")]


@pytest.mark.parametrize("name,scale", [("simple_plane_wave", 0.08), ("wave_3d", 0.12)])
def test_example_script_identical_on_both_engines(built_lib, name, scale):
    from tools import run_examples as R
    ref = R.run_example(name, "reference", scale)
    host = R.run_example(name, "fw25-host", scale)
    dev = R.run_example(name, "fw25-device", scale)
    assert ref["steps"] == host["steps"] == dev["steps"] > 100 and ref["sensors"] > 10_000
    assert np.abs(ref["out"]).max() > 0
    np.testing.assert_array_equal(host["out"], ref["out"])
    c = R.compare(dev["out"], ref["out"])
    assert c["rel_l2"] <= 1e-5, c
