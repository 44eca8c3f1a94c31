"""CPU oracle (oracle/fw25_oracle.c) known-answer and property tests.  The oracle is test
infrastructure; these tests make sure the checker itself behaves like a wave solver and like the
reference's loop."""

import numpy as np
import pytest

from fullwave25_b200 import synthetic
from oracle import oracle
from tests import cases


def test_zero_source_stays_zero():
    pb = cases.make("het3d")
    pb.icmat[:] = 0
    g, f = oracle.run(pb, return_fields=True)
    assert not g.any() and not f["p"].any() and not f["u"].any()


def test_frames_layout_and_count():
    pb = cases.make("het3d_ragged")          # nT = 40, modT = 3 -> frames at t = 0, 3, ..., 39
    g = oracle.run(pb)
    assert g.shape == (14, pb.ncoordsout)
    st = oracle.Stepper(pb)
    frames = []
    for t in range(pb.nT):
        st.step()
        if t % pb.modT == 0:
            frames.append(st.record())
    np.testing.assert_array_equal(np.stack(frames), g)


@pytest.mark.parametrize("name", ["hom2d", "hom3d"])
def test_plane_wave_speed(name):
    """Homogeneous medium: the plane pulse must travel at c -- the lag that maximises the
    cross-correlation of two sensors 20 cells apart is 20 dX / (c dT) steps (within one step)."""
    kw = dict(cases.CASES[name])
    kw.update(nT=280, modT=1, n_pml=8, n_trans=4)
    kw["shape"] = (96, 88, 88) if name == "hom3d" else (110, 70)
    pb = synthetic.make_problem(**kw)
    pb.beta[:] = 0.5                       # 1 - 2 beta = 0: linear
    nb = 8 + 8 + 4
    mid = [s // 2 for s in pb.shape[1:]]
    x1, x2 = nb + 10, nb + 10 + 20
    pb.outc = np.array([[x1, *mid], [x2, *mid]], np.int32)
    g = oracle.run(pb).astype(np.float64)
    lags = np.arange(60, 140)
    xc = [np.dot(g[: len(g) - L, 0], g[L:, 1]) for L in lags]
    lag = lags[int(np.argmax(xc))]
    c = float(pb.extra["c"].flat[0])
    expected = 20 * pb.dX / c / pb.dT
    assert abs(lag - expected) <= 1.0, (lag, expected)
    peak = np.abs(g).max(axis=0)
    assert 0.7 * peak[0] < peak[1] < 1.3 * peak[0]          # plane wave: no geometric spreading


def test_attenuation_reduces_amplitude():
    kw = dict(shape=(110, 70), nT=260, modT=1, seed=4, homogeneous=True, n_air=0, n_pml=8, n_trans=4)
    pb = synthetic.make_problem(**kw)
    nb = 20
    pb.outc = np.array([[nb + 10, 35], [nb + 50, 35]], np.int32)
    g = np.abs(oracle.run(pb)).max(axis=0)
    assert g[1] < 0.999 * g[0]             # relaxation mechanisms attenuate the travelling pulse


def test_air_voxels_are_pressure_release():
    pb = cases.make("het3d")
    st = oracle.Stepper(pb)
    for _ in range(30):
        st.step()
    st.inject()                             # state at the start of step 30: air voxels forced to 0
    p = st.field("p")
    a = pb.icczero
    assert not p[a[:, 0], a[:, 1], a[:, 2]].any()


def test_rim_never_updated_and_rim_sensor_reads_zero():
    pb = cases.make("het2d")
    pb.outc = np.vstack([pb.outc, [[3, 40], [40, 2]]]).astype(np.int32)
    g, f = oracle.run(pb, return_fields=True)
    assert not g[:, -2:].any()
    for k in "puv":
        assert not f[k][:8].any() and not f[k][-8:].any() and not f[k][:, :8].any() and not f[k][:, -8:].any()


def test_sub_range_sweeps_compose():
    """Sweeping x in pieces equals one full sweep (what the slab / boundary-first schedule relies on)."""
    pb = cases.make("het3d")
    a, b = oracle.Stepper(pb), oracle.Stepper(pb)
    for t in range(12):
        a.step()
        b.inject()
        for lo, hi in ((30, 48), (0, 17), (17, 30)):
            b.sweep_u(lo, hi)
        for lo, hi in ((25, 48), (0, 25)):
            b.sweep_p(lo, hi)
        b.t += 1
    np.testing.assert_array_equal(a.mem, b.mem)


def test_portable_and_fma_builds_agree():
    import ctypes as C
    from pathlib import Path
    oracle.build()
    pb = cases.make("het2d")
    s, keep = oracle._marshal(pb)
    outs = []
    for name in ("libfw25_oracle.so", "libfw25_oracle_fma.so"):
        if name.endswith("_fma.so") and not oracle._cpu_has_fma():
            pytest.skip("no FMA on this CPU")
        lib = C.CDLL(str(Path(oracle.__file__).parent / name))
        lib.fw25o_run.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        g = np.zeros((oracle.n_frames(pb), s.ncoordsout), np.float32)
        assert lib.fw25o_run(C.byref(s), g.ctypes.data, None) == 0
        outs.append(g)
    np.testing.assert_array_equal(outs[0], outs[1])
