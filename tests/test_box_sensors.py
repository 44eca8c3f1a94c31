"""Box sensors (include/fw25.h `out_box`): every point of a box is a sensor, row-major -- a
rectangular `Sensor(mask)` (sensor.py:24-50) or `Solver.run(record_whole_domain=True)` (solver.py:709-731).  The
engine records them without a coordinate or index list; results must be bit-identical to listing the coordinates."""

import copy

import numpy as np
import pytest

from fullwave25_b200 import engine
from fullwave25_b200.problem import Problem
from oracle import oracle
from tests import cases


def box_coords(lo, hi):
    grids = np.meshgrid(*[np.arange(a, b) for a, b in zip(lo, hi)], indexing="ij")
    return np.stack([g.reshape(-1) for g in grids], axis=1).astype(np.int32)


def with_box(name, lo, hi, **kw):
    pb = cases.make(name)
    for k, v in kw.items():
        setattr(pb, k, v)
    listed = copy.copy(pb)
    listed.outc = box_coords(lo, hi)
    boxed = copy.copy(pb)
    boxed.out_box = tuple(lo) + tuple(hi)
    return listed, boxed


def test_problem_box_is_the_row_major_coordinate_list(tmp_path):
    listed, boxed = with_box("het3d", (5, 0, 9), (9, 7, 12))
    boxed.normalise()
    assert boxed.ncoordsout == 4 * 7 * 3 == listed.outc.shape[0]
    np.testing.assert_array_equal(boxed.sensor_coords(), listed.outc)
    mask = np.zeros(boxed.shape, bool)
    mask[5:9, 0:7, 9:12] = True                       # what Sensor(mask).outcoords is upstream: np.where order
    np.testing.assert_array_equal(boxed.sensor_coords(), np.stack(np.where(mask), axis=1))
    back = Problem.from_dat_dir(boxed.to_dat_dir(tmp_path / "d"))       # the .dat protocol has no boxes: listed
    np.testing.assert_array_equal(back.outc, listed.outc)
    assert boxed.slab(8, 30).out_box == boxed.out_box
    with pytest.raises(ValueError):
        bad = copy.copy(boxed)
        bad.out_box = (1, 2, 3)
        bad.normalise()


@pytest.mark.gpu
@pytest.mark.parametrize("name,lo,hi", [
    ("het3d", (10, 9, 12), (30, 20, 31)),           # interior box, 4180 points
    ("het3d_ragged", (0, 0, 0), (45, 47, 53)),      # the whole extended grid (record_whole_domain): rim reads 0
    ("het2d", (3, 5), (70, 88)),                    # 2D, touches the rim on three sides
    ("het2d_long", (0, 0), (120, 100)),             # 2D whole domain, graph-replayed steps, nT % modT != 0
])
def test_box_sensors_equal_listed_sensors_and_oracle(built_lib, name, lo, hi, monkeypatch):
    listed, boxed = with_box(name, lo, hi)
    if name == "het2d_long":
        listed.nT = boxed.nT = 203
    want = oracle.run(listed)
    got_box, _ = engine.run(boxed)
    np.testing.assert_array_equal(got_box, want)
    # a listed box of >= 4096 coordinates is recognised and takes the same path; force the index path by breaking
    # the row-major order (swap two rows) and compare column-permuted results
    got_list, _ = engine.run(listed)
    np.testing.assert_array_equal(got_list, want)
    perm = np.arange(len(listed.outc))
    perm[[1, 2]] = perm[[2, 1]]
    swapped = copy.copy(listed)
    swapped.outc = listed.outc[perm]
    got_swapped, _ = engine.run(swapped)
    np.testing.assert_array_equal(got_swapped[:, perm], want)
    assert np.abs(want).max() > 0
    monkeypatch.setenv("FW25_GRAPH", "0")            # launched steps instead of graph replay
    np.testing.assert_array_equal(engine.run(boxed)[0], want)


@pytest.mark.gpu
def test_box_errors_and_empty_box(built_lib):
    _, boxed = with_box("het2d", (3, 5), (3, 88))    # empty along x
    got, _ = engine.run(boxed)
    assert got.shape == (boxed.n_frames, 0)
    _, boxed = with_box("het2d", (3, 5), (81, 88))   # hi beyond nX = 80
    with pytest.raises(engine.EngineError, match="out_box"):
        engine.run(boxed)


@pytest.mark.gpu
def test_box_sensors_on_slabs_keep_global_order(built_lib):
    """Two x-slabs in one process (fw25_run with a device list) need two GPUs; the handle API shows the same thing
    on one: each slab engine owns a contiguous run of the box's rows."""
    listed, boxed = with_box("het3d", (10, 9, 12), (30, 20, 31))
    boxed.normalise()
    nX = boxed.nX
    cut = 24
    lo_eng = engine.Engine(boxed.slab(0, cut + 8), slab=(nX, 0, 0, cut))
    hi_eng = engine.Engine(boxed.slab(cut - 8, nX), slab=(nX, cut - 8, cut, nX))
    per_plane = 11 * 19
    assert lo_eng.n_local_sensors == (cut - 10) * per_plane and hi_eng.n_local_sensors == (30 - cut) * per_plane
    np.testing.assert_array_equal(lo_eng.local_sensor_ids(), np.arange(0, (cut - 10) * per_plane))
    np.testing.assert_array_equal(hi_eng.local_sensor_ids(), np.arange((cut - 10) * per_plane, 20 * per_plane))
    lo_eng.close(); hi_eng.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name,cap,stream", [("het2d_long", 0, "2"), ("het2d_long", 7, "2"), ("het2d_long", 7, "0"),
                                             ("het3d_long", 3, "2"), ("het2d_long", 1, "2")])
def test_frames_streamed_during_the_loop_equal_frames_read_at_the_end(built_lib, name, cap, stream, monkeypatch):
    """Large recordings leave the device while the time loop runs (FrameStreamer: copier thread, events, the loop
    stalls only when the ring is full).  Forced on here for small problems, with small rings (wrap-around, ring-full
    waits) and small batches; also the synchronous ring-full path (`stream` = "0")."""
    pb = cases.make(name)
    shape = pb.shape
    pb.out_box = (0,) * pb.ndim + tuple(shape)          # whole-domain recording
    pb.modT = 2
    want = oracle.run(_listed(pb))
    monkeypatch.setenv("FW25_STREAM_FRAMES", stream)
    monkeypatch.setenv("FW25_STREAM_BATCH_KB", "64")
    if cap:
        monkeypatch.setenv("FW25_FRAMES_CAP", str(cap))
    got, stats = engine.run(pb)
    np.testing.assert_array_equal(got, want)
    assert stats["d2h_bytes"] == want.nbytes and np.abs(want).max() > 0
    with engine.Engine(pb) as e:                        # fw25_run_engine on a live engine, twice (transmit events)
        a, _ = e.run()
        e.reset(pb.icc, pb.icmat)
        b, _ = e.run()
    np.testing.assert_array_equal(a, want)
    np.testing.assert_array_equal(b, want)


def _listed(pb):
    q = copy.copy(pb)
    q.outc = pb.sensor_coords()
    q.out_box = None
    return q
