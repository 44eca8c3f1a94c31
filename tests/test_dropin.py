"""Drop-in boundary against the REFERENCE's own Python layer (needs the reference package: /root/reference in the
build container, baseline/_ref on the GPU box; skipped when neither is present).

CPU: what `Problem.from_fullwave_objects` assembles in memory is byte-identical to the directory the reference's
`InputFileWriter` writes for the same PML-extended objects, and `Problem.from_dat_dir` reads that directory back.
GPU: `fullwave.Solver.run` gives identical sensor data with (a) the reference's shipped sm_100 binary, (b) the
`fw25_engine` executable passed as path_fullwave_simulation_bin, (c) `launcher.install()`, (d) `run_solver`."""

import shutil
import tempfile
from pathlib import Path

import numpy as np
import pytest

from fullwave25_b200 import launcher
from fullwave25_b200.problem import MAP_NAMES, Problem

try:
    from tools import ref_objects
    _fw = ref_objects.import_fullwave()
    HAVE_REF = True
except Exception:  # noqa: BLE001
    HAVE_REF = False

needs_ref = pytest.mark.skipif(not HAVE_REF, reason="reference package not available")


def _write_with_reference(shape, td, **kw):
    fw, grid, medium, source, sensor = ref_objects.build(shape, **kw)
    from fullwave.solver.input_file_writer import InputFileWriter
    from fullwave.solver.pml_builder import PMLBuilder
    pml = PMLBuilder(grid=grid, medium=medium, source=source, sensor=sensor, m_spatial_order=8,
                     n_pml_layer=6, n_transition_layer=4, use_isotropic_relaxation=True)
    ext = pml.run(use_pml=True)
    w = InputFileWriter(work_dir=Path(td), grid=pml.extended_grid, medium=ext, source=pml.extended_source,
                        sensor=pml.extended_sensor, path_fullwave_simulation_bin=ref_objects.ref_bin(len(shape)),
                        use_exponential_attenuation=False, use_isotropic_relaxation=True)
    sim_dir = w.run("txrx_0", is_static_map=False, recalculate_pml=True)
    return pml, ext, Path(sim_dir)


@needs_ref
@pytest.mark.parametrize("shape", [(20, 24), (12, 14, 16)])
def test_in_memory_problem_equals_reference_dat_directory(shape):
    with tempfile.TemporaryDirectory() as td:
        pml, ext, sim_dir = _write_with_reference(shape, td, n_steps=30, n_sensors=9, n_air=5)
        from_files = Problem.from_dat_dir(sim_dir)
        in_mem = Problem.from_fullwave_objects(pml.extended_grid, ext, pml.extended_source, pml.extended_sensor)
    for k in ("ndim", "nX", "nY", "nZ", "nT", "nTic", "modT", "ndmap", "dX", "dT"):
        assert getattr(from_files, k) == getattr(in_mem, k), k
    for k in MAP_NAMES + ("dmap", "dcmap", "icc", "icmat", "outc", "icczero"):
        a, b = getattr(from_files, k), getattr(in_mem, k)
        assert a.dtype == b.dtype and a.shape == b.shape, k
        assert a.tobytes() == b.tobytes(), k
    assert 0 < from_files.ncoordszero <= 5 and from_files.ncoordsout == 9


@needs_ref
@pytest.mark.parametrize("shape", [(20, 24), (12, 14, 16)])
def test_lean_lists_and_lazy_pml_builder_equal_the_reference_objects(shape):
    """What the in-memory device path uses instead of the padded objects: coordinate lists shifted by the boundary
    width == the reference's extended source / sensor / air lists; the lazy PMLBuilder exposes the same grid and layer
    attributes without padding anything, and materialises the real builder on demand."""
    from fullwave.solver.pml_builder import PMLBuilder
    from fullwave25_b200 import mapgen
    fw, grid, medium, source, sensor = ref_objects.build(shape, n_steps=30, n_sensors=9, n_air=5)
    kw = dict(m_spatial_order=8, n_pml_layer=6, n_transition_layer=4, use_isotropic_relaxation=True)
    real = PMLBuilder(grid, medium, source, sensor, **kw)
    lazy = launcher.lazy_pml_builder_class(PMLBuilder)(grid, medium, source, sensor, **kw)
    src, sen, air = launcher.lean_lists(lazy)
    np.testing.assert_array_equal(src.incoords, real.extended_source.incoords)
    np.testing.assert_array_equal(src.icmat, real.extended_source.icmat)
    np.testing.assert_array_equal(sen.outcoords, real.extended_sensor.outcoords)
    assert sen.sampling_modulus_time == real.extended_sensor.sampling_modulus_time
    np.testing.assert_array_equal(air, np.stack(np.nonzero(real.extended_medium.air_map), axis=1))
    for a in ("nx", "ny", "nt", "num_boundary_points", "pml_layer_m", "transition_layer_m", "n_polynomial",
              "theoritical_reflection_coefficient", "is_3d"):
        assert getattr(lazy, a) == getattr(real, a), a
    for a in ("dt", "dx", "c0", "cfl", "nx", "ny", "nt"):
        assert getattr(lazy.extended_grid, a) == getattr(real.extended_grid, a), a
    s1, s2 = mapgen.MediumSpec.from_pml_builder(lazy), mapgen.MediumSpec.from_pml_builder(real)
    assert s1.extended_shape == s2.extended_shape and s1.d_target_pml() == s2.d_target_pml()
    assert isinstance(lazy, PMLBuilder) and lazy._full is None            # nothing padded so far
    ext = lazy.run(use_pml=True)                                         # the host path still works: materialised now
    want = real.run(use_pml=True)
    for k, v in want.relaxation_param_dict_for_fw2.items():
        np.testing.assert_array_equal(ext.relaxation_param_dict_for_fw2[k], v, err_msg=k)
    np.testing.assert_array_equal(lazy.extended_source.incoords, real.extended_source.incoords)


def test_device_id_forms_match_reference_launcher():
    assert launcher.parse_cuda_device_id(None) == "0"
    assert launcher.parse_cuda_device_id(2) == "2"
    assert launcher.parse_cuda_device_id("3") == "3"
    assert launcher.parse_cuda_device_id([0, 1, 2]) == "0,1,2"
    assert launcher.device_ids_of([1, 3]) == (1, 3)
    for bad in (-1, "a", [0, -1], 1.5, "0,1"):
        with pytest.raises(ValueError):
            launcher.parse_cuda_device_id(bad)
    if HAVE_REF:
        from fullwave.solver.launcher import Launcher as RefLauncher
        for v in (None, 2, "3", [0, 1, 2]):
            assert RefLauncher._parse_cuda_device_id(v) == launcher.parse_cuda_device_id(v)


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(40, 48), (20, 24, 28)])
def test_solver_run_identical_through_every_boundary(built_lib, shape):
    from fullwave25_b200 import build
    fw, grid, medium, source, sensor = ref_objects.build(shape, n_steps=60, n_sensors=12, n_air=6, modT=3)
    kw = dict(pml_layer_thickness_px=6, n_transition_layer=4)
    ndim = len(shape)
    out = {}
    with tempfile.TemporaryDirectory() as td:
        s_ref = fw.Solver(Path(td) / "ref", grid, medium, source, sensor,
                          path_fullwave_simulation_bin=ref_objects.ref_bin(ndim), **kw)
        out["reference binary"] = s_ref.run()
        s_cli = fw.Solver(Path(td) / "cli", grid, medium, source, sensor,
                          path_fullwave_simulation_bin=build.CLI, **kw)
        out["fw25_engine executable"] = s_cli.run()
        undo = launcher.install()
        try:
            s_ins = fw.Solver(Path(td) / "ins", grid, medium, source, sensor,
                              path_fullwave_simulation_bin=build.CLI, **kw)
            assert isinstance(s_ins.fullwave_launcher, launcher.Launcher)
            out["launcher.install()"] = s_ins.run()
        finally:
            undo()
        out["run_solver (no disk)"] = launcher.run_solver(s_ref)
    want = out.pop("reference binary")
    assert want.shape == (12, 20) and np.abs(want).max() > 0
    for name, got in out.items():
        assert got.shape == want.shape, name
        np.testing.assert_array_equal(got, want, err_msg=name)


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("shape,iso", [((40, 48), True), ((20, 24, 28), True), ((18, 20, 22), False)])
def test_run_solver_with_gpu_built_maps_matches_reference_binary(built_lib, shape, iso):
    """`run_solver(solver, maps="device")`: PMLBuilder.run and the extended-grid host arrays are replaced by
    fw25_mapgen.  a / b may differ from the reference's files by one float32 ulp (exp), so the traces are compared
    with the north star's tolerance, relative L2 <= 1e-5, against the reference's own binary; whole-domain recording
    and a transmit-event session ride on the same maps."""
    fw, grid, medium, source, sensor = ref_objects.build(shape, n_steps=60, n_sensors=12, n_air=6, modT=3)
    kw = dict(pml_layer_thickness_px=6, n_transition_layer=4, use_isotropic_relaxation=iso)
    if not iso:
        medium.use_isotropic_relaxation = False      # (else Solver logs a warning whose own format string is broken)
    with tempfile.TemporaryDirectory() as td:
        s_ref = fw.Solver(Path(td) / "ref", grid, medium, source, sensor,
                          path_fullwave_simulation_bin=ref_objects.ref_bin(len(shape), isotropic=iso), **kw)
        want = s_ref.run()
        got, stats = launcher.run_solver(s_ref, maps="device", return_stats=True)
        assert stats["maps"] == "device" and stats["mapgen_kernel_ms"] > 0
        assert got.shape == want.shape and np.abs(want).max() > 0
        rel = np.linalg.norm(got.astype(np.float64) - want) / np.linalg.norm(want.astype(np.float64))
        assert rel <= 1e-5, rel
        ref_wd = s_ref.run(record_whole_domain=True, sampling_modulus_time_whole_domain=20)    # the reference binary
        host = launcher.run_solver(s_ref, record_whole_domain=True, sampling_modulus_time_whole_domain=20)
        np.testing.assert_array_equal(host, ref_wd)          # box recording == the reference's whole-grid sensor list
        dev = launcher.run_solver(s_ref, record_whole_domain=True, sampling_modulus_time_whole_domain=20, maps="device")
        assert dev.shape == host.shape == (int(np.prod([n + 2 * 18 for n in shape])), 3)
        assert np.linalg.norm(dev.astype(np.float64) - host) <= 1e-5 * np.linalg.norm(host.astype(np.float64))
        with launcher.Session() as ses:
            first = launcher.run_solver(s_ref, maps="device", session=ses)
            again = launcher.run_solver(s_ref, maps="device", session=ses)     # maps stay resident, sources re-sent
        np.testing.assert_array_equal(first, got)
        np.testing.assert_array_equal(again, got)
        # a device list: every slab builds its own planes of the maps (fw25_run_medium_multi); same bits as one device
        two, st2 = launcher.run_solver(s_ref, maps="device", cuda_device_id=[0, 0], return_stats=True)
        assert st2["n_devices"] == 2 and st2["maps"] == "device"
        np.testing.assert_array_equal(two, got)


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(40, 48), (20, 24, 28)])
def test_static_map_transmit_events_reuse_device_maps(built_lib, shape):
    """The reference's multi-transmit loop -- a new Solver per event, run(is_static_map=True, recalculate_pml=k==0)
    -- through launcher.install(): the second and third events reuse the engine of the first (maps stay in HBM,
    only the sources are replaced) and still match the reference binary event by event; so does the no-disk
    Session path.

    Reference behaviour pinned here: in static-map mode the reference does NOT link icczero.dat / ncoordszero.dat
    into the simulation directory (input_file_writer.py:647-706 lists the linked files), so its binary -- and
    any engine that honours the directory -- runs WITHOUT the air voxels; the no-disk Session path takes the air
    map from the medium object and therefore matches the reference's non-static runs."""
    from fullwave25_b200 import build
    fw, grid, medium, source, sensor = ref_objects.build(shape, n_steps=48, n_sensors=10, n_air=4, modT=2)
    kw = dict(pml_layer_thickness_px=6, n_transition_layer=4)
    ndim = len(shape)
    p0 = np.asarray(source.p0)
    sources = [source, fw.Source(-0.5 * p0, source.mask), fw.Source(np.roll(p0, 5, axis=1), source.mask)]

    def events(td, bin_path, static=True):
        outs, launchers = [], []
        for k, src in enumerate(sources):
            s = fw.Solver(Path(td), grid, medium, src, sensor, path_fullwave_simulation_bin=bin_path, **kw)
            outs.append(s.run(f"txrx_{k}", is_static_map=static, recalculate_pml=(k == 0) or not static))
            launchers.append(s.fullwave_launcher)
        return outs, launchers

    with tempfile.TemporaryDirectory() as td_ref, tempfile.TemporaryDirectory() as td_new, \
            tempfile.TemporaryDirectory() as td_full:
        want, _ = events(td_ref, ref_objects.ref_bin(ndim))
        want_air, _ = events(td_full, ref_objects.ref_bin(ndim), static=False)
        undo = launcher.install()
        try:
            got, las = events(td_new, build.CLI)
        finally:
            undo()
            launcher.release()
        assert [la.last_stats["maps_reused"] for la in las] == [False, True, True]
        with launcher.Session() as ses:
            nodisk = []
            for src in sources:
                s = fw.Solver(Path(td_new) / "nodisk", grid, medium, src, sensor,
                              path_fullwave_simulation_bin=build.CLI, **kw)
                nodisk.append(launcher.run_solver(s, session=ses))
            # a session holds sensors and recording period of its first event: a different Sensor is refused, not
            # silently answered with the old sensor's traces
            smask = np.asarray(sensor.mask).copy()
            smask.flat[np.flatnonzero(~smask)[:3]] = True
            other = fw.Solver(Path(td_new) / "nodisk2", grid, medium, sources[0],
                              fw.Sensor(mask=smask, sampling_modulus_time=2), path_fullwave_simulation_bin=build.CLI, **kw)
            with pytest.raises(ValueError, match="session"):
                launcher.run_solver(other, session=ses)
            slower = fw.Solver(Path(td_new) / "nodisk3", grid, medium, sources[0],
                               fw.Sensor(mask=np.asarray(sensor.mask), sampling_modulus_time=3),
                               path_fullwave_simulation_bin=build.CLI, **kw)
            with pytest.raises(ValueError, match="session"):
                launcher.run_solver(slower, session=ses)
    assert np.abs(want[0]).max() > 0 and not np.array_equal(want[0], want[1])
    assert not np.array_equal(want[0], want_air[0])          # the static-map directory lost the air voxels
    for k in range(3):
        np.testing.assert_array_equal(got[k], want[k], err_msg=f"event {k} (launcher.install, static maps)")
        np.testing.assert_array_equal(nodisk[k], want_air[k], err_msg=f"event {k} (Session)")


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(40, 48), (20, 24, 28)])
def test_anisotropic_solver_run_identical_through_every_boundary(built_lib, shape):
    """`Solver(use_isotropic_relaxation=False)`: the reference writes the per-axis file set and runs its anisotropic
    binary (which has no air-voxel kernel); the executable, the swapped launcher and the no-disk path give the same
    bits.  The reference's Python layer writes identical maps on every axis, so the engine runs its isotropic kernels."""
    from fullwave25_b200 import build
    fw, grid, medium, source, sensor = ref_objects.build(shape, n_steps=60, n_sensors=12, n_air=6, modT=3)
    medium.use_isotropic_relaxation = False      # (else Solver logs a warning whose own format string is broken)
    kw = dict(pml_layer_thickness_px=6, n_transition_layer=4, use_isotropic_relaxation=False)
    ndim = len(shape)
    out = {}
    with tempfile.TemporaryDirectory() as td:
        s_ref = fw.Solver(Path(td) / "ref", grid, medium, source, sensor,
                          path_fullwave_simulation_bin=ref_objects.ref_bin(ndim, isotropic=False), **kw)
        out["reference anisotropic binary"] = s_ref.run()
        assert (Path(td) / "ref" / "txrx_0" / "kappay.dat").exists()
        s_cli = fw.Solver(Path(td) / "cli", grid, medium, source, sensor, path_fullwave_simulation_bin=build.CLI, **kw)
        out["fw25_engine executable"] = s_cli.run()
        undo = launcher.install()
        try:
            s_ins = fw.Solver(Path(td) / "ins", grid, medium, source, sensor, path_fullwave_simulation_bin=build.CLI, **kw)
            out["launcher.install()"] = s_ins.run()
        finally:
            undo()
        out["run_solver (no disk)"] = launcher.run_solver(s_ref)
    want = out.pop("reference anisotropic binary")
    assert want.shape == (12, 20) and np.abs(want).max() > 0
    for name, got in out.items():
        np.testing.assert_array_equal(got, want, err_msg=name)


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(40, 48), (20, 24, 28)])
def test_patched_solver_run_in_memory(built_lib, shape):
    """`launcher.install(in_memory=True)`: the user's script keeps calling `fullwave.Solver(...).run(...)` -- same
    signature, same return value -- and nothing touches the disk.  maps="host" is bit-identical to the reference
    binary, maps="device" within 1e-5 relative L2; a static-map transmit sequence reuses the engine; calls that need the
    directory (load_results=False) still go through the original method."""
    import os
    from fullwave25_b200 import build
    fw, grid, medium, source, sensor = ref_objects.build(shape, n_steps=60, n_sensors=12, n_air=0, modT=3)
    kw = dict(pml_layer_thickness_px=6, n_transition_layer=4)
    with tempfile.TemporaryDirectory() as td:
        want = fw.Solver(Path(td) / "ref", grid, medium, source, sensor,
                         path_fullwave_simulation_bin=ref_objects.ref_bin(len(shape)), **kw).run()
        for maps in ("host", "device"):
            undo = launcher.install(in_memory=True, maps=maps)
            try:
                s = fw.Solver(Path(td) / f"mem_{maps}", grid, medium, source, sensor,
                              path_fullwave_simulation_bin=build.CLI, **kw)
                got = s.run()
                assert not (Path(td) / f"mem_{maps}" / "txrx_0").exists()          # no simulation directory
                if maps == "device":                                               # ... and nothing padded on the host
                    assert type(s.pml_builder).__name__ == "PMLBuilder" and s.pml_builder._full is None
                if maps == "host":
                    np.testing.assert_array_equal(got, want)
                else:
                    assert np.linalg.norm(got.astype(np.float64) - want) <= 1e-5 * np.linalg.norm(want.astype(np.float64))
                # a transmit sequence on a static map: event 0 builds, events 1, 2 reuse the engine
                seq = [fw.Solver(Path(td) / f"seq_{maps}", grid, medium, source, sensor,
                                 path_fullwave_simulation_bin=build.CLI, **kw)
                       .run(f"txrx_{k}", is_static_map=True, recalculate_pml=(k == 0)) for k in range(3)]
                assert launcher._StaticSession.session is not None and launcher._StaticSession.session.eng is not None
                for ev in seq:
                    np.testing.assert_array_equal(ev, got)
                path = s.run("to_disk", load_results=False)                          # falls through: directory + genout.dat
                assert os.path.exists(path) and str(path).endswith("genout.dat")
            finally:
                undo()
            assert launcher._StaticSession.session is None
            assert fw.Solver.run.__module__.startswith("fullwave")


@needs_ref
def test_install_in_memory_patches_and_restores_the_reference(built_lib, tmp_path):
    """CPU: `install(in_memory=True, maps="device")` swaps Launcher, `Solver.run` and the PMLBuilder `Solver.__init__`
    instantiates; constructing a Solver then pads nothing; `uninstall()` puts everything back.  (Without a GPU the run
    itself stops at the first CUDA call with the reference's error type.)"""
    import importlib
    from fullwave25_b200 import build, engine
    sol = importlib.import_module("fullwave.solver.solver")
    before = (sol.Launcher, sol.Solver.run, sol.PMLBuilder)
    fw, grid, medium, source, sensor = ref_objects.build((20, 24), n_steps=30, n_sensors=9, n_air=3)
    undo = launcher.install(in_memory=True, maps="device")
    try:
        assert sol.Launcher is launcher.Launcher and sol.Solver.run is not before[1]
        assert issubclass(sol.PMLBuilder, before[2]) and sol.PMLBuilder is not before[2]
        s = fw.Solver(tmp_path / "w", grid, medium, source, sensor, path_fullwave_simulation_bin=build.CLI,
                      pml_layer_thickness_px=6, n_transition_layer=4)
        assert s.pml_builder._full is None and s.pml_builder.extended_grid.nx == 20 + 2 * 18
        try:
            import torch
            has_gpu = torch.cuda.is_available()
        except Exception:  # noqa: BLE001
            has_gpu = False
        if not has_gpu:
            with pytest.raises(launcher.SimulationError):
                s.run()
            assert s.pml_builder._full is None            # the device path never asked for the padded objects
        with pytest.raises(ValueError):
            launcher.install(maps="gpu")
    finally:
        undo()
    assert (sol.Launcher, sol.Solver.run, sol.PMLBuilder) == before
    assert engine.lib() is not None


@needs_ref
def test_installed_errors_are_caught_as_the_reference_exception(built_lib, tmp_path):
    """CPU: after install() an engine failure raises a class that `except fullwave.solver.launcher.SimulationError`
    catches (and `except fullwave25_b200.launcher.SimulationError` too); uninstall() restores the plain class."""
    import importlib
    lau = importlib.import_module("fullwave.solver.launcher")
    ref_err = lau.SimulationError
    undo = launcher.install()
    try:
        (tmp_path / "nX.dat").write_bytes(b"")                   # a directory the engine cannot read
        la = launcher.Launcher(None, is_3d=False, use_gpu=True, cuda_device_id=0)
        with pytest.raises(ref_err) as info:
            la.run(tmp_path)
        assert isinstance(info.value, launcher.SimulationError)
    finally:
        undo()
    assert launcher._raise_cls is launcher.SimulationError and launcher._reference_launcher_cls is None


def test_exponential_attenuation_directory_goes_to_the_reference_launcher(built_lib, tmp_path, monkeypatch):
    """CPU: a simulation directory of the exponential-attenuation engine (a_exp.dat, no relaxation maps) is handed to
    the reference's own Launcher with the same arguments -- what install() promises for such solvers."""
    calls = []

    class FakeReferenceLauncher:
        def __init__(self, path, *, is_3d, use_gpu, cuda_device_id):
            calls.append(("init", path, is_3d, use_gpu, cuda_device_id))

        def run(self, simulation_dir, *, load_results=True):
            calls.append(("run", Path(simulation_dir), load_results))
            return "reference result"

    (tmp_path / "a_exp.dat").write_bytes(b"\0" * 16)
    la = launcher.Launcher(Path("/some/exp_binary"), is_3d=True, use_gpu=True, cuda_device_id=[0, 1])
    with pytest.raises(NotImplementedError):                     # not installed: nobody to hand the run to
        la.run(tmp_path)
    monkeypatch.setattr(launcher, "_reference_launcher_cls", FakeReferenceLauncher)
    assert la.run(tmp_path, load_results=False) == "reference result"
    assert calls == [("init", Path("/some/exp_binary"), True, True, [0, 1]), ("run", tmp_path.absolute(), False)]
    (tmp_path / "kappax.dat").write_bytes(b"")                  # relaxation maps present: this engine's directory
    assert not launcher._is_exponential_attenuation_dir(tmp_path)


def test_static_key_sees_rewritten_maps(tmp_path):
    """CPU: the identity of a static-map directory includes inode and change time of every linked map (and the
    anisotropic stems), so maps rewritten in place are not mistaken for the ones on the device."""
    work, sim = tmp_path / "work", tmp_path / "sim"
    work.mkdir(); sim.mkdir()
    for stem in ("rho", "kappay", "c"):
        (work / f"{stem}.dat").write_bytes(b"\1" * 64)
        (sim / f"{stem}.dat").symlink_to(work / f"{stem}.dat")
    k0 = launcher._static_key(sim, (0,))
    assert k0 is not None and launcher._static_key(sim, (0,)) == k0
    for stem in ("kappay", "c"):
        tmp = work / f"{stem}.tmp"
        tmp.write_bytes(b"\2" * 64)                               # same size, new inode
        tmp.replace(work / f"{stem}.dat")
        assert launcher._static_key(sim, (0,)) != k0
        k0 = launcher._static_key(sim, (0,))
    (sim / "rho.dat").unlink()
    (sim / "rho.dat").write_bytes(b"\1" * 64)                    # not a link: not the static-map layout
    assert launcher._static_key(sim, (0,)) is None
