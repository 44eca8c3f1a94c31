"""Host-side mirror of the reference's launcher layer -- the seam where the new engine plugs in.

Upstream, `fullwave.Solver.run` hands a directory of `.dat` files to
`fullwave.solver.launcher.Launcher.run` (/root/reference/fullwave/solver/launcher.py:160-254), which
exec's a pre-compiled CUDA binary there and reads `genout.dat` back.  This module keeps that interface
(same class name, constructor arguments, `run(simulation_dir, load_results=...)`, `SimulationError`,
`cuda_device_id` forms None / int / "N" / [ids]) and computes the result with libfw25.so in-process.

Three ways to use it:
  1. `install()`  -- swap the launcher class inside the reference package: `fullwave.Solver` then runs on
                     the new engine with no other change (the .dat directory is still written, genout.dat too);
  2. `run_solver(solver, ...)` -- same result without the disk round trip: the engine input is built
                     straight from the solver's PML-extended objects;
  3. the `fw25_engine` executable -- pass it as `Solver(path_fullwave_simulation_bin=...)`; zero edits.
There is no CPU path: like the reference (`use_gpu=False` raises NotImplementedError, launcher.py:191-194).
"""

from __future__ import annotations

import logging
import time
from pathlib import Path

import numpy as np

from . import engine
from .problem import MAP_NAMES, Problem

logger = logging.getLogger("__main__." + __name__)


class SimulationError(Exception):
    """Raised when the engine fails (the reference raises its own SimulationError on a non-zero exit status).
    After `install()` the raised class ALSO derives from `fullwave.solver.launcher.SimulationError`, so user code
    that catches the reference's exception keeps working."""


last_run_stats: dict | None = None   # engine statistics of the most recent run_solver call (fw25_stats + extras)
_raise_cls = SimulationError          # install() swaps in a subclass of both error types
_reference_launcher_cls = None        # the reference's own Launcher, kept by install() for the runs this engine does not cover


def _sim_error(msg: str) -> Exception:
    return _raise_cls(msg)


def _is_exponential_attenuation_dir(simulation_dir: Path) -> bool:
    """The exponential-attenuation file set (input_file_writer.py:629-650, :746-751): a_exp.dat, no relaxation maps."""
    return (simulation_dir / "a_exp.dat").exists() and not (simulation_dir / "kappax.dat").exists()


def parse_cuda_device_id(cuda_device_id) -> str:
    """None -> "0", 3 -> "3", "3" -> "3", [0, 1] -> "0,1"; same errors as launcher.py:63-105."""
    if cuda_device_id is None:
        return "0"
    if isinstance(cuda_device_id, bool):
        raise ValueError("CUDA device ID must be an integer, string, list, or None.")
    if isinstance(cuda_device_id, int):
        if cuda_device_id < 0:
            raise ValueError("CUDA device ID must be a non-negative integer.")
        return str(cuda_device_id)
    if isinstance(cuda_device_id, str):
        if not cuda_device_id.isdigit() or int(cuda_device_id) < 0:
            raise ValueError("CUDA device ID string must represent a non-negative integer.")
        return cuda_device_id
    if isinstance(cuda_device_id, list):
        if not all(isinstance(i, int) and not isinstance(i, bool) and i >= 0 for i in cuda_device_id):
            raise ValueError("All CUDA device IDs in the list must be non-negative integers.")
        return ",".join(str(i) for i in cuda_device_id)
    raise ValueError("CUDA device ID must be an integer, string, list, or None.")


def device_ids_of(cuda_device_id) -> tuple[int, ...]:
    return tuple(int(v) for v in parse_cuda_device_id(cuda_device_id).split(","))


# ------------------------------------------------------------------------------------------------------------
# Static maps: several transmit events on one medium.  Upstream every `Solver.run(is_static_map=True,
# recalculate_pml=False)` writes only icmat.dat and symlinks the ~20 map files of the work directory into the
# new simulation directory (input_file_writer.py:146-175, :647-714) -- and every launch of the binary reads and
# uploads all of them again.  Here the engine of the previous event stays alive: when the next directory's map
# files resolve to the same files (same real path, size and mtime), only the source list is replaced
# (fw25_reset) and the maps never leave HBM.

_STATIC_FILES = (MAP_NAMES + Problem.aniso_stems(3) +
                 ("dcmap", "dmap", "c", "outc", "icczero", "nX", "nY", "nZ", "modT", "dX", "dT", "ndmap", "ncoordsout",
                  "ncoordszero"))


def _static_key(simulation_dir: Path, device_ids) -> tuple | None:
    """Identity of everything but the sources, or None when the directory is not in the static-map layout
    (its map files are not symlinks)."""
    import os
    if not (simulation_dir / "rho.dat").is_symlink():
        return None
    key = [tuple(device_ids)]
    for stem in _STATIC_FILES:
        f = simulation_dir / f"{stem}.dat"
        if not f.exists():
            key.append((stem, None))
            continue
        st = f.stat()      # (follows the link: identity and change times of the work directory's file)
        key.append((stem, os.path.realpath(f), st.st_ino, st.st_size, st.st_mtime_ns, st.st_ctime_ns))
    return tuple(key)


class _LiveEngine:
    """At most one engine kept alive between `Launcher.run` calls (bounded device memory)."""
    key: tuple | None = None
    eng: "engine.Engine | None" = None

    @classmethod
    def release(cls) -> None:
        if cls.eng is not None:
            cls.eng.close()
        cls.key, cls.eng = None, None


def release() -> None:
    """Free the engine (and its device-resident maps) kept alive for static-map reuse."""
    _LiveEngine.release()
    _StaticSession.release()


def _run_dat_dir(simulation_dir: Path, device_ids) -> tuple[np.ndarray, dict]:
    key = _static_key(simulation_dir, device_ids) if len(device_ids) == 1 else None
    if key is not None and key == _LiveEngine.key:
        ndim = _LiveEngine.eng.pb.ndim
        i32 = lambda stem: int(np.fromfile(simulation_dir / f"{stem}.dat", dtype=np.int32)[0])  # noqa: E731
        ncoords, nTic, nT = i32("ncoords"), i32("nTic"), i32("nT")
        icc = np.fromfile(simulation_dir / "icc.dat", dtype=np.int32)[: ncoords * ndim].reshape(ncoords, ndim)
        icmat = np.fromfile(simulation_dir / "icmat.dat", dtype=np.float32)[: ncoords * nTic].reshape(ncoords, nTic)
        _LiveEngine.eng.reset(icc, icmat, nT)
        genout, stats = _LiveEngine.eng.run()
        stats["maps_reused"] = True
        return genout, stats
    pb = Problem.from_dat_dir(simulation_dir)
    if key is None:
        release()                 # a kept engine holds its maps + state in HBM: this run needs the memory
        return engine.run(pb, device_ids=device_ids)
    _LiveEngine.release()
    eng = engine.Engine(pb, device=device_ids[0])
    _LiveEngine.key, _LiveEngine.eng = key, eng
    genout, stats = eng.run()
    stats["maps_reused"] = False
    return genout, stats


class Session:
    """Several transmit events on one medium without the disk: the first `run_solver(solver, session=s)` uploads
    the maps, later calls (new `Solver` objects over the SAME grid, medium and sensor, different `Source`) only
    replace the source list."""

    def __init__(self):
        self.eng = None
        self.shape = None

    def close(self):
        if self.eng is not None:
            self.eng.close()
            self.eng = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


class Launcher:
    """Drop-in for `fullwave.solver.launcher.Launcher`, backed by libfw25.so."""

    def __init__(self, path_fullwave_simulation_bin: Path | None = None, *, is_3d: bool = False,
                 use_gpu: bool = True, cuda_device_id=None) -> None:
        self._path_fullwave_simulation_bin = path_fullwave_simulation_bin   # used only by runs handed to the reference
        self.is_3d = is_3d
        self.use_gpu = use_gpu
        self.cuda_device_id = parse_cuda_device_id(cuda_device_id)
        self.last_stats: dict | None = None
        engine.lib()                                                        # fail at construction if not built

    def run(self, simulation_dir: Path, *, load_results: bool = True):
        simulation_dir = Path(simulation_dir).absolute()
        if not self.use_gpu:
            raise NotImplementedError("Currently, only GPU version is supported.")
        if _is_exponential_attenuation_dir(simulation_dir):
            # not this engine's physics (out of scope): hand the run to the reference's own launcher
            # and binary, which is what `install()` promises for such solvers
            ref_cls = _reference_launcher_cls
            if ref_cls is None:
                raise NotImplementedError("exponential-attenuation simulations run on the reference engine only; "
                                          "call fullwave25_b200.launcher.install() or use fullwave's Launcher")
            ref = ref_cls(self._path_fullwave_simulation_bin, is_3d=self.is_3d, use_gpu=self.use_gpu,
                          cuda_device_id=[int(v) for v in self.cuda_device_id.split(",")]
                          if "," in self.cuda_device_id else int(self.cuda_device_id))
            return ref.run(simulation_dir, load_results=load_results)
        log = simulation_dir / "fw2_execution.log"
        t0 = time.time()
        try:
            if bool(self.is_3d) != (simulation_dir / "nZ.dat").exists():
                raise ValueError(f"launcher is_3d={self.is_3d} but the directory holds a "
                                 f"{3 if (simulation_dir / 'nZ.dat').exists() else 2}D problem")
            genout, stats = _run_dat_dir(simulation_dir, device_ids_of(self.cuda_device_id))
        except Exception as e:  # noqa: BLE001
            log.write_text(f"fw25 engine failed: {type(e).__name__}: {e}\n")
            msg = ("Simulation failed. please check the simulation log file for more information.\n"
                   f"The log file is located at:\n{log}")
            logger.exception(msg)
            raise _sim_error(msg) from e
        self.last_stats = stats
        genout.tofile(simulation_dir / "genout.dat")
        log.write_text(f"fw25 engine (libfw25.so, sm_100a)\n{stats}\nSimulation completed in {time.time() - t0:.2e} s\n")
        if not load_results:
            return simulation_dir / "genout.dat"
        flat = genout.reshape(-1)
        if np.isnan(flat).any():
            logger.warning("The simulation contains NaN values. Check the simulation domains or PML settings.")
        if np.isinf(flat).any():
            logger.warning("The simulation contains Inf values. Check the simulation domains or PML settings.")
        return flat


class _StaticSession:
    """The live engine of a static-map transmit sequence run through the patched `Solver.run` (install(in_memory=True))."""
    session: "Session | None" = None
    medium_ref = None

    @classmethod
    def release(cls):
        if cls.session is not None:
            cls.session.close()
        cls.session, cls.medium_ref = None, None


def install(fullwave_module=None, *, in_memory: bool = False, maps: str = "host"):
    """Make the reference's `Solver` use this engine.  Returns an `uninstall` callable.

    Default: replaces `fullwave.solver.solver.Launcher` (the name `Solver.__init__` instantiates, solver.py:510-515):
    `Solver.run` still writes its .dat directory and reads genout.dat back, only the launch is in-process.

    in_memory=True additionally replaces `Solver.run` (solver.py:620-778) itself, same signature and return value:
    nothing touches the disk, and with maps="device" `PMLBuilder.run` is replaced by the GPU map builder as well
    (`run_solver`).  `is_static_map=True` keeps the engine and its maps alive between transmit events exactly like
    upstream keeps the .dat files: `recalculate_pml=True` starts a sequence, `False` reuses it.  Calls that need the
    directory (`load_results=False`, exponential attenuation) fall through to the original method."""
    import importlib
    import weakref
    if maps not in ("host", "device"):
        raise ValueError('maps must be "host" or "device"')
    global _raise_cls, _reference_launcher_cls
    sol = importlib.import_module("fullwave.solver.solver")
    lau = importlib.import_module("fullwave.solver.launcher")
    saved = (sol.Launcher, lau.Launcher, sol.Solver.run, sol.PMLBuilder)
    saved_err = (_raise_cls, _reference_launcher_cls)
    if lau.Launcher is not Launcher:
        _reference_launcher_cls = lau.Launcher
    ref_err = getattr(lau, "SimulationError", None)
    if isinstance(ref_err, type) and not issubclass(SimulationError, ref_err):
        _raise_cls = type("SimulationError", (SimulationError, ref_err), {})
    sol.Launcher = Launcher
    lau.Launcher = Launcher

    if in_memory:
        original_run = sol.Solver.run

        def run(self, simulation_dir_name="txrx_0", *, is_static_map=False, recalculate_pml=True,
                record_whole_domain=False, sampling_modulus_time_whole_domain=1, load_results=True):
            if not load_results or getattr(self, "use_exponential_attenuation", False):
                return original_run(self, simulation_dir_name, is_static_map=is_static_map,
                                    recalculate_pml=recalculate_pml, record_whole_domain=record_whole_domain,
                                    sampling_modulus_time_whole_domain=sampling_modulus_time_whole_domain,
                                    load_results=load_results)
            session = None
            if is_static_map:
                med = self.pml_builder.medium_org
                same = (_StaticSession.session is not None and _StaticSession.medium_ref is not None
                        and _StaticSession.medium_ref() is med)
                if recalculate_pml or not same:
                    _StaticSession.release()
                    _StaticSession.session, _StaticSession.medium_ref = Session(), weakref.ref(med)
                session = _StaticSession.session
            return run_solver(self, record_whole_domain=record_whole_domain,
                              sampling_modulus_time_whole_domain=sampling_modulus_time_whole_domain,
                              maps=maps, session=session)

        sol.Solver.run = run
        if maps == "device":       # `Solver.__init__` then no longer pads anything on the host (solver.py:527-536)
            sol.PMLBuilder = lazy_pml_builder_class(sol.PMLBuilder)

    def uninstall():
        global _raise_cls, _reference_launcher_cls
        sol.Launcher, lau.Launcher, sol.Solver.run, sol.PMLBuilder = saved
        _raise_cls, _reference_launcher_cls = saved_err
        _StaticSession.release()
    return uninstall


def lean_lists(pmlb):
    """(source, sensor, air coordinates) on the EXTENDED grid straight from the user-grid objects: the reference pads
    the masks with zeros (`_extend_map_for_pml(..., fill_edge=False)`, pml_builder.py:243-262) and lists the non-zero
    cells in row-major order (utils/coordinates.py:28-53), which is the original list shifted by the boundary width."""
    from types import SimpleNamespace
    nb = int(pmlb.num_boundary_points)
    src, sen, med = pmlb.source_org, pmlb.sensor_org, pmlb.medium_org
    source = SimpleNamespace(incoords=np.asarray(src.incoords) + nb, icmat=src.icmat)
    sensor = SimpleNamespace(outcoords=np.asarray(sen.outcoords) + nb,
                             sampling_modulus_time=sen.sampling_modulus_time)
    air = np.asarray(med.air_map)
    icczero = (np.stack(np.nonzero(air != 0), axis=1) + nb) if air.any() else np.zeros((0, air.ndim), np.int64)
    return source, sensor, icczero


def lazy_pml_builder_class(PMLBuilder):
    """A `PMLBuilder` with the same constructor and attributes whose expensive part -- padding the medium, source and
    sensor to the extended grid in `__init__` (pml_builder.py:222-262; 7 s for 120^3) -- happens only when something
    asks for `extended_medium / extended_source / extended_sensor / pml_mask_*` or calls `run()`.  The GPU map builder
    needs none of them (`MediumSpec.from_pml_builder` reads the user-grid medium, `lean_lists` the user-grid masks)."""
    import fullwave

    class LazyPMLBuilder(PMLBuilder):
        def __init__(self, grid, medium, source, sensor, *, m_spatial_order=8, n_pml_layer=40, n_transition_layer=40,
                     use_isotropic_relaxation=False):
            self._ctor = (grid, medium, source, sensor, dict(
                m_spatial_order=m_spatial_order, n_pml_layer=n_pml_layer, n_transition_layer=n_transition_layer,
                use_isotropic_relaxation=use_isotropic_relaxation))
            self._full = None
            self.grid_org, self.medium_org, self.source_org, self.sensor_org = grid, medium, source, sensor
            self.is_3d = grid.is_3d
            self.use_isotropic_relaxation = use_isotropic_relaxation
            self.m_spatial_order, self.n_pml_layer = m_spatial_order, n_pml_layer
            self.n_transition_layer = n_transition_layer
            nb = self.num_boundary_points
            deltas = (grid.dx, grid.dy, grid.dz) if self.is_3d else (grid.dx, grid.dy)
            domain_size = tuple((n + 2 * nb) * d for n, d in zip(np.asarray(medium.sound_speed).shape, deltas))
            self.extended_grid = fullwave.Grid(domain_size=domain_size, f0=grid.f0, duration=grid.duration,
                                               c0=grid.c0, ppw=grid.ppw, cfl=grid.cfl)
            self.pml_layer_m = self.extended_grid.dx * n_pml_layer
            self.transition_layer_m = self.extended_grid.dx * n_transition_layer
            self.n_polynomial = 2
            self.theoritical_reflection_coefficient = 10 ** (-30)
            if self.n_pml_layer == 0:
                self.n_transition_layer = 0

        def _materialise(self):
            if self._full is None:
                g, m, s, r, kw = self._ctor
                self._full = PMLBuilder(g, m, s, r, **kw)
            return self._full

        extended_medium = property(lambda self: self._materialise().extended_medium)
        extended_source = property(lambda self: self._materialise().extended_source)
        extended_sensor = property(lambda self: self._materialise().extended_sensor)
        pml_mask_x = property(lambda self: self._materialise().pml_mask_x)
        pml_mask_y = property(lambda self: self._materialise().pml_mask_y)
        pml_mask_z = property(lambda self: self._materialise().pml_mask_z)

        def run(self, *, use_pml=True):
            return self._materialise().run(use_pml=use_pml)

    LazyPMLBuilder.__name__ = "PMLBuilder"
    return LazyPMLBuilder


def _sensor_and_box(solver, record_whole_domain: bool, modulus: int, lean: bool = False):
    """(sensor, out_box).  `record_whole_domain` (solver.py:709-731) upstream builds a `Sensor` whose mask is the whole
    extended grid -- one coordinate row per grid point.  Here it is a box (fw25.h, out_box): no mask, no coordinate
    list on the host, no index list on the device; the frames come back in the same row-major order."""
    if not record_whole_domain:
        return (lean_lists(solver.pml_builder)[1] if lean else solver.pml_builder.extended_sensor), None
    from types import SimpleNamespace
    eg = solver.pml_builder.extended_grid
    shape = (int(eg.nx), int(eg.ny), int(eg.nz)) if solver.is_3d else (int(eg.nx), int(eg.ny))
    return SimpleNamespace(sampling_modulus_time=int(modulus), outcoords=None), (0,) * len(shape) + shape


def _run_solver_device_maps(solver, sensor, out_box, device: int, session: Session | None):
    """The engine input without `PMLBuilder.run` and without any extended-grid array on the host: the thirteen
    coefficient maps and dcmap are generated on the GPU from the solver's ORIGINAL (user-grid) medium
    (`mapgen.MapSet`, fw25_mapgen) and adopted by the engine in place."""
    from . import mapgen
    pmlb = solver.pml_builder
    t0 = time.perf_counter()
    spec = mapgen.MediumSpec.from_pml_builder(pmlb, use_pml=solver.use_pml, dcmap_full3d=False)
    ms = mapgen.MapSet(spec, device=device)
    if ms.invalid_count:
        logger.warning("Warning: Some attenuation values correspond to invalid relaxation parameters. "
                       "This is due to the limitations of the precomputed lookup table. "
                       "Please change the attenuation values.\nNumber of invalid points: %d.", ms.invalid_count)
    source, _, icczero = lean_lists(pmlb)
    if not getattr(solver, "use_isotropic_relaxation", True):
        icczero = icczero[:0]      # the reference's anisotropic binaries have no inject_source_zero kernel (fw25.h)
    pb = Problem.for_device_maps(ms, pmlb.extended_grid, source, sensor, out_box=out_box, icczero=icczero)
    eng = engine.Engine(pb, device=device, device_maps=ms.device_maps())
    setup_s = time.perf_counter() - t0
    try:
        genout, stats = eng.run()
    finally:
        if session is None:
            eng.close()
            ms.close()
    if session is not None:
        session.eng, session.shape = eng, pb.shape
    stats.update(mapgen_upload_ms=ms.upload_ms, mapgen_kernel_ms=ms.kernel_ms, host_setup_s=setup_s, maps="device")
    return genout, stats, pb


def _run_solver_device_maps_multi(solver, sensor, out_box, device_ids):
    """maps="device" on several GPUs: no `PMLBuilder.run`, no extended-grid array on the host, and no single device
    ever holds the whole grid -- `mapgen.run_medium(device_ids=...)` (fw25_run_medium_multi)."""
    from types import SimpleNamespace

    from . import mapgen
    pmlb = solver.pml_builder
    t0 = time.perf_counter()
    spec = mapgen.MediumSpec.from_pml_builder(pmlb, use_pml=solver.use_pml, dcmap_full3d=False)
    d_table, dmap, ndmap, _ = spec.stencil_tables()
    source, _, icczero = lean_lists(pmlb)
    if not getattr(solver, "use_isotropic_relaxation", True):
        icczero = icczero[:0]
    shape_only = SimpleNamespace(shape=spec.extended_shape, ndmap=ndmap, dmap=dmap, d_table=d_table)
    pb = Problem.for_device_maps(shape_only, pmlb.extended_grid, source, sensor, out_box=out_box, icczero=icczero)
    genout, stats = mapgen.run_medium(spec, pb, device_ids=device_ids)
    stats.update(host_setup_s=time.perf_counter() - t0 - stats["native_call_ms"] / 1e3, maps="device")
    return genout, stats, pb


def _remember(stats: dict) -> None:
    global last_run_stats
    last_run_stats = dict(stats)


def run_solver(solver, *, record_whole_domain: bool = False, sampling_modulus_time_whole_domain: int = 1,
               cuda_device_id=None, return_stats: bool = False, session: Session | None = None,
               maps: str = "host"):
    """`Solver.run` without the disk.  maps="host": PMLBuilder stays the reference's own Python (solver.py:694), the
    engine input is assembled in memory (what InputFileWriter would have written, input_file_writer.py:563-881) --
    bit-identical to the reference binary.  maps="device": the coefficient maps are built on the GPU from the
    user-grid medium instead (fw25_mapgen; a / b within one float32 ulp of the reference's files, everything else
    bit-identical; a device list makes every GPU build its own x-slab of the maps).  The sensor traces come back as [n_sensors, n_frames] exactly like
    `Solver._reshape_sensor_data`."""
    if maps not in ("host", "device"):
        raise ValueError('maps must be "host" or "device"')
    ids = device_ids_of(cuda_device_id if cuda_device_id is not None else getattr(solver, "cuda_device_id", None))
    if maps == "device" and len(ids) == 1 and not (session is not None and session.eng is not None):
        sensor, out_box = _sensor_and_box(solver, record_whole_domain, sampling_modulus_time_whole_domain, lean=True)
        if session is None:
            release()             # an engine kept for static-map reuse would only be in the way
        try:
            genout, stats, pb = _run_solver_device_maps(solver, sensor, out_box, ids[0], session)
        except (engine.EngineError, ValueError) as e:
            raise _sim_error(str(e)) from e
        result = genout.reshape(-1, pb.ncoordsout).T
        _remember(stats)
        return (result, stats) if return_stats else result
    if maps == "device" and len(ids) > 1 and session is None:
        # several GPUs: every device builds its own x-slab of the maps from the user-grid medium (fw25_run_medium_multi)
        from . import mapgen
        sensor, out_box = _sensor_and_box(solver, record_whole_domain, sampling_modulus_time_whole_domain, lean=True)
        release()
        try:
            genout, stats, pb = _run_solver_device_maps_multi(solver, sensor, out_box, ids)
        except (engine.EngineError, ValueError) as e:
            raise _sim_error(str(e)) from e
        result = genout.reshape(-1, pb.ncoordsout).T
        _remember(stats)
        return (result, stats) if return_stats else result
    if session is not None and session.eng is not None:      # next transmit event: only the sources change
        src = lean_lists(solver.pml_builder)[0]
        eg = solver.pml_builder.extended_grid
        shape = (eg.nx, eg.ny, eg.nz) if solver.is_3d else (eg.nx, eg.ny)
        if tuple(shape) != tuple(session.shape):
            raise ValueError(f"session holds a {session.shape} grid, this solver has {shape}")
        # only the SOURCE may change between the events of a session: sensors, recording period and (by contract)
        # the medium stay on the device from the first event
        sensor, out_box = _sensor_and_box(solver, record_whole_domain, sampling_modulus_time_whole_domain, lean=True)
        held = session.eng.pb
        n_new = int(np.prod(np.asarray(out_box).reshape(2, -1)[1] - np.asarray(out_box).reshape(2, -1)[0])) \
            if out_box is not None else len(np.asarray(sensor.outcoords))
        if (n_new != held.ncoordsout or int(sensor.sampling_modulus_time) != held.modT or
                (None if out_box is None else tuple(int(v) for v in out_box)) != held.out_box or
                (out_box is None and not np.array_equal(np.asarray(sensor.outcoords, np.int32).reshape(-1, held.ndim),
                                                        held.outc))):
            raise ValueError("session: the sensor, its sampling modulus or the recording mode differs from the event "
                             "that created the session; close it and start a new one")
        try:
            session.eng.reset(np.asarray(src.incoords), np.asarray(src.icmat), int(eg.nt))
            genout, stats = session.eng.run()
        except (engine.EngineError, ValueError) as e:
            raise _sim_error(str(e)) from e
        result = genout.reshape(-1, session.eng.pb.ncoordsout).T
        _remember(stats)
        return (result, stats) if return_stats else result
    extended_medium = solver.pml_builder.run(use_pml=solver.use_pml)
    sensor, out_box = _sensor_and_box(solver, record_whole_domain, sampling_modulus_time_whole_domain)
    pb = Problem.from_fullwave_objects(solver.pml_builder.extended_grid, extended_medium,
                                       solver.pml_builder.extended_source, sensor, out_box=out_box)
    try:
        if session is not None and len(ids) == 1:
            session.eng = engine.Engine(pb, device=ids[0])
            session.shape = pb.shape
            genout, stats = session.eng.run()
        else:
            release()
            genout, stats = engine.run(pb, device_ids=ids)
    except (engine.EngineError, ValueError) as e:
        raise _sim_error(str(e)) from e
    result = genout.reshape(-1, pb.ncoordsout).T          # solver.py:600-618
    _remember(stats)
    return (result, stats) if return_stats else result
