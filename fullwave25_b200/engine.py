"""ctypes binding of libfw25.so (include/fw25.h) -- the only way Python reaches the CUDA engine.

There is NO fallback: if the shared library is missing or no CUDA device answers, every entry point
raises.  (The reference behaves the same way: without its GPU binary `Launcher.run` raises,
/root/reference/fullwave/solver/launcher.py:191-194, :221-241.)
"""

from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

from .problem import MAP_NAMES, Problem

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libfw25.so"
if os.environ.get("FW25_LIB"):          # tuning sweeps only (tools/build_variants.py): another build of the same library
    LIB_PATH = Path(os.environ["FW25_LIB"]).resolve()
_F = C.POINTER(C.c_float)
_I = C.POINTER(C.c_int32)


class EngineError(RuntimeError):
    """The native engine reported a failure (non-zero return code + fw25_last_error())."""


class CAniso(C.Structure):
    """fw25_aniso: per-axis maps of the anisotropic file set; [axis] or [axis][nu]."""
    _fields_ = [("kappa_vel", C.c_void_p * 3), ("a_vel", (C.c_void_p * 2) * 3), ("b_vel", (C.c_void_p * 2) * 3),
                ("kappa_prs", C.c_void_p * 3), ("a_prs", (C.c_void_p * 2) * 3), ("b_prs", (C.c_void_p * 2) * 3)]


class CProblem(C.Structure):
    _fields_ = [
        ("ndim", C.c_int32), ("nX", C.c_int32), ("nY", C.c_int32), ("nZ", C.c_int32),
        ("nT", C.c_int32), ("nTic", C.c_int32), ("modT", C.c_int32), ("ndmap", C.c_int32),
        ("dX", C.c_float), ("dT", C.c_float),
        ("rho", C.c_void_p), ("K", C.c_void_p), ("beta", C.c_void_p),
        ("kappax", C.c_void_p), ("kappau", C.c_void_p),
        ("apmlx1", C.c_void_p), ("bpmlx1", C.c_void_p), ("apmlx2", C.c_void_p), ("bpmlx2", C.c_void_p),
        ("apmlu1", C.c_void_p), ("bpmlu1", C.c_void_p), ("apmlu2", C.c_void_p), ("bpmlu2", C.c_void_p),
        ("dmap", C.c_void_p), ("dcmap", C.c_void_p),
        ("ncoords", C.c_int32), ("icc", C.c_void_p), ("icmat", C.c_void_p),
        ("ncoordsout", C.c_int32), ("outc", C.c_void_p),
        ("ncoordszero", C.c_int32), ("icczero", C.c_void_p),
        ("maps_on_device", C.c_int32), ("map_pitch", C.c_int32), ("dcmap_full3d", C.c_int32),
        ("ext_p", C.c_void_p), ("ext_u", C.c_void_p), ("ext_v", C.c_void_p), ("ext_w", C.c_void_p),
        ("aniso", C.POINTER(CAniso)),
        ("out_box", C.c_void_p),
    ]


class CSlab(C.Structure):
    _fields_ = [("nX_global", C.c_int32), ("gx0", C.c_int32), ("own_lo", C.c_int32), ("own_hi", C.c_int32)]


class CStats(C.Structure):
    _fields_ = [("setup_ms", C.c_double), ("loop_ms", C.c_double), ("d2h_ms", C.c_double),
                ("kernel_launches", C.c_int64), ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
                ("point_updates", C.c_int64), ("halo_bytes", C.c_int64), ("n_devices", C.c_int32),
                ("skewed_steps", C.c_int32)]


_lib = None

_SIGS = {
    "fw25_run": (C.c_int, [C.POINTER(CProblem), _I, C.c_int32, _F, C.c_size_t, C.POINTER(CStats)]),
    "fw25_create": (C.c_int, [C.POINTER(CProblem), C.POINTER(CSlab), C.c_int32, C.POINTER(C.c_void_p)]),
    "fw25_destroy": (None, [C.c_void_p]),
    "fw25_inject": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    "fw25_sweep_u": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "fw25_sweep_p": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "fw25_record": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    "fw25_step": (C.c_int, [C.c_void_p, C.c_int32]),
    "fw25_sync": (C.c_int, [C.c_void_p]),
    "fw25_step_timed": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_double)]),
    "fw25_n_local_sensors": (C.c_int32, [C.c_void_p]),
    "fw25_local_sensor_ids": (C.c_int, [C.c_void_p, _I]),
    "fw25_read_frames": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, _F]),
    "fw25_read_field": (C.c_int, [C.c_void_p, C.c_char_p, _F]),
    "fw25_field_ptr": (C.c_void_p, [C.c_void_p, C.c_char_p]),
    "fw25_pitch": (C.c_int32, [C.c_int32]),
    "fw25_current_step": (C.c_int32, [C.c_void_p]),
    "fw25_launch_count": (C.c_int64, [C.c_void_p]),
    "fw25_set_kernel_variant": (C.c_int, [C.c_void_p, C.c_int32]),
    "fw25_reset": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, _I, _F]),
    "fw25_run_engine": (C.c_int, [C.c_void_p, _F, C.c_size_t, C.POINTER(CStats)]),
    "fw25_device_count": (C.c_int32, []),
    "fw25_last_error": (C.c_char_p, []),
    "fw25_abi_version": (C.c_int32, []),
}

EXPORTS = tuple(_SIGS)


def lib() -> C.CDLL:
    """Load libfw25.so (built in-tree by fullwave25_b200.build); raises if it is absent."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise EngineError(
                f"{LIB_PATH} is missing: build it with `python -m fullwave25_b200.build` "
                "(there is no CPU or PyTorch fallback for the engine)")
        _lib = C.CDLL(str(LIB_PATH))
        for name, (res, args) in _SIGS.items():
            fn = getattr(_lib, name)
            fn.restype = res
            fn.argtypes = args
    return _lib


def _check(rc: int) -> None:
    if rc != 0:
        raise EngineError(f"fw25 engine error {rc}: {lib().fw25_last_error().decode(errors='replace')}")


def _ptr(a) -> int:
    """Host numpy array or device pointer (int) -> address."""
    if isinstance(a, (int, np.integer)):
        return int(a)
    return a.ctypes.data


def marshal(pb: Problem, *, device_maps: dict | None = None, ext_state: dict | None = None):
    """Problem -> (CProblem, keepalive).  device_maps: {name: device address} for the 13 maps + dcmap
    (maps_on_device = 1).  ext_state: {"p","u","v","w": device address}."""
    s = CProblem()
    s.ndim, s.nX, s.nY, s.nZ = pb.ndim, pb.nX, pb.nY, pb.nZ if pb.ndim == 3 else 1
    s.nT, s.nTic, s.modT, s.ndmap = pb.nT, pb.nTic, pb.modT, pb.ndmap
    s.dX, s.dT = pb.dX, pb.dT
    s.dcmap_full3d = int(bool(pb.dcmap_full3d))
    keep = [pb]
    if device_maps is not None:
        for name in MAP_NAMES + ("dcmap",):
            setattr(s, name, int(device_maps[name]))
        s.maps_on_device = 1
        s.map_pitch = int(device_maps.get("pitch", 0))
    else:
        for name in MAP_NAMES + ("dcmap",):
            a = getattr(pb, name)
            assert a.flags.c_contiguous and a.dtype == (np.int32 if name == "dcmap" else np.float32), name
            setattr(s, name, a.ctypes.data)
    s.dmap = pb.dmap.ctypes.data
    s.ncoords, s.icc, s.icmat = pb.ncoords, pb.icc.ctypes.data, pb.icmat.ctypes.data
    s.ncoordsout, s.outc = pb.ncoordsout, pb.outc.ctypes.data
    if pb.out_box is not None:          # box sensors: no coordinate list at all
        box = np.asarray(pb.out_box, np.int32).reshape(2 * pb.ndim)
        keep.append(box)
        s.out_box, s.outc = box.ctypes.data, None
    s.ncoordszero, s.icczero = pb.ncoordszero, pb.icczero.ctypes.data
    if ext_state:
        s.ext_p, s.ext_u = ext_state.get("p"), ext_state.get("u")
        s.ext_v, s.ext_w = ext_state.get("v"), ext_state.get("w")
    dev_aniso = (device_maps or {}).get("aniso")
    if pb.aniso is not None or dev_aniso:
        # anisotropic file set: host arrays (pb.aniso) or, with device-resident maps, {stem: device address} under
        # device_maps["aniso"] in the same [nX][nY][pitch] layout
        if device_maps is not None and not dev_aniso:
            raise EngineError("device-resident maps of an anisotropic problem need device_maps['aniso']")
        addr = (lambda stem: int(dev_aniso[stem])) if dev_aniso else (lambda stem: pb.aniso[stem].ctypes.data)
        an = CAniso()
        vel = ("x", "y", "z")[: pb.ndim]
        prs = ("u", "w") if pb.ndim == 2 else ("u", "v", "w")
        for ax in range(pb.ndim):
            an.kappa_vel[ax] = addr("kappa" + vel[ax])
            an.kappa_prs[ax] = addr("kappa" + prs[ax])
            for nu in range(2):
                an.a_vel[ax][nu] = addr(f"apml{vel[ax]}{nu + 1}")
                an.b_vel[ax][nu] = addr(f"bpml{vel[ax]}{nu + 1}")
                an.a_prs[ax][nu] = addr(f"apml{prs[ax]}{nu + 1}")
                an.b_prs[ax][nu] = addr(f"bpml{prs[ax]}{nu + 1}")
        keep.append(an)
        s.aniso = C.pointer(an)
    return s, keep


def run(pb: Problem, device_ids=(0,)) -> tuple[np.ndarray, dict]:
    """Whole job with HOST buffers through fw25_run (C-ABI).  Several devices (the reference's
    `cuda_device_id=[0, 1, ...]`): x-slabs driven from this process by the native runner with peer-to-peer
    halo copies.  Returns (genout [n_frames, ncoordsout], stats)."""
    pb.normalise()
    s, keep = marshal(pb)
    genout = np.zeros((pb.n_frames, pb.ncoordsout), np.float32)
    ids = np.asarray(list(device_ids), np.int32)
    st = CStats()
    _check(lib().fw25_run(C.byref(s), ids.ctypes.data_as(_I), len(ids), genout.ctypes.data_as(_F),
                          genout.size, C.byref(st)))
    del keep
    return genout, {f: getattr(st, f) for f, _ in CStats._fields_}


class Engine:
    """Handle API: create once, step, read fields / frames (used by tests and the slab driver)."""

    def __init__(self, pb: Problem, device: int = 0, slab: tuple[int, int, int, int] | None = None,
                 device_maps: dict | None = None, ext_state: dict | None = None, variant: int = 0):
        if device_maps is None or pb.rho is None:
            pb.normalise()
        self.pb = pb
        s, self._keep = marshal(pb, device_maps=device_maps, ext_state=ext_state)
        self._keep.append((device_maps, ext_state))
        h = C.c_void_p()
        cs = None
        if slab is not None:
            cs = CSlab(*slab)
        _check(lib().fw25_create(C.byref(s), C.byref(cs) if cs is not None else None, device, C.byref(h)))
        self._h = h
        if variant:
            self.set_variant(variant)

    def close(self) -> None:
        if getattr(self, "_h", None) and _lib is not None:   # (_lib is None during interpreter shutdown)
            _lib.fw25_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def reset(self, icc: np.ndarray, icmat: np.ndarray, nT: int | None = None) -> None:
        """Next transmit event on the same medium (fw25_reset): zero wave field, t = 0, new sources
        icc int32 [ncoords, ndim], icmat float32 [ncoords, nTic]; maps and sensors stay on the device."""
        icc = np.ascontiguousarray(icc, np.int32).reshape(-1, self.pb.ndim)
        icmat = np.ascontiguousarray(icmat, np.float32).reshape(len(icc), -1)
        nT = self.pb.nT if nT is None else int(nT)
        _check(lib().fw25_reset(self._h, nT, icmat.shape[1], len(icc), icc.ctypes.data_as(_I), icmat.ctypes.data_as(_F)))
        self._nT = nT

    def run(self) -> tuple[np.ndarray, dict]:
        """Steps [t, nT) with the frames read back (fw25_run_engine) -> (genout [n_frames, ncoordsout], stats)."""
        nT = getattr(self, "_nT", self.pb.nT)
        n_frames = -(-nT // self.pb.modT) if nT > 0 else 0
        genout = np.zeros((n_frames, self.pb.ncoordsout), np.float32)
        st = CStats()
        _check(lib().fw25_run_engine(self._h, genout.ctypes.data_as(_F), genout.size, C.byref(st)))
        return genout, {f: getattr(st, f) for f, _ in CStats._fields_}

    def set_variant(self, v: int) -> None:
        _check(lib().fw25_set_kernel_variant(self._h, v))

    def step(self, n: int = 1) -> None:
        _check(lib().fw25_step(self._h, n))

    def step_timed(self, n: int, detail: bool = False) -> dict:
        """n steps timed with CUDA events on the engine's stream -> {total_ms, sweep_u_ms, sweep_p_ms, other_ms}."""
        out = (C.c_double * 4)()
        _check(lib().fw25_step_timed(self._h, n, int(detail), out))
        return {"total_ms": out[0], "sweep_u_ms": out[1], "sweep_p_ms": out[2], "other_ms": out[3]}

    def sync(self) -> None:
        _check(lib().fw25_sync(self._h))

    def inject(self, t: int, stream: int = 0) -> None:
        _check(lib().fw25_inject(self._h, t, stream or None))

    def sweep_u(self, lo: int, hi: int, stream: int = 0) -> None:
        _check(lib().fw25_sweep_u(self._h, lo, hi, stream or None))

    def sweep_p(self, lo: int, hi: int, stream: int = 0) -> None:
        _check(lib().fw25_sweep_p(self._h, lo, hi, stream or None))

    def record(self, frame: int, stream: int = 0) -> None:
        _check(lib().fw25_record(self._h, frame, stream or None))

    @property
    def t(self) -> int:
        return lib().fw25_current_step(self._h)

    @property
    def launches(self) -> int:
        return lib().fw25_launch_count(self._h)

    @property
    def n_local_sensors(self) -> int:
        return lib().fw25_n_local_sensors(self._h)

    def local_sensor_ids(self) -> np.ndarray:
        ids = np.zeros(self.n_local_sensors, np.int32)
        _check(lib().fw25_local_sensor_ids(self._h, ids.ctypes.data_as(_I)))
        return ids

    def read_frames(self, f0: int, f1: int) -> np.ndarray:
        out = np.zeros((f1 - f0, self.n_local_sensors), np.float32)
        _check(lib().fw25_read_frames(self._h, f0, f1, out.ctypes.data_as(_F)))
        return out

    def field(self, name: str) -> np.ndarray:
        pb = self.pb
        out = np.zeros(pb.shape, np.float32)
        _check(lib().fw25_read_field(self._h, name.encode(), out.ctypes.data_as(_F)))
        return out

    def field_ptr(self, name: str) -> int:
        return lib().fw25_field_ptr(self._h, name.encode())
