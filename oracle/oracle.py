"""ctypes loader for the CPU oracle (oracle/fw25_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module, and only as the checker -- never the product package (fullwave25_b200/).

The oracle restates the arithmetic of the reference's binary-only engine (see fw25_oracle.h for the
PTX/SASS ranges).  It takes any object exposing the fields of the engine's file protocol
(/root/reference/fullwave/solver/input_file_writer.py:563-881): ndim, nX, nY, nZ, nT, nTic, modT,
ndmap, dX, dT, rho, K, beta, kappax, kappau, apml{x,u}{1,2}, bpml{x,u}{1,2}, dmap, dcmap, icc, icmat,
outc, icczero.
"""

from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_F = C.POINTER(C.c_float)
_I = C.POINTER(C.c_int32)


class _Problem(C.Structure):
    _fields_ = [
        ("ndim", C.c_int32), ("nX", C.c_int32), ("nY", C.c_int32), ("nZ", C.c_int32),
        ("nT", C.c_int32), ("nTic", C.c_int32), ("modT", C.c_int32), ("ndmap", C.c_int32),
        ("dX", C.c_float), ("dT", C.c_float),
        ("rho", _F), ("K", _F), ("beta", _F), ("kappax", _F), ("kappau", _F),
        ("apmlx1", _F), ("bpmlx1", _F), ("apmlx2", _F), ("bpmlx2", _F),
        ("apmlu1", _F), ("bpmlu1", _F), ("apmlu2", _F), ("bpmlu2", _F),
        ("dmap", _F), ("dcmap", _I),
        ("ncoords", C.c_int32), ("icc", _I), ("icmat", _F),
        ("ncoordsout", C.c_int32), ("outc", _I),
        ("ncoordszero", C.c_int32), ("icczero", _I),
        ("dcmap_full3d", C.c_int32), ("nX_dcmap", C.c_int32),
        ("aniso", C.c_int32),
        ("kappa_vel", _F * 3), ("a_vel", (_F * 2) * 3), ("b_vel", (_F * 2) * 3),
        ("kappa_prs", _F * 3), ("a_prs", (_F * 2) * 3), ("b_prs", (_F * 2) * 3),
    ]


class _State(C.Structure):
    _fields_ = [("p", _F), ("u", _F), ("v", _F), ("w", _F), ("psi", _F * 6), ("phi", _F * 6)]


_MAPS = ("rho", "K", "beta", "kappax", "kappau", "apmlx1", "bpmlx1", "apmlx2", "bpmlx2",
         "apmlu1", "bpmlu1", "apmlu2", "bpmlu2")

_lib = None


def build(force: bool = False) -> None:
    """Compile the oracle with oracle/Makefile (gcc only)."""
    # make decides staleness (sources newer than the .so) -- cheap when up to date
    subprocess.run(["make", "-C", str(_HERE), "-B" if force else "-s"], check=True, stdout=subprocess.DEVNULL)


def _cpu_has_fma() -> bool:
    try:
        for line in Path("/proc/cpuinfo").read_text().splitlines():
            if line.startswith("flags"):
                fl = line.split(":")[1].split()
                return "fma" in fl and "avx2" in fl
    except OSError:
        pass
    return False


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        name = "libfw25_oracle_fma.so" if _cpu_has_fma() else "libfw25_oracle.so"
        _lib = C.CDLL(str(_HERE / name))
        _lib.fw25o_run.argtypes = [C.POINTER(_Problem), _F, _F]
        _lib.fw25o_run.restype = C.c_int
        for fn in ("fw25o_sweep_u", "fw25o_sweep_p"):
            getattr(_lib, fn).argtypes = [C.POINTER(_Problem), C.POINTER(_State), C.c_int, C.c_int]
            getattr(_lib, fn).restype = None
        _lib.fw25o_inject.argtypes = [C.POINTER(_Problem), C.POINTER(_State), C.c_int]
        _lib.fw25o_inject.restype = None
        _lib.fw25o_record.argtypes = [C.POINTER(_Problem), C.POINTER(_State), _F]
        _lib.fw25o_record.restype = None
    return _lib


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int32)


def _marshal(pb):
    """Returns (struct, keepalive list)."""
    keep = []
    s = _Problem()
    s.ndim, s.nX, s.nY = int(pb.ndim), int(pb.nX), int(pb.nY)
    s.nZ = int(pb.nZ) if pb.ndim == 3 else 1
    s.nT, s.nTic, s.modT, s.ndmap = int(pb.nT), int(pb.nTic), int(pb.modT), int(pb.ndmap)
    s.dX, s.dT = float(pb.dX), float(pb.dT)
    s.dcmap_full3d = int(bool(getattr(pb, "dcmap_full3d", False)))
    s.nX_dcmap = s.nX
    n = s.nX * s.nY * s.nZ
    for name in _MAPS:
        a = _f32(getattr(pb, name)).reshape(-1)
        assert a.size == n, (name, a.size, n)
        keep.append(a)
        setattr(s, name, a.ctypes.data_as(_F))
    aniso = getattr(pb, "aniso", None)
    if aniso:
        # file stems of the anisotropic protocol (input_file_writer.py:592-620) -> per-axis slots
        s.aniso = 1
        vel, prs = aniso_axis_names(s.ndim)

        def ptr(stem):
            a = _f32(aniso[stem]).reshape(-1)
            assert a.size == n, (stem, a.size, n)
            keep.append(a)
            return a.ctypes.data_as(_F)
        for ax in range(s.ndim):
            s.kappa_vel[ax] = ptr("kappa" + vel[ax])
            s.kappa_prs[ax] = ptr("kappa" + prs[ax])
            for nu in range(2):
                s.a_vel[ax][nu] = ptr(f"apml{vel[ax]}{nu + 1}")
                s.b_vel[ax][nu] = ptr(f"bpml{vel[ax]}{nu + 1}")
                s.a_prs[ax][nu] = ptr(f"apml{prs[ax]}{nu + 1}")
                s.b_prs[ax][nu] = ptr(f"bpml{prs[ax]}{nu + 1}")
    a = _f32(pb.dmap).reshape(-1)
    assert a.size == 18 * s.ndmap
    keep.append(a)
    s.dmap = a.ctypes.data_as(_F)
    a = _i32(pb.dcmap).reshape(-1)
    assert a.size == n
    keep.append(a)
    s.dcmap = a.ctypes.data_as(_I)
    for cnt, name in (("ncoords", "icc"), ("ncoordsout", "outc"), ("ncoordszero", "icczero")):
        a = _i32(getattr(pb, name)).reshape(-1, s.ndim)
        if name == "icczero" and aniso:
            a = a[:0]            # the anisotropic binaries have no inject_source_zero kernel
        keep.append(a)
        setattr(s, cnt, a.shape[0])
        setattr(s, name, a.ctypes.data_as(_I))
    a = _f32(pb.icmat).reshape(s.ncoords, s.nTic) if s.ncoords else np.zeros((0, max(s.nTic, 1)), np.float32)
    keep.append(a)
    s.icmat = a.ctypes.data_as(_F)
    return s, keep


def aniso_axis_names(ndim: int):
    """File-stem letters per axis (x, y[, z]) of the anisotropic protocol: (velocity sweep, pressure sweep)."""
    return (("x", "y", "z")[:ndim], ("u", "w") if ndim == 2 else ("u", "v", "w"))


def n_frames(pb) -> int:
    return -(-int(pb.nT) // int(pb.modT))


def run(pb, return_fields: bool = False):
    """Run nT steps from zero state.  Returns genout [n_frames, ncoordsout] (and dict of final
    p,u,v,w fields shaped like the grid when return_fields)."""
    s, keep = _marshal(pb)
    genout = np.zeros((n_frames(pb), s.ncoordsout), np.float32)
    n = s.nX * s.nY * s.nZ
    final = np.zeros((4, n), np.float32) if return_fields else None
    rc = lib().fw25o_run(C.byref(s), genout.ctypes.data_as(_F),
                         final.ctypes.data_as(_F) if return_fields else None)
    if rc != 0:
        raise MemoryError("oracle allocation failed")
    del keep
    if not return_fields:
        return genout
    shape = (s.nX, s.nY, s.nZ) if s.ndim == 3 else (s.nX, s.nY)
    return genout, {k: final[i].reshape(shape) for i, k in enumerate("puvw")}


class Stepper:
    """Step-by-step oracle with explicit state (used by the slab / halo tests and the CPU baseline)."""

    def __init__(self, pb):
        self.s, self._keep = _marshal(pb)
        n = self.s.nX * self.s.nY * self.s.nZ
        self.shape = (self.s.nX, self.s.nY, self.s.nZ) if self.s.ndim == 3 else (self.s.nX, self.s.nY)
        self.mem = np.zeros((16, n), np.float32)
        st = _State()
        st.p, st.u, st.v, st.w = (self.mem[i].ctypes.data_as(_F) for i in range(4))
        for k in range(6):
            st.psi[k] = self.mem[4 + k].ctypes.data_as(_F)
            st.phi[k] = self.mem[10 + k].ctypes.data_as(_F)
        self.st = st
        self.t = 0

    def field(self, name: str) -> np.ndarray:
        return self.mem["puvw".index(name)].reshape(self.shape)

    def inject(self, t=None):
        lib().fw25o_inject(C.byref(self.s), C.byref(self.st), self.t if t is None else t)

    def sweep_u(self, x_lo=0, x_hi=None):
        lib().fw25o_sweep_u(C.byref(self.s), C.byref(self.st), x_lo, self.s.nX if x_hi is None else x_hi)

    def sweep_p(self, x_lo=0, x_hi=None):
        lib().fw25o_sweep_p(C.byref(self.s), C.byref(self.st), x_lo, self.s.nX if x_hi is None else x_hi)

    def record(self) -> np.ndarray:
        out = np.zeros(self.s.ncoordsout, np.float32)
        lib().fw25o_record(C.byref(self.s), C.byref(self.st), out.ctypes.data_as(_F))
        return out

    def step(self):
        self.inject()
        self.sweep_u()
        self.sweep_p()
        self.t += 1
