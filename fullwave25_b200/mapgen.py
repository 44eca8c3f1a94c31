"""Medium maps built on the GPU: the host side of `fw25_mapgen` (include/fw25.h).

Upstream, `Solver.run` first calls `PMLBuilder.run` (/root/reference/fullwave/solver/pml_builder.py:812-840), which
ramps the ten relaxation-parameter maps towards their PML targets and turns them into a / b maps over the EXTENDED
grid in float64 numpy, then `InputFileWriter` casts them to float32 and derives K and dcmap
(input_file_writer.py:95-103, :563-627, :716-745).  Here the host only gathers what that code reads -- the USER-grid
maps of the `Medium` / `MediumRelaxationMaps` the solver was given, the layer counts, dt, and three 1-D transition
functions sampled with the reference's own numpy expressions -- and one CUDA kernel writes the thirteen float32 maps
and dcmap straight into the engine's HBM layout.  There is no CPU path: without libfw25.so this raises.
"""

from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import engine, stencil
from .problem import MAP_NAMES

_D = C.POINTER(C.c_double)

# the reference's dictionary order (fullwave/solver/utils.py:68-86) == last axis of the look-up database
RELAX_KEYS = ("kappa_x1", "kappa_x2", "d_x1_nu1", "alpha_x1_nu1", "d_x2_nu1", "alpha_x2_nu1",
              "d_x1_nu2", "alpha_x1_nu2", "d_x2_nu2", "alpha_x2_nu2")


class CMedium(C.Structure):
    """fw25_medium."""
    _fields_ = [
        ("ndim", C.c_int32), ("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32),
        ("m_spatial_order", C.c_int32), ("n_pml_layer", C.c_int32), ("n_transition_layer", C.c_int32),
        ("use_pml", C.c_int32),
        ("dt", C.c_double), ("d_target_pml", C.c_double),
        ("tf_polynomial", C.c_void_p), ("tf_linear", C.c_void_p), ("tf_cosine", C.c_void_p),
        ("sound_speed", C.c_void_p), ("density", C.c_void_p), ("beta", C.c_void_p),
        ("relax", C.c_void_p * 10),
        ("alpha_coeff", C.c_void_p), ("alpha_power", C.c_void_p),
        ("lut", C.c_void_p), ("lut_alpha", C.c_void_p), ("lut_power", C.c_void_p),
        ("lut_na", C.c_int32), ("lut_np", C.c_int32),
        ("alpha_min", C.c_double), ("alpha_max", C.c_double), ("power_min", C.c_double), ("power_max", C.c_double),
        ("lut_invalid", C.c_void_p),
        ("c_round_min", C.c_int32), ("dcmap_full3d", C.c_int32),
        ("input_f32", C.c_int32), ("reserved_", C.c_int32),
    ]


SIGS = {
    "fw25_mapgen": (C.c_int, [C.POINTER(CMedium), C.c_int32, C.POINTER(C.c_void_p), _D]),
    "fw25_mapgen_slab": (C.c_int, [C.POINTER(CMedium), C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                   C.POINTER(C.c_void_p), _D]),
    "fw25_mapgen_slab_begin": (C.c_int, [C.POINTER(CMedium), C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                         C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "fw25_mapgen_finish": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), _D]),
    "fw25_mapset_problem": (C.c_int, [C.c_void_p, C.POINTER(engine.CProblem)]),
    "fw25_mapset_read": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p]),
    "fw25_mapset_invalid_count": (C.c_int64, [C.c_void_p]),
    "fw25_mapset_destroy": (None, [C.c_void_p]),
    "fw25_run_medium": (C.c_int, [C.POINTER(CMedium), C.POINTER(engine.CProblem), C.c_int32, engine._F, C.c_size_t,
                                  C.POINTER(engine.CStats)]),
    "fw25_run_medium_multi": (C.c_int, [C.POINTER(CMedium), C.POINTER(engine.CProblem), engine._I, C.c_int32, engine._F,
                                        C.c_size_t, C.POINTER(engine.CStats)]),
}


@dataclass
class LookupTable:
    """The relaxation-parameter database `Medium.build()` indexes (utils/relaxation_parameters.py:115-158)."""
    database: np.ndarray          # [nA, nP, 10]
    alpha_list: np.ndarray        # [nA] ascending
    power_list: np.ndarray        # [nP] ascending
    invalid_matrix: np.ndarray | None = None

    @classmethod
    def from_mat(cls, path) -> "LookupTable":
        from scipy.io import loadmat
        db = loadmat(path)
        return cls(db["database"], db["alpha_0_list"][0], db["power_list"][0], db.get("invalid_matrix"))


@dataclass
class MediumSpec:
    """Everything `PMLBuilder.run` + `InputFileWriter` read to build the engine's maps."""
    user_shape: tuple
    dt: float                      # extended_grid.dt
    dx: float                      # extended_grid.dx
    c0: float
    cfl: float
    sound_speed: np.ndarray
    density: np.ndarray
    beta: np.ndarray
    relax: dict | None = None      # RELAX_KEYS -> user-grid map (MediumRelaxationMaps.relaxation_param_dict)
    alpha_coeff: np.ndarray | None = None
    alpha_power: np.ndarray | None = None
    lut: LookupTable | None = None
    m_spatial_order: int = 8
    n_pml_layer: int = 40
    n_transition_layer: int = 40
    use_pml: bool = True
    n_polynomial: int = 2
    theoretical_reflection_coefficient: float = 1e-30
    dcmap_full3d: bool = False
    extra: dict = field(default_factory=dict)
    # x-sharded runs: the maps hold only the user-grid x planes [u0, u0 + n) of a grid whose full extent is user_shape
    # (each rank keeps the planes its slab reads); extra must then carry c_min / c_max of the WHOLE medium
    user_planes: tuple | None = None

    @property
    def ndim(self) -> int:
        return len(self.user_shape)

    @property
    def num_boundary_points(self) -> int:
        return self.n_transition_layer + self.n_pml_layer + self.m_spatial_order

    @property
    def extended_shape(self) -> tuple:
        return tuple(int(n) + 2 * self.num_boundary_points for n in self.user_shape)

    # -- host scalars and 1-D tables, the reference's own numpy expressions --------------------------------
    def d_target_pml(self) -> float:
        """pml_builder.py:1063-1070 (identical in `_apply_pml`, :881-888)."""
        pml_layer_m = self.dx * self.n_pml_layer
        transition_layer_m = self.dx * self.n_transition_layer
        return float(-(self.n_polynomial + 1) * self.c0 * np.log(self.theoretical_reflection_coefficient)
                     / (2 * (pml_layer_m + transition_layer_m)))

    def transition_tables(self):
        """(polynomial, linear, cosine) samples (pml_builder.py:30-39, :1296-1338)."""
        x_full = np.linspace(0, 1, self.n_pml_layer + self.n_transition_layer + 1)
        x_tr = np.linspace(0, 1, self.n_transition_layer + 1)
        return (np.ascontiguousarray(x_full ** self.n_polynomial), np.ascontiguousarray(x_full),
                np.ascontiguousarray(0.5 * (1 - np.cos(np.pi * x_tr))))

    def stencil_tables(self):
        """d, dmap, ndmap and round(min c) from the USER grid: padding by edge replication leaves min / max of
        the sound speed unchanged (input_file_writer.py:95-103, :183-559)."""
        c = self.sound_speed
        if "c_min" in self.extra and "c_max" in self.extra:      # known to the caller (a pass over a multi-GB map otherwise)
            c_min, c_max = np.float64(self.extra["c_min"]), np.float64(self.extra["c_max"])
        else:
            c_min, c_max = np.float64(c.min()), np.float64(c.max())
        dim = int(stencil.matlab_round(c_max) - stencil.matlab_round(c_min))
        dm = stencil.d_map(c_min, dim, self.dt, self.dx, is_3d=self.ndim == 3)
        ndmap = 1 if dim == 0 else dm.shape[2]
        return (stencil.d_table(self.cfl, is_3d=self.ndim == 3), np.ascontiguousarray(dm.astype(np.float32)), ndmap,
                int(stencil.matlab_round(c_min)))

    @classmethod
    def from_pml_builder(cls, pml_builder, *, use_pml: bool = True, dcmap_full3d: bool = False,
                         lut: LookupTable | None = None) -> "MediumSpec":
        """From the reference's `PMLBuilder` (as `Solver.__init__` builds it, solver.py:527-536): its ORIGINAL
        medium, layer counts and extended grid.  `Medium` objects need the look-up database (`lut`, default: the
        file the medium names)."""
        med, eg = pml_builder.medium_org, pml_builder.extended_grid
        kw = dict(user_shape=tuple(np.asarray(med.sound_speed).shape), dt=float(eg.dt), dx=float(eg.dx),
                  c0=float(eg.c0), cfl=float(eg.cfl), sound_speed=med.sound_speed, density=med.density,
                  beta=med.beta, m_spatial_order=int(pml_builder.m_spatial_order),
                  n_pml_layer=int(pml_builder.n_pml_layer), n_transition_layer=int(pml_builder.n_transition_layer),
                  use_pml=bool(use_pml), n_polynomial=getattr(pml_builder, "n_polynomial", 2),
                  theoretical_reflection_coefficient=getattr(pml_builder, "theoritical_reflection_coefficient", 1e-30),
                  dcmap_full3d=dcmap_full3d)
        if hasattr(med, "relaxation_param_dict") and not hasattr(med, "alpha_coeff"):
            return cls(relax={k: med.relaxation_param_dict[k] for k in RELAX_KEYS}, **kw)
        if lut is None:
            lut = LookupTable.from_mat(med.path_relaxation_parameters_database)
        return cls(alpha_coeff=med.alpha_coeff, alpha_power=med.alpha_power, lut=lut, **kw)


def _lib():
    lib = engine.lib()
    if not getattr(lib, "_fw25_mapgen_bound", False):
        for name, (res, args) in SIGS.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        lib._fw25_mapgen_bound = True
    return lib


def _f64(a, shape) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float64)
    if a.shape != tuple(shape):
        raise ValueError(f"map shape error: {a.shape} != {tuple(shape)}")
    return a


def _user_map(a, shape) -> np.ndarray:
    """A user-grid map as the C-ABI takes it: C-contiguous float64 (the reference's Medium) or float32 as is."""
    a = np.asarray(a)
    a = np.ascontiguousarray(a, dtype=np.float32 if a.dtype == np.float32 else np.float64)
    if a.shape != tuple(shape):
        raise ValueError(f"map shape error: {a.shape} != {tuple(shape)}")
    return a


def marshal_medium(spec: MediumSpec):
    """MediumSpec -> (CMedium, keepalive list, (d_table, dmap, ndmap)).  float32 user maps go through unconverted
    (fw25_medium.input_f32) when EVERY user-grid map is float32; otherwise everything is float64."""
    if spec.ndim not in (2, 3):
        raise ValueError("the medium must be 2D or 3D")
    full_shape = tuple(int(n) for n in spec.user_shape)
    shape = full_shape if spec.user_planes is None else (int(spec.user_planes[1]),) + full_shape[1:]
    if spec.user_planes is not None and not ("c_min" in spec.extra and "c_max" in spec.extra):
        raise ValueError("a MediumSpec holding a plane range needs extra['c_min'] / extra['c_max'] of the whole medium")
    md = CMedium()
    keep = []
    user = [spec.sound_speed, spec.density, spec.beta]
    if spec.relax is not None:
        user += [spec.relax[k] for k in RELAX_KEYS]
    else:
        if spec.lut is None or spec.alpha_coeff is None or spec.alpha_power is None:
            raise ValueError("MediumSpec needs either the relaxation maps or alpha_coeff / alpha_power + a table")
        user += [spec.alpha_coeff, spec.alpha_power]
    f32 = all(np.asarray(a).dtype == np.float32 for a in user)

    def put(a, shp=shape, user_map=False):
        a = _user_map(a, shp) if (user_map and f32) else _f64(a, shp)
        keep.append(a)
        return a.ctypes.data

    md.ndim = spec.ndim
    md.nx, md.ny, md.nz = full_shape[0], full_shape[1], full_shape[2] if spec.ndim == 3 else 1
    md.m_spatial_order, md.n_pml_layer = spec.m_spatial_order, spec.n_pml_layer
    md.n_transition_layer, md.use_pml = spec.n_transition_layer, int(spec.use_pml)
    md.dt = spec.dt
    md.input_f32 = int(f32)
    if spec.use_pml:
        md.d_target_pml = spec.d_target_pml()
        tp, tl, tc = spec.transition_tables()
        md.tf_polynomial, md.tf_linear, md.tf_cosine = put(tp, tp.shape), put(tl, tl.shape), put(tc, tc.shape)
    md.sound_speed, md.density, md.beta = (put(a, user_map=True) for a in user[:3])
    if spec.relax is not None:
        for i, k in enumerate(RELAX_KEYS):
            md.relax[i] = put(spec.relax[k], user_map=True)
    else:
        t = spec.lut
        db = np.ascontiguousarray(t.database, dtype=np.float64)
        if db.ndim != 3:
            raise ValueError("look_up_table must have 3 dimensions.")
        if db.shape[2] != 10:
            raise ValueError("look_up_table must have 4 * n_relaxation_mechanisms + 2 columns.")
        if np.isnan(db).any():
            raise ValueError("look_up_table must not contain NaN values.")
        al, pl = np.asarray(t.alpha_list, np.float64).reshape(-1), np.asarray(t.power_list, np.float64).reshape(-1)
        md.alpha_coeff, md.alpha_power = put(spec.alpha_coeff, user_map=True), put(spec.alpha_power, user_map=True)
        md.lut = put(db, db.shape)
        md.lut_alpha, md.lut_power = put(al.round(10), al.shape), put(pl.round(10), pl.shape)
        md.lut_na, md.lut_np = db.shape[0], db.shape[1]
        md.alpha_min, md.alpha_max = float(al.min()), float(al.max())
        md.power_min, md.power_max = float(pl.min()), float(pl.max().round(4))
        if t.invalid_matrix is not None:
            inv = np.ascontiguousarray(np.asarray(t.invalid_matrix) != 0, dtype=np.uint8)
            keep.append(inv)
            md.lut_invalid = inv.ctypes.data
    tables = spec.stencil_tables()
    md.c_round_min = tables[3]
    md.dcmap_full3d = int(bool(spec.dcmap_full3d))
    return md, keep, tables[:3]


def run_medium(spec: MediumSpec, pb, device: int = 0, device_ids=None):
    """Whole job from the USER-grid medium through fw25_run_medium (C-ABI): the medium is uploaded block by block, the
    maps are generated as the blocks land and the first time steps already run underneath.  pb: a Problem holding the
    step counts and the coordinate lists on the EXTENDED grid (`Problem.for_device_maps`-style; its maps are unused).
    device_ids with several entries: x-slabs over those GPUs, each building its own slab of the maps
    (fw25_run_medium_multi).  Returns (genout [n_frames, ncoordsout], stats)."""
    import time
    t0 = time.perf_counter()
    md, keep, (d_table, dmap, ndmap) = marshal_medium(spec)
    pb.normalise()
    s, keep2 = engine.marshal(pb, device_maps={name: 0 for name in MAP_NAMES + ("dcmap",)})
    s.maps_on_device = 0
    s.dmap, s.ndmap = dmap.ctypes.data, ndmap
    genout = np.zeros((pb.n_frames, pb.ncoordsout), np.float32)
    st = engine.CStats()
    t1 = time.perf_counter()
    if device_ids is not None and len(device_ids) > 1:
        ids = np.asarray(list(device_ids), np.int32)
        engine._check(_lib().fw25_run_medium_multi(C.byref(md), C.byref(s), ids.ctypes.data_as(engine._I), len(ids),
                                                   genout.ctypes.data_as(engine._F), genout.size, C.byref(st)))
    else:
        dev0 = int(device_ids[0]) if device_ids else device
        engine._check(_lib().fw25_run_medium(C.byref(md), C.byref(s), dev0, genout.ctypes.data_as(engine._F), genout.size,
                                             C.byref(st)))
    t2 = time.perf_counter()
    del keep, keep2
    stats = {f: getattr(st, f) for f, _ in engine.CStats._fields_}
    stats.update(host_marshal_ms=(t1 - t0) * 1e3, native_call_ms=(t2 - t1) * 1e3)
    return genout, stats


class MapSet:
    """Device-resident maps of one medium (fw25_mapset).  Keep it alive as long as an engine uses it."""

    def __init__(self, spec: MediumSpec, device: int = 0, planes: tuple | None = None, background: bool = False):
        """planes = (gx0, gx1): build the extended x planes [gx0, gx1) only (one x-slab incl. its ghost planes,
        fw25_mapgen_slab); spec.user_planes says which user-grid planes the host arrays hold.  background=True
        (fw25_mapgen_slab_begin): returns at once with the final device pointers -- an engine can be created on them --
        while the maps are still being uploaded and generated; call wait() before anything steps on them."""
        self.spec = spec
        self.device = device
        md, keep, (self.d_table, self.dmap, self.ndmap) = marshal_medium(spec)
        h = C.c_void_p()
        ms = (C.c_double * 2)()
        self._job = None
        if planes is None and spec.user_planes is None and not background:
            engine._check(_lib().fw25_mapgen(C.byref(md), device, C.byref(h), ms))
            self.shape = spec.extended_shape
        else:
            gx0, gx1 = planes if planes is not None else (0, spec.extended_shape[0])
            u0, un = spec.user_planes if spec.user_planes is not None else (0, spec.user_shape[0])
            if background:
                job = C.c_void_p()
                engine._check(_lib().fw25_mapgen_slab_begin(C.byref(md), device, int(gx0), int(gx1), int(u0), int(un),
                                                            C.byref(h), C.byref(job)))
                self._job, self._keep = job, keep          # the host arrays stay alive until wait()
            else:
                engine._check(_lib().fw25_mapgen_slab(C.byref(md), device, int(gx0), int(gx1), int(u0), int(un),
                                                      C.byref(h), ms))
            self.shape = (int(gx1) - int(gx0),) + tuple(spec.extended_shape[1:])
        del keep
        self._h = h
        self.upload_ms, self.kernel_ms = ms[0], ms[1]
        self.invalid_count = int(_lib().fw25_mapset_invalid_count(h)) if self._job is None else 0

    def wait(self) -> None:
        """background=True: block until every plane has been generated (fw25_mapgen_finish)."""
        if self._job is None:
            return
        job, self._job = self._job, None
        out = C.c_void_p()
        ms = (C.c_double * 2)()
        rc = _lib().fw25_mapgen_finish(job, C.byref(out), ms)
        self._keep = None
        if rc:
            self._h = None                                   # a failed job has freed the set
            engine._check(rc)
        self.upload_ms, self.kernel_ms = ms[0], ms[1]
        self.invalid_count = int(_lib().fw25_mapset_invalid_count(self._h))

    def fill(self, cpb: "engine.CProblem") -> None:
        """Point a fw25_problem at these maps (fw25_mapset_problem)."""
        engine._check(_lib().fw25_mapset_problem(self._h, C.byref(cpb)))

    def device_maps(self) -> dict:
        """{name: device address} + pitch, the form `engine.Engine(device_maps=...)` takes."""
        s = engine.CProblem()
        self.fill(s)
        out = {name: getattr(s, name) for name in MAP_NAMES + ("dcmap",)}
        out["pitch"] = s.map_pitch
        out["owner"] = self
        return out

    def read(self, name: str) -> np.ndarray:
        """One map back on the host, dense: the bytes the reference would have written to <name>.dat."""
        out = np.zeros(self.shape, np.int32 if name == "dcmap" else np.float32)
        engine._check(_lib().fw25_mapset_read(self._h, name.encode(), out.ctypes.data))
        return out

    def close(self) -> None:
        if getattr(self, "_job", None) is not None and engine._lib is not None:
            try:
                self.wait()
            except engine.EngineError:
                pass
        if getattr(self, "_h", None) and engine._lib is not None:
            engine._lib.fw25_mapset_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
