"""The engine's input, as one in-memory object: the contents of a reference ``simulation_dir``.

Field names are the reference's ``.dat`` file stems
(/root/reference/fullwave/solver/input_file_writer.py:563-627, :753-821).  ``from_dat_dir`` /
``to_dat_dir`` speak the reference's file protocol byte for byte; ``from_fullwave_objects`` builds the
same fields straight from the reference's Python objects (what ``InputFileWriter`` would have written)
so that the in-process path can skip the disk round trip.
"""

from __future__ import annotations

from dataclasses import dataclass, field
from pathlib import Path

import numpy as np

from . import stencil

MAP_NAMES = ("rho", "K", "beta", "kappax", "kappau",
             "apmlx1", "bpmlx1", "apmlx2", "bpmlx2",
             "apmlu1", "bpmlu1", "apmlu2", "bpmlu2")

# fullwave python name -> .dat stem, isotropic variant (input_file_writer.py:581-591)
_RELAX_RENAME = {
    "kappa_x": "kappax", "kappa_u": "kappau",
    "a_pml_x1": "apmlx1", "b_pml_x1": "bpmlx1", "a_pml_x2": "apmlx2", "b_pml_x2": "bpmlx2",
    "a_pml_u1": "apmlu1", "b_pml_u1": "bpmlu1", "a_pml_u2": "apmlu2", "b_pml_u2": "bpmlu2",
}


_CAST_PARALLEL_MIN = 1 << 23      # elements; below this a plain numpy cast is as fast


def cast_c(a, dtype) -> np.ndarray:
    """`np.ascontiguousarray(a, dtype=dtype)` with the element conversion spread over the host cores for large arrays.
    The transducer examples hand over a float64 `icmat` of 1.6 GB (convex_transducer: 62 100 sources x 3244 steps) and
    the host-maps path thirteen float64 maps of the extended grid; one numpy thread converts ~1.5 GB/s, which was 1.07 s
    of a 1.5 s `Solver.run`.  numpy releases the GIL inside `copyto`, so row blocks convert concurrently.  Same values:
    every element goes through the same IEEE conversion."""
    a = np.asarray(a)
    dtype = np.dtype(dtype)
    if a.dtype == dtype and a.flags["C_CONTIGUOUS"]:
        return a
    if a.size < _CAST_PARALLEL_MIN or a.ndim == 0 or a.shape[0] < 2:
        return np.ascontiguousarray(a, dtype=dtype)
    import os
    from concurrent.futures import ThreadPoolExecutor
    out = np.empty(a.shape, dtype)
    n = a.shape[0]
    workers = max(1, min(16, os.cpu_count() or 1, n))
    edges = np.linspace(0, n, 4 * workers + 1).astype(np.int64)

    def part(k):
        lo, hi = int(edges[k]), int(edges[k + 1])
        if hi > lo:
            np.copyto(out[lo:hi], a[lo:hi], casting="unsafe")
    with ThreadPoolExecutor(max_workers=workers) as ex:
        list(ex.map(part, range(len(edges) - 1)))
    return out


@dataclass
class Problem:
    ndim: int
    nX: int
    nY: int
    nZ: int
    nT: int
    nTic: int
    modT: int
    ndmap: int
    dX: float
    dT: float
    rho: np.ndarray
    K: np.ndarray
    beta: np.ndarray
    kappax: np.ndarray
    kappau: np.ndarray
    apmlx1: np.ndarray
    bpmlx1: np.ndarray
    apmlx2: np.ndarray
    bpmlx2: np.ndarray
    apmlu1: np.ndarray
    bpmlu1: np.ndarray
    apmlu2: np.ndarray
    bpmlu2: np.ndarray
    dmap: np.ndarray            # float32 [9, 2, ndmap]
    dcmap: np.ndarray           # int32 grid
    icc: np.ndarray             # int32 [ncoords, ndim]
    icmat: np.ndarray           # float32 [ncoords, nTic]
    outc: np.ndarray            # int32 [ncoordsout, ndim]
    icczero: np.ndarray         # int32 [ncoordszero, ndim]
    extra: dict = field(default_factory=dict)  # dY, dZ, c0, c, d: written by the reference, unused by the engine
    # False = the reference 3D binary's behaviour: only the first nX*nY entries of dcmap are honoured, the
    # rest read 0 (include/fw25.h, fw25_problem.dcmap_full3d).  No effect in 2D.
    dcmap_full3d: bool = False
    # Anisotropic-relaxation protocol (use_isotropic_relaxation=False upstream; input_file_writer.py:592-620):
    # {file stem: map} for kappa{x,y[,z]}, kappa{u,w[,v]}, apml/bpml{x,y[,z],u,w[,v]}{1,2}.  None = isotropic.
    # When set, the isotropic fields kappax .. bpmlu2 above hold the x-axis members (kappax, kappau, apmlx*, apmlu*).
    aniso: dict | None = None
    # Box sensors (include/fw25.h, out_box): (lo..., hi...) of a box whose every point is a sensor, row-major -- a
    # rectangular `Sensor(mask)` or `record_whole_domain`.  outc is then unused (may be an empty list).
    out_box: tuple | None = None

    # ------------------------------------------------------------------ basics
    @property
    def shape(self) -> tuple[int, ...]:
        return (self.nX, self.nY, self.nZ) if self.ndim == 3 else (self.nX, self.nY)

    @property
    def n_points(self) -> int:
        return int(np.prod(self.shape))

    @property
    def ncoords(self) -> int:
        return int(self.icc.shape[0])

    @property
    def ncoordsout(self) -> int:
        if self.out_box is not None:
            b = np.asarray(self.out_box, np.int64).reshape(2, self.ndim)
            return int(np.prod(b[1] - b[0]))
        return int(self.outc.shape[0])

    def sensor_coords(self) -> np.ndarray:
        """int32 [ncoordsout, ndim]: outc, or the box's points in row-major order."""
        if self.out_box is None:
            return self.outc
        b = np.asarray(self.out_box, np.int64).reshape(2, self.ndim)
        grids = np.meshgrid(*[np.arange(lo, hi) for lo, hi in zip(b[0], b[1])], indexing="ij")
        return np.stack([g.reshape(-1) for g in grids], axis=1).astype(np.int32)

    @property
    def ncoordszero(self) -> int:
        return int(self.icczero.shape[0])

    @property
    def n_frames(self) -> int:
        return -(-self.nT // self.modT) if self.nT > 0 else 0

    @staticmethod
    def aniso_stems(ndim: int) -> tuple[str, ...]:
        """File stems of the anisotropic protocol, per-axis letters: velocity sweep x, y[, z]; pressure sweep
        u, w (2D) / u, v, w (3D) for axes x, y[, z]."""
        letters = ("x", "y", "u", "w") if ndim == 2 else ("x", "y", "z", "u", "v", "w")
        return tuple(f"kappa{l}" for l in letters) + tuple(f"{ab}pml{l}{nu}" for l in letters for nu in (1, 2)
                                                           for ab in ("a", "b"))

    def normalise(self) -> "Problem":
        """Cast every array to the protocol dtype / shape (C-contiguous) and validate sizes."""
        if self.ndim not in (2, 3):
            raise ValueError("ndim must be 2 or 3")
        if self.ndim == 2:
            self.nZ = 1
        if self.aniso is not None:
            for stem in self.aniso_stems(self.ndim):
                if stem not in self.aniso:
                    raise ValueError(f"anisotropic problem lacks {stem}")
                a = cast_c(self.aniso[stem], np.float32)
                if a.size != self.n_points:
                    raise ValueError(f"{stem}: {a.size} values, grid has {self.n_points}")
                self.aniso[stem] = a.reshape(self.shape)
        for name in MAP_NAMES:
            if getattr(self, name) is None:      # maps resident on the device (mapgen.MapSet): nothing to cast
                continue
            a = cast_c(getattr(self, name), np.float32)
            if a.size != self.n_points:
                raise ValueError(f"{name}: {a.size} values, grid has {self.n_points}")
            setattr(self, name, a.reshape(self.shape))
        if self.dcmap is not None:
            self.dcmap = cast_c(self.dcmap, np.int32).reshape(self.shape)
        self.dmap = np.ascontiguousarray(self.dmap, dtype=np.float32).reshape(9, 2, -1)
        if self.dmap.shape[2] < self.ndmap:
            raise ValueError("dmap has fewer columns than ndmap")
        if self.dmap.shape[2] != self.ndmap:
            # the reference writes ndmap = 1 for a homogeneous medium; keep the columns the engine reads
            self.dmap = np.ascontiguousarray(self.dmap[:, :, : self.ndmap])
        if self.dcmap is not None and self.dcmap.size and (self.dcmap.min() < 0 or self.dcmap.max() >= self.ndmap):
            raise ValueError("dcmap entries must lie in [0, ndmap)")
        self.icc = np.ascontiguousarray(self.icc, dtype=np.int32).reshape(-1, self.ndim)
        if self.out_box is not None:
            self.out_box = tuple(int(v) for v in np.asarray(self.out_box).reshape(-1))
            if len(self.out_box) != 2 * self.ndim:
                raise ValueError("out_box must be (lo..., hi...) with 2 * ndim entries")
            self.outc = np.zeros((0, self.ndim), np.int32)
        self.outc = np.ascontiguousarray(self.outc, dtype=np.int32).reshape(-1, self.ndim)
        self.icczero = np.ascontiguousarray(self.icczero, dtype=np.int32).reshape(-1, self.ndim)
        self.icmat = cast_c(self.icmat, np.float32).reshape(
            self.ncoords, self.nTic if self.ncoords == 0 else -1)
        if self.ncoords and self.icmat.shape[1] != self.nTic:
            raise ValueError("icmat must be [ncoords, nTic]")
        if self.modT < 1:
            raise ValueError("modT must be >= 1")
        return self

    # ------------------------------------------------------------------ reference file protocol
    @classmethod
    def from_dat_dir(cls, sim_dir: str | Path) -> "Problem":
        """Read a reference simulation directory (the files the shipped binary opens)."""
        d = Path(sim_dir)

        def i32(name, default=None):
            f = d / f"{name}.dat"
            if not f.exists():
                if default is None:
                    raise FileNotFoundError(f)
                return default
            return int(np.fromfile(f, dtype=np.int32)[0])

        def f32(name):
            return float(np.fromfile(d / f"{name}.dat", dtype=np.float32)[0])

        ndim = 3 if (d / "nZ.dat").exists() else 2
        nX, nY = i32("nX"), i32("nY")
        nZ = i32("nZ") if ndim == 3 else 1
        shape = (nX, nY, nZ) if ndim == 3 else (nX, nY)
        n = int(np.prod(shape))

        def fmap(name):
            a = np.fromfile(d / f"{name}.dat", dtype=np.float32)
            if a.size != n:
                raise ValueError(f"{name}.dat holds {a.size} floats, expected {n}")
            return a.reshape(shape)

        ncoords, ncoordsout = i32("ncoords"), i32("ncoordsout")
        ncoordszero = i32("ncoordszero", 0)
        nTic = i32("nTic")
        ndmap = i32("ndmap")

        def coords(name, cnt):
            f = d / f"{name}.dat"
            if cnt == 0 or not f.exists():
                return np.zeros((0, ndim), np.int32)
            return np.fromfile(f, dtype=np.int32)[: cnt * ndim].reshape(cnt, ndim)

        maps = {name: fmap(name) for name in MAP_NAMES}
        aniso = None
        if (d / "kappay.dat").exists():      # the anisotropic file set (dangling static-map links do not count)
            aniso = {stem: fmap(stem) for stem in cls.aniso_stems(ndim)}
        pb = cls(
            ndim=ndim, nX=nX, nY=nY, nZ=nZ, nT=i32("nT"), nTic=nTic, modT=i32("modT"), ndmap=ndmap,
            dX=f32("dX"), dT=f32("dT"), **maps,
            dmap=np.fromfile(d / "dmap.dat", dtype=np.float32).reshape(9, 2, -1),
            dcmap=np.fromfile(d / "dcmap.dat", dtype=np.int32).reshape(shape),
            icc=coords("icc", ncoords),
            icmat=np.fromfile(d / "icmat.dat", dtype=np.float32)[: ncoords * nTic].reshape(ncoords, nTic),
            outc=coords("outc", ncoordsout),
            icczero=coords("icczero", ncoordszero),
            aniso=aniso,
        )
        return pb.normalise()

    def to_dat_dir(self, sim_dir: str | Path) -> Path:
        """Write the directory the reference's InputFileWriter would have produced
        (input_file_writer.py:716-821); used to feed the reference binary identical bytes."""
        d = Path(sim_dir)
        d.mkdir(parents=True, exist_ok=True)
        for name in MAP_NAMES:
            getattr(self, name).astype(np.float32).tofile(d / f"{name}.dat")
        for stem, a in (self.aniso or {}).items():
            a.astype(np.float32).tofile(d / f"{stem}.dat")
        ex = self.extra
        np.asarray(ex.get("c", np.zeros(self.shape, np.float32)), dtype=np.float32).tofile(d / "c.dat")
        np.asarray(ex.get("d", np.zeros((9, 2))), dtype=np.float32).tofile(d / "d.dat")
        self.dmap.astype(np.float32).tofile(d / "dmap.dat")
        self.dcmap.astype(np.int32).tofile(d / "dcmap.dat")
        self.icc.astype(np.int32).tofile(d / "icc.dat")
        self.sensor_coords().astype(np.int32).tofile(d / "outc.dat")
        self.icczero.astype(np.int32).tofile(d / "icczero.dat")
        self.icmat.astype(np.float32).tofile(d / "icmat.dat")
        ints = {"nX": self.nX, "nY": self.nY, "nT": self.nT, "ncoords": self.ncoords,
                "ncoordsout": self.ncoordsout, "ncoordszero": self.ncoordszero, "nTic": self.nTic,
                "modT": self.modT, "ndmap": self.ndmap}
        floats = {"dX": self.dX, "dY": ex.get("dY", self.dX), "dT": self.dT, "c0": ex.get("c0", 1540.0)}
        if self.ndim == 3:
            ints["nZ"] = self.nZ
            floats["dZ"] = ex.get("dZ", self.dX)
        for k, v in ints.items():
            np.array(v).astype(np.int32).tofile(d / f"{k}.dat")
        for k, v in floats.items():
            np.array(v).astype(np.float32).tofile(d / f"{k}.dat")
        return d

    # ------------------------------------------------------------------ reference Python objects
    @classmethod
    def from_fullwave_objects(cls, grid, medium, source, sensor, out_box=None) -> "Problem":
        """Build the engine input from the reference's (already PML-extended) objects, i.e. what
        ``InputFileWriter(...).run`` would write (input_file_writer.py:95-103, :150-153, :563-627,
        :753-821), without touching the disk.  Duck-typed: needs grid.{nx,ny,nz,nt,dx,dy,dz,dt,c0,cfl,is_3d},
        medium.{sound_speed,density,beta,bulk_modulus,air_map,relaxation_param_dict_for_fw2},
        source.{incoords,icmat}, sensor.{outcoords,sampling_modulus_time}."""
        is_3d = bool(grid.is_3d)
        c = np.asarray(medium.sound_speed)
        d_tab, dmap, dcmap, ndmap = stencil.tables(c, dt=grid.dt, dx=grid.dx, cfl=grid.cfl, is_3d=is_3d)
        relax = medium.relaxation_param_dict_for_fw2
        maps = {dat: relax[py] for py, dat in _RELAX_RENAME.items()}
        aniso = None
        if "kappa_y" in relax:       # use_isotropic_relaxation=False upstream: per-axis maps (a_pml_y1 -> apmly1.dat)
            aniso = {key.replace("_", ""): relax[key] for key in relax
                     if key.replace("_", "") in cls.aniso_stems(3 if is_3d else 2)}
        icmat = np.asarray(source.icmat)
        air = np.asarray(medium.air_map)
        pb = cls(
            ndim=3 if is_3d else 2, nX=int(grid.nx), nY=int(grid.ny), nZ=int(grid.nz) if is_3d else 1,
            nT=int(grid.nt), nTic=int(icmat.shape[1]), modT=int(sensor.sampling_modulus_time), ndmap=ndmap,
            dX=float(np.float32(grid.dx)), dT=float(np.float32(grid.dt)),
            rho=medium.density, K=medium.bulk_modulus, beta=medium.beta, **maps,
            dmap=dmap, dcmap=dcmap,
            icc=np.asarray(source.incoords), icmat=icmat,
            outc=np.zeros((0, c.ndim), np.int32) if out_box is not None else np.asarray(sensor.outcoords),
            icczero=np.stack(np.nonzero(air != 0), axis=1) if air.any() else np.zeros((0, c.ndim), np.int32),
            extra={"c": c, "d": d_tab, "dY": grid.dy, "dZ": getattr(grid, "dz", grid.dx), "c0": grid.c0},
            aniso=aniso, out_box=out_box,
        )
        return pb.normalise()

    @classmethod
    def for_device_maps(cls, mapset, grid, source, sensor, air_map=None, out_box=None, icczero=None) -> "Problem":
        """Engine input whose 13 maps + dcmap already sit in HBM (`mapgen.MapSet`): only the step counts, the
        stencil table and the source / sensor / air-voxel lists come from the host.  grid, source, sensor: the
        reference's PML-extended objects (as in `from_fullwave_objects`) or anything with the same attributes;
        air voxels: the extended air map, or their coordinates (`icczero`) directly."""
        is_3d = len(mapset.shape) == 3
        icmat = np.asarray(source.icmat)
        nd = 3 if is_3d else 2
        air = None if air_map is None else np.asarray(air_map)
        if icczero is not None:
            air = None
        none_maps = {name: None for name in MAP_NAMES}
        pb = cls(
            ndim=nd, nX=int(mapset.shape[0]), nY=int(mapset.shape[1]), nZ=int(mapset.shape[2]) if is_3d else 1,
            nT=int(grid.nt), nTic=int(icmat.shape[1]), modT=int(sensor.sampling_modulus_time), ndmap=mapset.ndmap,
            dX=float(np.float32(grid.dx)), dT=float(np.float32(grid.dt)), **none_maps,
            dmap=mapset.dmap, dcmap=None,
            icc=np.asarray(source.incoords), icmat=icmat,
            outc=np.zeros((0, nd), np.int32) if out_box is not None else np.asarray(sensor.outcoords),
            icczero=(np.asarray(icczero) if icczero is not None else
                     np.stack(np.nonzero(air != 0), axis=1) if air is not None and air.any()
                     else np.zeros((0, nd), np.int32)),
            extra={"d": mapset.d_table, "c0": getattr(grid, "c0", 1540.0)},
            dcmap_full3d=True,       # a reference-style truncation is already inside the generated dcmap
            out_box=out_box,
        )
        return pb.normalise()

    # ------------------------------------------------------------------ slabs
    def slab(self, gx0: int, gx1: int) -> "Problem":
        """Planes [gx0, gx1) of every grid array (views); coordinates stay global."""
        kw = {name: getattr(self, name)[gx0:gx1] for name in MAP_NAMES}
        if self.aniso is not None:
            kw["aniso"] = {k: v[gx0:gx1] for k, v in self.aniso.items()}
        return Problem(ndim=self.ndim, nX=gx1 - gx0, nY=self.nY, nZ=self.nZ, nT=self.nT, nTic=self.nTic,
                       modT=self.modT, ndmap=self.ndmap, dX=self.dX, dT=self.dT, **kw, dmap=self.dmap,
                       dcmap=self.dcmap[gx0:gx1], icc=self.icc, icmat=self.icmat, outc=self.outc,
                       icczero=self.icczero, extra={}, dcmap_full3d=self.dcmap_full3d, out_box=self.out_box)
