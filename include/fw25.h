/*
 * fw25.h -- C-ABI of the B200-native Fullwave 2.5 time-stepping engine (libfw25.so).
 *
 * This is the drop-in boundary.  Upstream, `fullwave.Solver.run` reaches the engine by writing a
 * directory of raw .dat files and exec'ing a pre-compiled CUDA binary:
 *     /root/reference/fullwave/solver/input_file_writer.py:105-179, :563-881   (the .dat protocol)
 *     /root/reference/fullwave/solver/launcher.py:160-254                      (subprocess.run + genout.dat)
 *     /root/reference/fullwave/solver/solver.py:734-759                        (call site)
 * The reference has no FFI for this path -- the "interface" is that file protocol -- so every entry
 * point below names the protocol element or reference behaviour it replaces.  INTEGRATION.md shows
 * the ctypes stub a maintainer would add to fullwave/solver/launcher.py.
 *
 * Plain C types only: no torch, no C++ in the signatures.  All arrays are little-endian,
 * C-contiguous; float maps are float32 over the EXTENDED grid [nX][nY][nZ] (nZ = 1 in 2D),
 * idx = (x*nY + y)*nZ + z, exactly as the reference writes them (input_file_writer.py:870-881).
 */
#ifndef FW25_H
#define FW25_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FW25_ABI_VERSION 6
#define FW25_M 8 /* stencil half-width: solver.py:296 (m_spatial_order = 8), kernels launched with M = 8 */

/* Anisotropic-relaxation file set (upstream `use_isotropic_relaxation=False`, input_file_writer.py:592-620; engine
 * family fullwave2_{2d,3d}_2_relax_multi_gpu_*): the same update with one kappa / a / b map PER AXIS.  Index 0, 1, 2 =
 * axis x, y, z (2D: 0, 1); [nu] = relaxation mechanism 1, 2.  Velocity sweep (fd_u): files kappa{x,y,z}.dat,
 * apml{x,y,z}{1,2}.dat, bpml{x,y,z}{1,2}.dat.  Pressure sweep (fd_p): kappa{u,w}.dat (2D: axes x, y) or
 * kappa{u,v,w}.dat (3D: axes x, y, z), apml/bpml likewise.  Same layout and residency rules as the isotropic maps. */
typedef struct fw25_aniso {
  const float *kappa_vel[3], *a_vel[3][2], *b_vel[3][2];
  const float *kappa_prs[3], *a_prs[3][2], *b_prs[3][2];
} fw25_aniso;

/* One simulation = the contents of one reference "simulation_dir".
 * Field names are the reference's .dat file stems (input_file_writer.py:766-821, :581-627). */
typedef struct fw25_problem {
  int32_t ndim;               /* 2 or 3 (reference: picks the 2d/ or 3d/ binary, solver.py:177-273) */
  int32_t nX, nY, nZ;         /* nX.dat nY.dat nZ.dat -- planes held by THIS problem (see fw25_slab) */
  int32_t nT, nTic, modT;     /* nT.dat nTic.dat modT.dat */
  int32_t ndmap;              /* ndmap.dat */
  float dX, dT;               /* dX.dat dT.dat (dY, dZ, c0, c.dat, d.dat are read by the reference but unused) */
  const float *rho, *K, *beta;                    /* rho.dat K.dat beta.dat */
  const float *kappax, *kappau;                   /* kappax.dat kappau.dat */
  const float *apmlx1, *bpmlx1, *apmlx2, *bpmlx2; /* feed the velocity sweep (fd_u) */
  const float *apmlu1, *bpmlu1, *apmlu2, *bpmlu2; /* feed the pressure sweep (fd_p) */
  const float *dmap;          /* dmap.dat  float32 [9][2][ndmap] */
  const int32_t *dcmap;       /* dcmap.dat int32   [nX*nY*nZ], 0-based */
  int32_t ncoords;            /* ncoords.dat */
  const int32_t *icc;         /* icc.dat   int32 [ncoords][ndim]   rows (x,y[,z]), GLOBAL coordinates */
  const float *icmat;         /* icmat.dat float32 [ncoords][nTic] */
  int32_t ncoordsout;         /* ncoordsout.dat */
  const int32_t *outc;        /* outc.dat  int32 [ncoordsout][ndim], GLOBAL coordinates */
  int32_t ncoordszero;        /* ncoordszero.dat */
  const int32_t *icczero;     /* icczero.dat int32 [ncoordszero][ndim], GLOBAL coordinates (air voxels) */
  /* Extensions (zero = reference behaviour):                                                     */
  int32_t maps_on_device;     /* 1: the 13 float maps + dcmap are DEVICE pointers on the engine's GPU */
  int32_t map_pitch;          /* floats per row (fastest axis) of the 13 maps + dcmap; 0 = dense (nZ, or
                                 nY in 2D).  Device maps with map_pitch == fw25_pitch(n_fast) are
                                 adopted without a copy (dcmap only when dcmap_full3d != 0). */
  int32_t dcmap_full3d;       /* 0: what the reference's 3D binary does -- its main() loads only the first
                                 nX*nY entries of dcmap.dat into a zeroed array (count = nX*nY at ASM
                                 0x40648c-0x4064c3; c.dat next to it gets nX*nY*nZ), so every 3D voxel
                                 whose flat index is >= nX*nY uses stencil-table column 0.  Verified
                                 bit-exactly on a B200 (DESIGN.md "dcmap").  1: honour dcmap per voxel
                                 (the documented intent; differs from the binary by ~4e-4 rel-L2). */
  float *ext_p, *ext_u, *ext_v, *ext_w; /* optional caller-owned DEVICE state arrays [nX][nY][pitch]
                                 (pitch = fw25_pitch(nZ)); lets a multi-process driver hand the halo
                                 planes to NCCL without a copy.  NULL: the engine allocates. */
  const fw25_aniso *aniso;    /* NULL: isotropic file set (above).  Else the per-axis maps; kappax .. bpmlu2 above are
                                 ignored, and -- like the reference's anisotropic binaries, which have no
                                 inject_source_zero kernel -- so are the air voxels (icczero).  When every axis holds
                                 identical values (what the reference's own Python layer writes, pml_builder.py:
                                 896-1005) the engine runs its isotropic kernels on one copy. */
  const int32_t *out_box;     /* optional (HOST pointer, 2*ndim ints: lo[ndim], hi[ndim]): the sensors are EVERY point
                                 of the box [lo, hi), in row-major order -- what a rectangular `Sensor(mask)` yields
                                 (sensor.py:24-50: np.where order) and what `Solver.run(record_whole_domain=True)` asks
                                 for (solver.py:709-731, the whole extended grid).  outc may then be NULL and
                                 ncoordsout must equal the box volume; frames come back in the same global order.
                                 The engine records such sensors without an index list, and
                                 recognises a box by itself in any outc list of >= 4096 coordinates. */
} fw25_problem;

/* x-slab owned by one engine when the grid is sharded along x (the reference's slab partitioner,
 * SURVEY.md 8(e); binary strings "GPU %d: global range [..)").  NULL slab = whole grid. */
typedef struct fw25_slab {
  int32_t nX_global;          /* planes of the whole extended grid */
  int32_t gx0;                /* global x of local plane 0 of the problem's arrays */
  int32_t own_lo, own_hi;     /* global owned range [own_lo, own_hi); requires
                                 gx0 <= max(own_lo-8,0) and gx0+nX >= min(own_hi+8, nX_global) */
} fw25_slab;

typedef struct fw25_stats {
  double setup_ms;            /* allocation + host->device upload + table preprocessing */
  double loop_ms;             /* the time loop, device time (CUDA events) */
  double d2h_ms;              /* genout device->host */
  int64_t kernel_launches;    /* kernels launched inside the time loop */
  int64_t h2d_bytes, d2h_bytes;
  int64_t point_updates;      /* nX*nY*nZ * nT (extended grid, the reference's count) */
  int64_t halo_bytes;         /* bytes moved between x-slabs (all interfaces, both directions); 0 on one device */
  int32_t n_devices;
  int32_t skewed_steps;       /* fw25_run_medium: time steps that ran block-wise under the upload of the medium */
} fw25_stats;

typedef struct fw25_engine fw25_engine; /* opaque */

/* ---- whole-job entry point: replaces `Launcher.run` (launcher.py:160-254) + reading genout.dat.
 * genout: caller-allocated float32 [ceil(nT/modT)][ncoordsout] == the bytes of genout.dat
 * (solver.py:600-618 reshapes it to [n_sensors, n_frames]).  device_ids mirrors CUDA_VISIBLE_DEVICES
 * (launcher.py:63-105, :206): n_devices > 1 shards x-slabs over those GPUs in this process.
 * Returns 0, or non-zero with fw25_last_error() set (the reference: non-zero exit status ->
 * SimulationError, launcher.py:221-241). */
int fw25_run(const fw25_problem *pb, const int32_t *device_ids, int32_t n_devices,
             float *genout, size_t genout_len, fw25_stats *stats);

/* ---- engine handle: what the reference's `main` does between loading the .dat files and the time
 * loop, kept alive so a caller can step, read fields and reuse uploads. */
int fw25_create(const fw25_problem *pb, const fw25_slab *slab, int32_t device, fw25_engine **out);
void fw25_destroy(fw25_engine *e);

/* ---- several transmit events on one medium.  Upstream this is `Solver.run(is_static_map=True,
 * recalculate_pml=False)` in a loop: only icmat.dat is rewritten per event, the maps are symlinked
 * (input_file_writer.py:146-175, :647-714) -- and still re-read and re-uploaded by every launch of the binary.
 * fw25_reset starts the next event on a live engine: wave field zeroed, t = 0, new source list and step counts;
 * maps, stencil tables, sensors stay resident in HBM.  fw25_run_engine then runs steps [t, nT) and writes
 * genout [ceil(nT/modT)][ncoordsout] like fw25_run (whole-grid engines only). */
int fw25_reset(fw25_engine *e, int32_t nT, int32_t nTic, int32_t ncoords, const int32_t *icc, const float *icmat);
int fw25_run_engine(fw25_engine *e, float *genout, size_t genout_len, fw25_stats *stats);

/* One reference time step is: inject(t) -> sweep_u -> sweep_p -> record(t) when t % modT == 0
 * (SURVEY.md 3.3).  x ranges are GLOBAL and are clamped to the engine's owned range. stream is a
 * cudaStream_t (NULL = the engine's own stream). */
int fw25_inject(fw25_engine *e, int32_t t, void *stream);                            /* inject_source + inject_source_zero */
int fw25_sweep_u(fw25_engine *e, int32_t gx_lo, int32_t gx_hi, void *stream);        /* fd_u */
int fw25_sweep_p(fw25_engine *e, int32_t gx_lo, int32_t gx_hi, void *stream);        /* fd_p */
int fw25_record(fw25_engine *e, int32_t frame, void *stream);                        /* compute_genout_frame_multi / extract_pressure_values */
int fw25_step(fw25_engine *e, int32_t n_steps);  /* n whole steps from the engine's current t, on its own stream */
int fw25_sync(fw25_engine *e);
/* n_steps whole steps bracketed by CUDA events on the engine's stream (blocks until done): out[0] = total
 * ms; detail != 0 also brackets every sweep launch: out[1] = sum fd_u ms, out[2] = sum fd_p ms, out[3] = rest
 * (the measurement hook behind bench.py's roofline numbers; the reference has no timers). */
int fw25_step_timed(fw25_engine *e, int32_t n_steps, int32_t detail, double *out4);

/* Results.  fw25_read_frames copies frames [f0, f1) of the owned sensors: out is
 * [f1-f0][fw25_n_local_sensors]; fw25_local_sensor_ids gives their row in the global outc list. */
int32_t fw25_n_local_sensors(const fw25_engine *e);
int fw25_local_sensor_ids(const fw25_engine *e, int32_t *ids);
int fw25_read_frames(fw25_engine *e, int32_t f0, int32_t f1, float *out);
/* name: "p","u","v","w"; out: float32 [nX][nY][nZ] of the engine's local planes (tests). */
int fw25_read_field(fw25_engine *e, const char *name, float *out);
/* device pointer of a state array ("p","u","v","w") and the row pitch (floats) of its layout */
void *fw25_field_ptr(fw25_engine *e, const char *name);
int32_t fw25_pitch(int32_t nZ);
int32_t fw25_current_step(const fw25_engine *e);
int64_t fw25_launch_count(const fw25_engine *e);
/* select the sweep implementation: 0 = auto (best available), 1 = simple (L1/L2-cached loads), 2 = TMA-tiled
 * x-marching, 3 = warp-specialised all-TMA x-marching
 * (3D only; fails with an error if the variant cannot run the problem) */
int fw25_set_kernel_variant(fw25_engine *e, int32_t variant);

/* number of CUDA devices visible to this process: the executable drop-in shards over all of them, like the
 * reference binary does with CUDA_VISIBLE_DEVICES (launcher.py:206; binary: cudaGetDeviceCount in main) */
int32_t fw25_device_count(void);

/* ---- Medium maps built ON THE GPU.  Upstream, every `Solver.run` first builds the engine's
 * 13 coefficient maps on the host in float64 numpy: `PMLBuilder.__init__` pads the user-grid maps by edge
 * replication (pml_builder.py:321-704), `PMLBuilder.run` ramps d / alpha of both relaxation mechanisms towards their
 * PML targets axis by axis (`_apply_pml`, `_apply_pml_3d`, `_apply_transition_and_pml`, :842-1498), turns them into
 * a, b (`_calc_a_and_b`, :794-810), and `InputFileWriter` derives K = c^2 rho, dcmap and casts everything to
 * float32 (input_file_writer.py:95-103, :563-627, :716-745).  fw25_mapgen does all of that in one kernel, from the
 * USER-grid float64 maps, straight into the engine's HBM layout -- same float64 operations in the same order, so
 * rho, K, beta, kappa and dcmap are bit-identical to the reference's files and a, b agree to 1 float32 ulp (exp()).
 * All pointers are HOST pointers to C-contiguous arrays over the user grid [nx][ny][nz] (nz = 1 in 2D). */
typedef struct fw25_medium {
  int32_t ndim;                      /* 2 or 3 */
  int32_t nx, ny, nz;                /* USER grid (Medium.sound_speed.shape) */
  int32_t m_spatial_order;           /* 8 (solver.py:296) */
  int32_t n_pml_layer;               /* PMLBuilder.n_pml_layer        (solver.py:483-486: 3 * ppw) */
  int32_t n_transition_layer;        /* PMLBuilder.n_transition_layer (same default) */
  int32_t use_pml;                   /* 0: `PMLBuilder.run(use_pml=False)` -- pad only, no ramps (pml_builder.py:838-840) */
  double dt;                         /* extended_grid.dt */
  double d_target_pml;               /* pml_builder.py:1063-1070, a host scalar */
  /* the reference's 1-D transition functions sampled by numpy on the host (pml_builder.py:1296-1338):
   * polynomial (d, nu = 1) and linear (alpha, nu = 1) over n_pml + n_transition + 1 points, cosine (nu = 2) over
   * n_transition + 1 points */
  const double *tf_polynomial, *tf_linear, *tf_cosine;
  const double *sound_speed, *density, *beta;
  /* relaxation parameters in the reference's dictionary order (solver/utils.py:68-86):
   * kappa_x1, kappa_x2, d_x1_nu1, alpha_x1_nu1, d_x2_nu1, alpha_x2_nu1, d_x1_nu2, alpha_x1_nu2, d_x2_nu2, alpha_x2_nu2
   * (`MediumRelaxationMaps.relaxation_param_dict`).  All NULL: look them up per voxel from the table below, which is
   * what `Medium.build()` does on the host (utils/relaxation_parameters.py:18-75, :189-242). */
  const double *relax[10];
  const double *alpha_coeff, *alpha_power;   /* user-grid maps for the look-up */
  const double *lut;                         /* database [lut_na][lut_np][10] */
  const double *lut_alpha, *lut_power;       /* alpha_0_list / power_list rounded to 10 decimals, ascending */
  int32_t lut_na, lut_np;
  double alpha_min, alpha_max, power_min, power_max;   /* clip bounds (relaxation_parameters.py:152-155, :216-217) */
  const uint8_t *lut_invalid;                /* invalid_matrix [lut_na][lut_np] or NULL: counted, like the warning */
  int32_t c_round_min;                       /* round(min(c) + 1e-9): dcmap = round(c + 1e-9) - c_round_min */
  int32_t dcmap_full3d;                      /* as fw25_problem.dcmap_full3d: 0 zeroes dcmap beyond the first nX*nY entries */
  int32_t input_f32;                         /* 0: the user-grid maps (sound_speed .. alpha_power) are float64, what the
                                                reference's Medium holds; 1: they are float32 arrays (converted exactly on
                                                the device; half the host->device traffic).  Tables stay float64. */
  int32_t reserved_;
} fw25_medium;

typedef struct fw25_mapset fw25_mapset; /* opaque: the 13 float maps + dcmap, device-resident, engine layout */

/* Builds the maps on `device`.  stats_ms (may be NULL): [0] = host->device upload of the user-grid maps,
 * [1] = the kernel (CUDA events). */
int fw25_mapgen(const fw25_medium *md, int32_t device, fw25_mapset **out, double *stats_ms);
/* One x-slab of the same maps: extended planes [gx0, gx1) only (a rank of an x-sharded run builds its own planes, ghost
 * planes included -- the reference's binary uploads whole maps and slices them per GPU, ASM 0x405f78-0x406204).  `md`
 * describes the WHOLE user grid (nx = its full extent), but its map pointers address host arrays that hold only the
 * user-grid planes [u_plane0, u_plane0 + u_planes); they must cover clamp(gx0 - nb) .. clamp(gx1 - 1 - nb), nb =
 * m_spatial_order + n_pml_layer + n_transition_layer.  The values are those fw25_mapgen gives the same planes. */
int fw25_mapgen_slab(const fw25_medium *md, int32_t device, int32_t gx0, int32_t gx1, int32_t u_plane0, int32_t u_planes,
                     fw25_mapset **out, double *stats_ms);
/* fw25_mapgen_slab in the background.  _begin allocates the set and returns at once: `*view` already carries the final
 * device pointers (fw25_mapset_problem / fw25_create may use them -- creating an engine never reads map contents), while
 * an uploader thread streams the user-grid planes block by block and generates the maps behind them.  _finish waits for
 * the last block, hands the set over (`*out` == the view) and destroys the job; nothing may STEP on the maps before it
 * returns.  The host arrays of `md` must stay alive until then; `md` itself need not.  The job's device scratch (an
 * upload ring of 3 x 32 user-grid planes per map) stays with the set and is released by fw25_mapset_destroy: freeing
 * device memory synchronises the device and would sit between map generation and the first time step. */
typedef struct fw25_mapjob fw25_mapjob;
int fw25_mapgen_slab_begin(const fw25_medium *md, int32_t device, int32_t gx0, int32_t gx1, int32_t u_plane0,
                           int32_t u_planes, fw25_mapset **view, fw25_mapjob **job);
int fw25_mapgen_finish(fw25_mapjob *job, fw25_mapset **out, double *stats_ms);
/* Fills nX nY nZ (the EXTENDED grid; nX = the planes held, for a slab), the 13 map pointers, dcmap, maps_on_device, map_pitch and dcmap_full3d of `pb`
 * so that fw25_create adopts the maps without a copy.  The mapset must outlive every engine created from it. */
int fw25_mapset_problem(const fw25_mapset *ms, fw25_problem *pb);
/* name: a .dat stem ("rho", "K", "beta", "kappax", ..., "bpmlu2", "dcmap"); out: dense [nX][nY][nZ] float32 (int32
 * for dcmap) on the host -- the bytes the reference would have written to <name>.dat. */
int fw25_mapset_read(const fw25_mapset *ms, const char *name, void *out);
int64_t fw25_mapset_invalid_count(const fw25_mapset *ms);   /* voxels that hit an invalid look-up entry */
void fw25_mapset_destroy(fw25_mapset *ms);

/* ---- Whole job from the USER-grid medium in ONE call: what `Solver.run` does between `PMLBuilder.run` and loading
 * genout.dat (solver.py:693-759) -- fw25_mapgen + fw25_run, pipelined.  The user-grid maps go up plane block by plane
 * block (x is the slowest axis of the reference's arrays), the coefficient maps of a block are generated as soon as its
 * planes have landed, and the first time steps already sweep the blocks that are ready -- x-marching sweeps, skewed in
 * time by two blocks per step -- while the rest of the medium is still on its way over PCIe.  Results are bit-identical
 * to fw25_mapgen followed by fw25_run.  pb supplies nT, nTic, modT, ndmap, dX, dT, dmap and the coordinate lists (GLOBAL,
 * extended-grid coordinates); its grid sizes and map pointers are ignored (they follow from md).  One device. */
int fw25_run_medium(const fw25_medium *md, const fw25_problem *pb, int32_t device, float *genout, size_t genout_len,
                    fw25_stats *stats);
/* The same job x-sharded over several GPUs of this process (`cuda_device_id=[0, 1, ...]`, launcher.py:63-105): every
 * device builds its own slab of the maps from the user-grid medium (fw25_mapgen_slab, all devices at once) and the
 * native multi-device runner of fw25_run steps them.  n_devices = 1 is fw25_run_medium. */
int fw25_run_medium_multi(const fw25_medium *md, const fw25_problem *pb, const int32_t *device_ids, int32_t n_devices,
                          float *genout, size_t genout_len, fw25_stats *stats);

const char *fw25_last_error(void);
int32_t fw25_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif
