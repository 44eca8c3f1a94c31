// fw25_internal.h -- host-side declarations shared by the engine translation units.
#pragma once
#include <cuda_runtime.h>
#include <string>
#include "fw25_kernels.cuh"

namespace fw25 {

// Launch with programmatic stream serialization (PDL): inside a stream capture this becomes a programmatic edge of the
// graph.  Only for kernels written for it (pdl_wait() before the first dependent access, fw25_kernels.cuh).
// FW25_PDL=0 turns the attribute off (plain serialization) for A/B runs.
bool pdl_enabled();
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args &&...args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// simple (L1/L2-cached) sweeps: fw25_sweeps_simple.cu.  a_lo/a_hi are LOCAL plane indices.
// aniso: read the per-axis maps (Fields::kv .. bp) instead of the per-sweep ones
void launch_sweep_u_simple(int ndim, const Fields &F, const Geom &G, int a_lo, int a_hi, cudaStream_t st, bool aniso = false);
void launch_sweep_p_simple(int ndim, const Fields &F, const Geom &G, int a_lo, int a_hi, cudaStream_t st, bool aniso = false);

// TMA-tiled x-marching sweeps (3D): fw25_sweeps_tiled.cu.  Return the number of kernels launched.
struct TiledPlan;
bool tiled_supported(int ndim, const Geom &G);
TiledPlan *tiled_plan_create(const Fields &F, const Geom &G, const float *host_dmap, cudaStream_t st, std::string *err);
void tiled_plan_destroy(TiledPlan *pl);
int launch_sweep_u_tiled(const TiledPlan *pl, const Fields &F, const Geom &G, int a_lo, int a_hi, cudaStream_t st);
int launch_sweep_p_tiled(const TiledPlan *pl, const Fields &F, const Geom &G, int a_lo, int a_hi, cudaStream_t st);

// warp-specialised all-TMA sweeps (3D): fw25_sweeps_ws.cu
struct WsPlan;
bool ws_supported(int ndim, const Geom &G);
// aniso: the anisotropic-relaxation family -- per-axis kappa / a / b tiles (Fields::kv .. bp), smaller tiles
WsPlan *ws_plan_create(const Fields &F, const Geom &G, const float *host_dmap, cudaStream_t st, std::string *err,
                       bool aniso = false);
void ws_plan_destroy(WsPlan *pl);
// push != nullptr: fused halo exchange -- the launch also stores its results into the neighbour's ghost planes
int launch_sweep_u_ws(const WsPlan *pl, const Fields &F, const Geom &G, int a_lo, int a_hi, cudaStream_t st,
                      const HaloPush *push = nullptr);
int launch_sweep_p_ws(const WsPlan *pl, const Fields &F, const Geom &G, int a_lo, int a_hi, cudaStream_t st,
                      const HaloPush *push = nullptr);

// TMA-tiled 2D sweeps: fw25_sweeps_2d.cu
struct Plan2D;
bool sweeps2d_supported(int ndim, const Geom &G);
bool sweeps2d_worthwhile(const Geom &G, int rows);   // auto mode: tiled only where it beats the simple sweeps
Plan2D *plan2d_create(const Fields &F, const Geom &G, const float *host_dmap, cudaStream_t st, std::string *err,
                      bool aniso = false);   // aniso: per-axis kappa / a / b maps (Fields::kv .. bp)
void plan2d_destroy(Plan2D *pl);
int launch_sweep_u_2d(const Plan2D *pl, const Fields &F, const Geom &G, int a_lo, int a_hi, cudaStream_t st);
int launch_sweep_p_2d(const Plan2D *pl, const Fields &F, const Geom &G, int a_lo, int a_hi, cudaStream_t st);

// box sensors: local origin (a0, b0, c0) and extent (wa, wb, wc) of the owned part of the box in the engine's layout,
// the local index ranges outside which a point lies in the never-updated rim (reads 0), array strides
struct SensBox {
  int a0, b0, c0, wa, wb, wc;
  int a_lo, a_hi, b_lo, b_hi, c_lo, c_hi;
  long long sA, sB;
};
// Fused 2D step (fw25_sweeps_2d.cu, k_sweep_p_2dc<TR, true>): fd_p(t) also records frame t and applies the injection
// / air zeroing of step t + 1.  Tiles are FUSE_TR rows x 128 columns over the full sweep range [a_rim_lo, a_rim_hi):
// tile = ((a - a_rim_lo) / FUSE_TR) * n_bx + c / 128 with n_bx = (nC - 8 + 127) / 128, cell = ((a - a_rim_lo) % FUSE_TR)
// * 128 + c % 128.  tile_ofs is the CSR over tiles of the special cells (sources, air voxels, listed sensors).
constexpr int FUSE_TR = 2;
constexpr int FUSE_SOURCE = 0, FUSE_AIR = 1, FUSE_SENSOR = 2;     // ent_kind
constexpr int FUSE_RECORD = 1, FUSE_INJECT = 2;                   // launch flags
struct Fuse2D {
  const int *tile_ofs;
  const unsigned short *ent_cell;
  const unsigned char *ent_kind;
  const int *ent_row;            // icmat row (source) / local sensor row (listed sensor)
  const float *icmat;
  int nTic;
  float *frames;                 // the frame ring [cap][n_sens]
  long long n_sens;
  int modT, cap;
  const int *d_t;                // device-side step counter; this launch is step *d_t + t_off
  SensBox box;                   // box sensors (use_box): rows follow from the cell's coordinates
  int use_box;
};
bool sweeps2d_fusable();
int launch_sweep_p_2d_fused(const Plan2D *pl, const Fields &F, const Geom &G, int a_lo, int a_hi, cudaStream_t st,
                            const Fuse2D &X, int t_off, int flags);

// streamed map generation (fw25_mapgen.cu): blocks of `block_planes` extended planes become valid in x order while the
// caller already steps.  mapstream_wait_recorded blocks the HOST until the block's event has been recorded and returns
// it (nullptr + g_err if the uploader failed); mapstream_finish joins the uploader and hands over the map set.
struct MapStream;
}  // namespace fw25
struct fw25_medium;
struct fw25_mapset;
namespace fw25 {
MapStream *mapstream_start(const fw25_medium *md, int device, int block_planes, fw25_mapset **ms_out);   // allocates
void mapstream_go(MapStream *S);                                                                          // starts the uploads
int mapstream_blocks(const MapStream *S);
cudaEvent_t mapstream_wait_recorded(MapStream *S, int block);
fw25_mapset *mapstream_finish(MapStream *S, double *stats_ms, int64_t *h2d_bytes);
void mapstream_bequeath(MapStream *S, fw25_mapset *ms);   // after _finish: the stream's device scratch is freed with the set
void mapstream_join(MapStream *S);      // the uploader thread is gone: nothing reads the caller's host arrays any more
void mapstream_destroy(MapStream *S);   // joins too; frees the upload ring (a synchronising cudaFree), streams, events

// point kernels: fw25_points.cu
void launch_dcmap_mask(int32_t *dcmap, long long cells, int pitch, int nC, int nB, long long first_plane,
                       long long limit, cudaStream_t st);
void launch_inject(float *p, const long long *src_idx, const int *src_row, const unsigned char *src_rim,
                   int n_src, const float *icmat, int nTic, int t, const long long *air_idx, int n_air,
                   cudaStream_t st, const int *d_t = nullptr);
void launch_record(const float *p, const long long *sens_idx, int n_sens, float *frame, cudaStream_t st);
// graph-replayed forms: the step number is *d_t + t_off (device-side counter, advanced by launch_tick)
void launch_record_dev(const float *p, const long long *sens_idx, int n_sens, float *frames, const int *d_t, int t_off,
                       int modT, int cap, cudaStream_t st);
void launch_record_box(const float *p, float *frames, long long n_sens, const int *d_t, int t_off, int modT, int cap,
                       const SensBox &B, cudaStream_t st);
// plane-range forms (time-skewed start): entries outside local planes [a_lo, a_hi) are left alone
void launch_inject_range(float *p, const long long *src_idx, const int *src_row, const unsigned char *src_flag, int n_src,
                         const float *icmat, int nTic, int t, const long long *air_idx, int n_air, long long sA, int a_lo,
                         int a_hi, cudaStream_t st);
void launch_record_range(const float *p, const long long *sens_idx, int n_sens, float *frame, long long sA, int a_lo,
                         int a_hi, cudaStream_t st);
void launch_tick(int *d_t, int set, int add, cudaStream_t st);   // *d_t = (set >= 0 ? set : *d_t) + add
int launches_per_inject(int n_src, int n_air, int t, int nTic, int n_src_rim);

}  // namespace fw25
