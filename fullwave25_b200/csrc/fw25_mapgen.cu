// fw25_mapgen.cu -- the engine's 13 coefficient maps + dcmap built on the GPU (include/fw25.h, fw25_mapgen).
//
// What it replaces, upstream (all float64 numpy on one host thread, over the EXTENDED grid):
//   pml_builder.py:321-704   _extend_map_for_pml        pad the user-grid maps by edge replication
//   pml_builder.py:842-1254  _apply_pml / _apply_pml_3d  per key: ramp d / alpha towards the PML target, axis by axis
//   pml_builder.py:1256-1498 _apply_transition_and_pml   one axis: rim + layers := target, 1-D ramp next to it
//   pml_builder.py:794-810   _calc_a_and_b               b = exp(-(d/kappa + alpha) dt), a = d/(kappa (d + kappa alpha) + 1e-10) (b - 1)
//   medium.py:256-259        bulk_modulus                K = c^2 rho
//   input_file_writer.py:95-103, :558-559, :870-881      dcmap = round(c + 1e-9) - round(min c + 1e-9); float32 casts
//   utils/relaxation_parameters.py:18-75                 (optional) nearest-bin look-up of the 10 relaxation parameters
//
// One thread per extended voxel.  The axis-by-axis in-place passes of the reference collapse to a closed form per
// voxel: every ramp takes its "face" value from the first user-grid cell along that axis, and the pad is an edge
// replication, so the value entering the ramps is the user-grid value at the CLAMPED coordinate, and the passes are
//   v <- target                         where the axis index lies in the rim / fully damped layers,
//   v <- v - tf * (v - target)          inside the 1-D ramp,
//   v unchanged                         elsewhere,
// applied for axis 0, 1[, 2] in that order.  Each step is the same float64 operation numpy performs (explicit _rn
// intrinsics, nothing contracted), so d / alpha after the ramps are bit-identical to the reference's and the float32
// maps differ only where exp() differs in its last float64 bit (<= 1 float32 ulp in a, b).
// HBM traffic per extended voxel: <= 13 float64 reads of the (13x smaller, mostly L2-resident) user grid + 14 x 4 B
// written: the kernel is bound by the 56 B/voxel it writes.

#include <climits>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/fw25.h"

namespace fw25 {
extern thread_local std::string g_err;
void reap_wait();   // fw25_run.cu: joins the deferred teardown of a previous whole-job call

namespace {

struct MgFail {
  int code;
};

#define MG_CUDA(expr)                                                                               \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess) {                                                                        \
      char _b[512];                                                                                 \
      snprintf(_b, sizeof _b, "CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__, __LINE__, \
               cudaGetErrorString(_e));                                                             \
      g_err = _b;                                                                                   \
      throw MgFail{2};                                                                              \
    }                                                                                               \
  } while (0)

void mg_fail(const std::string &msg) {
  g_err = msg;
  throw MgFail{1};
}

// ramp codes in the per-axis weight tables
constexpr double W_KEEP = 2.0;     // outside every layer: value unchanged
constexpr double W_TARGET = -1.0;  // rim / fully damped layers: value := target
// anything else: the transition-function sample tf in [0, 1]

struct MgParams {
  int ndim;
  int nA, nB, nC, pitch;   // extended grid in the engine's layout: 3D (x, y, z); 2D (x, -, y) with nB = 1
  int uA, uB, uC;          // user grid, same axis mapping
  int nb;                  // boundary points per side: M + n_pml + n_transition
  int use_pml;
  double dt, d_target;
  // weight tables [3 ramp kinds][axis length], one per engine axis (B unused in 2D): kind 0 polynomial (d nu1),
  // 1 linear (alpha nu1), 2 cosine (d, alpha nu2)
  const double *wA, *wB, *wC;
  // user-grid inputs: float64 (the reference's Medium) or float32 (in_f32), planes [u_plane0, ...) of the user grid
  const void *c, *rho, *beta;
  const void *relax[10];
  const void *alpha_coeff, *alpha_power;
  int in_f32, u_plane0;
  int a0;                  // first extended plane of this launch (blockIdx.z = 0)
  int a_base;              // extended plane stored at index 0 of the outputs (0, or an x-slab's first plane)
  const double *lut, *lut_alpha, *lut_power;
  const unsigned char *lut_invalid;
  int lut_na, lut_np;
  double alpha_min, alpha_max, power_min, power_max;
  int c_round_min;
  long long dcmap_limit;   // flat dense indices >= limit read 0 (reference 3D binary); < 0: no limit
  float *out[13];          // rho K beta kappax kappau apmlx1 bpmlx1 apmlx2 bpmlx2 apmlu1 bpmlu1 apmlu2 bpmlu2
  int32_t *dcmap;
  unsigned long long *invalid_count;
};

__device__ __forceinline__ double ramp(double v, double w, double target) {
  if (w == W_KEEP) return v;
  if (w == W_TARGET) return target;
  return __dsub_rn(v, __dmul_rn(w, __dsub_rn(v, target)));   // up_vals - tf * (up_vals - value_target)
}

// np.searchsorted(list, v) (side = "left") clipped to [0, n - 1]: first i with list[i] >= v
__device__ __forceinline__ int search_left(const double *__restrict__ list, int n, double v) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(list + mid) < v) lo = mid + 1; else hi = mid;
  }
  return min(lo, n - 1);
}

__device__ __forceinline__ double clip(double v, double lo, double hi) { return fmin(fmax(v, lo), hi); }   // np.clip

__device__ __forceinline__ void a_and_b(double d, double kappa, double alpha, double dt, float &a, float &b) {
  // b = np.exp(-(d_x / kappa_x + alpha_x) * dt)
  const double bb = exp(__dmul_rn(-__dadd_rn(__ddiv_rn(d, kappa), alpha), dt));
  // a = d_x / (kappa_x * (d_x + kappa_x * alpha_x) + eps) * (b - 1)
  const double den = __dadd_rn(__dmul_rn(kappa, __dadd_rn(d, __dmul_rn(kappa, alpha))), 1e-10);
  const double aa = __dmul_rn(__ddiv_rn(d, den), __dsub_rn(bb, 1.0));
  a = __double2float_rn(aa);
  b = __double2float_rn(bb);
}

__device__ __forceinline__ double user(const void *ptr, long long i, int in_f32) {
  return in_f32 ? (double)__ldg(static_cast<const float *>(ptr) + i) : __ldg(static_cast<const double *>(ptr) + i);
}

__global__ void __launch_bounds__(256) k_mapgen(const MgParams P) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;   // contiguous axis (incl. the padding columns)
  const int b = blockIdx.y;
  const int a = blockIdx.z + P.a0;
  if (c >= P.pitch) return;
  const long long o = ((long long)(a - P.a_base) * P.nB + b) * P.pitch + c;
  if (c >= P.nC) {   // padding: zeros, like the engine's own uploads
#pragma unroll
    for (int i = 0; i < 13; ++i) P.out[i][o] = 0.0f;
    P.dcmap[o] = 0;
    return;
  }
  const int ua = min(max(a - P.nb, 0), P.uA - 1);
  const int ub = min(max(b - P.nb, 0), P.uB - 1);
  const int uc = min(max(c - P.nb, 0), P.uC - 1);
  const long long u = ((long long)(ua - P.u_plane0) * P.uB + ub) * P.uC + uc;
  const int f32 = P.in_f32;

  const double cs = user(P.c, u, f32), rho = user(P.rho, u, f32);
  P.out[0][o] = __double2float_rn(rho);
  P.out[1][o] = __double2float_rn(__dmul_rn(__dmul_rn(cs, cs), rho));   // np.multiply(sound_speed**2, density)
  P.out[2][o] = __double2float_rn(user(P.beta, u, f32));
  {
    const long long dense = ((long long)a * P.nB + b) * P.nC + c;
    int dc = (int)rint(__dadd_rn(cs, 1e-9)) - P.c_round_min;   // np.round(c + 1e-9): ties to even, like rint
    if (P.dcmap_limit >= 0 && dense >= P.dcmap_limit) dc = 0;
    P.dcmap[o] = dc;
  }

  double r[10];
  if (P.relax[0]) {
#pragma unroll
    for (int i = 0; i < 10; ++i) r[i] = user(P.relax[i], u, f32);
  } else {
    const double al = clip(user(P.alpha_coeff, u, f32), P.alpha_min, P.alpha_max);
    const double pw = clip(user(P.alpha_power, u, f32), P.power_min, P.power_max);
    const int ia = search_left(P.lut_alpha, P.lut_na, al);
    const int ip = search_left(P.lut_power, P.lut_np, pw);
    const long long e = (long long)ia * P.lut_np + ip;
    if (P.lut_invalid && P.lut_invalid[e]) atomicAdd(P.invalid_count, 1ULL);
#pragma unroll
    for (int i = 0; i < 10; ++i) r[i] = __ldg(P.lut + e * 10 + i);
  }

  if (P.use_pml) {
    // axis order 0, 1[, 2] == engine axes A, (B,) C; 2D has no B
    const double *w[3] = {P.wA, P.ndim == 3 ? P.wB : nullptr, P.wC};
    const int n[3] = {P.nA, P.nB, P.nC};
    const int at[3] = {a, b, c};
#pragma unroll
    for (int ax = 0; ax < 3; ++ax) {
      if (!w[ax]) continue;
      const double wp = __ldg(w[ax] + at[ax]), wl = __ldg(w[ax] + n[ax] + at[ax]), wc = __ldg(w[ax] + 2 * n[ax] + at[ax]);
      r[2] = ramp(r[2], wp, P.d_target);  r[4] = ramp(r[4], wp, P.d_target);   // d_x1_nu1, d_x2_nu1: polynomial -> d_target_pml
      r[3] = ramp(r[3], wl, 0.0);         r[5] = ramp(r[5], wl, 0.0);          // alpha_*_nu1: linear -> 0
      r[6] = ramp(r[6], wc, 0.0);         r[8] = ramp(r[8], wc, 0.0);          // d_*_nu2: cosine -> 0
      r[7] = ramp(r[7], wc, 0.0);         r[9] = ramp(r[9], wc, 0.0);          // alpha_*_nu2: cosine -> 0
    }
  }
  // kappa_x <- kappa_x2, kappa_u <- kappa_x1; "x" family (velocity sweep) <- x2, "u" family (pressure sweep) <- x1
  P.out[3][o] = __double2float_rn(r[1]);
  P.out[4][o] = __double2float_rn(r[0]);
  float fa, fb;
  a_and_b(r[4], r[1], r[5], P.dt, fa, fb);  P.out[5][o] = fa;  P.out[6][o] = fb;    // apmlx1, bpmlx1
  a_and_b(r[8], r[1], r[9], P.dt, fa, fb);  P.out[7][o] = fa;  P.out[8][o] = fb;    // apmlx2, bpmlx2
  a_and_b(r[2], r[0], r[3], P.dt, fa, fb);  P.out[9][o] = fa;  P.out[10][o] = fb;   // apmlu1, bpmlu1
  a_and_b(r[6], r[0], r[7], P.dt, fa, fb);  P.out[11][o] = fa; P.out[12][o] = fb;   // apmlu2, bpmlu2
}

// Per-axis weight table of one ramp kind (pml_builder.py:1282-1294, :1340-1375): n = extended axis length,
// thickness / offset in cells, tf = the transition function sampled on thickness + 1 points.
void fill_weights(double *w, int n, int m, int thickness, int offset, const double *tf) {
  for (int i = 0; i < n; ++i) w[i] = W_KEEP;
  const int lo_set = m + offset + thickness;                  // input_array[: M + offset + thickness] = target
  const int hi_set = n - m - thickness - offset;              // input_array[n - M - thickness - offset :] = target
  for (int i = 0; i < n; ++i)
    if (i < lo_set || i >= hi_set) w[i] = W_TARGET;
  const int up_start = m + offset - 1, up_end = m + offset + thickness;
  for (int i = up_start; i < up_end; ++i)                     // reversed transition function
    if (i >= 0 && i < n) w[i] = tf[thickness - (i - up_start)];
  const int down_start = n - m - thickness - offset - 1, down_end = n - m - offset;
  for (int i = down_start; i < down_end; ++i)                 // forward transition function (written second: it wins)
    if (i >= 0 && i < n) w[i] = tf[i - down_start];
}

}  // namespace
}  // namespace fw25

using namespace fw25;

struct fw25_mapset {
  int device = 0;
  int ndim = 3, nX = 0, nY = 0, nZ = 1, pitch = 0;   // nX: planes held (an x-slab: fewer than the extended grid's)
  int x0 = 0;                                          // extended plane of the first plane held
  int dcmap_full3d = 1;
  size_t cells = 0;
  float *block = nullptr;    // one allocation: the 13 maps, then dcmap
  float *maps[13] = {};
  int32_t *dcmap = nullptr;
  long long invalid = 0;
  std::vector<void *> attic;   // device scratch of the job that built the set (upload ring, tables), released with the set:
                               // a cudaFree synchronises the device and would sit in front of the run's first step
  ~fw25_mapset() {
    cudaSetDevice(device);
    for (void *p : attic) cudaFree(p);
    if (block) cudaFree(block);
  }
};

static const char *const kMapNames[13] = {"rho",    "K",      "beta",   "kappax", "kappau", "apmlx1", "bpmlx1",
                                          "apmlx2", "bpmlx2", "apmlu1", "bpmlu1", "apmlu2", "bpmlu2"};

namespace fw25 {
namespace {

// A validated medium: kernel parameters, the small tables on the device, the output mapset.  The user-grid maps
// themselves are uploaded by the caller -- all at once (fw25_mapgen) or plane block by plane block (MapStream).
struct MapgenPlan {
  MgParams P{};
  std::unique_ptr<fw25_mapset> ms;
  void *tables = nullptr;                  // device: weight tables, look-up database, invalid counter
  int n_user = 0;                          // user-grid maps: c, rho, beta + 10 relaxation maps or alpha_coeff, alpha_power
  const void *host_user[13] = {};
  size_t slot_off[13] = {};                // byte offset inside MgParams of each user map's pointer
  size_t elem = 8;                         // bytes per user-grid value
  size_t user_plane = 0;                   // values per user-grid x plane

  ~MapgenPlan() {
    if (tables) cudaFree(tables);
  }

  // a_first / a_count: the extended x planes to hold ([0, nA) by default; an x-slab otherwise)
  void create(const fw25_medium *md, int device, cudaStream_t st, int a_first = 0, int a_count = -1) {
    if (md->ndim != 2 && md->ndim != 3) mg_fail("fw25_mapgen: ndim must be 2 or 3");
    const int ndim = md->ndim;
    const int nz_u = ndim == 3 ? md->nz : 1;
    if (md->nx <= 0 || md->ny <= 0 || nz_u <= 0) mg_fail("fw25_mapgen: user grid dimensions must be positive");
    const int m = md->m_spatial_order, npml = md->n_pml_layer, ntr = md->n_transition_layer;
    if (m < 0 || npml < 0 || ntr < 0) mg_fail("fw25_mapgen: negative layer count");
    if (md->use_pml && ntr == 0)   // pml_builder.py:1275-1281 (the nu = 2 keys transit within the transition layer)
      mg_fail("Transition layer is not defined. Set transit_within_transition_layer to False or define n_transition_layer.");
    if (!md->sound_speed || !md->density || !md->beta) mg_fail("fw25_mapgen: sound_speed / density / beta is NULL");
    const bool direct = md->relax[0] != nullptr;
    if (direct) {
      for (int i = 0; i < 10; ++i)
        if (!md->relax[i]) mg_fail("fw25_mapgen: a relaxation-parameter map is NULL");
    } else {
      if (!md->alpha_coeff || !md->alpha_power || !md->lut || !md->lut_alpha || !md->lut_power || md->lut_na <= 0 ||
          md->lut_np <= 0)
        mg_fail("fw25_mapgen: neither relaxation maps nor a complete look-up table were given");
    }
    if (md->use_pml && (!md->tf_polynomial || !md->tf_linear || !md->tf_cosine))
      mg_fail("fw25_mapgen: a transition-function table is NULL");
    const int nb = m + npml + ntr;
    const long long ex = md->nx + 2LL * nb, ey = md->ny + 2LL * nb, ez = ndim == 3 ? nz_u + 2LL * nb : 1;
    if (ex > INT32_MAX || ey > INT32_MAX || ez > INT32_MAX) mg_fail("fw25_mapgen: extended grid too large");

    MG_CUDA(cudaSetDevice(device));
    ms.reset(new fw25_mapset());
    ms->device = device;
    if (a_count < 0) a_count = (int)ex - a_first;
    if (a_first < 0 || a_count <= 0 || a_first + (long long)a_count > ex) mg_fail("fw25_mapgen: plane range outside the extended grid");
    ms->ndim = ndim; ms->nX = a_count; ms->x0 = a_first; ms->nY = (int)ey; ms->nZ = (int)ez;
    ms->dcmap_full3d = md->dcmap_full3d != 0 || ndim == 2;

    P.ndim = ndim;
    P.nA = (int)ex; P.nB = ndim == 3 ? (int)ey : 1; P.nC = ndim == 3 ? (int)ez : (int)ey;
    P.uA = md->nx; P.uB = ndim == 3 ? md->ny : 1; P.uC = ndim == 3 ? nz_u : md->ny;
    P.pitch = fw25_pitch(P.nC);
    P.a_base = a_first;
    ms->pitch = P.pitch;
    P.nb = nb; P.use_pml = md->use_pml != 0;
    P.dt = md->dt; P.d_target = md->d_target_pml;
    P.c_round_min = md->c_round_min;
    P.dcmap_limit = (ndim == 3 && !md->dcmap_full3d) ? ex * ey : -1;
    P.lut_na = md->lut_na; P.lut_np = md->lut_np;
    P.alpha_min = md->alpha_min; P.alpha_max = md->alpha_max; P.power_min = md->power_min; P.power_max = md->power_max;
    P.in_f32 = md->input_f32 != 0;
    elem = P.in_f32 ? 4 : 8;
    user_plane = (size_t)P.uB * P.uC;
    if (P.nB > 65535 || P.nA > 65535) mg_fail("fw25_mapgen: more than 65535 rows per axis");

    auto user_map = [&](const void *host, const void **slot) {
      host_user[n_user] = host;
      slot_off[n_user] = (size_t)(reinterpret_cast<const char *>(slot) - reinterpret_cast<const char *>(&P));
      ++n_user;
    };
    user_map(md->sound_speed, &P.c);
    user_map(md->density, &P.rho);
    user_map(md->beta, &P.beta);
    if (direct) {
      for (int i = 0; i < 10; ++i) user_map(md->relax[i], &P.relax[i]);
    } else {
      user_map(md->alpha_coeff, &P.alpha_coeff);
      user_map(md->alpha_power, &P.alpha_power);
    }

    // the small tables: one allocation, copied before anything else on `st`
    struct Item { const void *host; size_t bytes; const void **slot; };
    std::vector<Item> items;
    auto want = [&](const void *host, size_t bytes, const void **slot) { items.push_back({host, bytes, slot}); };
    if (!direct) {
      want(md->lut, (size_t)md->lut_na * md->lut_np * 10 * 8, (const void **)&P.lut);
      want(md->lut_alpha, (size_t)md->lut_na * 8, (const void **)&P.lut_alpha);
      want(md->lut_power, (size_t)md->lut_np * 8, (const void **)&P.lut_power);
      if (md->lut_invalid) want(md->lut_invalid, (size_t)md->lut_na * md->lut_np, (const void **)&P.lut_invalid);
    }
    std::vector<double> w[3];                       // per-axis weight tables, alive until the copies are done
    if (P.use_pml) {
      const int len[3] = {P.nA, P.nB, P.nC};
      const double **slot[3] = {&P.wA, &P.wB, &P.wC};
      for (int ax = 0; ax < 3; ++ax) {
        if (ndim == 2 && ax == 1) continue;
        w[ax].resize((size_t)3 * len[ax]);
        fill_weights(w[ax].data(), len[ax], m, npml + ntr, 0, md->tf_polynomial);
        fill_weights(w[ax].data() + len[ax], len[ax], m, npml + ntr, 0, md->tf_linear);
        fill_weights(w[ax].data() + 2 * (size_t)len[ax], len[ax], m, ntr, npml, md->tf_cosine);
        want(w[ax].data(), w[ax].size() * 8, (const void **)slot[ax]);
      }
    }
    const unsigned long long zero = 0;
    want(&zero, 8, (const void **)&P.invalid_count);
    size_t total = 0;
    for (auto &it : items) total += (it.bytes + 255) / 256 * 256;
    MG_CUDA(cudaMalloc(&tables, total));
    size_t off = 0;
    for (auto &it : items) {
      char *d = static_cast<char *>(tables) + off;
      MG_CUDA(cudaMemcpyAsync(d, it.host, it.bytes, cudaMemcpyHostToDevice, st));
      *it.slot = d;
      off += (it.bytes + 255) / 256 * 256;
    }
    MG_CUDA(cudaStreamSynchronize(st));             // the host vectors above go out of scope

    ms->cells = (size_t)a_count * P.nB * P.pitch;
    MG_CUDA(cudaMalloc((void **)&ms->block, ms->cells * 4 * 14));
    for (int i = 0; i < 13; ++i) P.out[i] = ms->maps[i] = ms->block + (size_t)i * ms->cells;
    P.dcmap = ms->dcmap = reinterpret_cast<int32_t *>(ms->block + (size_t)13 * ms->cells);
  }

  // extended planes [a_lo, a_hi) from user-grid planes that start at u_plane0 in the buffers `dev_user[i]`
  void launch(int a_lo, int a_hi, void *const *dev_user, int u_plane0, cudaStream_t st) {
    if (a_hi <= a_lo) return;
    MgParams Q = P;
    for (int i = 0; i < n_user; ++i) *reinterpret_cast<const void **>(reinterpret_cast<char *>(&Q) + slot_off[i]) = dev_user[i];
    Q.u_plane0 = u_plane0;
    Q.a0 = a_lo;
    dim3 grid((Q.pitch + 255) / 256, Q.nB, a_hi - a_lo);
    k_mapgen<<<grid, 256, 0, st>>>(Q);
    MG_CUDA(cudaGetLastError());
  }

  // user planes the extended planes [a_lo, a_hi) read: [first, last]
  void user_range(int a_lo, int a_hi, int &first, int &last) const {
    first = std::min(std::max(a_lo - P.nb, 0), P.uA - 1);
    last = std::min(std::max(a_hi - 1 - P.nb, 0), P.uA - 1);
  }

  long long read_invalid(cudaStream_t st) {
    unsigned long long inv = 0;
    MG_CUDA(cudaMemcpyAsync(&inv, P.invalid_count, 8, cudaMemcpyDeviceToHost, st));
    MG_CUDA(cudaStreamSynchronize(st));
    return (long long)inv;
  }
};

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// MapStream: the maps become valid plane block by plane block, in x order, while the caller already steps.  An
// uploader thread sends the user-grid planes each block reads (a ring of three slots in HBM, the whole user grid
// never sits on the device) and launches k_mapgen for the block on its own stream; ready[b] is recorded behind it.
struct MapStream {
  MapgenPlan plan;
  int device = 0, block = 32, n_blocks = 0;
  cudaStream_t up = nullptr, gen = nullptr;
  std::vector<cudaEvent_t> ready, landed, slot_free;
  static constexpr int kSlots = 3;
  char *ring = nullptr;
  size_t slot_bytes = 0, map_bytes = 0;   // per slot; per user map inside a slot
  std::thread th;
  std::mutex m;
  std::condition_variable cv;
  int recorded = 0;                        // events ready[0 .. recorded) have been recorded
  bool failed = false;
  std::string err;
  double upload_ms = 0, kernel_ms = 0;
  int64_t h2d_bytes = 0;
  cudaEvent_t t_begin = nullptr, t_end = nullptr;
  int a_first = 0, a_count = 0;            // extended planes built (an x-slab, or the whole grid)
  int u_plane0 = 0;                        // user-grid plane the host arrays start at

  ~MapStream() {
    if (th.joinable()) th.join();
    cudaSetDevice(device);
    for (auto &v : {&ready, &landed, &slot_free})
      for (cudaEvent_t e : *v)
        if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : {t_begin, t_end})
      if (e) cudaEventDestroy(e);
    if (up) { cudaStreamSynchronize(up); cudaStreamDestroy(up); }
    if (gen) { cudaStreamSynchronize(gen); cudaStreamDestroy(gen); }
    if (ring) cudaFree(ring);
  }

  // gx0 / gx1 (gx1 < 0: the whole grid): the extended planes to build; u0 / un (un < 0: the whole user grid): the
  // user-grid planes the host arrays of `md` hold
  void start(const fw25_medium *md, int dev, int block_planes, int gx0 = 0, int gx1 = -1, int u0 = 0, int un = -1) {
    device = dev;
    block = block_planes;
    MG_CUDA(cudaSetDevice(device));
    MG_CUDA(cudaStreamCreateWithFlags(&up, cudaStreamNonBlocking));
    // the generator kernel shares the GPU with the sweeps of blocks that are already valid: its CTAs go first
    int lo_pri = 0, hi_pri = 0;
    MG_CUDA(cudaDeviceGetStreamPriorityRange(&lo_pri, &hi_pri));
    MG_CUDA(cudaStreamCreateWithPriority(&gen, cudaStreamNonBlocking, hi_pri));
    plan.create(md, device, up, gx0, gx1 < 0 ? -1 : gx1 - gx0);
    a_first = plan.ms->x0;
    a_count = plan.ms->nX;
    if (un < 0) { u0 = 0; un = plan.P.uA; }
    u_plane0 = u0;
    {
      int first, last;
      plan.user_range(a_first, a_first + a_count, first, last);
      if (first < u0 || last >= u0 + un)
        mg_fail("fw25_mapgen_slab: the host arrays do not hold the user-grid planes this slab reads");
    }
    n_blocks = (a_count + block - 1) / block;
    map_bytes = (size_t)block * plan.user_plane * plan.elem;
    map_bytes = (map_bytes + 255) / 256 * 256;
    slot_bytes = map_bytes * plan.n_user;
    MG_CUDA(cudaMalloc((void **)&ring, slot_bytes * kSlots));
    ready.assign(n_blocks, nullptr);
    landed.assign(n_blocks, nullptr);
    slot_free.assign(n_blocks, nullptr);
    for (int b = 0; b < n_blocks; ++b) {
      MG_CUDA(cudaEventCreateWithFlags(&ready[b], cudaEventDisableTiming));
      MG_CUDA(cudaEventCreateWithFlags(&landed[b], cudaEventDisableTiming));
    }
    MG_CUDA(cudaEventCreate(&t_begin));
    MG_CUDA(cudaEventCreate(&t_end));
  }
  void go() { th = std::thread([this] { body(); }); }

  void body() {
    try {
      MG_CUDA(cudaSetDevice(device));
      MG_CUDA(cudaEventRecord(t_begin, up));
      for (int b = 0; b < n_blocks; ++b) {
        const int a_lo = a_first + b * block, a_hi = std::min(a_lo + block, a_first + a_count);
        int u0, u1;
        plan.user_range(a_lo, a_hi, u0, u1);
        char *slot = ring + (size_t)(b % kSlots) * slot_bytes;
        // At most three blocks of copies are queued ahead: the copy engine serves its queue in order, and other
        // host->device traffic of the process must not wait behind the whole medium.
        if (b >= 3) MG_CUDA(cudaEventSynchronize(landed[b - 3]));
        if (b >= kSlots) MG_CUDA(cudaStreamWaitEvent(up, ready[b - kSlots], 0));   // the slot's last reader is done
        void *dev_user[13];
        const size_t bytes = (size_t)(u1 - u0 + 1) * plan.user_plane * plan.elem;
        for (int i = 0; i < plan.n_user; ++i) {
          dev_user[i] = slot + (size_t)i * map_bytes;
          MG_CUDA(cudaMemcpyAsync(dev_user[i], static_cast<const char *>(plan.host_user[i]) +
                                  (size_t)(u0 - u_plane0) * plan.user_plane * plan.elem, bytes, cudaMemcpyHostToDevice, up));
        }
        h2d_bytes += (int64_t)bytes * plan.n_user;
        MG_CUDA(cudaEventRecord(landed[b], up));
        MG_CUDA(cudaStreamWaitEvent(gen, landed[b], 0));
        plan.launch(a_lo, a_hi, dev_user, u0, gen);
        MG_CUDA(cudaEventRecord(ready[b], gen));
        {
          std::lock_guard<std::mutex> lk(m);
          recorded = b + 1;
        }
        cv.notify_all();
      }
      MG_CUDA(cudaEventRecord(t_end, up));
    } catch (const MgFail &) {
      std::lock_guard<std::mutex> lk(m);
      failed = true;
      err = g_err;                                   // (thread-local: hand it to the waiting thread)
      cv.notify_all();
    }
  }

  // blocks until ready[b] has been RECORDED (waiting on an unrecorded event would be a no-op); nullptr on failure
  cudaEvent_t wait_recorded(int b) {
    std::unique_lock<std::mutex> lk(m);
    cv.wait(lk, [&] { return failed || recorded > b; });
    if (failed) { g_err = err; return nullptr; }
    return ready[b];
  }

  void finish(fw25_mapset *ms) {
    if (th.joinable()) th.join();
    if (failed) { g_err = err; throw MgFail{2}; }
    MG_CUDA(cudaSetDevice(device));
    ms->invalid = plan.read_invalid(gen);
    MG_CUDA(cudaStreamSynchronize(up));
    float a = 0;
    MG_CUDA(cudaEventElapsedTime(&a, t_begin, t_end));
    upload_ms = a;
  }
};

MapStream *mapstream_start(const fw25_medium *md, int device, int block_planes, fw25_mapset **ms_out) {
  reap_wait();
  std::unique_ptr<MapStream> S(new MapStream());
  try {
    S->start(md, device, block_planes);
  } catch (const MgFail &) {
    return nullptr;
  }
  *ms_out = S->plan.ms.get();                        // owned by the stream until mapstream_release_mapset
  return S.release();
}
void mapstream_go(MapStream *S) { S->go(); }
int mapstream_blocks(const MapStream *S) { return S->n_blocks; }
cudaEvent_t mapstream_wait_recorded(MapStream *S, int b) { return S->wait_recorded(b); }
fw25_mapset *mapstream_finish(MapStream *S, double *stats_ms, int64_t *h2d_bytes) {
  fw25_mapset *ms = S->plan.ms.get();
  try {
    S->finish(ms);
  } catch (const MgFail &) {
    return nullptr;
  }
  if (stats_ms) { stats_ms[0] = S->upload_ms; stats_ms[1] = 0; }
  if (h2d_bytes) *h2d_bytes = S->h2d_bytes;
  return S->plan.ms.release();
}
void mapstream_bequeath(MapStream *S, fw25_mapset *ms) {
  if (S->ring) { ms->attic.push_back(S->ring); S->ring = nullptr; }
  if (S->plan.tables) { ms->attic.push_back(S->plan.tables); S->plan.tables = nullptr; }
}
void mapstream_join(MapStream *S) {
  if (S->th.joinable()) S->th.join();
}
void mapstream_destroy(MapStream *S) { delete S; }

}  // namespace fw25

extern "C" {

// extended planes [gx0, gx1) from host arrays that hold the user-grid planes [u_plane0, u_plane0 + u_planes)
static int mapgen_impl(const fw25_medium *md, int32_t device, int gx0, int gx1, int u_plane0, int u_planes,
                       fw25_mapset **out, double *stats_ms) {
  if (!md || !out) { g_err = "fw25_mapgen: NULL argument"; return 1; }
  *out = nullptr;
  fw25::reap_wait();                         // a previous whole-job call may still be giving its memory back
  MapgenPlan plan;
  void *scratch = nullptr;   // device scratch: the user-grid inputs
  cudaStream_t st = nullptr;
  cudaEvent_t ev[3] = {};
  int rc = 0;
  try {
    MG_CUDA(cudaSetDevice(device));
    MG_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    for (auto &e : ev) MG_CUDA(cudaEventCreate(&e));
    plan.create(md, device, st, gx0, gx1 < 0 ? -1 : gx1 - gx0);
    const int a_lo = plan.ms->x0, a_hi = plan.ms->x0 + plan.ms->nX;
    int first, last;
    plan.user_range(a_lo, a_hi, first, last);
    if (u_planes < 0) { u_plane0 = 0; u_planes = plan.P.uA; }
    if (first < u_plane0 || last >= u_plane0 + u_planes)
      mg_fail("fw25_mapgen_slab: the host arrays do not hold the user-grid planes this slab reads");
    // One scratch allocation for the user-grid inputs and one output allocation: a handful of driver calls whatever
    // the number of maps.
    const size_t plane_bytes = plan.user_plane * plan.elem, need = (size_t)(last - first + 1) * plane_bytes;
    const size_t per = (need + 255) / 256 * 256;
    MG_CUDA(cudaMalloc(&scratch, per * plan.n_user));
    MG_CUDA(cudaEventRecord(ev[0], st));
    void *dev_user[13];
    for (int i = 0; i < plan.n_user; ++i) {
      dev_user[i] = static_cast<char *>(scratch) + (size_t)i * per;
      MG_CUDA(cudaMemcpyAsync(dev_user[i], static_cast<const char *>(plan.host_user[i]) + (size_t)(first - u_plane0) * plane_bytes,
                              need, cudaMemcpyHostToDevice, st));
    }
    MG_CUDA(cudaEventRecord(ev[1], st));
    plan.launch(a_lo, a_hi, dev_user, first, st);
    MG_CUDA(cudaEventRecord(ev[2], st));
    plan.ms->invalid = plan.read_invalid(st);
    if (stats_ms) {
      float a = 0, b = 0;
      MG_CUDA(cudaEventElapsedTime(&a, ev[0], ev[1]));
      MG_CUDA(cudaEventElapsedTime(&b, ev[1], ev[2]));
      stats_ms[0] = a; stats_ms[1] = b;
    }
  } catch (const MgFail &f) {
    rc = f.code;
  }
  if (st) { cudaStreamSynchronize(st); cudaStreamDestroy(st); }
  if (scratch) cudaFree(scratch);
  for (auto &e : ev)
    if (e) cudaEventDestroy(e);
  if (rc) return rc;
  *out = plan.ms.release();
  return 0;
}

int fw25_mapgen(const fw25_medium *md, int32_t device, fw25_mapset **out, double *stats_ms) {
  return mapgen_impl(md, device, 0, -1, 0, -1, out, stats_ms);
}

int fw25_mapgen_slab(const fw25_medium *md, int32_t device, int32_t gx0, int32_t gx1, int32_t u_plane0, int32_t u_planes,
                     fw25_mapset **out, double *stats_ms) {
  if (gx1 <= gx0 || u_planes <= 0) { g_err = "fw25_mapgen_slab: empty plane range"; return 1; }
  return mapgen_impl(md, device, gx0, gx1, u_plane0, u_planes, out, stats_ms);
}

// ---- the same, in the background: the set is allocated and its device pointers are valid at once (an engine can be
// created on them), an uploader thread streams the user-grid planes block by block and generates the maps behind them
struct fw25_mapjob {
  fw25::MapStream *S = nullptr;
};

int fw25_mapgen_slab_begin(const fw25_medium *md, int32_t device, int32_t gx0, int32_t gx1, int32_t u_plane0,
                           int32_t u_planes, fw25_mapset **view, fw25_mapjob **job) {
  if (!md || !view || !job) { g_err = "fw25_mapgen_slab_begin: NULL argument"; return 1; }
  *view = nullptr; *job = nullptr;
  if (gx1 <= gx0 || u_planes <= 0) { g_err = "fw25_mapgen_slab_begin: empty plane range"; return 1; }
  fw25::reap_wait();
  std::unique_ptr<fw25::MapStream> S(new fw25::MapStream());
  try {
    S->start(md, device, 32, gx0, gx1, u_plane0, u_planes);
  } catch (const MgFail &f) {
    return f.code;
  }
  *view = S->plan.ms.get();
  S->go();
  *job = new fw25_mapjob{S.release()};
  return 0;
}

int fw25_mapgen_finish(fw25_mapjob *job, fw25_mapset **out, double *stats_ms) {
  if (!job || !out) { g_err = "fw25_mapgen_finish: NULL argument"; return 1; }
  *out = nullptr;
  double st[2] = {0, 0};
  fw25_mapset *ms = fw25::mapstream_finish(job->S, st, nullptr);   // joins the uploader, waits for the last block
  if (ms) fw25::mapstream_bequeath(job->S, ms);                     // the ring and the tables are freed with the set
  fw25::mapstream_destroy(job->S);                                  // (frees the set too if the job failed)
  delete job;
  if (!ms) return 2;
  if (stats_ms) { stats_ms[0] = st[0]; stats_ms[1] = st[1]; }
  *out = ms;
  return 0;
}

int fw25_mapset_problem(const fw25_mapset *ms, fw25_problem *pb) {
  if (!ms || !pb) { g_err = "fw25_mapset_problem: NULL argument"; return 1; }
  pb->ndim = ms->ndim; pb->nX = ms->nX; pb->nY = ms->nY; pb->nZ = ms->nZ;
  const float **slot[13] = {&pb->rho,    &pb->K,      &pb->beta,   &pb->kappax, &pb->kappau, &pb->apmlx1, &pb->bpmlx1,
                            &pb->apmlx2, &pb->bpmlx2, &pb->apmlu1, &pb->bpmlu1, &pb->apmlu2, &pb->bpmlu2};
  for (int i = 0; i < 13; ++i) *slot[i] = ms->maps[i];
  pb->dcmap = ms->dcmap;
  pb->maps_on_device = 1;
  pb->map_pitch = ms->pitch;
  pb->dcmap_full3d = 1;   // the reference binary's truncation, when asked for, is already in the generated dcmap
  pb->aniso = nullptr;
  return 0;
}

int fw25_mapset_read(const fw25_mapset *ms, const char *name, void *out) {
  if (!ms || !name || !out) { g_err = "fw25_mapset_read: NULL argument"; return 1; }
  const void *src = nullptr;
  if (!strcmp(name, "dcmap")) src = ms->dcmap;
  for (int i = 0; i < 13 && !src; ++i)
    if (!strcmp(name, kMapNames[i])) src = ms->maps[i];
  if (!src) { g_err = std::string("fw25_mapset_read: unknown map '") + name + "'"; return 1; }
  try {
    MG_CUDA(cudaSetDevice(ms->device));
    const int nC = ms->ndim == 3 ? ms->nZ : ms->nY;
    const size_t rows = ms->ndim == 3 ? (size_t)ms->nX * ms->nY : (size_t)ms->nX;
    MG_CUDA(cudaMemcpy2D(out, (size_t)nC * 4, src, (size_t)ms->pitch * 4, (size_t)nC * 4, rows, cudaMemcpyDeviceToHost));
  } catch (const MgFail &f) {
    return f.code;
  }
  return 0;
}

int64_t fw25_mapset_invalid_count(const fw25_mapset *ms) { return ms ? ms->invalid : -1; }

void fw25_mapset_destroy(fw25_mapset *ms) { delete ms; }

}  // extern "C"
