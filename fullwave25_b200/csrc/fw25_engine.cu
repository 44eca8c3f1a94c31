// fw25_engine.cu -- engine object + C-ABI (include/fw25.h) of the B200-native Fullwave 2.5 engine.
//
// Replaces what the reference's binary-only `main` does: load maps,
// allocate state, run the time loop inject -> fd_u -> fd_p -> record, return the sensor frames.
// Differences by design: one time level (in-place leapfrog, no proceed_time copies), 64-bit indexing,
// row-padded layout for 16-byte vector / TMA access, coordinate lists resolved to linear indices once.
#include <sys/mman.h>
#include <unistd.h>

#include <algorithm>
#include <mutex>
#include <deque>
#include <condition_variable>
#include <atomic>
#include <chrono>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <unordered_set>
#include <numeric>
#include <vector>

#include "../../include/fw25.h"
#include "fw25_internal.h"

namespace fw25 {

thread_local std::string g_err;

struct Fail {
  int code;
};

#define FW_CUDA(expr)                                                                               \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess) {                                                                        \
      char _b[512];                                                                                 \
      snprintf(_b, sizeof _b, "CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__, __LINE__, \
               cudaGetErrorString(_e));                                                             \
      g_err = _b;                                                                                   \
      throw Fail{2};                                                                                \
    }                                                                                               \
  } while (0)

static void fail(int code, const std::string &msg) {
  g_err = msg;
  throw Fail{code};
}

static inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

// Is the coordinate list exactly the points of a box in row-major order?  (A rectangular `Sensor(mask)`:
// np.where order, sensor.py:24-50.)  box: lo[nd], hi[nd].
static bool detect_box(const int32_t *c, int n, int nd, int32_t *box) {
  long long vol = 1;
  for (int k = 0; k < nd; ++k) {
    box[k] = c[k];
    box[nd + k] = c[(size_t)(n - 1) * nd + k] + 1;
    if (box[nd + k] <= box[k]) return false;
    vol *= box[nd + k] - box[k];
  }
  if (vol != n) return false;
  int32_t cur[3] = {box[0], box[1], nd == 3 ? box[2] : 0};
  for (int i = 0; i < n; ++i) {
    const int32_t *ci = c + (size_t)i * nd;
    for (int k = 0; k < nd; ++k)
      if (ci[k] != cur[k]) return false;
    for (int k = nd - 1; k >= 0; --k) {        // next point, last axis fastest
      if (++cur[k] < box[nd + k]) break;
      cur[k] = box[k];
    }
  }
  return true;
}

struct Engine {
  int device = 0;
  int ndim = 3;
  int nXl = 0, nY = 0, nZ = 1;  // API dims of the local problem
  int nT = 0, nTic = 0, modT = 1;
  int nX_global = 0, gx0 = 0, own_lo = 0, own_hi = 0;
  Geom G{};
  Fields F{};
  cudaStream_t stream = nullptr;
  std::vector<void *> owned;
  size_t cells = 0;  // padded cells per array

  long long *d_src_idx = nullptr;
  int *d_src_row = nullptr;
  unsigned char *d_src_rim = nullptr;
  float *d_icmat = nullptr;
  int n_src = 0, n_src_rim = 0;
  long long *d_air_idx = nullptr;
  int n_air = 0;
  long long *d_sens_idx = nullptr;
  int n_sens = 0, n_sens_global = 0;
  std::vector<int32_t> sens_ids;             // global outc row of each local sensor (box sensors: filled on demand)
  // fused 2D step (k_sweep_p_2dc<2, true>): host copies of the point lists and the per-tile CSR built from them
  std::vector<long long> h_src_idx, h_air_idx, h_sens_idx;
  std::vector<int> h_src_row;
  std::vector<unsigned char> h_src_flag;
  bool fuse_ok = false;
  Fuse2D fuse{};
  std::vector<void *> fuse_owned;
  bool sens_box = false;                     // the sensors are every point of a box: no index list (fw25.h, out_box)
  SensBox box{};
  int sens_first = 0;                        // box sensors: global outc row of local sensor 0 (rows are contiguous)
  std::vector<void *> src_owned;             // source-list allocations (replaced by reset())
  std::unordered_set<long long> air_set;     // linear indices of the air voxels held locally
  float *d_frames = nullptr;
  int frames_cap = 0, n_frames = 0;

  int t = 0;
  int64_t launches = 0;
  int variant = 0;
  int64_t h2d_bytes = 0;
  bool aniso = false;            // per-axis relaxation maps that really differ: ANISO simple sweeps only
  bool aniso_protocol = false;   // the problem came as the anisotropic file set (its binaries ignore air voxels)

  // Whole steps replayed from a CUDA graph (small grids are launch-bound: the 2D examples step in ~10 us).  The step
  // number then lives on the device (*d_t, advanced by the graph's last node) so that one instantiated graph
  // serves every period.
  int *d_t = nullptr;
  int d_t_host = -1;             // value *d_t holds once the stream drains (-1: never set)
  struct StepGraph {
    cudaGraphExec_t exec = nullptr;
    int steps = 0, nodes = 0, frames = 0;
    bool with_inject = false, records = false, fused = false;
    int variant = -1;
  } sg;
  int graph_mode = -1;           // -1: auto (on for grids <= 2^27 cells), 0: off, 1: on
  TiledPlan *plan = nullptr;   // TMA-tiled sweeps (3D)
  WsPlan *ws = nullptr;        // warp-specialised all-TMA sweeps (3D)
  Plan2D *p2d = nullptr;       // TMA-tiled sweeps (2D)

  ~Engine() {
    cudaSetDevice(device);
    if (sg.exec) cudaGraphExecDestroy(sg.exec);
    tiled_plan_destroy(plan);
    ws_plan_destroy(ws);
    plan2d_destroy(p2d);
    if (stream) release_staging();
    for (void *p : fuse_owned) cudaFree(p);
    for (void *p : owned) cudaFree(p);
    for (void *p : src_owned) cudaFree(p);
    if (stream) cudaStreamDestroy(stream);
  }

  std::vector<float *> state_pool;   // state arrays allocated ahead, under the map uploads (init)
  double malloc_ms = 0;        // host time spent inside cudaMalloc (FW25_SETUP_TRACE=1 prints the setup phases)
  template <class T>
  T *dalloc(size_t n) {
    void *p = nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    FW_CUDA(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
    malloc_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    owned.push_back(p);
    return static_cast<T *>(p);
  }

  bool is_rim(int x, int y, int z) const {
    if (x < M || x >= nX_global - M) return true;
    if (y < M || y >= nY - M) return true;
    if (ndim == 3 && (z < M || z >= nZ - M)) return true;
    return false;
  }
  // linear index in the padded local layout of GLOBAL coordinate (x,y,z)
  long long lin(int x, int y, int z) const {
    if (ndim == 3) return (long long)(x - gx0) * G.sA + (long long)y * G.sB + z;
    return (long long)(x - gx0) * G.sA + y;
  }

  // src_pitch: floats per row of the caller's array.  A device array already in the engine's padded layout
  // is adopted as is (no copy) unless `must_copy`.
  const float *upload_map(const float *src, bool on_device, int src_pitch, bool must_copy = false) {
    const size_t rows = (size_t)G.nA * G.nB;
    if (on_device && src_pitch == G.pitch && !must_copy) return src;
    float *dst = dalloc<float>(cells);
    if (G.pitch != G.nC) FW_CUDA(cudaMemsetAsync(dst, 0, cells * sizeof(float), stream));
    if (!on_device && src_pitch == G.nC && G.pitch != G.nC && rows * (size_t)G.nC * 4 >= ((size_t)32 << 20)) {
      upload_dense_rows(dst, src, rows);
    } else {
      FW_CUDA(cudaMemcpy2DAsync(dst, (size_t)G.pitch * 4, src, (size_t)src_pitch * 4, (size_t)G.nC * 4, rows,
                                on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, stream));
    }
    if (!on_device) h2d_bytes += (int64_t)rows * G.nC * 4;
    return dst;
  }

  // Dense host rows into the padded layout.  A pitched host-to-device copy runs at 33 GB/s on a B200 (one DMA
  // descriptor per 5 KB row), a dense one at 54 GB/s (tools/native/probe_h2d.cu): so the rows go up densely into one
  // of two staging buffers and are re-pitched by a device-side 2-D copy on a second stream while the next chunk is
  // in flight.
  struct Staging {
    float *buf[2] = {nullptr, nullptr};
    cudaStream_t s2 = nullptr;
    cudaEvent_t up[2] = {nullptr, nullptr}, placed[2] = {nullptr, nullptr};
    size_t rows_per_chunk = 0;
    int used = 0;
  } stg;
  void upload_dense_rows(float *dst, const float *src, size_t rows) {
    const size_t row_b = (size_t)G.nC * 4;
    if (!stg.buf[0]) {
      size_t chunk_mb = 128;
      if (const char *ev = getenv("FW25_STAGE_MB")) chunk_mb = std::max(1, atoi(ev));     // tests: force many chunks
      stg.rows_per_chunk = std::max<size_t>(1, (chunk_mb << 20) / row_b);
      for (int k = 0; k < 2; ++k) {
        FW_CUDA(cudaMalloc((void **)&stg.buf[k], stg.rows_per_chunk * row_b));
        FW_CUDA(cudaEventCreateWithFlags(&stg.up[k], cudaEventDisableTiming));
        FW_CUDA(cudaEventCreateWithFlags(&stg.placed[k], cudaEventDisableTiming));
      }
      FW_CUDA(cudaStreamCreateWithFlags(&stg.s2, cudaStreamNonBlocking));
    }
    FW_CUDA(cudaEventRecord(stg.placed[0], stream));       // the memset of dst precedes the first placement
    FW_CUDA(cudaStreamWaitEvent(stg.s2, stg.placed[0], 0));
    for (size_t r0 = 0; r0 < rows; r0 += stg.rows_per_chunk) {
      const size_t n = std::min(stg.rows_per_chunk, rows - r0);
      const int b = stg.used & 1;
      if (stg.used >= 2) FW_CUDA(cudaStreamWaitEvent(stream, stg.placed[b], 0));    // buffer b is free again
      FW_CUDA(cudaMemcpyAsync(stg.buf[b], src + r0 * G.nC, n * row_b, cudaMemcpyHostToDevice, stream));
      FW_CUDA(cudaEventRecord(stg.up[b], stream));
      FW_CUDA(cudaStreamWaitEvent(stg.s2, stg.up[b], 0));
      FW_CUDA(cudaMemcpy2DAsync(dst + r0 * G.pitch, (size_t)G.pitch * 4, stg.buf[b], row_b, row_b, n,
                                cudaMemcpyDeviceToDevice, stg.s2));
      FW_CUDA(cudaEventRecord(stg.placed[b], stg.s2));
      ++stg.used;
    }
    for (int b = 0; b < 2; ++b) FW_CUDA(cudaStreamWaitEvent(stream, stg.placed[b], 0));
  }
  void release_staging() {
    if (!stg.buf[0]) return;
    cudaStreamSynchronize(stream);
    cudaStreamSynchronize(stg.s2);
    for (int k = 0; k < 2; ++k) {
      cudaFree(stg.buf[k]); cudaEventDestroy(stg.up[k]); cudaEventDestroy(stg.placed[k]);
      stg.buf[k] = nullptr;
    }
    cudaStreamDestroy(stg.s2);
    stg = Staging{};
  }

  template <class T>
  T *salloc(size_t n) {
    void *p = nullptr;
    FW_CUDA(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
    src_owned.push_back(p);
    return static_cast<T *>(p);
  }

  bool coord_ok(const int32_t *c) const {
    if (c[0] < 0 || c[0] >= nX_global || c[1] < 0 || c[1] >= nY) return false;
    if (ndim == 3 && (c[2] < 0 || c[2] >= nZ)) return false;
    return true;
  }

  // sources: every source whose plane is held locally (ghost planes included, so that the neighbour's copy
  // of an injected cell stays consistent without an extra exchange).
  // flag bit 0: sits in the never-updated rim; bit 1: also an air voxel (zeroing always wins, fw25_points.cu)
  void setup_sources(int ncoords, const int32_t *icc, const float *icmat) {
    const int nd = ndim;
    for (void *p : src_owned) cudaFree(p);
    src_owned.clear();
    n_src = n_src_rim = 0;
    std::vector<long long> idx; std::vector<int> row; std::vector<unsigned char> flag;
    if (ncoords > 0 && (!icc || (!icmat && nTic > 0))) fail(1, "icc / icmat pointer is NULL");
    for (int i = 0; i < ncoords; ++i) {
      const int32_t *c = icc + (size_t)i * nd;
      if (!coord_ok(c)) fail(1, "icc: source coordinate outside the grid");
      if (c[0] < gx0 || c[0] >= gx0 + nXl) continue;
      const long long li = lin(c[0], c[1], nd == 3 ? c[2] : 0);
      idx.push_back(li);
      row.push_back(i);
      const bool r = is_rim(c[0], c[1], nd == 3 ? c[2] : M);
      const bool dead = air_set.count(li) != 0;
      flag.push_back((unsigned char)((r ? 1 : 0) | (dead ? 2 : 0)));
      n_src_rim += (r && !dead);
    }
    n_src = (int)idx.size();
    if (ndim == 2) { h_src_idx = idx; h_src_row = row; h_src_flag = flag; }
    d_src_idx = salloc<long long>(n_src); d_src_row = salloc<int>(n_src); d_src_rim = salloc<unsigned char>(n_src);
    d_icmat = nullptr;
    if (n_src) {
      FW_CUDA(cudaMemcpyAsync(d_src_idx, idx.data(), n_src * sizeof(long long), cudaMemcpyHostToDevice, stream));
      FW_CUDA(cudaMemcpyAsync(d_src_row, row.data(), n_src * sizeof(int), cudaMemcpyHostToDevice, stream));
      FW_CUDA(cudaMemcpyAsync(d_src_rim, flag.data(), n_src, cudaMemcpyHostToDevice, stream));
      const size_t nic = (size_t)ncoords * nTic;
      d_icmat = salloc<float>(nic);
      if (nic) FW_CUDA(cudaMemcpyAsync(d_icmat, icmat, nic * 4, cudaMemcpyHostToDevice, stream));
      h2d_bytes += (int64_t)nic * 4;
    }
    FW_CUDA(cudaStreamSynchronize(stream));  // host vectors go out of scope
  }

  // Next transmit event on the same medium: zero the wave field, t = 0, new source list.  Maps, stencil
  // tables, tensor maps, sensors and the frame ring stay resident.
  void reset(int nT_, int nTic_, int ncoords, const int32_t *icc, const float *icmat) {
    if (nT_ < 0 || nTic_ < 0 || ncoords < 0) fail(1, "reset: negative count");
    FW_CUDA(cudaStreamSynchronize(stream));
    nT = nT_; nTic = nTic_;
    setup_sources(ncoords, icc, icmat);
    build_fuse_lists();
    float *st[16] = {F.p, F.q[0], F.q[1], F.q[2], F.psi[0][0], F.psi[0][1], F.psi[1][0], F.psi[1][1], F.psi[2][0],
                     F.psi[2][1], F.phi[0][0], F.phi[0][1], F.phi[1][0], F.phi[1][1], F.phi[2][0], F.phi[2][1]};
    for (float *a : st)
      if (a) FW_CUDA(cudaMemsetAsync(a, 0, cells * sizeof(float), stream));
    t = 0;
    d_t_host = -1;
    n_frames = nT > 0 ? (nT + modT - 1) / modT : 0;
    if (sg.exec) { cudaGraphExecDestroy(sg.exec); sg.exec = nullptr; }   // the graph holds the old source pointers
  }

  void init(const fw25_problem &pb, const fw25_slab *slab, int dev) {
    const auto tr0 = std::chrono::steady_clock::now();
    device = dev;
    FW_CUDA(cudaSetDevice(device));
    if (pb.ndim != 2 && pb.ndim != 3) fail(1, "ndim must be 2 or 3");
    ndim = pb.ndim;
    nXl = pb.nX; nY = pb.nY; nZ = ndim == 3 ? pb.nZ : 1;
    nT = pb.nT; nTic = pb.nTic; modT = pb.modT;
    if (nXl <= 0 || nY <= 0 || nZ <= 0) fail(1, "grid dimensions must be positive");
    if (modT <= 0) fail(1, "modT must be >= 1");
    if (nT < 0 || nTic < 0) fail(1, "nT / nTic must be >= 0");
    if (pb.ndmap <= 0) fail(1, "ndmap must be >= 1");
    if (pb.ncoords < 0 || pb.ncoordsout < 0 || pb.ncoordszero < 0) fail(1, "negative coordinate count");
    if (slab) {
      nX_global = slab->nX_global; gx0 = slab->gx0; own_lo = slab->own_lo; own_hi = slab->own_hi;
      if (own_lo < 0 || own_hi > nX_global || own_lo > own_hi) fail(1, "bad slab owned range");
      if (gx0 > std::max(own_lo - M, 0) || gx0 + nXl < std::min(own_hi + M, nX_global) || gx0 < 0 ||
          gx0 + nXl > nX_global)
        fail(1, "slab arrays must cover the owned range plus 8 ghost planes per interior side");
    } else {
      nX_global = nXl; gx0 = 0; own_lo = 0; own_hi = nXl;
    }
    FW_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));

    G.nA = nXl;
    G.nB = ndim == 3 ? nY : 1;
    G.nC = ndim == 3 ? nZ : nY;
    G.pitch = round_up(G.nC, 32);   // 128-byte rows: one warp = one line, TMA-legal strides
    G.sB = G.pitch;
    G.sA = (long long)G.nB * G.pitch;
    G.ndmap = pb.ndmap;
    G.dX = pb.dX; G.dT = pb.dT;
    G.a_rim_lo = std::max(own_lo, M) - gx0;
    G.a_rim_hi = std::min(own_hi, nX_global - M) - gx0;
    cells = (size_t)G.nA * G.nB * G.pitch;

    const bool dev_maps = pb.maps_on_device != 0;
    const int mp = pb.map_pitch > 0 ? pb.map_pitch : G.nC;
    if (mp < G.nC) fail(1, "map_pitch is smaller than the fastest axis");
    // anisotropic file set: per-axis maps.  When every axis holds the same values (what the reference's Python
    // layer writes) the isotropic kernels run on the x-axis copy; otherwise the ANISO simple sweeps read all of them.
    const fw25_aniso *an = pb.aniso;
    aniso_protocol = an != nullptr;
    if (an) {
      for (int ax = 0; ax < ndim; ++ax) {
        bool ok = an->kappa_vel[ax] && an->kappa_prs[ax];
        for (int nu = 0; nu < 2; ++nu)
          ok = ok && an->a_vel[ax][nu] && an->b_vel[ax][nu] && an->a_prs[ax][nu] && an->b_prs[ax][nu];
        if (!ok) fail(1, "an anisotropic map pointer is NULL");
      }
      aniso = dev_maps;                      // device maps are not compared
      if (!dev_maps) {
        const size_t bytes = (size_t)nXl * nY * nZ * sizeof(float);
        auto same = [&](const float *a, const float *b) { return a == b || memcmp(a, b, bytes) == 0; };
        for (int ax = 1; ax < ndim && !aniso; ++ax) {
          aniso = !same(an->kappa_vel[ax], an->kappa_vel[0]) || !same(an->kappa_prs[ax], an->kappa_prs[0]);
          for (int nu = 0; nu < 2 && !aniso; ++nu)
            aniso = !same(an->a_vel[ax][nu], an->a_vel[0][nu]) || !same(an->b_vel[ax][nu], an->b_vel[0][nu]) ||
                    !same(an->a_prs[ax][nu], an->a_prs[0][nu]) || !same(an->b_prs[ax][nu], an->b_prs[0][nu]);
        }
      }
    }
    const float *maps[13] = {pb.rho, pb.K, pb.beta,
                             an ? an->kappa_vel[0] : pb.kappax, an ? an->kappa_prs[0] : pb.kappau,
                             an ? an->a_vel[0][0] : pb.apmlx1, an ? an->b_vel[0][0] : pb.bpmlx1,
                             an ? an->a_vel[0][1] : pb.apmlx2, an ? an->b_vel[0][1] : pb.bpmlx2,
                             an ? an->a_prs[0][0] : pb.apmlu1, an ? an->b_prs[0][0] : pb.bpmlu1,
                             an ? an->a_prs[0][1] : pb.apmlu2, an ? an->b_prs[0][1] : pb.bpmlu2};
    for (auto m : maps)
      if (!m) fail(1, "a medium map pointer is NULL");
    if (!pb.dmap || !pb.dcmap) fail(1, "dmap / dcmap pointer is NULL");
    // Host maps: every upload keeps the copy engine busy for tens of milliseconds, every cudaMalloc blocks this thread
    // for about as long per 5 GB.  So the state arrays are allocated BETWEEN the uploads, under the copies in flight
    // (at 800 x 1240 x 1240: 16 x 30 ms that used to follow the last upload).
    const int n_state_needed = (pb.ext_p ? 0 : 1) + (pb.ext_u ? 0 : 1) + (pb.ext_v ? 0 : 1) +
                               (ndim == 3 ? (pb.ext_w ? 0 : 1) + 12 : 8);
    int uploads_done = 0;
    auto up = [&](const float *src) {
      const float *d = upload_map(src, dev_maps, mp);
      ++uploads_done;
      if (!dev_maps)
        while ((int)state_pool.size() < std::min(n_state_needed, (n_state_needed * uploads_done + 12) / 13))
          state_pool.push_back(dalloc<float>(cells));
      return d;
    };
    F.rho = up(maps[0]);
    F.K = up(maps[1]);
    F.beta = up(maps[2]);
    F.kappax = up(maps[3]);
    F.kappau = up(maps[4]);
    F.ax1 = up(maps[5]); F.bx1 = up(maps[6]);
    F.ax2 = up(maps[7]); F.bx2 = up(maps[8]);
    F.au1 = up(maps[9]); F.bu1 = up(maps[10]);
    F.au2 = up(maps[11]); F.bu2 = up(maps[12]);
    for (int slot = 0; slot < 3; ++slot) {   // per-axis slots: alias the per-sweep maps unless truly anisotropic
      F.kv[slot] = F.kappax; F.kp[slot] = F.kappau;
      F.av[slot][0] = F.ax1; F.bv[slot][0] = F.bx1; F.av[slot][1] = F.ax2; F.bv[slot][1] = F.bx2;
      F.ap[slot][0] = F.au1; F.bp[slot][0] = F.bu1; F.ap[slot][1] = F.au2; F.bp[slot][1] = F.bu2;
    }
    if (aniso) {
      for (int ax = 1; ax < ndim; ++ax) {
        const int slot = (ndim == 2) ? 2 : ax;   // 2D: the reference's y is the engine's contiguous axis C
        F.kv[slot] = upload_map(an->kappa_vel[ax], dev_maps, mp);
        F.kp[slot] = upload_map(an->kappa_prs[ax], dev_maps, mp);
        for (int nu = 0; nu < 2; ++nu) {
          F.av[slot][nu] = upload_map(an->a_vel[ax][nu], dev_maps, mp);
          F.bv[slot][nu] = upload_map(an->b_vel[ax][nu], dev_maps, mp);
          F.ap[slot][nu] = upload_map(an->a_prs[ax][nu], dev_maps, mp);
          F.bp[slot][nu] = upload_map(an->b_prs[ax][nu], dev_maps, mp);
        }
      }
    }
    {
      // Reference 3D behaviour: only the first nX*nY entries of dcmap are honoured (fw25.h, dcmap_full3d).
      const bool mask = ndim == 3 && !pb.dcmap_full3d;
      int32_t *dc = const_cast<int32_t *>(reinterpret_cast<const int32_t *>(
          upload_map(reinterpret_cast<const float *>(pb.dcmap), dev_maps, mp, /*must_copy=*/mask)));
      if (mask)
        launch_dcmap_mask(dc, (long long)cells, G.pitch, G.nC, G.nB, gx0, (long long)nX_global * nY, stream);
      F.dcmap = dc;
    }
    {
      float *d = dalloc<float>((size_t)18 * pb.ndmap);
      // dmap is a small host table in both modes
      FW_CUDA(cudaMemcpyAsync(d, pb.dmap, (size_t)18 * pb.ndmap * 4, cudaMemcpyDefault, stream));
      F.dmap = d;
    }
    const bool trace = getenv("FW25_SETUP_TRACE") != nullptr;
    auto trace_point = [&](const char *what) {
      if (!trace) return;
      cudaStreamSynchronize(stream);
      fprintf(stderr, "[fw25 setup] %-28s t = %8.1f ms   (cudaMalloc so far %.1f ms, h2d %.2f GB)\n", what,
              std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tr0).count() , malloc_ms,
              h2d_bytes / 1e9);
    };
    trace_point("maps uploaded");
    // state
    auto state = [&](float *ext) {
      float *d = ext;
      if (!d && !state_pool.empty()) { d = state_pool.back(); state_pool.pop_back(); }
      if (!d) d = dalloc<float>(cells);
      FW_CUDA(cudaMemsetAsync(d, 0, cells * sizeof(float), stream));
      return d;
    };
    F.p = state(pb.ext_p);
    if (ndim == 3) {
      F.q[0] = state(pb.ext_u); F.q[1] = state(pb.ext_v); F.q[2] = state(pb.ext_w);
    } else {  // 2D: reference u -> axis A, reference v -> axis C
      F.q[0] = state(pb.ext_u); F.q[1] = nullptr; F.q[2] = state(pb.ext_v);
    }
    for (int ax = 0; ax < 3; ++ax)
      for (int nu = 0; nu < 2; ++nu) {
        const bool used = ndim == 3 || ax != 1;
        F.psi[ax][nu] = used ? state(nullptr) : nullptr;
        F.phi[ax][nu] = used ? state(nullptr) : nullptr;
      }

    if (!aniso && tiled_supported(ndim, G)) {
      std::string perr;
      std::vector<float> hd((size_t)18 * pb.ndmap);
      memcpy(hd.data(), pb.dmap, hd.size() * 4);
      plan = tiled_plan_create(F, G, hd.data(), stream, &perr);
      if (!plan) fail(2, "tiled sweep setup failed: " + perr);
      if (ws_supported(ndim, G)) {
        ws = ws_plan_create(F, G, hd.data(), stream, &perr);
        if (!ws) fail(2, "warp-specialised sweep setup failed: " + perr);
      }
    }

    if (!aniso && sweeps2d_supported(ndim, G)) {
      std::string perr;
      std::vector<float> hd((size_t)18 * pb.ndmap);
      memcpy(hd.data(), pb.dmap, hd.size() * 4);
      p2d = plan2d_create(F, G, hd.data(), stream, &perr);
      if (!p2d) fail(2, "2D sweep setup failed: " + perr);
    }

    // ---- coordinate lists -> linear indices (bit-exact integer maps)
    const int nd = ndim;
    {  // air voxels (ghost planes included)
      std::vector<long long> idx;
      if (pb.ncoordszero > 0 && !pb.icczero) fail(1, "icczero pointer is NULL");
      for (int i = 0; i < (aniso_protocol ? 0 : pb.ncoordszero); ++i) {   // (the anisotropic binaries have no air kernel)
        const int32_t *c = pb.icczero + (size_t)i * nd;
        if (!coord_ok(c)) fail(1, "icczero: air coordinate outside the grid");
        if (c[0] < gx0 || c[0] >= gx0 + nXl) continue;
        idx.push_back(lin(c[0], c[1], nd == 3 ? c[2] : 0));
      }
      n_air = (int)idx.size();
      if (ndim == 2) h_air_idx = idx;
      d_air_idx = dalloc<long long>(n_air);
      if (n_air) FW_CUDA(cudaMemcpyAsync(d_air_idx, idx.data(), n_air * sizeof(long long), cudaMemcpyHostToDevice, stream));
      FW_CUDA(cudaStreamSynchronize(stream));
      air_set.insert(idx.begin(), idx.end());
    }
    setup_sources(pb.ncoords, pb.icc, pb.icmat);
    int32_t found_box[6];
    const int32_t *obox = pb.out_box;
    if (!obox && pb.outc && pb.ncoordsout >= 4096 && detect_box(pb.outc, pb.ncoordsout, nd, found_box)) obox = found_box;
    if (obox) {  // box sensors: the owned planes of the box are a contiguous run of global outc rows
      const int dims[3] = {nX_global, nY, nZ};
      long long vol = 1, per_plane = 1;
      for (int k = 0; k < nd; ++k) {
        if (obox[k] < 0 || obox[nd + k] > dims[k] || obox[k] > obox[nd + k]) fail(1, "out_box: box outside the grid");
        vol *= obox[nd + k] - obox[k];
        if (k > 0) per_plane *= obox[nd + k] - obox[k];
      }
      if (vol != pb.ncoordsout) fail(1, "out_box: ncoordsout is not the box volume");
      const int x0 = std::max(obox[0], own_lo), x1 = std::min(obox[nd], own_hi);
      sens_box = true;
      box.wa = vol > 0 ? std::max(x1 - x0, 0) : 0;
      box.a0 = x0 - gx0;
      box.b0 = nd == 3 ? obox[1] : 0;            box.wb = nd == 3 ? obox[4] - obox[1] : 1;
      box.c0 = nd == 3 ? obox[2] : obox[1];      box.wc = nd == 3 ? obox[5] - obox[2] : obox[3] - obox[1];
      box.a_lo = M - gx0;                        box.a_hi = nX_global - M - gx0;
      box.b_lo = nd == 3 ? M : 0;                box.b_hi = nd == 3 ? nY - M : 1;
      box.c_lo = M;                              box.c_hi = G.nC - M;
      box.sA = G.sA; box.sB = G.sB;
      if ((long long)box.wa * per_plane > INT32_MAX || box.wb > 65535 || box.wa > 65535) fail(1, "out_box: box too large");
      n_sens = (int)((long long)box.wa * per_plane);
      sens_first = (int)((long long)(x0 - obox[0]) * per_plane);
      d_sens_idx = nullptr;
    } else {  // sensors owned by this slab, in global outc order
      std::vector<long long> idx;
      if (pb.ncoordsout > 0 && !pb.outc) fail(1, "outc pointer is NULL");
      for (int i = 0; i < pb.ncoordsout; ++i) {
        const int32_t *c = pb.outc + (size_t)i * nd;
        if (!coord_ok(c)) fail(1, "outc: sensor coordinate outside the grid");
        if (c[0] < own_lo || c[0] >= own_hi) continue;
        sens_ids.push_back(i);
        idx.push_back(is_rim(c[0], c[1], nd == 3 ? c[2] : M) ? -1 : lin(c[0], c[1], nd == 3 ? c[2] : 0));
      }
      n_sens = (int)idx.size();
      if (ndim == 2) h_sens_idx = idx;
      d_sens_idx = dalloc<long long>(n_sens);
      if (n_sens) FW_CUDA(cudaMemcpyAsync(d_sens_idx, idx.data(), n_sens * sizeof(long long), cudaMemcpyHostToDevice, stream));
      FW_CUDA(cudaStreamSynchronize(stream));
    }
    n_sens_global = pb.ncoordsout;
    n_frames = nT > 0 ? (nT + modT - 1) / modT : 0;
    {
      size_t free_b = 0, total_b = 0;
      FW_CUDA(cudaMemGetInfo(&free_b, &total_b));
      const size_t budget = std::min<size_t>(free_b / 4, (size_t)16 << 30);
      const size_t per = std::max<size_t>((size_t)n_sens * 4, 4);
      frames_cap = (int)std::max<size_t>(1, std::min<size_t>((size_t)std::max(n_frames, 1), budget / per));
      if (const char *ev = getenv("FW25_FRAMES_CAP")) frames_cap = std::max(1, std::min(frames_cap, atoi(ev)));   // tests: small ring
      d_frames = dalloc<float>((size_t)frames_cap * std::max(n_sens, 1));
      // rim sensors read 0: the fused 2D step never writes their columns, so the ring starts out zeroed
      FW_CUDA(cudaMemsetAsync(d_frames, 0, (size_t)frames_cap * std::max(n_sens, 1) * sizeof(float), stream));
    }
    d_t = dalloc<int>(1);
    build_fuse_lists();
    trace_point("state, plans, lists");
    // The two staging buffers stay until the engine is destroyed: cudaFree synchronises the device and was measured
    // at ~145 ms per buffer next to 40 GB of live allocations -- more than uploading 5 GB of maps.
    if (const char *g = getenv("FW25_GRAPH")) graph_mode = atoi(g) != 0;
    if (const char *v = getenv("FW25_VARIANT")) {   // tuning / cross-checks: force a sweep implementation
      const int want = atoi(v);
      if (want == 1 || (want == 2 && (plan || p2d)) || (want == 3 && ws)) variant = want;
    }
    FW_CUDA(cudaStreamSynchronize(stream));
  }

  // Per-tile lists of the special cells of the fused 2D step (fw25_internal.h, Fuse2D).  Whole-grid 2D engines
  // only; a source inside the never-updated rim keeps the separate injection kernel.  Opt-in (FW25_FUSE2D=1):
  // measured on a B200 the fused step launches 2.1 kernels per step instead of 3.2-3.6 but is no faster (14.9 vs
  // 14.9 us at 628 x 628, 98.2 vs 96.4 us at 1457 x 2178) -- with programmatic dependent launch the two point kernels
  // already hide behind the sweeps (profiles/README.md).
  void build_fuse_lists() {
    for (void *p : fuse_owned) cudaFree(p);
    fuse_owned.clear();
    fuse_ok = false;
    const char *ev = getenv("FW25_FUSE2D");
    if (!ev || atoi(ev) == 0) return;
    if (ndim != 2 || !p2d || own_lo != 0 || own_hi != nX_global || n_src_rim > 0 || !sweeps2d_fusable()) return;
    const int a_lo = G.a_rim_lo, a_hi = G.a_rim_hi;
    if (a_hi <= a_lo) return;
    const int n_bx = (G.nC - M + 127) / 128, n_by = (a_hi - a_lo + FUSE_TR - 1) / FUSE_TR;
    const int n_tiles = n_bx * n_by;
    struct Ent { int tile; unsigned short cell; unsigned char kind; int row; };
    std::vector<Ent> ents;
    auto add = [&](long long li, int kind, int row) {
      const int a = (int)(li / G.sA), c = (int)(li % G.sA);
      if (a < a_lo || a >= a_hi || c < M || c >= G.nC - M) return;       // rim cells are never updated
      ents.push_back({((a - a_lo) / FUSE_TR) * n_bx + c / 128,
                      (unsigned short)(((a - a_lo) % FUSE_TR) * 128 + c % 128), (unsigned char)kind, row});
    };
    for (size_t i = 0; i < h_src_idx.size(); ++i)
      if (!(h_src_flag[i] & 2)) add(h_src_idx[i], FUSE_SOURCE, h_src_row[i]);   // (also an air voxel: zeroing wins)
    for (long long li : h_air_idx) add(li, FUSE_AIR, 0);
    if (!sens_box)
      for (size_t i = 0; i < h_sens_idx.size(); ++i)
        if (h_sens_idx[i] >= 0) add(h_sens_idx[i], FUSE_SENSOR, (int)i);
    std::vector<int> ofs(n_tiles + 1, 0);
    for (const Ent &e : ents) ++ofs[e.tile + 1];
    for (int k = 0; k < n_tiles; ++k) ofs[k + 1] += ofs[k];
    std::vector<int> pos(ofs.begin(), ofs.end() - 1), row(ents.size());
    std::vector<unsigned short> cell(ents.size());
    std::vector<unsigned char> kind(ents.size());
    for (const Ent &e : ents) {                                            // stable: list order within a tile
      const int k = pos[e.tile]++;
      cell[k] = e.cell; kind[k] = e.kind; row[k] = e.row;
    }
    auto up = [&](const void *h, size_t bytes) {
      void *d = nullptr;
      FW_CUDA(cudaMalloc(&d, std::max<size_t>(bytes, 8)));
      fuse_owned.push_back(d);
      if (bytes) FW_CUDA(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, stream));
      return d;
    };
    fuse.tile_ofs = (const int *)up(ofs.data(), ofs.size() * sizeof(int));
    fuse.ent_cell = (const unsigned short *)up(cell.data(), cell.size() * sizeof(unsigned short));
    fuse.ent_kind = (const unsigned char *)up(kind.data(), kind.size());
    fuse.ent_row = (const int *)up(row.data(), row.size() * sizeof(int));
    FW_CUDA(cudaStreamSynchronize(stream));                                // host vectors go out of scope
    fuse.icmat = d_icmat; fuse.nTic = nTic;
    fuse.frames = d_frames; fuse.n_sens = n_sens; fuse.modT = modT; fuse.cap = frames_cap;
    fuse.d_t = d_t;
    fuse.box = box; fuse.use_box = sens_box ? 1 : 0;
    fuse_ok = true;
  }
  bool use_fused_2d() const {
    return fuse_ok && use_2d(G.a_rim_hi - G.a_rim_lo) && !use_ws() && !use_tiled();
  }

  bool use_ws() const { return ws != nullptr && (variant == 0 || variant == 3); }
  bool use_tiled() const { return plan != nullptr && variant != 1; }
  // variant 2 forces the tiled 2D sweeps; auto picks them for launches of >= ~0.8 M cells
  bool use_2d(int rows) const { return p2d != nullptr && (variant == 2 || (variant == 0 && sweeps2d_worthwhile(G, rows))); }

  void clamp(int gx_lo, int gx_hi, int &a_lo, int &a_hi) const {
    a_lo = std::max(gx_lo - gx0, G.a_rim_lo);
    a_hi = std::min(gx_hi - gx0, G.a_rim_hi);
  }

  void inject(int tt, cudaStream_t st) {
    launch_inject(F.p, d_src_idx, d_src_row, d_src_rim, (tt < nTic || n_src_rim > 0) ? n_src : 0, d_icmat, nTic,
                  tt, d_air_idx, n_air, st);
    launches += launches_per_inject(n_src, n_air, tt, nTic, n_src_rim);
  }
  // push (ws sweeps only): fused halo exchange, see HaloPush
  void sweep_u(int gx_lo, int gx_hi, cudaStream_t st, const HaloPush *push = nullptr) {
    int a_lo, a_hi;
    clamp(gx_lo, gx_hi, a_lo, a_hi);
    if (a_hi <= a_lo) return;
    if (use_ws()) { launches += launch_sweep_u_ws(ws, F, G, a_lo, a_hi, st, push); return; }
    if (push) fail(3, "fused halo push needs the warp-specialised sweeps");
    if (use_tiled()) { launches += launch_sweep_u_tiled(plan, F, G, a_lo, a_hi, st); return; }
    if (use_2d(a_hi - a_lo)) { launches += launch_sweep_u_2d(p2d, F, G, a_lo, a_hi, st); return; }
    launch_sweep_u_simple(ndim, F, G, a_lo, a_hi, st, aniso);
    launches += (a_hi - a_lo + 32767) / 32768;
  }
  void sweep_p(int gx_lo, int gx_hi, cudaStream_t st, const HaloPush *push = nullptr) {
    int a_lo, a_hi;
    clamp(gx_lo, gx_hi, a_lo, a_hi);
    if (a_hi <= a_lo) return;
    if (use_ws()) { launches += launch_sweep_p_ws(ws, F, G, a_lo, a_hi, st, push); return; }
    if (push) fail(3, "fused halo push needs the warp-specialised sweeps");
    if (use_tiled()) { launches += launch_sweep_p_tiled(plan, F, G, a_lo, a_hi, st); return; }
    if (use_2d(a_hi - a_lo)) { launches += launch_sweep_p_2d(p2d, F, G, a_lo, a_hi, st); return; }
    launch_sweep_p_simple(ndim, F, G, a_lo, a_hi, st, aniso);
    launches += (a_hi - a_lo + 32767) / 32768;
  }
  void record(int frame, cudaStream_t st) {
    if (n_sens == 0) return;
    float *slot = d_frames + (size_t)(frame % frames_cap) * n_sens;
    if (sens_box) launch_record_box(F.p, slot, n_sens, nullptr, 0, 1, 1, box, st);
    else launch_record(F.p, d_sens_idx, n_sens, slot, st);
    launches += 1;
  }
  // global outc rows of the local sensors; box sensors keep only the first row until somebody asks
  const std::vector<int32_t> &sensor_ids() {
    if (sens_box && (int)sens_ids.size() != n_sens) {
      sens_ids.resize(n_sens);
      std::iota(sens_ids.begin(), sens_ids.end(), sens_first);
    }
    return sens_ids;
  }
  void step_once() {
    inject(t, stream);
    sweep_u(0, nX_global, stream);
    sweep_p(0, nX_global, stream);
    if (t % modT == 0) record(t / modT, stream);
    ++t;
  }

  bool graph_enabled() const {
    if (graph_mode >= 0) return graph_mode != 0;
    return cells <= ((size_t)1 << 27) && own_lo == 0 && own_hi == nX_global;
  }
  // Capture `steps` whole steps (inject -> fd_u -> fd_p -> record) into one graph.  Graphs that record
  // (modT <= 32) cover whole recording periods and must start at t % modT == 0; for a long period the graph
  // is 16 record-free steps.
  void build_graph(bool with_inject) {
    if (sg.exec) { cudaGraphExecDestroy(sg.exec); sg.exec = nullptr; }
    sg.records = modT <= 32;
    sg.steps = sg.records ? modT * std::max(1, 16 / modT) : 16;
    sg.with_inject = with_inject;
    sg.variant = variant;
    sg.frames = 0;
    const int64_t l0 = launches;
    FW_CUDA(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
    // 2D: fd_p(t) records frame t and applies the injection of step t + 1 itself (two kernels per step); only the
    // first step of the graph is injected by k_inject and the last fd_p does not inject, so the state between graph
    // launches is the same as after launched steps.
    const bool fused = use_fused_2d();
    sg.fused = fused;
    for (int j = 0; j < sg.steps; ++j) {
      if (!fused || j == 0) {
        launch_inject(F.p, d_src_idx, d_src_row, d_src_rim, with_inject ? n_src : 0, d_icmat, nTic, j, d_air_idx,
                      n_air, stream, d_t);
        launches += ((with_inject && n_src > 0) || n_air > 0) ? 1 : 0;
      }
      sweep_u(0, nX_global, stream);
      const bool rec = sg.records && j % modT == 0 && n_sens > 0;
      if (fused) {
        launches += launch_sweep_p_2d_fused(p2d, F, G, G.a_rim_lo, G.a_rim_hi, stream, fuse, j,
                                            (rec ? FUSE_RECORD : 0) | (j + 1 < sg.steps ? FUSE_INJECT : 0));
      } else {
        sweep_p(0, nX_global, stream);
        if (rec) {
          if (sens_box) launch_record_box(F.p, d_frames, n_sens, d_t, j, modT, frames_cap, box, stream);
          else launch_record_dev(F.p, d_sens_idx, n_sens, d_frames, d_t, j, modT, frames_cap, stream);
          ++launches;
        }
      }
      if (sg.records && j % modT == 0) ++sg.frames;
    }
    launch_tick(d_t, -1, sg.steps, stream);
    ++launches;
    cudaGraph_t g = nullptr;
    FW_CUDA(cudaStreamEndCapture(stream, &g));
    sg.nodes = (int)(launches - l0);
    launches = l0;
    cudaError_t e = cudaGraphInstantiate(&sg.exec, g, 0);
    cudaGraphDestroy(g);
    FW_CUDA(e);
  }
  // Advance by one step, or by a whole graph of steps when one fits: returns the number of steps taken.
  // frame_room = frames the ring can still take before the caller must read them out.
  int advance(int max_steps, int frame_room) {
    if (max_steps <= 0) return 0;
    if (graph_enabled()) {
      const bool wi = n_src > 0 && (n_src_rim > 0 || t < nTic);
      const bool records = modT <= 32;
      const int steps = records ? modT * std::max(1, 16 / modT) : 16;
      const bool phase_ok = records ? (t % modT == 0) : (t % modT != 0 && (t % modT) + steps <= modT);
      const int frames = records ? steps / modT : 0;
      if (phase_ok && steps <= max_steps && frames <= frame_room) {
        if (!sg.exec || sg.with_inject != wi || sg.variant != variant || sg.fused != use_fused_2d()) build_graph(wi);
        if (d_t_host != t) { launch_tick(d_t, t, 0, stream); ++launches; }
        FW_CUDA(cudaGraphLaunch(sg.exec, stream));
        launches += sg.nodes;
        t += sg.steps;
        d_t_host = t;
        return sg.steps;
      }
    }
    step_once();
    return 1;
  }
  void read_frames(int f0, int f1, float *out) {
    if (f0 < 0 || f1 < f0 || f1 - f0 > frames_cap) fail(1, "read_frames: bad frame range");
    if (n_sens == 0 || f1 == f0) return;
    FW_CUDA(cudaStreamSynchronize(stream));
    int f = f0;
    while (f < f1) {  // the ring may wrap
      const int slot = f % frames_cap;
      const int run = std::min(f1 - f, frames_cap - slot);
      FW_CUDA(cudaMemcpy(out + (size_t)(f - f0) * n_sens, d_frames + (size_t)slot * n_sens,
                         (size_t)run * n_sens * 4, cudaMemcpyDeviceToHost));
      f += run;
    }
  }
  float *field(const char *name) const {
    if (!strcmp(name, "p")) return F.p;
    if (!strcmp(name, "u")) return F.q[0];
    if (!strcmp(name, "v")) return ndim == 3 ? F.q[1] : F.q[2];
    if (!strcmp(name, "w")) return ndim == 3 ? F.q[2] : nullptr;
    return nullptr;
  }
};

}  // namespace fw25

using fw25::Engine;
using fw25::Fail;
using fw25::g_err;

struct fw25_engine {
  Engine e;
};

#define FW_TRY(body)                          \
  try {                                       \
    body;                                     \
    return 0;                                 \
  } catch (const Fail &f) {                   \
    return f.code;                            \
  } catch (const std::exception &ex) {        \
    g_err = std::string("exception: ") + ex.what(); \
    return 3;                                 \
  }

// ---------------------------------------------------------------------------------------------- whole job
namespace {

constexpr int M = fw25::M;
using fw25::HaloPush;

int n_frames_of(const fw25_problem *pb) {
  return pb->nT > 0 ? (pb->nT + pb->modT - 1) / std::max(pb->modT, 1) : 0;
}

// frames [f0, f1) of engine e -> columns sens_ids of genout [n_frames][ncoordsout]
void scatter_frames(Engine &e, int f0, int f1, float *genout, int ncoordsout, std::vector<float> &tmp) {
  if (e.n_sens == 0 || f1 <= f0) return;
  if (e.n_sens == ncoordsout) {            // one slab owns every sensor: rows are already in global order
    e.read_frames(f0, f1, genout + (size_t)f0 * ncoordsout);
    return;
  }
  tmp.resize((size_t)(f1 - f0) * e.n_sens);
  e.read_frames(f0, f1, tmp.data());
  for (int f = f0; f < f1; ++f) {
    const float *src = tmp.data() + (size_t)(f - f0) * e.n_sens;
    float *dst = genout + (size_t)f * ncoordsout;
    if (e.sens_box) {                        // a slab's share of a box is one contiguous run of rows
      memcpy(dst + e.sens_first, src, (size_t)e.n_sens * sizeof(float));
      continue;
    }
    for (int i = 0; i < e.n_sens; ++i) dst[e.sens_ids[i]] = src[i];
  }
}

// One slab per device, all driven from this host thread in lockstep -- the reference's own model (a single
// process looping over cudaSetDevice) and what `cuda_device_id=[0, 1, ...]` /
// CUDA_VISIBLE_DEVICES="0,1,..." select.  Same partition rule as the reference, boundary-first schedule, and
// only the planes the stencils read cross an interface: u (8 planes), v, w (1 plane each) after fd_u, p (8) after
// fd_p -- 18 planes per direction per step against the reference's 16 arrays x 8.  Transfers are peer-to-peer
// copies (NVLink) queued on the sender's boundary stream and overlapped with the interior sweeps.
struct MultiRun {
  struct Dev {
    fw25_engine *h = nullptr;
    int device = 0;
    int own_lo = 0, own_hi = 0, gx0 = 0, gx1 = 0;
    bool has_lo = false, has_hi = false;
    cudaStream_t bnd = nullptr;
    cudaEvent_t ev_main = nullptr, ev_bu = nullptr, ev_bp = nullptr, ev_end = nullptr, ev_sent = nullptr, ev_in = nullptr;
  };
  std::vector<Dev> d;
  int t = 0;
  int64_t halo_bytes = 0;
  // fused: the boundary sweeps store their results straight into the neighbour's ghost planes over NVLink
  // (HaloPush) -- no copies.  Needs the warp-specialised 3D sweeps and peer access on every interface.
  bool fused = false;
  // concurrent (default): boundary sweeps on the high-priority stream next to the interior sweep.  serial
  // (FW25_SLAB_SCHEDULE=serial): boundary planes first on the engine's own stream, then the interior -- one sweep
  // kernel on the GPU at a time; only copies use the boundary stream.  Measured equal on a B200 pair.
  bool serial = false;

  ~MultiRun() {
    for (auto &x : d) {
      cudaSetDevice(x.device);
      if (x.h) cudaStreamSynchronize(x.h->e.stream);
      if (x.bnd) { cudaStreamSynchronize(x.bnd); cudaStreamDestroy(x.bnd); }
      for (cudaEvent_t ev : {x.ev_main, x.ev_bu, x.ev_bp, x.ev_end, x.ev_sent, x.ev_in})
        if (ev) cudaEventDestroy(ev);
      if (x.h) fw25_destroy(x.h);
    }
  }

  void init(const fw25_problem &pb, const int32_t *device_ids, int n) {
    const int nX = pb.nX, base = nX / n, rem = nX % n;
    if (base < 2 * M) fw25::fail(1, "x-slabs would be thinner than two halos (16 planes): use fewer GPUs");
    if (pb.ext_p || pb.ext_u || pb.ext_v || pb.ext_w) fw25::fail(1, "caller-owned state arrays need a single device");
    const size_t row = pb.ndim == 3 ? (size_t)(pb.map_pitch > 0 ? pb.map_pitch : pb.nZ) * pb.nY
                                    : (size_t)(pb.map_pitch > 0 ? pb.map_pitch : pb.nY);   // map elements per x plane
    d.resize(n);
    int lo = 0;
    for (int r = 0; r < n; ++r) {
      Dev &x = d[r];
      x.device = device_ids[r];
      x.own_lo = lo;
      x.own_hi = lo + base + (r < rem ? 1 : 0);
      x.has_lo = r > 0;
      x.has_hi = r < n - 1;
      x.gx0 = x.has_lo ? x.own_lo - M : 0;
      x.gx1 = x.has_hi ? x.own_hi + M : nX;
      lo = x.own_hi;
      fw25_problem sub = pb;
      sub.nX = x.gx1 - x.gx0;
      const size_t off = (size_t)x.gx0 * row;
      const float **maps[13] = {&sub.rho, &sub.K, &sub.beta, &sub.kappax, &sub.kappau, &sub.apmlx1, &sub.bpmlx1,
                                &sub.apmlx2, &sub.bpmlx2, &sub.apmlu1, &sub.bpmlu1, &sub.apmlu2, &sub.bpmlu2};
      for (auto m : maps)
        if (*m) *m += off;
      if (sub.dcmap) sub.dcmap += off;
      fw25_aniso an_sub;
      if (pb.aniso) {
        an_sub = *pb.aniso;
        for (int ax = 0; ax < 3; ++ax) {
          if (an_sub.kappa_vel[ax]) an_sub.kappa_vel[ax] += off;
          if (an_sub.kappa_prs[ax]) an_sub.kappa_prs[ax] += off;
          for (int nu = 0; nu < 2; ++nu) {
            if (an_sub.a_vel[ax][nu]) an_sub.a_vel[ax][nu] += off;
            if (an_sub.b_vel[ax][nu]) an_sub.b_vel[ax][nu] += off;
            if (an_sub.a_prs[ax][nu]) an_sub.a_prs[ax][nu] += off;
            if (an_sub.b_prs[ax][nu]) an_sub.b_prs[ax][nu] += off;
          }
        }
        sub.aniso = &an_sub;
      }
      fw25_slab sl{nX, x.gx0, x.own_lo, x.own_hi};
      const int rc = fw25_create(&sub, &sl, x.device, &x.h);
      if (rc) throw Fail{rc};
      FW_CUDA(cudaSetDevice(x.device));
      int lo_pri = 0, hi_pri = 0;
      FW_CUDA(cudaDeviceGetStreamPriorityRange(&lo_pri, &hi_pri));
      FW_CUDA(cudaStreamCreateWithPriority(&x.bnd, cudaStreamNonBlocking, hi_pri));
      for (cudaEvent_t *ev : {&x.ev_main, &x.ev_bu, &x.ev_bp, &x.ev_end, &x.ev_sent, &x.ev_in})
        FW_CUDA(cudaEventCreateWithFlags(ev, cudaEventDisableTiming));
    }
    bool all_peer = true;
    for (int r = 0; r + 1 < n; ++r) {      // neighbours talk over NVLink when the platform allows it
      const int a = d[r].device, b = d[r + 1].device;
      int ab = 0, ba = 0;
      if (a == b) continue;                // (tests: two slabs on one device)
      cudaDeviceCanAccessPeer(&ab, a, b);
      cudaDeviceCanAccessPeer(&ba, b, a);
      if (ab) { cudaSetDevice(a); cudaDeviceEnablePeerAccess(b, 0); }
      if (ba) { cudaSetDevice(b); cudaDeviceEnablePeerAccess(a, 0); }
      cudaGetLastError();                  // cudaErrorPeerAccessAlreadyEnabled is fine; copies are staged otherwise
      all_peer = all_peer && ab && ba;
    }
    fused = all_peer;
    for (int r = 0; r < n; ++r) fused = fused && E(r).use_ws();
    if (const char *ev = getenv("FW25_FUSED_HALO")) fused = fused && atoi(ev) != 0;
    if (const char *ev = getenv("FW25_SLAB_SCHEDULE")) serial = std::string(ev) == "serial";
  }

  // what a boundary sweep of slab r next to neighbour `to` pushes: the neighbour's arrays, shifted so that r's
  // element index lands on the same global plane; v, w only for the plane adjacent to the interface
  HaloPush push_to(int r, int to, bool velocities) {
    Engine &me = E(r), &nb = E(to);
    const long long shift = (long long)(me.gx0 - nb.gx0) * me.G.sA;
    HaloPush h{};
    const int g8 = to < r ? d[r].own_lo : d[r].own_hi - M;       // the 8 planes next to the interface
    h.lo0 = g8 - me.gx0;
    h.hi0 = h.lo0 + M;
    if (velocities) {
      for (int k = 0; k < 3; ++k) h.a[k] = nb.F.q[k] + shift;
      const int g1 = to < r ? d[r].own_lo : d[r].own_hi - 1;     // the one plane of v, w the neighbour reads
      h.lo1 = g1 - me.gx0;
      h.hi1 = h.lo1 + 1;
    } else {
      h.a[0] = nb.F.p + shift;
    }
    return h;
  }

  Engine &E(int r) { return d[r].h->e; }

  // my outermost owned planes [lo, lo+w) of `name` -> the same global planes (ghosts) of neighbour `to`
  void send_planes(int r, int to, int which, int g_lo, int w) {
    Engine &src = E(r), &dst = E(to);
    float *s = which < 0 ? src.F.p : src.F.q[which];
    float *t_ = which < 0 ? dst.F.p : dst.F.q[which];
    const size_t plane = (size_t)src.G.sA;
    const size_t bytes = (size_t)w * plane * sizeof(float);
    FW_CUDA(cudaMemcpyPeerAsync(t_ + (size_t)(g_lo - dst.gx0) * plane, d[to].device,
                                s + (size_t)(g_lo - src.gx0) * plane, d[r].device, bytes, d[r].bnd));
    halo_bytes += (int64_t)bytes;
  }

  // exchange after a sweep: ready[r] = event on r's boundary stream after which r's boundary planes are final and
  // r's ghost planes are no longer being read
  void exchange(bool velocities) {
    const int n = (int)d.size();
    if (fused) {                             // the boundary sweeps already pushed: only order the streams
      for (int r = 0; r < n; ++r) {
        Dev &x = d[r];
        cudaEvent_t Dev::*done = velocities ? &Dev::ev_bu : &Dev::ev_bp;
        FW_CUDA(cudaSetDevice(x.device));
        if (x.has_lo) FW_CUDA(cudaStreamWaitEvent(x.bnd, d[r - 1].*done, 0));
        if (x.has_hi) FW_CUDA(cudaStreamWaitEvent(x.bnd, d[r + 1].*done, 0));
      }
      return;
    }
    for (int r = 0; r < n; ++r) {
      Dev &x = d[r];
      FW_CUDA(cudaSetDevice(x.device));
      cudaEvent_t Dev::*ready = velocities ? &Dev::ev_bu : &Dev::ev_bp;
      const int nd = E(r).ndim;
      for (int side = 0; side < 2; ++side) {
        if (side == 0 ? !x.has_lo : !x.has_hi) continue;
        const int to = side == 0 ? r - 1 : r + 1;
        FW_CUDA(cudaStreamWaitEvent(x.bnd, d[to].*ready, 0));
        auto lo_of = [&](int w) { return side == 0 ? x.own_lo : x.own_hi - w; };
        if (velocities) {
          send_planes(r, to, 0, lo_of(M), M);                 // u: x-stencil of fd_p
          if (nd == 3) send_planes(r, to, 1, lo_of(1), 1);    // v, w: cross terms only
          send_planes(r, to, 2, lo_of(1), 1);
        } else {
          send_planes(r, to, -1, lo_of(M), M);                // p: x-stencil of fd_u
        }
      }
      FW_CUDA(cudaEventRecord(x.ev_sent, x.bnd));
    }
    for (int r = 0; r < n; ++r) {
      Dev &x = d[r];
      FW_CUDA(cudaSetDevice(x.device));
      if (x.has_lo) FW_CUDA(cudaStreamWaitEvent(x.bnd, d[r - 1].ev_sent, 0));
      if (x.has_hi) FW_CUDA(cudaStreamWaitEvent(x.bnd, d[r + 1].ev_sent, 0));
    }
  }

  // Planes swept by a boundary launch: only the outer 8 cross the interface, but an 8-plane launch pays the
  // x-marching kernels' chunk prologue for 8 planes of work, a full 32-plane chunk does not.
  int bw(const Dev &x) const {
    const int sides = (x.has_lo ? 1 : 0) + (x.has_hi ? 1 : 0);
    return std::max(M, std::min(32, (x.own_hi - x.own_lo) / std::max(sides, 1)));
  }
  template <class Fn>
  void boundary(int r, Fn &&fn) {            // fn(lo, hi, neighbour)
    Dev &x = d[r];
    const int w = bw(x);
    if (x.has_lo) fn(x.own_lo, std::min(x.own_lo + w, x.own_hi), r - 1);
    if (x.has_hi) fn(std::max(x.own_hi - w, x.own_lo), x.own_hi, r + 1);
  }
  void plane_bytes(int r, int planes) { halo_bytes += (int64_t)planes * E(r).G.sA * (int64_t)sizeof(float); }

  void wait_neighbours(cudaStream_t st, int r, cudaEvent_t Dev::*ev) {
    if (d[r].has_lo) FW_CUDA(cudaStreamWaitEvent(st, d[r - 1].*ev, 0));
    if (d[r].has_hi) FW_CUDA(cudaStreamWaitEvent(st, d[r + 1].*ev, 0));
  }
  // copies of one exchange on the boundary streams; `done` is recorded on each boundary stream once the planes of
  // both neighbours have landed.  final_ev: the sender's planes are final AND (the same event of the neighbour) the
  // neighbour's ghost planes are no longer read.
  void copy_exchange(bool velocities, cudaEvent_t Dev::*final_ev, cudaEvent_t Dev::*done) {
    const int n = (int)d.size();
    for (int r = 0; r < n; ++r) {
      Dev &x = d[r];
      FW_CUDA(cudaSetDevice(x.device));
      FW_CUDA(cudaStreamWaitEvent(x.bnd, x.*final_ev, 0));
      const int nd = E(r).ndim;
      for (int side = 0; side < 2; ++side) {
        if (side == 0 ? !x.has_lo : !x.has_hi) continue;
        const int to = side == 0 ? r - 1 : r + 1;
        FW_CUDA(cudaStreamWaitEvent(x.bnd, d[to].*final_ev, 0));
        auto lo_of = [&](int w) { return side == 0 ? x.own_lo : x.own_hi - w; };
        if (velocities) {
          send_planes(r, to, 0, lo_of(M), M);
          if (nd == 3) send_planes(r, to, 1, lo_of(1), 1);
          send_planes(r, to, 2, lo_of(1), 1);
        } else {
          send_planes(r, to, -1, lo_of(M), M);
        }
      }
      FW_CUDA(cudaEventRecord(x.ev_sent, x.bnd));
    }
    for (int r = 0; r < n; ++r) {
      Dev &x = d[r];
      FW_CUDA(cudaSetDevice(x.device));
      wait_neighbours(x.bnd, r, &Dev::ev_sent);
      FW_CUDA(cudaEventRecord(x.*done, x.bnd));
    }
  }

  void step_serial() {
    const int n = (int)d.size();
    for (int r = 0; r < n; ++r) {          // inject, boundary planes of fd_u
      Dev &x = d[r];
      Engine &e = E(r);
      FW_CUDA(cudaSetDevice(x.device));
      if (t > 0) {                         // my ghost p planes are in; the neighbours' ghost velocities were read
        if (fused) wait_neighbours(e.stream, r, &Dev::ev_bp);
        else FW_CUDA(cudaStreamWaitEvent(e.stream, x.ev_end, 0));
      }
      e.inject(t, e.stream);
      boundary(r, [&](int lo, int hi, int to) {
        if (!fused) { e.sweep_u(lo, hi, e.stream); return; }
        const HaloPush h = push_to(r, to, true);
        e.sweep_u(lo, hi, e.stream, &h);
        plane_bytes(r, M + 2);
      });
      FW_CUDA(cudaEventRecord(x.ev_bu, e.stream));
    }
    if (!fused) copy_exchange(true, &Dev::ev_bu, &Dev::ev_in);
    for (int r = 0; r < n; ++r) {          // interior fd_u (the transfers overlap it), boundary planes of fd_p
      Dev &x = d[r];
      Engine &e = E(r);
      FW_CUDA(cudaSetDevice(x.device));
      e.sweep_u(x.own_lo + (x.has_lo ? bw(x) : 0), x.own_hi - (x.has_hi ? bw(x) : 0), e.stream);
      if (fused) wait_neighbours(e.stream, r, &Dev::ev_bu);     // pushed velocities landed; their p ghosts were read
      else FW_CUDA(cudaStreamWaitEvent(e.stream, x.ev_in, 0));
      boundary(r, [&](int lo, int hi, int to) {
        if (!fused) { e.sweep_p(lo, hi, e.stream); return; }
        const HaloPush h = push_to(r, to, false);
        e.sweep_p(lo, hi, e.stream, &h);
        plane_bytes(r, M);
      });
      FW_CUDA(cudaEventRecord(x.ev_bp, e.stream));
    }
    if (!fused) copy_exchange(false, &Dev::ev_bp, &Dev::ev_end);
    for (int r = 0; r < n; ++r) {          // interior fd_p (the p transfers overlap it), sensors
      Dev &x = d[r];
      Engine &e = E(r);
      FW_CUDA(cudaSetDevice(x.device));
      e.sweep_p(x.own_lo + (x.has_lo ? bw(x) : 0), x.own_hi - (x.has_hi ? bw(x) : 0), e.stream);
      if (t % e.modT == 0) e.record(t / e.modT, e.stream);
      e.t = t + 1;
    }
    ++t;
  }

  void step() {
    if (serial) { step_serial(); return; }
    const int n = (int)d.size();
    for (int r = 0; r < n; ++r) {          // inject, boundary planes of fd_u first
      Dev &x = d[r];
      Engine &e = E(r);
      FW_CUDA(cudaSetDevice(x.device));
      if (t > 0) FW_CUDA(cudaStreamWaitEvent(e.stream, x.ev_end, 0));   // ghost p planes of the previous step are in
      e.inject(t, e.stream);
      FW_CUDA(cudaEventRecord(x.ev_main, e.stream));
      FW_CUDA(cudaStreamWaitEvent(x.bnd, x.ev_main, 0));
      if (fused && t > 0) {                  // the neighbours' boundary fd_p of the previous step read their ghosts
        if (x.has_lo) FW_CUDA(cudaStreamWaitEvent(x.bnd, d[r - 1].ev_bp, 0));
        if (x.has_hi) FW_CUDA(cudaStreamWaitEvent(x.bnd, d[r + 1].ev_bp, 0));
      }
      boundary(r, [&](int lo, int hi, int to) {
        if (!fused) { e.sweep_u(lo, hi, x.bnd); return; }
        const HaloPush h = push_to(r, to, true);
        e.sweep_u(lo, hi, x.bnd, &h);
        plane_bytes(r, M + 2);
      });
      FW_CUDA(cudaEventRecord(x.ev_bu, x.bnd));
    }
    exchange(true);
    for (int r = 0; r < n; ++r) {          // interior fd_u overlaps the transfers; then boundary planes of fd_p
      Dev &x = d[r];
      Engine &e = E(r);
      FW_CUDA(cudaSetDevice(x.device));
      e.sweep_u(x.own_lo + (x.has_lo ? bw(x) : 0), x.own_hi - (x.has_hi ? bw(x) : 0), e.stream);
      FW_CUDA(cudaEventRecord(x.ev_main, e.stream));
      FW_CUDA(cudaStreamWaitEvent(x.bnd, x.ev_main, 0));
      boundary(r, [&](int lo, int hi, int to) {   // (fused: exchange(true) made this stream wait for the
        if (!fused) { e.sweep_p(lo, hi, x.bnd); return; }   //  neighbours' boundary fd_u, the last readers of their p ghosts)
        const HaloPush h = push_to(r, to, false);
        e.sweep_p(lo, hi, x.bnd, &h);
        plane_bytes(r, M);
      });
      FW_CUDA(cudaEventRecord(x.ev_bp, x.bnd));
    }
    exchange(false);
    for (int r = 0; r < n; ++r) {          // interior fd_p overlaps the p transfers
      Dev &x = d[r];
      Engine &e = E(r);
      FW_CUDA(cudaSetDevice(x.device));
      FW_CUDA(cudaEventRecord(x.ev_end, x.bnd));
      FW_CUDA(cudaStreamWaitEvent(e.stream, x.ev_bu, 0));     // interior fd_p reads the boundary planes' velocities
      e.sweep_p(x.own_lo + (x.has_lo ? bw(x) : 0), x.own_hi - (x.has_hi ? bw(x) : 0), e.stream);
      if (t % e.modT == 0) {
        FW_CUDA(cudaStreamWaitEvent(e.stream, x.ev_bp, 0));
        e.record(t / e.modT, e.stream);
      }
      e.t = t + 1;
    }
    ++t;
  }

  void sync_all() {
    for (auto &x : d) {
      FW_CUDA(cudaSetDevice(x.device));
      FW_CUDA(cudaStreamSynchronize(x.bnd));
      FW_CUDA(cudaStreamSynchronize(x.h->e.stream));
      FW_CUDA(cudaGetLastError());
    }
  }
};

int run_multi(const fw25_problem *pb, const int32_t *device_ids, int n, float *genout, fw25_stats *stats) {
  using clk = std::chrono::steady_clock;
  auto ms_since = [](clk::time_point a) { return std::chrono::duration<double, std::milli>(clk::now() - a).count(); };
  MultiRun mr;
  const auto t0 = clk::now();
  mr.init(*pb, device_ids, n);
  mr.sync_all();
  const double setup_ms = ms_since(t0);
  const int n_frames = n_frames_of(pb);
  int cap = INT_MAX;
  for (int r = 0; r < n; ++r) cap = std::min(cap, mr.E(r).frames_cap);
  std::vector<float> tmp;
  double d2h_ms = 0;
  int flushed = 0;
  auto flush = [&](int upto) {
    const auto a = clk::now();
    mr.sync_all();
    for (int r = 0; r < n; ++r) {
      FW_CUDA(cudaSetDevice(mr.d[r].device));
      scatter_frames(mr.E(r), flushed, upto, genout, pb->ncoordsout, tmp);
    }
    flushed = upto;
    d2h_ms += ms_since(a);
  };
  const auto t1 = clk::now();
  double flush_in_loop = 0;
  for (int t = 0; t < pb->nT; ++t) {
    const int have = (t + pb->modT - 1) / pb->modT;          // frames recorded by steps 0 .. t-1
    if (t % pb->modT == 0 && have - flushed >= cap) { const double b = d2h_ms; flush(have); flush_in_loop += d2h_ms - b; }
    mr.step();
  }
  mr.sync_all();
  const double loop_ms = ms_since(t1) - flush_in_loop;
  flush(n_frames);
  if (stats) {
    stats->setup_ms = setup_ms;
    stats->loop_ms = loop_ms;
    stats->d2h_ms = d2h_ms;
    stats->kernel_launches = 0;
    stats->h2d_bytes = 0;
    for (int r = 0; r < n; ++r) { stats->kernel_launches += mr.E(r).launches; stats->h2d_bytes += mr.E(r).h2d_bytes; }
    stats->d2h_bytes = (int64_t)n_frames * pb->ncoordsout * 4;
    stats->point_updates = (int64_t)pb->nX * pb->nY * (pb->ndim == 3 ? pb->nZ : 1) * (int64_t)pb->nT;
    stats->halo_bytes = mr.halo_bytes;
    stats->n_devices = n;
  }
  return 0;
}

// Whole-domain recordings return gigabytes of frames into memory the caller has just allocated (numpy.zeros, a fresh
// file mapping): the device-to-host copies would then crawl at page-fault speed (measured 4 GB/s for 1.2 GB).  A few
// helper threads fault the pages in (MADV_POPULATE_WRITE: contents untouched) while the GPU runs the time loop.
struct Prefault {
  std::vector<std::thread> th;
  char *lo = nullptr, *hi = nullptr;
  size_t stripe = (size_t)16 << 20, n_stripes = 0;
  std::atomic<size_t> next{0};
  std::unique_ptr<std::atomic<unsigned char>[]> done;
  size_t mark = 0;                       // stripes [0, mark) are known to be populated (reader side)
  void start(void *ptr, size_t bytes) {
#ifdef MADV_POPULATE_WRITE
    if (!ptr || bytes < ((size_t)64 << 20)) return;
    if (const char *ev = getenv("FW25_PREFAULT")) { if (atoi(ev) == 0) return; }
    const size_t page = (size_t)sysconf(_SC_PAGESIZE);
    lo = reinterpret_cast<char *>(((uintptr_t)ptr + page - 1) / page * page);
    hi = reinterpret_cast<char *>(((uintptr_t)ptr + bytes) / page * page);
    if (hi <= lo) { lo = hi = nullptr; return; }
#ifdef MADV_HUGEPAGE
    (void)madvise(lo, (size_t)(hi - lo), MADV_HUGEPAGE);   // 2 MB pages where the kernel allows: 512x fewer faults
#endif
    n_stripes = ((size_t)(hi - lo) + stripe - 1) / stripe;
    done.reset(new std::atomic<unsigned char>[n_stripes]);
    for (size_t i = 0; i < n_stripes; ++i) done[i].store(0);
    // stripes are handed out in address order, so the front of the buffer -- the first frames -- is ready first
    for (int i = 0; i < 6; ++i)
      th.emplace_back([this] {
        for (;;) {
          const size_t k = next.fetch_add(1);
          if (k >= n_stripes) return;
          char *a = lo + k * stripe, *b = std::min(hi, a + stripe);
          (void)madvise(a, (size_t)(b - a), MADV_POPULATE_WRITE);
          done[k].store(1, std::memory_order_release);
        }
      });
#else
    (void)ptr; (void)bytes;
#endif
  }
  // block until every page below `end` has been populated (no-op when nothing was started)
  void wait_until(const void *end) {
    if (!lo) return;
    const char *e = std::min<const char *>(static_cast<const char *>(end), hi);
    if (e <= lo) return;
    const size_t need = ((size_t)(e - lo) + stripe - 1) / stripe;
    while (mark < std::min(need, n_stripes)) {
      if (done[mark].load(std::memory_order_acquire)) ++mark;
      else std::this_thread::sleep_for(std::chrono::microseconds(50));
    }
  }
  void join() { for (auto &t : th) if (t.joinable()) t.join(); th.clear(); }
  ~Prefault() { join(); }
};

// Frames leave the device WHILE the time loop runs: whole-domain / whole-user-grid recordings (every shipped example)
// produce gigabytes of frames, and copying them after the loop costs more than the loop itself (468 x 468 sensors every
// 2nd step of 2805: loop 51 ms, copy 73 ms).  A copier thread waits for the event recorded after the steps that
// complete a batch of frames and copies the batch out of the ring on its own stream; the loop only stalls when the
// ring is full.  Single whole-grid engines (rows already in global order).
struct FrameStreamer {
  Engine &e;
  float *genout;
  Prefault &pf;
  struct Job { int f0, f1; cudaEvent_t ev; };
  std::deque<Job> q;
  std::mutex m;
  std::condition_variable cv_job, cv_done;
  std::thread th;
  bool closing = false, failed = false;
  std::string err;
  std::atomic<int> flushed{0};
  double ms = 0;
  cudaStream_t cs = nullptr;
  FrameStreamer(Engine &e_, float *g, Prefault &p) : e(e_), genout(g), pf(p) {
    FW_CUDA(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
    th = std::thread([this] { body(); });
  }
  void body() {
    cudaSetDevice(e.device);
    for (;;) {
      Job j;
      {
        std::unique_lock<std::mutex> lk(m);
        cv_job.wait(lk, [&] { return closing || !q.empty(); });
        if (q.empty()) return;
        j = q.front(); q.pop_front();
      }
      cudaError_t rc = cudaEventSynchronize(j.ev);
      const size_t n = (size_t)e.n_sens;
      pf.wait_until(genout + (size_t)j.f1 * n);
      const auto t0 = std::chrono::steady_clock::now();
      for (int f = j.f0; f < j.f1 && rc == cudaSuccess;) {          // the ring may wrap
        const int slot = f % e.frames_cap, run = std::min(j.f1 - f, e.frames_cap - slot);
        rc = cudaMemcpyAsync(genout + (size_t)f * n, e.d_frames + (size_t)slot * n, (size_t)run * n * 4,
                             cudaMemcpyDeviceToHost, cs);
        f += run;
      }
      if (rc == cudaSuccess) rc = cudaStreamSynchronize(cs);
      ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
      cudaEventDestroy(j.ev);
      {
        std::lock_guard<std::mutex> lk(m);
        if (rc != cudaSuccess) { failed = true; err = cudaGetErrorString(rc); }
        flushed.store(j.f1);
      }
      cv_done.notify_all();
    }
  }
  void push(int f0, int f1) {                                      // frames [f0, f1) are complete once the work queued so far is
    cudaEvent_t ev = nullptr;
    FW_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    FW_CUDA(cudaEventRecord(ev, e.stream));
    { std::lock_guard<std::mutex> lk(m); q.push_back({f0, f1, ev}); }
    cv_job.notify_one();
  }
  void wait_flushed(int upto) {
    std::unique_lock<std::mutex> lk(m);
    cv_done.wait(lk, [&] { return failed || flushed.load() >= upto; });
  }
  void close() {
    { std::lock_guard<std::mutex> lk(m); closing = true; }
    cv_job.notify_all();
    if (th.joinable()) th.join();
    if (cs) { cudaStreamDestroy(cs); cs = nullptr; }
  }
  ~FrameStreamer() { close(); }
};

// The time loop over an existing engine, from its current step to nT: frames are read out of the device ring when
// it fills up and at the end.  genout: [n_frames][n_sens_global].
void run_loop(Engine &e, float *genout, fw25_stats *stats, double setup_ms) {
  struct Ev {
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    ~Ev() { for (auto x : ev) if (x) cudaEventDestroy(x); }
  } H;
  FW_CUDA(cudaSetDevice(e.device));
  for (auto &x : H.ev) FW_CUDA(cudaEventCreate(&x));
  std::vector<float> tmp;
  double d2h_ms = 0, flush_in_loop = 0;
  int flushed = 0;
  const int64_t l0 = e.launches, h0 = e.h2d_bytes;
  const int t_begin = e.t;
  Prefault pf;
  pf.start(genout, (size_t)e.n_frames * e.n_sens_global * sizeof(float));
  auto flush = [&](int upto) {
    pf.join();
    FW_CUDA(cudaEventRecord(H.ev[2], e.stream));
    scatter_frames(e, flushed, upto, genout, e.n_sens_global, tmp);
    FW_CUDA(cudaEventRecord(H.ev[3], e.stream));
    FW_CUDA(cudaEventSynchronize(H.ev[3]));
    float ms = 0;
    FW_CUDA(cudaEventElapsedTime(&ms, H.ev[2], H.ev[3]));
    d2h_ms += ms;
    flushed = upto;
  };
  // large recordings stream out while the loop runs (FrameStreamer); small ones are read at the end
  const size_t frame_b = (size_t)e.n_sens * sizeof(float);
  bool stream_out = e.n_sens > 0 && e.n_sens == e.n_sens_global && (size_t)e.n_frames * frame_b >= ((size_t)64 << 20);
  if (const char *ev = getenv("FW25_STREAM_FRAMES")) {        // 0: never, 2: whenever there are frames (tests)
    const int v = atoi(ev);
    stream_out = v == 0 ? false : v == 2 ? (e.n_sens > 0 && e.n_sens == e.n_sens_global && e.n_frames > 0) : stream_out;
  }
  std::unique_ptr<FrameStreamer> fs;
  if (stream_out) fs.reset(new FrameStreamer(e, genout, pf));
  size_t batch_b = (size_t)32 << 20;
  if (const char *ev = getenv("FW25_STREAM_BATCH_KB")) batch_b = (size_t)std::max(1, atoi(ev)) << 10;
  const int batch = (int)std::max<size_t>(1, batch_b / std::max<size_t>(frame_b, 1));
  int queued = 0;                                            // frames handed to the streamer
  FW_CUDA(cudaEventRecord(H.ev[0], e.stream));
  while (e.t < e.nT) {
    const int have = (e.t + e.modT - 1) / e.modT;            // frames recorded by steps 0 .. t-1
    if (fs) {
      int room = e.frames_cap - (have - fs->flushed.load());
      if (room <= 0 && e.t % e.modT == 0) {                  // ring full: hand over what is complete, wait for space
        if (have > queued) { fs->push(queued, have); queued = have; }
        fs->wait_flushed(have - e.frames_cap / 2);             // until half of the ring is free again
        if (fs->failed) fw25::fail(2, "frame streamer: " + fs->err);
        room = e.frames_cap - (have - fs->flushed.load());
      }
      e.advance(e.nT - e.t, room);
      const int now = (e.t + e.modT - 1) / e.modT;
      if (now - queued >= batch) { fs->push(queued, now); queued = now; }
      continue;
    }
    int room = e.frames_cap - (have - flushed);
    if (room <= 0 && e.t % e.modT == 0) { const double b = d2h_ms; flush(have); flush_in_loop += d2h_ms - b; room = e.frames_cap; }
    e.advance(e.nT - e.t, room);
  }
  FW_CUDA(cudaEventRecord(H.ev[1], e.stream));
  if (fs) {
    if (e.n_frames > queued) fs->push(queued, e.n_frames);
    fs->wait_flushed(e.n_frames);
    if (fs->failed) fw25::fail(2, "frame streamer: " + fs->err);
    fs->close();
    d2h_ms = fs->ms;
    flushed = e.n_frames;
  }
  FW_CUDA(cudaEventSynchronize(H.ev[1]));
  FW_CUDA(cudaGetLastError());
  float loop_ms = 0;
  FW_CUDA(cudaEventElapsedTime(&loop_ms, H.ev[0], H.ev[1]));
  if (!fs) flush(e.n_frames);
  if (stats) {
    stats->setup_ms = setup_ms;
    stats->loop_ms = loop_ms - flush_in_loop;
    stats->d2h_ms = d2h_ms;
    stats->kernel_launches = e.launches - l0;
    stats->h2d_bytes = setup_ms > 0 ? e.h2d_bytes : e.h2d_bytes - h0;
    stats->d2h_bytes = (int64_t)e.n_frames * e.n_sens * 4;
    stats->point_updates = (int64_t)e.nXl * e.nY * e.nZ * (int64_t)(e.nT - t_begin);
    stats->halo_bytes = 0;
    stats->n_devices = 1;
  }
}

int run_single(const fw25_problem *pb, int dev0, float *genout, fw25_stats *stats) {
  struct Holder {
    fw25_engine *h = nullptr;
    ~Holder() { if (h) fw25_destroy(h); }
  } H;
  using clk = std::chrono::steady_clock;
  const auto t0 = clk::now();
  const int rc = fw25_create(pb, nullptr, dev0, &H.h);
  if (rc) throw Fail{rc};
  Engine &e = H.h->e;
  FW_CUDA(cudaStreamSynchronize(e.stream));
  const double setup_ms = std::chrono::duration<double, std::milli>(clk::now() - t0).count();
  run_loop(e, genout, stats, setup_ms);
  if (stats) stats->kernel_launches = e.launches;
  return 0;
}

}  // namespace

extern "C" {

const char *fw25_last_error(void) { return g_err.c_str(); }
int32_t fw25_abi_version(void) { return FW25_ABI_VERSION; }
int32_t fw25_pitch(int32_t n_fast) { return fw25::round_up(n_fast, 32); }

int fw25_create(const fw25_problem *pb, const fw25_slab *slab, int32_t device, fw25_engine **out) {
  if (!pb || !out) { g_err = "fw25_create: NULL argument"; return 1; }
  *out = nullptr;
  std::unique_ptr<fw25_engine> h(new fw25_engine());
  try {
    h->e.init(*pb, slab, device);
  } catch (const Fail &f) {
    return f.code;
  } catch (const std::exception &ex) {
    g_err = std::string("exception: ") + ex.what();
    return 3;
  }
  *out = h.release();
  return 0;
}

void fw25_destroy(fw25_engine *h) { delete h; }

static cudaStream_t pick(fw25_engine *h, void *s) { return s ? static_cast<cudaStream_t>(s) : h->e.stream; }

int fw25_inject(fw25_engine *h, int32_t t, void *s) { FW_TRY((cudaSetDevice(h->e.device), h->e.inject(t, pick(h, s)))) }
int fw25_sweep_u(fw25_engine *h, int32_t lo, int32_t hi, void *s) { FW_TRY((cudaSetDevice(h->e.device), h->e.sweep_u(lo, hi, pick(h, s)))) }
int fw25_sweep_p(fw25_engine *h, int32_t lo, int32_t hi, void *s) { FW_TRY((cudaSetDevice(h->e.device), h->e.sweep_p(lo, hi, pick(h, s)))) }
int fw25_record(fw25_engine *h, int32_t frame, void *s) { FW_TRY((cudaSetDevice(h->e.device), h->e.record(frame, pick(h, s)))) }

int fw25_step(fw25_engine *h, int32_t n) {
  FW_TRY({
    FW_CUDA(cudaSetDevice(h->e.device));
    for (int left = n; left > 0;) left -= h->e.advance(left, INT_MAX);
    FW_CUDA(cudaGetLastError());
  })
}
// n whole steps timed with CUDA events on the engine's stream.  out[0] = total ms; with detail != 0 every
// sweep launch is bracketed too: out[1] = sum over fd_u launches, out[2] = fd_p, out[3] = rest
// (injection, recording, gaps).  Blocks until the steps are done.
int fw25_step_timed(fw25_engine *h, int32_t n, int32_t detail, double *out) {
  FW_TRY({
    Engine &e = h->e;
    FW_CUDA(cudaSetDevice(e.device));
    std::vector<cudaEvent_t> ev((size_t)(detail ? 4 * n : 0) + 2);
    for (auto &x : ev) FW_CUDA(cudaEventCreate(&x));
    FW_CUDA(cudaEventRecord(ev[0], e.stream));
    for (int left = detail ? 0 : n; left > 0;) left -= e.advance(left, INT_MAX);
    for (int i = 0; i < (detail ? n : 0); ++i) {
      cudaEvent_t *q = &ev[2 + 4 * (size_t)i];
      e.inject(e.t, e.stream);
      FW_CUDA(cudaEventRecord(q[0], e.stream));
      e.sweep_u(0, e.nX_global, e.stream);
      FW_CUDA(cudaEventRecord(q[1], e.stream));
      FW_CUDA(cudaEventRecord(q[2], e.stream));
      e.sweep_p(0, e.nX_global, e.stream);
      FW_CUDA(cudaEventRecord(q[3], e.stream));
      if (e.t % e.modT == 0) e.record(e.t / e.modT, e.stream);
      ++e.t;
    }
    FW_CUDA(cudaEventRecord(ev[1], e.stream));
    FW_CUDA(cudaEventSynchronize(ev[1]));
    FW_CUDA(cudaGetLastError());
    float ms = 0;
    FW_CUDA(cudaEventElapsedTime(&ms, ev[0], ev[1]));
    out[0] = ms; out[1] = out[2] = out[3] = 0;
    if (detail) {
      for (int i = 0; i < n; ++i) {
        cudaEvent_t *q = &ev[2 + 4 * (size_t)i];
        FW_CUDA(cudaEventElapsedTime(&ms, q[0], q[1])); out[1] += ms;
        FW_CUDA(cudaEventElapsedTime(&ms, q[2], q[3])); out[2] += ms;
      }
      out[3] = out[0] - out[1] - out[2];
    }
    for (auto &x : ev) cudaEventDestroy(x);
  })
}
int fw25_sync(fw25_engine *h) {
  FW_TRY({
    FW_CUDA(cudaSetDevice(h->e.device));
    FW_CUDA(cudaStreamSynchronize(h->e.stream));
    FW_CUDA(cudaGetLastError());
  })
}

int32_t fw25_n_local_sensors(const fw25_engine *h) { return h->e.n_sens; }
int fw25_local_sensor_ids(const fw25_engine *h, int32_t *ids) {
  const std::vector<int32_t> &v = const_cast<fw25_engine *>(h)->e.sensor_ids();
  std::copy(v.begin(), v.end(), ids);
  return 0;
}
int fw25_read_frames(fw25_engine *h, int32_t f0, int32_t f1, float *out) {
  FW_TRY((cudaSetDevice(h->e.device), h->e.read_frames(f0, f1, out)))
}
int fw25_read_field(fw25_engine *h, const char *name, float *out) {
  FW_TRY({
    Engine &e = h->e;
    FW_CUDA(cudaSetDevice(e.device));
    const float *d = e.field(name);
    if (!d) fw25::fail(1, std::string("unknown field: ") + name);
    FW_CUDA(cudaStreamSynchronize(e.stream));
    FW_CUDA(cudaMemcpy2D(out, (size_t)e.G.nC * 4, d, (size_t)e.G.pitch * 4, (size_t)e.G.nC * 4,
                         (size_t)e.G.nA * e.G.nB, cudaMemcpyDeviceToHost));
  })
}
void *fw25_field_ptr(fw25_engine *h, const char *name) { return h->e.field(name); }
int32_t fw25_current_step(const fw25_engine *h) { return h->e.t; }
int64_t fw25_launch_count(const fw25_engine *h) { return h->e.launches; }
int fw25_set_kernel_variant(fw25_engine *h, int32_t v) {
  if (v < 0 || v > 3) { g_err = "fw25_set_kernel_variant: variant must be 0..3"; return 1; }
  if (v == 2 && !h->e.plan && !h->e.p2d) { g_err = "fw25_set_kernel_variant: the TMA-tiled sweeps cannot run this problem"; return 1; }
  if (v == 3 && !h->e.ws) { g_err = "fw25_set_kernel_variant: the warp-specialised sweeps need a 3D problem with < 2^32 cells per array"; return 1; }
  h->e.variant = v;
  return 0;
}

int fw25_run(const fw25_problem *pb, const int32_t *device_ids, int32_t n_devices, float *genout,
             size_t genout_len, fw25_stats *stats) {
  if (!pb) { g_err = "fw25_run: NULL problem"; return 1; }
  if (pb->modT <= 0) { g_err = "modT must be >= 1"; return 1; }
  const int32_t dev0 = 0;
  if (!device_ids || n_devices <= 0) { device_ids = &dev0; n_devices = 1; }
  const int n_frames = n_frames_of(pb);
  if (genout_len < (size_t)n_frames * (size_t)std::max(pb->ncoordsout, 0)) {
    g_err = "fw25_run: genout buffer too small";
    return 1;
  }
  if ((size_t)n_frames * std::max(pb->ncoordsout, 0) > 0 && !genout) { g_err = "fw25_run: NULL genout"; return 1; }
  try {
    if (n_devices > 1) return run_multi(pb, device_ids, n_devices, genout, stats);
    return run_single(pb, device_ids[0], genout, stats);
  } catch (const Fail &f) {
    return f.code;
  } catch (const std::exception &ex) {
    g_err = std::string("exception: ") + ex.what();
    return 3;
  }
}

int fw25_reset(fw25_engine *h, int32_t nT, int32_t nTic, int32_t ncoords, const int32_t *icc, const float *icmat) {
  if (!h) { g_err = "fw25_reset: NULL engine"; return 1; }
  FW_TRY((cudaSetDevice(h->e.device), h->e.reset(nT, nTic, ncoords, icc, icmat)))
}

int fw25_run_engine(fw25_engine *h, float *genout, size_t genout_len, fw25_stats *stats) {
  if (!h) { g_err = "fw25_run_engine: NULL engine"; return 1; }
  Engine &e = h->e;
  if (e.own_lo != 0 || e.own_hi != e.nX_global) { g_err = "fw25_run_engine: the engine holds one slab of a sharded grid"; return 1; }
  if (genout_len < (size_t)e.n_frames * (size_t)e.n_sens_global) { g_err = "fw25_run_engine: genout buffer too small"; return 1; }
  if ((size_t)e.n_frames * e.n_sens_global > 0 && !genout) { g_err = "fw25_run_engine: NULL genout"; return 1; }
  FW_TRY(run_loop(e, genout, stats, 0.0))
}

int32_t fw25_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

}  // extern "C"
