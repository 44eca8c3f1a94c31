// fw25_engine.h -- the engine object shared by the engine translation units (fw25_engine.cu: setup and stepping;
// fw25_multi.cu: several slabs in one process; fw25_run.cu: whole-job loops and frame streaming; fw25_cabi.cu: the
// C-ABI of include/fw25.h).
#pragma once
#include <algorithm>
#include <chrono>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <numeric>
#include <string>
#include <thread>
#include <unordered_set>
#include <vector>

#include "../../include/fw25.h"
#include "fw25_internal.h"

namespace fw25 {

extern thread_local std::string g_err;

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

[[noreturn]] inline void fail(int code, const std::string &msg) {
  g_err = msg;
  throw Fail{code};
}

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

// std::vector allocator whose resize() leaves trivially constructible elements uninitialised
template <class T>
struct DefaultInit : std::allocator<T> {
  template <class U> struct rebind { using other = DefaultInit<U>; };
  template <class U, class... A>
  void construct(U *p, A &&...a) {
    if constexpr (sizeof...(A) == 0) ::new (static_cast<void *>(p)) U;
    else ::new (static_cast<void *>(p)) U(std::forward<A>(a)...);
  }
};

bool detect_box(const int32_t *c, int n, int nd, int32_t *box);

// deferred teardown of whole-job runs (fw25_run.cu): one reaper thread; allocating entry points join it first
void reap_wait();
void reap_async(std::function<void()> fn);

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
  // (the source lists are millions of entries that every element of is written right after the resize: no zero-fill)
  std::vector<long long, DefaultInit<long long>> h_src_idx;
  std::vector<long long> h_air_idx, h_sens_idx;
  std::vector<int, DefaultInit<int>> h_src_row;
  std::vector<unsigned char, DefaultInit<unsigned char>> h_src_flag;
  bool fuse_ok = false;
  Fuse2D fuse{};
  std::vector<void *> fuse_owned;
  bool sens_box = false;                     // the sensors are every point of a box: no index list (fw25.h, out_box)
  SensBox box{};
  int sens_first = 0;                        // box sensors: global outc row of local sensor 0 (rows are contiguous)
  std::vector<void *> src_owned;             // source-list allocations (replaced by reset())
  std::unordered_set<long long> air_set;     // linear indices of the air voxels held locally
  std::vector<unsigned char> plane_src, plane_air, plane_sens;   // per local plane: holds a live source / air voxel / sensor
  float *d_frames = nullptr;
  int frames_cap = 0, n_frames = 0;

  int t = 0;
  int64_t launches = 0;
  int variant = 0;
  int64_t h2d_bytes = 0;
  bool aniso = false;            // per-axis relaxation maps that really differ: the ANISO instantiations of the sweeps
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

  // Full-size arrays come out of ONE allocation: on a B200 thirty cudaMalloc calls of 4.9 GB cost 180 ms and 970 ms to
  // free again, one of 148 GB costs 55 ms + 60 ms (profiles/probe_alloc_r02.txt).
  char *arena = nullptr;
  size_t arena_slices = 0, arena_used = 0;
  float *big() {
    if (arena_used < arena_slices) return reinterpret_cast<float *>(arena + (arena_used++) * cells * sizeof(float));
    return dalloc<float>(cells);
  }
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
    float *dst = big();
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
  void upload_dense_rows(float *dst, const float *src, size_t rows);
  void release_staging();

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
  void setup_sources(int ncoords, const int32_t *icc, const float *icmat);

  // Next transmit event on the same medium: zero the wave field, t = 0, new source list.  Maps, stencil
  // tables, tensor maps, sensors and the frame ring stay resident.
  void reset(int nT_, int nTic_, int ncoords, const int32_t *icc, const float *icmat);

  void init(const fw25_problem &pb, const fw25_slab *slab, int dev);

  // Time-skewed start (fw25_pipeline.cu): steps [0, Ts) over blocks of `block` planes as their maps become valid.
  // avail(b) blocks until the event behind which block b's maps are valid has been recorded and returns it.
  bool skew_supported(int Ts) const;
  void run_skewed(int Ts, int block, const std::function<cudaEvent_t(int)> &avail);

  // Per-tile lists of the special cells of the fused 2D step (fw25_internal.h, Fuse2D).  Whole-grid 2D engines
  // only; a source inside the never-updated rim keeps the separate injection kernel.  Opt-in (FW25_FUSE2D=1):
  // measured on a B200 the fused step launches 2.1 kernels per step instead of 3.2-3.6 but is no faster (14.9 vs
  // 14.9 us at 628 x 628, 98.2 vs 96.4 us at 1457 x 2178) -- with programmatic dependent launch the two point kernels
  // already hide behind the sweeps (profiles/README.md).
  void build_fuse_lists();
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
  void build_graph(bool with_inject);
  // Advance by one step, or by a whole graph of steps when one fits: returns the number of steps taken.
  // frame_room = frames the ring can still take before the caller must read them out.
  int advance(int max_steps, int frame_room);
  void read_frames(int f0, int f1, float *out);
  float *field(const char *name) const {
    if (!strcmp(name, "p")) return F.p;
    if (!strcmp(name, "u")) return F.q[0];
    if (!strcmp(name, "v")) return ndim == 3 ? F.q[1] : F.q[2];
    if (!strcmp(name, "w")) return ndim == 3 ? F.q[2] : nullptr;
    return nullptr;
  }
};

}  // namespace fw25

struct fw25_engine {
  fw25::Engine e;
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

namespace fw25 {

int n_frames_of(const fw25_problem *pb);
// frames [f0, f1) of engine e -> columns sens_ids of genout [n_frames][ncoordsout]
void scatter_frames(Engine &e, int f0, int f1, float *genout, int ncoordsout, std::vector<float> &tmp);
// whole jobs (fw25_run): one engine / one slab per device of this process
int run_single(const fw25_problem *pb, int dev0, float *genout, fw25_stats *stats);
// mapsets (optional): per-slab device-resident map sets (fw25_mapgen_slab) adopted instead of pb's host maps
int run_multi(const fw25_problem *pb, const int32_t *device_ids, int n, float *genout, fw25_stats *stats,
              fw25_mapset *const *mapsets = nullptr);
// the time loop over an existing engine, from its current step to nT (fw25_run_engine)
void run_loop(Engine &e, float *genout, fw25_stats *stats, double setup_ms);
// fw25_mapgen + fw25_run pipelined (fw25_pipeline.cu)
int run_medium(const fw25_medium *md, const fw25_problem *pb, int device, float *genout, fw25_stats *stats);
int run_medium_multi(const fw25_medium *md, const fw25_problem *pb, const int32_t *device_ids, int n, float *genout,
                     fw25_stats *stats);

}  // namespace fw25
