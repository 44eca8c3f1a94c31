// fw25_engine.cu -- engine object + C-ABI (include/fw25.h) of the B200-native Fullwave 2.5 engine.
//
// Replaces what the reference's binary-only `main` does: load maps,
// allocate state, run the time loop inject -> fd_u -> fd_p -> record, return the sensor frames.
// Differences by design: one time level (in-place leapfrog, no proceed_time copies), 64-bit indexing,
// row-padded layout for 16-byte vector / TMA access, coordinate lists resolved to linear indices once.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
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
  int n_sens = 0;
  std::vector<int32_t> sens_ids;
  float *d_frames = nullptr;
  int frames_cap = 0, n_frames = 0;

  int t = 0;
  int64_t launches = 0;
  int variant = 0;
  int64_t h2d_bytes = 0;
  TiledPlan *plan = nullptr;   // TMA-tiled sweeps (3D)
  WsPlan *ws = nullptr;        // warp-specialised all-TMA sweeps (3D)

  ~Engine() {
    cudaSetDevice(device);
    tiled_plan_destroy(plan);
    ws_plan_destroy(ws);
    for (void *p : owned) cudaFree(p);
    if (stream) cudaStreamDestroy(stream);
  }

  template <class T>
  T *dalloc(size_t n) {
    void *p = nullptr;
    FW_CUDA(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
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
    FW_CUDA(cudaMemcpy2DAsync(dst, (size_t)G.pitch * 4, src, (size_t)src_pitch * 4, (size_t)G.nC * 4, rows,
                              on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, stream));
    if (!on_device) h2d_bytes += (int64_t)rows * G.nC * 4;
    return dst;
  }

  void init(const fw25_problem &pb, const fw25_slab *slab, int dev) {
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
    const float *maps[13] = {pb.rho, pb.K, pb.beta, pb.kappax, pb.kappau, pb.apmlx1, pb.bpmlx1,
                             pb.apmlx2, pb.bpmlx2, pb.apmlu1, pb.bpmlu1, pb.apmlu2, pb.bpmlu2};
    for (auto m : maps)
      if (!m) fail(1, "a medium map pointer is NULL");
    if (!pb.dmap || !pb.dcmap) fail(1, "dmap / dcmap pointer is NULL");
    F.rho = upload_map(pb.rho, dev_maps, mp);
    F.K = upload_map(pb.K, dev_maps, mp);
    F.beta = upload_map(pb.beta, dev_maps, mp);
    F.kappax = upload_map(pb.kappax, dev_maps, mp);
    F.kappau = upload_map(pb.kappau, dev_maps, mp);
    F.ax1 = upload_map(pb.apmlx1, dev_maps, mp); F.bx1 = upload_map(pb.bpmlx1, dev_maps, mp);
    F.ax2 = upload_map(pb.apmlx2, dev_maps, mp); F.bx2 = upload_map(pb.bpmlx2, dev_maps, mp);
    F.au1 = upload_map(pb.apmlu1, dev_maps, mp); F.bu1 = upload_map(pb.bpmlu1, dev_maps, mp);
    F.au2 = upload_map(pb.apmlu2, dev_maps, mp); F.bu2 = upload_map(pb.bpmlu2, dev_maps, mp);
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
    // state
    auto state = [&](float *ext) {
      float *d = ext ? ext : dalloc<float>(cells);
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

    if (tiled_supported(ndim, G)) {
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

    // ---- coordinate lists -> linear indices (bit-exact integer maps)
    const int nd = ndim;
    auto coord_ok = [&](const int32_t *c) {
      if (c[0] < 0 || c[0] >= nX_global || c[1] < 0 || c[1] >= nY) return false;
      if (nd == 3 && (c[2] < 0 || c[2] >= nZ)) return false;
      return true;
    };
    {  // sources: every source whose plane is held locally (ghost planes included, so that the
       // neighbour's copy of an injected cell stays consistent without an extra exchange)
      std::vector<long long> idx; std::vector<int> row; std::vector<unsigned char> rim;
      if (pb.ncoords > 0 && (!pb.icc || (!pb.icmat && nTic > 0))) fail(1, "icc / icmat pointer is NULL");
      for (int i = 0; i < pb.ncoords; ++i) {
        const int32_t *c = pb.icc + (size_t)i * nd;
        if (!coord_ok(c)) fail(1, "icc: source coordinate outside the grid");
        if (c[0] < gx0 || c[0] >= gx0 + nXl) continue;
        idx.push_back(lin(c[0], c[1], nd == 3 ? c[2] : 0));
        row.push_back(i);
        const bool r = is_rim(c[0], c[1], nd == 3 ? c[2] : M);
        rim.push_back(r);
        n_src_rim += r;
      }
      n_src = (int)idx.size();
      d_src_idx = dalloc<long long>(n_src); d_src_row = dalloc<int>(n_src); d_src_rim = dalloc<unsigned char>(n_src);
      if (n_src) {
        FW_CUDA(cudaMemcpyAsync(d_src_idx, idx.data(), n_src * sizeof(long long), cudaMemcpyHostToDevice, stream));
        FW_CUDA(cudaMemcpyAsync(d_src_row, row.data(), n_src * sizeof(int), cudaMemcpyHostToDevice, stream));
        FW_CUDA(cudaMemcpyAsync(d_src_rim, rim.data(), n_src, cudaMemcpyHostToDevice, stream));
        const size_t nic = (size_t)pb.ncoords * nTic;
        d_icmat = dalloc<float>(nic);
        if (nic) FW_CUDA(cudaMemcpyAsync(d_icmat, pb.icmat, nic * 4, cudaMemcpyHostToDevice, stream));
        h2d_bytes += (int64_t)nic * 4;
      }
      FW_CUDA(cudaStreamSynchronize(stream));  // host vectors go out of scope
    }
    {  // air voxels
      std::vector<long long> idx;
      if (pb.ncoordszero > 0 && !pb.icczero) fail(1, "icczero pointer is NULL");
      for (int i = 0; i < pb.ncoordszero; ++i) {
        const int32_t *c = pb.icczero + (size_t)i * nd;
        if (!coord_ok(c)) fail(1, "icczero: air coordinate outside the grid");
        if (c[0] < gx0 || c[0] >= gx0 + nXl) continue;
        idx.push_back(lin(c[0], c[1], nd == 3 ? c[2] : 0));
      }
      n_air = (int)idx.size();
      d_air_idx = dalloc<long long>(n_air);
      if (n_air) FW_CUDA(cudaMemcpyAsync(d_air_idx, idx.data(), n_air * sizeof(long long), cudaMemcpyHostToDevice, stream));
      FW_CUDA(cudaStreamSynchronize(stream));
    }
    {  // sensors owned by this slab, in global outc order
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
      d_sens_idx = dalloc<long long>(n_sens);
      if (n_sens) FW_CUDA(cudaMemcpyAsync(d_sens_idx, idx.data(), n_sens * sizeof(long long), cudaMemcpyHostToDevice, stream));
      FW_CUDA(cudaStreamSynchronize(stream));
    }
    n_frames = nT > 0 ? (nT + modT - 1) / modT : 0;
    {
      size_t free_b = 0, total_b = 0;
      FW_CUDA(cudaMemGetInfo(&free_b, &total_b));
      const size_t budget = std::min<size_t>(free_b / 4, (size_t)16 << 30);
      const size_t per = std::max<size_t>((size_t)n_sens * 4, 4);
      frames_cap = (int)std::max<size_t>(1, std::min<size_t>((size_t)std::max(n_frames, 1), budget / per));
      d_frames = dalloc<float>((size_t)frames_cap * std::max(n_sens, 1));
    }
    FW_CUDA(cudaStreamSynchronize(stream));
  }

  bool use_ws() const { return ws != nullptr && (variant == 0 || variant == 3); }
  bool use_tiled() const { return plan != nullptr && variant != 1; }

  void clamp(int gx_lo, int gx_hi, int &a_lo, int &a_hi) const {
    a_lo = std::max(gx_lo - gx0, G.a_rim_lo);
    a_hi = std::min(gx_hi - gx0, G.a_rim_hi);
  }

  void inject(int tt, cudaStream_t st) {
    launch_inject(F.p, d_src_idx, d_src_row, d_src_rim, (tt < nTic || n_src_rim > 0) ? n_src : 0, d_icmat, nTic,
                  tt, d_air_idx, n_air, st);
    launches += launches_per_inject(n_src, n_air, tt, nTic, n_src_rim);
  }
  void sweep_u(int gx_lo, int gx_hi, cudaStream_t st) {
    int a_lo, a_hi;
    clamp(gx_lo, gx_hi, a_lo, a_hi);
    if (a_hi <= a_lo) return;
    if (use_ws()) { launches += launch_sweep_u_ws(ws, F, G, a_lo, a_hi, st); return; }
    if (use_tiled()) { launches += launch_sweep_u_tiled(plan, F, G, a_lo, a_hi, st); return; }
    launch_sweep_u_simple(ndim, F, G, a_lo, a_hi, st);
    launches += (a_hi - a_lo + 32767) / 32768;
  }
  void sweep_p(int gx_lo, int gx_hi, cudaStream_t st) {
    int a_lo, a_hi;
    clamp(gx_lo, gx_hi, a_lo, a_hi);
    if (a_hi <= a_lo) return;
    if (use_ws()) { launches += launch_sweep_p_ws(ws, F, G, a_lo, a_hi, st); return; }
    if (use_tiled()) { launches += launch_sweep_p_tiled(plan, F, G, a_lo, a_hi, st); return; }
    launch_sweep_p_simple(ndim, F, G, a_lo, a_hi, st);
    launches += (a_hi - a_lo + 32767) / 32768;
  }
  void record(int frame, cudaStream_t st) {
    if (n_sens == 0) return;
    launch_record(F.p, d_sens_idx, n_sens, d_frames + (size_t)(frame % frames_cap) * n_sens, st);
    launches += 1;
  }
  void step_once() {
    inject(t, stream);
    sweep_u(0, nX_global, stream);
    sweep_p(0, nX_global, stream);
    if (t % modT == 0) record(t / modT, stream);
    ++t;
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
    for (int i = 0; i < n; ++i) h->e.step_once();
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
    for (int i = 0; i < n; ++i) {
      if (!detail) { e.step_once(); continue; }
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
  std::copy(h->e.sens_ids.begin(), h->e.sens_ids.end(), ids);
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
  if (v == 2 && !h->e.plan) { g_err = "fw25_set_kernel_variant: the TMA-tiled sweeps need a 3D problem"; return 1; }
  if (v == 3 && !h->e.ws) { g_err = "fw25_set_kernel_variant: the warp-specialised sweeps need a 3D problem with < 2^32 cells per array"; return 1; }
  h->e.variant = v;
  return 0;
}

int fw25_run(const fw25_problem *pb, const int32_t *device_ids, int32_t n_devices, float *genout,
             size_t genout_len, fw25_stats *stats) {
  if (!pb) { g_err = "fw25_run: NULL problem"; return 1; }
  const int dev0 = (device_ids && n_devices > 0) ? device_ids[0] : 0;
  if (n_devices > 1) { g_err = "fw25_run: in-process multi-device sharding is not built yet; use the torchrun driver"; return 4; }
  const int n_frames = pb->nT > 0 ? (pb->nT + pb->modT - 1) / std::max(pb->modT, 1) : 0;
  if (genout_len < (size_t)n_frames * (size_t)std::max(pb->ncoordsout, 0)) {
    g_err = "fw25_run: genout buffer too small";
    return 1;
  }
  if ((size_t)n_frames * pb->ncoordsout > 0 && !genout) { g_err = "fw25_run: NULL genout"; return 1; }
  fw25_engine *h = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  int rc = 0;
  try {
    FW_CUDA(cudaSetDevice(dev0));
    FW_CUDA(cudaEventCreate(&ev0));
    FW_CUDA(cudaEventCreate(&ev1));
    cudaEvent_t s0, s1;
    FW_CUDA(cudaEventCreate(&s0));
    FW_CUDA(cudaEventCreate(&s1));
    FW_CUDA(cudaEventRecord(s0, 0));
    rc = fw25_create(pb, nullptr, dev0, &h);
    if (rc) throw Fail{rc};
    Engine &e = h->e;
    FW_CUDA(cudaEventRecord(s1, 0));
    FW_CUDA(cudaEventSynchronize(s1));
    float setup_ms = 0, loop_ms = 0;
    FW_CUDA(cudaEventElapsedTime(&setup_ms, s0, s1));
    cudaEventDestroy(s0); cudaEventDestroy(s1);
    double d2h_ms = 0;
    int flushed = 0;  // frames already copied out
    FW_CUDA(cudaEventRecord(ev0, e.stream));
    for (int t = 0; t < e.nT; ++t) {
      e.step_once();
      const int have = (t / e.modT) + 1;  // frames recorded so far
      if (have - flushed == e.frames_cap && t % e.modT == 0 && have < e.n_frames) {
        cudaEvent_t a, b;
        FW_CUDA(cudaEventCreate(&a)); FW_CUDA(cudaEventCreate(&b));
        FW_CUDA(cudaEventRecord(a, e.stream));
        e.read_frames(flushed, have, genout + (size_t)flushed * e.n_sens);
        FW_CUDA(cudaEventRecord(b, e.stream));
        FW_CUDA(cudaEventSynchronize(b));
        float ms = 0; cudaEventElapsedTime(&ms, a, b); d2h_ms += ms;
        cudaEventDestroy(a); cudaEventDestroy(b);
        flushed = have;
      }
    }
    FW_CUDA(cudaEventRecord(ev1, e.stream));
    FW_CUDA(cudaEventSynchronize(ev1));
    FW_CUDA(cudaGetLastError());
    FW_CUDA(cudaEventElapsedTime(&loop_ms, ev0, ev1));
    {
      cudaEvent_t a, b;
      FW_CUDA(cudaEventCreate(&a)); FW_CUDA(cudaEventCreate(&b));
      FW_CUDA(cudaEventRecord(a, e.stream));
      e.read_frames(flushed, e.n_frames, genout + (size_t)flushed * e.n_sens);
      FW_CUDA(cudaEventRecord(b, e.stream));
      FW_CUDA(cudaEventSynchronize(b));
      float ms = 0; cudaEventElapsedTime(&ms, a, b); d2h_ms += ms;
      cudaEventDestroy(a); cudaEventDestroy(b);
    }
    if (stats) {
      stats->setup_ms = setup_ms;
      stats->loop_ms = loop_ms - (e.n_frames > e.frames_cap ? d2h_ms : 0.0);
      stats->d2h_ms = d2h_ms;
      stats->kernel_launches = e.launches;
      stats->h2d_bytes = e.h2d_bytes;
      stats->d2h_bytes = (int64_t)e.n_frames * e.n_sens * 4;
      stats->point_updates = (int64_t)pb->nX * pb->nY * (pb->ndim == 3 ? pb->nZ : 1) * (int64_t)pb->nT;
    }
  } catch (const Fail &f) {
    rc = f.code;
  } catch (const std::exception &ex) {
    g_err = std::string("exception: ") + ex.what();
    rc = 3;
  }
  if (h) fw25_destroy(h);
  if (ev0) cudaEventDestroy(ev0);
  if (ev1) cudaEventDestroy(ev1);
  return rc;
}

}  // extern "C"
