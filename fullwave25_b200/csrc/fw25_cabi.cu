// fw25_cabi.cu -- the C-ABI of include/fw25.h over the engine object (the drop-in boundary).
#include <memory>

#include "fw25_engine.h"

using fw25::Engine;
using fw25::Fail;
using fw25::g_err;
using fw25::n_frames_of;
using fw25::run_loop;
using fw25::run_multi;
using fw25::run_single;

extern "C" {

const char *fw25_last_error(void) { return g_err.c_str(); }
int32_t fw25_abi_version(void) { return FW25_ABI_VERSION; }
int32_t fw25_pitch(int32_t n_fast) { return fw25::round_up(n_fast, 32); }

int fw25_create(const fw25_problem *pb, const fw25_slab *slab, int32_t device, fw25_engine **out) {
  if (!pb || !out) { g_err = "fw25_create: NULL argument"; return 1; }
  *out = nullptr;
  fw25::reap_wait();                         // a previous whole-job call may still be giving its memory back
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

static int run_medium_checked(const fw25_medium *md, const fw25_problem *pb, const int32_t *device_ids, int32_t n_devices,
                              float *genout, size_t genout_len, fw25_stats *stats) {
  if (!md || !pb) { g_err = "fw25_run_medium: NULL argument"; return 1; }
  if (pb->modT <= 0) { g_err = "modT must be >= 1"; return 1; }
  if (pb->aniso) { g_err = "fw25_run_medium: the GPU map builder produces the isotropic file set"; return 1; }
  if (n_devices < 1 || !device_ids) { g_err = "fw25_run_medium: at least one device id is needed"; return 1; }
  const int n_frames = n_frames_of(pb);
  if (genout_len < (size_t)n_frames * (size_t)std::max(pb->ncoordsout, 0)) {
    g_err = "fw25_run_medium: genout buffer too small";
    return 1;
  }
  if ((size_t)n_frames * std::max(pb->ncoordsout, 0) > 0 && !genout) { g_err = "fw25_run_medium: NULL genout"; return 1; }
  try {
    if (n_devices > 1) return fw25::run_medium_multi(md, pb, device_ids, n_devices, genout, stats);
    return fw25::run_medium(md, pb, device_ids[0], genout, stats);
  } catch (const Fail &f) {
    return f.code;
  } catch (const std::exception &ex) {
    g_err = std::string("exception: ") + ex.what();
    return 3;
  }
}

int fw25_run_medium(const fw25_medium *md, const fw25_problem *pb, int32_t device, float *genout, size_t genout_len,
                    fw25_stats *stats) {
  return run_medium_checked(md, pb, &device, 1, genout, genout_len, stats);
}

int fw25_run_medium_multi(const fw25_medium *md, const fw25_problem *pb, const int32_t *device_ids, int32_t n_devices,
                          float *genout, size_t genout_len, fw25_stats *stats) {
  return run_medium_checked(md, pb, device_ids, n_devices, genout, genout_len, stats);
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
