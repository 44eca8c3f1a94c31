// fw25_pipeline.cu -- the first time steps run WHILE the medium is still arriving (fw25_run_medium).
//
// Upstream `Solver.run` is strictly sequential: build the maps on the host (`PMLBuilder.run`), write ~20 files, start
// the binary, which reads and uploads everything and only then steps (solver.py:693-759).  Here the
// user-grid medium goes to the GPU plane block by plane block -- x is the slowest axis of the reference's arrays, so a
// block of planes is one contiguous piece of every map -- `k_mapgen` turns each block into the engine's 14 maps as it
// lands (MapStream, fw25_mapgen.cu), and the x-marching sweeps start on the blocks that are ready.
//
// Time skew.  fd_u(t) at plane x reads p_t at x-7 .. x+8, fd_p(t) reads u_{t+1} at x-8 .. x+7 (and v, w at x +- 1), and
// both update in place.  With blocks of B = 32 planes, U(m) = [B m, B (m+1)) and the half-block-shifted ranges
// P(m) = [B m - 16, B (m+1) - 16), step t can run ONE block behind step t-1:
//
//     stage k:   for t = 0, 1, 2, ...   m = k - t
//                  fd_u(t)      on U(m)      reads p_t on [B m - 7, B (m+1) + 8]: P(m) (final and injected since stage
//                                            k-1, overwritten by fd_p(t) only below) and P(m+1) (fd_p(t-1) and
//                                            inject(t) passed it earlier in THIS stage)
//                  fd_p(t)      on P(m)      reads u_{t+1} on [B m - 24, B m + 23]: U(m-1), U(m); U(m-1) is overwritten
//                                            by fd_u(t+1) only later in this stage
//                  record(t)    on P(m)      the freshly updated p, before the next injection touches it
//                  inject(t+1)  on P(m)      p_{t+1} there is final now
//
// Stage k touches maps of blocks <= k only, so it can be queued as soon as block k has been generated; everything runs
// on ONE stream in this order, which is what makes the in-place update safe.  Every cell sees exactly the operations,
// in the order, of whole-grid sweeps -- the kernels are the same, launched over plane ranges -- so the result is
// bit-identical to the sequential path (tests/test_pipeline.py).
#include <memory>

#include "fw25_engine.h"

namespace fw25 {

bool Engine::skew_supported(int Ts) const {
  // listed sensors whose frames all fit the ring (no read-out in the middle of a skewed phase), whole-grid engine,
  // plane-range sweeps (the warp-specialised 3D kernels or the simple ones), no graph replay
  const int frames = Ts > 0 ? (Ts + modT - 1) / modT : 0;
  return ndim == 3 && !sens_box && frames <= frames_cap && own_lo == 0 && own_hi == nX_global && !graph_enabled() &&
         Ts > 0;
}

void Engine::run_skewed(int Ts, int block, const std::function<cudaEvent_t(int)> &avail) {
  if (block < 4 * M || block % 2) fail(1, "run_skewed: blocks must hold an even number of >= 32 planes");
  if (t != 0) fail(1, "run_skewed: the engine has already stepped");
  const int NB = (G.nA + block - 1) / block, H = block / 2;
  struct Range { int lo, hi; };                                     // local planes
  auto U = [&](int m) { return Range{m * block, std::min((m + 1) * block, G.nA)}; };
  auto P = [&](int m) { return Range{std::min(std::max(m * block - H, 0), G.nA), std::min((m + 1) * block - H, G.nA)}; };
  auto any = [&](const std::vector<unsigned char> &flag, Range r) {
    for (int a = r.lo; a < r.hi; ++a)
      if (flag[a]) return true;
    return false;
  };
  auto inject_range = [&](int tt, Range r) {
    if (r.hi <= r.lo) return;
    const bool src = n_src > 0 && (tt < nTic || n_src_rim > 0) && any(plane_src, r);
    const bool air = n_air > 0 && any(plane_air, r);
    if (!src && !air) return;
    launch_inject_range(F.p, d_src_idx, d_src_row, d_src_rim, src ? n_src : 0, d_icmat, nTic, tt, d_air_idx,
                        air ? n_air : 0, G.sA, r.lo, r.hi, stream);
    ++launches;
  };
  for (int k = 0; k <= NB + Ts - 1; ++k) {
    if (k < NB) {
      cudaEvent_t ev = avail(k);
      if (!ev) throw Fail{2};
      FW_CUDA(cudaStreamWaitEvent(stream, ev, 0));
    }
    for (int tt = 0; tt < Ts; ++tt) {
      const int m = k - tt;
      if (m < 0) break;
      if (m > NB) continue;                                         // step tt is complete
      if (tt == 0) {                                                // nobody injects for step 0 behind a fd_p
        if (m == 0) inject_range(0, P(0));
        if (m + 1 <= NB) inject_range(0, P(m + 1));
      }
      if (m < NB) { const Range r = U(m); sweep_u(gx0 + r.lo, gx0 + r.hi, stream); }
      const Range r = P(m);
      if (r.hi <= r.lo) continue;
      sweep_p(gx0 + r.lo, gx0 + r.hi, stream);
      if (tt % modT == 0 && n_sens > 0 && any(plane_sens, r)) {
        launch_record_range(F.p, d_sens_idx, n_sens, d_frames + (size_t)((tt / modT) % frames_cap) * n_sens, G.sA, r.lo,
                            r.hi, stream);
        ++launches;
      }
      if (tt + 1 < Ts) inject_range(tt + 1, r);
    }
  }
  FW_CUDA(cudaGetLastError());
  t = Ts;
}

// fw25_mapgen + fw25_run in one pipelined call (include/fw25.h, fw25_run_medium)
int run_medium(const fw25_medium *md, const fw25_problem *pb_in, int device, float *genout, fw25_stats *stats) {
  using clk = std::chrono::steady_clock;
  const auto t0 = clk::now();
  auto ms_since = [&](clk::time_point a) { return std::chrono::duration<double, std::milli>(clk::now() - a).count(); };
  constexpr int kBlock = 32;                 // one x-marching chunk of the warp-specialised sweeps
  const bool trace = getenv("FW25_SETUP_TRACE") != nullptr;
  auto tp = [&](const char *what) {
    if (trace) fprintf(stderr, "[fw25 run_medium] %-34s t = %8.1f ms\n", what, ms_since(t0));
  };
  struct Holder {
    MapStream *S = nullptr;
    fw25_engine *h = nullptr;
    fw25_mapset *ms = nullptr;               // set once ownership has passed from the stream to this call
    std::function<void(const char *)> tp;
    // The uploader thread is joined here (the caller's host arrays are free again); the device memory -- engine, map set
    // and the stream's 2 GB upload ring -- goes back on the reaper thread (fw25_run.cu) after this call has returned.
    // Destroying the stream here (a synchronising cudaFree next to 150 GB of live allocations) was the 85-360 ms between
    // "frames on the host" and the return of a 1.1-1.4 s job; now 0.3 ms (profiles/trace_run_medium_r02.txt).
    ~Holder() {
      MapStream *s = S;
      if (s) mapstream_join(s);
      if (tp) tp("uploader joined");
      fw25_engine *e = h;
      fw25_mapset *m = ms;
      if (e || m || s)
        reap_async([e, m, s] {
          if (e) fw25_destroy(e);
          if (m) fw25_mapset_destroy(m);
          if (s) mapstream_destroy(s);       // (still owns the map set if the job failed before mapstream_finish)
        });
      if (tp) tp("teardown handed over");
    }
  } H;
  H.tp = tp;
  fw25_mapset *ms_view = nullptr;
  H.S = mapstream_start(md, device, kBlock, &ms_view);
  if (!H.S) return 2;
  tp("map set allocated");
  fw25_problem pb = *pb_in;
  if (fw25_mapset_problem(ms_view, &pb) != 0) return 1;
  {
    const int rc = fw25_create(&pb, nullptr, device, &H.h);     // adopts the (still empty) maps; uploads lists, builds plans
    if (rc) return rc;
  }
  Engine &e = H.h->e;
  // The medium starts to flow only now: the engine's own uploads (coordinate lists from pageable memory, source
  // signals) would otherwise queue chunk by chunk behind the medium's copies on the one host->device copy engine
  // (measured: 260 ms instead of 20).
  // (Measured again in round 2 with the medium flowing from the start: engine creation 90 -> 310 ms, whole job
  // 1.04-1.12 -> 1.18-1.21 s for 20 steps at 800 x 1240 x 1240.)
  FW_CUDA(cudaStreamSynchronize(e.stream));
  mapstream_go(H.S);
  tp("engine created, uploads started");
  const double setup_ms = ms_since(t0);
  const int NB = mapstream_blocks(H.S);
  // Steps run block-wise under the upload; the rest run as whole-grid sweeps.  Every skewed step adds one block-step
  // of GPU work per arriving block, and block-wise launches pay ~5 % in launch tails: NB / 3 steps keep the GPU busy
  // from the first third of the upload on (a block arrives in ~11 ms over PCIe 5, a block-step takes ~1.8 ms), more
  // only adds small launches (measured at 800 x 1240 x 1240: 8 steps 1.0x s, all 20 steps +50 ms).
  int Ts = std::min(e.nT, std::max(2, NB / 3));
  if (const char *ev = getenv("FW25_SKEW_STEPS")) Ts = std::min(e.nT, std::max(0, atoi(ev)));   // tests / A-B runs
  const int64_t l0 = e.launches;
  cudaEvent_t ev_a = nullptr, ev_b = nullptr;
  FW_CUDA(cudaEventCreate(&ev_a));
  FW_CUDA(cudaEventCreate(&ev_b));
  FW_CUDA(cudaEventRecord(ev_a, e.stream));
  if (e.skew_supported(Ts)) {
    e.run_skewed(Ts, kBlock, [&](int b) { return mapstream_wait_recorded(H.S, b); });
  } else {                                   // whole-grid stepping needs every block
    cudaEvent_t last = mapstream_wait_recorded(H.S, NB - 1);
    if (!last) return 2;
    FW_CUDA(cudaStreamWaitEvent(e.stream, last, 0));
    Ts = 0;
  }
  FW_CUDA(cudaEventRecord(ev_b, e.stream));
  tp("skewed steps queued");
  fw25_stats st{};
  run_loop(e, genout, &st, 0.0);                              // remaining steps + frames to the host
  tp("loop done, frames on the host");
  double gen_ms[2] = {0, 0};
  int64_t h2d = 0;
  H.ms = mapstream_finish(H.S, gen_ms, &h2d);                 // joins the uploader; reports a failed upload
  if (!H.ms) return 2;
  tp("map stream finished");
  float skew_ms = 0;
  FW_CUDA(cudaEventElapsedTime(&skew_ms, ev_a, ev_b));
  cudaEventDestroy(ev_a);
  cudaEventDestroy(ev_b);
  if (stats) {
    *stats = st;
    stats->setup_ms = setup_ms;
    stats->loop_ms = st.loop_ms + skew_ms;
    stats->kernel_launches = e.launches - l0;
    stats->h2d_bytes = e.h2d_bytes + h2d;
    stats->point_updates = (int64_t)e.nXl * e.nY * e.nZ * (int64_t)e.nT;
    stats->n_devices = 1;
    stats->skewed_steps = Ts;                                    // steps run time-skewed under the upload
  }
  return 0;
}

}  // namespace fw25
