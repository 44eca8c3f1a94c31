// fw25_run.cu -- whole-job loops over one engine: frame ring read-out, frames streamed to the host while the loop
// runs, pre-faulted destination memory (the genout kernels + writer thread of the reference).
#include <sys/mman.h>
#include <unistd.h>

#include <atomic>
#include <condition_variable>
#include <deque>
#include <memory>
#include <mutex>
#include <thread>

#include "fw25_engine.h"

namespace fw25 {

namespace {

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

}  // namespace

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

// ---- deferred teardown.  Giving 150 GB of device memory back takes the driver 50-300 ms (unmapping the page tables;
// profiles/README.md), as long as six time steps of the largest grid.  The whole-job entry points therefore return as
// soon as the frames are on the host and hand the engine / map set to one reaper thread; every entry point that
// allocates device memory joins it first (reap_wait), so the memory is back before anybody can miss it, and the
// library joins it at unload.  FW25_ASYNC_TEARDOWN=0 frees before returning.
namespace {
struct Reaper {
  std::mutex m;
  std::thread th;
  void wait() {
    std::lock_guard<std::mutex> lk(m);
    if (th.joinable()) th.join();
  }
  void run(std::function<void()> fn) {
    static const bool on = [] { const char *e = getenv("FW25_ASYNC_TEARDOWN"); return !e || atoi(e) != 0; }();
    if (!on) { fn(); return; }
    std::lock_guard<std::mutex> lk(m);
    if (th.joinable()) th.join();
    th = std::thread(std::move(fn));
  }
  ~Reaper() { if (th.joinable()) th.join(); }
} g_reaper;
}  // namespace

void reap_wait() { g_reaper.wait(); }
void reap_async(std::function<void()> fn) { g_reaper.run(std::move(fn)); }

int run_single(const fw25_problem *pb, int dev0, float *genout, fw25_stats *stats) {
  struct Holder {
    fw25_engine *h = nullptr;
    ~Holder() {
      if (!h) return;
      fw25_engine *e = h;
      reap_async([e] { fw25_destroy(e); });
    }
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

}  // namespace fw25
