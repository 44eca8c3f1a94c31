// fw25_multi.cu -- several x-slabs driven from ONE process (fw25_run with a device list).
#include "fw25_engine.h"

namespace fw25 {

namespace {

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

  // mapsets (optional): one device-resident map set per slab, planes [gx0, gx1) of that slab (fw25_mapgen_slab) --
  // adopted in place instead of slicing and uploading pb's host maps
  void init(const fw25_problem &pb, const int32_t *device_ids, int n, fw25_mapset *const *mapsets = nullptr) {
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
      const size_t off = mapsets ? 0 : (size_t)x.gx0 * row;
      if (mapsets) {
        if (fw25_mapset_problem(mapsets[r], &sub) != 0) throw Fail{1};
        if (sub.nX != x.gx1 - x.gx0) fw25::fail(1, "a per-slab map set does not hold the slab's planes");
      }
      const float **maps[13] = {&sub.rho, &sub.K, &sub.beta, &sub.kappax, &sub.kappau, &sub.apmlx1, &sub.bpmlx1,
                                &sub.apmlx2, &sub.bpmlx2, &sub.apmlu1, &sub.bpmlu1, &sub.apmlu2, &sub.bpmlu2};
      for (auto m : maps)
        if (*m) *m += off;
      if (sub.dcmap) sub.dcmap += off;
      fw25_aniso an_sub;
      if (pb.aniso && !mapsets) {
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

}  // namespace

int run_multi(const fw25_problem *pb, const int32_t *device_ids, int n, float *genout, fw25_stats *stats,
              fw25_mapset *const *mapsets) {
  using clk = std::chrono::steady_clock;
  auto ms_since = [](clk::time_point a) { return std::chrono::duration<double, std::milli>(clk::now() - a).count(); };
  MultiRun mr;
  const auto t0 = clk::now();
  mr.init(*pb, device_ids, n, mapsets);
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

// fw25_run_medium on several devices: every device builds its own x-slab of the maps (fw25_mapgen_slab, all devices at
// once, one host thread each) from the user-grid medium, the slabs are adopted in place and the native multi-device
// runner steps them.  The partition is MultiRun::init's (the reference's rule).
int run_medium_multi(const fw25_medium *md, const fw25_problem *pb_in, const int32_t *device_ids, int n, float *genout,
                     fw25_stats *stats) {
  using clk = std::chrono::steady_clock;
  const auto t0 = clk::now();
  const int nb = md->m_spatial_order + md->n_pml_layer + md->n_transition_layer;
  const int nX = md->nx + 2 * nb, base = nX / n, rem = nX % n;
  if (base < 2 * M) fw25::fail(1, "x-slabs would be thinner than two halos (16 planes): use fewer GPUs");
  struct Sets {
    std::vector<fw25_mapset *> ms;
    ~Sets() { for (auto *m : ms) if (m) fw25_mapset_destroy(m); }
  } S;
  S.ms.assign(n, nullptr);
  std::vector<std::string> errs(n);
  std::vector<int> rcs(n, 0);
  std::vector<double> gen_ms((size_t)2 * n, 0.0);
  {
    std::vector<std::thread> th;
    int lo = 0;
    for (int r = 0; r < n; ++r) {
      const int own_hi = lo + base + (r < rem ? 1 : 0);
      const int gx0 = r > 0 ? lo - M : 0, gx1 = r < n - 1 ? own_hi + M : nX;
      lo = own_hi;
      th.emplace_back([&, r, gx0, gx1] {
        rcs[r] = fw25_mapgen_slab(md, device_ids[r], gx0, gx1, 0, md->nx, &S.ms[r], &gen_ms[2 * r]);
        if (rcs[r]) errs[r] = fw25_last_error();          // g_err is thread-local
      });
    }
    for (auto &t : th) t.join();
  }
  for (int r = 0; r < n; ++r)
    if (rcs[r]) { g_err = errs[r]; throw Fail{rcs[r]}; }
  const double mapgen_wall_ms = std::chrono::duration<double, std::milli>(clk::now() - t0).count();
  fw25_problem pb = *pb_in;
  pb.ndim = md->ndim;
  pb.nX = nX; pb.nY = md->ny + 2 * nb; pb.nZ = md->ndim == 3 ? md->nz + 2 * nb : 1;
  pb.aniso = nullptr;
  const int rc = run_multi(&pb, device_ids, n, genout, stats, S.ms.data());
  if (stats) stats->setup_ms += mapgen_wall_ms;
  return rc;
}

}  // namespace fw25
