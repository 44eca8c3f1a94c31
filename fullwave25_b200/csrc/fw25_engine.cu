// fw25_engine.cu -- engine object of the B200-native Fullwave 2.5 engine (declared in fw25_engine.h).
//
// Replaces what the reference's binary-only `main` does: load maps,
// allocate state, run the time loop inject -> fd_u -> fd_p -> record, return the sensor frames.
// Differences by design: one time level (in-place leapfrog, no proceed_time copies), 64-bit indexing,
// row-padded layout for 16-byte vector / TMA access, coordinate lists resolved to linear indices once.
#include "fw25_engine.h"

namespace fw25 {

thread_local std::string g_err;

// Is the coordinate list exactly the points of a box in row-major order?  (A rectangular `Sensor(mask)`:
// np.where order, sensor.py:24-50.)  box: lo[nd], hi[nd].
bool detect_box(const int32_t *c, int n, int nd, int32_t *box) {
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

void Engine::upload_dense_rows(float *dst, const float *src, size_t rows) {
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

void Engine::release_staging() {
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

void Engine::setup_sources(int ncoords, const int32_t *icc, const float *icmat) {
  const int nd = ndim;
  const bool trace = getenv("FW25_SETUP_TRACE") != nullptr;
  const auto ts0 = std::chrono::steady_clock::now();
  auto tp = [&](const char *what) {
    if (trace) fprintf(stderr, "[fw25 sources] %-24s +%7.1f ms\n", what,
                       std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - ts0).count());
  };
  FW_CUDA(cudaStreamSynchronize(stream));            // the previous lists may still be in flight (reset)
  tp("stream drained");
  for (void *p : src_owned) cudaFree(p);
  src_owned.clear();
  n_src = n_src_rim = 0;
  plane_src.assign((size_t)nXl, 0);
  if (ncoords > 0 && (!icc || (!icmat && nTic > 0))) fail(1, "icc / icmat pointer is NULL");
  // A plane source of a full-size 3D grid is millions of coordinates: the list is resolved by a few threads, each
  // over a contiguous range of rows, and concatenated in row order.
  struct Part {
    std::vector<unsigned char> plane;
    int rim = 0, kept = 0; bool bad = false;
  };
  const int n_thr = ncoords >= (1 << 17) ? (int)std::min<unsigned>(8, std::max(1u, std::thread::hardware_concurrency())) : 1;
  std::vector<Part> parts(n_thr);
  // every thread writes its rows in place; slab engines drop the sources of other slabs and compact afterwards
  h_src_idx.resize((size_t)ncoords); h_src_row.resize((size_t)ncoords); h_src_flag.resize((size_t)ncoords);
  tp("host lists allocated");
  auto work = [&](int k) {
    Part &P = parts[k];
    const int i0 = (int)((long long)ncoords * k / n_thr), i1 = (int)((long long)ncoords * (k + 1) / n_thr);
    P.plane.assign((size_t)nXl, 0);
    unsigned char *plane = P.plane.data();
    int o = i0, rim = 0;                               // (counters stay thread-local: the Parts share cache lines)
    for (int i = i0; i < i1; ++i) {
      const int32_t *c = icc + (size_t)i * nd;
      if (!coord_ok(c)) { P.bad = true; return; }
      if (c[0] < gx0 || c[0] >= gx0 + nXl) continue;
      const long long li = lin(c[0], c[1], nd == 3 ? c[2] : 0);
      const bool r = is_rim(c[0], c[1], nd == 3 ? c[2] : M);
      const bool dead = plane_air[c[0] - gx0] && air_set.count(li) != 0;
      h_src_idx[o] = li;
      h_src_row[o] = i;
      h_src_flag[o] = (unsigned char)((r ? 1 : 0) | (dead ? 2 : 0));
      ++o;
      rim += (r && !dead);
      if (!dead) plane[c[0] - gx0] = 1;
    }
    P.rim = rim;
    P.kept = o - i0;
  };
  {
    std::vector<std::thread> th;
    for (int k = 1; k < n_thr; ++k) th.emplace_back(work, k);
    work(0);
    for (auto &t_ : th) t_.join();
  }
  tp("coordinates resolved");
  size_t out = 0;
  for (int k = 0; k < n_thr; ++k) {
    Part &P = parts[k];
    if (P.bad) fail(1, "icc: source coordinate outside the grid");
    const size_t i0 = (size_t)((long long)ncoords * k / n_thr);
    if (out != i0 && P.kept) {
      std::move(h_src_idx.begin() + i0, h_src_idx.begin() + i0 + P.kept, h_src_idx.begin() + out);
      std::move(h_src_row.begin() + i0, h_src_row.begin() + i0 + P.kept, h_src_row.begin() + out);
      std::move(h_src_flag.begin() + i0, h_src_flag.begin() + i0 + P.kept, h_src_flag.begin() + out);
    }
    out += P.kept;
    n_src_rim += P.rim;
    for (int a = 0; a < nXl; ++a) plane_src[a] |= P.plane[a];
  }
  h_src_idx.resize(out); h_src_row.resize(out); h_src_flag.resize(out);
  n_src = (int)h_src_idx.size();
  tp("lists compacted");
  d_src_idx = salloc<long long>(n_src); d_src_row = salloc<int>(n_src); d_src_rim = salloc<unsigned char>(n_src);
  d_icmat = nullptr;
  tp("device lists allocated");
  if (n_src) {   // the host lists are members: nothing here waits for the copies
    FW_CUDA(cudaMemcpyAsync(d_src_idx, h_src_idx.data(), n_src * sizeof(long long), cudaMemcpyHostToDevice, stream));
    FW_CUDA(cudaMemcpyAsync(d_src_row, h_src_row.data(), n_src * sizeof(int), cudaMemcpyHostToDevice, stream));
    FW_CUDA(cudaMemcpyAsync(d_src_rim, h_src_flag.data(), n_src, cudaMemcpyHostToDevice, stream));
    const size_t nic = (size_t)ncoords * nTic;
    d_icmat = salloc<float>(nic);
    tp("index lists queued");
    if (nic) FW_CUDA(cudaMemcpyAsync(d_icmat, icmat, nic * 4, cudaMemcpyHostToDevice, stream));
    h2d_bytes += (int64_t)nic * 4;
    tp("signals queued");
  }
}

void Engine::reset(int nT_, int nTic_, int ncoords, const int32_t *icc, const float *icmat) {
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

void Engine::init(const fw25_problem &pb, const fw25_slab *slab, int dev) {
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
  // layer writes) the isotropic kernels run on the x-axis copy; otherwise the ANISO instantiations read all of them.
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
  const int n_state_needed = (pb.ext_p ? 0 : 1) + (pb.ext_u ? 0 : 1) + (pb.ext_v ? 0 : 1) +
                             (ndim == 3 ? (pb.ext_w ? 0 : 1) + 12 : 8);
  {
    const bool adopt = dev_maps && mp == G.pitch;
    const bool dc_copy = !adopt || (ndim == 3 && !pb.dcmap_full3d);
    const size_t want = (size_t)n_state_needed + (adopt ? 0 : 13 + (aniso ? 10 * (ndim - 1) : 0)) + (dc_copy ? 1 : 0);
    void *a = nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    if (cudaMalloc(&a, want * cells * sizeof(float)) == cudaSuccess) {
      arena = static_cast<char *>(a);
      arena_slices = want;
      owned.push_back(a);
    } else {
      cudaGetLastError();                  // no contiguous block of that size: fall back to one allocation per array
    }
    malloc_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  }
  auto up = [&](const float *src) { return upload_map(src, dev_maps, mp); };
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
  auto trace_point = [&](const char *what) {   // host clock only: no synchronisation, the trace must not change the run
    if (!trace) return;
    fprintf(stderr, "[fw25 setup] %-28s t = %8.1f ms   (cudaMalloc so far %.1f ms, h2d %.2f GB)\n", what,
            std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tr0).count() , malloc_ms,
            h2d_bytes / 1e9);
  };
  trace_point("maps uploaded");
  // state
  auto state = [&](float *ext) {
    float *d = ext ? ext : big();
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
  if (aniso && ws_supported(ndim, G)) {     // truly per-axis maps: the warp-specialised sweeps with 10 more tiles per plane
    std::string perr;
    std::vector<float> hd((size_t)18 * pb.ndmap);
    memcpy(hd.data(), pb.dmap, hd.size() * 4);
    ws = ws_plan_create(F, G, hd.data(), stream, &perr, /*aniso=*/true);
    if (!ws) fail(2, "warp-specialised anisotropic sweep setup failed: " + perr);
  }

  if (sweeps2d_supported(ndim, G)) {
    std::string perr;
    std::vector<float> hd((size_t)18 * pb.ndmap);
    memcpy(hd.data(), pb.dmap, hd.size() * 4);
    p2d = plan2d_create(F, G, hd.data(), stream, &perr, aniso);
    if (!p2d) fail(2, "2D sweep setup failed: " + perr);
  }

  // ---- coordinate lists -> linear indices (bit-exact integer maps)
  const int nd = ndim;
  plane_air.assign((size_t)nXl, 0);
  plane_sens.assign((size_t)nXl, 0);
  {  // air voxels (ghost planes included)
    std::vector<long long> idx;
    if (pb.ncoordszero > 0 && !pb.icczero) fail(1, "icczero pointer is NULL");
    for (int i = 0; i < (aniso_protocol ? 0 : pb.ncoordszero); ++i) {   // (the anisotropic binaries have no air kernel)
      const int32_t *c = pb.icczero + (size_t)i * nd;
      if (!coord_ok(c)) fail(1, "icczero: air coordinate outside the grid");
      if (c[0] < gx0 || c[0] >= gx0 + nXl) continue;
      idx.push_back(lin(c[0], c[1], nd == 3 ? c[2] : 0));
      plane_air[c[0] - gx0] = 1;
    }
    n_air = (int)idx.size();
    h_air_idx = idx;                         // (a member: the copy below need not be waited for)
    d_air_idx = dalloc<long long>(n_air);
    if (n_air) FW_CUDA(cudaMemcpyAsync(d_air_idx, h_air_idx.data(), n_air * sizeof(long long), cudaMemcpyHostToDevice, stream));
    air_set.insert(idx.begin(), idx.end());
  }
  trace_point("state, plans, air list");
  setup_sources(pb.ncoords, pb.icc, pb.icmat);
  trace_point("source list");
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
      if (idx.back() >= 0) plane_sens[c[0] - gx0] = 1;
    }
    n_sens = (int)idx.size();
    h_sens_idx = idx;
    d_sens_idx = dalloc<long long>(n_sens);
    if (n_sens) FW_CUDA(cudaMemcpyAsync(d_sens_idx, h_sens_idx.data(), n_sens * sizeof(long long), cudaMemcpyHostToDevice, stream));
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

void Engine::build_fuse_lists() {
  for (void *p : fuse_owned) cudaFree(p);
  fuse_owned.clear();
  fuse_ok = false;
  const char *ev = getenv("FW25_FUSE2D");
  if (!ev || atoi(ev) == 0) return;
  if (ndim != 2 || !p2d || aniso || own_lo != 0 || own_hi != nX_global || n_src_rim > 0 || !sweeps2d_fusable()) return;
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

void Engine::build_graph(bool with_inject) {
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

int Engine::advance(int max_steps, int frame_room) {
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

void Engine::read_frames(int f0, int f1, float *out) {
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

}  // namespace fw25
