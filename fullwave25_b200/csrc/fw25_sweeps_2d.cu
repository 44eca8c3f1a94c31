// fw25_sweeps_2d.cu -- 2D sweeps for sm_100a: TMA-staged stencil tiles.
//
// The 2D grids of the shipped examples are 0.4-3 M points: the whole problem lives in L2 or
// nearly so, a step takes tens of microseconds, and there is no third axis to march along.  ncu on the
// one-thread-per-cell L1/L2-path kernels (profiles/ncu_r01_2d.txt): 400 instructions per 32 cells, ~36 global loads
// of p per cell, long-scoreboard stalls of 19 warps per issue, DRAM at 41 % -- latency-bound.  Here the CTA is a
// small tile whose haloed stencil field arrives with ONE TMA load (cp.async.bulk.tensor.2d, out-of-range elements
// zero-filled) while the threads already fetch their point-wise operands (coefficients, memory variables, old
// fields: each touched exactly once, straight from global memory); all stencil taps are then LDS.
//   * k_sweep_*_2dc<TR>  (default, TR = 2): TR rows x 128 columns, one cell per thread -- the shortest dependency
//     chain, 32-41 registers, full occupancy; the (TR + 15)-row tile is shared by the TR rows.
//   * k_sweep_*_2d<RPT>  (FW25_2D_TR=0): RPT rows x 128 columns, 128 threads, each thread walks down RPT rows with
//     the 16-point x column in registers and loads the next row's operands ahead.  Fewer halo bytes per cell, but
//     less parallelism: on a B200 every RPT > 1 lost to RPT = 1 at every 2D size (profiles/sweep_2d_r01.txt).
//
// Arithmetic: operation for operation the reference's (2D PTX L38-463 / L465-891; fw25_kernels.cuh);
// bit-identical to k_sweep_*_simple<2> and the oracle.
#include <cuda.h>

#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "fw25_internal.h"
#include "fw25_kernels.cuh"
#include "fw25_tma.cuh"

#ifndef FW25_2D_TR_DEFAULT
#define FW25_2D_TR_DEFAULT 2   // 2 | 4 | 8: one-cell-per-thread tiles of that height (2 measured best); 0: marching tiles (pick_rpt)
#endif
#ifndef FW25_WS_2D_MINB
#define FW25_WS_2D_MINB 6   // resident CTAs per SM the register budget is sized for (80 registers, no spills)
#endif

namespace fw25 {

namespace {

constexpr int TC2 = 128;    // columns per CTA = threads per CTA

struct StencilTab2 {
  float4 d03, d47, e;
};

struct PwU {   // point-wise operands of one cell of fd_u
  float rho, K, kx, a1, b1, a2, b2, mA1, mA2, mC1, mC2, qA, qC;
  int ci;
};
struct PwAx {  // anisotropic family: kappa / a / b of the contiguous axis (the reference's y); axis x sits in Pw*
  float k, a1, b1, a2, b2;
};

// read-only maps (never written during a run: safe before pdl_wait) ...
__device__ __forceinline__ void load_maps_u(PwU &w, const Fields &F, long long i) {
  w.ci = __ldg(F.dcmap + i);
  w.rho = __ldg(F.rho + i); w.K = __ldg(F.K + i); w.kx = __ldg(F.kappax + i);
  w.a1 = __ldg(F.ax1 + i); w.b1 = __ldg(F.bx1 + i); w.a2 = __ldg(F.ax2 + i); w.b2 = __ldg(F.bx2 + i);
}
__device__ __forceinline__ PwAx load_axis_c_u(const Fields &F, long long i) {
  return PwAx{__ldg(F.kv[2] + i), __ldg(F.av[2][0] + i), __ldg(F.bv[2][0] + i), __ldg(F.av[2][1] + i), __ldg(F.bv[2][1] + i)};
}
__device__ __forceinline__ PwAx load_axis_c_p(const Fields &F, long long i) {
  return PwAx{__ldg(F.kp[2] + i), __ldg(F.ap[2][0] + i), __ldg(F.bp[2][0] + i), __ldg(F.ap[2][1] + i), __ldg(F.bp[2][1] + i)};
}
// ... and the state the previous sweeps wrote
__device__ __forceinline__ void load_state_u(PwU &w, const Fields &F, long long i) {
  w.mA1 = __ldcs(F.psi[0][0] + i); w.mA2 = __ldcs(F.psi[0][1] + i);
  w.mC1 = __ldcs(F.psi[2][0] + i); w.mC2 = __ldcs(F.psi[2][1] + i);
  w.qA = __ldcs(F.q[0] + i); w.qC = __ldcs(F.q[2] + i);
}
__device__ __forceinline__ PwU load_pw_u(const Fields &F, long long i) {
  PwU w;
  load_maps_u(w, F, i);
  load_state_u(w, F, i);
  return w;
}

struct PwP {   // point-wise operands of one cell of fd_p
  float K, beta, ku, a1, b1, a2, b2, fA1, fA2, fC1, fC2, p;
  int ci;
};

__device__ __forceinline__ void load_maps_p(PwP &w, const Fields &F, long long i) {
  w.ci = __ldg(F.dcmap + i);
  w.K = __ldg(F.K + i); w.beta = __ldg(F.beta + i); w.ku = __ldg(F.kappau + i);
  w.a1 = __ldg(F.au1 + i); w.b1 = __ldg(F.bu1 + i); w.a2 = __ldg(F.au2 + i); w.b2 = __ldg(F.bu2 + i);
}
__device__ __forceinline__ void load_state_p(PwP &w, const Fields &F, long long i) {
  w.fA1 = __ldcs(F.phi[0][0] + i); w.fA2 = __ldcs(F.phi[0][1] + i);
  w.fC1 = __ldcs(F.phi[2][0] + i); w.fC2 = __ldcs(F.phi[2][1] + i);
  w.p = F.p[i];
}
__device__ __forceinline__ PwP load_pw_p(const Fields &F, long long i) {
  PwP w;
  load_maps_p(w, F, i);
  load_state_p(w, F, i);
  return w;
}

// ------------------------------------------------------------------------------------------ fd_u (2D)
template <int RPT>
__global__ void __launch_bounds__(TC2, FW25_WS_2D_MINB)
    k_sweep_u_2d(const __grid_constant__ CUtensorMap map_p, const Fields F, const Geom G,
                 const StencilTab2 *__restrict__ tab, int a_lo, int a_hi) {
  constexpr int HR = RPT + 15, HC = TC2 + 16;
  __shared__ alignas(128) float tile[HR][HC];   // tile[r][j] = p[a0 - 7 + r][c0 - 8 + j]
  __shared__ uint64_t bar;
  const int tc = threadIdx.x;
  const int c0 = blockIdx.x * TC2;
  const int a0 = a_lo + blockIdx.y * RPT;
  if (tc == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
    mbar_arrive_expect_tx(&bar, HR * HC * 4);
    tma_load_2d(&tile[0][0], &map_p, c0 - 8, a0 - 7, &bar);
  }
  __syncthreads();                               // the initialised barrier is visible to every waiter

  const int c = c0 + tc;
  const bool act = c >= M && c < G.nC - M;
  const int rows = min(RPT, a_hi - a0);
  long long i = (long long)a0 * G.sA + c;
  PwU cur{};
  if (act) cur = load_pw_u(F, i);                // row 0's operands fly while the tile lands
  mbar_wait(&bar, 0);
  if (!act) return;

  const int sc = tc + 8;
  float pc[16];                                  // pc[j] = p[a - 7 + j][c]
#pragma unroll
  for (int j = 0; j < 15; ++j) pc[j + 1] = tile[j][sc];

#pragma unroll
  for (int r = 0; r < RPT; ++r) {
    if (r >= rows) break;
    const int sr = r + 7;
#pragma unroll
    for (int j = 0; j < 15; ++j) pc[j] = pc[j + 1];
    pc[15] = tile[sr + 8][sc];
    PwU nxt{};
    if (r + 1 < rows) nxt = load_pw_u(F, i + G.sA);

    const StencilTab2 T = tab[cur.ci];
    const float D[9] = {0.f, T.d03.x, T.d03.y, T.d03.z, T.d03.w, T.d47.x, T.d47.y, T.d47.z, T.d47.w};
    const float E = T.e.x;
    float yv[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) yv[j] = (j == 7) ? pc[7] : tile[sr][sc + j - 7];
    float gA = 0.f, gC = 0.f;
#pragma unroll
    for (int k = 1; k <= M; ++k) {
      gA = fma_(D[k], sub_(pc[7 + k], pc[8 - k]), gA);
      gC = fma_(D[k], sub_(yv[7 + k], yv[8 - k]), gC);
    }
    const float p11 = tile[sr + 1][sc + 1];
    float cA = sub_(p11, yv[8]);
    cA = add_(cA, tile[sr + 1][sc - 1]); cA = sub_(cA, yv[6]);
    float cC = sub_(p11, pc[8]);
    cC = add_(cC, tile[sr - 1][sc + 1]); cC = sub_(cC, pc[6]);
    const float dX = G.dX;
    gA = div_(fma_(E, cA, gA), dX);
    gC = div_(fma_(E, cC, gC), dX);

    const float s = div_(div_(G.dT, cur.rho), fma_(rcp_(cur.K), pc[7], 1.0f));
    const float mA1 = fma_(cur.b1, cur.mA1, mul_(gA, cur.a1));
    const float mA2 = fma_(cur.b2, cur.mA2, mul_(gA, cur.a2));
    const float mC1 = fma_(cur.b1, cur.mC1, mul_(gC, cur.a1));
    const float mC2 = fma_(cur.b2, cur.mC2, mul_(gC, cur.a2));
    const float qA = fma_(-s, add_(add_(div_(gA, cur.kx), mA1), mA2), cur.qA);
    const float qC = fma_(-s, add_(add_(div_(gC, cur.kx), mC1), mC2), cur.qC);
    __stcs(F.psi[0][0] + i, mA1); __stcs(F.psi[0][1] + i, mA2);
    __stcs(F.psi[2][0] + i, mC1); __stcs(F.psi[2][1] + i, mC2);
    F.q[0][i] = qA; F.q[2][i] = qC;              // the next sweep's stencil fields: default caching
    cur = nxt;
    i += G.sA;
  }
}

// ------------------------------------------------------------------------------------------ fd_p (2D)
template <int RPT>
__global__ void __launch_bounds__(TC2, FW25_WS_2D_MINB)
    k_sweep_p_2d(const __grid_constant__ CUtensorMap map_u, const __grid_constant__ CUtensorMap map_v, const Fields F,
                 const Geom G, const StencilTab2 *__restrict__ tab, int a_lo, int a_hi) {
  constexpr int UR = RPT + 15, UC = TC2 + 8;     // tu[r][j] = u[a0 - 8 + r][c0 - 4 + j]
  constexpr int VR = RPT + 2, VC = TC2 + 16;     // tv[r][j] = v[a0 - 1 + r][c0 - 8 + j]
  __shared__ alignas(128) float tu[UR][UC];
  __shared__ alignas(128) float tv[VR][VC];
  __shared__ uint64_t bar;
  const int tc = threadIdx.x;
  const int c0 = blockIdx.x * TC2;
  const int a0 = a_lo + blockIdx.y * RPT;
  if (tc == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
    mbar_arrive_expect_tx(&bar, (UR * UC + VR * VC) * 4);
    tma_load_2d(&tu[0][0], &map_u, c0 - 4, a0 - 8, &bar);
    tma_load_2d(&tv[0][0], &map_v, c0 - 8, a0 - 1, &bar);
  }
  __syncthreads();

  const int c = c0 + tc;
  const bool act = c >= M && c < G.nC - M;
  const int rows = min(RPT, a_hi - a0);
  long long i = (long long)a0 * G.sA + c;
  PwP cur{};
  if (act) cur = load_pw_p(F, i);
  mbar_wait(&bar, 0);
  if (!act) return;

  const int su = tc + 4, sv = tc + 8;
  float uc[16];                                  // uc[j] = u[a - 8 + j][c]
#pragma unroll
  for (int j = 0; j < 15; ++j) uc[j + 1] = tu[j][su];

#pragma unroll
  for (int r = 0; r < RPT; ++r) {
    if (r >= rows) break;
    const int ur = r + 8, vr = r + 1;
#pragma unroll
    for (int j = 0; j < 15; ++j) uc[j] = uc[j + 1];
    uc[15] = tu[ur + 7][su];
    PwP nxt{};
    if (r + 1 < rows) nxt = load_pw_p(F, i + G.sA);

    const StencilTab2 T = tab[cur.ci];
    const float D[9] = {0.f, T.d03.x, T.d03.y, T.d03.z, T.d03.w, T.d47.x, T.d47.y, T.d47.z, T.d47.w};
    const float E = T.e.x;
    float hA = 0.f, hC = 0.f;
#pragma unroll
    for (int k = 1; k <= M; ++k) {
      hA = fma_(D[k], sub_(uc[7 + k], uc[8 - k]), hA);
      hC = fma_(D[k], sub_(tv[vr][sv + k - 1], tv[vr][sv - k]), hC);
    }
    float cA = sub_(tu[ur][su + 1], tu[ur - 1][su + 1]);
    cA = add_(cA, tu[ur][su - 1]); cA = sub_(cA, tu[ur - 1][su - 1]);
    float cC = sub_(tv[vr + 1][sv], tv[vr + 1][sv - 1]);
    cC = add_(cC, tv[vr - 1][sv]); cC = sub_(cC, tv[vr - 1][sv - 1]);
    const float dX = G.dX;
    hA = div_(fma_(E, cA, hA), dX);
    hC = div_(fma_(E, cC, hC), dX);

    const float fA1 = fma_(cur.b1, cur.fA1, mul_(hA, cur.a1));
    const float fA2 = fma_(cur.b2, cur.fA2, mul_(hA, cur.a2));
    const float fC1 = fma_(cur.b1, cur.fC1, mul_(hC, cur.a1));
    const float fC2 = fma_(cur.b2, cur.fC2, mul_(hC, cur.a2));
    float S = add_(div_(hA, cur.ku), div_(hC, cur.ku));
    S = add_(fA1, S); S = add_(fA2, S); S = add_(fC1, S); S = add_(fC2, S);
    const float At = mul_(mul_(G.dT, cur.K), S);
    const float Bt = fma_(cur.p, mul_(rcp_(cur.K), sub_(1.0f, add_(cur.beta, cur.beta))), 1.0f);
    __stcs(F.phi[0][0] + i, fA1); __stcs(F.phi[0][1] + i, fA2);
    __stcs(F.phi[2][0] + i, fC1); __stcs(F.phi[2][1] + i, fC2);
    F.p[i] = fma_(-At, Bt, cur.p);
    cur = nxt;
    i += G.sA;
  }
}

// ------------------------------------------------------------------------------------------ one cell per thread
// TR rows x 128 columns per CTA, TR * 128 threads, one cell each: the haloed tile ((TR + 15) rows) is shared by the
// TR rows, so a taller CTA moves fewer halo bytes per cell than TR single-row CTAs while every thread keeps the
// shortest possible dependency chain (no marching).  The tile-height sweep on a B200 (profiles/sweep_2d_r01.txt)
// showed the single-row marching tiles beating the taller marching ones at every 2D size -- parallelism matters
// more than halo traffic there -- so this form keeps one cell per thread and shares the tile instead.
template <int TR, bool ANISO>
__global__ void __launch_bounds__(TC2 * TR)
    k_sweep_u_2dc(const __grid_constant__ CUtensorMap map_p, const Fields F, const Geom G,
                  const StencilTab2 *__restrict__ tab, int a_lo, int a_hi) {
  constexpr int HR = TR + 15, HC = TC2 + 16;
  __shared__ alignas(128) float tile[HR][HC];   // tile[r][j] = p[a0 - 7 + r][c0 - 8 + j]
  __shared__ uint64_t bar;
  const int tc = threadIdx.x, tr = threadIdx.y;
  const int c0 = blockIdx.x * TC2;
  const int a0 = a_lo + blockIdx.y * TR;
  pdl_trigger();                                 // the next kernel may be scheduled once every CTA is here
  if (tc == 0 && tr == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  __syncthreads();
  const int c = c0 + tc, a = a0 + tr;
  const bool act = c >= M && c < G.nC - M && a < a_hi;
  const long long i = (long long)a * G.sA + c;
  PwU w{};
  PwAx wc{};
  StencilTab2 T{};
  if (act) {                                     // coefficient maps and the stencil table: in flight before ...
    load_maps_u(w, F, i);
    if constexpr (ANISO) wc = load_axis_c_u(F, i);
    T = tab[w.ci];
  }
  pdl_wait();                                    // ... the previous kernel (injection / fd_p: they write p) is done
  if (tc == 0 && tr == 0) {
    mbar_arrive_expect_tx(&bar, HR * HC * 4);
    tma_load_2d(&tile[0][0], &map_p, c0 - 8, a0 - 7, &bar);
  }
  if (act) load_state_u(w, F, i);
  mbar_wait(&bar, 0);
  if (!act) return;
  const int sr = tr + 7, sc = tc + 8;
  const float D[9] = {0.f, T.d03.x, T.d03.y, T.d03.z, T.d03.w, T.d47.x, T.d47.y, T.d47.z, T.d47.w};
  const float E = T.e.x;
  float gA = 0.f, gC = 0.f;
#pragma unroll
  for (int k = 1; k <= M; ++k) {
    gA = fma_(D[k], sub_(tile[sr + k][sc], tile[sr + 1 - k][sc]), gA);
    gC = fma_(D[k], sub_(tile[sr][sc + k], tile[sr][sc + 1 - k]), gC);
  }
  const float pcen = tile[sr][sc], p11 = tile[sr + 1][sc + 1];
  float cA = sub_(p11, tile[sr][sc + 1]);
  cA = add_(cA, tile[sr + 1][sc - 1]); cA = sub_(cA, tile[sr][sc - 1]);
  float cC = sub_(p11, tile[sr + 1][sc]);
  cC = add_(cC, tile[sr - 1][sc + 1]); cC = sub_(cC, tile[sr - 1][sc]);
  const float dX = G.dX;
  gA = div_(fma_(E, cA, gA), dX);
  gC = div_(fma_(E, cC, gC), dX);
  const float s = div_(div_(G.dT, w.rho), fma_(rcp_(w.K), pcen, 1.0f));
  if constexpr (!ANISO) wc = PwAx{w.kx, w.a1, w.b1, w.a2, w.b2};
  const float mA1 = fma_(w.b1, w.mA1, mul_(gA, w.a1));
  const float mA2 = fma_(w.b2, w.mA2, mul_(gA, w.a2));
  const float mC1 = fma_(wc.b1, w.mC1, mul_(gC, wc.a1));
  const float mC2 = fma_(wc.b2, w.mC2, mul_(gC, wc.a2));
  const float qA = fma_(-s, add_(add_(div_(gA, w.kx), mA1), mA2), w.qA);
  const float qC = fma_(-s, add_(add_(div_(gC, wc.k), mC1), mC2), w.qC);
  __stcs(F.psi[0][0] + i, mA1); __stcs(F.psi[0][1] + i, mA2);
  __stcs(F.psi[2][0] + i, mC1); __stcs(F.psi[2][1] + i, mC2);
  F.q[0][i] = qA; F.q[2][i] = qC;
}

// FUSED: the epilogue also records this step's sensor frame and applies the NEXT step's source injection / air
// zeroing (the reference's launch order is inject(t) -> fd_u -> fd_p -> record(t); with the injection of step t + 1
// folded into fd_p(t) and record(t) taken from the freshly computed p' a graph-replayed 2D step is two kernels instead
// of four).  Sensors: box sensors by arithmetic on the cell's own coordinates; listed sensors, sources and air voxels
// through a per-tile CSR list built at setup (most tiles have none: two int loads per CTA).  A cell that is both
// sensor and source is recorded from the register / shared copy of p' before the injected value lands.
template <int TR, bool FUSED, bool ANISO>
__global__ void __launch_bounds__(TC2 * TR)
    k_sweep_p_2dc(const __grid_constant__ CUtensorMap map_u, const __grid_constant__ CUtensorMap map_v, const Fields F,
                  const Geom G, const StencilTab2 *__restrict__ tab, int a_lo, int a_hi, const Fuse2D X, int t_off,
                  int flags) {
  constexpr int UR = TR + 15, UC = TC2 + 8;      // tu[r][j] = u[a0 - 8 + r][c0 - 4 + j]
  constexpr int VR = TR + 2, VC = TC2 + 16;      // tv[r][j] = v[a0 - 1 + r][c0 - 8 + j]
  __shared__ alignas(128) float tu[UR][UC];
  __shared__ alignas(128) float tv[VR][VC];
  __shared__ uint64_t bar;
  const int tc = threadIdx.x, tr = threadIdx.y;
  const int c0 = blockIdx.x * TC2;
  const int a0 = a_lo + blockIdx.y * TR;
  pdl_trigger();
  if (tc == 0 && tr == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  __syncthreads();
  const int c = c0 + tc, a = a0 + tr;
  const bool act = c >= M && c < G.nC - M && a < a_hi;
  const long long i = (long long)a * G.sA + c;
  PwP w{};
  PwAx wc{};
  StencilTab2 T{};
  int e0 = 0, e1 = 0;
  if (act) {
    load_maps_p(w, F, i);
    if constexpr (ANISO) wc = load_axis_c_p(F, i);
    T = tab[w.ci];
  }
  if (FUSED) {                                   // the tile's list bounds are setup data too
    const int tile = blockIdx.y * gridDim.x + blockIdx.x;
    e0 = __ldg(X.tile_ofs + tile);
    e1 = __ldg(X.tile_ofs + tile + 1);
  }
  pdl_wait();                                    // fd_u (it writes u, v) is done
  if (tc == 0 && tr == 0) {
    mbar_arrive_expect_tx(&bar, (UR * UC + VR * VC) * 4);
    tma_load_2d(&tu[0][0], &map_u, c0 - 4, a0 - 8, &bar);
    tma_load_2d(&tv[0][0], &map_v, c0 - 8, a0 - 1, &bar);
  }
  if (act) load_state_p(w, F, i);
  mbar_wait(&bar, 0);
  float pn = 0.f;
  if (act) {
    const int su = tc + 4, sv = tc + 8, ur = tr + 8, vr = tr + 1;
    const float D[9] = {0.f, T.d03.x, T.d03.y, T.d03.z, T.d03.w, T.d47.x, T.d47.y, T.d47.z, T.d47.w};
    const float E = T.e.x;
    float hA = 0.f, hC = 0.f;
#pragma unroll
    for (int k = 1; k <= M; ++k) {
      hA = fma_(D[k], sub_(tu[ur + k - 1][su], tu[ur - k][su]), hA);
      hC = fma_(D[k], sub_(tv[vr][sv + k - 1], tv[vr][sv - k]), hC);
    }
    float cA = sub_(tu[ur][su + 1], tu[ur - 1][su + 1]);
    cA = add_(cA, tu[ur][su - 1]); cA = sub_(cA, tu[ur - 1][su - 1]);
    float cC = sub_(tv[vr + 1][sv], tv[vr + 1][sv - 1]);
    cC = add_(cC, tv[vr - 1][sv]); cC = sub_(cC, tv[vr - 1][sv - 1]);
    const float dX = G.dX;
    hA = div_(fma_(E, cA, hA), dX);
    hC = div_(fma_(E, cC, hC), dX);
    if constexpr (!ANISO) wc = PwAx{w.ku, w.a1, w.b1, w.a2, w.b2};
    const float fA1 = fma_(w.b1, w.fA1, mul_(hA, w.a1));
    const float fA2 = fma_(w.b2, w.fA2, mul_(hA, w.a2));
    const float fC1 = fma_(wc.b1, w.fC1, mul_(hC, wc.a1));
    const float fC2 = fma_(wc.b2, w.fC2, mul_(hC, wc.a2));
    float S = add_(div_(hA, w.ku), div_(hC, wc.k));
    S = add_(fA1, S); S = add_(fA2, S); S = add_(fC1, S); S = add_(fC2, S);
    const float At = mul_(mul_(G.dT, w.K), S);
    const float Bt = fma_(w.p, mul_(rcp_(w.K), sub_(1.0f, add_(w.beta, w.beta))), 1.0f);
    __stcs(F.phi[0][0] + i, fA1); __stcs(F.phi[0][1] + i, fA2);
    __stcs(F.phi[2][0] + i, fC1); __stcs(F.phi[2][1] + i, fC2);
    pn = fma_(-At, Bt, w.p);
    F.p[i] = pn;
  }
  if (!FUSED) return;

  const int t = *X.d_t + t_off;
  float *frame = X.frames + (size_t)((t / X.modT) % X.cap) * X.n_sens;
  if ((flags & FUSE_RECORD) && X.use_box && act) {           // box sensor: this cell's row follows from (a, c)
    const int ba = a - X.box.a0, bc = c - X.box.c0;
    if (ba >= 0 && ba < X.box.wa && bc >= 0 && bc < X.box.wc) frame[(size_t)ba * X.box.wc + bc] = pn;
  }
  if (e1 > e0) {                                             // uniform over the CTA
    float *pnew = &tu[0][0];                                 // the u tile is dead: reuse it for the tile's p'
    __syncthreads();
    pnew[tr * TC2 + tc] = pn;
    __syncthreads();                                         // p' stored (shared AND global) by every owner thread
    for (int e = e0 + tr * TC2 + tc; e < e1; e += TR * TC2) {
      const int kind = X.ent_kind[e], cell = X.ent_cell[e], row = X.ent_row[e];
      if (kind == FUSE_SENSOR) {
        if (flags & FUSE_RECORD) frame[row] = pnew[cell];
      } else if (flags & FUSE_INJECT) {
        const long long idx = (long long)(a0 + cell / TC2) * G.sA + c0 + cell % TC2;
        if (kind == FUSE_AIR) F.p[idx] = 0.0f;
        else if (t + 1 < X.nTic) F.p[idx] = X.icmat[(size_t)row * X.nTic + t + 1];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ host
using EncodeFn = CUresult (*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                              const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                              CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn encode_fn_2d() {
  static EncodeFn fn = [] {
    void *sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      return reinterpret_cast<EncodeFn>(sym);
    return (EncodeFn) nullptr;
  }();
  return fn;
}

bool tmap2d(CUtensorMap *m, const void *base, const Geom &G, int box_c, int box_a, std::string *err) {
  EncodeFn fn = encode_fn_2d();
  if (!fn) { *err = "cuTensorMapEncodeTiled is not available from this driver"; return false; }
  const cuuint64_t dims[2] = {(cuuint64_t)G.pitch, (cuuint64_t)G.nA};
  const cuuint64_t strides[1] = {(cuuint64_t)G.sA * 4};
  const cuuint32_t box[2] = {(cuuint32_t)box_c, (cuuint32_t)box_a};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char b[160];
    snprintf(b, sizeof b, "cuTensorMapEncodeTiled (2D) failed with CUresult %d (pitch %d, box %dx%d)", (int)r, G.pitch,
             box_c, box_a);
    *err = b;
    return false;
  }
  return true;
}

}  // namespace

constexpr int kRpt[4] = {1, 2, 4, 8};    // rows per CTA the kernels are instantiated for
static int rpt_slot(int rpt) { return rpt == 1 ? 0 : rpt == 2 ? 1 : rpt == 4 ? 2 : 3; }

struct Plan2D {
  bool aniso = false;            // per-axis kappa / a / b maps: the ANISO one-cell-per-thread tiles (TR = 2) only
  StencilTab2 *tab = nullptr;
  CUtensorMap p[4], u[4], v[4];   // one set of boxes per entry of kRpt
};

bool sweeps2d_supported(int ndim, const Geom &G) {
  return ndim == 2 && G.nB == 1 && G.pitch % 32 == 0 && G.nC > 2 * M && encode_fn_2d() != nullptr;
}

// The tiled sweeps beat the L1/L2-path kernels at every 2D size measured (0.39 M .. 3.2 M cells: +26 .. +40 %,
// profiles/sweep_2d_r01.txt), so auto mode always takes them; FW25_2D_MIN_CELLS sets a floor for experiments.
bool sweeps2d_worthwhile(const Geom &G, int rows) {
  const char *ev = getenv("FW25_2D_MIN_CELLS");
  const long long thr = ev ? atoll(ev) : 0;
  return (long long)rows * G.nC >= thr;
}

Plan2D *plan2d_create(const Fields &F, const Geom &G, const float *host_dmap, cudaStream_t st, std::string *err,
                      bool aniso) {
  auto *pl = new Plan2D();
  pl->aniso = aniso;
  bool ok = true;
  for (int v = 0; v < 4 && ok; ++v) {
    const int rpt = kRpt[v];
    ok = tmap2d(&pl->p[v], F.p, G, TC2 + 16, rpt + 15, err) && tmap2d(&pl->u[v], F.q[0], G, TC2 + 8, rpt + 15, err) &&
         tmap2d(&pl->v[v], F.q[2], G, TC2 + 16, rpt + 2, err);
  }
  if (!ok) { delete pl; return nullptr; }
  const int nd = G.ndmap;
  std::vector<StencilTab2> t(nd);
  for (int c = 0; c < nd; ++c) {
    auto Dk = [&](int k) { return host_dmap[(size_t)(2 * k) * nd + c]; };
    t[c].d03 = make_float4(Dk(1), Dk(2), Dk(3), Dk(4));
    t[c].d47 = make_float4(Dk(5), Dk(6), Dk(7), Dk(8));
    t[c].e = make_float4(host_dmap[(size_t)3 * nd + c], 0.f, 0.f, 0.f);
  }
  if (cudaMalloc(&pl->tab, sizeof(StencilTab2) * nd) != cudaSuccess ||
      cudaMemcpyAsync(pl->tab, t.data(), sizeof(StencilTab2) * nd, cudaMemcpyHostToDevice, st) != cudaSuccess ||
      cudaStreamSynchronize(st) != cudaSuccess) {
    *err = std::string("2D sweep plan: ") + cudaGetErrorString(cudaGetLastError());
    if (pl->tab) cudaFree(pl->tab);
    delete pl;
    return nullptr;
  }
  return pl;
}

void plan2d_destroy(Plan2D *pl) {
  if (!pl) return;
  if (pl->tab) cudaFree(pl->tab);
  delete pl;
}

// rows per thread of the marching tiles.  Measured on a B200 at every 2D size from 0.39 M to 3.2 M cells
// (profiles/sweep_2d_r01.txt): 1 row beats 2 beats 4 beats 8 -- the 16-row haloed tile comes out of L2 through TMA
// for next to nothing, parallelism and short dependency chains are what pays.  FW25_2D_RPT overrides.
static int pick_rpt(const Geom &G, int rows) {
  (void)G; (void)rows;
  const char *ev = getenv("FW25_2D_RPT");
  const int env = ev ? atoi(ev) : 0;
  if (env == 1 || env == 2 || env == 4 || env == 8) return env;
  return 1;
}

// FW25_2D_TR = 2 | 4 | 8: one-cell-per-thread tiles of that many rows (default, see pick_tr); 0: marching tiles
static int pick_tr() {
  const char *ev = getenv("FW25_2D_TR");
  const int v = ev ? atoi(ev) : FW25_2D_TR_DEFAULT;
  return (v == 2 || v == 4 || v == 8) ? v : 0;
}

int launch_sweep_u_2d(const Plan2D *pl, const Fields &F, const Geom &G, int a_lo, int a_hi, cudaStream_t st) {
  if (a_hi <= a_lo) return 0;
  if (const int tr = pl->aniso ? 2 : pick_tr()) {
    dim3 grd((G.nC - M + TC2 - 1) / TC2, (a_hi - a_lo + tr - 1) / tr, 1), blk(TC2, tr, 1);
    const CUtensorMap &mp = pl->p[rpt_slot(tr)];
    if (pl->aniso) launch_pdl(k_sweep_u_2dc<2, true>, dim3(grd.x, (a_hi - a_lo + 1) / 2, 1), dim3(TC2, 2, 1), 0, st,
                              pl->p[rpt_slot(2)], F, G, pl->tab, a_lo, a_hi);
    else if (tr == 8) launch_pdl(k_sweep_u_2dc<8, false>, grd, blk, 0, st, mp, F, G, pl->tab, a_lo, a_hi);
    else if (tr == 4) launch_pdl(k_sweep_u_2dc<4, false>, grd, blk, 0, st, mp, F, G, pl->tab, a_lo, a_hi);
    else launch_pdl(k_sweep_u_2dc<2, false>, grd, blk, 0, st, mp, F, G, pl->tab, a_lo, a_hi);
    return 1;
  }
  const int rpt = pick_rpt(G, a_hi - a_lo);
  dim3 grd((G.nC - M + TC2 - 1) / TC2, (a_hi - a_lo + rpt - 1) / rpt, 1);
  const CUtensorMap &mp = pl->p[rpt_slot(rpt)];
  if (rpt == 8) k_sweep_u_2d<8><<<grd, TC2, 0, st>>>(mp, F, G, pl->tab, a_lo, a_hi);
  else if (rpt == 4) k_sweep_u_2d<4><<<grd, TC2, 0, st>>>(mp, F, G, pl->tab, a_lo, a_hi);
  else if (rpt == 2) k_sweep_u_2d<2><<<grd, TC2, 0, st>>>(mp, F, G, pl->tab, a_lo, a_hi);
  else k_sweep_u_2d<1><<<grd, TC2, 0, st>>>(mp, F, G, pl->tab, a_lo, a_hi);
  return 1;
}

int launch_sweep_p_2d(const Plan2D *pl, const Fields &F, const Geom &G, int a_lo, int a_hi, cudaStream_t st) {
  if (a_hi <= a_lo) return 0;
  if (const int tr = pl->aniso ? 2 : pick_tr()) {
    dim3 grd((G.nC - M + TC2 - 1) / TC2, (a_hi - a_lo + tr - 1) / tr, 1), blk(TC2, tr, 1);
    const CUtensorMap &mu = pl->u[rpt_slot(tr)], &mv = pl->v[rpt_slot(tr)];
    const Fuse2D none{};
    if (pl->aniso) launch_pdl(k_sweep_p_2dc<2, false, true>, dim3(grd.x, (a_hi - a_lo + 1) / 2, 1), dim3(TC2, 2, 1), 0, st,
                              pl->u[rpt_slot(2)], pl->v[rpt_slot(2)], F, G, pl->tab, a_lo, a_hi, none, 0, 0);
    else if (tr == 8) launch_pdl(k_sweep_p_2dc<8, false, false>, grd, blk, 0, st, mu, mv, F, G, pl->tab, a_lo, a_hi, none, 0, 0);
    else if (tr == 4) launch_pdl(k_sweep_p_2dc<4, false, false>, grd, blk, 0, st, mu, mv, F, G, pl->tab, a_lo, a_hi, none, 0, 0);
    else launch_pdl(k_sweep_p_2dc<2, false, false>, grd, blk, 0, st, mu, mv, F, G, pl->tab, a_lo, a_hi, none, 0, 0);
    return 1;
  }
  const int rpt = pick_rpt(G, a_hi - a_lo);
  dim3 grd((G.nC - M + TC2 - 1) / TC2, (a_hi - a_lo + rpt - 1) / rpt, 1);
  const CUtensorMap &mu = pl->u[rpt_slot(rpt)], &mv = pl->v[rpt_slot(rpt)];
  if (rpt == 8) k_sweep_p_2d<8><<<grd, TC2, 0, st>>>(mu, mv, F, G, pl->tab, a_lo, a_hi);
  else if (rpt == 4) k_sweep_p_2d<4><<<grd, TC2, 0, st>>>(mu, mv, F, G, pl->tab, a_lo, a_hi);
  else if (rpt == 2) k_sweep_p_2d<2><<<grd, TC2, 0, st>>>(mu, mv, F, G, pl->tab, a_lo, a_hi);
  else k_sweep_p_2d<1><<<grd, TC2, 0, st>>>(mu, mv, F, G, pl->tab, a_lo, a_hi);
  return 1;
}

// ---- fused fd_p (2D, graph-replayed whole-grid steps): see k_sweep_p_2dc<TR, FUSED>
bool sweeps2d_fusable() { return pick_tr() == FUSE_TR; }

int launch_sweep_p_2d_fused(const Plan2D *pl, const Fields &F, const Geom &G, int a_lo, int a_hi, cudaStream_t st,
                            const Fuse2D &X, int t_off, int flags) {
  if (a_hi <= a_lo) return 0;
  dim3 grd((G.nC - M + TC2 - 1) / TC2, (a_hi - a_lo + FUSE_TR - 1) / FUSE_TR, 1), blk(TC2, FUSE_TR, 1);
  const CUtensorMap &mu = pl->u[rpt_slot(FUSE_TR)], &mv = pl->v[rpt_slot(FUSE_TR)];
  launch_pdl(k_sweep_p_2dc<FUSE_TR, true, false>, grd, blk, 0, st, mu, mv, F, G, pl->tab, a_lo, a_hi, X, t_off, flags);
  return 1;
}

}  // namespace fw25
