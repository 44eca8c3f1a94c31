// fw25_sweeps_tiled.cu -- 3D sweeps for sm_100a: 2.5-D marching along x (the slowest, slab axis) over
// (y,z) tiles of TY x 32 cells.
//
//   * the 16-point x column of the stencil field lives in registers and shifts one plane per step;
//   * the (y,z) neighbourhood of the current plane (8-cell halo) and of the next plane (cross terms)
//     comes from a 4-deep shared-memory ring of haloed tiles filled by TMA (cp.async.bulk.tensor.3d,
//     zero fill outside the array) two planes ahead, completion tracked with mbarriers;
//   * the 17/18 point-wise arrays are read once and the 9/7 results written once with streaming
//     (evict-first) accesses, 128 B per warp, so the only re-read traffic is the tile halo, served by L2;
//   * no second time level and no proceed_time copy (each sweep writes only arrays it reads point-wise).
//
// Replaces the reference's fd_u / fd_p launches (3D PTX L38-675 / L677-1323).
// The arithmetic is operation-for-operation the reference's (see fw25_kernels.cuh); only the order in
// which operands are fetched differs.
#include <cuda.h>

#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>
#include <cstddef>

#include "fw25_internal.h"
#include "fw25_kernels.cuh"
#include "fw25_tma.cuh"

namespace fw25 {

namespace {

constexpr int TZ = 32;   // one warp spans the contiguous axis: 128-byte rows
constexpr int NS = 4;    // ring depth: planes x, x+1 in use, x+2, x+3 in flight

struct StencilTab {      // per sound-speed column: D1..D8, E (dmap[2k][c], dmap[3][c]), padded to 48 B
  float4 d03, d47, e;
};

// ------------------------------------------------------------------------------------------ fd_u
template <int TY, int MINB>
__global__ void __launch_bounds__(TY *TZ, MINB)
    k_sweep_u_tiled(const __grid_constant__ CUtensorMap tm_p, const Fields F, const Geom G,
                    const StencilTab *__restrict__ tab, int a_lo, int a_hi, int Lx) {
  constexpr int HY = TY + 2 * M, HZ = TZ + 2 * M;
  constexpr uint32_t TILE_BYTES = HY * HZ * sizeof(float);
  __shared__ __align__(128) float sp[NS][HY][HZ];
  __shared__ __align__(8) uint64_t bar[NS];

  const int tz = threadIdx.x, ty = threadIdx.y;
  const int z0 = blockIdx.x * TZ, y0 = M + blockIdx.y * TY;
  const int xa = a_lo + blockIdx.z * Lx;
  const int xb = min(xa + Lx, a_hi);
  const int z = z0 + tz, y = y0 + ty;
  const bool act = (z >= M) && (z < G.nC - M) && (y < G.nB - M);
  const bool lead = (tz == 0) && (ty == 0);
  const int sy = ty + M, sz = tz + M;

  if (lead) {
    tma_prefetch_desc(&tm_p);
#pragma unroll
    for (int s = 0; s < NS; ++s) mbar_init(&bar[s], 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (lead) {
    for (int P = xa; P <= min(xa + NS - 1, xb); ++P) {
      mbar_arrive_expect_tx(&bar[P & 3], TILE_BYTES);
      tma_load_3d(&sp[P & 3][0][0], &tm_p, z0 - M, y0 - M, P, &bar[P & 3]);
    }
  }

  const long long sA = G.sA, sB = G.sB;
  long long i = (long long)xa * sA + (long long)y * sB + z;
  const float *__restrict__ p = F.p;

  // x column: pc[j] = p[x - 7 + j]; pc[15] (= p[x + 8]) is fetched at the top of every step
  float pc[16];
  float pm_y1 = 0.f, pm_z1 = 0.f;   // p[x-1, y+1, z], p[x-1, y, z+1] carried from the previous step
  if (act) {
#pragma unroll
    for (int j = 0; j < 15; ++j) pc[j] = p[i + (j - 7) * sA];
    pm_y1 = p[i - sA + sB];
    pm_z1 = p[i - sA + 1];
  } else {
#pragma unroll
    for (int j = 0; j < 15; ++j) pc[j] = 0.f;
  }
  pc[15] = 0.f;

  mbar_wait(&bar[xa & 3], 0);

#pragma unroll 4
  for (int x = xa; x < xb; ++x, i += sA) {
    float rho = 1.f, Kc = 1.f, kx = 1.f, a1 = 0.f, b1 = 0.f, a2 = 0.f, b2 = 0.f;
    float q0 = 0.f, q1 = 0.f, q2 = 0.f, m00 = 0.f, m01 = 0.f, m10 = 0.f, m11 = 0.f, m20 = 0.f, m21 = 0.f;
    int ci = 0;
    if (act) {
      pc[15] = p[i + 8 * sA];
      ci = __ldcs(F.dcmap + i);
      rho = __ldcs(F.rho + i); Kc = __ldcs(F.K + i); kx = __ldcs(F.kappax + i);
      a1 = __ldcs(F.ax1 + i); b1 = __ldcs(F.bx1 + i); a2 = __ldcs(F.ax2 + i); b2 = __ldcs(F.bx2 + i);
      q0 = __ldcs(F.q[0] + i); q1 = __ldcs(F.q[1] + i); q2 = __ldcs(F.q[2] + i);
      m00 = __ldcs(F.psi[0][0] + i); m01 = __ldcs(F.psi[0][1] + i);
      m10 = __ldcs(F.psi[1][0] + i); m11 = __ldcs(F.psi[1][1] + i);
      m20 = __ldcs(F.psi[2][0] + i); m21 = __ldcs(F.psi[2][1] + i);
    }
    const int n1 = x + 1 - xa;
    mbar_wait(&bar[(x + 1) & 3], (n1 >> 2) & 1);   // plane x+1 has landed (plane x was waited earlier)

    if (act) {
      const float(*c0)[HZ] = sp[x & 3];
      const float(*c1)[HZ] = sp[(x + 1) & 3];
      const StencilTab T = tab[ci];
      const float D[9] = {0.f, T.d03.x, T.d03.y, T.d03.z, T.d03.w, T.d47.x, T.d47.y, T.d47.z, T.d47.w};
      const float E = T.e.x;
      const float pcen = pc[7];

      // taps of the current plane: yv[j] = P(0, j-7, 0), zv[j] = P(0, 0, j-7); j = 7 is the centre
      float yv[16], zv[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        yv[j] = (j == 7) ? pcen : c0[sy + j - 7][sz];
        zv[j] = (j == 7) ? pcen : c0[sy][sz + j - 7];
      }
      float gA = 0.f, gB = 0.f, gC = 0.f;
#pragma unroll
      for (int k = 1; k <= M; ++k) {   // ascending k, fma accumulate (PTX L201-319)
        gA = fma_(D[k], sub_(pc[7 + k], pc[8 - k]), gA);
        gB = fma_(D[k], sub_(yv[7 + k], yv[8 - k]), gB);
        gC = fma_(D[k], sub_(zv[7 + k], zv[8 - k]), gC);
      }
      // transverse corrections, strictly left to right (PTX L472-547)
      const float p110 = c1[sy + 1][sz], p1m0 = c1[sy - 1][sz], p101 = c1[sy][sz + 1], p10m = c1[sy][sz - 1];
      const float p011 = c0[sy + 1][sz + 1], p01m = c0[sy + 1][sz - 1], p0m1 = c0[sy - 1][sz + 1];
      const float p100 = pc[8], pm00 = pc[6];
      const float p010 = yv[8], p0m0 = yv[6], p001 = zv[8], p00m = zv[6];
      float cA = sub_(p110, p010);
      cA = add_(cA, p1m0); cA = sub_(cA, p0m0);
      cA = add_(cA, p101); cA = sub_(cA, p001);
      cA = add_(cA, p10m); cA = sub_(cA, p00m);
      float cB = sub_(p110, p100);
      cB = add_(cB, pm_y1); cB = sub_(cB, pm00);
      cB = add_(cB, p011); cB = sub_(cB, p001);
      cB = add_(cB, p01m); cB = sub_(cB, p00m);
      float cC = sub_(p101, p100);
      cC = add_(cC, pm_z1); cC = sub_(cC, pm00);
      cC = add_(cC, p011); cC = sub_(cC, p010);
      cC = add_(cC, p0m1); cC = sub_(cC, p0m0);
      pm_y1 = p010; pm_z1 = p001;

      const float dX = G.dX;
      gA = div_(fma_(E, cA, gA), dX);
      gB = div_(fma_(E, cB, gB), dX);
      gC = div_(fma_(E, cC, gC), dX);

      const float s = div_(div_(G.dT, rho), fma_(rcp_(Kc), pcen, 1.0f));
      m00 = fma_(b1, m00, mul_(gA, a1)); m01 = fma_(b2, m01, mul_(gA, a2));
      m10 = fma_(b1, m10, mul_(gB, a1)); m11 = fma_(b2, m11, mul_(gB, a2));
      m20 = fma_(b1, m20, mul_(gC, a1)); m21 = fma_(b2, m21, mul_(gC, a2));
      q0 = fma_(-s, add_(add_(div_(gA, kx), m00), m01), q0);
      q1 = fma_(-s, add_(add_(div_(gB, kx), m10), m11), q1);
      q2 = fma_(-s, add_(add_(div_(gC, kx), m20), m21), q2);
      __stcs(F.psi[0][0] + i, m00); __stcs(F.psi[0][1] + i, m01);
      __stcs(F.psi[1][0] + i, m10); __stcs(F.psi[1][1] + i, m11);
      __stcs(F.psi[2][0] + i, m20); __stcs(F.psi[2][1] + i, m21);
      __stcs(F.q[0] + i, q0); __stcs(F.q[1] + i, q1); __stcs(F.q[2] + i, q2);
#pragma unroll
      for (int j = 0; j < 15; ++j) pc[j] = pc[j + 1];
    }
    __syncthreads();   // every warp is done with plane x: its slot can be refilled
    if (lead && x + NS <= xb) {
      mbar_arrive_expect_tx(&bar[x & 3], TILE_BYTES);
      tma_load_3d(&sp[x & 3][0][0], &tm_p, z0 - M, y0 - M, x + NS, &bar[x & 3]);
    }
  }
}

// ------------------------------------------------------------------------------------------ fd_p
template <int TY>
struct alignas(128) PStage {        // one plane of the three velocity tiles (TMA destinations: 128 B aligned)
  alignas(128) float u[TY + 2][TZ + 8];          // rows y0-1 .. y0+TY,   z0-4 .. z0+TZ+3
  alignas(128) float v[TY + 2 * M][TZ + 8];      // rows y0-8 .. y0+TY+7, z0-4 .. z0+TZ+3
  alignas(128) float w[TY + 2][TZ + 2 * M];      // rows y0-1 .. y0+TY,   z0-8 .. z0+TZ+7
};

template <int TY, int MINB>
__global__ void __launch_bounds__(TY *TZ, MINB)
    k_sweep_p_tiled(const __grid_constant__ CUtensorMap tm_u, const __grid_constant__ CUtensorMap tm_v,
                    const __grid_constant__ CUtensorMap tm_w, const Fields F, const Geom G,
                    const StencilTab *__restrict__ tab, int a_lo, int a_hi, int Lx) {
  constexpr int UY = TY + 2, UZ = TZ + 8, VY = TY + 2 * M, VZ = TZ + 8, WY = TY + 2, WZ = TZ + 2 * M;
  constexpr uint32_t U_BYTES = UY * UZ * 4, V_BYTES = VY * VZ * 4, W_BYTES = WY * WZ * 4;
  static_assert(sizeof(PStage<TY>) % 128 == 0 && offsetof(PStage<TY>, v) % 128 == 0 &&
                    offsetof(PStage<TY>, w) % 128 == 0, "TMA destinations must be 128-byte aligned");
  __shared__ __align__(128) PStage<TY> st[NS];
  __shared__ __align__(8) uint64_t bar[NS];

  const int tz = threadIdx.x, ty = threadIdx.y;
  const int z0 = blockIdx.x * TZ, y0 = M + blockIdx.y * TY;
  const int xa = a_lo + blockIdx.z * Lx;
  const int xb = min(xa + Lx, a_hi);
  const int z = z0 + tz, y = y0 + ty;
  const bool act = (z >= M) && (z < G.nC - M) && (y < G.nB - M);
  const bool lead = (tz == 0) && (ty == 0);

  auto issue = [&](int P) {
    PStage<TY> &S = st[P & 3];
    mbar_arrive_expect_tx(&bar[P & 3], U_BYTES + V_BYTES + W_BYTES);
    tma_load_3d(&S.u[0][0], &tm_u, z0 - 4, y0 - 1, P, &bar[P & 3]);
    tma_load_3d(&S.v[0][0], &tm_v, z0 - 4, y0 - M, P, &bar[P & 3]);
    tma_load_3d(&S.w[0][0], &tm_w, z0 - M, y0 - 1, P, &bar[P & 3]);
  };

  if (lead) {
    tma_prefetch_desc(&tm_u); tma_prefetch_desc(&tm_v); tma_prefetch_desc(&tm_w);
#pragma unroll
    for (int s = 0; s < NS; ++s) mbar_init(&bar[s], 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (lead)
    for (int P = xa; P <= min(xa + NS - 1, xb); ++P) issue(P);

  const long long sA = G.sA, sB = G.sB;
  long long i = (long long)xa * sA + (long long)y * sB + z;
  const float *__restrict__ u = F.q[0];
  const float *__restrict__ v = F.q[1];
  const float *__restrict__ w = F.q[2];

  // x column of u: uc[j] = u[x - 8 + j]; uc[15] (= u[x + 7]) is fetched at the top of every step
  float uc[16];
  // carried from the previous step (plane x-1): U(-1,+-1,0), U(-1,0,+-1); V(-1,0,0), V(-1,-1,0); W(-1,0,0), W(-1,0,-1)
  float um_y1 = 0.f, um_ym = 0.f, um_z1 = 0.f, um_zm = 0.f, vm_0 = 0.f, vm_m = 0.f, wm_0 = 0.f, wm_m = 0.f;
  // current plane centre pair, filled from plane x+1 of the previous step: V(0,0,0), V(0,-1,0), W(0,0,0), W(0,0,-1)
  float v0_0 = 0.f, v0_m = 0.f, w0_0 = 0.f, w0_m = 0.f;
  if (act) {
#pragma unroll
    for (int j = 0; j < 15; ++j) uc[j] = u[i + (j - 8) * sA];
    const long long im = i - sA;
    um_y1 = u[im + sB]; um_ym = u[im - sB]; um_z1 = u[im + 1]; um_zm = u[im - 1];
    vm_0 = v[im]; vm_m = v[im - sB]; wm_0 = w[im]; wm_m = w[im - 1];
    v0_0 = v[i]; v0_m = v[i - sB]; w0_0 = w[i]; w0_m = w[i - 1];
  } else {
#pragma unroll
    for (int j = 0; j < 15; ++j) uc[j] = 0.f;
  }
  uc[15] = 0.f;

  mbar_wait(&bar[xa & 3], 0);

#pragma unroll 4
  for (int x = xa; x < xb; ++x, i += sA) {
    float Kc = 1.f, bt = 0.f, ku = 1.f, a1 = 0.f, b1 = 0.f, a2 = 0.f, b2 = 0.f, pc = 0.f;
    float f00 = 0.f, f01 = 0.f, f10 = 0.f, f11 = 0.f, f20 = 0.f, f21 = 0.f;
    int ci = 0;
    if (act) {
      uc[15] = u[i + 7 * sA];
      ci = __ldcs(F.dcmap + i);
      Kc = __ldcs(F.K + i); bt = __ldcs(F.beta + i); ku = __ldcs(F.kappau + i);
      a1 = __ldcs(F.au1 + i); b1 = __ldcs(F.bu1 + i); a2 = __ldcs(F.au2 + i); b2 = __ldcs(F.bu2 + i);
      pc = __ldcs(F.p + i);
      f00 = __ldcs(F.phi[0][0] + i); f01 = __ldcs(F.phi[0][1] + i);
      f10 = __ldcs(F.phi[1][0] + i); f11 = __ldcs(F.phi[1][1] + i);
      f20 = __ldcs(F.phi[2][0] + i); f21 = __ldcs(F.phi[2][1] + i);
    }
    const int n1 = x + 1 - xa;
    mbar_wait(&bar[(x + 1) & 3], (n1 >> 2) & 1);

    if (act) {
      const PStage<TY> &S0 = st[x & 3];
      const PStage<TY> &S1 = st[(x + 1) & 3];
      const StencilTab T = tab[ci];
      const float D[9] = {0.f, T.d03.x, T.d03.y, T.d03.z, T.d03.w, T.d47.x, T.d47.y, T.d47.z, T.d47.w};
      const float E = T.e.x;
      const int uy = ty + 1, uz = tz + 4;      // centre in the u tile
      const int vy = ty + M, vz = tz + 4;      // centre in the v tile
      const int wy = ty + 1, wz = tz + M;      // centre in the w tile

      // v taps along y: vv[j] = V(0, j-8, 0), j = 0..15 (8 = centre, 7 = y-1: registers)
      // w taps along z: wv[j] = W(0, 0, j-8)
      float vv[16], wv[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        vv[j] = (j == 8) ? v0_0 : (j == 7) ? v0_m : S0.v[vy + j - 8][vz];
        wv[j] = (j == 8) ? w0_0 : (j == 7) ? w0_m : S0.w[wy][wz + j - 8];
      }
      float hA = 0.f, hB = 0.f, hC = 0.f;
#pragma unroll
      for (int k = 1; k <= M; ++k) {   // PTX L799-966
        hA = fma_(D[k], sub_(uc[7 + k], uc[8 - k]), hA);
        hB = fma_(D[k], sub_(vv[7 + k], vv[8 - k]), hB);
        hC = fma_(D[k], sub_(wv[7 + k], wv[8 - k]), hC);
      }
      // cross terms (PTX L1120-1219)
      const float u0_y1 = S0.u[uy + 1][uz], u0_ym = S0.u[uy - 1][uz], u0_z1 = S0.u[uy][uz + 1], u0_zm = S0.u[uy][uz - 1];
      float cA = sub_(u0_y1, um_y1);
      cA = add_(cA, u0_ym); cA = sub_(cA, um_ym);
      cA = add_(cA, u0_z1); cA = sub_(cA, um_z1);
      cA = add_(cA, u0_zm); cA = sub_(cA, um_zm);
      const float vp_0 = S1.v[vy][vz], vp_m = S1.v[vy - 1][vz];
      const float v0_z1 = S0.v[vy][vz + 1], vm_z1 = S0.v[vy - 1][vz + 1], v0_zm = S0.v[vy][vz - 1], vm_zm = S0.v[vy - 1][vz - 1];
      float cB = sub_(vp_0, vp_m);
      cB = add_(cB, vm_0); cB = sub_(cB, vm_m);
      cB = add_(cB, v0_z1); cB = sub_(cB, vm_z1);
      cB = add_(cB, v0_zm); cB = sub_(cB, vm_zm);
      const float wp_0 = S1.w[wy][wz], wp_m = S1.w[wy][wz - 1];
      const float w0_y1 = S0.w[wy + 1][wz], wm_y1 = S0.w[wy + 1][wz - 1], w0_ym = S0.w[wy - 1][wz], wm_ym = S0.w[wy - 1][wz - 1];
      float cC = sub_(wp_0, wp_m);
      cC = add_(cC, wm_0); cC = sub_(cC, wm_m);
      cC = add_(cC, w0_y1); cC = sub_(cC, wm_y1);
      cC = add_(cC, w0_ym); cC = sub_(cC, wm_ym);
      // carry to the next step
      um_y1 = u0_y1; um_ym = u0_ym; um_z1 = u0_z1; um_zm = u0_zm;
      vm_0 = v0_0; vm_m = v0_m; wm_0 = w0_0; wm_m = w0_m;
      v0_0 = vp_0; v0_m = vp_m; w0_0 = wp_0; w0_m = wp_m;

      const float dX = G.dX;
      hA = div_(fma_(E, cA, hA), dX);
      hB = div_(fma_(E, cB, hB), dX);
      hC = div_(fma_(E, cC, hC), dX);

      f00 = fma_(b1, f00, mul_(hA, a1)); f01 = fma_(b2, f01, mul_(hA, a2));
      f10 = fma_(b1, f10, mul_(hB, a1)); f11 = fma_(b2, f11, mul_(hB, a2));
      f20 = fma_(b1, f20, mul_(hC, a1)); f21 = fma_(b2, f21, mul_(hC, a2));
      float Ssum = add_(div_(hA, ku), div_(hB, ku));      // PTX L1290-1305
      Ssum = add_(div_(hC, ku), Ssum);
      Ssum = add_(f00, Ssum); Ssum = add_(f01, Ssum); Ssum = add_(f10, Ssum); Ssum = add_(f11, Ssum);
      Ssum = add_(f20, Ssum); Ssum = add_(f21, Ssum);
      const float At = mul_(mul_(G.dT, Kc), Ssum);
      const float Bt = fma_(pc, mul_(rcp_(Kc), sub_(1.0f, add_(bt, bt))), 1.0f);
      __stcs(F.phi[0][0] + i, f00); __stcs(F.phi[0][1] + i, f01);
      __stcs(F.phi[1][0] + i, f10); __stcs(F.phi[1][1] + i, f11);
      __stcs(F.phi[2][0] + i, f20); __stcs(F.phi[2][1] + i, f21);
      F.p[i] = fma_(-At, Bt, pc);   // p is the next sweep's stencil field: keep it in L2
#pragma unroll
      for (int j = 0; j < 15; ++j) uc[j] = uc[j + 1];
    }
    __syncthreads();
    if (lead && x + NS <= xb) issue(x + NS);
  }
}

// ------------------------------------------------------------------------------------------ host
using EncodeFn = CUresult (*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                              const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                              CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn encode_fn() {
  static EncodeFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void *sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeFn>(sym);
  });
  return fn;
}

bool make_tmap3d(CUtensorMap *m, const float *base, const Geom &G, int box_c, int box_b, std::string *err) {
  EncodeFn fn = encode_fn();
  if (!fn) { *err = "cuTensorMapEncodeTiled is not available from this driver"; return false; }
  const cuuint64_t dims[3] = {(cuuint64_t)G.pitch, (cuuint64_t)G.nB, (cuuint64_t)G.nA};
  const cuuint64_t strides[2] = {(cuuint64_t)G.sB * 4, (cuuint64_t)G.sA * 4};
  const cuuint32_t box[3] = {(cuuint32_t)box_c, (cuuint32_t)box_b, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char b[160];
    snprintf(b, sizeof b, "cuTensorMapEncodeTiled failed with CUresult %d (pitch %d, box %dx%d)", (int)r, G.pitch,
             box_c, box_b);
    *err = b;
    return false;
  }
  return true;
}

constexpr int TY_U = 16, TY_P = 16;

int pick_chunk(const Geom &G, int planes, int ty) {
  // enough CTAs for >= ~24 waves of 148 SMs x 2 resident CTAs, chunks no shorter than 16 planes
  const long long tiles = (long long)((G.nC - M + TZ - 1) / TZ) * ((G.nB - 2 * M + ty - 1) / ty);
  const long long want = 148LL * 2 * 24;
  long long chunks = (want + tiles - 1) / tiles;
  int Lx = (int)((planes + chunks - 1) / chunks);
  if (Lx < 16) Lx = 16;
  if (Lx > planes) Lx = planes;
  return Lx;
}

}  // namespace

struct TiledPlan {
  CUtensorMap tm_p, tm_u, tm_v, tm_w;
  StencilTab *tab = nullptr;
};

bool tiled_supported(int ndim, const Geom &G) {
  return ndim == 3 && G.pitch % 32 == 0 && G.nB > 2 * M && G.nC > 2 * M && encode_fn() != nullptr;
}

TiledPlan *tiled_plan_create(const Fields &F, const Geom &G, const float *host_dmap, cudaStream_t st, std::string *err) {
  auto *pl = new TiledPlan();
  bool ok = make_tmap3d(&pl->tm_p, F.p, G, TZ + 2 * M, TY_U + 2 * M, err) &&
            make_tmap3d(&pl->tm_u, F.q[0], G, TZ + 8, TY_P + 2, err) &&
            make_tmap3d(&pl->tm_v, F.q[1], G, TZ + 8, TY_P + 2 * M, err) &&
            make_tmap3d(&pl->tm_w, F.q[2], G, TZ + 2 * M, TY_P + 2, err);
  if (!ok) { delete pl; return nullptr; }
  std::vector<StencilTab> h(G.ndmap);
  const int nd = G.ndmap;
  for (int c = 0; c < nd; ++c) {
    auto Dk = [&](int k) { return host_dmap[(size_t)(2 * k) * nd + c]; };
    h[c].d03 = make_float4(Dk(1), Dk(2), Dk(3), Dk(4));
    h[c].d47 = make_float4(Dk(5), Dk(6), Dk(7), Dk(8));
    h[c].e = make_float4(host_dmap[(size_t)3 * nd + c], 0.f, 0.f, 0.f);
  }
  if (cudaMalloc(&pl->tab, sizeof(StencilTab) * nd) != cudaSuccess ||
      cudaMemcpyAsync(pl->tab, h.data(), sizeof(StencilTab) * nd, cudaMemcpyHostToDevice, st) != cudaSuccess ||
      cudaStreamSynchronize(st) != cudaSuccess) {
    *err = std::string("tiled plan: ") + cudaGetErrorString(cudaGetLastError());
    if (pl->tab) cudaFree(pl->tab);
    delete pl;
    return nullptr;
  }
  return pl;
}

void tiled_plan_destroy(TiledPlan *pl) {
  if (!pl) return;
  if (pl->tab) cudaFree(pl->tab);
  delete pl;
}

int launch_sweep_u_tiled(const TiledPlan *pl, const Fields &F, const Geom &G, int a_lo, int a_hi, cudaStream_t st) {
  if (a_hi <= a_lo) return 0;
  const int Lx = pick_chunk(G, a_hi - a_lo, TY_U);
  dim3 blk(TZ, TY_U, 1);
  dim3 grd((G.nC - M + TZ - 1) / TZ, (G.nB - 2 * M + TY_U - 1) / TY_U, (a_hi - a_lo + Lx - 1) / Lx);
  k_sweep_u_tiled<TY_U, 2><<<grd, blk, 0, st>>>(pl->tm_p, F, G, pl->tab, a_lo, a_hi, Lx);
  return 1;
}

int launch_sweep_p_tiled(const TiledPlan *pl, const Fields &F, const Geom &G, int a_lo, int a_hi, cudaStream_t st) {
  if (a_hi <= a_lo) return 0;
  const int Lx = pick_chunk(G, a_hi - a_lo, TY_P);
  dim3 blk(TZ, TY_P, 1);
  dim3 grd((G.nC - M + TZ - 1) / TZ, (G.nB - 2 * M + TY_P - 1) / TY_P, (a_hi - a_lo + Lx - 1) / Lx);
  k_sweep_p_tiled<TY_P, 2><<<grd, blk, 0, st>>>(pl->tm_u, pl->tm_v, pl->tm_w, F, G, pl->tab, a_lo, a_hi, Lx);
  return 1;
}

}  // namespace fw25
