// fw25_sweeps_ws.cu -- 3D sweeps for sm_100a, warp-specialised: every global READ goes through TMA.
//
// Same 2.5-D scheme as fw25_sweeps_tiled.cu (march along x over (y,z) tiles of TY x 32 cells, x column of the
// stencil field in registers), but:
//   * a dedicated PRODUCER warp (one elected lane) issues all loads as cp.async.bulk.tensor.3d: the haloed
//     stencil tiles into a 4-slot ring AND the 16/18 point-wise arrays of the plane (centre tiles TY x 32) into
//     a 2-stage buffer, one plane ahead, completion on "full" mbarriers;
//   * TY CONSUMER warps (one per tile row, a warp = 32 contiguous z = one 128-byte line) read everything with
//     LDS at compile-time offsets, do the arithmetic, write the 9 (7) results with streaming stores, and hand
//     the buffers back through "empty" mbarriers (one arrival per warp).  No __syncthreads in the loop, no
//     long-scoreboard waits on global loads, no per-array address arithmetic for loads, no staging registers.
//
// Arithmetic: operation for operation the reference's (fw25_kernels.cuh); bit-identical to the other variants.
// Replaces fd_u / fd_p (3D PTX L38-675 / L677-1323).
#include <cuda.h>

#include <algorithm>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "fw25_internal.h"
#include "fw25_kernels.cuh"
#include "fw25_tma.cuh"

#ifndef FW25_WS_TY
#define FW25_WS_TY 16     // tile rows = consumer warps per CTA (sweep on a B200: profiles/sweep_ws_r01.txt)
#endif
#ifndef FW25_WS_MINB
#define FW25_WS_MINB 2    // resident CTAs per SM the register budget is sized for
#endif
#ifndef FW25_WS_UNROLL
#define FW25_WS_UNROLL 4  // x steps per unrolled loop body (ring slots become compile-time offsets at 4)
#endif
#define FW25_PRAGMA_(x) _Pragma(#x)
#define FW25_UNROLL_X(n) FW25_PRAGMA_(unroll n)

namespace fw25 {

namespace {

constexpr int TZ = 32;
constexpr int NH = 4;   // halo-tile ring slots: planes x, x+1 in use, two in flight
constexpr int NP = 2;   // point-wise stages: plane x in use, x+1 in flight

struct StencilTab {
  float4 d03, d47, e;
};

// tensor-map slots in the plan's device array
enum : int {
  // fd_u
  MU_PHALO = 0, MU_PNEW, MU_DC, MU_RHO, MU_K, MU_KX, MU_A1, MU_B1, MU_A2, MU_B2, MU_Q0, MU_Q1, MU_Q2,
  MU_M00, MU_M01, MU_M10, MU_M11, MU_M20, MU_M21, MU_END,
  // fd_p
  MP_UHALO = MU_END, MP_VHALO, MP_WHALO, MP_UNEW, MP_DC, MP_K, MP_BETA, MP_KU, MP_A1, MP_B1, MP_A2, MP_B2, MP_P,
  MP_F00, MP_F01, MP_F10, MP_F11, MP_F20, MP_F21, MP_END,
};
constexpr int NPW_U = MU_END - MU_PNEW;   // 18 point-wise tiles per plane in fd_u
constexpr int NPW_P = MP_END - MP_UNEW;   // 16 in fd_p
// Anisotropic-relaxation family (one kappa / a / b map per axis, input_file_writer.py:592-620): the per-sweep slots
// above hold axis A (x); axes B and C add 5 point-wise tiles each (kappa, a1, b1, a2, b2) behind the isotropic ones.
constexpr int NX_ANISO = 10;
constexpr int MXU = MP_END, MXP = MXU + NX_ANISO, M_END_ANISO = MXP + NX_ANISO;   // tensor-map slots of the extras
enum : int { X_KAPPA = 0, X_A1, X_B1, X_A2, X_B2 };
__host__ __device__ constexpr int npw_u(bool aniso) { return NPW_U + (aniso ? NX_ANISO : 0); }
__host__ __device__ constexpr int npw_p(bool aniso) { return NPW_P + (aniso ? NX_ANISO : 0); }
__host__ __device__ constexpr int xslot(int npw_iso, int axis, int which) { return npw_iso + (axis - 1) * 5 + which; }   // axis 1 (B), 2 (C)

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}

// Tile order.  CTAs are handed out in blockIdx.x-fastest order and about 2 x 148 of them are resident: a CTA starts
// when the one ~296 launches before it retires, so two tiles that are d launches apart run ~ d / 296 chunks out of
// phase.  With z fastest over a full row of tiles (39 at nZ = 1240), y-neighbours are 39 launches = 4 planes apart and
// the y halo rows one of them brought into L2 are gone (~13 MB stream through L2 per plane step) by the time the other
// asks for them.  Strips of `strip` z-tiles (z fastest inside the strip, then y, then the next strip) put y-neighbours
// `strip` launches apart; only the strip edges (1 in `strip` z-halos) are out of phase.
__device__ __forceinline__ void tile_of_block(int strip, int &bx, int &by) {
  bx = blockIdx.x; by = blockIdx.y;
  if (strip <= 0 || strip >= (int)gridDim.x) return;
  const int L = blockIdx.x + gridDim.x * blockIdx.y;
  const int per = strip * gridDim.y;
  const int s = L / per;
  const int w = min(strip, (int)gridDim.x - s * strip);     // the last strip may be narrower
  const int r = L - s * per;
  by = r / w;
  bx = s * strip + r - by * w;
}

template <int TY, int NPW>
struct alignas(128) SmemU {
  float halo[NH][TY + 2 * M][TZ + 2 * M];
  float pw[NP][NPW][TY][TZ];
  uint64_t full_h[NH], empty_h[NH], full_p[NP], empty_p[NP];
};

template <int TY, int NPW>
struct alignas(128) SmemP {
  struct alignas(128) Halo {
    alignas(128) float u[TY + 2][TZ + 8];
    alignas(128) float v[TY + 2 * M][TZ + 8];
    alignas(128) float w[TY + 2][TZ + 2 * M];
  } halo[NH];
  float pw[NP][NPW][TY][TZ];
  uint64_t full_h[NH], empty_h[NH], full_p[NP], empty_p[NP];
};

// ------------------------------------------------------------------------------------------ fd_u
// PUSH: the launch covers the planes next to an x-slab interface and also stores its results into the neighbour
// GPU's ghost planes through a peer-mapped pointer (NVLink): the halo exchange is part of the sweep, tile by
// tile, instead of a copy after it (u: every plane of the launch; v, w: only the plane the cross terms read).
template <int TY, int MINB, bool PUSH, bool ANISO>
__global__ void __launch_bounds__((TY + 1) * TZ, MINB)
    k_sweep_u_ws(const CUtensorMap *__restrict__ maps, const Fields F, const Geom G,
                 const StencilTab *__restrict__ tab, int a_lo, int a_hi, int Lx, int hint, const HaloPush push) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int NPW = npw_u(ANISO);
  SmemU<TY, NPW> &S = *reinterpret_cast<SmemU<TY, NPW> *>(smem_raw);
  constexpr int HY = TY + 2 * M, HZ = TZ + 2 * M;
  constexpr uint32_t HALO_BYTES = HY * HZ * 4, PW_BYTES = NPW * TY * TZ * 4;

  const int tz = threadIdx.x, ty = threadIdx.y;
  int bx, by;
  tile_of_block(hint >> 8, bx, by);
  hint &= 0xff;
  const int z0 = bx * TZ, y0 = M + by * TY;
  const int xa = a_lo + blockIdx.z * Lx;
  const int xb = min(xa + Lx, a_hi);
  const int len = xb - xa;

  if (ty == TY && tz == 0) {
#pragma unroll
    for (int s = 0; s < NH; ++s) { mbar_init(&S.full_h[s], 1); mbar_init(&S.empty_h[s], TY); }
#pragma unroll
    for (int s = 0; s < NP; ++s) { mbar_init(&S.full_p[s], 1); mbar_init(&S.empty_p[s], TY); }
    mbar_fence_init();
  }
  __syncthreads();

  if (ty == TY) {
    // ---------------- producer warp: one lane issues every load of the chunk
    if (tz == 0) {
      const uint64_t pol_stream = l2_policy_evict_first();
      const uint64_t pol_keep = l2_policy_evict_last();
      // halo tiles of planes xa .. xb (ring index m), point-wise tiles of xa .. xb-1 one plane behind them, so
      // that waiting for a point-wise stage to drain never delays the next halo tile
      for (int m = 0; m <= len + 1; ++m) {
        if (m <= len) {
          const int s = m & (NH - 1), k = m / NH;
          if (k > 0) mbar_wait(&S.empty_h[s], (k - 1) & 1);
          mbar_arrive_expect_tx(&S.full_h[s], HALO_BYTES);
          if (hint & 2) tma_load_3d_hint(&S.halo[s][0][0], &maps[MU_PHALO], z0 - M, y0 - M, xa + m, &S.full_h[s], pol_keep);
          else tma_load_3d(&S.halo[s][0][0], &maps[MU_PHALO], z0 - M, y0 - M, xa + m, &S.full_h[s]);
        }
        const int n = m - 1;
        if (n >= 0 && n < len) {
          const int P = xa + n, s = n & (NP - 1), k = n / NP;
          if (k > 0) mbar_wait(&S.empty_p[s], (k - 1) & 1);
          mbar_arrive_expect_tx(&S.full_p[s], PW_BYTES);
          // p[x+8] is this plane's first touch from DRAM; the haloed tile of the same plane follows 8 steps later
          if (hint & 2) tma_load_3d_hint(&S.pw[s][0][0][0], &maps[MU_PNEW], z0, y0, P + M, &S.full_p[s], pol_keep);
          else tma_load_3d(&S.pw[s][0][0][0], &maps[MU_PNEW], z0, y0, P + M, &S.full_p[s]);
          if (hint & 1) {
#pragma unroll
            for (int a = 1; a < NPW; ++a)
              tma_load_3d_hint(&S.pw[s][a][0][0], &maps[a < NPW_U ? MU_PNEW + a : MXU + a - NPW_U], z0, y0, P,
                               &S.full_p[s], pol_stream);
          } else {
#pragma unroll
            for (int a = 1; a < NPW; ++a)
              tma_load_3d(&S.pw[s][a][0][0], &maps[a < NPW_U ? MU_PNEW + a : MXU + a - NPW_U], z0, y0, P, &S.full_p[s]);
          }
        }
      }
    }
    return;
  }

  // ---------------- consumer warps
  const int z = z0 + tz, y = y0 + ty;
  const bool act = (z >= M) && (z < G.nC - M) && (y < G.nB - M);
  const int sy = ty + M, sz = tz + M;
  const unsigned sA = (unsigned)G.sA, sB = (unsigned)G.sB;
  unsigned gi = (unsigned)xa * sA + (unsigned)y * sB + (unsigned)z;   // element index (< 2^32, checked on the host)
  const float *__restrict__ p = F.p;

  float pc[16];
  float pm_y1 = 0.f, pm_z1 = 0.f;
  if (act) {
#pragma unroll
    for (int j = 0; j < 15; ++j) pc[j] = p[gi + (unsigned)(j - 7) * sA];
    pm_y1 = p[gi - sA + sB];
    pm_z1 = p[gi - sA + 1];
  } else {
#pragma unroll
    for (int j = 0; j < 15; ++j) pc[j] = 0.f;
  }
  pc[15] = 0.f;

  mbar_wait(&S.full_h[0], 0);

FW25_UNROLL_X(FW25_WS_UNROLL)
  for (int n = 0; n < len; ++n, gi += sA) {
    mbar_wait(&S.full_h[(n + 1) & (NH - 1)], ((n + 1) / NH) & 1);
    mbar_wait(&S.full_p[n & (NP - 1)], (n / NP) & 1);

    if (act) {
      const float(*c0)[HZ] = S.halo[n & (NH - 1)];
      const float(*c1)[HZ] = S.halo[(n + 1) & (NH - 1)];
      const float(*W)[TY][TZ] = S.pw[n & (NP - 1)];
      pc[15] = W[0][ty][tz];
      const int ci = __float_as_int(W[MU_DC - MU_PNEW][ty][tz]);
      const StencilTab T = tab[ci];
      const float D[9] = {0.f, T.d03.x, T.d03.y, T.d03.z, T.d03.w, T.d47.x, T.d47.y, T.d47.z, T.d47.w};
      const float E = T.e.x;
      const float pcen = pc[7];

      float yv[16], zv[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        yv[j] = (j == 7) ? pcen : c0[sy + j - 7][sz];
        zv[j] = (j == 7) ? pcen : c0[sy][sz + j - 7];
      }
      float gA = 0.f, gB = 0.f, gC = 0.f;
#pragma unroll
      for (int k = 1; k <= M; ++k) {
        gA = fma_(D[k], sub_(pc[7 + k], pc[8 - k]), gA);
        gB = fma_(D[k], sub_(yv[7 + k], yv[8 - k]), gB);
        gC = fma_(D[k], sub_(zv[7 + k], zv[8 - k]), gC);
      }
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

      const float rho = W[MU_RHO - MU_PNEW][ty][tz], Kc = W[MU_K - MU_PNEW][ty][tz], kx = W[MU_KX - MU_PNEW][ty][tz];
      const float a1 = W[MU_A1 - MU_PNEW][ty][tz], b1 = W[MU_B1 - MU_PNEW][ty][tz];
      const float a2 = W[MU_A2 - MU_PNEW][ty][tz], b2 = W[MU_B2 - MU_PNEW][ty][tz];
      const float s = div_(div_(G.dT, rho), fma_(rcp_(Kc), pcen, 1.0f));
      float kxB = kx, a1B = a1, b1B = b1, a2B = a2, b2B = b2, kxC = kx, a1C = a1, b1C = b1, a2C = a2, b2C = b2;
      if constexpr (ANISO) {
        kxB = W[xslot(NPW_U, 1, X_KAPPA)][ty][tz]; kxC = W[xslot(NPW_U, 2, X_KAPPA)][ty][tz];
        a1B = W[xslot(NPW_U, 1, X_A1)][ty][tz]; b1B = W[xslot(NPW_U, 1, X_B1)][ty][tz];
        a2B = W[xslot(NPW_U, 1, X_A2)][ty][tz]; b2B = W[xslot(NPW_U, 1, X_B2)][ty][tz];
        a1C = W[xslot(NPW_U, 2, X_A1)][ty][tz]; b1C = W[xslot(NPW_U, 2, X_B1)][ty][tz];
        a2C = W[xslot(NPW_U, 2, X_A2)][ty][tz]; b2C = W[xslot(NPW_U, 2, X_B2)][ty][tz];
      }
      const float m00 = fma_(b1, W[MU_M00 - MU_PNEW][ty][tz], mul_(gA, a1));
      const float m01 = fma_(b2, W[MU_M01 - MU_PNEW][ty][tz], mul_(gA, a2));
      const float m10 = fma_(b1B, W[MU_M10 - MU_PNEW][ty][tz], mul_(gB, a1B));
      const float m11 = fma_(b2B, W[MU_M11 - MU_PNEW][ty][tz], mul_(gB, a2B));
      const float m20 = fma_(b1C, W[MU_M20 - MU_PNEW][ty][tz], mul_(gC, a1C));
      const float m21 = fma_(b2C, W[MU_M21 - MU_PNEW][ty][tz], mul_(gC, a2C));
      const float q0 = fma_(-s, add_(add_(div_(gA, kx), m00), m01), W[MU_Q0 - MU_PNEW][ty][tz]);
      const float q1 = fma_(-s, add_(add_(div_(gB, kxB), m10), m11), W[MU_Q1 - MU_PNEW][ty][tz]);
      const float q2 = fma_(-s, add_(add_(div_(gC, kxC), m20), m21), W[MU_Q2 - MU_PNEW][ty][tz]);
      __stcs(F.psi[0][0] + gi, m00); __stcs(F.psi[0][1] + gi, m01);
      __stcs(F.psi[1][0] + gi, m10); __stcs(F.psi[1][1] + gi, m11);
      __stcs(F.psi[2][0] + gi, m20); __stcs(F.psi[2][1] + gi, m21);
      __stcs(F.q[0] + gi, q0); __stcs(F.q[1] + gi, q1); __stcs(F.q[2] + gi, q2);
      if constexpr (PUSH) {
        if (xa + n >= push.lo0 && xa + n < push.hi0) push.a[0][gi] = q0;
        if (xa + n >= push.lo1 && xa + n < push.hi1) { push.a[1][gi] = q1; push.a[2][gi] = q2; }
      }
#pragma unroll
      for (int j = 0; j < 15; ++j) pc[j] = pc[j + 1];
    }
    __syncwarp();
    if (tz == 0) {
      mbar_arrive(&S.empty_h[n & (NH - 1)]);
      mbar_arrive(&S.empty_p[n & (NP - 1)]);
    }
  }
}

// ------------------------------------------------------------------------------------------ fd_p
template <int TY, int MINB, bool PUSH, bool ANISO>
__global__ void __launch_bounds__((TY + 1) * TZ, MINB)
    k_sweep_p_ws(const CUtensorMap *__restrict__ maps, const Fields F, const Geom G,
                 const StencilTab *__restrict__ tab, int a_lo, int a_hi, int Lx, int hint, const HaloPush push) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int NPW = npw_p(ANISO);
  SmemP<TY, NPW> &S = *reinterpret_cast<SmemP<TY, NPW> *>(smem_raw);
  constexpr uint32_t U_BYTES = (TY + 2) * (TZ + 8) * 4, V_BYTES = (TY + 2 * M) * (TZ + 8) * 4,
                     W_BYTES = (TY + 2) * (TZ + 2 * M) * 4, PW_BYTES = NPW * TY * TZ * 4;

  const int tz = threadIdx.x, ty = threadIdx.y;
  int bx, by;
  tile_of_block(hint >> 8, bx, by);
  hint &= 0xff;
  const int z0 = bx * TZ, y0 = M + by * TY;
  const int xa = a_lo + blockIdx.z * Lx;
  const int xb = min(xa + Lx, a_hi);
  const int len = xb - xa;

  if (ty == TY && tz == 0) {
#pragma unroll
    for (int s = 0; s < NH; ++s) { mbar_init(&S.full_h[s], 1); mbar_init(&S.empty_h[s], TY); }
#pragma unroll
    for (int s = 0; s < NP; ++s) { mbar_init(&S.full_p[s], 1); mbar_init(&S.empty_p[s], TY); }
    mbar_fence_init();
  }
  __syncthreads();

  if (ty == TY) {
    // producer: halo tiles of planes xa-1 .. xb (ring index m = plane - (xa-1)), point-wise tiles of xa .. xb-1
    if (tz == 0) {
      const uint64_t pol_stream = l2_policy_evict_first();
      const uint64_t pol_keep = l2_policy_evict_last();
      for (int m = 0; m <= len + 1; ++m) {
        {
          const int P = xa - 1 + m, s = m & (NH - 1), k = m / NH;
          if (k > 0) mbar_wait(&S.empty_h[s], (k - 1) & 1);
          mbar_arrive_expect_tx(&S.full_h[s], U_BYTES + V_BYTES + W_BYTES);
          if (hint & 2) {
            tma_load_3d_hint(&S.halo[s].u[0][0], &maps[MP_UHALO], z0 - 4, y0 - 1, P, &S.full_h[s], pol_keep);
            tma_load_3d_hint(&S.halo[s].v[0][0], &maps[MP_VHALO], z0 - 4, y0 - M, P, &S.full_h[s], pol_keep);
            tma_load_3d_hint(&S.halo[s].w[0][0], &maps[MP_WHALO], z0 - M, y0 - 1, P, &S.full_h[s], pol_keep);
          } else {
            tma_load_3d(&S.halo[s].u[0][0], &maps[MP_UHALO], z0 - 4, y0 - 1, P, &S.full_h[s]);
            tma_load_3d(&S.halo[s].v[0][0], &maps[MP_VHALO], z0 - 4, y0 - M, P, &S.full_h[s]);
            tma_load_3d(&S.halo[s].w[0][0], &maps[MP_WHALO], z0 - M, y0 - 1, P, &S.full_h[s]);
          }
        }
        const int n = m - 1;           // point-wise plane xa + n goes out right after halo plane xa + n
        if (n >= 0 && n < len) {
          const int P = xa + n, s = n & (NP - 1), k = n / NP;
          if (k > 0) mbar_wait(&S.empty_p[s], (k - 1) & 1);
          mbar_arrive_expect_tx(&S.full_p[s], PW_BYTES);
          if (hint & 2) tma_load_3d_hint(&S.pw[s][0][0][0], &maps[MP_UNEW], z0, y0, P + M - 1, &S.full_p[s], pol_keep);
          else tma_load_3d(&S.pw[s][0][0][0], &maps[MP_UNEW], z0, y0, P + M - 1, &S.full_p[s]);   // u[x+7]
          if (hint & 1) {
#pragma unroll
            for (int a = 1; a < NPW; ++a)
              tma_load_3d_hint(&S.pw[s][a][0][0], &maps[a < NPW_P ? MP_UNEW + a : MXP + a - NPW_P], z0, y0, P,
                               &S.full_p[s], pol_stream);
          } else {
#pragma unroll
            for (int a = 1; a < NPW; ++a)
              tma_load_3d(&S.pw[s][a][0][0], &maps[a < NPW_P ? MP_UNEW + a : MXP + a - NPW_P], z0, y0, P, &S.full_p[s]);
          }
        }
      }
    }
    return;
  }

  const int z = z0 + tz, y = y0 + ty;
  const bool act = (z >= M) && (z < G.nC - M) && (y < G.nB - M);
  const unsigned sA = (unsigned)G.sA, sB = (unsigned)G.sB;
  unsigned gi = (unsigned)xa * sA + (unsigned)y * sB + (unsigned)z;
  const float *__restrict__ u = F.q[0];

  float uc[16];   // uc[j] = u[x - 8 + j]
  if (act) {
#pragma unroll
    for (int j = 0; j < 15; ++j) uc[j] = u[gi + (unsigned)(j - 8) * sA];
  } else {
#pragma unroll
    for (int j = 0; j < 15; ++j) uc[j] = 0.f;
  }
  uc[15] = 0.f;

  mbar_wait(&S.full_h[0], 0);
  mbar_wait(&S.full_h[1], 0);

FW25_UNROLL_X(FW25_WS_UNROLL)
  for (int n = 0; n < len; ++n, gi += sA) {
    // ring index of plane x-1 is n, of x is n+1, of x+1 is n+2
    mbar_wait(&S.full_h[(n + 2) & (NH - 1)], ((n + 2) / NH) & 1);
    mbar_wait(&S.full_p[n & (NP - 1)], (n / NP) & 1);

    if (act) {
      const typename SmemP<TY, NPW>::Halo &Sm = S.halo[n & (NH - 1)];
      const typename SmemP<TY, NPW>::Halo &S0 = S.halo[(n + 1) & (NH - 1)];
      const typename SmemP<TY, NPW>::Halo &S1 = S.halo[(n + 2) & (NH - 1)];
      const float(*W)[TY][TZ] = S.pw[n & (NP - 1)];
      uc[15] = W[0][ty][tz];
      const int ci = __float_as_int(W[MP_DC - MP_UNEW][ty][tz]);
      const StencilTab T = tab[ci];
      const float D[9] = {0.f, T.d03.x, T.d03.y, T.d03.z, T.d03.w, T.d47.x, T.d47.y, T.d47.z, T.d47.w};
      const float E = T.e.x;
      const int uy = ty + 1, uz = tz + 4;
      const int vy = ty + M, vz = tz + 4;
      const int wy = ty + 1, wz = tz + M;

      float hA = 0.f, hB = 0.f, hC = 0.f;
#pragma unroll
      for (int k = 1; k <= M; ++k) {   // PTX L799-966: ascending k
        hA = fma_(D[k], sub_(uc[7 + k], uc[8 - k]), hA);
        hB = fma_(D[k], sub_(S0.v[vy + k - 1][vz], S0.v[vy - k][vz]), hB);
        hC = fma_(D[k], sub_(S0.w[wy][wz + k - 1], S0.w[wy][wz - k]), hC);
      }
      float cA = sub_(S0.u[uy + 1][uz], Sm.u[uy + 1][uz]);
      cA = add_(cA, S0.u[uy - 1][uz]); cA = sub_(cA, Sm.u[uy - 1][uz]);
      cA = add_(cA, S0.u[uy][uz + 1]); cA = sub_(cA, Sm.u[uy][uz + 1]);
      cA = add_(cA, S0.u[uy][uz - 1]); cA = sub_(cA, Sm.u[uy][uz - 1]);
      float cB = sub_(S1.v[vy][vz], S1.v[vy - 1][vz]);
      cB = add_(cB, Sm.v[vy][vz]); cB = sub_(cB, Sm.v[vy - 1][vz]);
      cB = add_(cB, S0.v[vy][vz + 1]); cB = sub_(cB, S0.v[vy - 1][vz + 1]);
      cB = add_(cB, S0.v[vy][vz - 1]); cB = sub_(cB, S0.v[vy - 1][vz - 1]);
      float cC = sub_(S1.w[wy][wz], S1.w[wy][wz - 1]);
      cC = add_(cC, Sm.w[wy][wz]); cC = sub_(cC, Sm.w[wy][wz - 1]);
      cC = add_(cC, S0.w[wy + 1][wz]); cC = sub_(cC, S0.w[wy + 1][wz - 1]);
      cC = add_(cC, S0.w[wy - 1][wz]); cC = sub_(cC, S0.w[wy - 1][wz - 1]);

      const float dX = G.dX;
      hA = div_(fma_(E, cA, hA), dX);
      hB = div_(fma_(E, cB, hB), dX);
      hC = div_(fma_(E, cC, hC), dX);

      const float Kc = W[MP_K - MP_UNEW][ty][tz], bt = W[MP_BETA - MP_UNEW][ty][tz], ku = W[MP_KU - MP_UNEW][ty][tz];
      const float a1 = W[MP_A1 - MP_UNEW][ty][tz], b1 = W[MP_B1 - MP_UNEW][ty][tz];
      const float a2 = W[MP_A2 - MP_UNEW][ty][tz], b2 = W[MP_B2 - MP_UNEW][ty][tz];
      const float pc = W[MP_P - MP_UNEW][ty][tz];
      float kuB = ku, a1B = a1, b1B = b1, a2B = a2, b2B = b2, kuC = ku, a1C = a1, b1C = b1, a2C = a2, b2C = b2;
      if constexpr (ANISO) {
        kuB = W[xslot(NPW_P, 1, X_KAPPA)][ty][tz]; kuC = W[xslot(NPW_P, 2, X_KAPPA)][ty][tz];
        a1B = W[xslot(NPW_P, 1, X_A1)][ty][tz]; b1B = W[xslot(NPW_P, 1, X_B1)][ty][tz];
        a2B = W[xslot(NPW_P, 1, X_A2)][ty][tz]; b2B = W[xslot(NPW_P, 1, X_B2)][ty][tz];
        a1C = W[xslot(NPW_P, 2, X_A1)][ty][tz]; b1C = W[xslot(NPW_P, 2, X_B1)][ty][tz];
        a2C = W[xslot(NPW_P, 2, X_A2)][ty][tz]; b2C = W[xslot(NPW_P, 2, X_B2)][ty][tz];
      }
      const float f00 = fma_(b1, W[MP_F00 - MP_UNEW][ty][tz], mul_(hA, a1));
      const float f01 = fma_(b2, W[MP_F01 - MP_UNEW][ty][tz], mul_(hA, a2));
      const float f10 = fma_(b1B, W[MP_F10 - MP_UNEW][ty][tz], mul_(hB, a1B));
      const float f11 = fma_(b2B, W[MP_F11 - MP_UNEW][ty][tz], mul_(hB, a2B));
      const float f20 = fma_(b1C, W[MP_F20 - MP_UNEW][ty][tz], mul_(hC, a1C));
      const float f21 = fma_(b2C, W[MP_F21 - MP_UNEW][ty][tz], mul_(hC, a2C));
      float Ssum = add_(div_(hA, ku), div_(hB, kuB));      // PTX L1290-1305
      Ssum = add_(div_(hC, kuC), Ssum);
      Ssum = add_(f00, Ssum); Ssum = add_(f01, Ssum); Ssum = add_(f10, Ssum); Ssum = add_(f11, Ssum);
      Ssum = add_(f20, Ssum); Ssum = add_(f21, Ssum);
      const float At = mul_(mul_(G.dT, Kc), Ssum);
      const float Bt = fma_(pc, mul_(rcp_(Kc), sub_(1.0f, add_(bt, bt))), 1.0f);
      __stcs(F.phi[0][0] + gi, f00); __stcs(F.phi[0][1] + gi, f01);
      __stcs(F.phi[1][0] + gi, f10); __stcs(F.phi[1][1] + gi, f11);
      __stcs(F.phi[2][0] + gi, f20); __stcs(F.phi[2][1] + gi, f21);
      const float pn = fma_(-At, Bt, pc);
      F.p[gi] = pn;                  // p is the next sweep's stencil field: default caching
      if constexpr (PUSH) {
        if (xa + n >= push.lo0 && xa + n < push.hi0) push.a[0][gi] = pn;
      }
#pragma unroll
      for (int j = 0; j < 15; ++j) uc[j] = uc[j + 1];
    }
    __syncwarp();
    if (tz == 0) {
      mbar_arrive(&S.empty_h[n & (NH - 1)]);     // plane x-1 is done
      mbar_arrive(&S.empty_p[n & (NP - 1)]);
    }
  }
}

// ------------------------------------------------------------------------------------------ host
using EncodeFn = CUresult (*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                              const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                              CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn encode_fn_ws() {
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

bool tmap3d(CUtensorMap *m, const void *base, const Geom &G, int box_c, int box_b, std::string *err) {
  EncodeFn fn = encode_fn_ws();
  if (!fn) { *err = "cuTensorMapEncodeTiled is not available from this driver"; return false; }
  const cuuint64_t dims[3] = {(cuuint64_t)G.pitch, (cuuint64_t)G.nB, (cuuint64_t)G.nA};
  const cuuint64_t strides[2] = {(cuuint64_t)G.sB * 4, (cuuint64_t)G.sA * 4};
  const cuuint32_t box[3] = {(cuuint32_t)box_c, (cuuint32_t)box_b, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void *>(base), dims, strides, box, estr,
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

// tile rows / resident CTAs per SM, per sweep (tuning: FW25_WS_TY_U=.. FW25_WS_MINB_U=.. python -m fullwave25_b200.build)
#ifndef FW25_WS_TY_U
#define FW25_WS_TY_U FW25_WS_TY
#endif
#ifndef FW25_WS_TY_P
#define FW25_WS_TY_P FW25_WS_TY
#endif
#ifndef FW25_WS_MINB_U
#define FW25_WS_MINB_U FW25_WS_MINB
#endif
#ifndef FW25_WS_MINB_P
#define FW25_WS_MINB_P FW25_WS_MINB
#endif
constexpr int TY_U = FW25_WS_TY_U, TY_P = FW25_WS_TY_P;
constexpr int MINB_U = FW25_WS_MINB_U, MINB_P = FW25_WS_MINB_P;
constexpr int MINB_WS = FW25_WS_MINB;     // anisotropic family
// Anisotropic family: 28 (fd_u) / 26 (fd_p) point-wise tiles per plane.  Tile rows chosen so that two CTAs still fit
// one SM's 227 KB of shared memory: fd_u 12 rows (107.6 KB), fd_p 10 rows (100.1 KB).
#ifndef FW25_WS_TY_AU
#define FW25_WS_TY_AU 12
#endif
#ifndef FW25_WS_TY_AP
#define FW25_WS_TY_AP 10
#endif
constexpr int TY_AU = FW25_WS_TY_AU, TY_AP = FW25_WS_TY_AP;
static_assert(sizeof(SmemU<TY_AU, npw_u(true)>) * 2 + 2048 <= 227 * 1024, "anisotropic fd_u: two CTAs per SM");
static_assert(sizeof(SmemP<TY_AP, npw_p(true)>) * 2 + 2048 <= 227 * 1024, "anisotropic fd_p: two CTAs per SM");

// bit 0: point-wise tiles are loaded with an L2 evict-first policy.  Helps long chunks, costs ~2 % at the default chunk
//        length (profiles/sweep_lx_r01.txt): off.
// bit 1: the stencil field's tiles (p in fd_u; u, v, w in fd_p) are loaded with an L2 evict-last policy: every plane of
//        the field is fetched twice, 7-8 plane steps apart (as the leading value of the register column, then as the
//        haloed tile), and ~100 MB of point-wise tiles stream through L2 in between.  Measured at 800 x 1240 x 1240:
//        27.66 -> 28.03 Gpt/s, DRAM bytes per updated point 113.5 -> 112.2 (fd_u), 107.1 -> 105.8 (fd_p)
//        (profiles/ncu_r02_ws.txt): on by default.
// (tried: cp.async.bulk.prefetch.tensor.L2 of the point-wise tiles 1-5 planes ahead: 27.6 -> 25.2 / 23.9 / 21.8 / 19.5
//  Gpt/s, monotonically worse -- profiles/README.md; not kept)
// bits 8..: strip width of the tile order (tile_of_block); 0 = plain z-fastest order
int ws_hint() {
  static const int h = [] {
    const char *e = getenv("FW25_WS_HINT");
    const char *s = getenv("FW25_WS_STRIP");
    const int strip = s ? atoi(s) : 0;      // measured: strips lose 4-10 % at 800 x 1240 x 1240 (profiles/README.md)
    return (e ? atoi(e) & 0xff : 2) | (std::max(strip, 0) << 8);
  }();
  return h;
}

int pick_chunk_ws(const Geom &G, int planes) {
  // Short chunks win on a B200 (profiles/sweep_lx_r01.txt: 88 % of the HBM roofline at 24-32 planes, 76 % at
  // 262): CTAs that start together stay within a few planes of each other, so neighbouring tiles find each
  // other's halo rows -- and the plane a column load touched seven steps earlier -- still in L2.  The
  // prologue (15 column loads + pipeline fill) costs < 2 % at 32.
  (void)G;
  static const int lx_env = [] { const char *e = getenv("FW25_WS_LX"); return e ? atoi(e) : 0; }();
  const int target = lx_env > 0 ? lx_env : 32;
  int chunks = (planes + target - 1) / target;
  // short launches (the interior of a 100-plane slab is 36 planes): one 36-plane chunk instead of two of 18, whose 15
  // prologue planes each would nearly double the column loads
  if (lx_env <= 0 && chunks > 1 && planes / chunks < 24) --chunks;
  return (planes + chunks - 1) / chunks;
}

}  // namespace

struct WsPlan {
  CUtensorMap *maps = nullptr;   // device array [MP_END], anisotropic: [M_END_ANISO]
  StencilTab *tab = nullptr;
  bool aniso = false;
};

bool ws_supported(int ndim, const Geom &G) {
  const unsigned long long cells = (unsigned long long)G.nA * G.nB * G.pitch;
  return ndim == 3 && G.pitch % 32 == 0 && G.nB > 2 * M && G.nC > 2 * M && cells < (1ull << 32) &&
         encode_fn_ws() != nullptr;
}

namespace {
template <class K>
bool allow_smem(K kernel, size_t bytes) {
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) == cudaSuccess;
}
}  // namespace

WsPlan *ws_plan_create(const Fields &F, const Geom &G, const float *host_dmap, cudaStream_t st, std::string *err,
                       bool aniso) {
  const int n_maps = aniso ? M_END_ANISO : MP_END;
  const int ty_u = aniso ? TY_AU : TY_U, ty_p = aniso ? TY_AP : TY_P;
  std::vector<CUtensorMap> h(n_maps);
  bool ok = true;
  auto centre_u = [&](int slot, const void *base) { ok = ok && tmap3d(&h[slot], base, G, TZ, ty_u, err); };
  auto centre_p = [&](int slot, const void *base) { ok = ok && tmap3d(&h[slot], base, G, TZ, ty_p, err); };
  ok = ok && tmap3d(&h[MU_PHALO], F.p, G, TZ + 2 * M, ty_u + 2 * M, err);
  centre_u(MU_PNEW, F.p); centre_u(MU_DC, F.dcmap); centre_u(MU_RHO, F.rho); centre_u(MU_K, F.K);
  centre_u(MU_KX, F.kv[0]); centre_u(MU_A1, F.av[0][0]); centre_u(MU_B1, F.bv[0][0]);
  centre_u(MU_A2, F.av[0][1]); centre_u(MU_B2, F.bv[0][1]);
  centre_u(MU_Q0, F.q[0]); centre_u(MU_Q1, F.q[1]); centre_u(MU_Q2, F.q[2]);
  centre_u(MU_M00, F.psi[0][0]); centre_u(MU_M01, F.psi[0][1]); centre_u(MU_M10, F.psi[1][0]);
  centre_u(MU_M11, F.psi[1][1]); centre_u(MU_M20, F.psi[2][0]); centre_u(MU_M21, F.psi[2][1]);
  ok = ok && tmap3d(&h[MP_UHALO], F.q[0], G, TZ + 8, ty_p + 2, err) &&
       tmap3d(&h[MP_VHALO], F.q[1], G, TZ + 8, ty_p + 2 * M, err) &&
       tmap3d(&h[MP_WHALO], F.q[2], G, TZ + 2 * M, ty_p + 2, err);
  centre_p(MP_UNEW, F.q[0]); centre_p(MP_DC, F.dcmap); centre_p(MP_K, F.K); centre_p(MP_BETA, F.beta);
  centre_p(MP_KU, F.kp[0]); centre_p(MP_A1, F.ap[0][0]); centre_p(MP_B1, F.bp[0][0]);
  centre_p(MP_A2, F.ap[0][1]); centre_p(MP_B2, F.bp[0][1]); centre_p(MP_P, F.p);
  centre_p(MP_F00, F.phi[0][0]); centre_p(MP_F01, F.phi[0][1]); centre_p(MP_F10, F.phi[1][0]);
  centre_p(MP_F11, F.phi[1][1]); centre_p(MP_F20, F.phi[2][0]); centre_p(MP_F21, F.phi[2][1]);
  if (aniso) {
    for (int ax = 1; ax <= 2; ++ax) {
      const int u0 = MXU + (ax - 1) * 5, p0 = MXP + (ax - 1) * 5;
      centre_u(u0 + X_KAPPA, F.kv[ax]);
      centre_u(u0 + X_A1, F.av[ax][0]); centre_u(u0 + X_B1, F.bv[ax][0]);
      centre_u(u0 + X_A2, F.av[ax][1]); centre_u(u0 + X_B2, F.bv[ax][1]);
      centre_p(p0 + X_KAPPA, F.kp[ax]);
      centre_p(p0 + X_A1, F.ap[ax][0]); centre_p(p0 + X_B1, F.bp[ax][0]);
      centre_p(p0 + X_A2, F.ap[ax][1]); centre_p(p0 + X_B2, F.bp[ax][1]);
    }
  }
  if (!ok) return nullptr;

  auto *pl = new WsPlan();
  pl->aniso = aniso;
  std::vector<StencilTab> t(G.ndmap);
  const int nd = G.ndmap;
  for (int c = 0; c < nd; ++c) {
    auto Dk = [&](int k) { return host_dmap[(size_t)(2 * k) * nd + c]; };
    t[c].d03 = make_float4(Dk(1), Dk(2), Dk(3), Dk(4));
    t[c].d47 = make_float4(Dk(5), Dk(6), Dk(7), Dk(8));
    t[c].e = make_float4(host_dmap[(size_t)3 * nd + c], 0.f, 0.f, 0.f);
  }
  cudaError_t e1 = cudaMalloc(&pl->maps, sizeof(CUtensorMap) * n_maps);
  cudaError_t e2 = cudaMalloc(&pl->tab, sizeof(StencilTab) * nd);
  bool attr;
  if (aniso) {
    constexpr size_t su = sizeof(SmemU<TY_AU, npw_u(true)>), sp = sizeof(SmemP<TY_AP, npw_p(true)>);
    attr = allow_smem(k_sweep_u_ws<TY_AU, MINB_WS, false, true>, su) && allow_smem(k_sweep_u_ws<TY_AU, MINB_WS, true, true>, su) &&
           allow_smem(k_sweep_p_ws<TY_AP, MINB_WS, false, true>, sp) && allow_smem(k_sweep_p_ws<TY_AP, MINB_WS, true, true>, sp);
  } else {
    constexpr size_t su = sizeof(SmemU<TY_U, NPW_U>), sp = sizeof(SmemP<TY_P, NPW_P>);
    attr = allow_smem(k_sweep_u_ws<TY_U, MINB_U, false, false>, su) && allow_smem(k_sweep_u_ws<TY_U, MINB_U, true, false>, su) &&
           allow_smem(k_sweep_p_ws<TY_P, MINB_P, false, false>, sp) && allow_smem(k_sweep_p_ws<TY_P, MINB_P, true, false>, sp);
  }
  if (e1 != cudaSuccess || e2 != cudaSuccess ||
      cudaMemcpyAsync(pl->maps, h.data(), sizeof(CUtensorMap) * n_maps, cudaMemcpyHostToDevice, st) != cudaSuccess ||
      cudaMemcpyAsync(pl->tab, t.data(), sizeof(StencilTab) * nd, cudaMemcpyHostToDevice, st) != cudaSuccess ||
      cudaStreamSynchronize(st) != cudaSuccess || !attr) {
    *err = std::string("warp-specialised plan: ") + cudaGetErrorString(cudaGetLastError());
    if (pl->maps) cudaFree(pl->maps);
    if (pl->tab) cudaFree(pl->tab);
    delete pl;
    return nullptr;
  }
  return pl;
}

void ws_plan_destroy(WsPlan *pl) {
  if (!pl) return;
  if (pl->maps) cudaFree(pl->maps);
  if (pl->tab) cudaFree(pl->tab);
  delete pl;
}

namespace {
template <int TY, int MINB, bool ANISO>
void launch_u(const WsPlan *pl, const Fields &F, const Geom &G, int a_lo, int a_hi, cudaStream_t st, const HaloPush *push) {
  const int Lx = pick_chunk_ws(G, a_hi - a_lo);
  constexpr size_t smem = sizeof(SmemU<TY, npw_u(ANISO)>);
  dim3 blk(TZ, TY + 1, 1);
  dim3 grd((G.nC - M + TZ - 1) / TZ, (G.nB - 2 * M + TY - 1) / TY, (a_hi - a_lo + Lx - 1) / Lx);
  if (push)
    k_sweep_u_ws<TY, MINB, true, ANISO><<<grd, blk, smem, st>>>(pl->maps, F, G, pl->tab, a_lo, a_hi, Lx, ws_hint(), *push);
  else
    k_sweep_u_ws<TY, MINB, false, ANISO><<<grd, blk, smem, st>>>(pl->maps, F, G, pl->tab, a_lo, a_hi, Lx, ws_hint(),
                                                                    HaloPush{});
}
template <int TY, int MINB, bool ANISO>
void launch_p(const WsPlan *pl, const Fields &F, const Geom &G, int a_lo, int a_hi, cudaStream_t st, const HaloPush *push) {
  const int Lx = pick_chunk_ws(G, a_hi - a_lo);
  constexpr size_t smem = sizeof(SmemP<TY, npw_p(ANISO)>);
  dim3 blk(TZ, TY + 1, 1);
  dim3 grd((G.nC - M + TZ - 1) / TZ, (G.nB - 2 * M + TY - 1) / TY, (a_hi - a_lo + Lx - 1) / Lx);
  if (push)
    k_sweep_p_ws<TY, MINB, true, ANISO><<<grd, blk, smem, st>>>(pl->maps, F, G, pl->tab, a_lo, a_hi, Lx, ws_hint(), *push);
  else
    k_sweep_p_ws<TY, MINB, false, ANISO><<<grd, blk, smem, st>>>(pl->maps, F, G, pl->tab, a_lo, a_hi, Lx, ws_hint(),
                                                                    HaloPush{});
}
}  // namespace

int launch_sweep_u_ws(const WsPlan *pl, const Fields &F, const Geom &G, int a_lo, int a_hi, cudaStream_t st,
                      const HaloPush *push) {
  if (a_hi <= a_lo) return 0;
  if (pl->aniso) launch_u<TY_AU, MINB_WS, true>(pl, F, G, a_lo, a_hi, st, push);
  else launch_u<TY_U, MINB_U, false>(pl, F, G, a_lo, a_hi, st, push);
  return 1;
}

int launch_sweep_p_ws(const WsPlan *pl, const Fields &F, const Geom &G, int a_lo, int a_hi, cudaStream_t st,
                      const HaloPush *push) {
  if (a_hi <= a_lo) return 0;
  if (pl->aniso) launch_p<TY_AP, MINB_WS, true>(pl, F, G, a_lo, a_hi, st, push);
  else launch_p<TY_P, MINB_P, false>(pl, F, G, a_lo, a_hi, st, push);
  return 1;
}

}  // namespace fw25
