// fw25_sweeps_simple.cu -- "simple" sweep kernels: one thread per cell, lanes along the contiguous
// axis, all stencil taps through L1/L2.  This is the correctness anchor and the fallback for shapes
// the TMA-tiled kernels do not take; it replaces the reference's fd_u / fd_p launch
// (3D PTX L38-675 / L677-1323, 2D PTX L38-461 / L465-889) with coalesced accesses and an in-place
// update (no second time level, no proceed_time copy).
#include "fw25_kernels.cuh"
#include "fw25_internal.h"

namespace fw25 {

// ANISO: per-axis kappa / a / b maps (the anisotropic engine family); otherwise one set per sweep.
template <int ND, bool ANISO>
__global__ void __launch_bounds__(256) k_sweep_u_simple(Fields F, Geom G, int a_lo, int a_hi) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = (ND == 3) ? (blockIdx.y * blockDim.y + threadIdx.y) : 0;
  const int a = a_lo + blockIdx.z;
  if (a >= a_hi) return;
  if (c < M || c >= G.nC - M) return;
  if (ND == 3 && (b < M || b >= G.nB - M)) return;
  const long long sA = G.sA, sB = G.sB;
  const long long i = a * sA + b * sB + c;
  const float *__restrict__ p = F.p;
  const int ci = F.dcmap[i];
  const float *dm = F.dmap + ci;
  const int nd = G.ndmap;
  float gA = 0.f, gB = 0.f, gC = 0.f;
#pragma unroll
  for (int k = 1; k <= M; ++k) {
    const float D = __ldg(dm + 2 * k * nd);
    gA = fma_(D, sub_(p[i + k * sA], p[i - (k - 1) * sA]), gA);
    if (ND == 3) gB = fma_(D, sub_(p[i + k * sB], p[i - (k - 1) * sB]), gB);
    gC = fma_(D, sub_(p[i + k], p[i - (k - 1)]), gC);
  }
  const float E = __ldg(dm + 3 * nd);
  float cA, cB = 0.f, cC;
  if (ND == 3) {
    cA = sub_(p[i + sA + sB], p[i + sB]);
    cA = add_(cA, p[i + sA - sB]); cA = sub_(cA, p[i - sB]);
    cA = add_(cA, p[i + sA + 1]);  cA = sub_(cA, p[i + 1]);
    cA = add_(cA, p[i + sA - 1]);  cA = sub_(cA, p[i - 1]);
    cB = sub_(p[i + sA + sB], p[i + sA]);
    cB = add_(cB, p[i - sA + sB]); cB = sub_(cB, p[i - sA]);
    cB = add_(cB, p[i + sB + 1]);  cB = sub_(cB, p[i + 1]);
    cB = add_(cB, p[i + sB - 1]);  cB = sub_(cB, p[i - 1]);
    cC = sub_(p[i + sA + 1], p[i + sA]);
    cC = add_(cC, p[i - sA + 1]);  cC = sub_(cC, p[i - sA]);
    cC = add_(cC, p[i + sB + 1]);  cC = sub_(cC, p[i + sB]);
    cC = add_(cC, p[i - sB + 1]);  cC = sub_(cC, p[i - sB]);
  } else {
    cA = sub_(p[i + sA + 1], p[i + 1]);
    cA = add_(cA, p[i + sA - 1]); cA = sub_(cA, p[i - 1]);
    cC = sub_(p[i + sA + 1], p[i + sA]);
    cC = add_(cC, p[i - sA + 1]); cC = sub_(cC, p[i - sA]);
  }
  const float dX = G.dX;
  gA = div_(fma_(E, cA, gA), dX);
  if (ND == 3) gB = div_(fma_(E, cB, gB), dX);
  gC = div_(fma_(E, cC, gC), dX);

  const float s = div_(div_(G.dT, F.rho[i]), fma_(rcp_(F.K[i]), p[i], 1.0f));
  float a1, b1, a2, b2, kx;
  if (!ANISO) { a1 = F.ax1[i]; b1 = F.bx1[i]; a2 = F.ax2[i]; b2 = F.bx2[i]; kx = F.kappax[i]; }
  {
    if (ANISO) { a1 = F.av[0][0][i]; b1 = F.bv[0][0][i]; a2 = F.av[0][1][i]; b2 = F.bv[0][1][i]; kx = F.kv[0][i]; }
    const float m1 = fma_(b1, F.psi[0][0][i], mul_(gA, a1));
    const float m2 = fma_(b2, F.psi[0][1][i], mul_(gA, a2));
    F.psi[0][0][i] = m1; F.psi[0][1][i] = m2;
    F.q[0][i] = fma_(-s, add_(add_(div_(gA, kx), m1), m2), F.q[0][i]);
  }
  if (ND == 3) {
    if (ANISO) { a1 = F.av[1][0][i]; b1 = F.bv[1][0][i]; a2 = F.av[1][1][i]; b2 = F.bv[1][1][i]; kx = F.kv[1][i]; }
    const float m1 = fma_(b1, F.psi[1][0][i], mul_(gB, a1));
    const float m2 = fma_(b2, F.psi[1][1][i], mul_(gB, a2));
    F.psi[1][0][i] = m1; F.psi[1][1][i] = m2;
    F.q[1][i] = fma_(-s, add_(add_(div_(gB, kx), m1), m2), F.q[1][i]);
  }
  {
    if (ANISO) { a1 = F.av[2][0][i]; b1 = F.bv[2][0][i]; a2 = F.av[2][1][i]; b2 = F.bv[2][1][i]; kx = F.kv[2][i]; }
    const float m1 = fma_(b1, F.psi[2][0][i], mul_(gC, a1));
    const float m2 = fma_(b2, F.psi[2][1][i], mul_(gC, a2));
    F.psi[2][0][i] = m1; F.psi[2][1][i] = m2;
    F.q[2][i] = fma_(-s, add_(add_(div_(gC, kx), m1), m2), F.q[2][i]);
  }
}

template <int ND, bool ANISO>
__global__ void __launch_bounds__(256) k_sweep_p_simple(Fields F, Geom G, int a_lo, int a_hi) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = (ND == 3) ? (blockIdx.y * blockDim.y + threadIdx.y) : 0;
  const int a = a_lo + blockIdx.z;
  if (a >= a_hi) return;
  if (c < M || c >= G.nC - M) return;
  if (ND == 3 && (b < M || b >= G.nB - M)) return;
  const long long sA = G.sA, sB = G.sB;
  const long long i = a * sA + b * sB + c;
  const float *__restrict__ u = F.q[0];
  const float *__restrict__ v = F.q[1];
  const float *__restrict__ w = F.q[2];
  const int ci = F.dcmap[i];
  const float *dm = F.dmap + ci;
  const int nd = G.ndmap;
  float hA = 0.f, hB = 0.f, hC = 0.f;
#pragma unroll
  for (int k = 1; k <= M; ++k) {
    const float D = __ldg(dm + 2 * k * nd);
    hA = fma_(D, sub_(u[i + (k - 1) * sA], u[i - k * sA]), hA);
    if (ND == 3) hB = fma_(D, sub_(v[i + (k - 1) * sB], v[i - k * sB]), hB);
    hC = fma_(D, sub_(w[i + (k - 1)], w[i - k]), hC);
  }
  const float E = __ldg(dm + 3 * nd);
  float cA, cB = 0.f, cC;
  if (ND == 3) {
    cA = sub_(u[i + sB], u[i - sA + sB]);
    cA = add_(cA, u[i - sB]); cA = sub_(cA, u[i - sA - sB]);
    cA = add_(cA, u[i + 1]);  cA = sub_(cA, u[i - sA + 1]);
    cA = add_(cA, u[i - 1]);  cA = sub_(cA, u[i - sA - 1]);
    cB = sub_(v[i + sA], v[i + sA - sB]);
    cB = add_(cB, v[i - sA]); cB = sub_(cB, v[i - sA - sB]);
    cB = add_(cB, v[i + 1]);  cB = sub_(cB, v[i - sB + 1]);
    cB = add_(cB, v[i - 1]);  cB = sub_(cB, v[i - sB - 1]);
    cC = sub_(w[i + sA], w[i + sA - 1]);
    cC = add_(cC, w[i - sA]); cC = sub_(cC, w[i - sA - 1]);
    cC = add_(cC, w[i + sB]); cC = sub_(cC, w[i + sB - 1]);
    cC = add_(cC, w[i - sB]); cC = sub_(cC, w[i - sB - 1]);
  } else {
    cA = sub_(u[i + 1], u[i - sA + 1]);
    cA = add_(cA, u[i - 1]); cA = sub_(cA, u[i - sA - 1]);
    cC = sub_(w[i + sA], w[i + sA - 1]);
    cC = add_(cC, w[i - sA]); cC = sub_(cC, w[i - sA - 1]);
  }
  const float dX = G.dX;
  hA = div_(fma_(E, cA, hA), dX);
  if (ND == 3) hB = div_(fma_(E, cB, hB), dX);
  hC = div_(fma_(E, cC, hC), dX);

  float a1, b1, a2, b2, kA, kB = 1.f, kC;
  if (!ANISO) { a1 = F.au1[i]; b1 = F.bu1[i]; a2 = F.au2[i]; b2 = F.bu2[i]; kA = kB = kC = F.kappau[i]; }
  else { a1 = F.ap[0][0][i]; b1 = F.bp[0][0][i]; a2 = F.ap[0][1][i]; b2 = F.bp[0][1][i]; kA = F.kp[0][i]; kC = F.kp[2][i]; }
  const float fA1 = fma_(b1, F.phi[0][0][i], mul_(hA, a1));
  const float fA2 = fma_(b2, F.phi[0][1][i], mul_(hA, a2));
  F.phi[0][0][i] = fA1; F.phi[0][1][i] = fA2;
  float fB1 = 0.f, fB2 = 0.f;
  if (ND == 3) {
    if (ANISO) { a1 = F.ap[1][0][i]; b1 = F.bp[1][0][i]; a2 = F.ap[1][1][i]; b2 = F.bp[1][1][i]; kB = F.kp[1][i]; }
    fB1 = fma_(b1, F.phi[1][0][i], mul_(hB, a1));
    fB2 = fma_(b2, F.phi[1][1][i], mul_(hB, a2));
    F.phi[1][0][i] = fB1; F.phi[1][1][i] = fB2;
  }
  if (ANISO) { a1 = F.ap[2][0][i]; b1 = F.bp[2][0][i]; a2 = F.ap[2][1][i]; b2 = F.bp[2][1][i]; }
  const float fC1 = fma_(b1, F.phi[2][0][i], mul_(hC, a1));
  const float fC2 = fma_(b2, F.phi[2][1][i], mul_(hC, a2));
  F.phi[2][0][i] = fC1; F.phi[2][1][i] = fC2;

  float S;
  if (ND == 3) {
    S = add_(div_(hA, kA), div_(hB, kB));
    S = add_(div_(hC, kC), S);
    S = add_(fA1, S); S = add_(fA2, S); S = add_(fB1, S); S = add_(fB2, S);
    S = add_(fC1, S); S = add_(fC2, S);
  } else {
    S = add_(div_(hA, kA), div_(hC, kC));
    S = add_(fA1, S); S = add_(fA2, S); S = add_(fC1, S); S = add_(fC2, S);
  }
  const float Kc = F.K[i], bt = F.beta[i], pc = F.p[i];
  const float Aterm = mul_(mul_(G.dT, Kc), S);
  const float Bterm = fma_(pc, mul_(rcp_(Kc), sub_(1.0f, add_(bt, bt))), 1.0f);
  F.p[i] = fma_(-Aterm, Bterm, pc);
}

static inline dim3 simple_block(int nd) { return nd == 3 ? dim3(64, 4, 1) : dim3(256, 1, 1); }

void launch_sweep_u_simple(int ndim, const Fields &F, const Geom &G, int a_lo, int a_hi, cudaStream_t st, bool aniso) {
  if (a_hi <= a_lo) return;
  dim3 blk = simple_block(ndim);
  for (int a0 = a_lo; a0 < a_hi; a0 += 32768) {  // gridDim.z <= 65535
    const int a1 = a0 + 32768 < a_hi ? a0 + 32768 : a_hi;
    dim3 grd((G.nC + blk.x - 1) / blk.x, (G.nB + blk.y - 1) / blk.y, a1 - a0);
    if (ndim == 3 && aniso) k_sweep_u_simple<3, true><<<grd, blk, 0, st>>>(F, G, a0, a1);
    else if (ndim == 3) k_sweep_u_simple<3, false><<<grd, blk, 0, st>>>(F, G, a0, a1);
    else if (aniso) k_sweep_u_simple<2, true><<<grd, blk, 0, st>>>(F, G, a0, a1);
    else k_sweep_u_simple<2, false><<<grd, blk, 0, st>>>(F, G, a0, a1);
  }
}

void launch_sweep_p_simple(int ndim, const Fields &F, const Geom &G, int a_lo, int a_hi, cudaStream_t st, bool aniso) {
  if (a_hi <= a_lo) return;
  dim3 blk = simple_block(ndim);
  for (int a0 = a_lo; a0 < a_hi; a0 += 32768) {  // gridDim.z <= 65535
    const int a1 = a0 + 32768 < a_hi ? a0 + 32768 : a_hi;
    dim3 grd((G.nC + blk.x - 1) / blk.x, (G.nB + blk.y - 1) / blk.y, a1 - a0);
    if (ndim == 3 && aniso) k_sweep_p_simple<3, true><<<grd, blk, 0, st>>>(F, G, a0, a1);
    else if (ndim == 3) k_sweep_p_simple<3, false><<<grd, blk, 0, st>>>(F, G, a0, a1);
    else if (aniso) k_sweep_p_simple<2, true><<<grd, blk, 0, st>>>(F, G, a0, a1);
    else k_sweep_p_simple<2, false><<<grd, blk, 0, st>>>(F, G, a0, a1);
  }
}

}  // namespace fw25
