/*
 * fw25_oracle.c -- CPU restatement of the reference engine's arithmetic.  TEST INFRASTRUCTURE ONLY
 * (see fw25_oracle.h for the rules and for the PTX/SASS line ranges this follows).
 *
 * Every floating-point operation below is a single IEEE-754 binary32 operation in the same order as
 * the reference's sm_100 SASS: fmaf() where the cubin has FFMA, separate * and + (compiled with
 * -ffp-contract=off) where it has FMUL/FADD, "/" where it has the div.rn sequence, 1.0f/x for rcp.rn.
 * The reference's two-time-level storage + proceed_time copy is collapsed to an in-place update:
 * each sweep only writes arrays it reads point-wise, so the results are identical.
 *
 * Build: oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp), twice: a portable build whose fmaf()
 * goes through libm (correctly rounded) and a -mfma build; both give the same bits, the loader
 * (oracle/oracle.py) picks the -mfma one when /proc/cpuinfo lists fma.
 */
#include "fw25_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define M 8


/* per-axis coefficient arrays of one sweep: the isotropic binaries use the same array on every axis */
typedef struct {
  const float *k[3], *a[3][2], *b[3][2];
} axis_coef;

static axis_coef coef_vel(const fw25o_problem *pb) {
  axis_coef c;
  for (int ax = 0; ax < 3; ++ax) {
    c.k[ax] = pb->aniso ? pb->kappa_vel[ax] : pb->kappax;
    c.a[ax][0] = pb->aniso ? pb->a_vel[ax][0] : pb->apmlx1; c.b[ax][0] = pb->aniso ? pb->b_vel[ax][0] : pb->bpmlx1;
    c.a[ax][1] = pb->aniso ? pb->a_vel[ax][1] : pb->apmlx2; c.b[ax][1] = pb->aniso ? pb->b_vel[ax][1] : pb->bpmlx2;
  }
  return c;
}

static axis_coef coef_prs(const fw25o_problem *pb) {
  axis_coef c;
  for (int ax = 0; ax < 3; ++ax) {
    c.k[ax] = pb->aniso ? pb->kappa_prs[ax] : pb->kappau;
    c.a[ax][0] = pb->aniso ? pb->a_prs[ax][0] : pb->apmlu1; c.b[ax][0] = pb->aniso ? pb->b_prs[ax][0] : pb->bpmlu1;
    c.a[ax][1] = pb->aniso ? pb->a_prs[ax][1] : pb->apmlu2; c.b[ax][1] = pb->aniso ? pb->b_prs[ax][1] : pb->bpmlu2;
  }
  return c;
}

static inline int in_rim(const fw25o_problem *pb, int x, int y, int z) {
  if (x < M || x >= pb->nX - M) return 1;
  if (y < M || y >= pb->nY - M) return 1;
  if (pb->ndim == 3 && (z < M || z >= pb->nZ - M)) return 1;
  return 0;
}

static inline size_t cell_index(const fw25o_problem *pb, const int32_t *c) {
  if (pb->ndim == 3) return ((size_t)c[0] * pb->nY + c[1]) * pb->nZ + c[2];
  return (size_t)c[0] * pb->nY + c[1];
}

/* inject_source (PTX L1325-1389): p[coord_i] = icmat[i*nTic + t] for t < nTic -- overwrite, not add.
 * inject_source_zero (PTX L1391-1443): p = 0 at air voxels, every step, after the sources.
 * Rim cells are never written by fd_p, and the reference's proceed_time copies the (zero) new half
 * over them each step, so an injected rim cell falls back to 0 once injection stops. */
void fw25o_inject(const fw25o_problem *pb, fw25o_state *st, int t) {
  const int nd = pb->ndim;
  for (int i = 0; i < pb->ncoords; ++i) {
    const int32_t *c = pb->icc + (size_t)i * nd;
    const int rim = in_rim(pb, c[0], c[1], nd == 3 ? c[2] : M);
    if (t < pb->nTic)
      st->p[cell_index(pb, c)] = pb->icmat[(size_t)i * pb->nTic + t];
    else if (rim)
      st->p[cell_index(pb, c)] = 0.0f;
  }
  for (int i = 0; i < pb->ncoordszero; ++i) st->p[cell_index(pb, pb->icczero + (size_t)i * nd)] = 0.0f;
}

/* ------------------------------------------------------------------ 3D */

/* The reference's 3D executable loads only the first nX*nY int32 of dcmap.dat -- its `main` passes
 * count = nX*nY to the reader (ASM 0x40648c-0x4064c3: mov 0x8d0(%rsp),%edx; imul 0x8d4(%rsp),%edx)
 * whereas every other 3D map gets nX*nY*nZ (e.g. c.dat, ASM 0x4064d0-0x406510, one more imul
 * 0x8d8(%rsp)) -- into a zero-filled host array.  So in 3D the kernels see dcmap[i] for the flat
 * index i < nX*nY and 0 (the table column of the minimum sound speed) everywhere else.  Confirmed on
 * a B200 against the binary (tools/bisect3d.py: bit-exact with this rule, rel-L2 3.6e-4 without).
 * dcmap_full3d != 0 selects the documented per-voxel behaviour instead (not what the binary does). */
static inline int dcmap_3d(const fw25o_problem *pb, ptrdiff_t i) {
  if (pb->dcmap_full3d || i < (ptrdiff_t)pb->nX_dcmap * pb->nY) return pb->dcmap[i];
  return 0;
}

/* fd_u, 3D PTX L38-675.  Arg roles: rho,K,dmap,dcmap,kappax,apmlx1,bpmlx1,apmlx2,bpmlx2,p,u,v,w,psi*. */
static void sweep_u_3d(const fw25o_problem *pb, fw25o_state *st, int x_lo, int x_hi) {
  const int nX = pb->nX, nY = pb->nY, nZ = pb->nZ, nd = pb->ndmap;
  const ptrdiff_t sY = nZ, sX = (ptrdiff_t)nY * nZ;
  const float dX = pb->dX, dT = pb->dT;
  const float *restrict p = st->p;
  const axis_coef C = coef_vel(pb);
  if (x_lo < M) x_lo = M;
  if (x_hi > nX - M) x_hi = nX - M;
#pragma omp parallel for collapse(2) schedule(static)
  for (int x = x_lo; x < x_hi; ++x)
    for (int y = M; y < nY - M; ++y)
      for (int z = M; z < nZ - M; ++z) {
        const ptrdiff_t i = (ptrdiff_t)x * sX + (ptrdiff_t)y * sY + z;
        const int c = dcmap_3d(pb, i);
        float gx = 0.0f, gy = 0.0f, gz = 0.0f;
        for (int k = 1; k <= M; ++k) { /* PTX L201-319: ascending k, fma accumulate */
          const float D = pb->dmap[(2 * k) * nd + c];
          gx = fmaf(D, p[i + k * sX] - p[i - (k - 1) * sX], gx);
          gy = fmaf(D, p[i + k * sY] - p[i - (k - 1) * sY], gy);
          gz = fmaf(D, p[i + k] - p[i - (k - 1)], gz);
        }
        const float E = pb->dmap[3 * nd + c];
        /* PTX L472-547: transverse corrections, summed strictly left to right */
        float cx = p[i + sX + sY] - p[i + sY];
        cx = cx + p[i + sX - sY]; cx = cx - p[i - sY];
        cx = cx + p[i + sX + 1];  cx = cx - p[i + 1];
        cx = cx + p[i + sX - 1];  cx = cx - p[i - 1];
        float cy = p[i + sX + sY] - p[i + sX];
        cy = cy + p[i - sX + sY]; cy = cy - p[i - sX];
        cy = cy + p[i + sY + 1];  cy = cy - p[i + 1];
        cy = cy + p[i + sY - 1];  cy = cy - p[i - 1];
        float cz = p[i + sX + 1] - p[i + sX];
        cz = cz + p[i - sX + 1];  cz = cz - p[i - sX];
        cz = cz + p[i + sY + 1];  cz = cz - p[i + sY];
        cz = cz + p[i - sY + 1];  cz = cz - p[i - sY];
        gx = fmaf(E, cx, gx) / dX; /* PTX L548-550 */
        gy = fmaf(E, cy, gy) / dX;
        gz = fmaf(E, cz, gz) / dX;
        /* PTX L551-608: psi' = fma(b, psi, g*a) */
        const float px1 = fmaf(C.b[0][0][i], st->psi[0][i], gx * C.a[0][0][i]);
        const float px2 = fmaf(C.b[0][1][i], st->psi[3][i], gx * C.a[0][1][i]);
        const float py1 = fmaf(C.b[1][0][i], st->psi[1][i], gy * C.a[1][0][i]);
        const float py2 = fmaf(C.b[1][1][i], st->psi[4][i], gy * C.a[1][1][i]);
        const float pz1 = fmaf(C.b[2][0][i], st->psi[2][i], gz * C.a[2][0][i]);
        const float pz2 = fmaf(C.b[2][1][i], st->psi[5][i], gz * C.a[2][1][i]);
        st->psi[0][i] = px1; st->psi[3][i] = px2;
        st->psi[1][i] = py1; st->psi[4][i] = py2;
        st->psi[2][i] = pz1; st->psi[5][i] = pz2;
        /* PTX L609-671 + SASS: s = (dT/rho) / fma(rcp(K), p, 1); q' = FFMA(-(s), t, q) */
        const float s = (dT / pb->rho[i]) / fmaf(1.0f / pb->K[i], p[i], 1.0f);
        st->u[i] = fmaf(-s, (gx / C.k[0][i] + px1) + px2, st->u[i]);
        st->v[i] = fmaf(-s, (gy / C.k[1][i] + py1) + py2, st->v[i]);
        st->w[i] = fmaf(-s, (gz / C.k[2][i] + pz1) + pz2, st->w[i]);
      }
}

/* fd_p, 3D PTX L677-1323.  Arg roles: K,beta,dmap,dcmap,kappau,apmlu1,bpmlu1,apmlu2,bpmlu2,p,u,v,w,phi*. */
static void sweep_p_3d(const fw25o_problem *pb, fw25o_state *st, int x_lo, int x_hi) {
  const int nX = pb->nX, nY = pb->nY, nZ = pb->nZ, nd = pb->ndmap;
  const ptrdiff_t sY = nZ, sX = (ptrdiff_t)nY * nZ;
  const float dX = pb->dX, dT = pb->dT;
  const float *restrict u = st->u, *restrict v = st->v, *restrict w = st->w;
  const axis_coef C = coef_prs(pb);
  if (x_lo < M) x_lo = M;
  if (x_hi > nX - M) x_hi = nX - M;
#pragma omp parallel for collapse(2) schedule(static)
  for (int x = x_lo; x < x_hi; ++x)
    for (int y = M; y < nY - M; ++y)
      for (int z = M; z < nZ - M; ++z) {
        const ptrdiff_t i = (ptrdiff_t)x * sX + (ptrdiff_t)y * sY + z;
        const int c = dcmap_3d(pb, i);
        float hx = 0.0f, hy = 0.0f, hz = 0.0f;
        for (int k = 1; k <= M; ++k) { /* PTX L799-966 */
          const float D = pb->dmap[(2 * k) * nd + c];
          hx = fmaf(D, u[i + (k - 1) * sX] - u[i - k * sX], hx);
          hy = fmaf(D, v[i + (k - 1) * sY] - v[i - k * sY], hy);
          hz = fmaf(D, w[i + (k - 1)] - w[i - k], hz);
        }
        const float E = pb->dmap[3 * nd + c];
        /* PTX L1120-1219 */
        float cu = u[i + sY] - u[i - sX + sY];
        cu = cu + u[i - sY]; cu = cu - u[i - sX - sY];
        cu = cu + u[i + 1];  cu = cu - u[i - sX + 1];
        cu = cu + u[i - 1];  cu = cu - u[i - sX - 1];
        float cv = v[i + sX] - v[i + sX - sY];
        cv = cv + v[i - sX]; cv = cv - v[i - sX - sY];
        cv = cv + v[i + 1];  cv = cv - v[i - sY + 1];
        cv = cv + v[i - 1];  cv = cv - v[i - sY - 1];
        float cw = w[i + sX] - w[i + sX - 1];
        cw = cw + w[i - sX]; cw = cw - w[i - sX - 1];
        cw = cw + w[i + sY]; cw = cw - w[i + sY - 1];
        cw = cw + w[i - sY]; cw = cw - w[i - sY - 1];
        hx = fmaf(E, cu, hx) / dX;
        hy = fmaf(E, cv, hy) / dX;
        hz = fmaf(E, cw, hz) / dX;
        const float fx1 = fmaf(C.b[0][0][i], st->phi[0][i], hx * C.a[0][0][i]);
        const float fx2 = fmaf(C.b[0][1][i], st->phi[3][i], hx * C.a[0][1][i]);
        const float fy1 = fmaf(C.b[1][0][i], st->phi[1][i], hy * C.a[1][0][i]);
        const float fy2 = fmaf(C.b[1][1][i], st->phi[4][i], hy * C.a[1][1][i]);
        const float fz1 = fmaf(C.b[2][0][i], st->phi[2][i], hz * C.a[2][0][i]);
        const float fz2 = fmaf(C.b[2][1][i], st->phi[5][i], hz * C.a[2][1][i]);
        st->phi[0][i] = fx1; st->phi[3][i] = fx2;
        st->phi[1][i] = fy1; st->phi[4][i] = fy2;
        st->phi[2][i] = fz1; st->phi[5][i] = fz2;
        /* PTX L1286-1319 + SASS tail */
        const float Kc = pb->K[i], bt = pb->beta[i], pc = st->p[i];
        float S = hx / C.k[0][i] + hy / C.k[1][i];
        S = hz / C.k[2][i] + S;
        S = fx1 + S; S = fx2 + S; S = fy1 + S; S = fy2 + S; S = fz1 + S; S = fz2 + S;
        const float A = (dT * Kc) * S;
        const float B = fmaf(pc, (1.0f / Kc) * (1.0f - (bt + bt)), 1.0f);
        st->p[i] = fmaf(-A, B, pc);
      }
}

/* ------------------------------------------------------------------ 2D */

/* fd_u, 2D PTX L38-461 */
static void sweep_u_2d(const fw25o_problem *pb, fw25o_state *st, int x_lo, int x_hi) {
  const int nX = pb->nX, nY = pb->nY, nd = pb->ndmap;
  const ptrdiff_t sX = nY;
  const float dX = pb->dX, dT = pb->dT;
  const float *restrict p = st->p;
  const axis_coef C = coef_vel(pb);
  if (x_lo < M) x_lo = M;
  if (x_hi > nX - M) x_hi = nX - M;
#pragma omp parallel for schedule(static)
  for (int x = x_lo; x < x_hi; ++x)
    for (int y = M; y < nY - M; ++y) {
      const ptrdiff_t i = (ptrdiff_t)x * sX + y;
      const int c = pb->dcmap[i];
      float gx = 0.0f, gy = 0.0f;
      for (int k = 1; k <= M; ++k) {
        const float D = pb->dmap[(2 * k) * nd + c];
        gx = fmaf(D, p[i + k * sX] - p[i - (k - 1) * sX], gx);
        gy = fmaf(D, p[i + k] - p[i - (k - 1)], gy);
      }
      const float E = pb->dmap[3 * nd + c];
      float cx = p[i + sX + 1] - p[i + 1];
      cx = cx + p[i + sX - 1]; cx = cx - p[i - 1];
      float cy = p[i + sX + 1] - p[i + sX];
      cy = cy + p[i - sX + 1]; cy = cy - p[i - sX];
      gx = fmaf(E, cx, gx) / dX;
      gy = fmaf(E, cy, gy) / dX;
      const float px1 = fmaf(C.b[0][0][i], st->psi[0][i], gx * C.a[0][0][i]);
      const float px2 = fmaf(C.b[0][1][i], st->psi[3][i], gx * C.a[0][1][i]);
      const float py1 = fmaf(C.b[1][0][i], st->psi[1][i], gy * C.a[1][0][i]);
      const float py2 = fmaf(C.b[1][1][i], st->psi[4][i], gy * C.a[1][1][i]);
      st->psi[0][i] = px1; st->psi[3][i] = px2;
      st->psi[1][i] = py1; st->psi[4][i] = py2;
      const float s = (dT / pb->rho[i]) / fmaf(1.0f / pb->K[i], p[i], 1.0f);
      st->u[i] = fmaf(-s, (gx / C.k[0][i] + px1) + px2, st->u[i]);
      st->v[i] = fmaf(-s, (gy / C.k[1][i] + py1) + py2, st->v[i]);
    }
}

/* fd_p, 2D PTX L465-889 */
static void sweep_p_2d(const fw25o_problem *pb, fw25o_state *st, int x_lo, int x_hi) {
  const int nX = pb->nX, nY = pb->nY, nd = pb->ndmap;
  const ptrdiff_t sX = nY;
  const float dX = pb->dX, dT = pb->dT;
  const float *restrict u = st->u, *restrict v = st->v;
  const axis_coef C = coef_prs(pb);
  if (x_lo < M) x_lo = M;
  if (x_hi > nX - M) x_hi = nX - M;
#pragma omp parallel for schedule(static)
  for (int x = x_lo; x < x_hi; ++x)
    for (int y = M; y < nY - M; ++y) {
      const ptrdiff_t i = (ptrdiff_t)x * sX + y;
      const int c = pb->dcmap[i];
      float hx = 0.0f, hy = 0.0f;
      for (int k = 1; k <= M; ++k) {
        const float D = pb->dmap[(2 * k) * nd + c];
        hx = fmaf(D, u[i + (k - 1) * sX] - u[i - k * sX], hx);
        hy = fmaf(D, v[i + (k - 1)] - v[i - k], hy);
      }
      const float E = pb->dmap[3 * nd + c];
      float cu = u[i + 1] - u[i - sX + 1];
      cu = cu + u[i - 1]; cu = cu - u[i - sX - 1];
      float cv = v[i + sX] - v[i + sX - 1];
      cv = cv + v[i - sX]; cv = cv - v[i - sX - 1];
      hx = fmaf(E, cu, hx) / dX;
      hy = fmaf(E, cv, hy) / dX;
      const float fx1 = fmaf(C.b[0][0][i], st->phi[0][i], hx * C.a[0][0][i]);
      const float fx2 = fmaf(C.b[0][1][i], st->phi[3][i], hx * C.a[0][1][i]);
      const float fy1 = fmaf(C.b[1][0][i], st->phi[1][i], hy * C.a[1][0][i]);
      const float fy2 = fmaf(C.b[1][1][i], st->phi[4][i], hy * C.a[1][1][i]);
      st->phi[0][i] = fx1; st->phi[3][i] = fx2;
      st->phi[1][i] = fy1; st->phi[4][i] = fy2;
      const float Kc = pb->K[i], bt = pb->beta[i], pc = st->p[i];
      float S = hx / C.k[0][i] + hy / C.k[1][i];
      S = fx1 + S; S = fx2 + S; S = fy1 + S; S = fy2 + S;
      const float A = (dT * Kc) * S;
      const float B = fmaf(pc, (1.0f / Kc) * (1.0f - (bt + bt)), 1.0f);
      st->p[i] = fmaf(-A, B, pc);
    }
}

void fw25o_sweep_u(const fw25o_problem *pb, fw25o_state *st, int x_lo, int x_hi) {
  if (pb->ndim == 3) sweep_u_3d(pb, st, x_lo, x_hi); else sweep_u_2d(pb, st, x_lo, x_hi);
}
void fw25o_sweep_p(const fw25o_problem *pb, fw25o_state *st, int x_lo, int x_hi) {
  if (pb->ndim == 3) sweep_p_3d(pb, st, x_lo, x_hi); else sweep_p_2d(pb, st, x_lo, x_hi);
}

/* compute_genout_frame_multi / extract_pressure_values (PTX L1477-1608): gather p' at outc;
 * sensors in the 8-cell rim read 0. */
void fw25o_record(const fw25o_problem *pb, const fw25o_state *st, float *frame) {
  const int nd = pb->ndim;
  for (int i = 0; i < pb->ncoordsout; ++i) {
    const int32_t *c = pb->outc + (size_t)i * nd;
    frame[i] = in_rim(pb, c[0], c[1], nd == 3 ? c[2] : M) ? 0.0f : st->p[cell_index(pb, c)];
  }
}

int fw25o_run(const fw25o_problem *pb, float *genout, float *final_puvw) {
  const size_t n = (size_t)pb->nX * pb->nY * (pb->ndim == 3 ? pb->nZ : 1);
  const int narr = 16;
  float *mem = (float *)calloc(n * narr, sizeof(float));
  if (!mem) return 1;
  fw25o_state st;
  st.p = mem; st.u = mem + n; st.v = mem + 2 * n; st.w = mem + 3 * n;
  for (int k = 0; k < 6; ++k) { st.psi[k] = mem + (4 + k) * n; st.phi[k] = mem + (10 + k) * n; }
  size_t frame = 0;
  for (int t = 0; t < pb->nT; ++t) { /* inject -> zero -> fd_u -> fd_p -> record */
    fw25o_inject(pb, &st, t);
    fw25o_sweep_u(pb, &st, 0, pb->nX);
    fw25o_sweep_p(pb, &st, 0, pb->nX);
    if (t % pb->modT == 0) {
      fw25o_record(pb, &st, genout + frame * (size_t)pb->ncoordsout);
      ++frame;
    }
  }
  if (final_puvw) memcpy(final_puvw, mem, 4 * n * sizeof(float));
  free(mem);
  return 0;
}
