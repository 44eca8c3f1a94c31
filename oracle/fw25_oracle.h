/*
 * fw25_oracle.h -- CPU restatement of the Fullwave 2.5 isotropic 2-relaxation time-stepping
 * engine (TEST INFRASTRUCTURE ONLY).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library, and only as the checker.  The product (fullwave25_b200/) never links,
 * imports or calls it.
 *
 * What it restates: the arithmetic of the reference's shipped CUDA executable
 *   /root/reference/fullwave/solver/bins/gpu/{2d,3d}/num_relax=2/
 *       fullwave2_{2d,3d}_2_relax_isotropic_multi_gpu_sm_100_cuda129
 * whose source is NOT in the reference repository.  The restatement was recovered from the
 * binary's embedded PTX *and* its sm_100 SASS (cuobjdump -ptx / -sass):
 *   fd_u                       3D PTX L38-675   (SASS: the final "q - s*t" is ONE FFMA)
 *   fd_p                       3D PTX L677-1323 (SASS: the final "p - a*b" is ONE FFMA)
 *   inject_source              3D PTX L1325-1389
 *   inject_source_zero         3D PTX L1391-1443
 *   compute_genout_frame_multi 3D PTX L1477-1570, extract_pressure_values L1572-1608
 *   2D twins                   2D PTX L38-461, L465-889, L893-1154
 * and the host loop order from SURVEY.md section 3.3 (inject -> zero -> fd_u -> fd_p -> record).
 *
 * PARITY PINNING: the reference ships no golden vectors for this path and cannot run without a
 * GPU, so in-container the oracle is "parity unpinned".  It is pinned by tests/golden/ref_*.npz:
 * sensor traces produced by the reference's own sm_100 binary on a B200 (tools/make_ref_golden.py,
 * run through gpurun), which tests/test_oracle_golden.py compares against this code.
 */
#ifndef FW25_ORACLE_H
#define FW25_ORACLE_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fw25o_problem {
  int32_t ndim;              /* 2 or 3 */
  int32_t nX, nY, nZ;        /* extended grid (PML included); nZ == 1 when ndim == 2 */
  int32_t nT, nTic, modT;
  int32_t ndmap;
  float dX, dT;
  const float *rho, *K, *beta;
  const float *kappax, *kappau;
  const float *apmlx1, *bpmlx1, *apmlx2, *bpmlx2; /* "x" family feeds the velocity sweep */
  const float *apmlu1, *bpmlu1, *apmlu2, *bpmlu2; /* "u" family feeds the pressure sweep */
  const float *dmap;         /* [9][2][ndmap] */
  const int32_t *dcmap;      /* [nX*nY*nZ], 0-based */
  int32_t ncoords;     const int32_t *icc;     const float *icmat; /* [ncoords][ndim], [ncoords][nTic] */
  int32_t ncoordsout;  const int32_t *outc;    /* [ncoordsout][ndim] */
  int32_t ncoordszero; const int32_t *icczero; /* [ncoordszero][ndim] */
  int32_t dcmap_full3d;      /* 0: the 3D binary's behaviour (only the first nX*nY dcmap entries are
                                loaded, the rest read 0 -- see dcmap_3d in fw25_oracle.c); 1: per-voxel */
  int32_t nX_dcmap;          /* nX of the WHOLE grid for that rule (== nX unless pb is a slab view) */
  /* Anisotropic-relaxation binaries (fullwave2_{2d,3d}_2_relax_multi_gpu_*; 2D PTX L38-934, 3D PTX L38-1427 of
   * those files): the same arithmetic with one kappa / a / b array PER AXIS instead of one per sweep.
   * aniso != 0: the velocity sweep uses kappa_vel[axis], a_vel[axis][nu], b_vel[axis][nu] (files kappa{x,y,z},
   * apml{x,y,z}{1,2}, bpml{x,y,z}{1,2}) and the pressure sweep kappa_prs / a_prs / b_prs (files kappa{u,..},
   * apml{u,..}{1,2}; 2D: u, w for axes x, y; 3D: u, v, w for x, y, z); the isotropic fields above are ignored. */
  int32_t aniso;
  const float *kappa_vel[3], *a_vel[3][2], *b_vel[3][2];
  const float *kappa_prs[3], *a_prs[3][2], *b_prs[3][2];
} fw25o_problem;

/* state: p,u,v,w + 6 psi (velocity-sweep memory variables: x1,y1,z1,x2,y2,z2) + 6 phi.
 * All arrays are [nX*nY*nZ] float32, updated in place.  In 2D w, *z1, *z2 may be NULL. */
typedef struct fw25o_state {
  float *p, *u, *v, *w;
  float *psi[6];
  float *phi[6];
} fw25o_state;

/* one sweep over x in [x_lo, x_hi) (global interior clamp [8, nX-8) is applied inside) */
void fw25o_inject(const fw25o_problem *pb, fw25o_state *st, int t);
void fw25o_sweep_u(const fw25o_problem *pb, fw25o_state *st, int x_lo, int x_hi);
void fw25o_sweep_p(const fw25o_problem *pb, fw25o_state *st, int x_lo, int x_hi);
void fw25o_record(const fw25o_problem *pb, const fw25o_state *st, float *frame);

/* full run from zero state.  genout: [ceil(nT/modT)][ncoordsout].  final (optional, may be NULL):
 * 4 arrays p,u,v,w concatenated, each nX*nY*nZ.  Returns 0 on success. */
int fw25o_run(const fw25o_problem *pb, float *genout, float *final_puvw);

#ifdef __cplusplus
}
#endif
#endif
