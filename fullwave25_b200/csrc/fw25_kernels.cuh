// fw25_kernels.cuh -- shared device-side definitions for the fw25 engine (sm_100a).
//
// Arithmetic contract: every floating-point operation is the same single IEEE-754 binary32 operation,
// in the same order, as the reference's sm_100 cubin (fd_u / fd_p of
// fullwave2_{2d,3d}_2_relax_isotropic_multi_gpu_sm_100_cuda129; PTX line ranges in SURVEY.md 8(a) and
// Appendix C; FFMA contraction of the final "q - s*t" taken from the SASS).  The _rn intrinsics are
// never re-associated or contracted by nvcc/ptxas, so the results are bit-identical to the reference
// wherever its own operations are (they are all .rn, no .ftz).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace fw25 {

constexpr int M = 8;

__device__ __forceinline__ float fma_(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ float mul_(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add_(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub_(float a, float b) { return __fsub_rn(a, b); }
// IEEE-754 division, correctly rounded.  A zero numerator over a non-zero, non-NaN denominator is +-0 with the
// xor of the signs (also for b = +-inf) -- answered directly, because div.rn's hardware range check sends
// zero numerators down its slow path and most of a pulsed-ultrasound domain is still exactly quiet.
__device__ __forceinline__ float div_(float a, float b) {
  if (a == 0.0f && (b < 0.0f || b > 0.0f))
    return __int_as_float((__float_as_int(a) ^ __float_as_int(b)) & 0x80000000);
  return __fdiv_rn(a, b);
}
__device__ __forceinline__ float rcp_(float a) { return __frcp_rn(a); }

// Programmatic dependent launch (griddepcontrol): a kernel launched with the programmatic-serialization attribute
// (launch_pdl, fw25_internal.h) may start while its predecessor in the stream / graph is still running.  It must not
// touch anything the predecessor writes -- or write anything the predecessor reads -- before pdl_wait(), which
// returns once the predecessor grid has completed and its memory operations are visible.  Every thread that does
// work calls pdl_wait(), so completion is transitive along a chain of such kernels.  pdl_trigger() lets the
// successor be scheduled as soon as every CTA of this grid has started.  Both are no-ops for plain launches.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// Internal layout: every field is [nA][nB][pitch] float32, C (fastest) axis padded to `pitch`
// (multiple of 4 floats => 16-byte aligned rows for float4 / TMA).  3D: (A,B,C) = (x,y,z).
// 2D: (A,B,C) = (x,-,y) with nB = 1, i.e. the reference's 2D "y" is our C axis; its v / *y* arrays
// live in the axis-C slots.
struct Fields {
  // medium (read-only)
  const float *rho, *K, *beta, *kappax, *kappau;
  const float *ax1, *bx1, *ax2, *bx2;  // apmlx*/bpmlx*: velocity sweep
  const float *au1, *bu1, *au2, *bu2;  // apmlu*/bpmlu*: pressure sweep
  const int32_t *dcmap;
  const float *dmap;                   // [9][2][ndmap]
  // anisotropic file set (per axis A,B,C; nu): only read by the ANISO instantiations of the simple sweeps
  const float *kv[3], *av[3][2], *bv[3][2];   // velocity sweep
  const float *kp[3], *ap[3][2], *bp[3][2];   // pressure sweep
  // state
  float *p;
  float *q[3];                         // velocity along A,B,C
  float *psi[3][2];                    // [axis][nu] velocity-sweep memory variables
  float *phi[3][2];                    // [axis][nu] pressure-sweep memory variables
};

struct Geom {
  int nA, nB, nC;       // local planes, rows, points per row
  int pitch;            // floats per row
  long long sA, sB;     // strides (floats) of A and B axes; sB == pitch
  int ndmap;
  float dX, dT;
  int a_rim_lo, a_rim_hi; // local A range that may be updated: global [8, nXglobal-8) ∩ owned
};

// Fused halo exchange (boundary sweeps of an x-slab): peer-mapped state arrays of the neighbour GPU, pre-offset
// so that THIS engine's element index addresses the same global plane there.  fd_u: a[0..2] = u, v, w; fd_p:
// a[0] = p.  a[0] is pushed for local planes [lo0, hi0) (the 8 planes next to the interface), v and w only for
// [lo1, hi1) (the one plane the cross terms read); the launch itself may cover more planes.
struct HaloPush {
  float *a[3];
  int lo0, hi0, lo1, hi1;
};

}  // namespace fw25
