// fw25_points.cu -- source injection, air (pressure-release) zeroing and sensor gather.
// Replaces inject_source (3D PTX L1325-1389), inject_source_zero (L1391-1443),
// compute_genout_frame_multi (L1477-1570) and extract_pressure_values (L1572-1608): the coordinate ->
// linear-index maps are resolved once at setup (bit-exact int arithmetic on the host), so the per-step
// kernels are pure scatter / gather; injection and zeroing share one launch.
#include <algorithm>
#include <cstdlib>

#include "fw25_internal.h"

namespace fw25 {

bool pdl_enabled() {
  static const bool on = [] {
    const char *e = getenv("FW25_PDL");
    return !e || atoi(e) != 0;
  }();
  return on;
}

// p[src] = icmat[row][t] while t < nTic (overwrite, not add); afterwards a source that sits in the
// never-updated 8-cell rim falls back to 0 (the reference's proceed_time copies the zero new half
// over it).  Air voxels are zeroed after the sources in the reference's launch order: a source that is
// also an air voxel is therefore always 0 -- the host marks it dead (flag bit 1) so that injection and
// zeroing can share ONE launch without a write race.  Threads [0, n_src) inject, [n_src, n_src + n_air) zero.
// t = t_off, or *d_t + t_off when the step number lives on the device (steps replayed from a CUDA graph).
__global__ void k_inject(float *__restrict__ p, const long long *__restrict__ src_idx,
                         const int *__restrict__ src_row, const unsigned char *__restrict__ src_flag,
                         int n_src, const float *__restrict__ icmat, int nTic, int t_off,
                         const int *__restrict__ d_t, const long long *__restrict__ air_idx, int n_air) {
  pdl_trigger();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_src) {
    const unsigned char f = src_flag[i];                 // index lists are written at setup only
    const long long s = src_idx[i];
    const int row = src_row[i];
    pdl_wait();                                          // p belongs to the previous kernel until here; d_t to k_tick
    const int t = d_t ? *d_t + t_off : t_off;
    if (f & 2) return;
    if (t < nTic) p[s] = icmat[(size_t)row * nTic + t];
    else if (f & 1) p[s] = 0.0f;
  } else if (i < n_src + n_air) {
    const long long s = air_idx[i - n_src];
    pdl_wait();
    p[s] = 0.0f;
  }
}

// frame[i] = p'[sensor_i]; sensors in the 8-cell rim (idx < 0) read 0.
__global__ void k_record(const float *__restrict__ p, const long long *__restrict__ sens_idx, int n_sens,
                         float *__restrict__ frame) {
  pdl_trigger();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_sens) return;
  const long long s = sens_idx[i];
  pdl_wait();
  frame[i] = s < 0 ? 0.0f : p[s];
}

// graph-replayed form: the frame slot comes from the device-side step counter, frame = ((*d_t + t_off) / modT) % cap
__global__ void k_record_dev(const float *__restrict__ p, const long long *__restrict__ sens_idx, int n_sens,
                             float *__restrict__ frames, const int *__restrict__ d_t, int t_off, int modT, int cap) {
  pdl_trigger();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_sens) return;
  const long long s = sens_idx[i];
  pdl_wait();
  const int f = ((*d_t + t_off) / modT) % cap;
  frames[(size_t)f * n_sens + i] = s < 0 ? 0.0f : p[s];
}

// Box sensors (every point of a box, row-major: rectangular sensor masks, whole-domain recording): no index list,
// one thread per box point, lanes along the contiguous axis; 4 B read + 4 B written per point instead of 16.
// frame slot: `frames` itself, or ((*d_t + t_off) / modT) % cap when the step number lives on the device.
__global__ void k_record_box(const float *__restrict__ p, float *__restrict__ frames, long long n_sens,
                             const int *__restrict__ d_t, int t_off, int modT, int cap, SensBox B) {
  pdl_trigger();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= B.wc) return;
  const int b = blockIdx.y, a = blockIdx.z;
  pdl_wait();
  float *frame = d_t ? frames + (size_t)(((*d_t + t_off) / modT) % cap) * n_sens : frames;
  const int la = B.a0 + a, lb = B.b0 + b, lc = B.c0 + c;          // local plane, row, column
  const bool rim = la < B.a_lo || la >= B.a_hi || lb < B.b_lo || lb >= B.b_hi || lc < B.c_lo || lc >= B.c_hi;
  frame[((size_t)a * B.wb + b) * B.wc + c] = rim ? 0.0f : p[(long long)la * B.sA + (long long)lb * B.sB + lc];
}

// Plane-range forms for the time-skewed start (Engine::run_skewed): only the entries whose plane index/sA lies in
// [a_lo, a_hi) are touched, so that injection and recording can follow the sweeps block by block.
__global__ void k_inject_range(float *__restrict__ p, const long long *__restrict__ src_idx,
                               const int *__restrict__ src_row, const unsigned char *__restrict__ src_flag, int n_src,
                               const float *__restrict__ icmat, int nTic, int t, const long long *__restrict__ air_idx,
                               int n_air, long long sA, int a_lo, int a_hi) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_src) {
    const long long s = src_idx[i];
    const int a = (int)(s / sA);
    if (a < a_lo || a >= a_hi) return;
    const unsigned char f = src_flag[i];
    if (f & 2) return;
    if (t < nTic) p[s] = icmat[(size_t)src_row[i] * nTic + t];
    else if (f & 1) p[s] = 0.0f;
  } else if (i < n_src + n_air) {
    const long long s = air_idx[i - n_src];
    const int a = (int)(s / sA);
    if (a < a_lo || a >= a_hi) return;
    p[s] = 0.0f;
  }
}

__global__ void k_record_range(const float *__restrict__ p, const long long *__restrict__ sens_idx, int n_sens,
                               float *__restrict__ frame, long long sA, int a_lo, int a_hi) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_sens) return;
  const long long s = sens_idx[i];
  if (s < 0) return;                                     // rim sensors read 0: the ring starts out zeroed
  const int a = (int)(s / sA);
  if (a < a_lo || a >= a_hi) return;
  frame[i] = p[s];
}

__global__ void k_tick(int *d_t, int set, int add) {
  pdl_trigger();
  pdl_wait();
  *d_t = (set >= 0 ? set : *d_t) + add;
}

// Reference 3D behaviour (include/fw25.h, dcmap_full3d): entries whose flat index in the WHOLE dense grid
// ((x*nY + y)*nZ + z) is >= limit (= nX*nY) read 0.  Runs once at setup on the engine's padded copy.
__global__ void k_dcmap_mask(int32_t *__restrict__ dc, long long cells, int pitch, int nC, int nB,
                             long long first_plane, long long limit) {
  for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < cells;
       j += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(j % pitch);
    const long long row = j / pitch;
    const int b = (int)(row % nB);
    const long long a = row / nB + first_plane;
    const long long flat = (a * nB + b) * nC + c;
    if (c < nC && flat >= limit) dc[j] = 0;
  }
}

void launch_dcmap_mask(int32_t *dcmap, long long cells, int pitch, int nC, int nB, long long first_plane,
                       long long limit, cudaStream_t st) {
  const int blocks = (int)std::min<long long>((cells + 255) / 256, 148LL * 16);
  k_dcmap_mask<<<blocks, 256, 0, st>>>(dcmap, cells, pitch, nC, nB, first_plane, limit);
}

int launches_per_inject(int n_src, int n_air, int t, int nTic, int n_src_rim) {
  return ((n_src > 0 && (t < nTic || n_src_rim > 0)) || n_air > 0) ? 1 : 0;
}

// n_src = 0 skips the injection part (the caller passes 0 once t >= nTic and no source sits in the rim)
void launch_inject(float *p, const long long *src_idx, const int *src_row, const unsigned char *src_flag,
                   int n_src, const float *icmat, int nTic, int t, const long long *air_idx, int n_air,
                   cudaStream_t st, const int *d_t) {
  const int n = n_src + n_air;
  if (n > 0)
    launch_pdl(k_inject, dim3((n + 255) / 256), dim3(256), 0, st, p, src_idx, src_row, src_flag, n_src, icmat, nTic, t, d_t,
               air_idx, n_air);
}

void launch_record(const float *p, const long long *sens_idx, int n_sens, float *frame, cudaStream_t st) {
  if (n_sens > 0) launch_pdl(k_record, dim3((n_sens + 255) / 256), dim3(256), 0, st, p, sens_idx, n_sens, frame);
}

void launch_record_dev(const float *p, const long long *sens_idx, int n_sens, float *frames, const int *d_t, int t_off,
                       int modT, int cap, cudaStream_t st) {
  if (n_sens > 0)
    launch_pdl(k_record_dev, dim3((n_sens + 255) / 256), dim3(256), 0, st, p, sens_idx, n_sens, frames, d_t, t_off, modT,
               cap);
}

void launch_record_box(const float *p, float *frames, long long n_sens, const int *d_t, int t_off, int modT, int cap,
                       const SensBox &B, cudaStream_t st) {
  if (n_sens <= 0) return;
  dim3 grid((B.wc + 255) / 256, B.wb, B.wa);
  launch_pdl(k_record_box, grid, dim3(256), 0, st, p, frames, n_sens, d_t, t_off, modT, cap, B);
}

void launch_inject_range(float *p, const long long *src_idx, const int *src_row, const unsigned char *src_flag, int n_src,
                         const float *icmat, int nTic, int t, const long long *air_idx, int n_air, long long sA, int a_lo,
                         int a_hi, cudaStream_t st) {
  const int n = n_src + n_air;
  if (n > 0)
    k_inject_range<<<(n + 255) / 256, 256, 0, st>>>(p, src_idx, src_row, src_flag, n_src, icmat, nTic, t, air_idx, n_air, sA,
                                                    a_lo, a_hi);
}

void launch_record_range(const float *p, const long long *sens_idx, int n_sens, float *frame, long long sA, int a_lo,
                         int a_hi, cudaStream_t st) {
  if (n_sens > 0) k_record_range<<<(n_sens + 255) / 256, 256, 0, st>>>(p, sens_idx, n_sens, frame, sA, a_lo, a_hi);
}

void launch_tick(int *d_t, int set, int add, cudaStream_t st) { launch_pdl(k_tick, dim3(1), dim3(1), 0, st, d_t, set, add); }

}  // namespace fw25
