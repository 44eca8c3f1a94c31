// probe_alloc.cu -- what does it cost to get 150 GB of device memory on a B200?  (bench.py e2e setup: 30 arrays of 4.9 GB)
//   nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o tools/native/probe_alloc tools/native/probe_alloc.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

using clk = std::chrono::steady_clock;
static double ms(clk::time_point a) { return std::chrono::duration<double, std::milli>(clk::now() - a).count(); }
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
#define CU(x) do { CUresult e = (x); if (e != CUDA_SUCCESS) { const char *s; cuGetErrorString(e, &s); printf("CU error %s at line %d\n", s, __LINE__); exit(1); } } while (0)

__global__ void touch(float *p, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = 0.f;
}

int main(int argc, char **argv) {
  const double gb_total = argc > 1 ? atof(argv[1]) : 148.0;
  const int n_arr = 30;
  const size_t per = (size_t)(gb_total / n_arr * 1e9) / 4096 * 4096;
  CK(cudaSetDevice(0));
  CK(cudaFree(0));
  size_t fr, tot;
  CK(cudaMemGetInfo(&fr, &tot));
  printf("free %.1f GB of %.1f GB\n", fr / 1e9, tot / 1e9);

  {  // 30 x cudaMalloc
    std::vector<void *> p(n_arr);
    auto t0 = clk::now();
    for (int i = 0; i < n_arr; ++i) CK(cudaMalloc(&p[i], per));
    printf("cudaMalloc x%d of %.2f GB: %.1f ms total\n", n_arr, per / 1e9, ms(t0));
    t0 = clk::now();
    for (int i = 0; i < n_arr; ++i) CK(cudaMemsetAsync(p[i], 0, per));
    CK(cudaDeviceSynchronize());
    printf("  memset all: %.1f ms\n", ms(t0));
    t0 = clk::now();
    for (int i = 0; i < n_arr; ++i) CK(cudaFree(p[i]));
    printf("  cudaFree x%d: %.1f ms\n", n_arr, ms(t0));
  }
  {  // one cudaMalloc
    void *p;
    auto t0 = clk::now();
    CK(cudaMalloc(&p, per * n_arr));
    printf("cudaMalloc x1 of %.1f GB: %.1f ms\n", per * n_arr / 1e9, ms(t0));
    t0 = clk::now();
    CK(cudaFree(p));
    printf("  cudaFree: %.1f ms\n", ms(t0));
  }
  {  // again (driver may cache)
    void *p;
    auto t0 = clk::now();
    CK(cudaMalloc(&p, per * n_arr));
    printf("cudaMalloc x1 again: %.1f ms\n", ms(t0));
    CK(cudaFree(p));
  }
  {  // stream-ordered pool
    cudaMemPool_t pool;
    CK(cudaDeviceGetDefaultMemPool(&pool, 0));
    uint64_t thr = UINT64_MAX;
    CK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
    cudaStream_t s;
    CK(cudaStreamCreate(&s));
    std::vector<void *> p(n_arr);
    auto t0 = clk::now();
    for (int i = 0; i < n_arr; ++i) CK(cudaMallocAsync(&p[i], per, s));
    printf("cudaMallocAsync x%d: host %.1f ms", n_arr, ms(t0));
    CK(cudaStreamSynchronize(s));
    printf(", synced %.1f ms\n", ms(t0));
    t0 = clk::now();
    for (int i = 0; i < n_arr; ++i) CK(cudaFreeAsync(p[i], s));
    CK(cudaStreamSynchronize(s));
    printf("  cudaFreeAsync: %.1f ms\n", ms(t0));
    t0 = clk::now();
    for (int i = 0; i < n_arr; ++i) CK(cudaMallocAsync(&p[i], per, s));
    CK(cudaStreamSynchronize(s));
    printf("cudaMallocAsync x%d from the warm pool: %.1f ms\n", n_arr, ms(t0));
    for (int i = 0; i < n_arr; ++i) CK(cudaFreeAsync(p[i], s));
    CK(cudaStreamSynchronize(s));
    CK(cudaMemPoolTrimTo(pool, 0));
  }
  {  // VMM: reserve + create + map, per array, and in 512 MB pieces from a second thread's point of view
    CUmemAllocationProp prop = {};
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = 0;
    size_t gran = 0;
    CU(cuMemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
    printf("VMM granularity (recommended) %.1f MB\n", gran / 1e6);
    const size_t sz = (per * n_arr + gran - 1) / gran * gran;
    CUdeviceptr va;
    auto t0 = clk::now();
    CU(cuMemAddressReserve(&va, sz, 0, 0, 0));
    double t_res = ms(t0);
    CUmemGenericAllocationHandle h;
    t0 = clk::now();
    CU(cuMemCreate(&h, sz, &prop, 0));
    double t_create = ms(t0);
    t0 = clk::now();
    CU(cuMemMap(va, sz, 0, h, 0));
    double t_map = ms(t0);
    CUmemAccessDesc ad = {};
    ad.location = prop.location;
    ad.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    t0 = clk::now();
    CU(cuMemSetAccess(va, sz, &ad, 1));
    double t_acc = ms(t0);
    printf("VMM one %.1f GB handle: reserve %.1f, create %.1f, map %.1f, setaccess %.1f ms\n", sz / 1e9, t_res, t_create,
           t_map, t_acc);
    touch<<<1184, 256>>>((float *)va, 1 << 20);
    CK(cudaDeviceSynchronize());
    t0 = clk::now();
    CU(cuMemUnmap(va, sz));
    CU(cuMemRelease(h));
    CU(cuMemAddressFree(va, sz));
    printf("  unmap + release: %.1f ms\n", ms(t0));
    // in 1 GB pieces: create/map/setaccess piecewise (could be pipelined with uploads)
    const size_t piece = ((size_t)1 << 30) / gran * gran;
    const int n_p = (int)(sz / piece);
    CU(cuMemAddressReserve(&va, (size_t)n_p * piece, 0, 0, 0));
    std::vector<CUmemGenericAllocationHandle> hs(n_p);
    t0 = clk::now();
    for (int i = 0; i < n_p; ++i) {
      CU(cuMemCreate(&hs[i], piece, &prop, 0));
      CU(cuMemMap(va + (size_t)i * piece, piece, 0, hs[i], 0));
      CU(cuMemSetAccess(va + (size_t)i * piece, piece, &ad, 1));
    }
    printf("VMM %d x 1 GB pieces: %.1f ms (%.2f ms per GB)\n", n_p, ms(t0), ms(t0) / n_p);
    for (int i = 0; i < n_p; ++i) { CU(cuMemUnmap(va + (size_t)i * piece, piece)); CU(cuMemRelease(hs[i])); }
    CU(cuMemAddressFree(va, (size_t)n_p * piece));
  }
  return 0;
}
