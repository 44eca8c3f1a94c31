// H2D rate of one pinned map into the engine's pitched layout: cudaMemcpy2DAsync (what upload_map does) vs a dense
// 1-D copy into a staging buffer followed by a device-side 2-D copy.   nvcc -O2 -o probe_h2d probe_h2d.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
int main(int argc, char **argv) {
  const size_t rows = argc > 1 ? atol(argv[1]) : 400 * 1240, nz = 1240, pitch = 1248;
  float *h, *d, *stage;
  CK(cudaMallocHost(&h, rows * nz * 4));
  CK(cudaMalloc(&d, rows * pitch * 4));
  CK(cudaMalloc(&stage, rows * nz * 4));
  for (size_t i = 0; i < rows * nz; i += 1024) h[i] = 1.f;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  float ms;
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(a);
    CK(cudaMemcpy2DAsync(d, pitch * 4, h, nz * 4, nz * 4, rows, cudaMemcpyHostToDevice, 0));
    cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
    printf("2D H2D:            %.1f ms  %.1f GB/s\n", ms, rows * nz * 4 / ms / 1e6);
    cudaEventRecord(a);
    CK(cudaMemcpyAsync(stage, h, rows * nz * 4, cudaMemcpyHostToDevice, 0));
    CK(cudaMemcpy2DAsync(d, pitch * 4, stage, nz * 4, nz * 4, rows, cudaMemcpyDeviceToDevice, 0));
    cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
    printf("1D H2D + 2D D2D:   %.1f ms  %.1f GB/s\n", ms, rows * nz * 4 / ms / 1e6);
  }
  return 0;
}
