// Measurement aid (not product code): plain streaming kernels to calibrate what HBM3e gives a
// write-only, read-only and copy stream on this B200, next to MEASURED_PEAKS.json's copy figure.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_write(double2* p, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    __stcg(p + i, make_double2(1.0, 2.0));
}
__global__ void k_read(const double2* p, size_t n, double* out) {
  double s = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    double2 v = __ldcg(p + i); s += v.x + v.y;
  }
  if (s == 12345.678) *out = s;
}
__global__ void k_copy(const double2* a, double2* b, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    __stcg(b + i, __ldcg(a + i));
}
// 9 separate write streams per block (like one pruning node: 8 partial rows + exponents)
__global__ void k_write9(double2* p, size_t n_per_row) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n_per_row) for (int r = 0; r < 9; ++r) __stcg(p + r * n_per_row + i, make_double2(1.0, r));
}
int main() {
  size_t bytes = (size_t)8 << 30, n = bytes / 16;
  double2 *a, *b; double* out;
  cudaMalloc(&a, bytes); cudaMalloc(&b, bytes); cudaMalloc(&out, 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0); k_write<<<148 * 16, 256>>>(a, n); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1); printf("write-only  %.1f GB/s\n", bytes / ms / 1e6);
    cudaEventRecord(e0); k_read<<<148 * 16, 256>>>(a, n, out); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1); printf("read-only   %.1f GB/s\n", bytes / ms / 1e6);
    cudaEventRecord(e0); k_copy<<<148 * 16, 256>>>(a, b, n); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1); printf("copy (r+w)  %.1f GB/s\n", 2.0 * bytes / ms / 1e6);
    size_t npr = n / 9;
    cudaEventRecord(e0); k_write9<<<(unsigned)((npr + 255) / 256), 256>>>(a, npr); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1); printf("write 9 rows %.1f GB/s\n", 9.0 * npr * 16 / ms / 1e6);
    cudaEventRecord(e0); cudaMemsetAsync(a, 0, bytes); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1); printf("cudaMemset  %.1f GB/s\n", bytes / ms / 1e6);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
