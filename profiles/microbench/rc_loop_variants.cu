// Measurement aid (not product code): variants of the register-carried contraction loop (kernels_dmma_rc.cuh),
// to find what keeps a lone warp below the DMMA issue rate.  S = 64: 8 state tiles, 256 DMMA per contraction.
#include <cstdio>
#include <cuda_runtime.h>
constexpr int NT = 8, PS = 64;
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int VAR>
__device__ __forceinline__ void contract(double (&acc)[2][NT][2], const double (&cur)[2][NT][2], const double* Pm, int g, int t4) {
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int n = 0; n < NT; ++n) acc[m][n][0] = acc[m][n][1] = 0.0;
  const double* p0 = Pm + g * PS + 2 * t4;
  if (VAR == 1) {  // no shared-memory loads in the loop
    const double2 b = *reinterpret_cast<const double2*>(p0);
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
      for (int n0 = 0; n0 < NT; n0 += 4) {
#pragma unroll
        for (int i = 0; i < 4; ++i) { dmma(acc[0][n0 + i][0], acc[0][n0 + i][1], cur[0][j][0], b.x); dmma(acc[1][n0 + i][0], acc[1][n0 + i][1], cur[1][j][0], b.x); }
#pragma unroll
        for (int i = 0; i < 4; ++i) { dmma(acc[0][n0 + i][0], acc[0][n0 + i][1], cur[0][j][1], b.y); dmma(acc[1][n0 + i][0], acc[1][n0 + i][1], cur[1][j][1], b.y); }
      }
  } else if (VAR == 0) {  // as in the product
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
      for (int n0 = 0; n0 < NT; n0 += 4) {
        double2 b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) b[i] = *reinterpret_cast<const double2*>(p0 + 8 * j + (8 * (n0 + i)) * PS);
#pragma unroll
        for (int i = 0; i < 4; ++i) { dmma(acc[0][n0 + i][0], acc[0][n0 + i][1], cur[0][j][0], b[i].x); dmma(acc[1][n0 + i][0], acc[1][n0 + i][1], cur[1][j][0], b[i].x); }
#pragma unroll
        for (int i = 0; i < 4; ++i) { dmma(acc[0][n0 + i][0], acc[0][n0 + i][1], cur[0][j][1], b[i].y); dmma(acc[1][n0 + i][0], acc[1][n0 + i][1], cur[1][j][1], b[i].y); }
      }
  } else if (VAR == 2) {  // explicit software pipeline: the next group's B fragments are requested before this group's DMMAs
    double2 b[4], nb[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) b[i] = *reinterpret_cast<const double2*>(p0 + (8 * i) * PS);
#pragma unroll
    for (int s = 0; s < 2 * NT; ++s) {
      const int j = s >> 1, n0 = (s & 1) * 4;
      if (s + 1 < 2 * NT) {
        const int j2 = (s + 1) >> 1, m0 = ((s + 1) & 1) * 4;
#pragma unroll
        for (int i = 0; i < 4; ++i) nb[i] = *reinterpret_cast<const double2*>(p0 + 8 * j2 + (8 * (m0 + i)) * PS);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) { dmma(acc[0][n0 + i][0], acc[0][n0 + i][1], cur[0][j][0], b[i].x); dmma(acc[1][n0 + i][0], acc[1][n0 + i][1], cur[1][j][0], b[i].x); }
#pragma unroll
      for (int i = 0; i < 4; ++i) { dmma(acc[0][n0 + i][0], acc[0][n0 + i][1], cur[0][j][1], b[i].y); dmma(acc[1][n0 + i][0], acc[1][n0 + i][1], cur[1][j][1], b[i].y); }
#pragma unroll
      for (int i = 0; i < 4; ++i) b[i] = nb[i];
    }
  } else if (VAR == 3) {  // one m-tile at a time (8 accumulator chains per pass, B fragments loaded twice)
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        double2 b[NT];
#pragma unroll
        for (int n = 0; n < NT; ++n) b[n] = *reinterpret_cast<const double2*>(p0 + 8 * j + (8 * n) * PS);
#pragma unroll
        for (int n = 0; n < NT; ++n) dmma(acc[m][n][0], acc[m][n][1], cur[m][j][0], b[n].x);
#pragma unroll
        for (int n = 0; n < NT; ++n) dmma(acc[m][n][0], acc[m][n][1], cur[m][j][1], b[n].y);
      }
  }
}
template <int VAR>
__global__ void __launch_bounds__(256, 1) k_rate(double* out, int iters, long long* cyc) {
  extern __shared__ __align__(16) double sm[];
  for (int i = threadIdx.x; i < 64 * PS; i += blockDim.x) sm[i] = 1.0 / 64 + i * 1e-9;
  __syncthreads();
  const int lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
  double cur[2][NT][2], acc[2][NT][2];
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int n = 0; n < NT; ++n) { cur[m][n][0] = 1.0 + lane * 1e-6; cur[m][n][1] = 0.5; }
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    contract<VAR>(acc, cur, sm, g, t4);
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
      for (int n = 0; n < NT; ++n) { cur[m][n][0] = acc[m][n][0] * 0.015; cur[m][n][1] = acc[m][n][1] * 0.015; }
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int n = 0; n < NT; ++n) s += cur[m][n][0] + cur[m][n][1];
  if (s == 123.456) out[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int VAR> static void run(int warps) {
  double* out; long long* cyc; cudaMalloc(&out, 8); cudaMalloc(&cyc, 8);
  const int iters = 200;
  size_t smem = 64 * PS * 8;
  cudaFuncSetAttribute(k_rate<VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k_rate<VAR><<<148, warps * 32, smem>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  k_rate<VAR><<<148, warps * 32, smem>>>(out, iters, cyc);
  cudaError_t e = cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  const double per = (double)h / iters, wps = warps / 4.0;
  printf("variant %d warps/SM=%2d: %7.0f cycles per contraction; pipe busy %.1f %%  %s\n", VAR, warps, per, 100.0 * 256 * 16 * wps / per, cudaGetErrorString(e));
}
int main() {
  run<0>(4); run<0>(8); run<1>(4); run<1>(8); run<2>(4); run<2>(8); run<3>(4); run<3>(8);
  return 0;
}
