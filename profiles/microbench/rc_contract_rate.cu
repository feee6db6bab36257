// Measurement aid (not product code): the contraction loop of kernels_dmma_rc.cuh in isolation --
// W warps per SM sub-partition each run rc_contract back to back on a resident stage; reports the DMMA
// rate as a fraction of the 16-cycle issue interval.  Used to separate "the loop itself" from the
// surrounding op pipeline when the kernel sits below the FP64 tensor peak.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../cybayes_b200/csrc/kernels_dmma_rc.cuh"
using namespace cb;

template <int MODE>
__global__ void __launch_bounds__(256, 1) k_rate(double* out, int iters, long long* cyc) {
  using Cfg = RcCfg<64>;
  extern __shared__ __align__(16) double sm[];
  for (int i = threadIdx.x; i < Cfg::MAT_D; i += blockDim.x) sm[i] = 1.0 / 64 + i * 1e-9;
  __syncthreads();
  const int lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
  double cur[2][Cfg::NT][2], acc[2][Cfg::NT][2];
  _Pragma("unroll") for (int m = 0; m < 2; ++m) _Pragma("unroll") for (int n = 0; n < Cfg::NT; ++n) { cur[m][n][0] = 1.0 + lane * 1e-6; cur[m][n][1] = 0.5; }
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    rc_contract<64>(acc, cur, sm, g, t4);
    _Pragma("unroll") for (int m = 0; m < 2; ++m) _Pragma("unroll") for (int n = 0; n < Cfg::NT; ++n) { cur[m][n][0] = acc[m][n][0] * 0.015; cur[m][n][1] = acc[m][n][1] * 0.015; }
  }
  const long long t1 = clock64();
  double s = 0;
  _Pragma("unroll") for (int m = 0; m < 2; ++m) _Pragma("unroll") for (int n = 0; n < Cfg::NT; ++n) s += cur[m][n][0] + cur[m][n][1];
  if (s == 123.456) out[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

static void run(int warps) {
  double* out; long long* cyc; cudaMalloc(&out, 8); cudaMalloc(&cyc, 8);
  const int iters = 200;
  size_t smem = RcCfg<64>::MAT_D * 8;
  cudaFuncSetAttribute(k_rate<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k_rate<0><<<148, warps * 32, smem>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  k_rate<0><<<148, warps * 32, smem>>>(out, iters, cyc);
  cudaError_t e = cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  const double per = (double)h / iters;                 // cycles per contraction per warp
  const double wps = warps / 4.0;                       // warps per sub-partition
  printf("warps/SM=%2d: %8.0f cycles per contraction per warp; pipe busy %.1f %% (256 DMMA x 16 cyc x %.2f warps/SMSP) %s\n",
         warps, per, 100.0 * 256 * 16 * wps / per, wps, cudaGetErrorString(e));
}
int main() {
  run(4); run(8);
  return 0;
}
