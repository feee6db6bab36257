// Measurement aid (not product code): what the FP64 pipes of this B200 sustain from registers only --
// DMMA (mma.sync.m8n8k4.f64) and plain DFMA, as a function of warps per SM and independent accumulator chains.
// SURVEY 8d asks for this number as the denominator of the S = 64 (C5) roofline.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dmma_peak dmma_peak.cu && ./dmma_peak
#include <cstdio>
#include <cuda_runtime.h>

template <int CH>
__global__ void k_dmma(double* out, int iters) {
  double c[CH][2];
  for (int i = 0; i < CH; ++i) c[i][0] = c[i][1] = 0.0;
  const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
  for (int i = 0; i < CH; ++i) s += c[i][0] + c[i][1];
  if (s == 123.456) out[0] = s;
}
template <int CH>
__global__ void k_dfma(double* out, int iters) {
  double c[CH];
  for (int i = 0; i < CH; ++i) c[i] = threadIdx.x;
  const double a = 1.0 + threadIdx.x * 1e-12, b = 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) c[i] = fma(c[i], a, b);
  }
  double s = 0;
  for (int i = 0; i < CH; ++i) s += c[i];
  if (s == 123.456) out[0] = s;
}
template <typename F>
static float time_ms(F launch) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  launch(); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) {
    cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  return best;
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  double* out; cudaMalloc(&out, 8);
  const int iters = 20000;
  printf("%s, %d SMs\n", p.name, sms);
  for (int warps : {4, 8, 16, 32}) {
    float ms;
    ms = time_ms([&] { k_dmma<1><<<sms, warps * 32>>>(out, iters); });
    printf("DMMA m8n8k4  warps/SM=%2d chains=1 : %7.2f TFLOP/s\n", warps, 2.0 * 256 * 1 * iters * warps * sms / ms / 1e9);
    ms = time_ms([&] { k_dmma<4><<<sms, warps * 32>>>(out, iters); });
    printf("DMMA m8n8k4  warps/SM=%2d chains=4 : %7.2f TFLOP/s\n", warps, 2.0 * 256 * 4 * iters * warps * sms / ms / 1e9);
    ms = time_ms([&] { k_dmma<16><<<sms, warps * 32>>>(out, iters); });
    printf("DMMA m8n8k4  warps/SM=%2d chains=16: %7.2f TFLOP/s\n", warps, 2.0 * 256 * 16 * iters * warps * sms / ms / 1e9);
    ms = time_ms([&] { k_dfma<8><<<sms, warps * 32>>>(out, iters); });
    printf("DFMA         warps/SM=%2d chains=8 : %7.2f TFLOP/s\n", warps, 2.0 * 32 * 8 * iters * warps * sms / ms / 1e9);
  }
  return 0;
}
