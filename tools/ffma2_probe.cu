// ffma2_probe.cu -- FP32 FMA issue: scalar FFMA vs packed fma.rn.f32x2 (FFMA2) on sm_100a
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256) k_ffma(float* out, int iters, float a, float b) {
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 0.001f + i;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], a, b);
  float s = 0; for (int i = 0; i < 16; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) k_ffma2(float* out, int iters, float a, float b) {
  unsigned long long acc[8], aa, bb;
  asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
  asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
#pragma unroll
  for (int i = 0; i < 8; ++i) { float x = threadIdx.x * 0.001f + i; asm("mov.b64 %0, {%1, %2};" : "=l"(acc[i]) : "f"(x), "f"(x + 0.5f)); }
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[i]) : "l"(aa), "l"(bb));
  float s = 0;
  for (int i = 0; i < 8; ++i) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[i])); s += lo + hi; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* out; cudaMalloc(&out, sms * 8 * 256 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 4000;
  for (int which = 0; which < 2; ++which) {
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
      cudaEventRecord(e0);
      if (which == 0) k_ffma<<<sms * 8, 256>>>(out, iters, 1.0001f, 0.5f); else k_ffma2<<<sms * 8, 256>>>(out, iters, 1.0001f, 0.5f);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
    }
    const double flops = 2.0 * sms * 8 * 256.0 * iters * 8 * 16;
    printf("%s: %.2f TFLOP/s (%.3f ms)\n", which ? "fma.rn.f32x2 (8 per 16 flop-lanes)" : "fma.rn.f32 (scalar)", flops / best * 1e-9, best);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
