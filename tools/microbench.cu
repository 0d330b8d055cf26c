// microbench.cu -- measures the FP32-FMA roofline denominator and shared-memory broadcast costs on the box.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
// Prints one JSON line: peak FFMA TFLOP/s, LDS.128 wavefront behaviour for broadcast patterns, TF32 mma.sync rate.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int ILP>
__global__ void __launch_bounds__(256) k_ffma(float* out, int iters, float a, float b) {
  float acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 0.001f + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < ILP; ++i) acc[i] = fmaf(acc[i], a, b);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// register-tiled outer product like an SGEMM inner loop: 8x8 accumulators, operands from registers
__global__ void __launch_bounds__(256) k_ffma_tile(float* out, int iters, float seed) {
  float acc[8][8], a[8], b[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = seed + i + threadIdx.x; b[i] = seed * 0.5f + i;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] += 1e-9f; }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) s += acc[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// LDS.128 with `distinct` different 16-byte chunks per warp (broadcast among the rest)
__global__ void __launch_bounds__(256) k_lds(float* out, int iters, int distinct, int stride_f4) {
  extern __shared__ float4 sm[];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = make_float4(i, 1, 2, 3);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  int idx = (lane % distinct) * stride_f4;
  float4 acc = make_float4(0, 0, 0, 0);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      float4 v = sm[(idx + r * 64) & 4095];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    idx = (idx + 1) & 4095;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
}

// SGEMM-like inner loop from shared memory: 8x4 tile, 3 LDS.128 per 32 FFMA (layout as in njode_tiled)
__global__ void __launch_bounds__(128) k_smem_gemm(float* out, int iters) {
  extern __shared__ float4 sm[];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = make_float4(i * 1e-3f, 1, 2, 3);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rg = lane & 7, cg = lane >> 3;
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const float4* A = sm + (warp & 1) * 8 * 9 + rg * 9;      // row r: 9 float4 stride (36 floats)
  const float4* W = sm + 2048 + cg * 4 * 9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int kc = 0; kc < 8; ++kc) {
      float4 a[8], w[4];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = A[i * 16 * 9 + kc];
#pragma unroll
      for (int j = 0; j < 4; ++j) w[j] = W[j * 9 + kc];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[i][j] = fmaf(a[i].x, w[j].x, acc[i][j]);
          acc[i][j] = fmaf(a[i].y, w[j].y, acc[i][j]);
          acc[i][j] = fmaf(a[i].z, w[j].z, acc[i][j]);
          acc[i][j] = fmaf(a[i].w, w[j].w, acc[i][j]);
        }
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) s += acc[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// TF32 mma.sync m16n8k8 throughput (legacy tensor path), 8 independent accumulators per warp
__global__ void __launch_bounds__(256) k_mma_tf32(float* out, int iters) {
  float c[8][4];
  uint32_t a[4] = {0x3f800000u + threadIdx.x, 0x3f000000u, 0x3e800000u, 0x3f400000u}, b[2] = {0x3f800000u, 0x3f000000u};
#pragma unroll
  for (int i = 0; i < 8; ++i) { c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static float time_ms(F launch, int reps) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  launch(); launch();
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  float* out; CK(cudaMalloc(&out, sizeof(float) * sms * 16 * 256));
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d", prop.name, sms, prop.clockRate);
  {
    const int iters = 20000, blocks = sms * 8;
    float ms = time_ms([&] { k_ffma<8><<<blocks, 256>>>(out, iters, 1.0001f, 0.5f); }, 5);
    double flops = 2.0 * blocks * 256.0 * iters * 8 * 8;
    printf(", \"ffma_ilp8_tflops\": %.2f", flops / ms * 1e-9);
    ms = time_ms([&] { k_ffma<16><<<blocks, 256>>>(out, iters, 1.0001f, 0.5f); }, 5);
    flops = 2.0 * blocks * 256.0 * iters * 8 * 16;
    printf(", \"ffma_ilp16_tflops\": %.2f", flops / ms * 1e-9);
  }
  {
    const int iters = 20000, blocks = sms * 4;
    float ms = time_ms([&] { k_ffma_tile<<<blocks, 256>>>(out, iters, 1.0f); }, 5);
    double flops = 2.0 * blocks * 256.0 * iters * 64;
    printf(", \"ffma_tile8x8_tflops\": %.2f", flops / ms * 1e-9);
  }
  {
    CK(cudaFuncSetAttribute(k_lds, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    const int iters = 4000, blocks = sms * 2;
    const int pats[][2] = {{1, 0}, {4, 1}, {8, 1}, {8, 9}, {32, 1}, {32, 9}};
    for (auto& p : pats) {
      float ms = time_ms([&] { k_lds<<<blocks, 256, 65536>>>(out, iters, p[0], p[1]); }, 5);
      // LDS.128 warp-instructions per SM per cycle (at the reported max clock)
      double instr = (double)blocks * 8 * iters * 16 / sms;
      printf(", \"lds128_d%d_s%d_cyc_per_instr\": %.2f", p[0], p[1], ms * 1e-3 * prop.clockRate * 1e3 / instr);
    }
  }
  {
    CK(cudaFuncSetAttribute(k_smem_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    const int iters = 4000;
    for (int occ = 1; occ <= 3; ++occ) {
      const int blocks = sms * occ;
      float ms = time_ms([&] { k_smem_gemm<<<blocks, 128, 65536>>>(out, iters); }, 5);
      double flops = 2.0 * blocks * 128.0 * iters * 8 * 128;
      printf(", \"smem_gemm_8x4_occ%d_tflops\": %.2f", occ, flops / ms * 1e-9);
    }
  }
  {
    const int iters = 20000, blocks = sms * 4;
    float ms = time_ms([&] { k_mma_tf32<<<blocks, 256>>>(out, iters); }, 5);
    double flops = 2.0 * 16 * 8 * 8 * 8.0 * iters * blocks * 8;
    printf(", \"mma_sync_tf32_tflops\": %.2f", flops / ms * 1e-9);
  }
  printf("}\n");
  return 0;
}
