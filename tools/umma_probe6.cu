// umma_probe6.cu -- what slows a stream of shared-memory-bound SS MMAs (M64 N72 K8 tf32, MN-major)?
// 16 worker warps generate one kind of background traffic while the issuer streams 512 MMAs:
//   0 idle, 1 mbarrier.try_wait polling, 2 tcgen05.ld, 3 tcgen05.st, 4 STS.128 (conflict-free), 5 LDS.128, 6 LDG (L2)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../neural-jump-ode_b200/csrc/njode_umma.cuh"
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
constexpr int TILE_B = 16384;

__global__ void __launch_bounds__(544) probe(long long* out, const float* gsrc, int n_mma, int mode) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  float* tiles = (float*)base;
  float* scratch = tiles + 10 * TILE_B / 4;          // 16 KB
  __shared__ uint64_t bar_done, bar_never;
  __shared__ uint32_t tmem_s;
  __shared__ volatile int stop;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 10 * TILE_B / 4 + 4096; i += blockDim.x) tiles[i] = 1.0f;
  if (tid == 0) { umma::mbar_init(&bar_done, 1); umma::mbar_init(&bar_never, 1); umma::fence_mbar_init(); stop = 0; }
  if (warp == 0) umma::tmem_alloc(&tmem_s, 512);
  umma::fence_async_smem();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = tmem_s;
  if (warp == 16) {
    if (umma::elect_one()) {
      constexpr uint32_t idesc = umma::idesc_tf32(64, 72, 1, 1);
      const uint64_t da = umma::desc_mn(umma::smem_u32(tiles), TILE_B), db = umma::desc_mn(umma::smem_u32(tiles + 4 * TILE_B / 4), TILE_B);
      const long long t0 = clock64();
      for (int m = 0; m < n_mma; m += 16) {
#pragma unroll
        for (int i = 0; i < 16; ++i) umma::mma_ss(tmem + 192, da + 64 * i, db + 64 * i, idesc, 1);
      }
      if (mode == 7) {
        // the reverse sweep's per-step batch: 12 chain MMAs (TS, M128 N32) + 48 weight-gradient MMAs (3 passes over
        // 16 K-slices of different tile pairs, fresh accumulator), commit, wait; 8 times
        constexpr uint32_t idc = umma::idesc_tf32(128, 32, 0, 0);
        const uint64_t dbk = umma::desc_k(umma::smem_u32(tiles + 9 * TILE_B / 4));
        const uint64_t dah = da, dal = da + 2 * (TILE_B / 16), dbh = db, dbl = db + 3 * (TILE_B / 16);
        uint32_t ph = 0;
        for (int rep = 0; rep < 8; ++rep) {
          const long long t1 = clock64();
          for (int ks = 0; ks < 12; ++ks) umma::mma_ts(tmem + 160, tmem + 64 + 8 * (ks & 3), dbk + 2 * (ks & 3), idc, ks > 0);
          for (int ks = 0; ks < 16; ++ks) umma::mma_ss(tmem + 192, dal + 64 * ks, dbh + 64 * ks, idesc, ks > 0);
          for (int ks = 0; ks < 16; ++ks) umma::mma_ss(tmem + 192, dah + 64 * ks, dbl + 64 * ks, idesc, 1);
          for (int ks = 0; ks < 16; ++ks) umma::mma_ss(tmem + 192, dah + 64 * ks, dbh + 64 * ks, idesc, 1);
          const long long t2 = clock64();
          umma::commit(&bar_done);
          umma::mbar_wait(&bar_done, ph); ph ^= 1;
          out[8 + 2 * rep] = t2 - t1;
          out[9 + 2 * rep] = clock64() - t1;
        }
        stop = 1;
        return;
      }
      umma::commit(&bar_done);
      out[0] = clock64() - t0;
      umma::mbar_wait(&bar_done, 0);
      out[1] = clock64() - t0;
      stop = 1;
    }
  } else {
    const int q = warp & 3, c = warp >> 2;
    const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16) + 8 * c;
    float sink = 0.0f;
    long long ops = 0;
    while (!stop) {
      if (mode == 1) {
        uint32_t done;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(umma::smem_u32(&bar_never)), "r"(0u) : "memory");
        sink += done;
      } else if (mode == 2) {
        float v[8], w[8];
        umma::tmem_ld8_nowait(lane_base + 64, v);
        umma::tmem_ld8_nowait(lane_base + 128, w);
        umma::wait_ld();
        sink += v[0] + w[7];
      } else if (mode == 3) {
        uint32_t u[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) u[i] = i;
        umma::tmem_st8_raw(lane_base + 0, u);
        umma::tmem_st8_raw(lane_base + 32, u);
        umma::wait_st();
      } else if (mode == 4) {
        float4* p = reinterpret_cast<float4*>(scratch) + (warp * 32 + lane);
        p[0] = make_float4(sink, 1, 2, 3);
        p[512] = make_float4(sink, 1, 2, 3);
      } else if (mode == 5) {
        const float4* p = reinterpret_cast<const float4*>(scratch) + (warp * 32 + lane);
        const float4 a = p[0], b = p[512];
        sink += a.x + b.y;
      } else if (mode == 6) {
        const float4* p = reinterpret_cast<const float4*>(gsrc) + (warp * 32 + lane) + (ops & 63) * 512;
        sink += __ldcg(p).x;
      } else {
        __nanosleep(100);
      }
      ++ops;
    }
    if (tid == 0) out[2] = ops;
    if (sink == 12345.678f) out[3] = 1;
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_free(tmem, 512);
}

int main() {
  long long* d_out; float* d_src;
  CK(cudaMalloc(&d_out, 64 * sizeof(long long)));
  CK(cudaMalloc(&d_src, 1 << 20));
  CK(cudaMemset(d_src, 0, 1 << 20));
  const int smem = 1024 + 10 * TILE_B + 16384;
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const char* names[7] = {"idle (nanosleep)", "mbarrier.try_wait polling", "tcgen05.ld x8 x2 loops", "tcgen05.st x8 x2 loops", "STS.128 x2 loops",
                          "LDS.128 x2 loops", "LDG.128 loops (L2)"};
  {
    CK(cudaMemset(d_out, 0, 64 * sizeof(long long)));
    probe<<<1, 544, smem>>>(d_out, d_src, 0, 7);
    CK(cudaDeviceSynchronize());
    long long h[32];
    CK(cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost));
    printf("per-step batch (12 TS chain + 48 SS wgrad MMAs), idle workers: issue / complete cycles:");
    for (int r = 0; r < 8; ++r) printf("  %lld/%lld", h[8 + 2 * r], h[9 + 2 * r]);
    printf("\n");
  }
  for (int mode = 0; mode < 7; ++mode) {
    CK(cudaMemset(d_out, 0, 64 * sizeof(long long)));
    probe<<<1, 544, smem>>>(d_out, d_src, 512, mode);
    CK(cudaDeviceSynchronize());
    long long h[4];
    CK(cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost));
    printf("%-28s: 512 MMAs issue %6lld complete %6lld cycles (%.1f / MMA); worker-thread iterations %lld\n", names[mode], h[0], h[1], h[1] / 512.0, h[2]);
  }
  return 0;
}
