// umma_probe5.cu -- what can the CUDA cores do while the tensor pipe streams shared-memory-bound MMAs?
// One issuer warp queues `n_mma` SS MN-major tf32 MMAs (M64 N72 K8, 32 cycles each = smem operand floor);
// four worker warps time single operations (tcgen05.st / tcgen05.ld / STS / LDS / LDG / local) issued while
// the MMAs are in flight, against the same operations with an idle tensor pipe.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/umma_probe5 tools/umma_probe5.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../neural-jump-ode_b200/csrc/njode_umma.cuh"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

constexpr int TILE_B = 16384;
constexpr int N_OPS = 7, N_REP = 12;

__global__ void __launch_bounds__(160) probe(long long* out, const float* gsrc, int n_mma, int paced, int ts_first, int fill) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  float* tiles = (float*)base;                      // 10 MN tiles
  float* scratch = tiles + 10 * TILE_B / 4;         // 8 KB worker scratch
  __shared__ uint64_t bar_done, bar_pace[2], bar_go;
  __shared__ uint32_t tmem_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 10 * TILE_B / 4 + 2048; i += blockDim.x) tiles[i] = fill ? __uint_as_float(0x3f000000u + (uint32_t)(i * 2654435761u >> 9)) : 0.0f;
  if (tid == 0) {
    umma::mbar_init(&bar_done, 1); umma::mbar_init(&bar_pace[0], 1); umma::mbar_init(&bar_pace[1], 1); umma::mbar_init(&bar_go, 1);
    umma::fence_mbar_init();
  }
  if (warp == 0) umma::tmem_alloc(&tmem_s, 512);
  umma::fence_async_smem();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = tmem_s;
  if (warp == 4) {
    if (umma::elect_one()) {
      constexpr uint32_t idesc = umma::idesc_tf32(64, 72, 1, 1);
      const uint64_t da = umma::desc_mn(umma::smem_u32(tiles), TILE_B), db = umma::desc_mn(umma::smem_u32(tiles + 4 * TILE_B / 4), TILE_B);
      uint32_t ph[2] = {0, 0};
      const long long t0 = clock64();
      if (ts_first) {   // a chain-style TS MMA (A from TMEM columns 0..31, K-major B) queued ahead of the SS stream
        constexpr uint32_t idc = umma::idesc_tf32(128, 32, 0, 0);
        const uint64_t dbk = umma::desc_k(umma::smem_u32(tiles + 8 * TILE_B / 4));
        for (int ks = 0; ks < 4; ++ks) umma::mma_ts(tmem + 128, tmem + 0 + 8 * ks, dbk + 2 * ks, idc, ks > 0);
        if (ts_first == 2) { umma::commit(&bar_go); umma::mbar_wait(&bar_go, 0); }   // ... and known to be complete
      }
      int ch = 0;
      for (int m = 0; m < n_mma; m += 8, ++ch) {
#pragma unroll
        for (int i = 0; i < 8; ++i) umma::mma_ss(tmem + 192, da + 64 * (i + (m & 8)), db + 64 * (i + (m & 8)), idesc, 1);
        if (paced) {
          umma::commit(&bar_pace[ch & 1]);
          if (ch >= 1) { umma::mbar_wait(&bar_pace[(ch - 1) & 1], ph[(ch - 1) & 1]); ph[(ch - 1) & 1] ^= 1; }
        }
      }
      umma::commit(&bar_done);
      out[N_OPS * N_REP] = clock64() - t0;           // issue time
      umma::mbar_wait(&bar_done, 0);
      out[N_OPS * N_REP + 1] = clock64() - t0;       // completion time
    }
  } else {
    const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
    float loc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) loc[i] = lane + i;
    // let the MMA stream get going
    const long long tstart = clock64();
    while (clock64() - tstart < 600) { }
    float sink = 0.0f;
    for (int op = 0; op < N_OPS; ++op) {
      for (int rep = 0; rep < N_REP; ++rep) {
        __syncwarp();
        const long long t0 = clock64();
        if (op == 0) {          // tcgen05.st x8 x2 + wait
          uint32_t u[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) u[i] = rep + i;
          umma::tmem_st8_raw(lane_base + 0, u);
          umma::tmem_st8_raw(lane_base + 32, u);
          umma::wait_st();
        } else if (op == 1) {   // tcgen05.ld x8 x2 + wait
          float v[8], w[8];
          umma::tmem_ld8_nowait(lane_base + 64, v);
          umma::tmem_ld8_nowait(lane_base + 96, w);
          umma::wait_ld();
          sink += v[0] + w[7];
        } else if (op == 2) {   // STS.128 x4 (conflict-free rows) + fence
          float4* p = reinterpret_cast<float4*>(scratch) + (warp * 32 + lane) * 4;
          p[0] = make_float4(rep, 1, 2, 3); p[1] = make_float4(rep, 1, 2, 3); p[2] = make_float4(rep, 1, 2, 3); p[3] = make_float4(rep, 1, 2, 3);
          umma::fence_async_smem();
        } else if (op == 3) {   // LDS.128 x4, dependent use
          const float4* p = reinterpret_cast<const float4*>(scratch) + (warp * 32 + lane) * 4;
          const float4 a = p[0], b = p[1], c = p[2], d = p[3];
          sink += a.x + b.y + c.z + d.w;
          asm volatile("" :: "f"(sink));
        } else if (op == 4) {   // LDG.128 x2 (L2 resident), dependent use
          const float4* p = reinterpret_cast<const float4*>(gsrc) + ((warp * 32 + lane) * 2 + rep * 512);
          const float4 a = __ldcg(p), b = __ldcg(p + 1);
          sink += a.x + b.y;
          asm volatile("" :: "f"(sink));
        } else if (op == 5) {   // local memory (dynamic index)
          const int idx = (rep + lane) & 15;
          volatile float* lp = loc;
          lp[idx] = sink;
          sink += lp[(idx + 5) & 15];
          asm volatile("" :: "f"(sink));
        } else {                // 64 dependent FFMAs (pure ALU reference)
#pragma unroll
          for (int i = 0; i < 64; ++i) sink = fmaf(sink, 1.0001f, 0.5f);
          asm volatile("" :: "f"(sink));
        }
        const long long t1 = clock64();
        if (tid == 0) out[op * N_REP + rep] = t1 - t0;
      }
    }
    if (sink == 12345.678f) out[N_OPS * N_REP + 2] = 1;
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_free(tmem, 512);
}

int main() {
  long long* d_out; float* d_src;
  CK(cudaMalloc(&d_out, 256 * sizeof(long long)));
  CK(cudaMalloc(&d_src, 1 << 20));
  CK(cudaMemset(d_src, 0, 1 << 20));
  const int smem = 1024 + 10 * TILE_B + 8192;
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const char* names[N_OPS] = {"tcgen05.st x8 x2 + wait::st", "tcgen05.ld x8 x2 + wait::ld", "STS.128 x4 + fence.async", "LDS.128 x4",
                              "LDG.128 x2 (L2)", "local st+ld", "64 dependent FFMA"};
  const int cfg[6][4] = {{0, 0, 0, 0}, {512, 0, 0, 0}, {512, 1, 0, 0}, {512, 0, 1, 0}, {512, 0, 2, 0}, {512, 0, 0, 1}};
  for (int c = 0; c < 6; ++c) {
    CK(cudaMemset(d_out, 0, 256 * sizeof(long long)));
    probe<<<1, 160, smem>>>(d_out, d_src, cfg[c][0], cfg[c][1], cfg[c][2], cfg[c][3]);
    CK(cudaDeviceSynchronize());
    long long h[256];
    CK(cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost));
    printf("== [random data: %d] [TS chain MMA first: %d] %d MMAs in flight (%s): issue %lld cycles, complete %lld cycles\n", cfg[c][3], cfg[c][2], cfg[c][0], cfg[c][1] ? "paced, <= 16 outstanding" : "queued at once",
           h[N_OPS * N_REP], h[N_OPS * N_REP + 1]);
    for (int op = 0; op < N_OPS; ++op) {
      printf("  %-30s:", names[op]);
      for (int r = 0; r < N_REP; ++r) printf(" %5lld", h[op * N_REP + r]);
      printf("\n");
    }
  }
  return 0;
}
