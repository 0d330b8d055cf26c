// umma_probe4.cu -- probes for the round-1 v2 reverse sweep:
//  (1) stacked single-pass weight-gradient MMA: [D1h|D0h|D1l|D0l]^T (M=128, MN-major, 4 tiles) x
//      [Zh|Ah|Zl|Al|X] (N=144, MN-major, 4.5 tiles), K = 128 rows = 16 k-steps: accumulator layout + cycles
//  (2) chain GEMM cost, A from TMEM: 12 x (M128 N32 K8) vs 4 x (M128 N64 K8) + 4 x (M128 N32 K8)
//  (3) two warps per TMEM lane quadrant (warp w and w+4), 16-column tcgen05.st / tcgen05.ld halves
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../neural-jump-ode_b200/csrc -o umma_probe4 umma_probe4.cu
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "njode_umma.cuh"
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); return 2; } } while (0)

constexpr int TILE_F = 128 * 32;

__device__ __forceinline__ void ld16(uint32_t taddr, uint32_t (&u)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
                 "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]) : "r"(taddr));
  umma::wait_ld();
}
__device__ __forceinline__ void st16(uint32_t taddr, const uint32_t (&u)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
               :: "r"(taddr), "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]), "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7]),
                 "r"(u[8]), "r"(u[9]), "r"(u[10]), "r"(u[11]), "r"(u[12]), "r"(u[13]), "r"(u[14]), "r"(u[15]) : "memory");
}

// A: [4][128][32] row-major source tiles, B: [5][128][32]; D out: [128 lanes][256 cols]; T out: cycles
__global__ void __launch_bounds__(256) probe(const float* A, const float* B, float* D, long long* T, float* D3) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  float* At = (float*)base;            // 4 tiles
  float* Bt = At + 4 * TILE_F;         // 5 tiles
  float* Wt = Bt + 5 * TILE_F;         // K-major weight tiles: hi (32 rows) then lo (32 rows), 128B swizzle
  __shared__ uint64_t mbar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 4 * TILE_F; i += 256) { const int t = i / TILE_F, r = (i / 32) % 128, c = i % 32; At[t * TILE_F + umma::swz_mn(r, c)] = A[i]; }
  for (int i = tid; i < 5 * TILE_F; i += 256) { const int t = i / TILE_F, r = (i / 32) % 128, c = i % 32; Bt[t * TILE_F + umma::swz_mn(r, c)] = B[i]; }
  for (int i = tid; i < 64 * 32; i += 256) { const int n = i / 32, k = i % 32; Wt[umma::swz_k(n, k)] = 0.01f * (float)((n * 7 + k * 3) % 17); }
  if (tid == 0) { umma::mbar_init(&mbar, 1); umma::fence_mbar_init(); }
  if (warp == 0) umma::tmem_alloc(&tmem_base_s, 512);
  umma::fence_async_smem(); umma::fence_before_sync(); __syncthreads(); umma::fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t phase = 0;
  // ---- (3) two warps per quadrant write their 16-column half of columns [256, 288), then read back the other half
  {
    uint32_t v[16];
    for (int i = 0; i < 16; ++i) v[i] = __float_as_uint((float)(tid * 100 + i));
    st16(lane_base + 256 + 16 * (warp >> 2), v);
    umma::wait_st(); umma::fence_before_sync(); __syncthreads(); umma::fence_after_sync();
    uint32_t u[16];
    ld16(lane_base + 256 + 16 * (1 - (warp >> 2)), u);
    for (int i = 0; i < 16; ++i) D3[tid * 16 + i] = __uint_as_float(u[i]);
    umma::fence_before_sync(); __syncthreads(); umma::fence_after_sync();
  }
  // ---- (1) stacked wgrad
  constexpr uint32_t idesc_w = umma::idesc_tf32(128, 144, 1, 1);
  for (int rep = 0; rep < 9; ++rep) {
    long long t0 = 0;
    if (tid == 0) {
      t0 = clock64();
      const uint64_t da = umma::desc_mn(umma::smem_u32(At), TILE_F * 4), db = umma::desc_mn(umma::smem_u32(Bt), TILE_F * 4);
      const int nrep = rep < 8 ? 1 : 8;
      for (int q = 0; q < nrep; ++q)
        for (int ks = 0; ks < 16; ++ks) umma::mma_ss(tmem, da + 64 * ks, db + 64 * ks, idesc_w, ks > 0);
      umma::commit(&mbar);
    }
    umma::mbar_wait(&mbar, phase); phase ^= 1;
    umma::fence_after_sync();
    if (tid == 0) T[rep] = clock64() - t0;
    __syncthreads();
  }
  if (warp < 4) {
    for (int c0 = 0; c0 < 144; c0 += 16) {
      uint32_t u[16];
      ld16(lane_base + c0, u);
      for (int i = 0; i < 16; ++i) D[tid * 256 + c0 + i] = __uint_as_float(u[i]);
    }
  }
  umma::fence_before_sync(); __syncthreads(); umma::fence_after_sync();
  // ---- (2) chain GEMM timing, A from TMEM columns [320,352) hi / [352,384) lo, accumulator [384, 448)
  {
    uint32_t v[16];
    for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(0.001f * (float)((tid + i) % 31));
    st16(lane_base + 320 + 16 * (warp >> 2), v);
    st16(lane_base + 352 + 16 * (warp >> 2), v);
    umma::wait_st(); umma::fence_before_sync(); __syncthreads(); umma::fence_after_sync();
    constexpr uint32_t id32 = umma::idesc_tf32(128, 32, 0, 0), id64 = umma::idesc_tf32(128, 64, 0, 0);
    const uint64_t dbh = umma::desc_k(umma::smem_u32(Wt)), dbl = umma::desc_k(umma::smem_u32(Wt + 32 * 32));
    for (int variant = 0; variant < 4; ++variant) {
      long long t0 = 0;
      if (tid == 0) {
        t0 = clock64();
        const int nrep = (variant & 1) ? 8 : 1;
        for (int q = 0; q < nrep; ++q) {
          if (variant < 2) {
            for (int ks = 0; ks < 4; ++ks) umma::mma_ts(tmem + 384, tmem + 352 + 8 * ks, dbh + 2 * ks, id32, ks > 0);
            for (int ks = 0; ks < 4; ++ks) umma::mma_ts(tmem + 384, tmem + 320 + 8 * ks, dbl + 2 * ks, id32, 1);
            for (int ks = 0; ks < 4; ++ks) umma::mma_ts(tmem + 384, tmem + 320 + 8 * ks, dbh + 2 * ks, id32, 1);
          } else {
            for (int ks = 0; ks < 4; ++ks) umma::mma_ts(tmem + 384, tmem + 352 + 8 * ks, dbh + 2 * ks, id32, ks > 0);
            for (int ks = 0; ks < 4; ++ks) umma::mma_ts(tmem + 384, tmem + 320 + 8 * ks, dbh + 2 * ks, id64, 1);
          }
        }
        umma::commit(&mbar);
      }
      umma::mbar_wait(&mbar, phase); phase ^= 1;
      umma::fence_after_sync();
      if (tid == 0) T[16 + variant] = clock64() - t0;
      __syncthreads();
    }
  }
  umma::fence_before_sync(); __syncthreads();
  if (warp == 0) umma::tmem_free(tmem, 512);
}

int main() {
  std::vector<float> A(4 * TILE_F), B(5 * TILE_F);
  srand(4242);
  for (auto& v : A) v = (float)((rand() % 255) - 127) / 64.f;
  for (auto& v : B) v = (float)((rand() % 255) - 127) / 32.f;
  float *dA, *dB, *dD, *dD3; long long* dT;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, 128 * 256 * 4)); CK(cudaMalloc(&dD3, 256 * 16 * 4));
  CK(cudaMalloc(&dT, 64 * 8)); CK(cudaMemset(dD, 0, 128 * 256 * 4)); CK(cudaMemset(dT, 0, 64 * 8));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  const int smem = 9 * TILE_F * 4 + 64 * 32 * 4 + 2048;
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  probe<<<1, 256, smem>>>(dA, dB, dD, dT, dD3);
  CK(cudaDeviceSynchronize());
  std::vector<float> D(128 * 256), D3(256 * 16); std::vector<long long> T(64);
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(D3.data(), dD3, D3.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(T.data(), dT, 64 * 8, cudaMemcpyDeviceToHost));
  // (3)
  int bad3 = 0;
  for (int t = 0; t < 256; ++t) { const int partner = t < 128 ? t + 128 : t - 128; for (int i = 0; i < 16; ++i) if (D3[t * 16 + i] != (float)(partner * 100 + i)) ++bad3; }
  printf("(3) two warps per TMEM quadrant, x16 halves: mismatches %d (expect 0)\n", bad3);
  // (1): expect D[i][n] at lane i, column n (tf32 inputs are exactly representable here)
  int bad = 0; double maxerr = 0;
  for (int i = 0; i < 128; ++i)
    for (int n = 0; n < 144; ++n) {
      const int ta = i / 32, ia = i % 32, tb = n / 32, nb = n % 32;
      double ref = 0;
      for (int r = 0; r < 128; ++r) ref += (double)A[ta * TILE_F + r * 32 + ia] * B[tb * TILE_F + r * 32 + nb];
      const double err = fabs(D[i * 256 + n] - ref);
      if (err > maxerr) maxerr = err;
      if (err > 1e-2) { if (bad < 5) printf("  D[%d][%d] = %f expected %f\n", i, n, D[i * 256 + n], ref); ++bad; }
    }
  printf("(1) stacked wgrad M128 N144 K128 MNxMN: bad cells %d of %d, max abs err %.3g\n", bad, 128 * 144, maxerr);
  printf("    cycles issue..complete, 16 MMAs: "); for (int r = 0; r < 8; ++r) printf("%lld ", T[r]); printf("| 8 x 16 MMAs: %lld (%.1f per MMA)\n", T[8], T[8] / 128.0);
  printf("(2) chain GEMM A from TMEM: 12 x N32: %lld cycles, x8: %lld (%.1f per GEMM) | 4 x N32 + 4 x N64: %lld, x8: %lld (%.1f per GEMM)\n",
         T[16], T[17], T[17] / 8.0, T[18], T[19], T[19] / 8.0);
  return bad || bad3;
}
