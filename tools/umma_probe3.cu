// umma_probe3.cu -- where does an M=64 accumulator live in TMEM?  One MN x MN MMA chain (M=64,N=64,K=128)
// over [D1|D0] x [Z|A] tiles; dumps 128 lanes x 64 columns and locates every expected D[i][n].
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); return 2; } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF); d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16; d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46; d |= (uint64_t)layout << 61;
  return d;
}
__device__ __host__ __forceinline__ int swz32(int r, int c) { return r * 32 + (((c >> 3) ^ (r & 3)) << 3) + (c & 7); }
__global__ void __launch_bounds__(128) probe(const float* A, const float* B, float* D, int M, int N) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  float* At = (float*)base; float* Bt = At + 2 * 4096;
  __shared__ uint64_t mbar; __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 2 * 4096; i += 128) { const int t = i / 4096, r = (i / 32) % 128, c = i % 32; At[t * 4096 + swz32(r, c)] = A[i]; Bt[t * 4096 + swz32(r, c)] = B[i]; }
  if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&mbar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (warp == 0) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(128u)); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;"); }
  asm volatile("fence.proxy.async.shared::cta;"); asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads(); asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_base_s;
  // zero the 128 x 128 TMEM window first so untouched cells read as 0
  { uint32_t z[32]; for (int c = 0; c < 32; ++c) z[c] = 0x7fc00000u;   // NaN marker = "not written"
    for (int c0 = 0; c0 < 128; c0 += 32) {
      const uint32_t ta = tmem + ((uint32_t)(warp * 32) << 16) + c0;
      asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
        :: "r"(ta), "r"(z[0]), "r"(z[1]), "r"(z[2]), "r"(z[3]), "r"(z[4]), "r"(z[5]), "r"(z[6]), "r"(z[7]), "r"(z[8]), "r"(z[9]), "r"(z[10]), "r"(z[11]), "r"(z[12]), "r"(z[13]), "r"(z[14]), "r"(z[15]),
           "r"(z[16]), "r"(z[17]), "r"(z[18]), "r"(z[19]), "r"(z[20]), "r"(z[21]), "r"(z[22]), "r"(z[23]), "r"(z[24]), "r"(z[25]), "r"(z[26]), "r"(z[27]), "r"(z[28]), "r"(z[29]), "r"(z[30]), "r"(z[31]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads(); asm volatile("tcgen05.fence::after_thread_sync;"); }
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    for (int ks = 0; ks < 16; ++ks) {
      const uint64_t da = make_desc(smem_u32(At) + ks * 1024, 16384, 512, 1), db = make_desc(smem_u32(Bt) + ks * 1024, 16384, 512, 1);
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" :: "r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(ks ? 1u : 0u) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&mbar)) : "memory");
  }
  uint32_t done = 0;
  for (int spin = 0; spin < (1 << 22) && !done; ++spin)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0u) : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;");
  for (int c0 = 0; c0 < 128; c0 += 32) {
    uint32_t v[32];
    const uint32_t ta = tmem + ((uint32_t)(warp * 32) << 16) + c0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]) : "r"(ta));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int c = 0; c < 32; ++c) D[tid * 128 + c0 + c] = __uint_as_float(v[c]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(128u));
}
int main(int argc, char** argv) {
  const int M = argc > 1 ? atoi(argv[1]) : 64, N = argc > 2 ? atoi(argv[2]) : 64;
  const int n = 2 * 4096;
  std::vector<float> A(n), B(n);
  srand(777);
  // make every D[i][n] unique: A[r][m] nonzero pattern with distinct values
  for (int i = 0; i < n; ++i) { A[i] = (float)((rand() % 255) - 127) / 64.f; B[i] = (float)((rand() % 255) - 127) / 32.f; }
  float *dA, *dB, *dD;
  CK(cudaMalloc(&dA, n * 4)); CK(cudaMalloc(&dB, n * 4)); CK(cudaMalloc(&dD, 128 * 128 * 4));
  CK(cudaMemcpy(dA, A.data(), n * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), n * 4, cudaMemcpyHostToDevice));
  const int smem = 4 * 16384 + 2048;
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  probe<<<1, 128, smem>>>(dA, dB, dD, M, N);
  CK(cudaDeviceSynchronize());
  std::vector<float> D(128 * 128);
  CK(cudaMemcpy(D.data(), dD, 128 * 128 * 4, cudaMemcpyDeviceToHost));
  int written = 0;
  for (float v : D) if (v == v) ++written;
  printf("M=%d N=%d: %d TMEM cells written (expect %d)\n", M, N, written, M * N);
  // locate D[i][c]
  int shown = 0, found_all = 1;
  for (int i = 0; i < M; ++i)
    for (int c = 0; c < N; ++c) {
      const int ti = i / 32, ii = i % 32, tc = c / 32, cc = c % 32;
      double ref = 0;
      for (int r = 0; r < 128; ++r) ref += (double)A[ti * 4096 + r * 32 + ii] * B[tc * 4096 + r * 32 + cc];
      int fl = -1, fc = -1, cnt = 0;
      for (int l = 0; l < 128; ++l) for (int k = 0; k < 128; ++k) if (fabs(D[l * 128 + k] - ref) < 1e-3) { if (!cnt) { fl = l; fc = k; } ++cnt; }
      if (cnt == 0) found_all = 0;
      if ((c == 0 || c == 1 || c == 33) && (i % 8 == 0 || i == 1 || i == 17 || i == 33) && shown < 40) { printf("  D[%2d][%2d] -> lane %3d col %3d (matches %d)\n", i, c, fl, fc, cnt); ++shown; }
    }
  printf("all found: %d\n", found_all);
  return 0;
}
