// umma_probe.cu -- validates the tcgen05 (UMMA) building blocks used by njode_tiled.cu on a real B200:
//   * shared-memory matrix descriptors for 128B-swizzled fp32/tf32 tiles, K-major and MN-major
//   * the kind::tf32 instruction descriptor, K-advance inside the swizzle atom, accumulate flag
//   * TMEM allocation, tcgen05.commit -> mbarrier, tcgen05.ld 32x32b.x32
//   * the 3xTF32 error-compensated split (hi*hi + hi*lo + lo*hi) against an fp64 reference
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/umma_probe tools/umma_probe.cu
// Run:   tools/umma_probe <test> [a_lbo a_sbo b_lbo b_sbo]      (byte offsets; each run = one variant)
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); return 2; } } while (0)

struct Params {
  int test;            // 1: K-major x K-major (M=128,N=32,K=32)  2: same, 3xTF32 split  3: MN x MN wgrad (K=128)
  uint32_t a_lbo, a_sbo, b_lbo, b_sbo;   // bytes
  int a_major, b_major;                  // 0 = K, 1 = MN
  int M, N, ksteps;
  uint32_t a_kstep_bytes, b_kstep_bytes;
  int n_split;                           // 1 or 3
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout = 2) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;                 // descriptor version 1 (sm_100)
  d |= (uint64_t)layout << 61;            // 2 = SWIZZLE_128B (K-major), 1 = SWIZZLE_128B_BASE32B (MN-major tf32)
  return d;
}

__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// element (r, c) of a [rows][32] fp32 tile with 128-byte rows and the 128B swizzle (16-byte chunk ^= row % 8)
__device__ __host__ __forceinline__ int swz(int r, int c) { return r * 32 + (((c >> 2) ^ (r & 7)) << 2) + (c & 3); }
// MN-major tf32 tiles: 128-byte rows, 32-byte chunk ^= row % 4 (SWIZZLE_128B_BASE32B, 4-row atoms)
__device__ __host__ __forceinline__ int swz32(int r, int c) { return r * 32 + (((c >> 3) ^ (r & 3)) << 3) + (c & 7); }

__global__ void __launch_bounds__(128) probe(Params P, const float* A, const float* B, float* D, int* status) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: 1024-aligned tiles of 16 KB: A_hi[4 tiles], A_lo[4], B_hi[2], B_lo[2]
  uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  float* Ah = (float*)base;
  float* Al = Ah + 4 * 4096;
  float* Bh = Al + 4 * 4096;
  float* Bl = Bh + 2 * 4096;
  __shared__ uint64_t mbar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;

  // fill: A is [4 tiles][128][32], B likewise (tests use what they need)
  for (int i = tid; i < 4 * 128 * 32; i += 128) {
    const int t = i / 4096, r = (i / 32) % 128, c = i % 32;
    const float a = A[i], b = B[i];
    const float ah = P.n_split == 3 ? tf32_rna(a) : a, bh = P.n_split == 3 ? tf32_rna(b) : b;
    const int e = P.a_major ? swz32(r, c) : swz(r, c);
    Ah[t * 4096 + e] = ah;
    Al[t * 4096 + e] = P.n_split == 3 ? tf32_rna(a - ah) : 0.f;
    if (t < 2) {
      Bh[t * 4096 + e] = bh;
      Bl[t * 4096 + e] = P.n_split == 3 ? tf32_rna(b - bh) : 0.f;
    }
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(128u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;");      // generic-proxy smem writes -> visible to the async proxy (UMMA)
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_base_s;

  if (tid == 0) {
    uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)P.a_major << 15) | ((uint32_t)P.b_major << 16) |
                     ((uint32_t)(P.N >> 3) << 17) | ((uint32_t)(P.M >> 4) << 24);
    int first = 1;
    for (int sp = 0; sp < P.n_split; ++sp) {
      // split order: (Al,Bh), (Ah,Bl), (Ah,Bh)  -- small terms first
      const float* a = (P.n_split == 1) ? Ah : (sp == 0 ? Al : Ah);
      const float* b = (P.n_split == 1) ? Bh : (sp == 1 ? Bl : Bh);
      for (int ks = 0; ks < P.ksteps; ++ks) {
        const uint64_t da = make_desc(smem_u32(a) + ks * P.a_kstep_bytes, P.a_lbo, P.a_sbo, P.a_major ? 1 : 2);
        const uint64_t db = make_desc(smem_u32(b) + ks * P.b_kstep_bytes, P.b_lbo, P.b_sbo, P.b_major ? 1 : 2);
        const uint32_t acc = first ? 0u : 1u;
        first = 0;
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                     :: "r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&mbar)) : "memory");
  }
  // wait (bounded spin)
  uint32_t done = 0;
  for (int spin = 0; spin < (1 << 22) && !done; ++spin) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0u) : "memory");
  }
  if (!done) { if (tid == 0) status[0] = -1; }
  asm volatile("tcgen05.fence::after_thread_sync;");
  if (done) {
    for (int c0 = 0; c0 < P.N; c0 += 32) {
      uint32_t v[32];
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                   "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                     "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                     "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                     "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                   : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int c = 0; c < 32; ++c) D[tid * P.N + c0 + c] = __uint_as_float(v[c]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(128u));
  if (tid == 0 && done) status[0] = 1;
}

int main(int argc, char** argv) {
  Params P = {};
  P.test = argc > 1 ? atoi(argv[1]) : 1;
  P.M = 128; P.N = 32; P.n_split = 1;
  if (P.test == 1 || P.test == 2) {            // D[r][n] = sum_k A[r][k] B[n][k]
    P.a_major = P.b_major = 0; P.ksteps = 4; P.a_kstep_bytes = P.b_kstep_bytes = 32;
    P.a_lbo = 16; P.a_sbo = 1024; P.b_lbo = 16; P.b_sbo = 1024;
    if (P.test == 2) P.n_split = 3;
  } else if (P.test == 3 || P.test == 4) {     // D[j][k] = sum_r A[r][j] B[r][k], r = 0..127
    P.a_major = P.b_major = 1; P.ksteps = 16; P.a_kstep_bytes = P.b_kstep_bytes = 1024;
    P.a_lbo = 16384; P.a_sbo = 512; P.b_lbo = 16384; P.b_sbo = 512;   // M=128: 4 tiles of 32 j; 4-row atoms
    if (P.test == 4) { P.N = 64; P.n_split = 3; }
  }
  if (argc > 5) { P.a_lbo = atoi(argv[2]); P.a_sbo = atoi(argv[3]); P.b_lbo = atoi(argv[4]); P.b_sbo = atoi(argv[5]); }

  const int n = 4 * 128 * 32;
  std::vector<float> A(n), B(n);
  srand(12345);
  for (int i = 0; i < n; ++i) {
    if (P.n_split == 3) { A[i] = (float)rand() / RAND_MAX * 2.f - 1.f; B[i] = (float)rand() / RAND_MAX * 2.f - 1.f; }
    else { A[i] = (float)((rand() % 17) - 8) * 0.125f; B[i] = (float)((rand() % 17) - 8) * 0.25f; }   // tf32-exact
  }
  float *dA, *dB, *dD; int* dS;
  CK(cudaMalloc(&dA, n * 4)); CK(cudaMalloc(&dB, n * 4)); CK(cudaMalloc(&dD, 128 * 256 * 4)); CK(cudaMalloc(&dS, 4));
  CK(cudaMemcpy(dA, A.data(), n * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), n * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0, 128 * 256 * 4)); CK(cudaMemset(dS, 0, 4));
  const int smem = 12 * 16384 + 2048;
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  probe<<<1, 128, smem>>>(P, dA, dB, dD, dS);
  CK(cudaDeviceSynchronize());
  std::vector<float> D(128 * P.N); int st = 0;
  CK(cudaMemcpy(D.data(), dD, 128 * P.N * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost));
  if (st != 1) { printf("test %d: status %d (mbarrier timeout)\n", P.test, st); return 1; }
  double maxerr = 0, maxref = 0; int bad = 0;
  for (int m = 0; m < 128; ++m)
    for (int c = 0; c < P.N; ++c) {
      double ref = 0;
      if (P.test <= 2) { for (int k = 0; k < 32; ++k) ref += (double)A[m * 32 + k] * B[c * 32 + k]; }
      else {           // A element (tile tj, row r, col j): M index = tj*32 + j ; B: N index = tk*32 + k
        const int tj = m / 32, j = m % 32, tk = c / 32, kk = c % 32;
        for (int r = 0; r < 128; ++r) ref += (double)A[tj * 4096 + r * 32 + j] * B[tk * 4096 + r * 32 + kk];
      }
      const double err = fabs(ref - D[m * P.N + c]);
      if (err > maxerr) maxerr = err;
      if (fabs(ref) > maxref) maxref = fabs(ref);
      if (err > 1e-3 * (1 + fabs(ref)) && bad < 6) { printf("  mismatch D[%d][%d] = %g, ref %g\n", m, c, D[m * P.N + c], ref); ++bad; }
    }
  printf("test %d (a_lbo=%u a_sbo=%u b_lbo=%u b_sbo=%u split=%d): max abs err %.3e, max |ref| %.3e, rel %.3e -> %s\n", P.test,
         P.a_lbo, P.a_sbo, P.b_lbo, P.b_sbo, P.n_split, maxerr, maxref, maxerr / maxref, maxerr <= 2e-6 * maxref * (P.n_split == 3 ? 1 : 1) + (P.n_split == 3 ? 0 : 1e-12) ? "OK" : "FAIL");
  return 0;
}
