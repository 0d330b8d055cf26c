// umma_probe2.cu -- (a) correctness of A-from-TMEM (TS) tf32 MMA, (b) issue-to-completion timing of the
// small-N MMA shapes the NJ-ODE sweep kernels use, to decide between smem (SS) and TMEM (TS) operands.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/umma_probe2 tools/umma_probe2.cu
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); return 2; } } while (0)

struct Params {
  int mode;            // 0 = TS correctness, 1 = timing
  int ts;              // A from TMEM?
  int a_major, b_major;
  int M, N, ksteps, reps;
  uint32_t a_lbo, a_sbo, b_lbo, b_sbo, a_kstep, b_kstep;   // bytes (a_kstep = columns in TS mode)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}
__device__ __host__ __forceinline__ int swz(int r, int c) { return r * 32 + (((c >> 2) ^ (r & 7)) << 2) + (c & 3); }
__device__ __host__ __forceinline__ int swz32(int r, int c) { return r * 32 + (((c >> 3) ^ (r & 3)) << 3) + (c & 7); }

#define TMEM_LD32(taddr, v) \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 " \
               "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];" \
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), \
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), \
                 "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), \
                 "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]) \
               : "r"(taddr))
#define TMEM_ST32(taddr, v) \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], " \
               "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" \
               :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), \
                 "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), \
                 "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), \
                 "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31]) : "memory")

__global__ void __launch_bounds__(128) probe(Params P, const float* A, const float* B, float* D, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  float* At = (float*)base;          // 4 tiles
  float* Bt = At + 4 * 4096;         // 2 tiles
  __shared__ uint64_t mbar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 4 * 128 * 32; i += 128) {
    const int t = i / 4096, r = (i / 32) % 128, c = i % 32;
    const int e = P.a_major ? swz32(r, c) : swz(r, c);
    At[t * 4096 + e] = A[i];
    if (t < 2) Bt[t * 4096 + (P.b_major ? swz32(r, c) : swz(r, c))] = B[i];
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_base_s;
  const uint32_t colA = 256;                         // A operand columns in TS mode
  if (P.ts) {                                        // thread r stores row r of A (tile 0) into TMEM lane r
    uint32_t v[32];
    for (int c = 0; c < 32; ++c) v[c] = __float_as_uint(A[tid * 32 + c]);
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + colA;
    TMEM_ST32(taddr, v);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
  }
  long long t0 = 0, t1 = 0;
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)P.a_major << 15) | ((uint32_t)P.b_major << 16) |
                           ((uint32_t)(P.N >> 3) << 17) | ((uint32_t)(P.M >> 4) << 24);
    t0 = clock64();
    int first = 1;
    for (int rep = 0; rep < P.reps; ++rep) {
      for (int ks = 0; ks < P.ksteps; ++ks) {
        const uint64_t db = make_desc(smem_u32(Bt) + ks * P.b_kstep, P.b_lbo, P.b_sbo, P.b_major ? 1 : 2);
        const uint32_t acc = first ? 0u : 1u;
        first = 0;
        if (P.ts) {
          const uint32_t ta = tmem + colA + ks * P.a_kstep;
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                       "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
                       :: "r"(tmem), "r"(ta), "l"(db), "r"(idesc), "r"(acc) : "memory");
        } else {
          const uint64_t da = make_desc(smem_u32(At) + ks * P.a_kstep, P.a_lbo, P.a_sbo, P.a_major ? 1 : 2);
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                       "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                       :: "r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
        }
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&mbar)) : "memory");
  }
  uint32_t done = 0;
  for (int spin = 0; spin < (1 << 24) && !done; ++spin) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0u) : "memory");
  }
  if (tid == 0) { t1 = clock64(); out[0] = done ? (t1 - t0) : -1; }
  asm volatile("tcgen05.fence::after_thread_sync;");
  if (done && P.mode == 0) {
    uint32_t v[32];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
    TMEM_LD32(taddr, v);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int c = 0; c < 32; ++c) D[tid * 32 + c] = __uint_as_float(v[c]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512u));
}


// lean issue loop: descriptors precomputed, k-steps fully unrolled with constant offsets
template <int KS, int TS>
__global__ void __launch_bounds__(128) probe_lean(Params P, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  float* At = (float*)base;
  float* Bt = At + 4 * 4096;
  __shared__ uint64_t mbar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 6 * 4096; i += 128) At[i] = 0.25f;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_base_s;
  long long t0 = 0;
  if (warp == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)P.a_major << 15) | ((uint32_t)P.b_major << 16) |
                           ((uint32_t)(P.N >> 3) << 17) | ((uint32_t)(P.M >> 4) << 24);
    const uint64_t da0 = make_desc(smem_u32(At), P.a_lbo, P.a_sbo, P.a_major ? 1 : 2);
    const uint64_t db0 = make_desc(smem_u32(Bt), P.b_lbo, P.b_sbo, P.b_major ? 1 : 2);
    const uint32_t ta0 = tmem + 256;
    const uint32_t astep = P.a_kstep, bstep = P.b_kstep >> 4;
    uint32_t elected = 0;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(elected));
    if (elected) {
      t0 = clock64();
      for (int rep = 0; rep < P.reps; ++rep) {
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          const uint64_t db = db0 + (uint64_t)(ks * bstep);
          if (TS) {
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
                         :: "r"(tmem), "r"(ta0 + ks * astep), "l"(db), "r"(idesc), "r"(1u));
          } else {
            const uint64_t da = da0 + (uint64_t)(ks * (astep >> 4));
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                         :: "r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(1u));
          }
        }
      }
      const long long t_issue = clock64();
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&mbar)) : "memory");
      out[1] = t_issue - t0;
    }
  }
  uint32_t done = 0;
  for (int spin = 0; spin < (1 << 24) && !done; ++spin) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0u) : "memory");
  }
  if (t0) out[0] = done ? (clock64() - t0) : -1;
  asm volatile("tcgen05.fence::after_thread_sync;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512u));
}

template <int KS, int TS>
static int run_lean(Params P, const char* name, long long* dO) {
  const int smem = 6 * 16384 + 2048;
  CK(cudaFuncSetAttribute(probe_lean<KS, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  probe_lean<KS, TS><<<1, 128, smem>>>(P, dO);
  CK(cudaDeviceSynchronize());
  long long cyc[2] = {0, 0};
  CK(cudaMemcpy(cyc, dO, 16, cudaMemcpyDeviceToHost));
  const int n_mma = P.reps * KS;
  printf("LEAN %-40s %6d MMAs: total %8lld (%6.2f/MMA)  issue-only %8lld (%6.2f/MMA)\n", name, n_mma, cyc[0],
         (double)cyc[0] / n_mma, cyc[1], (double)cyc[1] / n_mma);
  return 0;
}

static int run(Params P, const char* name, const float* dA, const float* dB, float* dD, long long* dO,
               const std::vector<float>& A, const std::vector<float>& B) {
  const int smem = 6 * 16384 + 2048;
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  probe<<<1, 128, smem>>>(P, dA, dB, dD, dO);
  CK(cudaDeviceSynchronize());
  long long cyc = 0;
  CK(cudaMemcpy(&cyc, dO, 8, cudaMemcpyDeviceToHost));
  if (P.mode == 0) {
    std::vector<float> D(128 * 32);
    CK(cudaMemcpy(D.data(), dD, 128 * 32 * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0;
    for (int m = 0; m < 128; ++m)
      for (int c = 0; c < 32; ++c) {
        double ref = 0;
        for (int k = 0; k < 32; ++k) ref += (double)A[m * 32 + k] * B[c * 32 + k];
        maxerr = fmax(maxerr, fabs(ref - D[m * 32 + c]));
      }
    printf("%-44s correctness: max abs err %.3e (cycles %lld)\n", name, maxerr, cyc);
  } else {
    const int n_mma = P.reps * P.ksteps;
    printf("%-44s %6d MMAs: %8lld cycles -> %7.2f cycles/MMA\n", name, n_mma, cyc, (double)cyc / n_mma);
  }
  return 0;
}

int main() {
  const int n = 4 * 128 * 32;
  std::vector<float> A(n), B(n);
  srand(12345);
  for (int i = 0; i < n; ++i) { A[i] = (float)((rand() % 17) - 8) * 0.125f; B[i] = (float)((rand() % 17) - 8) * 0.25f; }
  float *dA, *dB, *dD; long long* dO;
  CK(cudaMalloc(&dA, n * 4)); CK(cudaMalloc(&dB, n * 4)); CK(cudaMalloc(&dD, 128 * 64 * 4)); CK(cudaMalloc(&dO, 16));
  CK(cudaMemcpy(dA, A.data(), n * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), n * 4, cudaMemcpyHostToDevice));
  Params P = {};
  // TS correctness
  P = Params{0, 1, 0, 0, 128, 32, 4, 1, 16, 1024, 16, 1024, 8, 32};
  run(P, "TS  K-major M128 N32 K32", dA, dB, dD, dO, A, B);
  P = Params{0, 0, 0, 0, 128, 32, 4, 1, 16, 1024, 16, 1024, 32, 32};
  run(P, "SS  K-major M128 N32 K32", dA, dB, dD, dO, A, B);
  // timing
  const int R = 256;
  P = Params{1, 0, 0, 0, 128, 32, 4, R, 16, 1024, 16, 1024, 32, 32};   run(P, "SS  K x K    M128 N32 (chain)", dA, dB, dD, dO, A, B);
  P = Params{1, 1, 0, 0, 128, 32, 4, R, 16, 1024, 16, 1024, 8, 32};    run(P, "TS  K x K    M128 N32 (chain, A in TMEM)", dA, dB, dD, dO, A, B);
  P = Params{1, 0, 0, 0, 64, 32, 4, R, 16, 1024, 16, 1024, 32, 32};    run(P, "SS  K x K    M64  N32", dA, dB, dD, dO, A, B);
  P = Params{1, 1, 0, 0, 64, 32, 4, R, 16, 1024, 16, 1024, 8, 32};     run(P, "TS  K x K    M64  N32", dA, dB, dD, dO, A, B);
  P = Params{1, 0, 1, 1, 128, 32, 16, R / 4, 16384, 512, 16384, 512, 1024, 1024}; run(P, "SS  MN x MN  M128 N32 (wgrad, 4 tiles)", dA, dB, dD, dO, A, B);
  P = Params{1, 0, 1, 1, 128, 32, 16, R / 4, 0, 512, 16384, 512, 1024, 1024};     run(P, "SS  MN x MN  M128 N32 (wgrad, LBO=0 alias)", dA, dB, dD, dO, A, B);
  P = Params{1, 0, 1, 1, 64, 32, 16, R / 4, 16384, 512, 16384, 512, 1024, 1024};  run(P, "SS  MN x MN  M64  N32 (wgrad)", dA, dB, dD, dO, A, B);
  P = Params{1, 0, 1, 1, 64, 64, 16, R / 4, 16384, 512, 16384, 512, 1024, 1024};  run(P, "SS  MN x MN  M64  N64 (2 wgrads stacked)", dA, dB, dD, dO, A, B);
  P = Params{1, 0, 1, 1, 128, 64, 16, R / 4, 16384, 512, 16384, 512, 1024, 1024}; run(P, "SS  MN x MN  M128 N64", dA, dB, dD, dO, A, B);
  P = Params{1, 0, 1, 1, 64, 16, 16, R / 4, 16384, 512, 16384, 512, 1024, 1024};  run(P, "SS  MN x MN  M64  N16", dA, dB, dD, dO, A, B);
  P = Params{1, 0, 1, 1, 64, 8, 16, R / 4, 16384, 512, 16384, 512, 1024, 1024};   run(P, "SS  MN x MN  M64  N8", dA, dB, dD, dO, A, B);
  P = Params{1, 0, 0, 0, 128, 64, 4, R, 16, 1024, 16, 1024, 32, 32};   run(P, "SS  K x K    M128 N64", dA, dB, dD, dO, A, B);
  P = Params{1, 1, 0, 0, 128, 64, 4, R, 16, 1024, 16, 1024, 8, 32};    run(P, "TS  K x K    M128 N64", dA, dB, dD, dO, A, B);
  P = Params{1, 0, 0, 0, 128, 128, 4, R, 16, 1024, 16, 1024, 32, 32};  run(P, "SS  K x K    M128 N128", dA, dB, dD, dO, A, B);
  P = Params{1, 1, 0, 0, 128, 128, 4, R, 16, 1024, 16, 1024, 8, 32};   run(P, "TS  K x K    M128 N128", dA, dB, dD, dO, A, B);
  P = Params{1, 0, 0, 0, 128, 256, 4, R, 16, 1024, 16, 1024, 32, 32};  run(P, "SS  K x K    M128 N256", dA, dB, dD, dO, A, B);
  // lean issue
  P = Params{1, 0, 0, 0, 128, 32, 4, R, 16, 1024, 16, 1024, 32, 32};   run_lean<4, 0>(P, "SS K x K M128 N32", dO);
  P = Params{1, 1, 0, 0, 128, 32, 4, R, 16, 1024, 16, 1024, 8, 32};    run_lean<4, 1>(P, "TS K x K M128 N32", dO);
  P = Params{1, 1, 0, 0, 128, 64, 4, R, 16, 1024, 16, 1024, 8, 32};    run_lean<4, 1>(P, "TS K x K M128 N64", dO);
  P = Params{1, 1, 0, 0, 128, 128, 4, R, 16, 1024, 16, 1024, 8, 32};   run_lean<4, 1>(P, "TS K x K M128 N128", dO);
  P = Params{1, 0, 0, 0, 128, 128, 4, R, 16, 1024, 16, 1024, 32, 32};  run_lean<4, 0>(P, "SS K x K M128 N128", dO);
  P = Params{1, 0, 0, 0, 128, 256, 4, R, 16, 1024, 16, 1024, 32, 32};  run_lean<4, 0>(P, "SS K x K M128 N256", dO);
  P = Params{1, 0, 1, 1, 128, 32, 16, R / 4, 16384, 512, 16384, 512, 1024, 1024}; run_lean<16, 0>(P, "SS MN x MN M128 N32 (wgrad)", dO);
  P = Params{1, 0, 1, 1, 128, 64, 16, R / 4, 16384, 512, 16384, 512, 1024, 1024}; run_lean<16, 0>(P, "SS MN x MN M128 N64 (wgrad x2)", dO);
  P = Params{1, 0, 1, 1, 64, 32, 16, R / 4, 16384, 512, 16384, 512, 1024, 1024};  run_lean<16, 0>(P, "SS MN x MN M64 N32 (wgrad)", dO);
  P = Params{1, 0, 1, 1, 128, 128, 16, R / 4, 16384, 512, 16384, 512, 1024, 1024}; run_lean<16, 0>(P, "SS MN x MN M128 N128", dO);
  return 0;
}
