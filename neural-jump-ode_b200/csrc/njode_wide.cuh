// njode_wide.cuh -- geometry and small device helpers shared by the "wide" tcgen05 kernels
// (njode_wide.cu: forward / reverse chain sweeps, njode_wgrad.cu: weight-gradient GEMM) for
// hidden_dim in {64, 128} and up to 3 hidden layers (BASELINE configs 4 and 5).
//
// Checkpoint geometry of this flavour.  A tile of 128 observation units owns (kmax + 1 + NJODE_WIDE_XSLOTS)
// checkpoint slots; a slot is (L + 1) planes of H x 128 floats.  The buffer holds two halves of equal size (plane
// layouts: aplane_off / dplane_off below):
//   half A (written by the forward sweep)              half D (written by the reverse sweep)
//   slot k <= kmax : plane 0 = h before Euler step k   slot k < kmax: plane l = d loss / d (pre-activation
//                    (after the last one for k = kmax)                of ODE layer l) of step k, l = 0..L
//                    plane l = output of ODE layer l-1
//   slot X1 = kmax+1 (readout at h_0), X2 = kmax+2 (readout at h_end):
//                    plane l (1..L) = output of out-net       plane l (0..L-1) = d / d (pre-activation of
//                    layer l-1                                out-net layer l); plane L = a copy of half A's
//                                                             plane L (written by the FORWARD sweep)
//   slot X3 = kmax+3 (jump net): plane l (0..L-1) = output    plane l (0..L) = d / d (pre-activation of jump
//                    of jump layer l                          layer l)
// The planes of half D are laid out for the weight-gradient kernel (njode_wgrad.cu), which contracts pairs (D plane,
// A plane) over rows with the D plane as its TMEM operand, lane = feature, 8 rows per thread and stage:
//     [16 row octets][H features][8 rows]
// so that a warp there (32 features, one octet) reads 1 KB contiguous.  (A plain [feature][row] layout made each of its
// LDG.256 touch 32 different lines, and the LSU time of those -- 80 wavefronts per warp and stage -- ate what taking
// the operand out of shared memory had saved.)  A row worker of the sweeps owns (row, 8 features): 8 scalar stores,
// each 4 x 32 contiguous bytes across the warp's 32 rows.
// A third region holds 8 aux floats per row and slot (the extra B columns of that GEMM).
#pragma once
#include <cstdlib>

#include "njode_common.cuh"
#include "njode_umma.cuh"

#define NJODE_WIDE_XSLOTS 3
#define NJODE_WIDE_LMAX 3

namespace wide {

constexpr int R = 128;                 // rows (observation units) per tile
constexpr int NWARP_W = 16;            // worker warps: warp w serves TMEM lane quadrant w % 4 and column group w / 4
constexpr int NT_W = NWARP_W * 32;     // 512 row workers
constexpr int MAX_DX = 2, MAX_O = 4;

template <int HW>
struct Cfg {
  static constexpr int CG = HW / 4;                   // columns per column group (one group per worker warp of a quadrant)
  static constexpr int NSUB = CG / 8;                 // sub-steps of 8 columns: the hand-over granularity of a layer
  static constexpr int STAGE_HALF = HW * 128;         // bytes of the hi (or lo) part of a weight stage: HW rows x 32 k
  static constexpr int STAGE_BYTES = 2 * STAGE_HALF;  // a stage = the 4 k-steps (one per column group) of one sub-step
  static constexpr int NSTAGE = 6;
  // Blocked GEMM order (NJODE_WIDE_BLOCKED): a chain GEMM runs as NBLK output blocks of CG columns -- block n is the
  // accumulator slice of column group n -- each over the whole K; a ring stage is one block's weights
  // ([hi | lo] x [HW / 32 k-chunks][CG rows][32 k]).
  static constexpr int NBLK = 4;
  static constexpr int BSTAGE_HALF = HW * CG * 4;     // bytes
  static constexpr int BSTAGE_BYTES = 2 * BSTAGE_HALF;
  static constexpr uint32_t A_HI = 0, A_LO = HW, ACC0 = 2 * HW, TMEM_COLS = 4 * HW;
  static constexpr int PL = R * HW;                   // floats per checkpoint plane
};

// matrices that run as chain GEMMs, in image order: jump layers 1..L, ODE layers 0..L, out-net layers 0..L-1
__host__ __device__ __forceinline__ int mat_jump(int L, int l) { (void)L; return l - 1; }
__host__ __device__ __forceinline__ int mat_ode(int L, int l) { return L + l; }
__host__ __device__ __forceinline__ int mat_out(int L, int l) { return 2 * L + 1 + l; }
__host__ __device__ __forceinline__ int n_mats(int L) { return 3 * L + 1; }

// ---- memory helpers (explicit address spaces: through plain pointers the compiler falls back to generic LD/ST) ----
__device__ __forceinline__ void ld8s(const float* __restrict__ src, float (&v)[8]) {     // 32-byte aligned, SHARED
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(src);
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(a));
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4+16];" : "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "r"(a));
}
__device__ __forceinline__ void ld8g(const float* __restrict__ src, float (&v)[8]) {     // one LDG.256, no L1 allocation
  asm volatile("ld.global.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(src));
}
__device__ __forceinline__ void st8g(float* __restrict__ dst, const float (&v)[8]) {     // one STG.256
  asm volatile("st.global.v8.f32 [%8], {%0,%1,%2,%3,%4,%5,%6,%7};"
               :: "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "l"(dst) : "memory");
}
// 8 consecutive features of one row into a half-D plane ([row octet][feature][8 rows]); dst points at (first feature, row)
__device__ __forceinline__ void st8t(float* __restrict__ dst, const float (&v)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) asm volatile("st.global.f32 [%0], %1;" :: "l"(dst + i * 8), "f"(v[i]) : "memory");
}
// offset of (8-feature chunk c, row r) inside a half-A plane of width HW: [4 row groups][HW/8 chunks][32 rows][8 floats]
// -- a warp moving one chunk of its 32 rows touches 1 KB contiguous, and the 32 rows of a weight-gradient stage (all
// chunks) are ONE contiguous block of HW * 128 bytes, i.e. one bulk copy
__host__ __device__ __forceinline__ int aplane_off(int HW, int c, int r) { return (((r >> 5) * (HW >> 3) + c) * 32 + (r & 31)) * 8; }
// offset of (feature f, row r) inside a half-D plane of width HW
__host__ __device__ __forceinline__ int dplane_off(int HW, int f, int r) { return (r >> 3) * (HW * 8) + f * 8 + (r & 7); }
__device__ __forceinline__ float ldg_na(const float* __restrict__ p) {
  float v;
  asm volatile("ld.global.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }

// ---- mbarrier / bulk-copy (TMA engine, 1-D) helpers ------------------------------------------------
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(umma::smem_u32(bar)), "r"(bytes) : "memory");
}
// global -> shared bulk copy executed by the TMA engine; completion is counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(umma::smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(umma::smem_u32(bar)) : "memory");
}

// mbarrier arrive WITHOUT release semantics.  The default (release) arrive waits until every earlier memory operation
// of the thread is performed -- including its checkpoint stores to global memory, an L2 round trip (~1000 cycles) that
// paced every sub-step of the sweeps (measured with the phase build: an epilogue sub-step took ~1000 cycles whatever
// its arithmetic).  What a hand-over publishes here is TMEM only: the tcgen05.st has completed (tcgen05.wait::st) and
// tcgen05.fence::before_thread_sync orders it before the arrive, so no memory release is needed.
__device__ __forceinline__ void mbar_arrive_relaxed(uint64_t* bar) {
  asm volatile("mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [%0];" :: "r"(umma::smem_u32(bar)) : "memory");
}

// sticky diagnostic word (njode_device_status): bit 2 / 3 / 4 = wide forward / reverse / weight-gradient kernel
// gave up on an mbarrier; bits 8.. = code of the wait site.  The kernel then TRAPS: a protocol failure must never
// produce silently wrong numbers.  Bring-up aid: with NJODE_NO_TRAP=1 in the environment the thread records the
// site and carries on without waiting any more (garbage results, but the status word can be read back).
// (each translation unit owns its status / no-trap words: no relocatable device code in this build)
struct Diag {
  unsigned* status;          // [0] = sticky word, [1] = first site that gave up: code | warp << 8 | block << 16
  unsigned notrap;           // bring-up mode (read once per role from the translation unit's g_notrap)
  unsigned bit;
  bool dead;
};
__device__ __forceinline__ bool mbar_wait_bounded(uint64_t* bar, uint32_t parity, uint32_t spins) {
  const uint32_t a = umma::smem_u32(bar);
#pragma unroll 1
  for (uint32_t spin = 0; spin < spins; ++spin) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(done) : "r"(a), "r"(parity) : "memory");
    if (done) return true;
  }
  return false;
}
// 32-bit shared-window address variants: a generic pointer into shared memory costs an S2R (window base) wherever the
// compiler rematerialises it, which it does inside register-capped stage loops
// (volatile: computed once and kept, never rematerialised)
__device__ __forceinline__ uint32_t smem_u32_once(const void* p) {
  uint32_t r;
  asm volatile("{\n\t.reg .u64 t;\n\tcvta.to.shared.u64 t, %1;\n\tcvt.u32.u64 %0, t;\n\t}\n" : "=r"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ bool mbar_wait_a(uint32_t a, uint32_t parity, uint32_t spins) {
#pragma unroll 1
  for (uint32_t spin = 0; spin < spins; ++spin) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(done) : "r"(a), "r"(parity) : "memory");
    if (done) return true;
  }
  return false;
}
__device__ __forceinline__ void wait_or_die_a(uint32_t bar_s, uint32_t parity, Diag& dg, unsigned code) {
  if (dg.dead) return;
  if (mbar_wait_a(bar_s, parity, dg.notrap ? (1u << 17) : (1u << 24))) return;
  atomicOr(dg.status, dg.bit | (1u << (8 + code)));
  atomicCAS(dg.status + 1, 0u, code | ((threadIdx.x >> 5) << 8) | (blockIdx.x << 16));
  __threadfence_system();
  if (dg.notrap) dg.dead = true; else __trap();
}
__device__ __forceinline__ void mbar_arrive_a(uint32_t bar_s) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar_s) : "memory");
}
__device__ __forceinline__ void ld8s_a(uint32_t a, float (&v)[8]) {
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(a));
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4+16];" : "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "r"(a));
}
__device__ __forceinline__ void wait_or_die(uint64_t* bar, uint32_t parity, Diag& dg, unsigned code) {
  if (dg.dead) return;
  if (dg.notrap ? mbar_wait_bounded(bar, parity, 1u << 17) : umma::mbar_wait(bar, parity)) return;    // 2^24 polls
  atomicOr(dg.status, dg.bit | (1u << (8 + code)));
  atomicCAS(dg.status + 1, 0u, code | ((threadIdx.x >> 5) << 8) | (blockIdx.x << 16));
  __threadfence_system();
  if (dg.notrap) dg.dead = true; else __trap();
}
// host side of the bring-up switch
static inline int njode_no_trap_env() {
  static const int v = [] { const char* e = getenv("NJODE_NO_TRAP"); return e ? atoi(e) : 0; }();
  return v;
}
static inline int njode_debug_sync_env() {
  static const int v = [] { const char* e = getenv("NJODE_DEBUG_SYNC"); return e ? atoi(e) : 0; }();
  return v;
}

// optional phase accounting (`make phase`, tools/phase_wide.py): thread 0 of every CTA adds up the SM cycles it spends
// in each phase of its role and stores the sums at the end.  PH(i) closes the current interval into bucket i.
#ifdef NJODE_PHASE
#define PH_DECL_AT(tid) long long ph_t__ = clock64(); long long ph_acc__[8] = {0, 0, 0, 0, 0, 0, 0, 0}; const bool ph_on__ = (threadIdx.x == (tid))
#define PH_DECL PH_DECL_AT(0)
#define PH(i) do { if (ph_on__) { const long long n__ = clock64(); ph_acc__[i] += n__ - ph_t__; ph_t__ = n__; } } while (0)
#define PH_STORE(buf) do { if (ph_on__ && blockIdx.x < 512) { for (int i__ = 0; i__ < 8; ++i__) (buf)[blockIdx.x][i__] = (unsigned long long)ph_acc__[i__]; } } while (0)
#else
#define PH_DECL_AT(tid) do { } while (0)
#define PH_DECL do { } while (0)
#define PH(i) do { } while (0)
#define PH_STORE(buf) do { } while (0)
#endif

}  // namespace wide

// flavour entry points (njode_wide.cu / njode_wgrad.cu)
int    njode_wide_supported(const NjodeDesc* d);
size_t njode_wide_image_bytes(const NjodeDesc* d);                 // one direction (forward or transposed images)
int    njode_wide_workers(const NjodeDesc* d, int64_t n_tiles);
int    njode_wide_forward(const SweepArgs& a, float* images, cudaStream_t st);
int    njode_wide_backward(const SweepArgs& a, float* images, cudaStream_t st);   // chain sweep, then weight gradients
int    njode_wide_wgrad(const SweepArgs& a, cudaStream_t st);
int    njode_wide_status(unsigned* out_host);
