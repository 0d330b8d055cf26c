// njode_tiled.cu -- tcgen05 / TMEM sweep kernels for hidden_dim = 32, one hidden layer (BASELINE configs 1-3).
//
// A CTA owns a tile of 128 observation units of one network stack.  Row r of the tile is TMEM lane r; FOUR
// threads share a row, each owning 8 of its 32 columns (512 threads: warp w serves TMEM lane quadrant w % 4
// and column slice w / 4).  The sweeps are latency chains (every Euler step is 2 dependent GEMM phases), so
// the per-phase CUDA-core work per thread is what sets the step time: 8 elements instead of 32 per thread.
//
// Every Linear layer of the forward sweep, the re-computation and the two data-gradient products of the
// reverse sweep are "chain" GEMMs  D[128 x 32] = A[128 x 32] * B[32 x 32]^T:
//     A  is written by the row owners straight into TMEM (tcgen05.st) -- no activation tile ever goes
//        through shared memory on the chain,
//     B  is the pre-split weight tile in shared memory (K-major, 128B swizzle),
//     D  accumulates in TMEM and is read back with tcgen05.ld for the fused epilogue
//        (bias + x/t/dt columns + activation + Euler update + checkpoint store).
// FP32 accuracy comes from the 3xTF32 split (Al*Bh + Ah*Bl + Ah*Bh, 12 MMAs of M128 N32 K8 per GEMM,
// 16.4 cycles each = tensor floor in TS mode).  Weight gradients are sums over rows, i.e. GEMMs whose
// contraction index is the row: they run as stacked MN-major MMAs  [D1|D0]^T (M=64) x [Z|A_in|aux] (N=72)
// over shared-memory tiles (32 cycles per MMA = the 128 B/cycle shared-memory operand floor); the aux
// columns (1, x, t, dt) give the bias and observation/time-column gradients in the same MMA.  The tensor
// core's accumulate is not round-to-nearest: ~3000 MMAs into one TMEM accumulator drifted by 6e-5
// (measured), so every step accumulates its 48 MMAs into a FRESH TMEM accumulator that the CUDA cores then
// merge into running sums (also TMEM-resident) with IEEE fp32 adds.  See DESIGN.md for the measurements.
#include <cstdlib>

#include "njode_common.cuh"
#include "njode_umma.cuh"

namespace {

constexpr int R = NJODE_TILED_TILE_ROWS;   // 128 rows per tile
constexpr int H = 32;
constexpr int CW = 8;                      // columns per thread
constexpr int NT = R * (H / CW);           // 512 threads
constexpr int TILE_F = R * 32;             // floats in a [128][32] tile (16 KB)
constexpr int WT_F = 32 * 32;              // floats in a weight tile (4 KB)
constexpr int MAX_DX = 2, MAX_O = 4;

struct __align__(16) SmallParams {
  float b_ode0[32], b_ode1[32], b_jump0[32], b_jump1[32], b_out0[32];
  float ext_ode0[MAX_DX + 2][32];   // [e][j]: columns H.. of the ODE first layer (x.., t_cur, dt)
  float w_jump0[MAX_DX][32];        // [e][j]
  float w_out1[MAX_O][32];          // [o][j]
  float b_out1[MAX_O];
  float red[4][R][MAX_O];           // forward: readout partial sums of the 4 column slices; backward: bias-gradient reduction
};

// sticky diagnostic word: bit 0 = forward, bit 1 = backward saw an mbarrier wait time out (njode_device_status)
__device__ unsigned g_tiled_status = 0;
// A wait that gives up is fatal: the role records it in the status word and traps, so that the next CUDA call of the
// process fails instead of training on predictions and gradients from unfinished MMAs (a spurious time-out is possible
// under time-slicing, MPS or a debugger).  NJODE_NO_TRAP=1 (bring-up): record and carry on, the status word stays readable.
__device__ unsigned g_tiled_notrap = 0;
__device__ __forceinline__ void tiled_gave_up(unsigned bit) {
  atomicOr(&g_tiled_status, bit);
  __threadfence_system();
  if (!g_tiled_notrap) __trap();
}

// optional phase trace (`make trace`): one thread per role of CTA 0 records (clock64 << 8 | id) at phase
// boundaries.  Forward: thread 0 writes straight to global memory.  Reverse sweep: worker thread 0 and the MMA
// issuer write into a shared-memory ring (a global store in front of the release-arrive of a hand-over would
// itself cost an L2 round trip and distort the trace) that is dumped when the role finishes.
#ifdef NJODE_TRACE
#define NJODE_TRACE_CAP 12288
#define NJODE_TRACE_SMEM_REC 512
__device__ long long g_trace[NJODE_TRACE_CAP];
__device__ int g_trace_n[3] = {0, 0, 0};
#define TR_DECL(part) const int tr_base__ = (part) * (NJODE_TRACE_CAP / 3); int tr_n__ = 0; \
    const bool tr_on__ = (threadIdx.x == ((part) == 2 ? NT : 0)) && blockIdx.x == 0; long long* tr_smem__ = nullptr; (void)tr_smem__
#define TR_SMEM(ptr) tr_smem__ = (ptr)
#define TR(id) do { if (tr_on__) { const long long v__ = (clock64() << 8) | (long long)(id); \
    if (tr_smem__) { tr_smem__[tr_n__ & (NJODE_TRACE_SMEM_REC - 1)] = v__; ++tr_n__; } \
    else if (tr_n__ < NJODE_TRACE_CAP / 3) { g_trace[tr_base__ + tr_n__] = v__; ++tr_n__; } } } while (0)
#define TR_END(part) do { if (tr_on__) { int n__ = tr_n__; \
    if (tr_smem__) { n__ = n__ < NJODE_TRACE_SMEM_REC ? n__ : NJODE_TRACE_SMEM_REC; \
      for (int i__ = 0; i__ < n__; ++i__) g_trace[tr_base__ + i__] = tr_smem__[(tr_n__ - n__ + i__) & (NJODE_TRACE_SMEM_REC - 1)]; } \
    g_trace_n[part] = n__; } } while (0)
#define NJODE_TRACE_SMEM_BYTES (2 * NJODE_TRACE_SMEM_REC * 8)
#define TRW_FLIP ph_tw ^= 1u     /* issuer: parity of the weight-gradient barrier (trace build measures its completion) */
#else
#define TRW_FLIP do { } while (0)
#define TR_DECL(part) do { } while (0)
#define TR_SMEM(ptr) do { } while (0)
#define TR(id) do { } while (0)
#define TR_END(part) do { } while (0)
#define NJODE_TRACE_SMEM_BYTES 0
#endif

struct Ctl {
  uint64_t bar_chain, bar_wgrad, bar_ops;
  uint32_t tmem_base;
  uint32_t timeout;
};

// weight tile W[n][k] (transpose: W^T) -> tf32 hi / lo, K-major 128B-swizzled B operand
__device__ __forceinline__ void load_wtile(float* hi, float* lo, const float* __restrict__ W, int ld, bool transpose, int nthreads) {
  for (int idx = threadIdx.x; idx < WT_F; idx += nthreads) {
    const int n = idx >> 5, k = idx & 31;
    const float v = transpose ? W[k * ld + n] : W[n * ld + k];
    const float h = umma::tf32_hi(v);
    hi[umma::swz_k(n, k)] = h;
    lo[umma::swz_k(n, k)] = umma::tf32_hi(v - h);
  }
}

// extension tile of the forward sweep's first ODE layer: B[n][k] = W0[n][H + k] for k < d_x + 2 (x.., t_cur, dt columns),
// the bias b0[n] at k = d_x + 2, zero beyond (K-major, 128B swizzle, tf32 hi / lo like every weight tile)
__device__ __forceinline__ void load_ext_tile(float* hi, float* lo, const ParamTable& T, const float* __restrict__ p, int nthreads) {
  const int ld0 = H + T.d_x + 2, ne = T.d_x + 2;
  for (int idx = threadIdx.x; idx < WT_F; idx += nthreads) {
    const int n = idx >> 5, k = idx & 31;
    const float v = k < ne ? p[T.w_off[NET_ODE][0] + n * ld0 + H + k] : (k == ne ? p[T.b_off[NET_ODE][0] + n] : 0.0f);
    const float h = umma::tf32_hi(v);
    hi[umma::swz_k(n, k)] = h;
    lo[umma::swz_k(n, k)] = umma::tf32_hi(v - h);
  }
}

__device__ __forceinline__ void load_small(SmallParams& sp, const ParamTable& T, const float* __restrict__ p) {
  const int t = threadIdx.x;
  if (t < 32) {
    const int ld0 = H + T.d_x + 2;
    sp.b_ode0[t] = p[T.b_off[NET_ODE][0] + t];
    sp.b_ode1[t] = p[T.b_off[NET_ODE][1] + t];
    sp.b_jump0[t] = p[T.b_off[NET_JUMP][0] + t];
    sp.b_jump1[t] = p[T.b_off[NET_JUMP][1] + t];
    sp.b_out0[t] = p[T.b_off[NET_OUT][0] + t];
    for (int e = 0; e < T.d_x + 2; ++e) sp.ext_ode0[e][t] = p[T.w_off[NET_ODE][0] + t * ld0 + H + e];
    for (int e = 0; e < T.d_x; ++e) sp.w_jump0[e][t] = p[T.w_off[NET_JUMP][0] + t * T.d_x + e];
    for (int o = 0; o < T.O; ++o) sp.w_out1[o][t] = p[T.w_off[NET_OUT][1] + o * H + t];
    if (t < T.O) sp.b_out1[t] = p[T.b_off[NET_OUT][1] + t];
  }
}

// 8 consecutive floats (32-byte aligned) from SHARED memory.  Explicit ld.shared: through the struct reference the
// compiler had lost the address space and emitted generic LD.E.128, which the LSU orders like a global access.
__device__ __forceinline__ void ld8(const float* __restrict__ src, float (&v)[8]) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(src);
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(a));
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4+16];" : "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "r"(a));
}
// one 256-bit global load / store per thread for its 8 checkpoint floats (LDG.256 / STG.256 on sm_100): one L2
// request per 32-byte sector instead of two, and no L1 allocation (the fills compete with the tensor core's
// shared-memory operand fetch: -6.5 % on the reverse sweep).  128-bit ld.global.cg pairs were measured slower.
__device__ __forceinline__ void ld8_cg(const float* __restrict__ src, float (&v)[8]) {
  asm volatile("ld.global.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(src));
}
// scalar global load that does not allocate in L1 (L1 fills share the SRAM data path with the tensor core's
// shared-memory operand fetch: checkpoint / knot loads in flight measurably slow the weight-gradient MMA batch)
__device__ __forceinline__ float ld_na(const float* __restrict__ p) {
  float v;
  asm volatile("ld.global.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ int ld_na_s32(const int32_t* __restrict__ p) {
  int v;
  asm volatile("ld.global.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ long long ld_na_s64(const int64_t* __restrict__ p) {
  long long v;
  asm volatile("ld.global.L1::no_allocate.s64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }
__device__ __forceinline__ void st8_stream(float* __restrict__ dst, const float (&v)[8]) {
  asm volatile("st.global.v8.f32 [%8], {%0,%1,%2,%3,%4,%5,%6,%7};"
               :: "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "l"(dst) : "memory");
}

// Operands are addressed through descriptor bases: every weight tile / MN tile sits at a compile-time offset
// from the first one, so a descriptor is base + constant (the 14-bit address field cannot overflow: shared
// memory is < 256 KB) and the issuer keeps only two 64-bit bases in registers.
constexpr uint64_t WT_DESC_STEP = WT_F * 4 / 16;     // one weight tile (4 KB) in descriptor address units
constexpr uint64_t MN_DESC_STEP = TILE_F * 4 / 16;   // one MN tile (16 KB)
__device__ __forceinline__ uint64_t wdesc(uint64_t wbase, int id, int lo) { return wbase + (uint64_t)(id * 2 + lo) * WT_DESC_STEP; }
__device__ __forceinline__ uint64_t tdesc(uint64_t tbase, int id) { return tbase + (uint64_t)id * MN_DESC_STEP; }

// 3xTF32 chain GEMM, A from TMEM: acc = A * B^T   (issued by one thread); dbh / dbl = K-major descriptors of B hi / lo
__device__ __forceinline__ void issue_chain(uint32_t tmem_acc, uint32_t tmem_a_hi, uint32_t tmem_a_lo, uint64_t dbh, uint64_t dbl) {
  constexpr uint32_t idesc = umma::idesc_tf32(128, 32, 0, 0);
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) umma::mma_ts(tmem_acc, tmem_a_lo + 8 * ks, dbh + 2 * ks, idesc, ks > 0);
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) umma::mma_ts(tmem_acc, tmem_a_hi + 8 * ks, dbl + 2 * ks, idesc, 1);
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) umma::mma_ts(tmem_acc, tmem_a_hi + 8 * ks, dbh + 2 * ks, idesc, 1);
}

// 3xTF32 row-contraction GEMM: acc[64 x N] = [A0|A1]^T (MN-major tiles, rows = contraction) * [B0|B1|..] over the
// first 8 * nks rows of the tiles (nks = 16 for full tiles; partially filled tiles contract over their units only)
template <int N>
__device__ __forceinline__ void issue_wgrad(uint32_t tmem_acc, uint64_t dah, uint64_t dal, uint64_t dbh, uint64_t dbl, int nks) {
  constexpr uint32_t idesc = umma::idesc_tf32(64, N, 1, 1);
#pragma unroll 4
  for (int ks = 0; ks < nks; ++ks) umma::mma_ss(tmem_acc, dal + 64 * ks, dbh + 64 * ks, idesc, ks > 0);   // fresh accumulator
#pragma unroll 4
  for (int ks = 0; ks < nks; ++ks) umma::mma_ss(tmem_acc, dah + 64 * ks, dbl + 64 * ks, idesc, 1);
#pragma unroll 4
  for (int ks = 0; ks < nks; ++ks) umma::mma_ss(tmem_acc, dah + 64 * ks, dbh + 64 * ks, idesc, 1);
}

// input scaling of 8 values with ONE warp-uniform switch (a per-element runtime switch bloats the loops)
__device__ __forceinline__ void scale8(int sc, float (&v)[8]) {
  if (sc == NJODE_SCALE_TANH) {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = tanhf(v[j]);
  } else if (sc == NJODE_SCALE_SIGMOID) {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 1.0f / (1.0f + expf(-v[j]));
  }
}
// g += acc * s'(.) expressed through the scaled value sv
__device__ __forceinline__ void scale_grad_acc8(int sc, const float (&acc)[8], const float (&sv)[8], float (&g)[8]) {
  if (sc == NJODE_SCALE_TANH) {
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] = fmaf(acc[j], 1.0f - sv[j] * sv[j], g[j]);
  } else if (sc == NJODE_SCALE_SIGMOID) {
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] = fmaf(acc[j], sv[j] * (1.0f - sv[j]), g[j]);
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] += acc[j];
  }
}

__device__ __forceinline__ int64_t pred_index(const ParamTable& T, int64_t obs, int s, int o) {
  return T.S == 1 ? obs * T.d_y * T.M + o : (obs * T.d_y + o) * T.M + s;
}

// tiles are sorted by step count (descending); worker w takes them in snake order so that every worker
// gets one long and one short tile per pair of rounds
__device__ __forceinline__ int64_t snake_tile(int64_t round, int worker, int n_workers) {
  return round * n_workers + ((round & 1) ? n_workers - 1 - worker : worker);
}

// ------------------------------------------------------------------------------------------------
// forward sweep
// ------------------------------------------------------------------------------------------------
// FW_EXT: a fifth k-step of the first ODE layer.  Its bias and its x / t / dt columns (jump_ode.py:57-61) are K columns 32.. of
// the same GEMM -- A columns (s(x).., t_cur, dt, 1), written by the row's first thread -- instead of 4 shared-memory vector loads
// and 32 FP32 instructions per thread and step in the epilogue: the forward sweep is bound by CUDA-core issue slots (2 CTAs per
// SM, ~60 % issue utilisation), the tensor pipe is 78 % idle, and 3 more MMAs cost ~85 cycles of issue.
enum { FW_ODE0 = 0, FW_ODE1, FW_JUMP1, FW_OUT0, FW_EXT, FW_COUNT };
constexpr uint32_t F_AK = 40;                       // A operand columns: 32 hidden units + 8 extension columns
constexpr uint32_t F_AHI = 0, F_ALO = F_AK, F_ACC = 2 * F_AK, F_TMEM_COLS = 128;
constexpr size_t FWD_SMEM = 1024 + FW_COUNT * 2 * WT_F * 4 + sizeof(SmallParams) + sizeof(Ctl) + 16 + NJODE_TRACE_SMEM_BYTES;

template <int ACT>
__global__ void __launch_bounds__(NT, 2) k_tiled_forward(SweepArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  float* wt = reinterpret_cast<float*>(base);                                  // [FW_COUNT][2][WT_F]
  SmallParams& sp = *reinterpret_cast<SmallParams*>(base + FW_COUNT * 2 * WT_F * 4);
  Ctl& ctl = *reinterpret_cast<Ctl*>(base + FW_COUNT * 2 * WT_F * 4 + sizeof(SmallParams));

  const ParamTable& T = a.T;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  TR_DECL(0);
  TR_SMEM(reinterpret_cast<long long*>((reinterpret_cast<uintptr_t>(&ctl + 1) + 15) & ~(uintptr_t)15));
  const int q = warp & 3, c = warp >> 2, row = q * 32 + lane, col0 = c * CW;
  const int s = blockIdx.x % T.S;
  const int worker = blockIdx.x / T.S, n_workers = gridDim.x / T.S;
  const float* p = a.params + (int64_t)s * T.stack_floats;
  const int dx = T.d_x, O = T.O, sc_kind = a.desc.input_scaling;
  // checkpoints: [stack][slot][plane][column chunk c][row][8]; plane 0 = hidden state before the step of this slot
  // (after the last step for the tile's final slot), plane 1 = hidden-layer activation z of that step (saves the
  // reverse sweep the re-computation GEMM and its epilogue).  Chunk-major inside a plane: a warp (one c, 32 rows)
  // stores / loads 1 KB contiguous = 8 full lines; row-major [row][32] made every 256-bit access of a warp touch 32
  // different lines, and the LSU time of those (512 line visits per plane and CTA) delayed the shared-memory loads
  // queued behind them on this latency-bound chain.
  float* ckpt = a.ckpt ? a.ckpt + (int64_t)s * a.total_slots * (2 * R * H) + (c * R + row) * CW : nullptr;

  load_wtile(wt + (FW_ODE0 * 2) * WT_F, wt + (FW_ODE0 * 2 + 1) * WT_F, p + T.w_off[NET_ODE][0], H + dx + 2, false, NT);
  load_wtile(wt + (FW_ODE1 * 2) * WT_F, wt + (FW_ODE1 * 2 + 1) * WT_F, p + T.w_off[NET_ODE][1], H, false, NT);
  load_wtile(wt + (FW_JUMP1 * 2) * WT_F, wt + (FW_JUMP1 * 2 + 1) * WT_F, p + T.w_off[NET_JUMP][1], H, false, NT);
  load_wtile(wt + (FW_OUT0 * 2) * WT_F, wt + (FW_OUT0 * 2 + 1) * WT_F, p + T.w_off[NET_OUT][0], H, false, NT);
  load_ext_tile(wt + (FW_EXT * 2) * WT_F, wt + (FW_EXT * 2 + 1) * WT_F, T, p, NT);
  load_small(sp, T, p);
  if (tid == 0) {
    umma::mbar_init(&ctl.bar_chain, 1);
    umma::fence_mbar_init();
    ctl.timeout = 0;
  }
  if (warp == 0) umma::tmem_alloc(&ctl.tmem_base, F_TMEM_COLS);
  umma::fence_async_smem();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = ctl.tmem_base;
  const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)col0;
  const uint64_t wbase = umma::desc_k(umma::smem_u32(wt));
  uint32_t phase = 0;
  bool ok = true;

  // acc = (in * W^T)[row][col0 .. col0+8)   (all 512 threads call this together)
  // ext != nullptr (first ODE layer): the row's extension columns (s(x).., t_cur, dt, 1), appended as a fifth k-step
  auto gemm = [&](const float (&in)[8], int wid, float (&acc)[8], const float* ext) {
    uint32_t hi[8], lo[8];
    TR(32 + 1);
    umma::split8(in, hi, lo);
    umma::tmem_st8_raw(lane_base + F_AHI, hi);
    umma::tmem_st8_raw(lane_base + F_ALO, lo);
    if (ext != nullptr && c == 0) {
      float e8[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) e8[j] = ext[j];
      umma::split8(e8, hi, lo);
      umma::tmem_st8_raw(lane_base + F_AHI + 32, hi);
      umma::tmem_st8_raw(lane_base + F_ALO + 32, lo);
    }
    umma::wait_st();
    TR(32 + 2);
    umma::fence_before_sync();
    __syncthreads();
    TR(32 + 3);
    if (warp == 0 && umma::elect_one()) {
      umma::fence_after_sync();
      if (ext != nullptr) {
        constexpr uint32_t idesc = umma::idesc_tf32(128, 32, 0, 0);
        const uint64_t dbh = wdesc(wbase, wid, 0), dbl = wdesc(wbase, wid, 1), dxh = wdesc(wbase, FW_EXT, 0), dxl = wdesc(wbase, FW_EXT, 1);
        const uint32_t a_hi = tmem + F_AHI, a_lo = tmem + F_ALO, d = tmem + F_ACC;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) umma::mma_ts(d, a_lo + 8 * ks, dbh + 2 * ks, idesc, ks > 0);
        umma::mma_ts(d, a_lo + 32, dxh, idesc, 1);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) umma::mma_ts(d, a_hi + 8 * ks, dbl + 2 * ks, idesc, 1);
        umma::mma_ts(d, a_hi + 32, dxl, idesc, 1);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) umma::mma_ts(d, a_hi + 8 * ks, dbh + 2 * ks, idesc, 1);
        umma::mma_ts(d, a_hi + 32, dxh, idesc, 1);
      } else {
        issue_chain(tmem + F_ACC, tmem + F_AHI, tmem + F_ALO, wdesc(wbase, wid, 0), wdesc(wbase, wid, 1));
      }
      umma::commit(&ctl.bar_chain);
    }
    TR(32 + 4);
    ok = ok && umma::mbar_wait(&ctl.bar_chain, phase);
    phase ^= 1;
    umma::fence_after_sync();
    TR(32 + 5);
    umma::tmem_ld8(lane_base + F_ACC, acc);
    TR(32 + 6);
  };

  for (int64_t round = 0; round * n_workers < a.n_tiles; ++round) {
    const int64_t tile = snake_tile(round, worker, n_workers);
    if (tile >= a.n_tiles) continue;
    const int kmax = a.tile_kmax[tile];
    const int u = a.perm[tile * R + row];
    // this thread's slice of the tile's checkpoint / knot slots (32-bit offsets per step from here on)
    float* const ck = ckpt ? ckpt + a.tile_slot_off[tile] * (2 * R * H) : nullptr;
    const float* const kn = a.knots + a.tile_slot_off[tile] * R + row;
    const int ke = u >= 0 ? a.kenc[u] : 0;
    const int K = ke >> 1;
    float x[MAX_DX], xs[MAX_DX];
#pragma unroll
    for (int e = 0; e < MAX_DX; ++e) {
      x[e] = (e < dx && u >= 0) ? a.values[(int64_t)u * dx + e] : 0.0f;
      xs[e] = scale_fwd_rt(sc_kind, x[e]);
    }
    float h[8], z[8], acc[8], cb[8], cw[8];
    // h = jump(x)                                                   jump_ode.py:169 / :176
    ld8(sp.b_jump0 + col0, z);
#pragma unroll
    for (int e = 0; e < MAX_DX; ++e) if (e < dx) {
      ld8(sp.w_jump0[e] + col0, cw);
#pragma unroll
      for (int j = 0; j < 8; ++j) z[j] = fmaf(cw[j], x[e], z[j]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) z[j] = act_fwd<ACT>(z[j]);
    gemm(z, FW_JUMP1, acc, nullptr);
    ld8(sp.b_jump1 + col0, cb);
#pragma unroll
    for (int j = 0; j < 8; ++j) h[j] = act_fwd<ACT>(acc[j] + cb[j]);
    if (ck) st8_stream(ck, h);

    // readout: y = out(h)                                           jump_ode.py:170 / :177, :205-212
    auto readout = [&](float* __restrict__ dst, int64_t obs, bool write) {
      gemm(h, FW_OUT0, acc, nullptr);
      ld8(sp.b_out0 + col0, cb);
      float zz[8], y[MAX_O];
#pragma unroll
      for (int j = 0; j < 8; ++j) zz[j] = act_fwd<ACT>(acc[j] + cb[j]);
#pragma unroll
      for (int o = 0; o < MAX_O; ++o) {
        y[o] = 0.0f;
        if (o < O) {
          ld8(sp.w_out1[o] + col0, cw);
#pragma unroll
          for (int j = 0; j < 8; ++j) y[o] = fmaf(zz[j], cw[j], y[o]);
        }
      }
      *reinterpret_cast<float4*>(sp.red[c][row]) = make_float4(y[0], y[1], y[2], y[3]);
      __syncthreads();
      if (c == 0 && write) {
#pragma unroll
        for (int o = 0; o < MAX_O; ++o) if (o < O)
          dst[pred_index(T, obs, s, o)] = sp.b_out1[o] + ((sp.red[0][row][o] + sp.red[1][row][o]) + (sp.red[2][row][o] + sp.red[3][row][o]));
      }
    };
    readout(a.preds, u, u >= 0);

    // Euler steps with x held constant                              jump_ode.py:188-203, :122-140
    // knots are loaded one step ahead: a load consumed in the step that issues it is an exposed global-memory
    // latency on this latency-bound chain (it was the top stall site of the forward kernel, 17 % of its samples)
    float tn = ld_na(kn);
    float tn_ahead = kmax > 0 ? ld_na(kn + R) : tn;
    for (int k = 0; k < kmax; ++k) {
      const float tc = tn;
      tn = tn_ahead;
      // (lands in its own register and is moved into tn_ahead at the END of the step: assigned here, the compiler's
      // register-rotation move followed the load directly and waited out the whole global-memory latency every step)
      const float tn_loaded = ld_na(kn + (k + 2 <= kmax ? k + 2 : kmax) * R);
      const float delta = __fsub_rn(tn, tc);
      TR(32 + 7);
#pragma unroll
      for (int j = 0; j < 8; ++j) z[j] = h[j];
      scale8(sc_kind, z);
      {
        // extension columns in the order of net.0.weight's columns H.. (x.., t_cur, dt), then the constant 1 of the bias
        static_assert(MAX_DX == 2, "extension column layout assumes d_x <= 2");
        const float ext[8] = {xs[0], dx > 1 ? xs[1] : tc, dx > 1 ? tc : delta, dx > 1 ? delta : 1.0f, dx > 1 ? 1.0f : 0.0f, 0.0f, 0.0f, 0.0f};
        gemm(z, FW_ODE0, acc, ext);
      }
      TR(32 + 9);
#pragma unroll
      for (int j = 0; j < 8; ++j) z[j] = act_fwd<ACT>(acc[j]);
      if (ck) st8_stream(ck + (k * 2 + 1) * (R * H), z);
      gemm(z, FW_ODE1, acc, nullptr);
      TR(32 + 8);
      if (k < K) {
        ld8(sp.b_ode1 + col0, cb);
#pragma unroll
        for (int j = 0; j < 8; ++j) h[j] = fmaf(delta, acc[j] + cb[j], h[j]);
      }
      if (ck) st8_stream(ck + (k + 1) * (2 * R * H), h);
      asm volatile("mov.f32 %0, %1;" : "=f"(tn_ahead) : "f"(tn_loaded));
    }
    readout(a.preds_before, (int64_t)u + 1, u >= 0 && (ke & 1));
  }

  TR_END(0);
  if (!ok && tid == 0) tiled_gave_up(1u);
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_free(tmem, F_TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------
// reverse sweep
// ------------------------------------------------------------------------------------------------
enum { WB_ODE1T = 0, WB_ODE0T, WB_OUT0, WB_OUT0T, WB_JUMP1T, WB_COUNT };
// shared-memory MN tiles; A = [D1M|D0M], B = [ZM|AM|XM] are consecutive so LBO = one tile
enum { T_D1M_HI = 0, T_D0M_HI, T_D1M_LO, T_D0M_LO, T_ZM_HI, T_AM_HI, T_XM_HI, T_ZM_LO, T_AM_LO, T_XM_LO, T_COUNT };
// TMEM columns: chain operands / accumulators, one fresh row-contraction accumulator (72 columns), and the
// running weight-gradient sums (40 useful columns per accumulator row, see merge below)
constexpr uint32_t B_AHI = 0, B_ALO = 32, B_DHI = 64, B_DLO = 96, B_ACCR = 128, B_ACCD = 160, B_SACC = 192,
                   B_RUN_ODE = 288, B_RUN_OUT = 328, B_RUN_J1 = 368, B_RUN_END = 408,   // (J1 rows 32-63 hold the first jump layer)
                   B_TMEM_COLS = 512;
constexpr size_t BWD_SMEM = 1024 + (size_t)T_COUNT * TILE_F * 4 + WB_COUNT * 2 * WT_F * 4 + sizeof(SmallParams) + sizeof(Ctl) + 16 + NJODE_TRACE_SMEM_BYTES;

constexpr int NT_B = NT + 128;  // 16 worker warps + one warpgroup whose first warp issues the MMAs

struct BwdSmem {
  float* tiles;        // [T_COUNT][TILE_F]
  float* wt;           // [WB_COUNT][2][WT_F]
  SmallParams* sp;
  Ctl* ctl;
};
__device__ __forceinline__ BwdSmem bwd_carve(uint8_t* smem_raw) {
  uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  BwdSmem m;
  m.tiles = reinterpret_cast<float*>(base);
  m.wt = m.tiles + (size_t)T_COUNT * TILE_F;
  m.sp = reinterpret_cast<SmallParams*>(reinterpret_cast<uint8_t*>(m.wt) + WB_COUNT * 2 * WT_F * 4);
  m.ctl = reinterpret_cast<Ctl*>(reinterpret_cast<uint8_t*>(m.sp) + sizeof(SmallParams));
  return m;
}

// ------------------------------------------------------------------------------------------------
// MMA issuer (one warp): mirrors the workers' sequence of hand-overs -- 2 per readout, 2 per Euler step,
// 2 for the jump net.  Everything it needs is re-derived here, after the register re-allocation.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void bwd_issuer(const SweepArgs& a, uint8_t* smem_raw) {
  const BwdSmem m = bwd_carve(smem_raw);
  Ctl& ctl = *m.ctl;
  const int S = a.T.S;
  const int worker = blockIdx.x / S, n_workers = gridDim.x / S;
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(&ctl.tmem_base);
  const uint64_t wbase0 = umma::desc_k(umma::smem_u32(m.wt)), tbase0 = umma::desc_mn(umma::smem_u32(m.tiles), TILE_F * 4);
  TR_DECL(2);
  TR_SMEM(reinterpret_cast<long long*>((reinterpret_cast<uintptr_t>(m.ctl + 1) + 15) & ~(uintptr_t)15) + NJODE_TRACE_SMEM_REC);
  bool ok = true;
  uint32_t ph_o = 0;
#ifdef NJODE_TRACE
  uint32_t ph_tw = 0;
#endif
  // (after the first time-out every later wait falls through, so a protocol bug costs seconds, not the GPU)
  auto wait_ops = [&]() { ok = ok && umma::mbar_wait(&ctl.bar_ops, ph_o); ph_o ^= 1; umma::fence_after_sync(); };
  // descriptors are base + constant; laundering the bases through an empty asm at every hand-over keeps the
  // compiler from hoisting ~200 loop-invariant 64-bit descriptors into (spilled) registers
  uint64_t wbase, tbase;
  auto fresh = [&]() { wbase = wbase0; tbase = tbase0; asm volatile("" : "+l"(wbase), "+l"(tbase)); };
  int nks = 0;                                    // K-slices of the row contraction: the current tile's units only
  auto issue_readout = [&]() {
    wait_ops();
    fresh();
    if (umma::elect_one()) {
      issue_chain(tmem + B_ACCR, tmem + B_AHI, tmem + B_ALO, wdesc(wbase, WB_OUT0, 0), wdesc(wbase, WB_OUT0, 1));
      umma::commit(&ctl.bar_chain);
    }
    __syncwarp();
    wait_ops();
    fresh();
    if (umma::elect_one()) {
      issue_chain(tmem + B_ACCD, tmem + B_DHI, tmem + B_DLO, wdesc(wbase, WB_OUT0T, 0), wdesc(wbase, WB_OUT0T, 1));
      umma::commit(&ctl.bar_chain);
      issue_wgrad<40>(tmem + B_SACC, tdesc(tbase, T_D1M_HI), tdesc(tbase, T_D1M_LO), tdesc(tbase, T_AM_HI), tdesc(tbase, T_AM_LO), nks);
      umma::commit(&ctl.bar_wgrad);
    }
    __syncwarp();
    TRW_FLIP;
  };
  for (int64_t round = 0; round * n_workers < a.n_tiles; ++round) {
    const int64_t tile = snake_tile(round, worker, n_workers);
    if (tile >= a.n_tiles) continue;
    const int kmax = a.tile_kmax[tile];
    nks = (tile < a.n_small_tiles ? a.tile_units_small : a.tile_units) / 8;
    issue_readout();
    for (int k = kmax - 1; k >= 0; --k) {
      TR(64);
      wait_ops();
      TR(65);
      fresh();
      if (umma::elect_one()) {
        issue_chain(tmem + B_ACCD, tmem + B_DHI, tmem + B_DLO, wdesc(wbase, WB_ODE1T, 0), wdesc(wbase, WB_ODE1T, 1));    // d z0
        umma::commit(&ctl.bar_chain);
      }
      __syncwarp();
      TR(66);
      wait_ops();
      TR(67);
      fresh();
      if (umma::elect_one()) {
        issue_chain(tmem + B_ACCD, tmem + B_DHI, tmem + B_DLO, wdesc(wbase, WB_ODE0T, 0), wdesc(wbase, WB_ODE0T, 1));    // d s(h)
        umma::commit(&ctl.bar_chain);
        issue_wgrad<72>(tmem + B_SACC, tdesc(tbase, T_D1M_HI), tdesc(tbase, T_D1M_LO), tdesc(tbase, T_ZM_HI), tdesc(tbase, T_ZM_LO), nks);
        umma::commit(&ctl.bar_wgrad);
      }
      __syncwarp();
      TRW_FLIP;
      TR(68);
#ifdef NJODE_TRACE
      umma::mbar_wait(&ctl.bar_wgrad, ph_tw ^ 1u);              // (waits do not consume the phase: the workers still see it)
      TR(69);
#endif
    }
    issue_readout();
    // jump net: the data-gradient chain first (its operand is in TMEM before the readout's weight-gradient MMAs are even
    // done), then ONE row-contraction batch for both layers: [d | d0]^T x [z | aux]
    wait_ops();
    fresh();
    if (umma::elect_one()) {
      issue_chain(tmem + B_ACCD, tmem + B_DHI, tmem + B_DLO, wdesc(wbase, WB_JUMP1T, 0), wdesc(wbase, WB_JUMP1T, 1));
      umma::commit(&ctl.bar_chain);
    }
    __syncwarp();
    wait_ops();
    fresh();
    if (umma::elect_one()) {
      issue_wgrad<40>(tmem + B_SACC, tdesc(tbase, T_D1M_HI), tdesc(tbase, T_D1M_LO), tdesc(tbase, T_AM_HI), tdesc(tbase, T_AM_LO), nks);
      umma::commit(&ctl.bar_wgrad);
    }
    __syncwarp();
    TRW_FLIP;
  }
  TR_END(2);
  if (!ok && (threadIdx.x & 31) == 0) tiled_gave_up(2u);
}

// ------------------------------------------------------------------------------------------------
// row workers (512 threads): load the weights, sweep the tiles, flush the weight-gradient sums
// ------------------------------------------------------------------------------------------------
template <int ACT>
__device__ __forceinline__ void bwd_worker(const SweepArgs& a, uint8_t* smem_raw) {
  const BwdSmem m = bwd_carve(smem_raw);
  float* const tiles = m.tiles;
  float* const wt = m.wt;
  SmallParams& sp = *m.sp;
  Ctl& ctl = *m.ctl;
  const ParamTable& T = a.T;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  TR_DECL(1);
  TR_SMEM(reinterpret_cast<long long*>((reinterpret_cast<uintptr_t>(m.ctl + 1) + 15) & ~(uintptr_t)15));
  const int q = warp & 3, c = warp >> 2, row = q * 32 + lane, col0 = c * CW;
  const int s = blockIdx.x % T.S;
  const int worker = blockIdx.x / T.S, n_workers = gridDim.x / T.S;
  const float* p = a.params + (int64_t)s * T.stack_floats;
  const int dx = T.d_x, O = T.O, sc_kind = a.desc.input_scaling;
  const float* ckpt = a.ckpt + (int64_t)s * a.total_slots * (2 * R * H);     // planes (h, z), see the forward kernel

  auto W = [&](int id, int lo) { return wt + (id * 2 + lo) * WT_F; };
  load_wtile(W(WB_ODE0T, 0), W(WB_ODE0T, 1), p + T.w_off[NET_ODE][0], H + dx + 2, true, NT);
  load_wtile(W(WB_ODE1T, 0), W(WB_ODE1T, 1), p + T.w_off[NET_ODE][1], H, true, NT);
  load_wtile(W(WB_OUT0, 0), W(WB_OUT0, 1), p + T.w_off[NET_OUT][0], H, false, NT);
  load_wtile(W(WB_OUT0T, 0), W(WB_OUT0T, 1), p + T.w_off[NET_OUT][0], H, true, NT);
  load_wtile(W(WB_JUMP1T, 0), W(WB_JUMP1T, 1), p + T.w_off[NET_JUMP][1], H, true, NT);
  load_small(sp, T, p);
  // MN tiles may hold anything at start; unused rows / columns only feed accumulator cells nobody reads,
  // but NaN * 0 must not leak into used cells: XM columns beyond the 8 used ones are never addressed (N = 72).
  for (int i = tid; i < T_COUNT * TILE_F / 4; i += NT) reinterpret_cast<float4*>(tiles)[i] = make_float4(0.f, 0.f, 0.f, 0.f);

  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(&ctl.tmem_base);
  const uint32_t quad_base = tmem + ((uint32_t)(q * 32) << 16);     // this warp's lane quadrant, column 0
  const uint32_t lane_base = quad_base + (uint32_t)col0;            // ... at this thread's column slice
  const uint32_t tiles_s = umma::smem_u32(tiles);
  // rows beyond a tile's units are padding: the row contraction stops before them, so they skip the tile stores
  bool has_unit_row = true;
  if (c == 0) {  // zero the persistent weight-gradient accumulators
    uint32_t zero[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) zero[i] = 0u;
    for (uint32_t col = 256; col < 448; col += 8) umma::tmem_st8_raw(quad_base + col, zero);   // covers B_RUN_*
    umma::wait_st();
  }
  umma::named_bar_sync(1, NT);     // weights, zeroed tiles and accumulators are in place (the issuer learns it from the first hand-over)

  bool ok = true;
  uint32_t ph_c = 0, ph_w = 0;
  float dbo[MAX_O];                 // readout-bias gradient: this row's dY summed over all tiles (c == 0 threads)
#pragma unroll
  for (int o = 0; o < MAX_O; ++o) dbo[o] = 0.0f;

  auto wait_chain = [&]() { ok = ok && umma::mbar_wait(&ctl.bar_chain, ph_c); ph_c ^= 1; umma::fence_after_sync(); };
  auto wait_wgrad = [&]() { ok = ok && umma::mbar_wait(&ctl.bar_wgrad, ph_w); ph_w ^= 1; umma::fence_after_sync(); };
  // make this thread's TMEM / smem operand writes visible to the tensor core, then tell the issuer
  // (one arrival per warp: 512 arrivals on one mbarrier word are 512 serialized shared-memory atomics per hand-over)
  auto hand_over = [&]() {
    umma::wait_st(); umma::fence_async_smem(); umma::fence_before_sync();
    __syncwarp();
    if (lane == 0) umma::mbar_arrive(&ctl.bar_ops);
  };
  // the same when only TMEM operands were written (no shared-memory tile since the last hand-over)
  auto hand_over_tmem = [&]() {
    umma::wait_st(); umma::fence_before_sync();
    __syncwarp();
    if (lane == 0) umma::mbar_arrive(&ctl.bar_ops);
  };
  // this thread's 8 columns -> TMEM A operand (hi, lo) and/or MN tile (hi, lo)
  auto put = [&](const float (&v)[8], bool to_tmem, uint32_t c_hi, uint32_t c_lo, int tile_hi, int tile_lo) {
    uint32_t hi[8], lo[8];
    umma::split8(v, hi, lo);
    if (to_tmem) { umma::tmem_st8_raw(lane_base + c_hi, hi); umma::tmem_st8_raw(lane_base + c_lo, lo); }
    if (tile_hi >= 0 && has_unit_row) { umma::chunk_to_mn_tile(tiles_s + tile_hi * (TILE_F * 4), row, c, hi); umma::chunk_to_mn_tile(tiles_s + tile_lo * (TILE_F * 4), row, c, lo); }
  };
  auto put_aux = [&](const float (&xv)[8]) {       // per-row scalars: written by the c == 0 thread of the row
    if (c == 0 && has_unit_row) {
      uint32_t hi[8], lo[8];
      umma::split8(xv, hi, lo);
      umma::chunk_to_mn_tile(tiles_s + T_XM_HI * (TILE_F * 4), row, 0, hi);
      umma::chunk_to_mn_tile(tiles_s + T_XM_LO * (TILE_F * 4), row, 0, lo);
    }
  };
  // running[run + col0 ..+8) += fresh[src + col0 ..+8) ; running[run+32+2c ..+2) += fresh[src8+2c ..+2)   (IEEE adds)
  auto merge = [&](uint32_t run, uint32_t src, uint32_t src8, bool wide) {
    float f2[2], q2[2];
    umma::tmem_ld2_nowait(quad_base + src8 + 2 * c, f2);
    umma::tmem_ld2_nowait(quad_base + run + (wide ? 32 : 0) + 2 * c, q2);
    if (wide) {
      float f[8], r8[8];
      umma::tmem_ld8_nowait(lane_base + src, f);
      umma::tmem_ld8_nowait(lane_base + run, r8);
      umma::wait_ld();
#pragma unroll
      for (int i = 0; i < 8; ++i) r8[i] += f[i];
      umma::tmem_st8(lane_base + run, r8);
    } else {
      umma::wait_ld();
    }
    q2[0] += f2[0];
    q2[1] += f2[1];
    umma::tmem_st2(quad_base + run + (wide ? 32 : 0) + 2 * c, q2);
    umma::wait_st();
  };

  // A tile's metadata (step count, first checkpoint slot, this row's unit) is requested ONE TILE AHEAD, in the jump phase of
  // the previous tile, and what hangs off it (the unit's kenc / observation / readout gradient, the h_end checkpoint) is
  // pulled into L2 at the end of that phase: at the top of a tile two dependent HBM round trips (~2000 cycles with the
  // tensor pipe idle) become one L2 hit.
  int kmax_pf = 0, u_pf = -1, ke_pf = 0;
  long long slot_pf = 0;
  auto request_tile = [&](int64_t t) {
    kmax_pf = ld_na_s32(a.tile_kmax + t);
    slot_pf = ld_na_s64(a.tile_slot_off + t);
    u_pf = ld_na_s32(a.perm + t * R + row);
  };
  // ... and what the tile's first phase consumes -- the unit's step code, observation, readout gradient (un-masked: the
  // mask is in kenc, which arrives together with it) and the h_end checkpoint -- is LOADED behind the last hand-over of the
  // previous tile, into registers that are dead there, so the ~1300 cycles of that tile's last MMA wait hide the latency.
  float hrow[8], dY_pf[MAX_O], x_pf[MAX_DX];
  auto load_prologue = [&]() {
    const int u = u_pf;
    ke_pf = u >= 0 ? ld_na_s32(a.kenc + u) : 0;
#pragma unroll
    for (int e = 0; e < MAX_DX; ++e) x_pf[e] = (e < dx && u >= 0) ? ld_na(a.values + (int64_t)u * dx + e) : 0.0f;
    const int64_t ob = (int64_t)u + 1 < a.N ? (int64_t)u + 1 : a.N - 1;          // (row u + 1 exists whenever kenc says so)
#pragma unroll
    for (int o = 0; o < MAX_O; ++o) dY_pf[o] = (u >= 0 && o < O) ? ld_na(a.grad_preds_before + pred_index(T, ob, s, o)) : 0.0f;
    ld8_cg(ckpt + (int64_t)slot_pf * (2 * R * H) + (c * R + row) * CW + (int64_t)kmax_pf * (2 * R * H), hrow);
  };
  // A readout's backward starts with its hidden layer re-computed from the hidden state in hrow (chain GEMM 1).  That needs
  // nothing but hrow, so it is started EARLY: for a tile's first readout at the end of the previous tile, for the second
  // one before the last Euler step's weight gradients are merged -- the chain then runs in the shadow of work that has to
  // happen anyway.  (Between two hand-overs there is always a wait for an MMA commit that the issuer only makes after it
  // has consumed the first of them, so the workers can never complete two phases of bar_ops unseen.)
  auto readout_begin = [&]() {
    put(hrow, true, B_AHI, B_ALO, -1, -1);
    TR(18);
    hand_over_tmem();
    TR(19);
  };
  if (snake_tile(0, worker, n_workers) < a.n_tiles) {
    request_tile(snake_tile(0, worker, n_workers));
    load_prologue();
    readout_begin();
  }

  for (int64_t round = 0; round * n_workers < a.n_tiles; ++round) {
    const int64_t tile = snake_tile(round, worker, n_workers);
    if (tile >= a.n_tiles) continue;
    TR(17);
    const int kmax = kmax_pf;
    has_unit_row = row < (tile < a.n_small_tiles ? a.tile_units_small : a.tile_units);
    const int u = u_pf;
    const int ke = ke_pf;
    // this thread's slice of the tile's checkpoint / knot slots (32-bit offsets per step from here on)
    const float* ck = ckpt + (int64_t)slot_pf * (2 * R * H) + (c * R + row) * CW;
    const float* kn = a.knots + (int64_t)slot_pf * R + row;
    const int64_t next_tile = snake_tile(round + 1, worker, n_workers);
    const bool has_next = next_tile < a.n_tiles;
    float xs[MAX_DX];
#pragma unroll
    for (int e = 0; e < MAX_DX; ++e) xs[e] = scale_fwd_rt(sc_kind, x_pf[e]);
    if (u >= 0) prefetch_l2(a.grad_preds + pred_index(T, u, s, 0));       // the second readout's gradient, for the end of the tile
    float g[8], z[8], acc[8], d[8];
    float tn = 0.0f, tc_next = 0.0f;                  // knots of the step being reversed (loaded inside the first readout)
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] = 0.0f;

    // ---- readout backward at a hidden state `hrow`; adds d loss / d hrow to g (no MMA in flight at entry) ----
    // dY = d loss / d readout of this row (loaded by the caller well ahead: a load issued here would be waited for by the
    // memory fence of the first hand-over, an exposed global-memory latency per readout).  `first_step` (readout at
    // h_end only): the operands of the first Euler step of the reverse loop are requested right behind the second
    // hand-over -- their registers are free from there on and the ~2000 cycles of MMA waits below hide the latency.
    auto out_backward = [&](const float (&dY)[MAX_O], bool first_step) {      // (after readout_begin; the tiles are free)
      float cw[8], cb[8];
      put(hrow, false, 0, 0, T_AM_HI, T_AM_LO);
      static_assert(MAX_O == 4, "aux column layout assumes <= 4 readout columns");
      const float xv[8] = {1.0f, dY[0], dY[1], dY[2], dY[3], 0.0f, 0.0f, 0.0f};
#pragma unroll
      for (int o = 0; o < MAX_O; ++o) dbo[o] += dY[o];
      put_aux(xv);
      // d (readout hidden pre-activation) needs sum_o dY[o] * w_out1[o][j]: prepare while the MMA runs
      float dz[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) dz[j] = 0.0f;
#pragma unroll
      for (int o = 0; o < MAX_O; ++o) if (o < O) {
        ld8(sp.w_out1[o] + col0, cw);
#pragma unroll
        for (int j = 0; j < 8; ++j) dz[j] = fmaf(dY[o], cw[j], dz[j]);
      }
      ld8(sp.b_out0 + col0, cb);
      wait_chain();
      TR(20);
      umma::tmem_ld8(lane_base + B_ACCR, acc);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        z[j] = act_fwd<ACT>(acc[j] + cb[j]);
        d[j] = dz[j] * act_grad_from_out<ACT>(z[j]);
      }
      put(d, true, B_DHI, B_DLO, T_D1M_HI, T_D1M_LO);
      put(z, false, 0, 0, T_D0M_HI, T_D0M_LO);
      hand_over();
      TR(21);
      if (first_step) {
        tn = ld_na(kn + kmax * R);
        tc_next = kmax > 0 ? ld_na(kn + (kmax - 1) * R) : tn;
        if (kmax > 0) {
          ld8_cg(ck + (kmax - 1) * (2 * R * H), hrow);
          ld8_cg(ck + (kmax - 1) * (2 * R * H) + R * H, z);
          if (kmax > 1) {
            prefetch_l2(ck + (kmax - 2) * (2 * R * H));
            prefetch_l2(ck + (kmax - 2) * (2 * R * H) + R * H);
          }
        }
      }
      wait_chain();
      TR(22);
      umma::tmem_ld8(lane_base + B_ACCD, acc);
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] += acc[j];
      // the readout's weight-gradient MMAs are still running: the caller merges them (merge_pending) when it next needs
      // the shared-memory tiles, and does tile-independent work until then
    };
    // fresh accumulator -> running sums for the batch in flight: 1 = an Euler step (rows 0-31 = d f rows: W1 block = columns
    // 0-31; rows 32-63 = d a0 rows: W0 block = columns 32-63; columns 64.. = bias, x, t, dt gradients for either),
    // 2 = a readout (hidden layer + readout weights)
    auto merge_pending = [&](int kind) {
      wait_wgrad();
      TR(4);
      const bool ode = kind == 1;
      merge(ode ? B_RUN_ODE : B_RUN_OUT, B_SACC + ((ode && q >= 2) ? 32 : 0), B_SACC + (ode ? 64 : 32), true);
    };
    auto load_dY = [&](float (&dY)[MAX_O], const float* __restrict__ gsrc, int64_t obs, bool live) {
#pragma unroll
      for (int o = 0; o < MAX_O; ++o) dY[o] = (live && o < O) ? ld_na(gsrc + pred_index(T, obs, s, o)) : 0.0f;
    };

    // ---- preds_before[u+1] = out(h_end) ----
    {
      float dY[MAX_O];                               // (hrow = h_end and dY_pf were loaded by load_prologue)
#pragma unroll
      for (int o = 0; o < MAX_O; ++o) dY[o] = (ke & 1) ? dY_pf[o] : 0.0f;
      out_backward(dY, true);
    }

    // ---- Euler steps, last to first (tn, tc_next, hrow, z of step kmax-1: requested inside the readout above) ----
    int pending = 2;                                  // weight-gradient MMAs still to be merged: 2 = the readout's, 1 = a step's
    for (int k = kmax - 1; k >= 0; --k) {
      TR(1);
      const float tc = tc_next;
      const float delta = __fsub_rn(tn, tc);          // 0 for rows that took fewer than k+1 steps
      tn = tc;
#pragma unroll
      for (int j = 0; j < 8; ++j) d[j] = delta * g[j];                       // d loss / d f(h)
      put(d, true, B_DHI, B_DLO, -1, -1);                                    // chain operand: straight to TMEM
      TR(2);
      hand_over_tmem();                                                      // -> d z0 = d1 * W1
      TR(3);
      // This step's checkpoints and the next knot.  Issued AFTER the hand-over on purpose: scoreboards are shared,
      // so anything that waits on an older load (delta above) or fences memory (the hand-over) would also wait for
      // these ~1000-cycle HBM loads.  They are consumed after the weight-gradient wait below.
      tc_next = ld_na(kn + (k > 0 ? k - 1 : 0) * R);
      if (k > 1) {   // pull the checkpoints of the step after the next into L2 (no destination register, no scoreboard)
        prefetch_l2(ck + (k - 2) * (2 * R * H));
        prefetch_l2(ck + (k - 2) * (2 * R * H) + R * H);
      }
      // the MN tiles are free once the previous step's weight-gradient MMAs are done
      merge_pending(pending);
      TR(5);
      scale8(sc_kind, hrow);
      TR(13);
      put(hrow, false, 0, 0, T_AM_HI, T_AM_LO);
      TR(14);
      put(d, false, 0, 0, T_D1M_HI, T_D1M_LO);
      TR(15);
      put(z, false, 0, 0, T_ZM_HI, T_ZM_LO);
      TR(16);
      {
        // aux columns (1, s(x).., t, dt); no dynamic indexing (a local-memory array costs an L2 round trip here)
        static_assert(MAX_DX == 2, "aux column layout assumes d_x <= 2");
        const float xv[8] = {1.0f, xs[0], dx > 1 ? xs[1] : tc, dx > 1 ? tc : delta, dx > 1 ? delta : 0.0f, 0.0f, 0.0f, 0.0f};
        put_aux(xv);
      }
      TR(6);
      wait_chain();
      TR(7);
      umma::tmem_ld8(lane_base + B_ACCD, acc);
      TR(8);
#pragma unroll
      for (int j = 0; j < 8; ++j) d[j] = acc[j] * act_grad_from_out<ACT>(z[j]);      // d loss / d a0
      put(d, true, B_DHI, B_DLO, T_D0M_HI, T_D0M_LO);
      TR(9);
      hand_over();                                                           // -> d s(h) = d0 * W0 ; weight gradients
      pending = 1;
      TR(10);
      // the NEXT step's checkpoints, as soon as their registers are free: an L2 hit is ~1000 cycles away and nothing
      // on the way to their first use (the tile stores after the weight-gradient wait) may have to wait for them
      if (k > 0) ld8_cg(ck + (k - 1) * (2 * R * H) + R * H, z);
      wait_chain();
      TR(11);
      umma::tmem_ld8(lane_base + B_ACCD, acc);
      scale_grad_acc8(sc_kind, acc, hrow, g);
      if (k > 0) ld8_cg(ck + (k - 1) * (2 * R * H), hrow);
      TR(12);
    }
    // ---- preds[u] = out(h0), then the jump net ----
    // h0, the readout gradient and the observation are requested BEFORE the wait for the last step's weight-gradient
    // MMAs (~1800 cycles with nothing else to do): by the time the merge is done they have landed
    float x[MAX_DX];
    {
      float dY[MAX_O];
      // h0 is the hidden state before step 0, i.e. what hrow holds after the loop (or, without steps, h_end itself) -- unless
      // an input scaling was applied to it in place
      const bool ident = sc_kind == NJODE_SCALE_IDENTITY;
      if (kmax > 0 && !ident) ld8_cg(ck, hrow);
      load_dY(dY, a.grad_preds, u, u >= 0);
#pragma unroll
      for (int e = 0; e < MAX_DX; ++e) x[e] = ident ? xs[e] : ((e < dx && u >= 0) ? ld_na(a.values + (int64_t)u * dx + e) : 0.0f);
      readout_begin();
      TR(25);
      merge_pending(pending);
      TR(26);
      out_backward(dY, false);
    }
    {
      // z = first jump layer (recomputed), d = d loss / d (pre-activation of h0)
      float cw[8];
      ld8(sp.b_jump0 + col0, z);
#pragma unroll
      for (int e = 0; e < MAX_DX; ++e) if (e < dx) {
        ld8(sp.w_jump0[e] + col0, cw);
#pragma unroll
        for (int j = 0; j < 8; ++j) z[j] = fmaf(cw[j], x[e], z[j]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        z[j] = act_fwd<ACT>(z[j]);
        d[j] = g[j] * act_grad_from_out<ACT>(hrow[j]);
      }
      put(d, true, B_DHI, B_DLO, -1, -1);
      hand_over_tmem();                                        // -> d z = d * W_jump1 (queued behind the readout's batch)
      TR(27);
      if (has_next) request_tile(next_tile);                   // lands under the waits below (this tile's values are dead)
      merge_pending(2);                                        // the readout's weight gradients; the tiles are free now
      put(d, false, 0, 0, T_D1M_HI, T_D1M_LO);
      put(z, false, 0, 0, T_AM_HI, T_AM_LO);
      const float xv[8] = {1.0f, x[0], x[1], 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};     // x[1] = 0 when d_x = 1
      put_aux(xv);
      TR(28);
      wait_chain();
      TR(29);
      umma::tmem_ld8(lane_base + B_ACCD, acc);
#pragma unroll
      for (int j = 0; j < 8; ++j) d[j] = acc[j] * act_grad_from_out<ACT>(z[j]);
      put(d, false, 0, 0, T_D0M_HI, T_D0M_LO);
      hand_over();
      TR(30);
      if (has_next) load_prologue();     // the next tile's first operands: they land under the wait below
      // ONE batch for both layers, [d | d0]^T (M = 64) x [z | aux] (N = 40): accumulator rows 0-31 = second layer (weights
      // and bias), rows 32-63 x aux columns = first layer (bias, x columns); both land in the B_RUN_J1 running sums
      wait_wgrad();
      if (has_next) readout_begin();     // the next tile's first chain GEMM runs under the merge and the tile's top
      merge(B_RUN_J1, B_SACC, B_SACC + 32, true);
      TR(31);
    }
  }

  // ---- flush the TMEM-resident weight-gradient accumulators into this CTA's partial buffer ----
  if (c == 0) *reinterpret_cast<float4*>(sp.red[0][row]) = make_float4(dbo[0], dbo[1], dbo[2], dbo[3]);
  umma::fence_before_sync();
  umma::named_bar_sync(1, NT);
  umma::fence_after_sync();
  {
    float* part = a.partials + (int64_t)blockIdx.x * T.stack_floats;
    const bool has_row = lane < 16;                 // M = 64 accumulator: row i lives in lane (i%16) + 32*(i/16)
    const int i = q * 16 + lane;                    // 0..63 when has_row
    const int j = i & 31;
    const int ld0 = H + dx + 2;
    float v[8], v8[8];
    // ODE net: running rows 0-31 = second layer (W1, b1), rows 32-63 = first layer (W0 incl. x/t/dt columns, b0)
    umma::tmem_ld8_nowait(lane_base + B_RUN_ODE, v);
    umma::tmem_ld8(quad_base + B_RUN_ODE + 32, v8);
    if (has_row && i < 32) {
      for (int k = 0; k < 8; ++k) part[T.w_off[NET_ODE][1] + j * H + col0 + k] = v[k];
      if (c == 0) part[T.b_off[NET_ODE][1] + j] = v8[0];
    }
    if (has_row && i >= 32) {
      for (int k = 0; k < 8; ++k) part[T.w_off[NET_ODE][0] + j * ld0 + col0 + k] = v[k];
      if (c == 0) {
        part[T.b_off[NET_ODE][0] + j] = v8[0];
        for (int e = 0; e < dx + 2; ++e) part[T.w_off[NET_ODE][0] + j * ld0 + H + e] = v8[1 + e];
      }
    }
    // output net: rows 0-31 = hidden layer (W, b); rows 32-63 (z rows) x dY columns = readout weights
    umma::tmem_ld8_nowait(lane_base + B_RUN_OUT, v);
    umma::tmem_ld8(quad_base + B_RUN_OUT + 32, v8);
    if (has_row && i < 32) {
      for (int k = 0; k < 8; ++k) part[T.w_off[NET_OUT][0] + j * H + col0 + k] = v[k];
      if (c == 0) part[T.b_off[NET_OUT][0] + j] = v8[0];
    }
    if (has_row && i >= 32 && c == 0) { for (int o = 0; o < O; ++o) part[T.w_off[NET_OUT][1] + o * H + j] = v8[1 + o]; }
    // jump net
    umma::tmem_ld8_nowait(lane_base + B_RUN_J1, v);
    umma::tmem_ld8(quad_base + B_RUN_J1 + 32, v8);
    if (has_row && i < 32) {
      for (int k = 0; k < 8; ++k) part[T.w_off[NET_JUMP][1] + j * H + col0 + k] = v[k];
      if (c == 0) part[T.b_off[NET_JUMP][1] + j] = v8[0];
    }
    if (has_row && i >= 32 && c == 0) {       // first layer: d0 rows x aux columns (1, x..) of the same running sums
      part[T.b_off[NET_JUMP][0] + j] = v8[0];
      for (int e = 0; e < dx; ++e) part[T.w_off[NET_JUMP][0] + j * dx + e] = v8[1 + e];
    }
    if (tid < O) {     // readout bias: fixed-order sum over the rows (deterministic)
      float sum = 0.0f;
      for (int r = 0; r < R; ++r) sum += sp.red[0][r][tid];
      part[T.b_off[NET_OUT][1] + tid] = sum;
    }
  }
  TR_END(1);
  if (!ok && lane == 0) tiled_gave_up(2u);
}

// Reverse sweep.  Warps 0-15 are row workers (see the header), warp 16 only issues MMAs: the workers hand
// operands over through an mbarrier (bar_ops, 512 arrivals) and never wait for the issue itself, so the
// ~30 cycles per tcgen05.mma of issue time and the back-pressure of the MMA queue stay off their path.
// The weight-gradient MMAs of step k (1536 tensor cycles, reading the shared-memory MN tiles) overlap the
// data-gradient chain and the first half of step k-1: the workers write the MN tiles of a step as late as
// possible (after waiting for the previous step's weight-gradient MMAs) and merge the previous fresh
// accumulator at the same point.  The hidden-layer activations come from the forward sweep's checkpoints.
// Register budget: the CTA is launched with 640 x 96 registers and setmaxnreg only moves registers inside that
// pool, so 512 x W + 128 x I <= 61440: workers W = 104, issuer warpgroup I = 56 (W = 112 / I = 40 dead-locks in
// setmaxnreg.inc).  The role bodies derive everything they need AFTER the re-allocation: whatever is live
// across it gets spilled under the 96-register budget and re-loaded from local memory inside the loops -- an L2
// round trip each with this shared-memory carve-out (measured: ~1000 cycles per step).
template <int ACT>
__global__ void __launch_bounds__(NT_B, 1) k_tiled_backward(SweepArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5;
  {
    Ctl& ctl = *bwd_carve(smem_raw).ctl;
    if (threadIdx.x == 0) {
      umma::mbar_init(&ctl.bar_chain, 1);
      umma::mbar_init(&ctl.bar_wgrad, 1);
      umma::mbar_init(&ctl.bar_ops, NT / 32);
      umma::fence_mbar_init();
      ctl.timeout = 0;
    }
    if (warp == 0) umma::tmem_alloc(&ctl.tmem_base, B_TMEM_COLS);
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
  }
  if (warp >= NT / 32) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == NT / 32) bwd_issuer(a, smem_raw);
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    bwd_worker<ACT>(a, smem_raw);
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_free(*reinterpret_cast<volatile uint32_t*>(&bwd_carve(smem_raw).ctl->tmem_base), B_TMEM_COLS);
}

template <int ACT>
int launch_tiled(const SweepArgs& a, cudaStream_t st, bool backward) {
  if (a.n_tiles == 0) return NJODE_OK;
  static const int no_trap = [] { const char* e = getenv("NJODE_NO_TRAP"); return e ? atoi(e) : 0; }();
  if (no_trap) { const unsigned one = 1; NJODE_CUDA_OK(cudaMemcpyToSymbolAsync(g_tiled_notrap, &one, sizeof(one), 0, cudaMemcpyHostToDevice, st)); }
  if (backward) {
    NJODE_CUDA_OK(cudaFuncSetAttribute(k_tiled_backward<ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BWD_SMEM));
    njode_timing_begin(2, st);
    k_tiled_backward<ACT><<<a.n_workers, NT_B, BWD_SMEM, st>>>(a);
    njode_timing_end(2, st);
    NJODE_LAUNCH_OK("k_tiled_backward");
  } else {
    NJODE_CUDA_OK(cudaFuncSetAttribute(k_tiled_forward<ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FWD_SMEM));
    njode_timing_begin(1, st);
    k_tiled_forward<ACT><<<a.n_workers, NT, FWD_SMEM, st>>>(a);
    njode_timing_end(1, st);
    NJODE_LAUNCH_OK("k_tiled_forward");
  }
  return NJODE_OK;
}

int dispatch_tiled(const SweepArgs& a, cudaStream_t st, bool backward) {
  switch (a.desc.activation) {
    case NJODE_ACT_RELU: return launch_tiled<NJODE_ACT_RELU>(a, st, backward);
    case NJODE_ACT_TANH: return launch_tiled<NJODE_ACT_TANH>(a, st, backward);
    case NJODE_ACT_SIGMOID: return launch_tiled<NJODE_ACT_SIGMOID>(a, st, backward);
    case NJODE_ACT_ELU: return launch_tiled<NJODE_ACT_ELU>(a, st, backward);
    case NJODE_ACT_LEAKY_RELU: return launch_tiled<NJODE_ACT_LEAKY_RELU>(a, st, backward);
    default: return launch_tiled<NJODE_ACT_SELU>(a, st, backward);
  }
}

}  // namespace

int njode_tiled_supported(const NjodeDesc* d) {
  const int O = d->shared_network ? d->d_y * d->num_moments : d->d_y;
  return d->hidden == H && d->n_hidden_layers == 1 && d->d_x <= MAX_DX && O <= MAX_O;
}

static int sm_count() {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms;
}

// CTAs of the reverse sweep (one per SM: it owns all 512 TMEM columns); a multiple of S
int njode_tiled_workers(const NjodeDesc* d, int64_t n_tiles) {
  const int S = d->shared_network ? 1 : d->num_moments;
  int64_t per_stack = sm_count() / S;
  if (per_stack < 1) per_stack = 1;
  if (per_stack > n_tiles) per_stack = n_tiles > 0 ? n_tiles : 1;
  return (int)(per_stack * S);
}

int njode_tiled_forward(const SweepArgs& a_in, cudaStream_t st) {
  SweepArgs a = a_in;
  // forward CTAs use 128 TMEM columns, ~45 KB smem and 512 threads: 2 per SM
  // (measured on the default workload, 320 tiles: one CTA per SM is 7 % slower -- the second resident CTA hides
  //  more latency than it costs even when every CTA has a single tile)
  const int S = a.T.S;
  const int ctas_per_sm = 2;
  int64_t per_stack = (int64_t)sm_count() * ctas_per_sm / S;
  if (per_stack > a.n_tiles) per_stack = a.n_tiles > 0 ? a.n_tiles : 1;
  a.n_workers = (int)(per_stack * S);
  return dispatch_tiled(a, st, false);
}
int njode_tiled_backward(const SweepArgs& a, cudaStream_t st) { return dispatch_tiled(a, st, true); }

// make TRACE=1 only: copy the phase trace of CTA 0 to the host and reset it; returns the number of records
extern "C" int njode_tiled_trace_fetch(long long* out_host, int cap) {
#ifdef NJODE_TRACE
  int n[3] = {0, 0, 0};
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(n, g_trace_n, sizeof(n));
  int total = 0;
  for (int half = 0; half < 3; ++half) {
    int cnt = n[half];
    if (total + cnt > cap) cnt = cap - total;
    if (cnt > 0) cudaMemcpyFromSymbol(out_host + total, g_trace, (size_t)cnt * sizeof(long long),
                                      (size_t)half * (NJODE_TRACE_CAP / 3) * sizeof(long long));
    total += cnt > 0 ? cnt : 0;
  }
  return total;
#else
  (void)out_host; (void)cap;
  return -1;
#endif
}

int njode_tiled_status(unsigned* out_host) {
  NJODE_CUDA_OK(cudaMemcpyFromSymbol(out_host, g_tiled_status, sizeof(unsigned)));
  return NJODE_OK;
}
