// njode_wide.cu -- tcgen05 / TMEM chain sweeps for hidden_dim in {64, 128}, 1..3 hidden layers
// (BASELINE configs 4 and 5: Heston H=128 / L=3 / tanh and the mixed H=64 batch).
//
// A CTA owns a tile of 128 observation units of one network stack; row r of the tile is TMEM lane r.  Every
// Linear of the path is a "chain" GEMM  D[128 x H] = A[128 x H] * W^T  with FP32 accuracy from the 3xTF32 split:
//     A   the activation (forward) / data gradient (reverse), written by the row workers straight into TMEM as
//         tf32 hi / lo parts (tcgen05.st) -- activations never pass through shared memory;
//     W   does NOT fit in shared memory (H=128, L=3: 4 ODE matrices x 128 KB of hi/lo parts), so it is STREAMED:
//         a producer warp walks the same GEMM program as the MMA issuer and pulls pre-split, pre-swizzled stages
//         of 2 x H x 32 floats from a weight image in L2 with cp.async.bulk (TMA engine) into a 6-stage ring,
//         full / empty mbarriers on either side (tcgen05.commit frees a stage);
//     D   accumulates in TMEM, double buffered, and is read back with tcgen05.ld for the fused epilogue.
// Epilogue and MMA of consecutive layers overlap: the 16 worker warps (warp w: lane quadrant w % 4, column
// group w / 4) produce the next A operand in sub-steps of 8 columns per thread; after each sub-step the issuer
// may run the 4 k-steps (one per column group) x 3 passes that only need those columns.  The K order of a
// weight image is permuted accordingly, so that a ring stage is exactly the weights of one sub-step.
// The reverse sweep is the same machine run backwards on transposed images: it propagates d loss / d(pre-
// activation) through every layer and writes those "d planes" next to the forward sweep's activation planes;
// all weight gradients are contractions of (d plane, activation plane) pairs over rows and run afterwards as one
// split-K tensor-core GEMM (njode_wgrad.cu) -- at H = 128 neither shared memory nor TMEM has room to do them here.
#include "njode_wide.cuh"

namespace {
using namespace wide;

__device__ unsigned g_status[2] = {0, 0};
__device__ unsigned g_notrap = 0;
// bring-up aid (njode_debug_cta_cycles): cycles and chain GEMMs of every CTA of the last sweep launch
__device__ unsigned long long g_cta_cycles[512][2];
__device__ unsigned long long g_phase12[512][8];    // `make phase`: per-CTA phase cycles of the worker role (thread 0)
__device__ unsigned long long g_phase_iss[512][8];  // ... and of the MMA issuer (lane 0): 0 operand wait, 1 weight wait, 2 issue

// GEMM order of the chain sweeps.
//   0: k-major.  A GEMM is NSUB passes, pass j = the 12 MMAs (N = H) over the A columns that sub-step j of the producing
//      epilogue completes; the next GEMM starts under the tail of the epilogue, but every accumulator column is complete
//      only after the last pass, so epilogue and MMAs of one layer never overlap (measured: 4100 + 3700 cycles).
//   1: blocked.  A GEMM is 4 output blocks (block n = the CG accumulator columns of column group n), each over the
//      whole K; the warps of column group n start their epilogue when block n is done, i.e. under the MMAs of blocks
//      n+1.., and hold the next operand in registers until the last block is done (the A operand region in TMEM is
//      still being read; at H = 128 TMEM has no room for a second one).
// Measured on B200 (profiles/r2_wide_blocked_order_experiment.log): 1 is SLOWER -- a tcgen05.mma with A from TMEM costs
// ~12 + 0.5 N cycles (N = 16: 19, 32: 29, 64: 50, 128: 76), so the 192 N = 32 MMAs of a blocked H = 128 GEMM take
// 5650 cycles against 3670 for the 48 N = 128 ones; the epilogue does drop from 4100 to 2760 cycles, the sweep goes
// from 3.97 to 5.58 ms.  Kept as a build option for the record; the product is 0.
#ifndef NJODE_WIDE_EARLY_SIGNAL
#define NJODE_WIDE_EARLY_SIGNAL 1
#endif
#ifndef NJODE_WIDE_BLOCKED
#define NJODE_WIDE_BLOCKED 0
#endif
constexpr bool BLOCKED = NJODE_WIDE_BLOCKED != 0;

constexpr int NT = NT_W + 128;         // 16 worker warps + one warpgroup holding the MMA issuer warp and the weight producer warp
// Register budget: the CTA launches with 640 x 96 registers and setmaxnreg moves registers inside that pool:
// 512 x 112 (workers: a 32-float state slice plus, in the reverse sweep, a 32-float activation slice per thread at
// H = 128) + 128 x 32 (issuer / producer: a handful of counters and two descriptors) = 61440.

template <int HW>
struct __align__(16) SmallW {
  float b_jump[NJODE_WIDE_LMAX + 1][HW], b_ode[NJODE_WIDE_LMAX + 1][HW], b_out[NJODE_WIDE_LMAX][HW];
  float ext_ode0[MAX_DX + 2][HW];      // [e][j]: columns H.. of the ODE first layer (x.., t_cur, dt)
  float w_jump0[MAX_DX][HW];           // [e][j]
  float w_out[MAX_O][HW];              // [o][j] readout layer (forward: W; reverse: the same)
  float b_outL[MAX_O];
  float red[4][R][MAX_O];              // readout partial sums of the 4 column groups
};

struct Ctl {
  uint64_t full[8], empty[8], ops[4], accd[4];      // (blocked order: ops[0] only, accd[n] = block n of the current GEMM)
  uint32_t tmem_base, pad;
};

template <int HW>
struct Smem {
  uint8_t* ring;
  SmallW<HW>* sw;
  Ctl* ctl;
  static constexpr size_t bytes() { return 1024 + (size_t)Cfg<HW>::NSTAGE * Cfg<HW>::STAGE_BYTES + sizeof(SmallW<HW>) + sizeof(Ctl) + 16; }
  __device__ __forceinline__ explicit Smem(uint8_t* raw) {
    uint8_t* base = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    ring = base;
    sw = reinterpret_cast<SmallW<HW>*>(base + (size_t)Cfg<HW>::NSTAGE * Cfg<HW>::STAGE_BYTES);
    ctl = reinterpret_cast<Ctl*>(reinterpret_cast<uint8_t*>(sw) + sizeof(SmallW<HW>));
  }
};

__device__ __forceinline__ int64_t pred_index(const ParamTable& T, int64_t obs, int s, int o) {
  return T.S == 1 ? obs * T.d_y * T.M + o : (obs * T.d_y + o) * T.M + s;
}
// the tiles of a worker, in the order it runs them: the schedule's longest-processing-time table (SweepArgs::tile_table)
struct TileList {
  const int32_t* list;
  int n;
  __device__ __forceinline__ TileList(const SweepArgs& a, int worker, int n_workers) {
    const int lo = a.tile_table[worker];
    n = a.tile_table[worker + 1] - lo;
    list = a.tile_table + n_workers + 1 + lo;
  }
};

// ------------------------------------------------------------------------------------------------
// weight images: for every stack and chain matrix, NSUB stages of [hi | lo] x [H rows n][32 k], K-major with the
// 128-byte swizzle.  Stage j holds, for column group g = 0..3, the 8 input features g*CG + 8j .. +8 (local k = 8g + i):
// exactly the columns of the A operand that sub-step j of the producing epilogue completes.
//   forward image:  B[n][k] = W[n][k]     (n = output feature)
//   reverse image:  B[n][k] = W[k][n]     (n = input feature < H: d loss / d input = d * W)
// ------------------------------------------------------------------------------------------------
template <int HW>
__global__ void k_wide_prep(ParamTable T, const float* __restrict__ params, float* __restrict__ img, int transpose) {
  using C = Cfg<HW>;
  const int L = T.L, NM = n_mats(L);
  const int64_t total = (int64_t)T.S * NM * HW * HW;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int k = (int)(idx % HW);
  const int n = (int)((idx / HW) % HW);
  const int m = (int)((idx / ((int64_t)HW * HW)) % NM);
  const int s = (int)(idx / ((int64_t)HW * HW * NM));
  int net, l;
  if (m < L) { net = NET_JUMP; l = m + 1; }
  else if (m <= 2 * L) { net = NET_ODE; l = m - L; }
  else { net = NET_OUT; l = m - 2 * L - 1; }
  const int ld = T.n_vec[net][l] + T.n_ext[net][l];
  const float* W = params + (int64_t)s * T.stack_floats + T.w_off[net][l];
  const float v = transpose ? W[k * ld + n] : W[n * ld + k];
  uint32_t hi, lo;
  umma::split1(v, hi, lo);
  if (BLOCKED) {
    // stage = output block n / CG; inside: k-chunk (k / 32) major, then a [CG rows][32 k] swizzled tile
    float* st = img + ((int64_t)(s * NM + m) * C::NBLK + n / C::CG) * (C::BSTAGE_BYTES / 4) + (k >> 5) * (C::CG * 32);
    st[umma::swz_k(n % C::CG, k & 31)] = __uint_as_float(hi);
    st[C::BSTAGE_HALF / 4 + umma::swz_k(n % C::CG, k & 31)] = __uint_as_float(lo);
    return;
  }
  const int g = k / C::CG, r = k % C::CG, j = r >> 3, i = r & 7, kk = g * 8 + i;
  float* st = img + ((int64_t)(s * NM + m) * C::NSUB + j) * (C::STAGE_BYTES / 4);
  st[umma::swz_k(n, kk)] = __uint_as_float(hi);
  st[C::STAGE_HALF / 4 + umma::swz_k(n, kk)] = __uint_as_float(lo);
}

// ------------------------------------------------------------------------------------------------
// roles shared by both sweeps
// ------------------------------------------------------------------------------------------------
// GEMMs of one tile, in program order: forward  J1..JL | O0..O(L-1) | kmax x (E0..EL) | O0..O(L-1)
//                                      reverse  O(L-1)..O0 | kmax x (EL..E0) | O(L-1)..O0 | JL..J1
template <int HW, bool BWD>
__device__ __forceinline__ void producer(const SweepArgs& a, uint8_t* raw, const float* __restrict__ img) {
  using C = Cfg<HW>;
  Smem<HW> sm(raw);
  Ctl& ctl = *sm.ctl;
  const int S = a.T.S, L = a.T.L;
  const int s = blockIdx.x % S, worker = blockIdx.x / S, n_workers = gridDim.x / S;
  constexpr int NST = BLOCKED ? C::NBLK : C::NSUB;                          // stages per matrix
  constexpr uint32_t SB = BLOCKED ? C::BSTAGE_BYTES : C::STAGE_BYTES;       // bytes per stage
  const uint8_t* simg = reinterpret_cast<const uint8_t*>(img) + (size_t)s * n_mats(L) * NST * SB;
  uint32_t sc = 0;
  Diag dg{g_status, g_notrap, BWD ? 8u : 4u, false};
  auto load = [&](int m) {
    const uint8_t* src = simg + (size_t)m * NST * SB;
#pragma unroll 1
    for (int j = 0; j < NST; ++j, ++sc) {
      const uint32_t stage = sc % C::NSTAGE, round = sc / C::NSTAGE;
      wait_or_die(&ctl.empty[stage], (round & 1u) ^ 1u, dg, 1);
      mbar_expect_tx(&ctl.full[stage], SB);
      bulk_g2s(sm.ring + (size_t)stage * SB, src + (size_t)j * SB, SB, &ctl.full[stage]);
    }
  };
  const TileList tl(a, worker, n_workers);
  for (int ti = 0; ti < tl.n; ++ti) {
    const int64_t tile = tl.list[ti];
    const int kmax = a.tile_kmax[tile];
    if (!BWD) {
      for (int l = 1; l <= L; ++l) load(mat_jump(L, l));
      for (int l = 0; l < L; ++l) load(mat_out(L, l));
      for (int k = 0; k < kmax; ++k)
        for (int l = 0; l <= L; ++l) load(mat_ode(L, l));
      for (int l = 0; l < L; ++l) load(mat_out(L, l));
    } else {
      for (int l = L - 1; l >= 0; --l) load(mat_out(L, l));
      for (int k = 0; k < kmax; ++k)
        for (int l = L; l >= 0; --l) load(mat_ode(L, l));
      for (int l = L - 1; l >= 0; --l) load(mat_out(L, l));
      for (int l = L; l >= 1; --l) load(mat_jump(L, l));
    }
  }
}

template <int HW, bool BWD>
__device__ __forceinline__ void issuer(const SweepArgs& a, uint8_t* raw) {
  using C = Cfg<HW>;
  Smem<HW> sm(raw);
  Ctl& ctl = *sm.ctl;
  const int S = a.T.S, L = a.T.L;
  const int worker = blockIdx.x / S, n_workers = gridDim.x / S;
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(&ctl.tmem_base);
  const uint32_t ring_s = umma::smem_u32(sm.ring);
  constexpr uint32_t idesc = umma::idesc_tf32(128, HW, 0, 0);
  Diag dg{g_status, g_notrap, BWD ? 8u : 4u, false};
  PH_DECL_AT(NT_W);
  uint32_t sc = 0, gi = 0;
  const TileList tl(a, worker, n_workers);
  for (int ti = 0; ti < tl.n; ++ti) {
    const int64_t tile = tl.list[ti];
    const int n_gemm = 3 * L + a.tile_kmax[tile] * (L + 1);
#pragma unroll 1
    for (int e = 0; e < n_gemm; ++e, ++gi) {
      if (BLOCKED) {
        constexpr uint32_t idesc_b = umma::idesc_tf32(128, C::CG, 0, 0);
        PH(2);
        wait_or_die(&ctl.ops[0], gi & 1u, dg, 2);              // the whole A operand of this GEMM is in TMEM
        PH(0);
#pragma unroll 1
        for (int nb = 0; nb < C::NBLK; ++nb, ++sc) {
          const uint32_t stage = sc % C::NSTAGE, sround = sc / C::NSTAGE;
          wait_or_die(&ctl.full[stage], sround & 1u, dg, 3);   // the block's weights are in the ring
          PH(1);
          umma::fence_after_sync();
          if (umma::elect_one()) {
            const uint32_t acc = tmem + C::ACC0 + nb * C::CG;
            const uint32_t sb = ring_s + stage * C::BSTAGE_BYTES;
            const uint32_t a_hi = tmem + C::A_HI, a_lo = tmem + C::A_LO;
            // k-step t = A columns 8t..8t+8 = k-chunk t / 4 (a [CG rows][32 k] tile), 32-byte step t % 4 inside it
#pragma unroll
            for (int t = 0; t < HW / 8; ++t)
              umma::mma_ts(acc, a_lo + 8 * t, umma::desc_k(sb + (t >> 2) * (C::CG * 128)) + 2 * (t & 3), idesc_b, t > 0 ? 1u : 0u);
#pragma unroll
            for (int t = 0; t < HW / 8; ++t)
              umma::mma_ts(acc, a_hi + 8 * t, umma::desc_k(sb + C::BSTAGE_HALF + (t >> 2) * (C::CG * 128)) + 2 * (t & 3), idesc_b, 1u);
#pragma unroll
            for (int t = 0; t < HW / 8; ++t)
              umma::mma_ts(acc, a_hi + 8 * t, umma::desc_k(sb + (t >> 2) * (C::CG * 128)) + 2 * (t & 3), idesc_b, 1u);
            umma::commit(&ctl.empty[stage]);                   // stage free once these MMAs have read it
            umma::commit(&ctl.accd[nb]);                       // accumulator block nb complete
          }
          __syncwarp();
          PH(2);
        }
        continue;
      }
      const uint32_t acc = tmem + C::ACC0 + (gi & 1u) * HW;
#pragma unroll 1
      for (int j = 0; j < C::NSUB; ++j, ++sc) {
        const uint32_t stage = sc % C::NSTAGE, sround = sc / C::NSTAGE;
        PH(2);
        wait_or_die(&ctl.ops[j], gi & 1u, dg, 2);            // the A columns of this sub-step are in TMEM
        PH(0);
        wait_or_die(&ctl.full[stage], sround & 1u, dg, 3);   // its weights are in the ring
        PH(1);
        umma::fence_after_sync();
        if (umma::elect_one()) {
          const uint64_t dbh = umma::desc_k(ring_s + stage * C::STAGE_BYTES);
          const uint64_t dbl = umma::desc_k(ring_s + stage * C::STAGE_BYTES + C::STAGE_HALF);
          const uint32_t a_hi = tmem + C::A_HI + 8 * j, a_lo = tmem + C::A_LO + 8 * j;
#pragma unroll
          for (int g = 0; g < 4; ++g) umma::mma_ts(acc, a_lo + g * C::CG, dbh + 2 * g, idesc, (j > 0 || g > 0) ? 1u : 0u);
#pragma unroll
          for (int g = 0; g < 4; ++g) umma::mma_ts(acc, a_hi + g * C::CG, dbl + 2 * g, idesc, 1u);
#pragma unroll
          for (int g = 0; g < 4; ++g) umma::mma_ts(acc, a_hi + g * C::CG, dbh + 2 * g, idesc, 1u);
          umma::commit(&ctl.empty[stage]);                            // stage free once these MMAs have read it
          if (j == C::NSUB - 1) umma::commit(&ctl.accd[gi & 1u]);     // accumulator complete
        }
        __syncwarp();
      }
    }
  }
  PH(2);
  PH_STORE(g_phase_iss);
}

// per-thread view of the worker role
template <int HW>
struct WorkerCtx {
  using C = Cfg<HW>;
  Ctl& ctl;
  SmallW<HW>& sw;
  int lane, q, g, row, col0;
  uint32_t lane_base;      // TMEM address of (this lane quadrant, column col0)
  uint32_t gi;             // GEMM counter (same sequence as the issuer's)
  float comp;              // accumulator read-back scale (SweepArgs::comp_chain)
  Diag dg;
  __device__ __forceinline__ WorkerCtx(Smem<HW>& sm, unsigned bit_) : ctl(*sm.ctl), sw(*sm.sw), gi(0), dg{g_status, g_notrap, bit_, false} {
    const int warp = threadIdx.x >> 5;
    lane = threadIdx.x & 31;
    q = warp & 3;
    g = warp >> 2;
    row = q * 32 + lane;
    col0 = g * C::CG;
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(&ctl.tmem_base);
    lane_base = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)col0;
  }
  // Both directions of the TMEM traffic are software-pipelined: an epilogue is a chain of short dependent steps
  // (tcgen05.ld -> wait -> arithmetic -> tcgen05.st -> wait -> fence -> arrive), and run back to back every sub-step
  // cost ~1000 cycles whatever its arithmetic (phase build: the reverse sweep, a tenth of the forward's arithmetic,
  // had the same epilogue time).  So the accumulator chunk j + 1 is requested before chunk j is processed, and the
  // hand-over of chunk j (wait::st, fence, arrive) is issued after the arithmetic of chunk j + 1, when its stores
  // have long landed.  wait_acc() flushes the last pending hand-over: every emission is followed by one.
  float nx[8];             // accumulator chunk in flight
  int pend = -1;           // emitted sub-chunk whose hand-over is still to be signalled
  // When a sub-chunk's hand-over is signalled (measured on B200, forward + reverse sweep):
  //   0: at the next emission (one sub-chunk of arithmetic later)          H=128: 3.98 + 4.46 ms   H=64: 1.49 + 1.82 ms
  //   1: after the next accumulator load has been waited for               H=128: 3.71 + 4.49      H=64: 1.41 + 1.84
  //   2: at once, behind its own tcgen05.st (the worker eats the latency)  H=128: 3.70 + 4.57      H=64: 1.36 + 1.79
  static constexpr int ES = NJODE_WIDE_EARLY_SIGNAL == 0 ? 0 : (HW == 64 ? 2 : 1);
  // Blocked order: the A operand region is being read until the current GEMM's last block is done.  The warps of the
  // last column group emit straight into it (their accumulator wait IS that event); the others park their raw FP32
  // operand in the spare TMEM columns (the k-major order's second accumulator) and move it over in flush_emits().
  static constexpr uint32_t STASH = C::ACC0 + HW;
  bool have = false;
  __device__ __forceinline__ void flush_emits() {
    if (!have) return;
    have = false;
    if (g == C::NBLK - 1) {
      umma::wait_st();
      umma::fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive_relaxed(&ctl.ops[0]);
      return;
    }
    umma::wait_st();                                             // the parked operand is in TMEM
    // every MMA that reads the A operand region must be complete: the last block of the previous GEMM of the sequence
    if (gi > 0) { wait_or_die(&ctl.accd[C::NBLK - 1], (gi - 1) & 1u, dg, 5); umma::fence_after_sync(); }
    float v[8], vn[8];
    umma::tmem_ld8_nowait(lane_base + STASH, vn);
#pragma unroll
    for (int j = 0; j < C::NSUB; ++j) {
      asm volatile("tcgen05.wait::ld.sync.aligned;" : "+f"(vn[0]), "+f"(vn[1]), "+f"(vn[2]), "+f"(vn[3]), "+f"(vn[4]), "+f"(vn[5]), "+f"(vn[6]), "+f"(vn[7]) :: "memory");
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = vn[i];
      if (j + 1 < C::NSUB) umma::tmem_ld8_nowait(lane_base + STASH + 8 * (j + 1), vn);
      uint32_t hi[8], lo[8];
      umma::split8(v, hi, lo);
      umma::tmem_st8_raw(lane_base + C::A_HI + 8 * j, hi);
      umma::tmem_st8_raw(lane_base + C::A_LO + 8 * j, lo);
    }
    umma::wait_st();
    umma::fence_before_sync();
    __syncwarp();
    if (lane == 0) mbar_arrive_relaxed(&ctl.ops[0]);
  }
  __device__ __forceinline__ void signal_pending() {
    if (BLOCKED) return;
    if (pend >= 0) {
      umma::wait_st();
      umma::fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive_relaxed(&ctl.ops[pend]);
      pend = -1;
    }
  }
  // sub-chunk j of the next GEMM's A operand: split, store to TMEM; the issuer is told one sub-chunk later
  __device__ __forceinline__ void emit(int j, const float (&v)[8]) {
    if (BLOCKED) {
      if (g == C::NBLK - 1) {
        uint32_t hi[8], lo[8];
        umma::split8(v, hi, lo);
        umma::tmem_st8_raw(lane_base + C::A_HI + 8 * j, hi);
        umma::tmem_st8_raw(lane_base + C::A_LO + 8 * j, lo);
      } else {
        uint32_t raw[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) raw[i] = __float_as_uint(v[i]);
        umma::tmem_st8_raw(lane_base + STASH + 8 * j, raw);
      }
      have = true;
      return;
    }
    uint32_t hi[8], lo[8];
    umma::split8(v, hi, lo);
    signal_pending();
    umma::tmem_st8_raw(lane_base + C::A_HI + 8 * j, hi);
    umma::tmem_st8_raw(lane_base + C::A_LO + 8 * j, lo);
    pend = j;
    if (ES == 2) signal_pending();          // at once: the worker eats the store latency, the issuer starts a sub-chunk earlier
  }
  // wait for the accumulator of GEMM gi (call once per GEMM, then acc_ld for sub-chunks 0, 1, .. in order, then done())
  __device__ __forceinline__ void wait_acc() {
    if (BLOCKED) {
      flush_emits();
      wait_or_die(&ctl.accd[g], gi & 1u, dg, 4);               // this column group's block of GEMM gi
      umma::fence_after_sync();
      umma::tmem_ld8_nowait(lane_base + C::ACC0, nx);
      return;
    }
    signal_pending();
    wait_or_die(&ctl.accd[gi & 1u], (gi >> 1) & 1u, dg, 4);
    umma::fence_after_sync();
    umma::tmem_ld8_nowait(lane_base + C::ACC0 + (gi & 1u) * HW, nx);
  }
  __device__ __forceinline__ void acc_ld(int j, float (&v)[8]) {
    // (the registers are operands of the wait so that no use of them can be scheduled in front of it)
    asm volatile("tcgen05.wait::ld.sync.aligned;" : "+f"(nx[0]), "+f"(nx[1]), "+f"(nx[2]), "+f"(nx[3]), "+f"(nx[4]), "+f"(nx[5]), "+f"(nx[6]), "+f"(nx[7]) :: "memory");
    // the previous sub-chunk's operand stores were issued before this load was waited for: they have landed, and the
    // issuer can start that sub-step's MMAs now rather than after this sub-chunk's arithmetic
    if (ES == 1) signal_pending();
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = nx[i] * comp;
    if (j + 1 < C::NSUB) umma::tmem_ld8_nowait(lane_base + C::ACC0 + (BLOCKED ? 0u : (gi & 1u) * HW) + 8 * (j + 1), nx);
  }
  __device__ __forceinline__ void done() { ++gi; }
};

template <int HW>
__device__ __forceinline__ void load_small(SmallW<HW>& sw, const ParamTable& T, const float* __restrict__ p) {
  const int L = T.L;
  for (int t = threadIdx.x; t < HW; t += NT_W) {
    for (int l = 0; l <= L; ++l) {
      sw.b_jump[l][t] = p[T.b_off[NET_JUMP][l] + t];
      sw.b_ode[l][t] = p[T.b_off[NET_ODE][l] + t];
      if (l < L) sw.b_out[l][t] = p[T.b_off[NET_OUT][l] + t];
    }
    const int ld0 = HW + T.d_x + 2;
    for (int e = 0; e < T.d_x + 2; ++e) sw.ext_ode0[e][t] = p[T.w_off[NET_ODE][0] + t * ld0 + HW + e];
    for (int e = 0; e < T.d_x; ++e) sw.w_jump0[e][t] = p[T.w_off[NET_JUMP][0] + t * T.d_x + e];
    for (int o = 0; o < T.O; ++o) sw.w_out[o][t] = p[T.w_off[NET_OUT][L] + o * HW + t];
    if (t < T.O) sw.b_outL[t] = p[T.b_off[NET_OUT][L] + t];
  }
}

// Activation of the epilogues.  tanhf() is ~20 instructions per element on two divergent paths.  A 6-instruction
// branch-free form, tanh(a) = sign(a) * (1 - 2 / (2^(2 |a| log2 e) + 1)) on MUFU.EX2 / MUFU.RCP, is available as a build
// option (NJODE_WIDE_EXACT_TANH=0).  Measured on B200, H = 128 / L = 3 / tanh: forward sweep 3.98 -> 3.68 ms (whole
// step -2.4 %), every golden still inside 1e-5, but the prediction error against the float64 oracle goes from 5e-7 to
// 2.0e-6 (absolute error 2.5e-7 per activation, relative error unbounded near 0).  Not worth a fifth of the parity
// margin: the product uses tanhf.
#ifndef NJODE_WIDE_EXACT_TANH
#define NJODE_WIDE_EXACT_TANH 1
#endif
template <int ACT>
__device__ __forceinline__ float act_w(float a) {
  if (ACT == NJODE_ACT_TANH && !NJODE_WIDE_EXACT_TANH) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fabsf(a) * 2.8853900817779268f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
    return copysignf(fmaf(-2.0f, r, 1.0f), a);
  }
  return act_fwd<ACT>(a);
}

__device__ __forceinline__ void scale8(int sc, float (&v)[8]) {
  if (sc == NJODE_SCALE_TANH) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = tanhf(v[i]);
  } else if (sc == NJODE_SCALE_SIGMOID) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = 1.0f / (1.0f + expf(-v[i]));
  }
}

// phase accounting (`make phase`): bucket 0 = everything between accumulator waits (epilogues, emissions, readouts),
// bucket 1 = waiting for an accumulator (the MMA tail the epilogue cannot hide + barrier latency)
#define WAIT_ACC() do { PH(0); w.wait_acc(); PH(1); } while (0)

// ------------------------------------------------------------------------------------------------
// forward sweep: row workers
// ------------------------------------------------------------------------------------------------
template <int HW, int ACT>
__device__ __forceinline__ void fwd_worker(const SweepArgs& a, uint8_t* raw) {
  using C = Cfg<HW>;
  constexpr int CG = C::CG, NSUB = C::NSUB;
  Smem<HW> sm(raw);
  WorkerCtx<HW> w(sm, 4u);
  w.comp = a.comp_chain;
  PH_DECL;
  SmallW<HW>& sw = w.sw;
  const ParamTable& T = a.T;
  const int L = T.L, dx = T.d_x, O = T.O, sc_kind = a.desc.input_scaling;
  const int s = blockIdx.x % T.S, worker = blockIdx.x / T.S, n_workers = gridDim.x / T.S;
  const int row = w.row, col0 = w.col0, g = w.g;
  const int64_t PL = C::PL, slotf = (int64_t)(L + 1) * PL;
  float* const ckpt_s = a.ckpt ? a.ckpt + (int64_t)s * a.total_slots * slotf + aplane_off(HW, col0 >> 3, row) : nullptr;

  const TileList tl(a, worker, n_workers);
  for (int ti = 0; ti < tl.n; ++ti) {
    const int64_t tile = tl.list[ti];
    const int kmax = a.tile_kmax[tile];
    const int u = a.perm[tile * R + row];
    float* const ck = ckpt_s ? ckpt_s + a.tile_slot_off[tile] * slotf : nullptr;
    float* const ckd = a.ckpt ? a.ckpt + ((int64_t)T.S * a.total_slots + (int64_t)s * a.total_slots + a.tile_slot_off[tile]) * slotf +
                                    dplane_off(HW, col0, row) : nullptr;       // half D, this thread's (first feature, row)
    // plane `pl` of slot `sl`, this thread's sub-chunk j
    auto cp = [&](int sl, int pl, int j) { return ck + ((int64_t)sl * (L + 1) + pl) * PL + j * 256; };      // (next chunk: + 32 rows x 8)
    const float* const kn = a.knots + a.tile_slot_off[tile] * R + row;
    const int ke = u >= 0 ? a.kenc[u] : 0;
    const int K = ke >> 1;
    const int X1 = kmax + 1, X2 = kmax + 2, X3 = kmax + 3;
    float x[MAX_DX], xs[MAX_DX];
#pragma unroll
    for (int e = 0; e < MAX_DX; ++e) {
      x[e] = (e < dx && u >= 0) ? a.values[(int64_t)u * dx + e] : 0.0f;
      xs[e] = scale_fwd_rt(sc_kind, x[e]);
    }
    float h[CG];

    // ---- h = jump(x)                                                  jump_ode.py:169 / :176 ----
#pragma unroll
    for (int j = 0; j < NSUB; ++j) {
      float z[8], cw[8];
      ld8s(sw.b_jump[0] + col0 + 8 * j, z);
#pragma unroll
      for (int e = 0; e < MAX_DX; ++e) if (e < dx) {
        ld8s(sw.w_jump0[e] + col0 + 8 * j, cw);
#pragma unroll
        for (int i = 0; i < 8; ++i) z[i] = fmaf(cw[i], x[e], z[i]);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) z[i] = act_w<ACT>(z[i]);
      w.emit(j, z);
      if (ck) st8g(cp(X3, 0, j), z);
    }
    for (int l = 1; l <= L; ++l) {
      WAIT_ACC();
#pragma unroll
      for (int j = 0; j < NSUB; ++j) {
        float z[8], cb[8];
        w.acc_ld(j, z);
        ld8s(sw.b_jump[l] + col0 + 8 * j, cb);
#pragma unroll
        for (int i = 0; i < 8; ++i) z[i] = act_w<ACT>(z[i] + cb[i]);
        w.emit(j, z);                   // next jump layer, or (l == L) the first out-net layer on h_0
        // (checkpoint stores come AFTER the hand-over: its release-arrive waits for every earlier global store of the
        //  thread to be acknowledged -- an L2 round trip per sub-step when the store goes first)
        if (l < L) {
          if (ck) st8g(cp(X3, l, j), z);
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) h[8 * j + i] = z[i];
          if (ck) st8g(cp(0, 0, j), z);
        }
      }
      w.done();
    }

    // ---- y = out(h): the A operand (h) has been emitted; L chain GEMMs, then the readout dot products ----
    auto readout = [&](int X, float* __restrict__ dst, int64_t obs, bool write) {
      float y[MAX_O];
#pragma unroll
      for (int o = 0; o < MAX_O; ++o) y[o] = 0.0f;
      for (int l = 0; l < L; ++l) {
        WAIT_ACC();
#pragma unroll
        for (int j = 0; j < NSUB; ++j) {
          float z[8], cb[8];
          w.acc_ld(j, z);
          ld8s(sw.b_out[l] + col0 + 8 * j, cb);
#pragma unroll
          for (int i = 0; i < 8; ++i) z[i] = act_w<ACT>(z[i] + cb[i]);
          if (l < L - 1) w.emit(j, z);
          if (ck) {
            st8g(cp(X, l + 1, j), z);
            // the readout layer's weight gradient contracts THIS activation with dY: the weight-gradient GEMM wants it as
            // its TMEM (feature-major) operand, so a second copy goes to half D, plane L of the slot
            if (l == L - 1) st8t(ckd + ((int64_t)X * (L + 1) + L) * PL + j * 64, z);
          }
          if (l < L - 1) {
          } else {
#pragma unroll
            for (int o = 0; o < MAX_O; ++o) if (o < O) {
              ld8s(sw.w_out[o] + col0 + 8 * j, cb);
#pragma unroll
              for (int i = 0; i < 8; ++i) y[o] = fmaf(z[i], cb[i], y[o]);
            }
          }
        }
        w.done();
      }
      *reinterpret_cast<float4*>(sw.red[g][row]) = make_float4(y[0], y[1], y[2], y[3]);
      umma::named_bar_sync(1, NT_W);
      if (g == 0 && write) {
#pragma unroll
        for (int o = 0; o < MAX_O; ++o) if (o < O)
          dst[pred_index(T, obs, s, o)] = sw.b_outL[o] + ((sw.red[0][row][o] + sw.red[1][row][o]) + (sw.red[2][row][o] + sw.red[3][row][o]));
      }
    };
    readout(X1, a.preds, u, u >= 0);

    // A operand of the next GEMM from the state in registers: s(h) for an Euler step, h itself for the readout
    auto emit_state = [&](bool scaled) {
#pragma unroll
      for (int j = 0; j < NSUB; ++j) {
        float z[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) z[i] = h[8 * j + i];
        if (scaled) scale8(sc_kind, z);
        w.emit(j, z);
      }
    };
    emit_state(kmax > 0);

    // ---- Euler steps with x held constant                              jump_ode.py:188-203, :122-140 ----
    float tn = ldg_na(kn);
    float tn_ahead = kmax > 0 ? ldg_na(kn + R) : tn;
    for (int k = 0; k < kmax; ++k) {
      const float tc = tn;
      tn = tn_ahead;
      const float tn_loaded = ldg_na(kn + (k + 2 <= kmax ? k + 2 : kmax) * R);
      const float delta = __fsub_rn(tn, tc);
      // first layer: bias + the x / t_cur / dt columns on the CUDA cores      jump_ode.py:57-61
      WAIT_ACC();
#pragma unroll
      for (int j = 0; j < NSUB; ++j) {
        float z[8], cw[8];
        w.acc_ld(j, z);
        ld8s(sw.b_ode[0] + col0 + 8 * j, cw);
#pragma unroll
        for (int i = 0; i < 8; ++i) z[i] += cw[i];
#pragma unroll
        for (int e = 0; e < MAX_DX; ++e) if (e < dx) {
          ld8s(sw.ext_ode0[e] + col0 + 8 * j, cw);
#pragma unroll
          for (int i = 0; i < 8; ++i) z[i] = fmaf(cw[i], xs[e], z[i]);
        }
        ld8s(sw.ext_ode0[dx] + col0 + 8 * j, cw);
#pragma unroll
        for (int i = 0; i < 8; ++i) z[i] = fmaf(cw[i], tc, z[i]);
        ld8s(sw.ext_ode0[dx + 1] + col0 + 8 * j, cw);
#pragma unroll
        for (int i = 0; i < 8; ++i) z[i] = act_w<ACT>(fmaf(cw[i], delta, z[i]));
        w.emit(j, z);
        if (ck) st8g(cp(k, 1, j), z);
      }
      w.done();
      for (int l = 1; l < L; ++l) {
        WAIT_ACC();
#pragma unroll
        for (int j = 0; j < NSUB; ++j) {
          float z[8], cb[8];
          w.acc_ld(j, z);
          ld8s(sw.b_ode[l] + col0 + 8 * j, cb);
#pragma unroll
          for (int i = 0; i < 8; ++i) z[i] = act_w<ACT>(z[i] + cb[i]);
          w.emit(j, z);
          if (ck) st8g(cp(k, l + 1, j), z);
        }
        w.done();
      }
      // last layer (no activation) and the Euler update                      jump_ode.py:130-131 / :138-139
      WAIT_ACC();
      const bool more = k + 1 < kmax;
#pragma unroll
      for (int j = 0; j < NSUB; ++j) {
        float f[8], cb[8];
        w.acc_ld(j, f);
        ld8s(sw.b_ode[L] + col0 + 8 * j, cb);
        if (k < K) {
#pragma unroll
          for (int i = 0; i < 8; ++i) h[8 * j + i] = fmaf(delta, f[i] + cb[i], h[8 * j + i]);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = h[8 * j + i];
        if (more && sc_kind != NJODE_SCALE_IDENTITY) {
          float fs[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) fs[i] = f[i];
          scale8(sc_kind, fs);
          w.emit(j, fs);
        } else {
          w.emit(j, f);
        }
        if (ck) st8g(cp(k + 1, 0, j), f);
      }
      w.done();
      asm volatile("mov.f32 %0, %1;" : "=f"(tn_ahead) : "f"(tn_loaded));
    }
    readout(X2, a.preds_before, (int64_t)u + 1, u >= 0 && (ke & 1));
  }
  PH(0);
  PH_STORE(g_phase12);
}

// ------------------------------------------------------------------------------------------------
// reverse sweep: row workers.  Propagates d loss / d (pre-activation) through every layer, last to first, and
// writes each of them as a "d plane" (half D of the checkpoint buffer) for the weight-gradient GEMM.
// ------------------------------------------------------------------------------------------------
template <int HW, int ACT>
__device__ __forceinline__ void bwd_worker(const SweepArgs& a, uint8_t* raw) {
  using C = Cfg<HW>;
  constexpr int CG = C::CG, NSUB = C::NSUB;
  Smem<HW> sm(raw);
  WorkerCtx<HW> w(sm, 8u);
  w.comp = a.comp_chain;
  PH_DECL;
  SmallW<HW>& sw = w.sw;
  const ParamTable& T = a.T;
  const int L = T.L, O = T.O, sc_kind = a.desc.input_scaling;
  const int s = blockIdx.x % T.S, worker = blockIdx.x / T.S, n_workers = gridDim.x / T.S;
  const int row = w.row, col0 = w.col0;
  const int64_t PL = C::PL, slotf = (int64_t)(L + 1) * PL;
  const int64_t half = (int64_t)T.S * a.total_slots * slotf;             // floats of half A
  const int64_t toff = (int64_t)s * a.total_slots * slotf + aplane_off(HW, col0 >> 3, row);

  const TileList tl(a, worker, n_workers);
  for (int ti = 0; ti < tl.n; ++ti) {
    const int64_t tile = tl.list[ti];
    const int kmax = a.tile_kmax[tile];
    const int u = a.perm[tile * R + row];
    const int ke = u >= 0 ? a.kenc[u] : 0;
    const float* const ca = a.ckpt + toff + a.tile_slot_off[tile] * slotf;       // activations (read)
    float* const cd = a.ckpt + half + (int64_t)s * a.total_slots * slotf + a.tile_slot_off[tile] * slotf +
                      dplane_off(HW, col0, row);                              // d planes (written; layout: njode_wide.cuh)
    auto pa = [&](int sl, int pl, int j) { return ca + ((int64_t)sl * (L + 1) + pl) * PL + j * 256; };
    // (half D planes are [row octet][feature][8 rows]: this thread's (first feature of sub-chunk j, row))
    auto pd = [&](int sl, int pl, int j) { return cd + ((int64_t)sl * (L + 1) + pl) * PL + j * 64; };
    // aux rows of a slot (8 floats per row, written by the column-group-0 thread of the row): the extra B columns of
    // the weight-gradient GEMM -- (1, s(x).., t, dt) for an Euler step, (1, dY..) for a readout, (1, x..) for the jump
    float* const cx = a.ckpt + 2 * half + ((int64_t)s * a.total_slots + a.tile_slot_off[tile]) * (R * 8) + row * 8;
    const float* const kn = a.knots + a.tile_slot_off[tile] * R + row;
    const int X1 = kmax + 1, X2 = kmax + 2, X3 = kmax + 3;
    float xr[MAX_DX], xs[MAX_DX];
#pragma unroll
    for (int e = 0; e < MAX_DX; ++e) {
      xr[e] = (e < T.d_x && u >= 0) ? a.values[(int64_t)u * T.d_x + e] : 0.0f;
      xs[e] = scale_fwd_rt(sc_kind, xr[e]);
    }
    float gr[CG];                      // d loss / d h (this thread's columns), running backwards in time
#pragma unroll
    for (int i = 0; i < CG; ++i) gr[i] = 0.0f;

    // One inner layer of a reverse chain: d_out = (accumulator = d_in * W) * act'(z), written as a d plane and handed on
    // as the next GEMM's operand.  ALL of the layer's z chunks are requested before the accumulator wait: the MMAs
    // (>= 3000 cycles at H = 128) hide the HBM latency, whereas a load issued inside the sub-step loop is consumed a
    // few hundred cycles later (measured: long-scoreboard stalls on act'(z) were the top stall of this kernel).
    auto chain_layer = [&](const float* __restrict__ zsrc, float* __restrict__ ddst, bool emit_next) {
      // (each load writes its final registers: through a temporary the compiler re-used one register octet for all
      //  chunks and every load waited for the previous one to land -- four serial HBM round trips per layer)
      float zb[CG];
#pragma unroll
      for (int j = 0; j < NSUB; ++j) ld8g(zsrc + j * 256, *reinterpret_cast<float(*)[8]>(&zb[8 * j]));
      WAIT_ACC();
#pragma unroll
      for (int j = 0; j < NSUB; ++j) {
        float acc[8];
        w.acc_ld(j, acc);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] *= act_grad_from_out<ACT>(zb[8 * j + i]);
        if (emit_next) w.emit(j, acc);
        st8t(ddst + j * 64, acc);
      }
      w.done();
    };

    // ---- readout backward at a hidden state: adds d loss / d h to gr ----
    auto out_backward = [&](int X, const float* __restrict__ gsrc, int64_t obs, bool live) {
      float dY[MAX_O];
#pragma unroll
      for (int o = 0; o < MAX_O; ++o) dY[o] = (live && o < O) ? gsrc[pred_index(T, obs, s, o)] : 0.0f;
      if (w.g == 0) {
        static_assert(MAX_O == 4, "aux row layout assumes <= 4 readout columns");
        const float xv[8] = {1.0f, dY[0], dY[1], dY[2], dY[3], 0.0f, 0.0f, 0.0f};
        st8g(cx + (int64_t)X * (R * 8), xv);
      }
      // d (pre-activation of out layer L-1) = (sum_o dY[o] * w_out[o][:]) * act'(z_L)
      float zl[CG];
#pragma unroll
      for (int j = 0; j < NSUB; ++j) ld8g(pa(X, L, j), *reinterpret_cast<float(*)[8]>(&zl[8 * j]));      // all chunks in flight at once
#pragma unroll
      for (int j = 0; j < NSUB; ++j) {
        float z[8], d[8], cw[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) z[i] = zl[8 * j + i];
#pragma unroll
        for (int i = 0; i < 8; ++i) d[i] = 0.0f;
#pragma unroll
        for (int o = 0; o < MAX_O; ++o) if (o < O) {
          ld8s(sw.w_out[o] + col0 + 8 * j, cw);
#pragma unroll
          for (int i = 0; i < 8; ++i) d[i] = fmaf(dY[o], cw[i], d[i]);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) d[i] *= act_grad_from_out<ACT>(z[i]);
        w.emit(j, d);                                            // -> d * W_out[L-1]
        st8t(pd(X, L - 1, j), d);
      }
      for (int l = L - 1; l >= 1; --l) chain_layer(pa(X, l, 0), pd(X, l - 1, 0), true);      // -> d * W_out[l-1]
      WAIT_ACC();                                              // d * W_out[0] = d loss / d h through this readout
#pragma unroll
      for (int j = 0; j < NSUB; ++j) {
        float acc[8];
        w.acc_ld(j, acc);
#pragma unroll
        for (int i = 0; i < 8; ++i) gr[8 * j + i] += acc[i];
      }
      w.done();
    };

    // ---- preds_before[u+1] = out(h_end) ----
    out_backward(X2, a.grad_preds_before, (int64_t)u + 1, u >= 0 && (ke & 1));

    // ---- Euler steps, last to first ----
    float tn = kmax > 0 ? ldg_na(kn + kmax * R) : 0.0f;
    float tc_next = kmax > 0 ? ldg_na(kn + (kmax - 1) * R) : 0.0f;
    float delta_k = 0.0f;                                        // dt of the step being differentiated
    if (kmax > 0) {
      // d loss / d f of the last step: delta * g
      const float delta = __fsub_rn(tn, tc_next);
      delta_k = delta;
#pragma unroll
      for (int j = 0; j < NSUB; ++j) {
        float d[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) d[i] = delta * gr[8 * j + i];
        w.emit(j, d);                                            // -> d * W_ode[L]
        st8t(pd(kmax - 1, L, j), d);
      }
    }
    for (int k = kmax - 1; k >= 0; --k) {
      tn = tc_next;                                              // t_k
      tc_next = ldg_na(kn + (k > 0 ? k - 1 : 0) * R);            // t_{k-1}
      if (w.g == 0) {
        static_assert(MAX_DX == 2, "aux row layout assumes d_x <= 2");
        const bool two = T.d_x > 1;
        const float xv[8] = {1.0f, xs[0], two ? xs[1] : tn, two ? tn : delta_k, two ? delta_k : 0.0f, 0.0f, 0.0f, 0.0f};
        st8g(cx + (int64_t)k * (R * 8), xv);
      }
      for (int l = L; l >= 1; --l) chain_layer(pa(k, l, 0), pd(k, l - 1, 0), true);           // -> d * W_ode[l-1]
      // d * W_ode[0][:, :H] = d loss / d s(h_k); fused with the next step's first operand delta_{k-1} * g
      const float delta_prev = __fsub_rn(tn, tc_next);           // (unused when k == 0)
      WAIT_ACC();
#pragma unroll
      for (int j = 0; j < NSUB; ++j) {
        float acc[8];
        w.acc_ld(j, acc);
        if (sc_kind != NJODE_SCALE_IDENTITY) {
          float hs[8];
          ld8g(pa(k, 0, j), hs);
          scale8(sc_kind, hs);
#pragma unroll
          for (int i = 0; i < 8; ++i) gr[8 * j + i] = fmaf(acc[i], scale_grad_rt(sc_kind, hs[i]), gr[8 * j + i]);
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) gr[8 * j + i] += acc[i];
        }
        if (k > 0) {
          float d[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) d[i] = delta_prev * gr[8 * j + i];
          w.emit(j, d);
          st8t(pd(k - 1, L, j), d);
        }
      }
      w.done();
      delta_k = delta_prev;
    }

    // ---- preds[u] = out(h_0), then the jump net ----
    out_backward(X1, a.grad_preds, u, u >= 0);
    if (w.g == 0) {
      const float xv[8] = {1.0f, xr[0], xr[1], 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};     // xr[1] = 0 when d_x = 1
      st8g(cx + (int64_t)X3 * (R * 8), xv);
    }
    // d (pre-activation of jump layer L) = g * act'(h_0)
    float h0[CG];
#pragma unroll
    for (int j = 0; j < NSUB; ++j) ld8g(pa(0, 0, j), *reinterpret_cast<float(*)[8]>(&h0[8 * j]));
#pragma unroll
    for (int j = 0; j < NSUB; ++j) {
      float d[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) d[i] = gr[8 * j + i] * act_grad_from_out<ACT>(h0[8 * j + i]);
      w.emit(j, d);                                              // -> d * W_jump[L]
      st8t(pd(X3, L, j), d);
    }
    for (int l = L; l >= 1; --l) chain_layer(pa(X3, l - 1, 0), pd(X3, l - 1, 0), l > 1);      // -> d * W_jump[l-1]
  }
  PH(0);
  PH_STORE(g_phase12);
}

template <int HW, int ACT, bool BWD>
__global__ void __launch_bounds__(NT, 1) k_wide_sweep(SweepArgs a, const float* __restrict__ img) {
  extern __shared__ uint8_t smem_raw[];
  using C = Cfg<HW>;
  const int warp = threadIdx.x >> 5;
  const long long t_start = clock64();
  {
    Smem<HW> sm(smem_raw);
    Ctl& ctl = *sm.ctl;
    if (threadIdx.x == 0) {
      for (int i = 0; i < C::NSTAGE; ++i) { umma::mbar_init(&ctl.full[i], 1); umma::mbar_init(&ctl.empty[i], 1); }
      for (int i = 0; i < 4; ++i) umma::mbar_init(&ctl.ops[i], NWARP_W);
      for (int i = 0; i < 4; ++i) umma::mbar_init(&ctl.accd[i], 1);
      umma::fence_mbar_init();
    }
    if (warp == 0) umma::tmem_alloc(&ctl.tmem_base, C::TMEM_COLS);
    if (warp < NWARP_W) load_small<HW>(*sm.sw, a.T, a.params + (int64_t)(blockIdx.x % a.T.S) * a.T.stack_floats);
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
  }
  if (warp < NWARP_W) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
    if (BWD) bwd_worker<HW, ACT>(a, smem_raw); else fwd_worker<HW, ACT>(a, smem_raw);
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
    if (warp == NWARP_W) issuer<HW, BWD>(a, smem_raw);
    else if (threadIdx.x == NT_W + 32) producer<HW, BWD>(a, smem_raw, img);
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    Smem<HW> sm(smem_raw);
    umma::tmem_free(*reinterpret_cast<volatile uint32_t*>(&sm.ctl->tmem_base), C::TMEM_COLS);
  }
  if (threadIdx.x == 0 && blockIdx.x < 512) {
    unsigned long long work = 0;
    const int S = a.T.S, nw = gridDim.x / S;
    const TileList tl(a, blockIdx.x / S, nw);
    for (int ti = 0; ti < tl.n; ++ti) work += 3 * a.T.L + a.tile_kmax[tl.list[ti]] * (a.T.L + 1);
    g_cta_cycles[blockIdx.x][0] = (unsigned long long)(clock64() - t_start);
    g_cta_cycles[blockIdx.x][1] = work;
  }
}

static int debug_sync(const char* what, cudaStream_t st) {
  if (!njode_debug_sync_env()) return NJODE_OK;
  const cudaError_t e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) NJODE_FAIL(NJODE_ECUDA, "NJODE_DEBUG_SYNC: %s failed: %s", what, cudaGetErrorString(e));
  fprintf(stderr, "NJODE_DEBUG_SYNC: %s ok\n", what);
  return NJODE_OK;
}

template <int HW, int ACT>
int launch_wide(const SweepArgs& a, const float* img, cudaStream_t st, bool backward) {
  if (a.n_tiles == 0) return NJODE_OK;
  const size_t smem = Smem<HW>::bytes();
  if (njode_no_trap_env()) { const unsigned one = 1; NJODE_CUDA_OK(cudaMemcpyToSymbolAsync(g_notrap, &one, sizeof(one), 0, cudaMemcpyHostToDevice, st)); }
  if (backward) {
    NJODE_CUDA_OK(cudaFuncSetAttribute(k_wide_sweep<HW, ACT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    njode_timing_begin(2, st);
    k_wide_sweep<HW, ACT, true><<<a.n_workers, NT, smem, st>>>(a, img);
    njode_timing_end(2, st);
    NJODE_LAUNCH_OK("k_wide_sweep<reverse>");
    return debug_sync("k_wide_sweep<reverse>", st);
  } else {
    NJODE_CUDA_OK(cudaFuncSetAttribute(k_wide_sweep<HW, ACT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    njode_timing_begin(1, st);
    k_wide_sweep<HW, ACT, false><<<a.n_workers, NT, smem, st>>>(a, img);
    njode_timing_end(1, st);
    NJODE_LAUNCH_OK("k_wide_sweep<forward>");
    return debug_sync("k_wide_sweep<forward>", st);
  }
  return NJODE_OK;
}

template <int HW>
int dispatch_wide(const SweepArgs& a, const float* img, cudaStream_t st, bool backward) {
  switch (a.desc.activation) {
    case NJODE_ACT_RELU: return launch_wide<HW, NJODE_ACT_RELU>(a, img, st, backward);
    case NJODE_ACT_TANH: return launch_wide<HW, NJODE_ACT_TANH>(a, img, st, backward);
    case NJODE_ACT_SIGMOID: return launch_wide<HW, NJODE_ACT_SIGMOID>(a, img, st, backward);
    case NJODE_ACT_ELU: return launch_wide<HW, NJODE_ACT_ELU>(a, img, st, backward);
    case NJODE_ACT_LEAKY_RELU: return launch_wide<HW, NJODE_ACT_LEAKY_RELU>(a, img, st, backward);
    default: return launch_wide<HW, NJODE_ACT_SELU>(a, img, st, backward);
  }
}

template <int HW>
int prep_images(const SweepArgs& a, float* img, cudaStream_t st, int transpose) {
  const int64_t total = (int64_t)a.T.S * n_mats(a.T.L) * HW * HW;
  k_wide_prep<HW><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(a.T, a.params, img, transpose);
  NJODE_LAUNCH_OK("k_wide_prep");
  return NJODE_OK;
}

}  // namespace

int njode_wide_supported(const NjodeDesc* d) {
  const int O = d->shared_network ? d->d_y * d->num_moments : d->d_y;
  return (d->hidden == 64 || d->hidden == 128) && d->n_hidden_layers >= 1 && d->n_hidden_layers <= NJODE_WIDE_LMAX &&
         d->d_x <= MAX_DX && O <= MAX_O;
}

size_t njode_wide_image_bytes(const NjodeDesc* d) {
  const int S = d->shared_network ? 1 : d->num_moments;
  return njode_align_up((size_t)S * n_mats(d->n_hidden_layers) * 2 * d->hidden * d->hidden * sizeof(float), 1024);
}

static int sm_count_wide() {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms;
}

// one CTA per SM (it owns the shared memory and, at H = 128, all of TMEM); a multiple of S
int njode_wide_workers(const NjodeDesc* d, int64_t n_tiles) {
  const int S = d->shared_network ? 1 : d->num_moments;
  int64_t per_stack = sm_count_wide() / S;
  if (per_stack < 1) per_stack = 1;
  if (per_stack > n_tiles) per_stack = n_tiles > 0 ? n_tiles : 1;
  return (int)(per_stack * S);
}

// Accumulator truncation compensation (SweepArgs::comp_*).  Model: every tcgen05.mma adds its K = 8 products into the
// FP32 accumulator with round-toward-zero; a CPU simulation of exactly the MMA order used here (tools/sim_3xtf32.py)
// gives a mean relative shrink of 8.6e-7 for a 48-MMA sum (H = 128 chain GEMM, every weight-gradient plane pair) and
// 3.5e-7 for a 24-MMA one (H = 64 chain GEMM) -- 2.1e-8 per MMA that lands on a full-size accumulator, the drift
// round 1 measured on the hardware.  Calibrated on the B200 against the float64 oracle (tools/wide_debug.py, mean
// signed error of the large gradient entries with the correction off / on: -2.5e-6 / +1.3e-6 at H = 128, L = 3 and
// -1.1e-6 / +2.8e-7 at H = 64): the hardware loses ~0.7 of the model's figure, which is what is applied.
// NJODE_WIDE_COMP scales the correction (0 = off) for such calibration runs.
static void set_comp(SweepArgs& a) {
  static const float scale = [] { const char* e = getenv("NJODE_WIDE_COMP"); return e ? (float)atof(e) : 1.0f; }();
  a.comp_chain = 1.0f + scale * (a.desc.hidden == 128 ? 6.0e-7f : 2.7e-7f);
  a.comp_wgrad = 1.0f + scale * 5.5e-7f;
}

static int check_table(const SweepArgs& a) {
  if (a.n_tiles > 0 && (!a.tile_table || a.table_workers * a.T.S != a.n_workers))
    NJODE_FAIL(NJODE_EINVAL, "wide kernels: the schedule's tile table was built for %d workers per stack, the launch has %d",
               a.table_workers, a.n_workers / a.T.S);
  return NJODE_OK;
}

int njode_wide_forward(const SweepArgs& a_in, float* images, cudaStream_t st) {
  SweepArgs a = a_in;
  if (int rc0 = check_table(a)) return rc0;
  set_comp(a);
  int rc = a.desc.hidden == 128 ? prep_images<128>(a, images, st, 0) : prep_images<64>(a, images, st, 0);
  if (rc) return rc;
  return a.desc.hidden == 128 ? dispatch_wide<128>(a, images, st, false) : dispatch_wide<64>(a, images, st, false);
}

int njode_wide_backward(const SweepArgs& a_in, float* images, cudaStream_t st) {
  SweepArgs a = a_in;
  if (int rc0 = check_table(a)) return rc0;
  set_comp(a);
  int rc = a.desc.hidden == 128 ? prep_images<128>(a, images, st, 1) : prep_images<64>(a, images, st, 1);
  if (rc) return rc;
  rc = a.desc.hidden == 128 ? dispatch_wide<128>(a, images, st, true) : dispatch_wide<64>(a, images, st, true);
  if (rc) return rc;
  return njode_wide_wgrad(a, st);
}

extern "C" int njode_debug_cta_cycles(unsigned long long* out_host, int n_ctas) {
  if (!out_host || n_ctas < 1 || n_ctas > 512) NJODE_FAIL(NJODE_EINVAL, "njode_debug_cta_cycles: need 1..512 CTAs");
  NJODE_CUDA_OK(cudaDeviceSynchronize());
  NJODE_CUDA_OK(cudaMemcpyFromSymbol(out_host, g_cta_cycles, (size_t)n_ctas * 2 * sizeof(unsigned long long)));
  return NJODE_OK;
}

int njode_sweep_phase_fetch(unsigned long long* out_host, int n_ctas, int issuer) {
  if (issuer) NJODE_CUDA_OK(cudaMemcpyFromSymbol(out_host, g_phase_iss, (size_t)n_ctas * 8 * sizeof(unsigned long long)));
  else NJODE_CUDA_OK(cudaMemcpyFromSymbol(out_host, g_phase12, (size_t)n_ctas * 8 * sizeof(unsigned long long)));
  return NJODE_OK;
}

int njode_wide_sweep_status(unsigned* out_host) {
  NJODE_CUDA_OK(cudaMemcpyFromSymbol(out_host, g_status, 2 * sizeof(unsigned)));
  return NJODE_OK;
}
