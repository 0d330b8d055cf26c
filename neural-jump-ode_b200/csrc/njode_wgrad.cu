// njode_wgrad.cu -- every weight gradient of the wide flavour (hidden_dim 64 / 128) as ONE split-K tensor-core GEMM.
//
// After the two chain sweeps (njode_wide.cu) the checkpoint buffer holds, for every Linear layer application of the
// batch, the layer's input activations (half A) and d loss / d(pre-activation) (half D) as planes of 128 rows.  A
// weight gradient is a contraction over rows:   dW[out][in] = sum_rows D[row][out] * A[row][in]
// i.e. a GEMM whose K index is the row, for ~(kmax + 3) x (L + 1) plane pairs per tile:
//     A operand = D plane transposed [out][rows]  (M = 128), from TENSOR MEMORY: the reverse sweep writes its d planes
//                 as [row octet][feature][8 rows]; a converter thread takes the 8 rows of its feature from the staged
//                 stage (ld.shared.v4 x 2), splits them into tf32 hi / lo and stores them straight into TMEM
//                 (tcgen05.st), lane = output feature;
//     B operand = [A plane | aux columns] [rows][in + 16]  (N = H + 16), MN-major tf32 tiles in shared memory
//                 (SWIZZLE_128B_BASE32B, 128-byte rows of 32 features), split by the converters on the way in; the
//                 forward sweep writes its activation planes as [row group of 32][8-feature chunk][32 rows][8], so
//                 that the 32 rows of a stage are ONE contiguous block of the plane (one bulk copy).
// The aux columns (1, s(x).., t, dt | 1, dY.. | 1, x..; written per slot by the reverse sweep) make the bias, the
// x / t / dt columns of the first ODE layer, the readout weights and the first jump layer fall out of the same MMAs.
// The first version took BOTH operands from shared memory and was bound by its port (8.5 KB per MMA = 68 cycles at
// 128 B/cycle plus 72 KB of tile stores per stage: ~1400 cycles per 32-row stage against 864 of tensor time); with
// the D operand in TMEM a stage moves 40 KB + 54 KB through shared memory and the tensor pipe is the bound.
// FP32 accuracy: 3xTF32 (P_lo*Q_hi + P_hi*Q_lo + P_hi*Q_hi).  The tensor core's accumulate is not round-to-nearest
// (round 1: ~3000 MMAs into one accumulator drifted a gradient by 6e-5), so each plane pair (48 MMAs) goes into a
// FRESH TMEM accumulator (double buffered) that the CUDA cores add into running sums held in REGISTERS (36 per
// thread: TMEM is taken by the operand ring and the two accumulators) with IEEE adds; the running sums are flushed
// per category into this CTA's partial buffer (reduced in a fixed order by k_reduce_partials: deterministic for a
// given schedule, no atomics).
// Pipeline: a producer thread streams the RAW planes, three bulk copies (TMA engine) per 32-row stage, into a 3-stage
// staging ring; 16 converter warps read a staged stage from shared memory, split it and write the operand stage (TMEM +
// MN tiles, 3-stage ring); one issuer warp runs 12 MMAs per stage; full / empty mbarriers between all three.
// The staging ring exists because of one instruction: the converters' hand-over needs fence.proxy.async (generic ->
// async proxy), which is a MEMBAR that waits for every global load the thread has in flight -- with the planes loaded
// straight into registers a stage could not take less than a memory round trip (measured: 2600 cycles per stage at
// H = 128 against 864 of tensor time, whatever the prefetch depth).  Bulk copies have no thread-side loads.
#include "njode_wide.cuh"

namespace {
using namespace wide;

__device__ unsigned g_status[2] = {0, 0};
__device__ unsigned g_notrap = 0;
__device__ unsigned long long g_phase3[512][8];     // `make phase`: per-CTA phase cycles of the converter role (thread 0)

// Converter groups: the 16 converter warps work as NGRP groups on alternating stages (group = warp / (16 / NGRP)).  With
// one group every warp takes part in every stage and a stage cannot be shorter than one warp's serial chain (wait ->
// ld.shared -> split -> tcgen05.st / tile stores -> fences -> arrive: ~2000 cycles, mostly latencies); with two groups
// a warp converts twice the data per stage, every other stage, and the chains of consecutive stages overlap.
#ifndef NJODE_K3_GROUPS
#define NJODE_K3_GROUPS 2
#endif
constexpr int NGRP = NJODE_K3_GROUPS;
constexpr int GWARPS = NWARP_W / NGRP;     // warps per group (a multiple of 4: every group covers the 4 TMEM lane quadrants)
static_assert(NGRP == 1 || NGRP == 2 || NGRP == 4, "converter groups");      // (coprime with the ring depths: see Ctl3::raw_full)
constexpr int NT3 = NT_W + 128;        // 16 converter / merge warps + one warpgroup: MMA issuer warp, producer warp (setmaxnreg: 112 / 32)
constexpr int NSTAGE3 = 3;             // operand stages (TMEM ring + MN tiles)
constexpr int NRAW = 3;                // staging stages (raw planes)
constexpr int SROWS = 32;              // rows (K) per stage
constexpr int BLK = SROWS * 128;       // bytes of one 32-feature block of a stage tile
// TMEM columns: operand ring (per stage 32 columns hi + 32 lo: 32 rows of the stage), then the two fresh accumulators
constexpr uint32_t A_RING = 0, FRESH0 = 192, FRESH1 = 352, TMEM3 = 512;
__host__ __device__ constexpr uint32_t a_hi_col(uint32_t stage) { return A_RING + 64u * stage; }
__host__ __device__ constexpr uint32_t a_lo_col(uint32_t stage) { return A_RING + 64u * stage + 32u; }

enum { CAT_ODE = 0, CAT_OUT, CAT_READOUT, CAT_JUMP, CAT_JUMP0 };

template <int HW>
struct G3 {
  static constexpr int NB = HW / 32;                       // 32-feature blocks per operand
  static constexpr int Q_HI = 0, X_HI = NB * BLK, Q_LO = X_HI + BLK, X_LO = Q_LO + NB * BLK, STAGE = X_LO + BLK;
  static constexpr int NACC = HW + 16;                     // accumulator columns (in-features + aux block)
  static constexpr int CPT = NACC / 4;                     // columns per merge thread (4 column groups)
  static constexpr int NLOAD = HW / 8;                     // converter warps that stage one 8-column chunk of the B operand
  // staging stage: 32 rows of the D plane ([4 octets][HW][8]), of the A plane ([HW/8 chunks][32][8]) and of the aux rows
  static constexpr int RAW_P = 0, RAW_Q = SROWS * HW * 4, RAW_X = 2 * SROWS * HW * 4, RAW = RAW_X + SROWS * 32;
};

struct Ctl3 {
  // raw_full is per (converter group, staging stage): a barrier shared by the groups would let a group that is done
  // with stage sc - 2 wait for stage sc while stage sc - 3 (the other group's, same staging slot) has not landed yet --
  // one phase too early for the parity test, which then passes at once (bulk copies complete out of order under load)
  uint64_t raw_full[4][NRAW], raw_empty[NRAW], full[NSTAGE3], empty[NSTAGE3], fresh_done[2], merged[2];
  uint32_t tmem_base, pad;
  float dbo[4][4];                       // readout-bias partial sums of the converter groups (flush)
};

__device__ __forceinline__ int n_cats(int L) { return 3 * L + 3; }
__device__ __forceinline__ void cat_decode(int L, int c, int& kind, int& l) {
  if (c <= L) { kind = CAT_ODE; l = c; }
  else if (c <= 2 * L) { kind = CAT_OUT; l = c - (L + 1); }
  else if (c == 2 * L + 1) { kind = CAT_READOUT; l = L; }
  else if (c <= 3 * L + 1) { kind = CAT_JUMP; l = c - (2 * L + 1); }
  else { kind = CAT_JUMP0; l = 0; }
}
__device__ __forceinline__ int cat_instances(int kind, int kmax) {
  return kind == CAT_ODE ? kmax : (kind == CAT_OUT || kind == CAT_READOUT) ? 2 : 1;
}
__device__ __forceinline__ bool cat_has_q(int kind) { return !(kind == CAT_READOUT || kind == CAT_JUMP0); }

// walks (category, tile of this CTA, instance = plane pair) in the order every role uses
struct Cursor {
  int c, inst, n_inst, kmax, r, so;      // category, instance within the tile, instances of the tile, its kmax, round, first slot
};
__device__ __forceinline__ bool cursor_valid(const Cursor& cu, int L) { return cu.c < n_cats(L); }
__device__ __forceinline__ void cursor_seek(Cursor& cu, const SweepArgs& a, int wi, int n_w) {
  const int L = a.T.L;
  while (cu.c < n_cats(L)) {
    const int lo = a.tile_table[wi], cnt = a.tile_table[wi + 1] - lo;      // this worker's tiles: the schedule's table, as in the sweeps
    if (cu.r < cnt) {
      const int tile = a.tile_table[n_w + 1 + lo + cu.r];
      int kind, l;
      cat_decode(L, cu.c, kind, l);
      cu.kmax = a.tile_kmax[tile];
      cu.n_inst = cat_instances(kind, cu.kmax);
      if (cu.n_inst > 0) { cu.so = (int)a.tile_slot_off[tile]; cu.inst = 0; return; }
      ++cu.r;
    } else {
      ++cu.c;
      cu.r = 0;
    }
  }
}
__device__ __forceinline__ void cursor_init(Cursor& cu, const SweepArgs& a, int wi, int n_w) {
  cu.c = 0; cu.r = 0; cu.inst = 0; cu.n_inst = 0; cu.kmax = 0; cu.so = 0;
  cursor_seek(cu, a, wi, n_w);
}
__device__ __forceinline__ void cursor_next(Cursor& cu, const SweepArgs& a, int wi, int n_w) {
  if (++cu.inst < cu.n_inst) return;
  ++cu.r;
  cursor_seek(cu, a, wi, n_w);
}
// slots (relative to the tile's first) of the P plane (half D), the Q plane (half A) and the aux rows of the instance
__device__ __forceinline__ void cursor_planes(const Cursor& cu, int L, int& p_slot, int& p_plane, int& q_slot, int& q_plane, int& x_slot) {
  int kind, l;
  cat_decode(L, cu.c, kind, l);
  const int X = cu.kmax + 1 + cu.inst, X3 = cu.kmax + 3;
  q_slot = q_plane = 0;
  switch (kind) {
    case CAT_ODE: p_slot = q_slot = x_slot = cu.inst; p_plane = q_plane = l; break;
    case CAT_OUT:
      p_slot = x_slot = X; p_plane = l;
      if (l == 0) { q_slot = cu.inst == 0 ? 0 : cu.kmax; q_plane = 0; } else { q_slot = X; q_plane = l; }
      break;
    case CAT_READOUT: p_slot = x_slot = X; p_plane = L; break;      // (half D, plane L: the forward sweep's copy of the last hidden layer)
    case CAT_JUMP: p_slot = q_slot = x_slot = X3; p_plane = l; q_plane = l - 1; break;
    default: p_slot = x_slot = X3; p_plane = 0; break;
  }
}

// (ring and control block are re-derived inside each role: a pointer computed before setmaxnreg is live across the
//  register re-allocation, gets spilled, and its re-load inside the stage loop is an L2 round trip)
template <int HW>
struct Carve3 {
  uint8_t* ring;      // operand stages (MN tiles)
  uint8_t* raw;       // staging stages
  Ctl3* ctl;
  __device__ __forceinline__ explicit Carve3(uint8_t* base) {
    ring = (uint8_t*)(((uintptr_t)base + 1023) & ~(uintptr_t)1023);
    raw = ring + (size_t)NSTAGE3 * G3<HW>::STAGE;
    ctl = reinterpret_cast<Ctl3*>(raw + (size_t)NRAW * G3<HW>::RAW);
  }
};

// ------------------------------------------------------------------------------------------------
// producer (one thread): raw planes -> staging ring, three bulk copies per 32-row stage
// ------------------------------------------------------------------------------------------------
template <int HW>
__device__ __forceinline__ void wg_producer(const SweepArgs& a, uint8_t* base) {
  using G = G3<HW>;
  Carve3<HW> cv(base);
  Ctl3& ctl = *cv.ctl;
  const int L = a.T.L, S = a.T.S;
  const int s = blockIdx.x % S, wi = blockIdx.x / S, n_w = gridDim.x / S;
  const int64_t PL = Cfg<HW>::PL, slotf = (int64_t)(L + 1) * PL;
  const int64_t half = (int64_t)S * a.total_slots * slotf;
  const float* const baseA = a.ckpt + (int64_t)s * a.total_slots * slotf;
  const float* const baseD = baseA + half;
  const float* const baseX = a.ckpt + 2 * half + (int64_t)s * a.total_slots * (R * 8);
  Diag dg{g_status, g_notrap, 16u, false};
  Cursor cu;
  cursor_init(cu, a, wi, n_w);
  uint32_t rc = 0;
  while (cursor_valid(cu, L)) {
    int kind, l, ps, pp, qsl, qp, xs;
    cat_decode(L, cu.c, kind, l);
    cursor_planes(cu, L, ps, pp, qsl, qp, xs);
    const bool has_q = cat_has_q(kind);
    const float* P = baseD + ((int64_t)(cu.so + ps) * (L + 1) + pp) * PL;
    const float* Q = baseA + ((int64_t)(cu.so + qsl) * (L + 1) + qp) * PL;
    const float* X = baseX + (int64_t)(cu.so + xs) * (R * 8);
#pragma unroll 1
    for (int qs = 0; qs < R / SROWS; ++qs, ++rc) {
      const uint32_t rs = rc % NRAW, rround = rc / NRAW;
      wait_or_die(&ctl.raw_empty[rs], (rround & 1u) ^ 1u, dg, 9);
      uint8_t* dst = cv.raw + (size_t)rs * G::RAW;
      uint64_t* const landed = &ctl.raw_full[rc % NGRP][rs];       // the barrier of the group that converts this stage
      mbar_expect_tx(landed, (uint32_t)(SROWS * HW * 4 * (has_q ? 2 : 1) + SROWS * 32));
      bulk_g2s(dst + G::RAW_P, P + (int64_t)qs * (SROWS * HW), SROWS * HW * 4, landed);     // 4 row octets x HW x 8
      if (has_q) bulk_g2s(dst + G::RAW_Q, Q + (int64_t)qs * (SROWS * HW), SROWS * HW * 4, landed);   // one row group: HW/8 chunks x 32 x 8
      bulk_g2s(dst + G::RAW_X, X + (int64_t)qs * (SROWS * 8), SROWS * 32, landed);
    }
    cursor_next(cu, a, wi, n_w);
  }
}

// ------------------------------------------------------------------------------------------------
// MMA issuer (one warp)
// ------------------------------------------------------------------------------------------------
template <int HW>
__device__ __forceinline__ void wg_issuer(const SweepArgs& a, uint8_t* base) {
  using G = G3<HW>;
  Carve3<HW> cv(base);
  Ctl3& ctl = *cv.ctl;
  const int S = a.T.S, L = a.T.L;
  const int wi = blockIdx.x / S, n_w = gridDim.x / S;
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(&ctl.tmem_base);
  const uint32_t ring_s = umma::smem_u32(cv.ring);
  // M = 128 also at H = 64: same tensor time as M = 64 (the floor is max(M, 128) * N / 256 cycles), standard accumulator
  // layout (row = lane); TMEM lanes 64..127 of the operand ring then hold zeros and accumulator rows 64..127 are unused.
  // A from TMEM (lane = output feature, column = row of the stage), B MN-major from shared memory.
  constexpr uint32_t idesc_q = umma::idesc_tf32(128, G::NACC, 0, 1), idesc_x = umma::idesc_tf32(128, 16, 0, 1);
  Cursor cu;
  cursor_init(cu, a, wi, n_w);
  uint32_t sc = 0, ic = 0;
  Diag dg{g_status, g_notrap, 16u, false};
  while (cursor_valid(cu, L)) {
    int kind, l;
    cat_decode(L, cu.c, kind, l);
    const bool has_q = cat_has_q(kind);
    const uint32_t b = ic & 1u, use = ic >> 1;
    if (use > 0) wait_or_die(&ctl.merged[b], (use - 1) & 1u, dg, 5);     // the accumulator's previous content is merged
    const uint32_t acc = tmem + (b ? FRESH1 : FRESH0);
#pragma unroll 1
    for (int qs = 0; qs < R / SROWS; ++qs, ++sc) {
      const uint32_t stage = sc % NSTAGE3, sround = sc / NSTAGE3;
      wait_or_die(&ctl.full[stage], sround & 1u, dg, 6);
      umma::fence_after_sync();
      if (umma::elect_one()) {
        const uint32_t sb = ring_s + stage * G::STAGE;
        const uint32_t p_hi = tmem + a_hi_col(stage), p_lo = tmem + a_lo_col(stage);
        const uint64_t q_hi = umma::desc_mn(sb + (has_q ? G::Q_HI : G::X_HI), BLK), q_lo = umma::desc_mn(sb + (has_q ? G::Q_LO : G::X_LO), BLK);
        const uint32_t idesc = has_q ? idesc_q : idesc_x;
#pragma unroll
        for (int ks = 0; ks < SROWS / 8; ++ks) umma::mma_ts(acc, p_lo + 8 * ks, q_hi + 64 * ks, idesc, (qs > 0 || ks > 0) ? 1u : 0u);
#pragma unroll
        for (int ks = 0; ks < SROWS / 8; ++ks) umma::mma_ts(acc, p_hi + 8 * ks, q_lo + 64 * ks, idesc, 1u);
#pragma unroll
        for (int ks = 0; ks < SROWS / 8; ++ks) umma::mma_ts(acc, p_hi + 8 * ks, q_hi + 64 * ks, idesc, 1u);
        umma::commit(&ctl.empty[stage]);
        if (qs == R / SROWS - 1) umma::commit(&ctl.fresh_done[b]);
      }
      __syncwarp();
    }
    cursor_next(cu, a, wi, n_w);
    ++ic;
  }
}

// ------------------------------------------------------------------------------------------------
// converters (16 warps): staged raw stage -> tf32 hi / lo operand stage; merge of finished accumulators; flush
// ------------------------------------------------------------------------------------------------
template <int HW>
__device__ __forceinline__ void wg_worker(const SweepArgs& a, uint8_t* base) {
  using G = G3<HW>;
  const ParamTable& T = a.T;
  const int L = T.L, S = T.S, dx = T.d_x, O = T.O;
  const int wi = blockIdx.x / S, n_w = gridDim.x / S;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = warp & 3, cg = warp >> 2;            // TMEM lane quadrant; accumulator column group (merge / flush)
  const int grp = warp / GWARPS, wl = warp % GWARPS; // converter group, warp within the group
  constexpr int OPW = 4 * NGRP / 4;                  // row octets of a stage per warp: 4 octets over GWARPS / 4 warps per quadrant
  const int oct0 = (wl >> 2) * OPW;                  // first of this warp's octets
  // shared-window (32-bit) addresses of the rings and the barriers (same carve as Carve3)
  const uint32_t ring_s = (smem_u32_once(base) + 1023u) & ~1023u;
  const uint32_t raw_s = ring_s + NSTAGE3 * G::STAGE, ctl_s = raw_s + NRAW * G::RAW;
  const uint32_t b_raw_full = ctl_s + offsetof(Ctl3, raw_full), b_raw_empty = ctl_s + offsetof(Ctl3, raw_empty);
  const uint32_t b_full = ctl_s + offsetof(Ctl3, full), b_empty = ctl_s + offsetof(Ctl3, empty);
  const uint32_t b_fresh = ctl_s + offsetof(Ctl3, fresh_done), b_merged = ctl_s + offsetof(Ctl3, merged);
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(ctl_s + (uint32_t)offsetof(Ctl3, tmem_base)));
  const uint32_t quad_t = tmem + ((uint32_t)(q * 32) << 16);                  // this warp's TMEM lane quadrant, column 0
  float* const part = a.partials + (int64_t)blockIdx.x * T.stack_floats;
  constexpr int CPW = (G::NLOAD + GWARPS - 1) / GWARPS;   // 8-column chunks of the B operand per warp (lane = row)
  const bool q_loader = wl < G::NLOAD;
  const int irow = q * 32 + lane;                   // this thread's output feature = its TMEM lane
  const bool has_row = irow < HW;                   // (at H = 64 lanes 64..127 hold nothing)
  const int sc_kind = a.desc.input_scaling;
  const float comp = a.comp_wgrad;
  Diag dg{g_status, g_notrap, 16u, false};
  // this thread's pieces of a staging stage
  const uint32_t rawP = raw_s + G::RAW_P + (uint32_t)(oct0 * HW + irow) * 32u;      // [octet][feature][8 rows]
  const uint32_t rawQ = raw_s + G::RAW_Q + (uint32_t)(wl * 32 + lane) * 32u;        // [chunk][row][8]: chunks wl, wl + GWARPS, ..
  const uint32_t rawX = raw_s + G::RAW_X + (uint32_t)lane * 32u;

  // running sums of the current category: accumulator row irow, columns [cg * CPT, +CPT)
  float run[G::CPT];
#pragma unroll
  for (int t = 0; t < G::CPT; ++t) run[t] = 0.0f;
  auto ld4 = [&](uint32_t addr, float (&v)[4]) {
    uint32_t u0, u1, u2, u3;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(u0), "=r"(u1), "=r"(u2), "=r"(u3) : "r"(addr));
    v[0] = __uint_as_float(u0); v[1] = __uint_as_float(u1); v[2] = __uint_as_float(u2); v[3] = __uint_as_float(u3);
  };
  // running += fresh[b]  (IEEE adds; comp: accumulator truncation compensation), then release the fresh accumulator
  auto merge = [&](uint32_t ic) {
    const uint32_t b = ic & 1u;
    wait_or_die_a(b_fresh + 8u * b, (ic >> 1) & 1u, dg, 7);
    umma::fence_after_sync();
    const uint32_t fr = quad_t + (b ? FRESH1 : FRESH0) + (uint32_t)(cg * G::CPT);
    constexpr int MB = 12;               // columns per batch of TMEM loads (one wait per batch: the loads' latencies overlap)
#pragma unroll
    for (int t0 = 0; t0 < G::CPT; t0 += MB) {
      float f[MB];
#pragma unroll
      for (int t = 0; t < MB; t += 4) {
        if (t0 + t < G::CPT) {
          float g[4];
          ld4(fr + t0 + t, g);
#pragma unroll
          for (int i = 0; i < 4; ++i) f[t + i] = g[i];
        }
      }
      umma::wait_ld();
#pragma unroll
      for (int t = 0; t < MB; ++t)
        if (t0 + t < G::CPT) run[t0 + t] = fmaf(f[t], comp, run[t0 + t]);
    }
    umma::fence_before_sync();
    __syncwarp();
    if (lane == 0) mbar_arrive_a(b_merged + 8u * b);
  };
  float dbo[MAX_O];                      // readout bias gradient: sum over rows of dY (aux converter warp, lane = row % 32)
#pragma unroll
  for (int o = 0; o < MAX_O; ++o) dbo[o] = 0.0f;
  // running sums of category (kind, l) -> this CTA's partial buffer (PyTorch layout), then clear them
  auto flush = [&](int kind, int l) {
    if (has_row) {
      const int net = kind == CAT_ODE ? NET_ODE : (kind == CAT_OUT || kind == CAT_READOUT) ? NET_OUT : NET_JUMP;
      const int ld = T.n_vec[net][l] + T.n_ext[net][l];
#pragma unroll
      for (int t = 0; t < G::CPT; ++t) {
        const int n = cg * G::CPT + t;
        const float v = run[t];
        if (kind == CAT_READOUT) {
          if (n >= 1 && n <= O) part[T.w_off[NET_OUT][L] + (n - 1) * HW + irow] = v;
        } else if (kind == CAT_JUMP0) {
          if (n == 0) part[T.b_off[NET_JUMP][0] + irow] = v;
          else if (n <= dx) part[T.w_off[NET_JUMP][0] + irow * dx + (n - 1)] = v;
        } else if (n < HW) {
          part[T.w_off[net][l] + irow * ld + n] = v;
        } else if (n == HW) {
          part[T.b_off[net][l] + irow] = v;
        } else if (kind == CAT_ODE && l == 0 && n - HW - 1 < dx + 2) {
          part[T.w_off[net][l] + irow * ld + HW + (n - HW - 1)] = v;
        }
      }
    }
    if (kind == CAT_READOUT && wl == 0) {          // the aux-converting warp of each group holds a share of the row sums
      static_assert(MAX_O == 4, "Ctl3::dbo");
#pragma unroll
      for (int o = 0; o < MAX_O; ++o) {
        float v = dbo[o];
        for (int sft = 16; sft > 0; sft >>= 1) v += __shfl_xor_sync(0xffffffffu, v, sft);
        if (lane == 0) asm volatile("st.shared.f32 [%0], %1;" :: "r"(ctl_s + (uint32_t)offsetof(Ctl3, dbo) + (uint32_t)(grp * 4 + o) * 4u), "f"(v) : "memory");
      }
#ifndef NJODE_K3_NO_DBO_BAR
      if (NGRP > 1) umma::named_bar_sync(2, 32 * NGRP);      // (the category occurs once per kernel)
      else __syncwarp();
#endif
      if (grp == 0 && lane < O) {
        float v = 0.0f;
        for (int g2 = 0; g2 < NGRP; ++g2) {
          float t;
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(t) : "r"(ctl_s + (uint32_t)offsetof(Ctl3, dbo) + (uint32_t)(g2 * 4 + lane) * 4u) : "memory");
          v += t;
        }
        part[T.b_off[NET_OUT][L] + lane] = v;
      }
    }
#pragma unroll
    for (int t = 0; t < G::CPT; ++t) run[t] = 0.0f;
  };

  // everything the MMAs may read must be finite: clear the operand tiles (aux columns 8..15 stay zero for good), the
  // TMEM operand ring (lanes of features >= H are never written again) and both accumulators
  for (uint32_t i = threadIdx.x; i < (uint32_t)(NSTAGE3 * G::STAGE / 16); i += NT_W) umma::st_shared_v4(ring_s + 16u * i, 0u, 0u, 0u, 0u);
  {
    uint32_t z[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
    for (uint32_t c = (uint32_t)cg * 8; c < TMEM3; c += 32) umma::tmem_st8_raw(quad_t + c, z);      // the 4 warps of a quadrant interleave
    umma::wait_st();
  }
  umma::fence_async_smem();
  umma::fence_before_sync();
  umma::named_bar_sync(1, NT_W);
  umma::fence_after_sync();

  Cursor cur;
  cursor_init(cur, a, wi, n_w);
  uint32_t sc = 0, ic = 0;
  bool pending = false;                  // instance ic - 1 not merged yet
  PH_DECL;
  while (cursor_valid(cur, L)) {
    int kind, l;
    cat_decode(L, cur.c, kind, l);
    const int c_now = cur.c;
    const bool has_q = cat_has_q(kind);
    const bool scale_q = kind == CAT_ODE && l == 0 && sc_kind != NJODE_SCALE_IDENTITY;
#pragma unroll 1
    for (int qs = 0; qs < R / SROWS; ++qs, ++sc) {
      if (NGRP > 1 && (int)(sc % NGRP) != grp) continue;           // the other group's stage
      const uint32_t stage = sc % NSTAGE3, sround = sc / NSTAGE3;
      const uint32_t rs = sc % NRAW;
      PH(0);
      // the raw stage has landed (bulk copies complete); (group, slot) pairs repeat every NRAW * NGRP stages
      wait_or_die_a(b_raw_full + 8u * (grp * NRAW + rs), (sc / (NRAW * NGRP)) & 1u, dg, 10);
      PH(1);
      const uint32_t roff = rs * (uint32_t)G::RAW;
      wait_or_die_a(b_empty + 8u * stage, (sround & 1u) ^ 1u, dg, 8);   // the operand stage is free (its MMAs are done)
      PH(2);
      const uint32_t sb = ring_s + stage * G::STAGE;
      if (has_row) {                                               // A operand: 8 rows per octet of this feature -> TMEM
#pragma unroll
        for (int i = 0; i < OPW; ++i) {
          float p[8];
          uint32_t hi[8], lo[8];
          ld8s_a(rawP + roff + (uint32_t)(i * HW) * 32u, p);
          umma::split8(p, hi, lo);
          umma::tmem_st8_raw(quad_t + a_hi_col(stage) + 8 * (oct0 + i), hi);
          umma::tmem_st8_raw(quad_t + a_lo_col(stage) + 8 * (oct0 + i), lo);
        }
      }
      if (q_loader && has_q) {                                     // B operand: 8-column chunks of 32 rows -> MN-major tiles
#pragma unroll
        for (int i = 0; i < CPW; ++i) {
          const int c = wl + i * GWARPS;
          if (c < G::NLOAD) {
            float qv[8];
            uint32_t hi[8], lo[8];
            ld8s_a(rawQ + roff + (uint32_t)(i * GWARPS) * 1024u, qv);
            if (scale_q) {
#pragma unroll
              for (int e = 0; e < 8; ++e) qv[e] = scale_fwd_rt(sc_kind, qv[e]);
            }
            umma::split8(qv, hi, lo);
            const uint32_t boff = (uint32_t)(c >> 2) * BLK;
            umma::chunk_to_mn_tile(sb + G::Q_HI + boff, lane, c & 3, hi);
            umma::chunk_to_mn_tile(sb + G::Q_LO + boff, lane, c & 3, lo);
          }
        }
      }
      if (wl == 0) {
        float x[8];
        uint32_t hi[8], lo[8];
        ld8s_a(rawX + roff, x);
        umma::split8(x, hi, lo);
        umma::chunk_to_mn_tile(sb + G::X_HI, lane, 0, hi);
        umma::chunk_to_mn_tile(sb + G::X_LO, lane, 0, lo);
        if (kind == CAT_READOUT) {
#pragma unroll
          for (int o = 0; o < MAX_O; ++o) dbo[o] += x[1 + o];
        }
      }
      PH(3);                                                       // loads + split + stores
      // One proxy fence covers both hand-overs: it orders the tile stores before the MMAs' reads AND the staging-stage
      // reads above before the producer's next bulk copy into the same bytes (releasing the staging stage earlier, on a
      // plain arrive, lost one stage in a few thousand to the refill: tools/k3_stress.py).
      umma::wait_st();
      umma::fence_before_sync();
      umma::fence_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive_a(b_full + 8u * stage);                        // operand stage ready for the issuer
        mbar_arrive_a(b_raw_empty + 8u * rs);                      // staging stage consumed: the producer may refill it
      }
      PH(4);                                                       // fences + hand-over
    }
    cursor_next(cur, a, wi, n_w);
    if (pending) merge(ic - 1);
    PH(5);                                                         // merge (incl. waiting for the accumulator)
    pending = true;
    ++ic;
    if (!cursor_valid(cur, L) || cur.c != c_now) {    // category finished: fold the last accumulator in and flush
      merge(ic - 1);
      pending = false;
      flush(kind, l);
      PH(6);                                                       // category end: last merge + flush
    }
  }
  PH_STORE(g_phase3);
}

template <int HW>
__global__ void __launch_bounds__(NT3, 1) k_wide_wgrad(SweepArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5;
  {
    Ctl3& ctl = *Carve3<HW>(smem_raw).ctl;
    if (threadIdx.x == 0) {
      for (int i = 0; i < NRAW; ++i) { for (int g = 0; g < 4; ++g) umma::mbar_init(&ctl.raw_full[g][i], 1); umma::mbar_init(&ctl.raw_empty[i], GWARPS); }
      for (int i = 0; i < NSTAGE3; ++i) { umma::mbar_init(&ctl.full[i], GWARPS); umma::mbar_init(&ctl.empty[i], 1); }
      for (int i = 0; i < 2; ++i) { umma::mbar_init(&ctl.fresh_done[i], 1); umma::mbar_init(&ctl.merged[i], NWARP_W); }
      umma::fence_mbar_init();
    }
    if (warp == 0) umma::tmem_alloc(&ctl.tmem_base, TMEM3);
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
  }
  if (warp < NWARP_W) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
    wg_worker<HW>(a, smem_raw);
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
    if (warp == NWARP_W) wg_issuer<HW>(a, smem_raw);
    else if (threadIdx.x == NT_W + 32) wg_producer<HW>(a, smem_raw);
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_free(*reinterpret_cast<volatile uint32_t*>(&Carve3<HW>(smem_raw).ctl->tmem_base), TMEM3);
}

template <int HW>
int launch_wgrad(const SweepArgs& a, cudaStream_t st) {
  const size_t smem = 1024 + (size_t)NSTAGE3 * G3<HW>::STAGE + (size_t)NRAW * G3<HW>::RAW + sizeof(Ctl3) + 16;
  if (njode_no_trap_env()) { const unsigned one = 1; NJODE_CUDA_OK(cudaMemcpyToSymbolAsync(g_notrap, &one, sizeof(one), 0, cudaMemcpyHostToDevice, st)); }
  NJODE_CUDA_OK(cudaFuncSetAttribute(k_wide_wgrad<HW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  njode_timing_begin(3, st);
  k_wide_wgrad<HW><<<a.n_workers, NT3, smem, st>>>(a);
  njode_timing_end(3, st);
  NJODE_LAUNCH_OK("k_wide_wgrad");
  if (njode_debug_sync_env()) {
    const cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) NJODE_FAIL(NJODE_ECUDA, "NJODE_DEBUG_SYNC: k_wide_wgrad failed: %s", cudaGetErrorString(e));
    fprintf(stderr, "NJODE_DEBUG_SYNC: k_wide_wgrad ok\n");
  }
  return NJODE_OK;
}

}  // namespace

int njode_wide_wgrad(const SweepArgs& a, cudaStream_t st) {
  if (a.n_tiles == 0) return NJODE_OK;
  return a.desc.hidden == 128 ? launch_wgrad<128>(a, st) : launch_wgrad<64>(a, st);
}

int njode_wgrad_phase_fetch(unsigned long long* out_host, int n_ctas) {
  NJODE_CUDA_OK(cudaMemcpyFromSymbol(out_host, g_phase3, (size_t)n_ctas * 8 * sizeof(unsigned long long)));
  return NJODE_OK;
}

int njode_wide_wgrad_status(unsigned* out_host) {
  NJODE_CUDA_OK(cudaMemcpyFromSymbol(out_host, g_status, 2 * sizeof(unsigned)));
  return NJODE_OK;
}
