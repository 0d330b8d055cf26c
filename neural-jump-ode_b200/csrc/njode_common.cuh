// njode_common.cuh -- shared host/device definitions for libnjode_b200 (sm_100a).
// Parameter layout, activation maths and error plumbing.  See include/njode.h for the ABI.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/njode.h"

#define NJODE_LMAX 7          // n_hidden_layers <= 7  (L+1 <= 8 Linear layers per net)
#define NJODE_WARP 32
#define NJODE_FULL 0xffffffffu

enum { NET_JUMP = 0, NET_ODE = 1, NET_OUT = 2 };

// Offsets (in floats, relative to the start of one stack's block) and shapes of every Linear.
// Layer (net,l): weight is (n_out x (n_vec+n_ext)) row-major; the first n_vec input columns are a
// hidden vector, the last n_ext columns are per-row scalars (x for jump layer 0; x,t,dt for ode layer 0).
struct ParamTable {
  int32_t w_off[3][NJODE_LMAX + 1];
  int32_t b_off[3][NJODE_LMAX + 1];
  int32_t n_vec[3][NJODE_LMAX + 1];
  int32_t n_ext[3][NJODE_LMAX + 1];
  int32_t n_out[3][NJODE_LMAX + 1];
  int32_t act[3][NJODE_LMAX + 1];   // activation applied after this layer?
  int32_t stack_floats;             // parameters per stack
  int32_t L, H, O, S, M, d_x, d_y;
};

static inline int njode_desc_ok(const NjodeDesc* d, const char** why) {
  if (!d) { *why = "null descriptor"; return 0; }
  if (d->d_x < 1 || d->d_y < 1 || d->hidden < 1 || d->num_moments < 1) { *why = "dimensions must be >= 1"; return 0; }
  if (d->n_hidden_layers < 1 || d->n_hidden_layers > NJODE_LMAX) { *why = "n_hidden_layers must be in 1..7"; return 0; }
  if (d->activation < 0 || d->activation > NJODE_ACT_SELU) { *why = "unknown activation code"; return 0; }
  if (d->input_scaling < 0 || d->input_scaling > NJODE_SCALE_SIGMOID) { *why = "unknown input_scaling code"; return 0; }
  if (d->has_dt && !(d->dt > 0.0f)) { *why = "dt_ode_step must be > 0"; return 0; }
  return 1;
}

static inline ParamTable njode_make_table(const NjodeDesc* d) {
  ParamTable T;
  const int L = d->n_hidden_layers, H = d->hidden;
  T.L = L; T.H = H; T.M = d->num_moments; T.d_x = d->d_x; T.d_y = d->d_y;
  T.S = d->shared_network ? 1 : d->num_moments;
  T.O = d->shared_network ? d->d_y * d->num_moments : d->d_y;
  int off = 0;
  for (int net = 0; net < 3; ++net) {
    for (int l = 0; l <= NJODE_LMAX; ++l) {
      T.w_off[net][l] = T.b_off[net][l] = T.n_vec[net][l] = T.n_ext[net][l] = T.n_out[net][l] = T.act[net][l] = 0;
    }
    for (int l = 0; l <= L; ++l) {
      int n_vec = H, n_ext = 0, n_out = H, act = 1;
      if (net == NET_JUMP) { if (l == 0) { n_vec = 0; n_ext = d->d_x; } act = 1; }
      if (net == NET_ODE)  { if (l == 0) { n_ext = d->d_x + 2; } act = (l < L); }
      if (net == NET_OUT)  { if (l == L) n_out = T.O; act = (l < L); }
      T.n_vec[net][l] = n_vec; T.n_ext[net][l] = n_ext; T.n_out[net][l] = n_out; T.act[net][l] = act;
      T.w_off[net][l] = off; off += n_out * (n_vec + n_ext);
      T.b_off[net][l] = off; off += n_out;
    }
  }
  T.stack_floats = off;
  return T;
}

// ------------------------------------------------------------------------------------------------
// activations (jump_ode.py:6-13) -- forward value and derivative expressed through the OUTPUT y,
// so the reverse sweep only needs the recomputed / stored post-activation values.
// ------------------------------------------------------------------------------------------------
#define NJODE_SELU_LAMBDA 1.0507009873554804934193349852946f
#define NJODE_SELU_ALPHA  1.6732632423543772848170429916717f

template <int ACT>
__device__ __forceinline__ float act_fwd(float a) {
  if (ACT == NJODE_ACT_RELU) {
    // NaN-propagating maximum, as torch.relu (clamp_min): the reference turns a NaN observation into NaN predictions,
    // loss and gradients (tests/golden/nan_observation_h32); fmaxf would silently drop the NaN
    float r;
    asm("max.NaN.f32 %0, %1, 0f00000000;" : "=f"(r) : "f"(a));
    return r;
  }
  if (ACT == NJODE_ACT_TANH) return tanhf(a);
  if (ACT == NJODE_ACT_SIGMOID) return 1.0f / (1.0f + expf(-a));
  if (ACT == NJODE_ACT_ELU) return a > 0.0f ? a : expm1f(a);
  if (ACT == NJODE_ACT_LEAKY_RELU) return a > 0.0f ? a : 0.01f * a;
  /* SELU */ return NJODE_SELU_LAMBDA * (a > 0.0f ? a : NJODE_SELU_ALPHA * expm1f(a));
}
template <int ACT>
__device__ __forceinline__ float act_grad_from_out(float y) {
  if (ACT == NJODE_ACT_RELU) return y > 0.0f ? 1.0f : 0.0f;
  if (ACT == NJODE_ACT_TANH) return 1.0f - y * y;
  if (ACT == NJODE_ACT_SIGMOID) return y * (1.0f - y);
  if (ACT == NJODE_ACT_ELU) return y > 0.0f ? 1.0f : y + 1.0f;
  if (ACT == NJODE_ACT_LEAKY_RELU) return y > 0.0f ? 1.0f : 0.01f;
  /* SELU */ return y > 0.0f ? NJODE_SELU_LAMBDA : y + NJODE_SELU_LAMBDA * NJODE_SELU_ALPHA;
}
__device__ __forceinline__ float act_fwd_rt(int act, float a) {
  switch (act) {
    case NJODE_ACT_RELU: return act_fwd<NJODE_ACT_RELU>(a);
    case NJODE_ACT_TANH: return act_fwd<NJODE_ACT_TANH>(a);
    case NJODE_ACT_SIGMOID: return act_fwd<NJODE_ACT_SIGMOID>(a);
    case NJODE_ACT_ELU: return act_fwd<NJODE_ACT_ELU>(a);
    case NJODE_ACT_LEAKY_RELU: return act_fwd<NJODE_ACT_LEAKY_RELU>(a);
    default: return act_fwd<NJODE_ACT_SELU>(a);
  }
}
__device__ __forceinline__ float act_grad_rt(int act, float y) {
  switch (act) {
    case NJODE_ACT_RELU: return act_grad_from_out<NJODE_ACT_RELU>(y);
    case NJODE_ACT_TANH: return act_grad_from_out<NJODE_ACT_TANH>(y);
    case NJODE_ACT_SIGMOID: return act_grad_from_out<NJODE_ACT_SIGMOID>(y);
    case NJODE_ACT_ELU: return act_grad_from_out<NJODE_ACT_ELU>(y);
    case NJODE_ACT_LEAKY_RELU: return act_grad_from_out<NJODE_ACT_LEAKY_RELU>(y);
    default: return act_grad_from_out<NJODE_ACT_SELU>(y);
  }
}
// input scaling s(.) of the ODE net (jump_ode.py:43-50, :57-58) and s' through the output
__device__ __forceinline__ float scale_fwd_rt(int sc, float v) {
  if (sc == NJODE_SCALE_TANH) return tanhf(v);
  if (sc == NJODE_SCALE_SIGMOID) return 1.0f / (1.0f + expf(-v));
  return v;
}
__device__ __forceinline__ float scale_grad_rt(int sc, float s) {
  if (sc == NJODE_SCALE_TANH) return 1.0f - s * s;
  if (sc == NJODE_SCALE_SIGMOID) return s * (1.0f - s);
  return 1.0f;
}

// ------------------------------------------------------------------------------------------------
// error plumbing (thread-local text, returned by njode_last_error)
// ------------------------------------------------------------------------------------------------
void njode_set_error(const char* fmt, ...);
#define NJODE_FAIL(code, ...) do { njode_set_error(__VA_ARGS__); return (code); } while (0)
#define NJODE_CUDA_OK(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { \
    njode_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
    return NJODE_ECUDA; } } while (0)
#define NJODE_LAUNCH_OK(what) do { njode_count_launch(1); cudaError_t e__ = cudaGetLastError(); if (e__ != cudaSuccess) { \
    njode_set_error("launch of %s failed: %s (%s:%d)", what, cudaGetErrorString(e__), __FILE__, __LINE__); \
    return NJODE_ECUDA; } } while (0)

// every kernel launch of the library bumps this counter (njode_kernel_launches: bench.py's gpu_launches claim)
void njode_count_launch(int n);

// one-shot CUDA events around the main sweep kernel (njode_set_kernel_timing)
void njode_timing_begin(int which, cudaStream_t st);
void njode_timing_end(int which, cudaStream_t st);

static inline size_t njode_align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ------------------------------------------------------------------------------------------------
// kernel flavours (implemented in njode_generic.cu / njode_tiled.cu)
// ------------------------------------------------------------------------------------------------
// Units packed into each tile (<= tile_rows; the remaining rows of a tile are padding).  The tcgen05 kernels are
// latency chains: with few tiles per SM their time is (steps of the longest tile) x (latency of one step), and the
// weight-gradient MMAs (contraction over the tile's rows) are the long pole of a step -- half- or quarter-filled
// tiles shorten that pole and spread the work over more SMs.  Every entry point derives this from (desc, N).
int32_t njode_tile_units(const NjodeDesc* d, int64_t N);
// The tiling of N units (sorted by step count, longest first): the first n_small tiles hold units_small units each,
// the others `units` (= njode_tile_units).  Small leading tiles keep the LONGEST units -- the critical path of a
// sweep with few tiles per SM -- in tiles whose reverse step is cheap (4 instead of 16 K-slices per weight-gradient
// batch, a quarter of the tile stores), while the bulk of the work stays in full tiles.
struct TilePlan {
  int32_t units, units_small;
  int64_t n_small, n_tiles;
  __host__ __device__ int32_t units_of(int64_t tile) const { return tile < n_small ? units_small : units; }
  __host__ __device__ int64_t first_unit(int64_t tile) const {
    return tile < n_small ? tile * units_small : n_small * units_small + (tile - n_small) * units;
  }
};
TilePlan njode_tile_plan(const NjodeDesc* d, int64_t N);

// checkpoint slots a tile owns beyond its kmax + 1 step slots (the wide flavour keeps the jump / readout
// activations of a tile in three extra slots, see njode_wide.cuh)
int32_t njode_slot_extra(const NjodeDesc* d);
// wide flavour: workers per stack the tile table is built for (0 = this flavour has no table) and the table's size
int32_t njode_table_workers(const NjodeDesc* d, int64_t n_tiles);
static inline size_t njode_table_ints(int32_t table_workers, int64_t n_tiles) {       // offsets, lists, scratch
  return table_workers > 0 ? (size_t)(table_workers + 1 + 2 * n_tiles) : 0;
}

struct SweepArgs {
  NjodeDesc desc;
  ParamTable T;
  const float* params;       // flat, PyTorch layout
  const float* params_t;     // flat, every weight transposed to (in x out) ("kernel layout")
  const float* times;
  const float* values;
  const int32_t* kenc;
  const int32_t* perm;
  const int32_t* tile_kmax;
  const int64_t* tile_slot_off;
  const float* knots;
  int64_t N, n_tiles, total_slots;
  int32_t tile_rows;
  int32_t tile_units;        // units per tile (rows >= tile_units of every tile are padding) ...
  int32_t tile_units_small;  // ... except in the first n_small_tiles tiles (njode_tile_plan)
  int64_t n_small_tiles;
  // forward outputs
  float* preds;
  float* preds_before;
  float* ckpt;               // may be NULL in forward (inference)
  // backward
  const float* grad_preds;
  const float* grad_preds_before;
  float* partials;           // [n_workers][stack_floats] per-CTA weight-gradient partial sums
  int32_t n_workers;         // CTAs (multiple of S); worker w serves stack w % S
  // wide flavour: the tensor core adds into its FP32 accumulator with truncation, so a sum built by n MMAs comes out
  // ~n x 2.1e-8 too small in magnitude (round 1 measured the drift, round 2 modelled it: DESIGN.md); the epilogues
  // scale what they read back by (1 + that) -- the factors for a chain GEMM and for a weight-gradient plane pair
  float comp_chain, comp_wgrad;
  // wide flavour: the tiles of worker w (of table_workers per stack), in the order it runs them:
  // tile_table[n_w + 1 + i] for i in [tile_table[w], tile_table[w + 1]).  Built with the schedule (njode_schedule.cu:
  // greedy longest-processing-time assignment).
  const int32_t* tile_table;
  int32_t table_workers;
};

#define NJODE_GENERIC_TILE_ROWS 32
int  njode_generic_supported(const NjodeDesc* d, const char** why);
int  njode_generic_workers(const NjodeDesc* d, int64_t n_tiles);
int  njode_generic_forward(const SweepArgs& a, cudaStream_t st);
int  njode_generic_backward(const SweepArgs& a, cudaStream_t st);
int  njode_generic_dense(const NjodeDesc* d, const float* params, const float* params_t, const float* times, const float* values,
                         const int64_t* off, int64_t B, int64_t N, const float* grid, int64_t G, float* dense, cudaStream_t st);

// row-tiled FP32 kernels (njode_rowtile.cu): hidden_dim in {32,64,96,128}, <= 3 hidden layers; same tile rows and
// checkpoint layout as the generic flavour
int  njode_rowtile_supported(const NjodeDesc* d);
int  njode_rowtile_workers(const NjodeDesc* d, int64_t n_tiles);
int  njode_rowtile_forward(const SweepArgs& a, cudaStream_t st);
int  njode_rowtile_backward(const SweepArgs& a, cudaStream_t st);

#define NJODE_TILED_TILE_ROWS 128
int  njode_tiled_supported(const NjodeDesc* d);
int  njode_tiled_workers(const NjodeDesc* d, int64_t n_tiles);
int  njode_tiled_forward(const SweepArgs& a, cudaStream_t st);
int  njode_tiled_backward(const SweepArgs& a, cudaStream_t st);
int  njode_tiled_status(unsigned* out_host);
