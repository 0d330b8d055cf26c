// njode_abi.cu -- extern "C" entry points of libnjode_b200.so (see include/njode.h).
// Argument validation, flavour selection, weight re-layout and the deterministic reduction of the
// per-CTA weight-gradient partial sums.  No torch headers, no persistent allocations.
#include <cstdlib>
#include <stdarg.h>
#include <string.h>

#include "njode_common.cuh"
#include "njode_wide.cuh"

int njode_wide_sweep_status(unsigned* out_host);
int njode_wide_wgrad_status(unsigned* out_host);

static thread_local char g_err[512] = "";

void njode_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* njode_last_error(void) { return g_err; }

// ---- measurement hooks -------------------------------------------------------------------------
#include <atomic>
static std::atomic<void*> g_ev[4][2];
extern "C" int njode_set_kernel_timing(int32_t which, void* ev_start, void* ev_stop) {
  if (which < 1 || which > 3) NJODE_FAIL(NJODE_EINVAL, "njode_set_kernel_timing: which must be 1 (forward), 2 (backward) or 3 (weight-gradient GEMM)");
  g_ev[which][0].store(ev_start);
  g_ev[which][1].store(ev_stop);
  return NJODE_OK;
}
void njode_timing_begin(int which, cudaStream_t st) {
  void* e = g_ev[which][0].exchange(nullptr);
  if (e) cudaEventRecord((cudaEvent_t)e, st);
}
void njode_timing_end(int which, cudaStream_t st) {
  void* e = g_ev[which][1].exchange(nullptr);
  if (e) cudaEventRecord((cudaEvent_t)e, st);
}

__global__ void __launch_bounds__(256) k_ffma_peak(float* out, int iters, float a, float b) {
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 0.001f + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], a, b);
  }
  float s = 0.0f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

extern "C" int njode_ffma_peak(float* tflops_host) {
  if (!tflops_host) NJODE_FAIL(NJODE_EINVAL, "njode_ffma_peak: null output");
  int dev = 0, sms = 0;
  NJODE_CUDA_OK(cudaGetDevice(&dev));
  NJODE_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int blocks = sms * 8, iters = 8000;
  float* out = nullptr;
  NJODE_CUDA_OK(cudaMalloc(&out, (size_t)blocks * 256 * sizeof(float)));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 6; ++rep) {
    cudaEventRecord(e0);
    k_ffma_peak<<<blocks, 256>>>(out, iters, 1.0001f, 0.5f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  NJODE_LAUNCH_OK("k_ffma_peak");
  const double flops = 2.0 * blocks * 256.0 * iters * 8 * 16;
  *tflops_host = (float)(flops / best * 1e-9);
  return NJODE_OK;
}
extern "C" int32_t njode_abi_version(void) { return NJODE_ABI_VERSION; }

static std::atomic<long long> g_launches{0};
void njode_count_launch(int n) { g_launches.fetch_add(n); }
extern "C" int64_t njode_kernel_launches(int32_t reset) {
  return reset ? g_launches.exchange(0) : g_launches.load();
}

extern "C" int njode_device_status(uint32_t* status_host) {
  if (!status_host) NJODE_FAIL(NJODE_EINVAL, "njode_device_status: null output");
  unsigned v = 0, w1[2] = {0, 0}, w2[2] = {0, 0};
  int rc = njode_tiled_status(&v);
  if (!rc) rc = njode_wide_sweep_status(w1);
  if (!rc) rc = njode_wide_wgrad_status(w2);
  *status_host = v | w1[0] | w2[0];
  return rc;
}

// `make phase` build only: per-CTA phase cycles, uint64[n_ctas][8]; which = 1: last sweep launch, worker thread 0 (buckets
// 0 epilogue, 1 accumulator wait), 2: its MMA issuer (0 operand wait, 1 weight wait, 2 issue), 3: weight-gradient loader role (0 bookkeeping, 1 stage wait, 2 split + stores, 3 fence + hand-over,
// 4 load issue, 5 merge, 6 flush).  All zeros in the product build.
int njode_sweep_phase_fetch(unsigned long long* out_host, int n_ctas, int issuer);
int njode_wgrad_phase_fetch(unsigned long long* out_host, int n_ctas);
extern "C" int njode_debug_phase(int32_t which, unsigned long long* out_host, int32_t n_ctas) {
  if (!out_host || n_ctas < 1 || n_ctas > 512) NJODE_FAIL(NJODE_EINVAL, "njode_debug_phase: need 1..512 CTAs");
  NJODE_CUDA_OK(cudaDeviceSynchronize());
  return which == 3 ? njode_wgrad_phase_fetch(out_host, n_ctas) : njode_sweep_phase_fetch(out_host, n_ctas, which == 2);
}

// bring-up detail of the wide kernels: {sweep status, first sweep site that gave up, weight-gradient status, its first site}
// (site = code | warp << 8 | block << 16; codes in njode_wide.cu / njode_wgrad.cu)
extern "C" int njode_device_status_detail(uint32_t* words_host) {
  if (!words_host) NJODE_FAIL(NJODE_EINVAL, "njode_device_status_detail: null output");
  int rc = njode_wide_sweep_status(words_host);
  if (!rc) rc = njode_wide_wgrad_status(words_host + 2);
  return rc;
}

extern "C" int64_t njode_params_per_stack(const NjodeDesc* d) {
  const char* why = nullptr;
  if (!njode_desc_ok(d, &why)) { njode_set_error("njode_params_per_stack: %s", why); return -1; }
  return njode_make_table(d).stack_floats;
}
extern "C" int32_t njode_num_stacks(const NjodeDesc* d) {
  const char* why = nullptr;
  if (!njode_desc_ok(d, &why)) { njode_set_error("njode_num_stacks: %s", why); return -1; }
  return d->shared_network ? 1 : d->num_moments;
}
extern "C" int64_t njode_param_count(const NjodeDesc* d) {
  const int64_t per = njode_params_per_stack(d);
  return per < 0 ? -1 : per * njode_num_stacks(d);
}

// which flavour runs for this descriptor: 0 = error, else NJODE_IMPL_GENERIC / _TILED / _ROWTILE / _WIDE.
// AUTO prefers the tcgen05 kernels (tiled: hidden 32, one layer; wide: hidden 64 / 128, <= 3 layers), then the
// row-tiled FP32 kernels, then generic.
static int pick_impl(const NjodeDesc* d, const char** why) {
  if (!njode_desc_ok(d, why)) return 0;
  if (d->impl == NJODE_IMPL_TILED) {
    if (!njode_tiled_supported(d)) { *why = "impl=TILED requested but this shape is not supported by the tiled kernels"; return 0; }
    return NJODE_IMPL_TILED;
  }
  if (d->impl == NJODE_IMPL_ROWTILE) {
    if (!njode_rowtile_supported(d)) { *why = "impl=ROWTILE requested but this shape is not supported by the row-tiled kernels"; return 0; }
    return NJODE_IMPL_ROWTILE;
  }
  if (d->impl == NJODE_IMPL_WIDE) {
    if (!njode_wide_supported(d)) { *why = "impl=WIDE requested but this shape is not supported by the wide tcgen05 kernels"; return 0; }
    return NJODE_IMPL_WIDE;
  }
  if (d->impl == NJODE_IMPL_AUTO && njode_tiled_supported(d)) return NJODE_IMPL_TILED;
  if (d->impl == NJODE_IMPL_AUTO && njode_wide_supported(d)) return NJODE_IMPL_WIDE;
  if (d->impl == NJODE_IMPL_AUTO && njode_rowtile_supported(d)) return NJODE_IMPL_ROWTILE;
  if (d->impl != NJODE_IMPL_AUTO && d->impl != NJODE_IMPL_GENERIC) { *why = "unknown impl code"; return 0; }
  if (!njode_generic_supported(d, why)) return 0;
  return NJODE_IMPL_GENERIC;
}

extern "C" int32_t njode_tile_rows(const NjodeDesc* d) {
  const char* why = nullptr;
  const int impl = pick_impl(d, &why);
  if (!impl) { njode_set_error("njode_tile_rows: %s", why); return -1; }
  return (impl == NJODE_IMPL_TILED || impl == NJODE_IMPL_WIDE) ? NJODE_TILED_TILE_ROWS : NJODE_GENERIC_TILE_ROWS;
}

extern "C" int32_t njode_selected_impl(const NjodeDesc* d) {
  const char* why = nullptr;
  const int impl = pick_impl(d, &why);
  if (!impl) { njode_set_error("njode_selected_impl: %s", why); return -1; }
  return impl;
}

int32_t njode_slot_extra(const NjodeDesc* d) {
  const char* why = nullptr;
  return pick_impl(d, &why) == NJODE_IMPL_WIDE ? NJODE_WIDE_XSLOTS : 0;
}

int32_t njode_table_workers(const NjodeDesc* d, int64_t n_tiles) {
  const char* why = nullptr;
  if (pick_impl(d, &why) != NJODE_IMPL_WIDE || n_tiles <= 0) return 0;
  const int S = d->shared_network ? 1 : d->num_moments;
  return njode_wide_workers(d, n_tiles) / S;
}

static int sm_count_abi() {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms;
}

int32_t njode_tile_units(const NjodeDesc* d, int64_t N) {
  const char* why = nullptr;
  const int impl = pick_impl(d, &why);
  if (impl == NJODE_IMPL_WIDE) return NJODE_TILED_TILE_ROWS;     // always full tiles: the weight stream is per tile-step
  if (impl != NJODE_IMPL_TILED) return NJODE_GENERIC_TILE_ROWS;
  const int S = d->shared_network ? 1 : d->num_moments;
  const int64_t full_tiles = (N + NJODE_TILED_TILE_ROWS - 1) / NJODE_TILED_TILE_ROWS;
  const int64_t sms_per_stack = sm_count_abi() / S > 0 ? sm_count_abi() / S : 1;
  // measured: partial tiles pay only while full tiles would leave SMs idle (default workload, 2.2 full tiles per SM:
  // half tiles are 13 % slower; configs[0] at batch 128, 0.14 tiles per SM: quarter tiles are 5 % faster)
  if (4 * full_tiles <= sms_per_stack) return NJODE_TILED_TILE_ROWS / 4;
  if (full_tiles <= sms_per_stack) return NJODE_TILED_TILE_ROWS / 2;
  return NJODE_TILED_TILE_ROWS;
}

TilePlan njode_tile_plan(const NjodeDesc* d, int64_t N) {
  TilePlan p;
  p.units = njode_tile_units(d, N);
  p.units_small = p.units;
  p.n_small = 0;
  const char* why = nullptr;
  if (pick_impl(d, &why) == NJODE_IMPL_TILED && p.units == NJODE_TILED_TILE_ROWS) {
    const int S = d->shared_network ? 1 : d->num_moments;
    const int64_t sms_per_stack = sm_count_abi() / S > 0 ? sm_count_abi() / S : 1;
    const int64_t full_tiles = (N + p.units - 1) / p.units;
    // Up to ~8 full tiles per SM the longest tile is a large part of a sweep's time (default workload: 2.2 full
    // tiles per SM, longest tile 78 steps against an average of 12 per CTA), which suggests one quarter tile per SM
    // for the longest units.  Measured: 0.3292 -> 0.3279 ms per step (a quarter tile's reverse step is 4000 cycles,
    // not the ~2900 its MMA and store counts suggest: all of its rows live in warps of one scheduler) for +43 % of
    // checkpoint bytes (padding rows are stored too).  Off unless NJODE_TAIL_TILES=1.
    static const int tail = [] { const char* e = getenv("NJODE_TAIL_TILES"); return e ? atoi(e) : 0; }();
    if (tail && full_tiles <= 8 * sms_per_stack) {
      p.units_small = NJODE_TILED_TILE_ROWS / 4;
      p.n_small = sms_per_stack;
      if (p.n_small * p.units_small > N) p.n_small = N / p.units_small;
    }
  }
  const int64_t rest = N - p.n_small * p.units_small;
  p.n_tiles = p.n_small + (rest + p.units - 1) / p.units;
  return p;
}

extern "C" int64_t njode_num_tiles(const NjodeDesc* d, int64_t N) {
  const char* why = nullptr;
  if (!pick_impl(d, &why)) { njode_set_error("njode_num_tiles: %s", why); return -1; }
  if (N < 0) { njode_set_error("njode_num_tiles: negative size"); return -1; }
  return njode_tile_plan(d, N).n_tiles;
}

extern "C" int64_t njode_ckpt_row_floats(const NjodeDesc* d) {
  const char* why = nullptr;
  const int impl = pick_impl(d, &why);
  if (!impl) { njode_set_error("njode_ckpt_row_floats: %s", why); return -1; }
  // the tiled kernels also keep the hidden-layer activation of every step (see njode_tiled.cu)
  if (impl == NJODE_IMPL_TILED) return 2 * d->hidden;
  // the row-tiled kernels keep every hidden-layer output of the ODE net (no re-computation in the reverse sweep)
  if (impl == NJODE_IMPL_ROWTILE) return (int64_t)(1 + d->n_hidden_layers) * d->hidden;
  // wide: (L + 1) activation planes + (L + 1) data-gradient planes + 8 aux floats per row and slot (njode_wide.cuh)
  if (impl == NJODE_IMPL_WIDE) return (int64_t)2 * (1 + d->n_hidden_layers) * d->hidden + 8;
  return d->hidden;
}

// ------------------------------------------------------------------------------------------------
// weight re-layout: params_t holds every Linear weight input-major ((n_in) x (n_out)); biases copied
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool locate(const ParamTable& T, int e, int& net, int& l, bool& is_w) {
  for (net = 0; net < 3; ++net)
    for (l = 0; l <= T.L; ++l) {
      const int nin = T.n_vec[net][l] + T.n_ext[net][l], nout = T.n_out[net][l];
      if (e >= T.w_off[net][l] && e < T.w_off[net][l] + nin * nout) { is_w = true; return true; }
      if (e >= T.b_off[net][l] && e < T.b_off[net][l] + nout) { is_w = false; return true; }
    }
  return false;
}

__global__ void k_transpose_params(ParamTable T, const float* __restrict__ p, float* __restrict__ pt, int64_t total) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int64_t s = i / T.stack_floats;
  const int e = (int)(i - s * T.stack_floats);      // index in input-major layout
  int net, l;
  bool is_w;
  if (!locate(T, e, net, l, is_w)) return;
  int64_t src = e;
  if (is_w) {
    const int nin = T.n_vec[net][l] + T.n_ext[net][l], nout = T.n_out[net][l];
    const int rel = e - T.w_off[net][l];
    const int k = rel / nout, j = rel - k * nout;
    src = T.w_off[net][l] + j * nin + k;
  }
  pt[i] = p[s * T.stack_floats + src];
}

// grad[i] = sum over the workers of stack s (fixed order) of their partial; partials are input-major
// (transposed = 1, generic / row-tiled flavours) or PyTorch layout (transposed = 0).
// Block = 32 elements x 32 worker groups: thread (e, g) adds the partials of workers g, g+32, ... of the element's
// stack (independent loads: the 8-group version was a chain of ~18 dependent L2 round trips, 9 us), the 32 group sums
// are combined through shared memory in a fixed order: deterministic for a given schedule.
#define RED_GROUPS 32
__global__ void __launch_bounds__(32 * RED_GROUPS) k_reduce_partials(ParamTable T, const float* __restrict__ partials, int n_workers,
                                                                   int transposed, float* __restrict__ grad, int64_t total) {
  __shared__ float red[RED_GROUPS][33];
  const int e_local = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int64_t i = (int64_t)blockIdx.x * 32 + e_local;
  float acc = 0.0f;
  int64_t s = 0;
  int e = 0;
  if (i < total) {
    s = i / T.stack_floats;
    e = (int)(i - s * T.stack_floats);
    const int per_stack = (n_workers - (int)s + T.S - 1) / T.S;       // workers of this stack: s, s + S, ...
    for (int j = g; j < per_stack; j += RED_GROUPS) acc += partials[(int64_t)((int)s + j * T.S) * T.stack_floats + e];
  }
  red[g][e_local] = acc;
  __syncthreads();
  if (g != 0 || i >= total) return;
  float sum = 0.0f;
#pragma unroll
  for (int k = 0; k < RED_GROUPS; k += 4)
    sum += (red[k][e_local] + red[k + 1][e_local]) + (red[k + 2][e_local] + red[k + 3][e_local]);
  int64_t dst = e;
  if (transposed) {
    int net, l;
    bool is_w;
    if (!locate(T, e, net, l, is_w)) return;
    if (is_w) {
      const int nin = T.n_vec[net][l] + T.n_ext[net][l], nout = T.n_out[net][l];
      const int rel = e - T.w_off[net][l];
      const int k = rel / nout, j = rel - k * nout;
      dst = T.w_off[net][l] + j * nin + k;
    }
  }
  grad[s * T.stack_floats + dst] = sum;
}

// ------------------------------------------------------------------------------------------------
static int check_common(const char* fn, const NjodeDesc* desc, const float* params, const float* times,
                        const float* values, int64_t B, int64_t N, int64_t n_tiles, int32_t tile_rows, int* impl) {
  const char* why = nullptr;
  *impl = pick_impl(desc, &why);
  if (!*impl) NJODE_FAIL(NJODE_EINVAL, "%s: %s", fn, why);
  if (!params || (N > 0 && (!times || !values))) NJODE_FAIL(NJODE_EINVAL, "%s: null input pointer", fn);
  if (B < 0 || N < 0) NJODE_FAIL(NJODE_EINVAL, "%s: negative size", fn);
  const int want = (*impl == NJODE_IMPL_TILED || *impl == NJODE_IMPL_WIDE) ? NJODE_TILED_TILE_ROWS : NJODE_GENERIC_TILE_ROWS;
  if (tile_rows != want) NJODE_FAIL(NJODE_EINVAL, "%s: schedule was built for tile_rows=%d, kernels need %d", fn, tile_rows, want);
  if (n_tiles != njode_tile_plan(desc, N).n_tiles) NJODE_FAIL(NJODE_EINVAL, "%s: n_tiles does not match N (njode_num_tiles)", fn);
  return NJODE_OK;
}

static SweepArgs make_args(const NjodeDesc* desc, const float* params, const float* params_t, const float* times,
                           const float* values, const int32_t* kenc, const int32_t* perm, const int32_t* tile_kmax,
                           const int64_t* tile_slot_off, const float* knots, int64_t N, int64_t n_tiles,
                           int64_t total_slots, int32_t tile_rows) {
  SweepArgs a;
  memset(&a, 0, sizeof(a));
  a.desc = *desc;
  a.T = njode_make_table(desc);
  a.params = params; a.params_t = params_t; a.times = times; a.values = values;
  a.kenc = kenc; a.perm = perm; a.tile_kmax = tile_kmax; a.tile_slot_off = tile_slot_off; a.knots = knots;
  a.N = N; a.n_tiles = n_tiles; a.total_slots = total_slots; a.tile_rows = tile_rows;
  const TilePlan plan = njode_tile_plan(desc, N);
  a.tile_units = plan.units; a.tile_units_small = plan.units_small; a.n_small_tiles = plan.n_small;
  a.table_workers = njode_table_workers(desc, n_tiles);
  a.tile_table = a.table_workers > 0 ? (const int32_t*)(tile_slot_off + n_tiles + 1) : nullptr;
  return a;
}

static size_t params_bytes(const NjodeDesc* d) {
  return njode_align_up((size_t)njode_param_count(d) * sizeof(float), 256);
}

// wide flavour: the workspace holds the split / swizzled weight images instead of the transposed parameters
static size_t relayout_bytes(const NjodeDesc* desc) {
  const char* why = nullptr;
  if (pick_impl(desc, &why) == NJODE_IMPL_WIDE) return njode_wide_image_bytes(desc) + 1024;
  return params_bytes(desc);
}
static float* image_ptr(void* workspace) { return (float*)(((uintptr_t)workspace + 1023) & ~(uintptr_t)1023); }

extern "C" size_t njode_forward_workspace_bytes(const NjodeDesc* desc) {
  if (njode_param_count(desc) < 0) return 0;
  return relayout_bytes(desc);
}

static int workers_for(const NjodeDesc* desc, int impl, int64_t n_tiles) {
  if (impl == NJODE_IMPL_TILED) return njode_tiled_workers(desc, n_tiles);
  if (impl == NJODE_IMPL_WIDE) return njode_wide_workers(desc, n_tiles);
  if (impl == NJODE_IMPL_ROWTILE) return njode_rowtile_workers(desc, n_tiles);
  return njode_generic_workers(desc, n_tiles);
}

extern "C" size_t njode_backward_workspace_bytes(const NjodeDesc* desc, int64_t n_tiles) {
  const char* why = nullptr;
  const int impl = pick_impl(desc, &why);
  if (!impl) { njode_set_error("njode_backward_workspace_bytes: %s", why); return 0; }
  const ParamTable T = njode_make_table(desc);
  const int nw = workers_for(desc, impl, n_tiles);
  return njode_align_up(relayout_bytes(desc), 256) + njode_align_up((size_t)nw * T.stack_floats * sizeof(float), 256);
}

static int relayout(const NjodeDesc* desc, const float* params, float* params_t, cudaStream_t st) {
  const ParamTable T = njode_make_table(desc);
  const int64_t total = (int64_t)T.stack_floats * T.S;
  k_transpose_params<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(T, params, params_t, total);
  NJODE_LAUNCH_OK("k_transpose_params");
  return NJODE_OK;
}

extern "C" int njode_forward(const NjodeDesc* desc, const float* params, const float* times, const float* values,
                             const int64_t* obs_offsets, int64_t B, int64_t N,
                             const int32_t* kenc, const int32_t* perm, const int32_t* tile_kmax,
                             const int64_t* tile_slot_off, const float* knots,
                             int64_t n_tiles, int64_t total_slots, int32_t tile_rows,
                             float* preds, float* preds_before, float* ckpt,
                             void* workspace, size_t workspace_bytes, void* stream) {
  int impl = 0;
  int rc = check_common("njode_forward", desc, params, times, values, B, N, n_tiles, tile_rows, &impl);
  if (rc) return rc;
  (void)obs_offsets;
  if (N > 0 && (!preds || !preds_before || !kenc || !perm || !tile_slot_off || !knots))
    NJODE_FAIL(NJODE_EINVAL, "njode_forward: null schedule/output pointer");
  if (workspace_bytes < njode_forward_workspace_bytes(desc)) NJODE_FAIL(NJODE_EWORKSPACE, "njode_forward: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  if (N == 0) return NJODE_OK;
  float* params_t = (float*)workspace;
  if (impl != NJODE_IMPL_TILED && impl != NJODE_IMPL_WIDE) {   // the tcgen05 kernels build their own (split, swizzled) weight tiles
    rc = relayout(desc, params, params_t, st);
    if (rc) return rc;
  }
  SweepArgs a = make_args(desc, params, params_t, times, values, kenc, perm, tile_kmax, tile_slot_off, knots, N, n_tiles,
                          total_slots, tile_rows);
  a.preds = preds; a.preds_before = preds_before; a.ckpt = ckpt;
  a.n_workers = workers_for(desc, impl, n_tiles);
  // preds_before of each trajectory's first observation stays 0 (jump_ode.py:161)
  NJODE_CUDA_OK(cudaMemsetAsync(preds_before, 0, (size_t)N * desc->d_y * desc->num_moments * sizeof(float), st));
  if (impl == NJODE_IMPL_TILED) return njode_tiled_forward(a, st);
  if (impl == NJODE_IMPL_WIDE) return njode_wide_forward(a, image_ptr(workspace), st);
  if (impl == NJODE_IMPL_ROWTILE) return njode_rowtile_forward(a, st);
  return njode_generic_forward(a, st);
}

// ------------------------------------------------------------------------------------------------
// dense-grid inference (njode_generic.cu: k_generic_dense)
extern "C" size_t njode_dense_workspace_bytes(const NjodeDesc* desc) {
  if (njode_param_count(desc) < 0) return 0;
  return params_bytes(desc);
}

extern "C" int njode_dense_forward(const NjodeDesc* desc, const float* params, const float* times, const float* values,
                                   const int64_t* obs_offsets, int64_t B, int64_t N, const float* grid, int64_t G,
                                   float* dense, void* workspace, size_t workspace_bytes, void* stream) {
  const char* why = nullptr;
  if (!njode_desc_ok(desc, &why)) NJODE_FAIL(NJODE_EINVAL, "njode_dense_forward: %s", why);
  if (!njode_generic_supported(desc, &why)) NJODE_FAIL(NJODE_EINVAL, "njode_dense_forward: %s", why);
  if (B < 0 || N < 0 || G < 0) NJODE_FAIL(NJODE_EINVAL, "njode_dense_forward: negative size");
  if (!params || !dense || (N > 0 && (!times || !values || !obs_offsets)) || (G > 0 && !grid))
    NJODE_FAIL(NJODE_EINVAL, "njode_dense_forward: null pointer");
  if (workspace_bytes < njode_dense_workspace_bytes(desc)) NJODE_FAIL(NJODE_EWORKSPACE, "njode_dense_forward: workspace too small");
  if (N * (desc->shared_network ? 1 : desc->num_moments) >= (1ll << 31)) NJODE_FAIL(NJODE_EINVAL, "njode_dense_forward: too many observations for one launch");
  cudaStream_t st = (cudaStream_t)stream;
  NJODE_CUDA_OK(cudaMemsetAsync(dense, 0, (size_t)B * G * desc->d_y * desc->num_moments * sizeof(float), st));
  if (N == 0 || G == 0) return NJODE_OK;
  float* params_t = (float*)workspace;
  int rc = relayout(desc, params, params_t, st);
  if (rc) return rc;
  return njode_generic_dense(desc, params, params_t, times, values, obs_offsets, B, N, grid, G, dense, st);
}

// ------------------------------------------------------------------------------------------------
// one call for an un-cached batch (see include/njode.h)
// ------------------------------------------------------------------------------------------------
extern "C" size_t njode_batch_arena_bytes(const NjodeDesc* desc, int64_t B, int64_t N, int64_t total_slots, int64_t* layout) {
  (void)B;
  const int tile_rows = njode_tile_rows(desc);
  const int64_t n_tiles = njode_num_tiles(desc, N);
  if (tile_rows < 1 || n_tiles < 0 || total_slots < 0) return 0;
  size_t o = 0;
  int64_t lay[NJODE_ARENA_WORDS] = {0};
  auto take = [&](int word, size_t bytes) { lay[word] = (int64_t)o; o += njode_align_up(bytes > 0 ? bytes : 4, 256); };
  take(NJODE_ARENA_KENC, (size_t)N * sizeof(int32_t));
  take(NJODE_ARENA_PERM, (size_t)n_tiles * tile_rows * sizeof(int32_t));
  take(NJODE_ARENA_TILE_KMAX, (size_t)n_tiles * sizeof(int32_t));
  // (+ the wide flavour's tile -> worker table behind the slot offsets)
  take(NJODE_ARENA_TILE_SLOT_OFF, (size_t)(n_tiles + 1) * sizeof(int64_t) +
                                      njode_table_ints(njode_table_workers(desc, n_tiles), n_tiles) * sizeof(int32_t));
  take(NJODE_ARENA_HEADER, NJODE_HDR_WORDS * sizeof(int64_t));
  take(NJODE_ARENA_KNOTS, (size_t)total_slots * tile_rows * sizeof(float));
  if (layout) memcpy(layout, lay, sizeof(lay));
  return o;
}

extern "C" size_t njode_batch_scratch_bytes(const NjodeDesc* desc, int64_t B, int64_t N) {
  const int tile_rows = njode_tile_rows(desc);
  if (tile_rows < 1) return 0;
  return njode_align_up(njode_schedule_workspace_bytes(B, N, tile_rows), 256) + njode_align_up(njode_forward_workspace_bytes(desc), 256);
}

// arena pointers of a batch (fixed part; knots follow)
struct ArenaView {
  int32_t *kenc, *perm, *tile_kmax;
  int64_t *tile_slot_off, *header;
  float* knots;
};
static ArenaView arena_view(void* arena, const int64_t* lay) {
  char* base = (char*)arena;
  ArenaView v;
  v.kenc = (int32_t*)(base + lay[NJODE_ARENA_KENC]);
  v.perm = (int32_t*)(base + lay[NJODE_ARENA_PERM]);
  v.tile_kmax = (int32_t*)(base + lay[NJODE_ARENA_TILE_KMAX]);
  v.tile_slot_off = (int64_t*)(base + lay[NJODE_ARENA_TILE_SLOT_OFF]);
  v.header = (int64_t*)(base + lay[NJODE_ARENA_HEADER]);
  v.knots = (float*)(base + lay[NJODE_ARENA_KNOTS]);
  return v;
}

extern "C" int njode_forward_batch_begin(const NjodeDesc* desc, const float* times, const int64_t* obs_offsets,
                                         int64_t B, int64_t N, void* arena, size_t arena_bytes,
                                         void* scratch, size_t scratch_bytes, int64_t* header_host, void* stream) {
  const int tile_rows = njode_tile_rows(desc);
  if (tile_rows < 1) return NJODE_EINVAL;                      // (njode_tile_rows set the error text)
  if (!arena || !scratch || !header_host) NJODE_FAIL(NJODE_EINVAL, "njode_forward_batch: null arena / scratch / header_host");
  int64_t lay[NJODE_ARENA_WORDS];
  const size_t fixed = njode_batch_arena_bytes(desc, B, N, 0, lay);
  if (arena_bytes < fixed) NJODE_FAIL(NJODE_EWORKSPACE, "njode_forward_batch: arena smaller than its size-independent part (njode_batch_arena_bytes(..., 0))");
  if (scratch_bytes < njode_batch_scratch_bytes(desc, B, N)) NJODE_FAIL(NJODE_EWORKSPACE, "njode_forward_batch: scratch too small");
  const ArenaView v = arena_view(arena, lay);
  const size_t sched_ws = njode_align_up(njode_schedule_workspace_bytes(B, N, tile_rows), 256);
  int rc = njode_schedule_build(desc, times, obs_offsets, B, N, tile_rows, v.kenc, v.perm, v.tile_kmax, v.tile_slot_off, v.header,
                                scratch, sched_ws, stream);
  if (rc) return rc;
  NJODE_CUDA_OK(cudaMemcpyAsync(header_host, v.header, NJODE_HDR_WORDS * sizeof(int64_t), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  return NJODE_OK;
}

extern "C" int njode_forward_batch_finish(const NjodeDesc* desc, const float* params, const float* times, const float* values,
                                          const int64_t* obs_offsets, int64_t B, int64_t N,
                                          void* arena, size_t arena_bytes, int32_t want_ckpt, float* ckpt, int64_t ckpt_floats,
                                          void* scratch, size_t scratch_bytes, int64_t* header_host,
                                          float* preds, float* preds_before, void* stream) {
  const int tile_rows = njode_tile_rows(desc);
  if (tile_rows < 1) return NJODE_EINVAL;
  if (!arena || !scratch || !header_host) NJODE_FAIL(NJODE_EINVAL, "njode_forward_batch: null arena / scratch / header_host");
  if (scratch_bytes < njode_batch_scratch_bytes(desc, B, N)) NJODE_FAIL(NJODE_EWORKSPACE, "njode_forward_batch: scratch too small");
  int64_t lay[NJODE_ARENA_WORDS];
  njode_batch_arena_bytes(desc, B, N, 0, lay);
  const ArenaView v = arena_view(arena, lay);
  NJODE_CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream));          // the schedule header is on the host now
  const int64_t total_slots = header_host[NJODE_HDR_TOTAL_SLOTS];
  const int64_t n_tiles = njode_num_tiles(desc, N);
  const int S = desc->shared_network ? 1 : desc->num_moments;
  const int64_t need_ckpt = want_ckpt ? (int64_t)S * total_slots * tile_rows * njode_ckpt_row_floats(desc) : 0;
  if (arena_bytes < njode_batch_arena_bytes(desc, B, N, total_slots, nullptr) || (want_ckpt && (ckpt_floats < need_ckpt || (!ckpt && need_ckpt > 0))))
    NJODE_FAIL(NJODE_ECAPACITY, "njode_forward_batch: this batch has %lld checkpoint slots; arena / ckpt are too small for it", (long long)total_slots);
  int rc = njode_schedule_knots(times, v.kenc, v.perm, v.tile_kmax, v.tile_slot_off, N, n_tiles, tile_rows, desc, v.knots, stream);
  if (rc) return rc;
  const size_t sched_ws = njode_align_up(njode_schedule_workspace_bytes(B, N, tile_rows), 256);
  return njode_forward(desc, params, times, values, obs_offsets, B, N, v.kenc, v.perm, v.tile_kmax, v.tile_slot_off, v.knots, n_tiles,
                       total_slots, tile_rows, preds, preds_before, want_ckpt ? ckpt : nullptr,
                       (char*)scratch + sched_ws, scratch_bytes - sched_ws, stream);
}

extern "C" int njode_forward_batch(const NjodeDesc* desc, const float* params, const float* times, const float* values,
                                   const int64_t* obs_offsets, int64_t B, int64_t N,
                                   void* arena, size_t arena_bytes, int32_t want_ckpt, float* ckpt, int64_t ckpt_floats,
                                   void* scratch, size_t scratch_bytes, int64_t* header_host,
                                   float* preds, float* preds_before, void* stream) {
  int rc = njode_forward_batch_begin(desc, times, obs_offsets, B, N, arena, arena_bytes, scratch, scratch_bytes, header_host, stream);
  if (rc) return rc;
  return njode_forward_batch_finish(desc, params, times, values, obs_offsets, B, N, arena, arena_bytes, want_ckpt, ckpt, ckpt_floats,
                                    scratch, scratch_bytes, header_host, preds, preds_before, stream);
}

extern "C" int njode_backward(const NjodeDesc* desc, const float* params, const float* times, const float* values,
                              const int64_t* obs_offsets, int64_t B, int64_t N,
                              const int32_t* kenc, const int32_t* perm, const int32_t* tile_kmax,
                              const int64_t* tile_slot_off, const float* knots,
                              int64_t n_tiles, int64_t total_slots, int32_t tile_rows,
                              const float* grad_preds, const float* grad_preds_before, float* ckpt,
                              float* grad_params, void* workspace, size_t workspace_bytes, void* stream) {
  int impl = 0;
  int rc = check_common("njode_backward", desc, params, times, values, B, N, n_tiles, tile_rows, &impl);
  if (rc) return rc;
  (void)obs_offsets;
  if (!grad_params) NJODE_FAIL(NJODE_EINVAL, "njode_backward: null grad_params");
  cudaStream_t st = (cudaStream_t)stream;
  const ParamTable T = njode_make_table(desc);
  const int64_t total = (int64_t)T.stack_floats * T.S;
  if (N == 0) { NJODE_CUDA_OK(cudaMemsetAsync(grad_params, 0, total * sizeof(float), st)); return NJODE_OK; }
  if (!grad_preds || !grad_preds_before || !ckpt || !kenc || !perm || !tile_slot_off || !knots)
    NJODE_FAIL(NJODE_EINVAL, "njode_backward: null pointer (checkpoints are required; run forward with ckpt != NULL)");
  if (workspace_bytes < njode_backward_workspace_bytes(desc, n_tiles))
    NJODE_FAIL(NJODE_EWORKSPACE, "njode_backward: workspace too small");
  float* params_t = (float*)workspace;
  float* partials = (float*)((char*)workspace + njode_align_up(relayout_bytes(desc), 256));
  if (impl != NJODE_IMPL_TILED && impl != NJODE_IMPL_WIDE) {
    rc = relayout(desc, params, params_t, st);
    if (rc) return rc;
  }
  SweepArgs a = make_args(desc, params, params_t, times, values, kenc, perm, tile_kmax, tile_slot_off, knots, N, n_tiles,
                          total_slots, tile_rows);
  a.grad_preds = grad_preds; a.grad_preds_before = grad_preds_before; a.ckpt = ckpt;
  a.partials = partials;
  a.n_workers = workers_for(desc, impl, n_tiles);
  NJODE_CUDA_OK(cudaMemsetAsync(partials, 0, (size_t)a.n_workers * T.stack_floats * sizeof(float), st));
  rc = impl == NJODE_IMPL_TILED ? njode_tiled_backward(a, st)
     : impl == NJODE_IMPL_WIDE ? njode_wide_backward(a, image_ptr(workspace), st)
     : impl == NJODE_IMPL_ROWTILE ? njode_rowtile_backward(a, st) : njode_generic_backward(a, st);
  if (rc) return rc;
  const int pytorch_layout = (impl == NJODE_IMPL_TILED || impl == NJODE_IMPL_WIDE) ? 1 : 0;
  k_reduce_partials<<<(unsigned)((total + 31) / 32), 32 * RED_GROUPS, 0, st>>>(T, partials, a.n_workers,
                                                                    pytorch_layout ? 0 : 1, grad_params, total);
  NJODE_LAUNCH_OK("k_reduce_partials");
  return NJODE_OK;
}
