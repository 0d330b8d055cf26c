// njode_schedule.cu -- the float32 Euler step rule of the reference (jump_ode.py:188-203) on device,
// and the K-descending tiling of observation units that the sweep kernels run on.
//
// The rule, reproduced bit for bit:   t = t_i;  while (fl32(t + dt) < t_next) t = fl32(t + dt);   // full steps
//                                     if (t < t_next) one closing step to exactly t_next
// About 65 % of intervals get a closing step of ~1e-8 because k float32 additions of dt land a few
// ulp short of t_next; "observation indexing must be bit-exact" therefore means exact __fadd_rn and
// the two exact comparisons -- never FMA contraction, never double.
#include <cub/device/device_radix_sort.cuh>

#include "njode_common.cuh"

#define NJODE_BINS 2048
#define NJODE_BIN_BITS 11
#define NJODE_KCAP (1 << 24)

__device__ __forceinline__ int count_steps(float t0, float t1, int has_dt, float dt) {
  if (!has_dt) return 1;                               // jump_ode.py:188-190
  int K = 0;
  float t = t0;
  while (true) {
    const float tn = __fadd_rn(t, dt);
    if (!(tn < t1)) break;                             // jump_ode.py:196
    if (tn == t || K >= NJODE_KCAP) break;             // the reference would never terminate here
    t = tn;
    ++K;
  }
  if (t < t1) ++K;                                     // jump_ode.py:201
  return K;
}

// one thread per trajectory: step counts of its intervals, the sort key / value of every unit, total step count.
// key = BINS-1 - min(K, BINS-1): an ascending STABLE sort of the keys puts the longest units first (longest-
// processing-time order) and keeps the batch order inside a step count, so the tiling -- and with it every
// summation order downstream -- is the same on every run (an atomic-cursor counting sort is not).
__global__ void k_count_steps(const float* __restrict__ times, const int64_t* __restrict__ off, int64_t B,
                              int has_dt, float dt, int32_t* __restrict__ kenc, uint32_t* __restrict__ keys,
                              int32_t* __restrict__ vals, unsigned long long* __restrict__ header) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long total = 0;
  if (b < B) {
    const int64_t lo = off[b], hi = off[b + 1];
    for (int64_t o = lo; o < hi; ++o) {
      int K = 0, has_next = 0;
      if (o + 1 < hi) {
        K = count_steps(times[o], times[o + 1], has_dt, dt);
        has_next = 1;
      }
      kenc[o] = (K << 1) | has_next;
      total += (unsigned long long)K;
      keys[o] = (uint32_t)(NJODE_BINS - 1 - min(K, NJODE_BINS - 1));
      vals[o] = (int32_t)o;
    }
  }
  // warp-reduce the step total before the single atomic
  for (int s = 16; s > 0; s >>= 1) total += __shfl_xor_sync(NJODE_FULL, total, s);
  if ((threadIdx.x & 31) == 0 && total) atomicAdd(&header[NJODE_HDR_TOTAL_STEPS], total);
}

// sorted unit list -> tiles of `tile_rows` rows holding plan.units_of(tile) units each (the other rows hold no unit: -1),
// and the largest step count of each tile.  One block per tile, one thread per row (tile_rows = 32 or 128).
__global__ void k_spread_perm(const int32_t* __restrict__ sorted, const int32_t* __restrict__ kenc, int64_t N, int tile_rows,
                              TilePlan plan, int32_t* __restrict__ perm, int32_t* __restrict__ tile_kmax) {
  __shared__ int warp_max[32];
  const int64_t tile = blockIdx.x;
  const int r = threadIdx.x;
  const int64_t j = plan.first_unit(tile) + r;
  const int u = (r < plan.units_of(tile) && j < N) ? sorted[j] : -1;
  perm[tile * tile_rows + r] = u;
  int km = u >= 0 ? (kenc[u] >> 1) : 0;
  for (int o = 16; o > 0; o >>= 1) km = max(km, __shfl_xor_sync(0xffffffffu, km, o));
  if ((r & 31) == 0) warp_max[r >> 5] = km;
  __syncthreads();
  if (r == 0) {
    for (int w = 1; w < (tile_rows + 31) / 32; ++w) km = max(km, warp_max[w]);
    tile_kmax[tile] = km;
  }
}

// single block: exclusive scan of (kmax + 1 + slots_extra) over tiles -> checkpoint slot offsets; header totals
__global__ void k_tile_scan(const int32_t* __restrict__ tile_kmax, int64_t n_tiles, int slots_extra,
                            int64_t* __restrict__ slot_off, long long* __restrict__ header) {
  __shared__ long long part[1024];
  __shared__ int kmx[1024];
  const int tid = threadIdx.x;
  const int64_t chunk = (n_tiles + blockDim.x - 1) / blockDim.x;
  const int64_t lo = min((int64_t)tid * chunk, n_tiles), hi = min(lo + chunk, n_tiles);
  long long s = 0;
  int km = 0;
  for (int64_t t = lo; t < hi; ++t) { s += tile_kmax[t] + 1 + slots_extra; km = max(km, tile_kmax[t]); }
  part[tid] = s;
  kmx[tid] = km;
  __syncthreads();
  if (tid == 0) {
    long long acc = 0;
    int m = 0;
    for (int i = 0; i < (int)blockDim.x; ++i) { long long v = part[i]; part[i] = acc; acc += v; m = max(m, kmx[i]); }
    slot_off[n_tiles] = acc;
    header[NJODE_HDR_TOTAL_SLOTS] = acc;
    header[NJODE_HDR_NUM_TILES] = n_tiles;
    header[NJODE_HDR_KMAX] = m;
  }
  __syncthreads();
  long long acc = part[tid];
  for (int64_t t = lo; t < hi; ++t) { slot_off[t] = acc; acc += tile_kmax[t] + 1 + slots_extra; }
}

// Tile -> worker assignment of the wide flavour.  Tiles are sorted by step count, and the longest few are outliers
// (config-4 shape, 2048 trajectories: kmax 244, 127, 116, ... against a mean of 21), so dealing them out in a fixed
// snake leaves the worker that got tile 0 with 1.5-1.7x the mean work and the sweep kernels wait for it (measured per
// CTA: tools/cta_balance.py).  Greedy longest-processing-time: every tile, in descending order of cost, goes to the
// worker with the least work so far (ties: lowest worker id, so the result is deterministic).  One warp: lane l keeps
// the loads of workers l, l + 32, ... in registers, the argmin is a shuffle reduction.
// table = [n_w + 1 offsets][n_tiles tile ids, grouped by worker, in processing order][n_tiles scratch]
__global__ void __launch_bounds__(32) k_lpt_assign(const int32_t* __restrict__ tile_kmax, int64_t n_tiles, int n_w, int cost_per_step,
                                                   int cost_fixed, int32_t* __restrict__ table) {
  const int lane = threadIdx.x;
  constexpr int PER = 8;                                   // up to 256 workers
  unsigned int load[PER], count[PER];
#pragma unroll
  for (int i = 0; i < PER; ++i) { load[i] = 0; count[i] = 0; }
  int32_t* const off = table;
  int32_t* const list = table + n_w + 1;
  int32_t* const owner = list + n_tiles;
  for (int64_t t = 0; t < n_tiles; ++t) {
    unsigned long long best = ~0ull;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int w = lane + 32 * i;
      if (w < n_w) {
        const unsigned long long k = ((unsigned long long)load[i] << 16) | (unsigned)w;
        best = k < best ? k : best;
      }
    }
    for (int sft = 16; sft > 0; sft >>= 1) {
      const unsigned long long o = __shfl_xor_sync(0xffffffffu, best, sft);
      best = o < best ? o : best;
    }
    const int w = (int)(best & 0xffffu);
    const unsigned cost = (unsigned)(tile_kmax[t] * cost_per_step + cost_fixed);
#pragma unroll
    for (int i = 0; i < PER; ++i)
      if (w == lane + 32 * i) { load[i] += cost; count[i] += 1; }
    if (lane == 0) owner[t] = w;
  }
  // offsets: exclusive scan of the counts in worker order (w = lane + 32 i: scan over i-major blocks of 32)
  unsigned int base = 0;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    unsigned int c = (lane + 32 * i < n_w) ? count[i] : 0, inc = c;
    for (int sft = 1; sft < 32; sft <<= 1) {
      const unsigned int o = __shfl_up_sync(0xffffffffu, inc, sft);
      if (lane >= sft) inc += o;
    }
    if (lane + 32 * i < n_w) off[lane + 32 * i] = (int32_t)(base + inc - c);
    base += __shfl_sync(0xffffffffu, inc, 31);
  }
  if (lane == 0) off[n_w] = (int32_t)n_tiles;
  __syncwarp();
  // scatter, keeping every worker's tiles in descending order of cost (= ascending tile id)
  for (int w0 = 0; w0 < n_w; w0 += 32) {                    // lane handles worker w0 + lane: walks the owner array once per 32 workers
    const int w = w0 + lane;
    int pos = w < n_w ? off[w] : 0;
    for (int64_t t = 0; t < n_tiles; ++t) {
      if (w < n_w && owner[t] == w) list[pos++] = (int32_t)t;
    }
  }
}

// one thread per (tile,row): the float32 knots t_0..t_kmax of that row
__global__ void k_fill_knots(const float* __restrict__ times, const int32_t* __restrict__ kenc,
                             const int32_t* __restrict__ perm, const int32_t* __restrict__ tile_kmax,
                             const int64_t* __restrict__ slot_off, int64_t Npad, int tile_rows, int slots_extra,
                             int has_dt, float dt, float* __restrict__ knots) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= Npad) return;
  const int64_t tile = idx / tile_rows;
  const int r = (int)(idx - tile * tile_rows);
  const int km = tile_kmax[tile];
  float* dst = knots + slot_off[tile] * tile_rows + r;
  const int u = perm[idx];
  if (u < 0) {
    for (int k = 0; k <= km + slots_extra; ++k) dst[(int64_t)k * tile_rows] = 0.0f;
    return;
  }
  const int ke = kenc[u];
  const int K = ke >> 1;
  float t = times[u];
  const float t1 = (ke & 1) ? times[u + 1] : t;
  for (int k = 0; k <= km; ++k) {
    dst[(int64_t)k * tile_rows] = t;
    if (k < K) t = (k == K - 1) ? t1 : __fadd_rn(t, dt);   // closing step lands exactly on t_next
  }
  for (int k = km + 1; k <= km + slots_extra; ++k) dst[(int64_t)k * tile_rows] = t;     // (a flavour's extra slots: defined, unused)
  (void)has_dt;
}

// ------------------------------------------------------------------------------------------------
static size_t sort_temp_bytes(int64_t N) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const int32_t*)nullptr,
                                  (int32_t*)nullptr, (int)N, 0, NJODE_BIN_BITS);
  return bytes;
}

// workspace: sort keys in / out, unit indices in (the sorted indices go straight to `perm`), cub scratch
extern "C" size_t njode_schedule_workspace_bytes(int64_t B, int64_t N, int32_t tile_rows) {
  (void)B; (void)tile_rows;
  if (N <= 0) return 256;
  return 4 * njode_align_up((size_t)N * sizeof(int32_t), 256) + njode_align_up(sort_temp_bytes(N), 256);
}

// NOTE: for the wide flavour `tile_slot_off` must have room for the tile table behind its n_tiles + 1 entries
// (njode_table_ints(njode_table_workers(..)) int32); njode_batch_arena_bytes lays the arena out that way.
extern "C" int njode_schedule_build(const NjodeDesc* desc, const float* times, const int64_t* obs_offsets,
                                    int64_t B, int64_t N, int32_t tile_rows,
                                    int32_t* kenc, int32_t* perm, int32_t* tile_kmax, int64_t* tile_slot_off,
                                    int64_t* header, void* workspace, size_t workspace_bytes, void* stream) {
  const char* why = nullptr;
  if (!njode_desc_ok(desc, &why)) NJODE_FAIL(NJODE_EINVAL, "njode_schedule_build: %s", why);
  if (B < 0 || N < 0 || tile_rows < 1) NJODE_FAIL(NJODE_EINVAL, "njode_schedule_build: bad sizes");
  if (N >= (1ll << 31) - 4096) NJODE_FAIL(NJODE_EINVAL, "njode_schedule_build: N must be < 2^31");
  if (workspace_bytes < njode_schedule_workspace_bytes(B, N, tile_rows))
    NJODE_FAIL(NJODE_EWORKSPACE, "njode_schedule_build: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const TilePlan plan = njode_tile_plan(desc, N);
  if (plan.units > tile_rows) NJODE_FAIL(NJODE_EINVAL, "njode_schedule_build: tile_rows does not match this descriptor (njode_tile_rows)");
  const int64_t n_tiles = plan.n_tiles;
  NJODE_CUDA_OK(cudaMemsetAsync(header, 0, NJODE_HDR_WORDS * sizeof(int64_t), st));
  if (N == 0) return NJODE_OK;
  const size_t seg = njode_align_up((size_t)N * sizeof(int32_t), 256);
  uint32_t* keys_in = (uint32_t*)workspace;
  uint32_t* keys_out = (uint32_t*)((char*)workspace + seg);
  int32_t* vals_in = (int32_t*)((char*)workspace + 2 * seg);
  int32_t* sorted = (int32_t*)((char*)workspace + 3 * seg);
  void* temp = (char*)workspace + 4 * seg;
  size_t temp_bytes = sort_temp_bytes(N);
  const int TB = 128;
  k_count_steps<<<(unsigned)((B + TB - 1) / TB), TB, 0, st>>>(times, obs_offsets, B, desc->has_dt, desc->dt, kenc, keys_in,
                                                             vals_in, (unsigned long long*)header);
  NJODE_LAUNCH_OK("k_count_steps");
  // stable LSD radix sort over the 11 key bits: perm = unit indices, longest first, batch order within a step count
  NJODE_CUDA_OK(cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys_in, keys_out, vals_in, sorted, (int)N, 0,
                                                NJODE_BIN_BITS, st));
  njode_count_launch(2);
  if (tile_rows > 1024 || tile_rows % 32 != 0) NJODE_FAIL(NJODE_EINVAL, "njode_schedule_build: tile_rows must be a multiple of 32, at most 1024");
  k_spread_perm<<<(unsigned)n_tiles, tile_rows, 0, st>>>(sorted, kenc, N, tile_rows, plan, perm, tile_kmax);
  NJODE_LAUNCH_OK("k_spread_perm");
  k_tile_scan<<<1, 1024, 0, st>>>(tile_kmax, n_tiles, njode_slot_extra(desc), tile_slot_off, (long long*)header);
  NJODE_LAUNCH_OK("k_tile_scan");
  const int n_w = njode_table_workers(desc, n_tiles);
  if (n_w > 0) {      // the table lives behind the n_tiles + 1 slot offsets (njode_batch_arena_bytes sizes the region for it)
    if (n_w > 256) NJODE_FAIL(NJODE_EINVAL, "njode_schedule_build: more than 256 workers per stack");
    k_lpt_assign<<<1, 32, 0, st>>>(tile_kmax, n_tiles, n_w, desc->n_hidden_layers + 1, 3 * desc->n_hidden_layers,
                                   (int32_t*)(tile_slot_off + n_tiles + 1));
    NJODE_LAUNCH_OK("k_lpt_assign");
  }
  return NJODE_OK;
}

extern "C" int njode_schedule_knots(const float* times, const int32_t* kenc, const int32_t* perm,
                                    const int32_t* tile_kmax, const int64_t* tile_slot_off,
                                    int64_t N, int64_t n_tiles, int32_t tile_rows, const NjodeDesc* desc,
                                    float* knots, void* stream) {
  const char* why = nullptr;
  if (!njode_desc_ok(desc, &why)) NJODE_FAIL(NJODE_EINVAL, "njode_schedule_knots: %s", why);
  if (N == 0 || n_tiles == 0) return NJODE_OK;
  const int64_t Npad = n_tiles * tile_rows;
  k_fill_knots<<<(unsigned)((Npad + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      times, kenc, perm, tile_kmax, tile_slot_off, Npad, tile_rows, njode_slot_extra(desc), desc->has_dt, desc->dt, knots);
  NJODE_LAUNCH_OK("k_fill_knots");
  return NJODE_OK;
}
