// njode_generic.cu -- shape-generic sweep kernels (any H <= 256, any L <= 7, any d_x/d_y/M).
//
// One warp per observation unit and per network stack.  A hidden vector is distributed over the
// lanes (lane holds elements lane, lane+32, ...); a Linear layer is a broadcast-by-shuffle matvec
// against weights stored input-major ("kernel layout", coalesced over output neurons).  The reverse
// sweep re-computes the layer activations of each Euler step from the checkpointed hidden state and
// accumulates weight gradients into a per-CTA private partial buffer (plain read-modify-write, one
// owner, no atomics); a final kernel sums the partials in a fixed order -> deterministic given the
// schedule.  This flavour covers every configuration the reference accepts; the tiled flavour
// (njode_tiled.cu) is the tuned FP32-FMA path for the benchmark shapes.
#include "njode_common.cuh"

#define GEN_NJ_MAX 8                 // H <= 256
#define GEN_EXT_MAX 10               // d_x + 2 <= 10
#define GEN_WARPS_PER_SM 24

template <int NJ>
struct Vec { float v[NJ]; };

// out = [act](b + Wt^T [in ; ext]);  Wt is (n_vec+n_ext) x n_out, input-major
template <int NJ>
__device__ __forceinline__ void layer_fwd(const ParamTable& T, int net, int l, const float* __restrict__ pt,
                                          const float* __restrict__ p, const Vec<NJ>& in,
                                          const float* ext, int act_kind, Vec<NJ>& out) {
  const int lane = threadIdx.x & 31;
  const int n_vec = T.n_vec[net][l], n_ext = T.n_ext[net][l], n_out = T.n_out[net][l];
  const float* __restrict__ Wt = pt + T.w_off[net][l];
  const float* __restrict__ b = p + T.b_off[net][l];
#pragma unroll
  for (int q = 0; q < NJ; ++q) { const int j = lane + 32 * q; out.v[q] = j < n_out ? b[j] : 0.0f; }
#pragma unroll
  for (int kq = 0; kq < NJ; ++kq) {
    for (int kl = 0; kl < 32; ++kl) {
      const int k = kq * 32 + kl;
      if (k >= n_vec) break;
      const float v = __shfl_sync(NJODE_FULL, in.v[kq], kl);
#pragma unroll
      for (int q = 0; q < NJ; ++q) {
        const int j = lane + 32 * q;
        if (j < n_out) out.v[q] = fmaf(Wt[(int64_t)k * n_out + j], v, out.v[q]);
      }
    }
  }
  for (int e = 0; e < n_ext; ++e) {
    const float v = ext[e];
#pragma unroll
    for (int q = 0; q < NJ; ++q) {
      const int j = lane + 32 * q;
      if (j < n_out) out.v[q] = fmaf(Wt[(int64_t)(n_vec + e) * n_out + j], v, out.v[q]);
    }
  }
  if (T.act[net][l]) {
#pragma unroll
    for (int q = 0; q < NJ; ++q) { const int j = lane + 32 * q; out.v[q] = j < n_out ? act_fwd_rt(act_kind, out.v[q]) : 0.0f; }
  }
}

// whole net forward, keeping every layer output z[0..L]
template <int NJ>
__device__ __forceinline__ void net_fwd(const ParamTable& T, int net, const float* pt, const float* p,
                                        const Vec<NJ>& in, const float* ext, int act_kind, Vec<NJ>* z) {
  for (int l = 0; l <= T.L; ++l) layer_fwd<NJ>(T, net, l, pt, p, l == 0 ? in : z[l - 1], l == 0 ? ext : nullptr, act_kind, z[l]);
}

// din = W^T d over the hidden-vector columns;  W is n_out x (n_vec+n_ext) row-major (PyTorch layout)
template <int NJ>
__device__ __forceinline__ void layer_dgrad(const ParamTable& T, int net, int l, const float* __restrict__ p,
                                            const Vec<NJ>& d, Vec<NJ>& din) {
  const int lane = threadIdx.x & 31;
  const int n_vec = T.n_vec[net][l], n_out = T.n_out[net][l];
  const int ld = n_vec + T.n_ext[net][l];
  const float* __restrict__ W = p + T.w_off[net][l];
#pragma unroll
  for (int q = 0; q < NJ; ++q) din.v[q] = 0.0f;
#pragma unroll
  for (int jq = 0; jq < NJ; ++jq) {
    for (int jl = 0; jl < 32; ++jl) {
      const int j = jq * 32 + jl;
      if (j >= n_out) break;
      const float dv = __shfl_sync(NJODE_FULL, d.v[jq], jl);
#pragma unroll
      for (int q = 0; q < NJ; ++q) {
        const int k = lane + 32 * q;
        if (k < n_vec) din.v[q] = fmaf(W[(int64_t)j * ld + k], dv, din.v[q]);
      }
    }
  }
}

// partial (input-major) += [in ; ext] (outer) d ;  bias partial += d
template <int NJ>
__device__ __forceinline__ void layer_wgrad(const ParamTable& T, int net, int l, float* __restrict__ part,
                                            const Vec<NJ>& in, const float* ext, const Vec<NJ>& d) {
  const int lane = threadIdx.x & 31;
  const int n_vec = T.n_vec[net][l], n_ext = T.n_ext[net][l], n_out = T.n_out[net][l];
  float* __restrict__ gWt = part + T.w_off[net][l];
  float* __restrict__ gb = part + T.b_off[net][l];
#pragma unroll
  for (int kq = 0; kq < NJ; ++kq) {
    for (int kl = 0; kl < 32; ++kl) {
      const int k = kq * 32 + kl;
      if (k >= n_vec) break;
      const float v = __shfl_sync(NJODE_FULL, in.v[kq], kl);
#pragma unroll
      for (int q = 0; q < NJ; ++q) {
        const int j = lane + 32 * q;
        if (j < n_out) gWt[(int64_t)k * n_out + j] += v * d.v[q];
      }
    }
  }
  for (int e = 0; e < n_ext; ++e) {
    const float v = ext[e];
#pragma unroll
    for (int q = 0; q < NJ; ++q) {
      const int j = lane + 32 * q;
      if (j < n_out) gWt[(int64_t)(n_vec + e) * n_out + j] += v * d.v[q];
    }
  }
#pragma unroll
  for (int q = 0; q < NJ; ++q) { const int j = lane + 32 * q; if (j < n_out) gb[j] += d.v[q]; }
}

// reverse of net_fwd: d is the gradient w.r.t. the net output on entry, w.r.t. the hidden-vector
// input on exit (left untouched garbage when the first layer has no vector input).
template <int NJ>
__device__ __forceinline__ void net_bwd(const ParamTable& T, int net, const float* p, float* part,
                                        const Vec<NJ>& in, const float* ext, int act_kind,
                                        const Vec<NJ>* z, Vec<NJ>& d) {
  for (int l = T.L; l >= 0; --l) {
    if (T.act[net][l]) {
#pragma unroll
      for (int q = 0; q < NJ; ++q) d.v[q] *= act_grad_rt(act_kind, z[l].v[q]);
    }
    layer_wgrad<NJ>(T, net, l, part, l == 0 ? in : z[l - 1], l == 0 ? ext : nullptr, d);
    if (T.n_vec[net][l] > 0) {
      Vec<NJ> din;
      layer_dgrad<NJ>(T, net, l, p, d, din);
      d = din;
    }
  }
}

template <int NJ>
__device__ __forceinline__ void load_row(const float* __restrict__ src, int H, Vec<NJ>& h) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int q = 0; q < NJ; ++q) { const int j = lane + 32 * q; h.v[q] = j < H ? src[j] : 0.0f; }
}
template <int NJ>
__device__ __forceinline__ void store_row(float* __restrict__ dst, int H, const Vec<NJ>& h) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int q = 0; q < NJ; ++q) { const int j = lane + 32 * q; if (j < H) dst[j] = h.v[q]; }
}

// write / read the O readouts of one stack into the (N, d_y, M) prediction tensors
__device__ __forceinline__ int64_t pred_index(const ParamTable& T, int64_t obs, int s, int o) {
  // separate nets: o = d, moment = s (jump_ode.py:179);  shared: flat o = d*M + m (jump_ode.py:172)
  return T.S == 1 ? obs * T.d_y * T.M + o : (obs * T.d_y + o) * T.M + s;
}

template <int NJ>
__global__ void __launch_bounds__(32) k_generic_forward(SweepArgs a) {
  const ParamTable& T = a.T;
  const int lane = threadIdx.x;
  const int s = blockIdx.x % T.S;
  const int worker = blockIdx.x / T.S, n_workers = gridDim.x / T.S;
  const float* p = a.params + (int64_t)s * T.stack_floats;
  const float* pt = a.params_t + (int64_t)s * T.stack_floats;
  const int H = T.H, R = a.tile_rows;
  const int act_kind = a.desc.activation, sc_kind = a.desc.input_scaling;
  float* ckpt = a.ckpt ? a.ckpt + (int64_t)s * a.total_slots * R * H : nullptr;

  Vec<NJ> z[NJODE_LMAX + 1];
  float ext[GEN_EXT_MAX];
  for (int64_t tile = worker; tile < a.n_tiles; tile += n_workers) {
    const int64_t slot0 = a.tile_slot_off[tile];
    for (int r = 0; r < R; ++r) {
      const int u = a.perm[tile * R + r];
      if (u < 0) continue;
      const int ke = a.kenc[u], K = ke >> 1;
      for (int e = 0; e < T.d_x; ++e) ext[e] = a.values[(int64_t)u * T.d_x + e];
      // h = jump(x_u)                                             jump_ode.py:169 / :176
      Vec<NJ> none;
#pragma unroll
      for (int q = 0; q < NJ; ++q) none.v[q] = 0.0f;
      net_fwd<NJ>(T, NET_JUMP, pt, p, none, ext, act_kind, z);
      Vec<NJ> h = z[T.L];
      if (ckpt) store_row<NJ>(ckpt + ((slot0 + 0) * R + r) * H, H, h);
      // preds[u] = out(h)                                          jump_ode.py:170 / :177
      net_fwd<NJ>(T, NET_OUT, pt, p, h, nullptr, act_kind, z);
      if (lane < T.O) a.preds[pred_index(T, u, s, lane)] = z[T.L].v[0];
      if (!(ke & 1)) continue;
      // integrate to t_{u+1} holding x_u constant                  jump_ode.py:188-203
      for (int e = 0; e < T.d_x; ++e) ext[e] = scale_fwd_rt(sc_kind, ext[e]);
      for (int k = 0; k < K; ++k) {
        const float tc = a.knots[(slot0 + k) * R + r], tn = a.knots[(slot0 + k + 1) * R + r];
        const float delta = __fsub_rn(tn, tc);                      // (t_next - t_last), jump_ode.py:138
        ext[T.d_x] = tc;
        ext[T.d_x + 1] = delta;
        Vec<NJ> sh;
#pragma unroll
        for (int q = 0; q < NJ; ++q) sh.v[q] = scale_fwd_rt(sc_kind, h.v[q]);
        net_fwd<NJ>(T, NET_ODE, pt, p, sh, ext, act_kind, z);
#pragma unroll
        for (int q = 0; q < NJ; ++q) h.v[q] = fmaf(delta, z[T.L].v[q], h.v[q]);   // jump_ode.py:139
        if (ckpt) store_row<NJ>(ckpt + ((slot0 + k + 1) * R + r) * H, H, h);
      }
      // preds_before[u+1] = out(h_end)                             jump_ode.py:205-212
      net_fwd<NJ>(T, NET_OUT, pt, p, h, nullptr, act_kind, z);
      if (lane < T.O) a.preds_before[pred_index(T, (int64_t)u + 1, s, lane)] = z[T.L].v[0];
    }
  }
}

template <int NJ>
__global__ void __launch_bounds__(32) k_generic_backward(SweepArgs a) {
  const ParamTable& T = a.T;
  const int lane = threadIdx.x;
  const int s = blockIdx.x % T.S;
  const int worker = blockIdx.x / T.S, n_workers = gridDim.x / T.S;
  const float* p = a.params + (int64_t)s * T.stack_floats;
  const float* pt = a.params_t + (int64_t)s * T.stack_floats;
  float* part = a.partials + (int64_t)blockIdx.x * T.stack_floats;
  const int H = T.H, R = a.tile_rows;
  const int act_kind = a.desc.activation, sc_kind = a.desc.input_scaling;
  const float* ckpt = a.ckpt + (int64_t)s * a.total_slots * R * H;

  Vec<NJ> z[NJODE_LMAX + 1];
  float ext[GEN_EXT_MAX];
  for (int64_t tile = worker; tile < a.n_tiles; tile += n_workers) {
    const int64_t slot0 = a.tile_slot_off[tile];
    for (int r = 0; r < R; ++r) {
      const int u = a.perm[tile * R + r];
      if (u < 0) continue;
      const int ke = a.kenc[u], K = ke >> 1;
      Vec<NJ> g;                                   // dLoss/dh, walked backwards in time
#pragma unroll
      for (int q = 0; q < NJ; ++q) g.v[q] = 0.0f;
      Vec<NJ> h;
      if (ke & 1) {
        // readout at the end of the interval -> preds_before[u+1]
        load_row<NJ>(ckpt + ((slot0 + K) * R + r) * H, H, h);
        net_fwd<NJ>(T, NET_OUT, pt, p, h, nullptr, act_kind, z);
        Vec<NJ> d;
#pragma unroll
        for (int q = 0; q < NJ; ++q) d.v[q] = 0.0f;
        if (lane < T.O) d.v[0] = a.grad_preds_before[pred_index(T, (int64_t)u + 1, s, lane)];
        net_bwd<NJ>(T, NET_OUT, p, part, h, nullptr, act_kind, z, d);
        g = d;
        for (int e = 0; e < T.d_x; ++e) ext[e] = scale_fwd_rt(sc_kind, a.values[(int64_t)u * T.d_x + e]);
        for (int k = K - 1; k >= 0; --k) {
          const float tc = a.knots[(slot0 + k) * R + r], tn = a.knots[(slot0 + k + 1) * R + r];
          const float delta = __fsub_rn(tn, tc);
          ext[T.d_x] = tc;
          ext[T.d_x + 1] = delta;
          load_row<NJ>(ckpt + ((slot0 + k) * R + r) * H, H, h);
          Vec<NJ> sh;
#pragma unroll
          for (int q = 0; q < NJ; ++q) sh.v[q] = scale_fwd_rt(sc_kind, h.v[q]);
          net_fwd<NJ>(T, NET_ODE, pt, p, sh, ext, act_kind, z);
          // h' = h + delta * f(h):  d f = delta * g ;  g <- g + s'(h) * (W0h^T ...)
          Vec<NJ> d2;
#pragma unroll
          for (int q = 0; q < NJ; ++q) d2.v[q] = delta * g.v[q];
          net_bwd<NJ>(T, NET_ODE, p, part, sh, ext, act_kind, z, d2);
#pragma unroll
          for (int q = 0; q < NJ; ++q) g.v[q] = fmaf(d2.v[q], scale_grad_rt(sc_kind, sh.v[q]), g.v[q]);
        }
      }
      // readout right after the jump -> preds[u]
      load_row<NJ>(ckpt + ((slot0 + 0) * R + r) * H, H, h);
      net_fwd<NJ>(T, NET_OUT, pt, p, h, nullptr, act_kind, z);
      {
        Vec<NJ> d;
#pragma unroll
        for (int q = 0; q < NJ; ++q) d.v[q] = 0.0f;
        if (lane < T.O) d.v[0] = a.grad_preds[pred_index(T, u, s, lane)];
        net_bwd<NJ>(T, NET_OUT, p, part, h, nullptr, act_kind, z, d);
#pragma unroll
        for (int q = 0; q < NJ; ++q) g.v[q] += d.v[q];
      }
      // jump net: h0 = jump(x_u)
      for (int e = 0; e < T.d_x; ++e) ext[e] = a.values[(int64_t)u * T.d_x + e];
      Vec<NJ> none;
#pragma unroll
      for (int q = 0; q < NJ; ++q) none.v[q] = 0.0f;
      net_fwd<NJ>(T, NET_JUMP, pt, p, none, ext, act_kind, z);
      net_bwd<NJ>(T, NET_JUMP, p, part, none, ext, act_kind, z, g);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Dense-grid inference: the model's prediction at EVERY time of a grid, not only at observation
// times -- what the reference's plotting code computes by driving jump_nns / euler_step / output_nns from Python
// (utils/plotting.py:133-256).  Its step rule differs from training: from the current time to each grid time t it takes
//     n_sub = max(1, int((t - t_cur) / dt_ode_step))     (1 if dt_ode_step is None)           plotting.py:166-169
// equal Euler sub-steps of (t - t_cur) / n_sub, all in float32 (t_cur accumulates), holding x_i constant; after a step
// sequence it evaluates the output net.  Grid times in [T_i, T_{i+1}] belong to observation i; the value at T_{i+1} is
// then overwritten by observation i+1 (after its jump) -- except when i+1 is the LAST observation, whose own loop only
// covers times > T_last (plotting.py:210): the grid point at the last observation keeps the pre-jump value.  Grid
// times before the first observation stay 0.  One warp per (observation, stack); raw readouts (mean, W) are written,
// the variance transform (W^2, or clamp(W - mean^2, 0)) is left to the caller (plotting.py:189-196).
template <int NJ>
__global__ void __launch_bounds__(32) k_generic_dense(NjodeDesc desc, ParamTable T, const float* __restrict__ params,
                                                      const float* __restrict__ params_t, const float* __restrict__ times,
                                                      const float* __restrict__ values, const int64_t* __restrict__ off, int64_t B,
                                                      int64_t N, const float* __restrict__ grid, int64_t G, float* __restrict__ dense) {
  const int lane = threadIdx.x;
  const int s = blockIdx.x % T.S;
  const int64_t u = blockIdx.x / T.S;
  if (u >= N) return;
  const float* p = params + (int64_t)s * T.stack_floats;
  const float* pt = params_t + (int64_t)s * T.stack_floats;
  const int act_kind = desc.activation, sc_kind = desc.input_scaling;
  // trajectory of this observation: last b with off[b] <= u
  int64_t lo = 0, hi = B;
  while (hi - lo > 1) { const int64_t mid = (lo + hi) >> 1; if (off[mid] <= u) lo = mid; else hi = mid; }
  const int64_t b = lo;
  const bool last = (u + 1 == off[b + 1]);
  const bool next_is_last = !last && (u + 2 == off[b + 1]);
  const float Ti = times[u], Tn = last ? 0.0f : times[u + 1];
  // first grid index of this observation: grid >= T_i (last observation: grid > T_last)
  int64_t g0 = 0, g1 = G;
  while (g0 < g1) { const int64_t mid = (g0 + g1) >> 1; const float t = grid[mid]; if (last ? (t <= Ti) : (t < Ti)) g0 = mid + 1; else g1 = mid; }

  Vec<NJ> z[NJODE_LMAX + 1];
  float ext[GEN_EXT_MAX];
  for (int e = 0; e < T.d_x; ++e) ext[e] = values[u * T.d_x + e];
  Vec<NJ> none;
#pragma unroll
  for (int q = 0; q < NJ; ++q) none.v[q] = 0.0f;
  net_fwd<NJ>(T, NET_JUMP, pt, p, none, ext, act_kind, z);
  Vec<NJ> h = z[T.L];
  for (int e = 0; e < T.d_x; ++e) ext[e] = scale_fwd_rt(sc_kind, ext[e]);
  const float dtf = desc.dt;
  float t_cur = Ti;
  for (int64_t g = g0; g < G; ++g) {
    const float tt = grid[g];
    if (!last && tt > Tn) break;
    const float span = __fsub_rn(tt, t_cur);
    int n_sub = 1;
    if (desc.has_dt) { n_sub = (int)__fdiv_rn(span, dtf); if (n_sub < 1) n_sub = 1; }
    const float dts = __fdiv_rn(span, (float)n_sub);
    for (int i = 0; i < n_sub; ++i) {
      const float t_new = __fadd_rn(t_cur, dts);
      const float delta = __fsub_rn(t_new, t_cur);
      ext[T.d_x] = t_cur;
      ext[T.d_x + 1] = delta;
      Vec<NJ> sh;
#pragma unroll
      for (int q = 0; q < NJ; ++q) sh.v[q] = scale_fwd_rt(sc_kind, h.v[q]);
      net_fwd<NJ>(T, NET_ODE, pt, p, sh, ext, act_kind, z);
#pragma unroll
      for (int q = 0; q < NJ; ++q) h.v[q] = fmaf(delta, z[T.L].v[q], h.v[q]);
      t_cur = t_new;
    }
    if (!last && tt == Tn && !next_is_last) continue;          // the next observation writes this grid point (after its jump)
    net_fwd<NJ>(T, NET_OUT, pt, p, h, nullptr, act_kind, z);
    if (lane < T.O) dense[pred_index(T, b * G + g, s, lane)] = z[T.L].v[0];
  }
}

int njode_generic_dense(const NjodeDesc* d, const float* params, const float* params_t, const float* times, const float* values,
                        const int64_t* off, int64_t B, int64_t N, const float* grid, int64_t G, float* dense, cudaStream_t st) {
  const ParamTable T = njode_make_table(d);
  if (N == 0 || G == 0) return NJODE_OK;
  const int nj = (T.H + 31) / 32;
  const unsigned blocks = (unsigned)(N * T.S);
#define NJODE_DENSE(NJ_) k_generic_dense<NJ_><<<blocks, 32, 0, st>>>(*d, T, params, params_t, times, values, off, B, N, grid, G, dense)
  if (nj <= 1) NJODE_DENSE(1); else if (nj <= 2) NJODE_DENSE(2); else if (nj <= 4) NJODE_DENSE(4); else NJODE_DENSE(8);
#undef NJODE_DENSE
  NJODE_LAUNCH_OK("k_generic_dense");
  return NJODE_OK;
}

// ------------------------------------------------------------------------------------------------
int njode_generic_supported(const NjodeDesc* d, const char** why) {
  if (d->hidden > 32 * GEN_NJ_MAX) { *why = "generic kernels support hidden_dim <= 256"; return 0; }
  if (d->d_x + 2 > GEN_EXT_MAX) { *why = "generic kernels support input_dim <= 8"; return 0; }
  const int O = d->shared_network ? d->d_y * d->num_moments : d->d_y;
  if (O > 32) { *why = "generic kernels support output_dim*num_moments <= 32"; return 0; }
  return 1;
}

int njode_generic_workers(const NjodeDesc* d, int64_t n_tiles) {
  const int S = d->shared_network ? 1 : d->num_moments;
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int64_t per_stack = (int64_t)sms * GEN_WARPS_PER_SM / S;
  if (per_stack < 1) per_stack = 1;
  if (per_stack > n_tiles) per_stack = n_tiles > 0 ? n_tiles : 1;
  // bound the partial-sum buffer to ~1 GiB
  const ParamTable T = njode_make_table(d);
  const int64_t cap = (1ll << 30) / ((int64_t)T.stack_floats * 4 * S);
  if (cap >= 1 && per_stack > cap) per_stack = cap;
  return (int)(per_stack * S);
}

template <int NJ>
static int launch_generic(const SweepArgs& a, cudaStream_t st, bool backward) {
  if (a.n_tiles == 0) return NJODE_OK;
  njode_timing_begin(backward ? 2 : 1, st);
  if (backward) k_generic_backward<NJ><<<a.n_workers, 32, 0, st>>>(a);
  else k_generic_forward<NJ><<<a.n_workers, 32, 0, st>>>(a);
  njode_timing_end(backward ? 2 : 1, st);
  NJODE_LAUNCH_OK(backward ? "k_generic_backward" : "k_generic_forward");
  return NJODE_OK;
}

static int dispatch_generic(const SweepArgs& a, cudaStream_t st, bool backward) {
  const int nj = (a.T.H + 31) / 32;
  if (nj <= 1) return launch_generic<1>(a, st, backward);
  if (nj <= 2) return launch_generic<2>(a, st, backward);
  if (nj <= 4) return launch_generic<4>(a, st, backward);
  return launch_generic<8>(a, st, backward);
}

int njode_generic_forward(const SweepArgs& a, cudaStream_t st) { return dispatch_generic(a, st, false); }
int njode_generic_backward(const SweepArgs& a, cudaStream_t st) { return dispatch_generic(a, st, true); }
