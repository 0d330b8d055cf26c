// njode_rowtile.cu -- row-tiled FP32 sweep kernels for hidden_dim in {32, 64, 96, 128}, up to 3 hidden layers
// (BASELINE configs 4 and 5: hidden 128 / 3 layers / tanh and hidden 64).
//
// A CTA (256 threads) owns a tile of 32 observation units of one network stack and runs ALL of them through
// every Linear layer together: a layer is a [32 x K] x [K x N] product with the layer's weights staged in
// shared memory and a 4-row x (H/32)-column register tile per thread,
//     thread (ty, tx):  rows 4*ty .. 4*ty+3,  columns tx, tx+32, ...      (ty = warp, tx = lane)
// Activations live in shared memory TRANSPOSED ([feature][36]: 32 rows + 4 floats of padding), so the four rows
// of a thread are one warp-uniform LDS.128 per k and a thread's results go back with conflict-free STS.128
// (144-byte feature stride = 9 x 16 B); weight reads are one conflict-free LDS.32 per column.  The per-row
// scalars of the first layers (x, t, dt) are extra "feature" rows of the same buffer, so every layer is the
// same loop.  The reverse sweep re-computes the hidden layers of a step from the checkpointed state, then walks
// the layers backwards: weight gradients are 4x4 register tiles of length-32 dot products between transposed
// buffers (contraction over the rows), added to this CTA's private partial buffer (input-major layout, plain
// read-modify-write by one owner thread: no atomics, deterministic for a given schedule); data gradients are
// the same GEMM loop against the PyTorch-layout weights.  FP32 FMA throughout -- the same arithmetic as the
// reference up to summation order.  (The tcgen05 path of njode_tiled.cu generalises to these shapes; this kernel
// is the first fast path for them: ~100x the warp-per-unit generic kernels.)
#include "njode_common.cuh"

namespace {

constexpr int RT = NJODE_GENERIC_TILE_ROWS;   // 32 rows per tile
constexpr int LDA = 36;                       // floats per feature row of a transposed activation buffer
constexpr int NTH = 256;
constexpr int EXT_MAX = 10;                   // d_x + 2 <= 10
constexpr int LMAX_RT = 3;                    // hidden layers
constexpr int O_MAX = 32;

__device__ __forceinline__ int64_t pred_index_rt(const ParamTable& T, int64_t obs, int s, int o) {
  return T.S == 1 ? obs * T.d_y * T.M + o : (obs * T.d_y + o) * T.M + s;
}

template <int NB>
struct Tile {                      // a thread's 4 x NB elements: rows 4*ty + i, columns tx + 32*j
  float v[4][NB];
};

// shared-memory carve-up (floats)
template <int NB>
struct Layout {
  static constexpr int H = 32 * NB;
  static constexpr int ACT_F = (H + EXT_MAX + 2) * LDA;       // one transposed activation buffer (+ ext rows)
  static constexpr int W_F = (H + EXT_MAX) * H;               // one layer's weights
  static constexpr int META_F = 6 * RT;                       // per-row scalars
  static constexpr size_t fwd_bytes() { return (size_t)(3 * ACT_F + W_F + META_F) * 4; }
  static constexpr size_t bwd_bytes() { return (size_t)((2 + LMAX_RT) * ACT_F + W_F + META_F) * 4; }
};

// stage `count` floats of global memory in shared memory (all threads; caller synchronises)
__device__ __forceinline__ void stage(float* __restrict__ dst, const float* __restrict__ src, int count) {
  int i = threadIdx.x;
  for (; i + 7 * NTH < count; i += 8 * NTH) {
    float t[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) t[u] = src[i + u * NTH];
#pragma unroll
    for (int u = 0; u < 8; ++u) dst[i + u * NTH] = t[u];
  }
  for (; i < count; i += NTH) dst[i] = src[i];
}

// 16 bytes from SHARED memory (explicit: through these pointers the compiler emitted generic LD.E.128)
__device__ __forceinline__ float4 lds128(const float* p) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"((uint32_t)__cvta_generic_to_shared(p)));
  return v;
}

__device__ __forceinline__ void sts128(float* p, float4 v) {
  asm volatile("st.shared.v4.f32 [%4], {%0,%1,%2,%3};" :: "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "r"((uint32_t)__cvta_generic_to_shared(p)) : "memory");
}

// acc[i][j] = sum_k in_T[k][4*ty + i] * W[k * ldw + tx + 32*j]   for columns < N   (all 256 threads)
template <int NB>
__device__ __forceinline__ void gemm(const float* __restrict__ in_T, int K, const float* __restrict__ W, int ldw, int N,
                                     Tile<NB>& acc) {
  const int ty = threadIdx.x >> 5, tx = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < NB; ++j) acc.v[i][j] = 0.0f;
  const float* a_ptr = in_T + 4 * ty;
  const bool full = N >= 32 * NB;
  if (full) {
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
      const float4 a = lds128(a_ptr + k * LDA);
      const float* w = W + k * ldw + tx;
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        const float wv = w[32 * j];
        acc.v[0][j] = fmaf(a.x, wv, acc.v[0][j]);
        acc.v[1][j] = fmaf(a.y, wv, acc.v[1][j]);
        acc.v[2][j] = fmaf(a.z, wv, acc.v[2][j]);
        acc.v[3][j] = fmaf(a.w, wv, acc.v[3][j]);
      }
    }
  } else {
    for (int k = 0; k < K; ++k) {
      const float4 a = lds128(a_ptr + k * LDA);
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        const int n = tx + 32 * j;
        const float wv = n < N ? W[k * ldw + n] : 0.0f;
        acc.v[0][j] = fmaf(a.x, wv, acc.v[0][j]);
        acc.v[1][j] = fmaf(a.y, wv, acc.v[1][j]);
        acc.v[2][j] = fmaf(a.z, wv, acc.v[2][j]);
        acc.v[3][j] = fmaf(a.w, wv, acc.v[3][j]);
      }
    }
  }
}

// a thread's elements <-> transposed buffer (columns < N)
template <int NB>
__device__ __forceinline__ void put_tile(float* __restrict__ buf, const Tile<NB>& t, int N) {
  const int ty = threadIdx.x >> 5, tx = threadIdx.x & 31;
#pragma unroll
  for (int j = 0; j < NB; ++j) {
    const int n = tx + 32 * j;
    if (n < N) sts128(buf + n * LDA + 4 * ty, make_float4(t.v[0][j], t.v[1][j], t.v[2][j], t.v[3][j]));
  }
}
template <int NB>
__device__ __forceinline__ void get_tile(const float* __restrict__ buf, Tile<NB>& t, int N) {
  const int ty = threadIdx.x >> 5, tx = threadIdx.x & 31;
#pragma unroll
  for (int j = 0; j < NB; ++j) {
    const int n = tx + 32 * j;
    float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n < N) q = lds128(buf + n * LDA + 4 * ty);
    t.v[0][j] = q.x; t.v[1][j] = q.y; t.v[2][j] = q.z; t.v[3][j] = q.w;
  }
}

// one Linear of net `net`: out = [act](b + [in ; ext] W^T) for the tile; input rows [row0, row0 + K) of `in_T`
template <int NB>
__device__ __forceinline__ void layer_forward(const ParamTable& T, int net, int l, const float* __restrict__ p,
                                              const float* __restrict__ pt, const float* __restrict__ in_T, float* wbuf,
                                              int act_kind, Tile<NB>& out) {
  const int n_vec = T.n_vec[net][l], n_ext = T.n_ext[net][l], n_out = T.n_out[net][l];
  const int K = n_vec + n_ext;
  __syncthreads();                                   // previous users of wbuf / producers of in_T are done
  stage(wbuf, pt + T.w_off[net][l], K * n_out);      // input-major: [k][n_out]
  __syncthreads();
  // the ext rows sit right after the H feature rows: a layer without vector input starts there
  gemm<NB>(in_T + (n_vec == 0 ? 32 * NB * LDA : 0), K, wbuf, n_out, n_out, out);
  const int tx = threadIdx.x & 31;
  const float* __restrict__ b = p + T.b_off[net][l];
#pragma unroll
  for (int j = 0; j < NB; ++j) {
    const int n = tx + 32 * j;
    const float bv = n < n_out ? b[n] : 0.0f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float v = out.v[i][j] + bv;
      if (T.act[net][l]) v = act_fwd_rt(act_kind, v);
      out.v[i][j] = n < n_out ? v : 0.0f;
    }
  }
}

// whole net forward; hidden-layer outputs z_0 .. z_{L-1} go to zbuf[l] (transposed), the last layer's output is
// returned in registers.  `in_T` holds the vector input in rows [0,H) and the ext scalars in rows [H, H+n_ext).
// a thread's elements <-> a row-major [32][H] plane in global memory (coalesced: lanes = consecutive columns)
template <int NB>
__device__ __forceinline__ void store_plane(float* __restrict__ dst, const Tile<NB>& t) {
  const int ty = threadIdx.x >> 5, tx = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < NB; ++j) dst[(4 * ty + i) * (32 * NB) + tx + 32 * j] = t.v[i][j];
}
template <int NB>
__device__ __forceinline__ void load_plane(const float* __restrict__ src, Tile<NB>& t) {
  const int ty = threadIdx.x >> 5, tx = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < NB; ++j) t.v[i][j] = src[(4 * ty + i) * (32 * NB) + tx + 32 * j];
}

template <int NB>
__device__ __forceinline__ void net_forward(const ParamTable& T, int net, const float* p, const float* pt,
                                            const float* in_T, float* const* zbuf, float* wbuf, int act_kind, Tile<NB>& out,
                                            float* __restrict__ zck = nullptr) {
  for (int l = 0; l <= T.L; ++l) {
    layer_forward<NB>(T, net, l, p, pt, l == 0 ? in_T : zbuf[l - 1], wbuf, act_kind, out);
    if (l < T.L) {
      put_tile<NB>(zbuf[l], out, T.n_out[net][l]);                // (the next layer_forward synchronises)
      if (zck) store_plane<NB>(zck + (int64_t)l * RT * 32 * NB, out);   // checkpoint plane 1 + l of this slot
    }
  }
}

// weight / bias gradient of layer (net, l): part (input-major) += in^T d over the 32 rows of the tile.
// d_T: [n_out][LDA] transposed gradient; in_T: rows [row0, row0+K) transposed input (vector + ext rows).
__device__ __forceinline__ void layer_wgrad(const ParamTable& T, int net, int l, float* __restrict__ part,
                                            const float* __restrict__ in_T, const float* __restrict__ d_T, int H) {
  const int n_vec = T.n_vec[net][l], n_ext = T.n_ext[net][l], n_out = T.n_out[net][l];
  const int K = n_vec + n_ext;
  const float* __restrict__ inp = in_T + (n_vec == 0 ? H * LDA : 0);
  float* __restrict__ gWt = part + T.w_off[net][l];
  const int JQ = (n_out + 3) >> 2;                   // a thread owns columns jq, jq + JQ, jq + 2 JQ, jq + 3 JQ
  const int KQ = (K + 3) >> 2;                       // ... and input rows 4 kq .. 4 kq + 3
  for (int t = threadIdx.x; t < JQ * KQ; t += NTH) {
    const int jq = t % JQ, kq = t / JQ;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int m = 0; m < 4; ++m) acc[i][m] = 0.0f;
    const float* dj[4];
    const float* ik[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { const int j = jq + JQ * i; dj[i] = d_T + (j < n_out ? j : 0) * LDA; }
#pragma unroll
    for (int m = 0; m < 4; ++m) { const int k = 4 * kq + m; ik[m] = inp + (k < K ? k : 0) * LDA; }
#pragma unroll 2
    for (int c = 0; c < RT; c += 4) {
      float4 dv[4], iv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) dv[i] = lds128(dj[i] + c);
#pragma unroll
      for (int m = 0; m < 4; ++m) iv[m] = lds128(ik[m] + c);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          acc[i][m] = fmaf(dv[i].x, iv[m].x, acc[i][m]);
          acc[i][m] = fmaf(dv[i].y, iv[m].y, acc[i][m]);
          acc[i][m] = fmaf(dv[i].z, iv[m].z, acc[i][m]);
          acc[i][m] = fmaf(dv[i].w, iv[m].w, acc[i][m]);
        }
    }
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const int k = 4 * kq + m;
      if (k >= K) continue;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int j = jq + JQ * i;
        // fire-and-forget reduction (RED.ADD): this thread is the only writer of the element in this CTA's private
        // partial buffer, so the order of the adds is its program order -- deterministic, and no load latency
        if (j < n_out) atomicAdd(&gWt[(int64_t)k * n_out + j], acc[i][m]);
      }
    }
  }
  // bias: column sums of d
  float* __restrict__ gb = part + T.b_off[net][l];
  for (int j = threadIdx.x; j < n_out; j += NTH) {
    const float* dr = d_T + j * LDA;
    float s = 0.0f;
#pragma unroll
    for (int c = 0; c < RT; c += 4) { const float4 q = lds128(dr + c); s += (q.x + q.y) + (q.z + q.w); }
    atomicAdd(&gb[j], s);
  }
}

// reverse of net_forward.  On entry `d` = gradient w.r.t. the net output (registers); on exit = gradient w.r.t.
// the vector input (undefined when the first layer has none).  zL_T: transposed post-activation output of the
// last layer (needed only when that layer has an activation: the jump net), may be null otherwise.
template <int NB>
__device__ __forceinline__ void net_backward(const ParamTable& T, int net, const float* p, float* part,
                                             const float* in_T, float* const* zbuf, const float* zL_T, float* dbuf,
                                             float* wbuf, int act_kind, Tile<NB>& d) {
  constexpr int H = 32 * NB;
  const int tx = threadIdx.x & 31;
  for (int l = T.L; l >= 0; --l) {
    const int n_vec = T.n_vec[net][l], n_ext = T.n_ext[net][l], n_out = T.n_out[net][l];
    if (T.act[net][l]) {
      Tile<NB> z;
      get_tile<NB>(l == T.L ? zL_T : zbuf[l], z, n_out);
#pragma unroll
      for (int j = 0; j < NB; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) d.v[i][j] *= act_grad_rt(act_kind, z.v[i][j]);
    }
    __syncthreads();                                  // previous readers of dbuf / wbuf are done
    put_tile<NB>(dbuf, d, n_out);
    if (n_vec > 0) stage(wbuf, p + T.w_off[net][l], n_out * (n_vec + n_ext));   // PyTorch layout [j][ld]
    __syncthreads();
    layer_wgrad(T, net, l, part, l == 0 ? in_T : zbuf[l - 1], dbuf, H);
    if (n_vec > 0) {
      gemm<NB>(dbuf, n_out, wbuf, n_vec + n_ext, n_vec, d);      // d_in[r][k] = sum_j d[r][j] W[j][k]
#pragma unroll
      for (int j = 0; j < NB; ++j)
        if (tx + 32 * j >= n_vec) {
#pragma unroll
          for (int i = 0; i < 4; ++i) d.v[i][j] = 0.0f;
        }
    }
  }
}

template <int NB>
__device__ __forceinline__ void scale_tile(int sc, Tile<NB>& t) {
  if (sc == NJODE_SCALE_IDENTITY) return;
#pragma unroll
  for (int j = 0; j < NB; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) t.v[i][j] = scale_fwd_rt(sc, t.v[i][j]);
}

// ------------------------------------------------------------------------------------------------
// forward sweep
// ------------------------------------------------------------------------------------------------
template <int NB>
__global__ void __launch_bounds__(NTH) k_rowtile_forward(SweepArgs a) {
  using L_ = Layout<NB>;
  constexpr int H = 32 * NB;
  extern __shared__ float smem_f[];
  float* hbuf = smem_f;                           // state / layer input (+ ext rows)
  float* zb[LMAX_RT] = {smem_f + L_::ACT_F, smem_f + 2 * L_::ACT_F, smem_f + L_::ACT_F};   // ping-pong (only z_{l-1} is live)
  float* wbuf = smem_f + 3 * L_::ACT_F;
  int* m_u = reinterpret_cast<int*>(wbuf + L_::W_F);
  int* m_ke = m_u + RT;
  float* m_delta = reinterpret_cast<float*>(m_ke + RT);

  const ParamTable& T = a.T;
  const int tid = threadIdx.x, ty = tid >> 5, tx = tid & 31;
  const int s = blockIdx.x % T.S;
  const int worker = blockIdx.x / T.S, n_workers = gridDim.x / T.S;
  const float* p = a.params + (int64_t)s * T.stack_floats;
  const float* pt = a.params_t + (int64_t)s * T.stack_floats;
  const int dx = T.d_x, act_kind = a.desc.activation, sc_kind = a.desc.input_scaling;
  // checkpoints: [stack][slot][plane][row][H]; plane 0 = hidden state before the step of this slot (after the last
  // step for the tile's final slot), planes 1..L = hidden-layer outputs z_0..z_{L-1} of the ODE net in that step
  const int planes = 1 + T.L;
  float* ckpt = a.ckpt ? a.ckpt + (int64_t)s * a.total_slots * planes * RT * H : nullptr;
  float* ext = hbuf + H * LDA;                    // ext rows: [e][row]

  auto store_ckpt = [&](int64_t slot, const Tile<NB>& h) {
    if (ckpt) store_plane<NB>(ckpt + slot * planes * RT * H, h);
  };
  // y = out(h) for the tile (h is in hbuf); rows with write[r] get their O readouts stored
  auto readout = [&](float* __restrict__ dst, bool before) {
    Tile<NB> y;
    net_forward<NB>(T, NET_OUT, p, pt, hbuf, zb, wbuf, act_kind, y);
    if (tx < T.O) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = 4 * ty + i, u = m_u[r];
        if (u >= 0 && (!before || (m_ke[r] & 1))) dst[pred_index_rt(T, (int64_t)u + (before ? 1 : 0), s, tx)] = y.v[i][0];
      }
    }
  };

  for (int64_t tile = worker; tile < a.n_tiles; tile += n_workers) {
    const int64_t slot0 = a.tile_slot_off[tile];
    const int kmax = a.tile_kmax[tile];
    __syncthreads();
    if (tid < RT) {
      const int u = a.perm[tile * RT + tid];
      m_u[tid] = u;
      m_ke[tid] = u >= 0 ? a.kenc[u] : 0;
      for (int e = 0; e < dx; ++e) ext[e * LDA + tid] = u >= 0 ? a.values[(int64_t)u * dx + e] : 0.0f;
    }
    // h = jump(x)                                                    jump_ode.py:169 / :176
    Tile<NB> h;
    net_forward<NB>(T, NET_JUMP, p, pt, hbuf, zb, wbuf, act_kind, h);
    __syncthreads();
    put_tile<NB>(hbuf, h, H);
    store_ckpt(slot0, h);
    // preds[u] = out(h)                                              jump_ode.py:170 / :177
    readout(a.preds, false);
    // Euler steps with x held constant                               jump_ode.py:188-203, :122-140
    __syncthreads();
    if (tid < RT) for (int e = 0; e < dx; ++e) ext[e * LDA + tid] = scale_fwd_rt(sc_kind, ext[e * LDA + tid]);
    for (int k = 0; k < kmax; ++k) {
      __syncthreads();
      if (tid < RT) {
        const float tc = a.knots[(slot0 + k) * RT + tid], tn = a.knots[(slot0 + k + 1) * RT + tid];
        const float delta = __fsub_rn(tn, tc);
        ext[dx * LDA + tid] = tc;
        ext[(dx + 1) * LDA + tid] = delta;
        m_delta[tid] = delta;
      }
      if (sc_kind != NJODE_SCALE_IDENTITY) {        // the ODE net sees s(h) (jump_ode.py:57); h itself stays in registers
        Tile<NB> sh = h;
        scale_tile<NB>(sc_kind, sh);
        put_tile<NB>(hbuf, sh, H);
      }
      Tile<NB> f;
      net_forward<NB>(T, NET_ODE, p, pt, hbuf, zb, wbuf, act_kind, f,
                      ckpt ? ckpt + ((slot0 + k) * planes + 1) * RT * H : nullptr);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = 4 * ty + i;
        const float delta = m_delta[r];
        if (k < (m_ke[r] >> 1)) {
#pragma unroll
          for (int j = 0; j < NB; ++j) h.v[i][j] = fmaf(delta, f.v[i][j], h.v[i][j]);     // jump_ode.py:139
        }
      }
      __syncthreads();
      put_tile<NB>(hbuf, h, H);
      store_ckpt(slot0 + k + 1, h);
    }
    // preds_before[u+1] = out(h_end)                                 jump_ode.py:205-212
    readout(a.preds_before, true);
  }
}

// ------------------------------------------------------------------------------------------------
// reverse sweep
// ------------------------------------------------------------------------------------------------
template <int NB>
__global__ void __launch_bounds__(NTH) k_rowtile_backward(SweepArgs a) {
  using L_ = Layout<NB>;
  constexpr int H = 32 * NB;
  extern __shared__ float smem_f[];
  float* hbuf = smem_f;                           // state / layer input (+ ext rows)
  float* zb[LMAX_RT] = {smem_f + L_::ACT_F, smem_f + 2 * L_::ACT_F, smem_f + 3 * L_::ACT_F};
  float* dbuf = smem_f + (1 + LMAX_RT) * L_::ACT_F;
  float* wbuf = smem_f + (2 + LMAX_RT) * L_::ACT_F;
  int* m_u = reinterpret_cast<int*>(wbuf + L_::W_F);
  int* m_ke = m_u + RT;
  float* m_delta = reinterpret_cast<float*>(m_ke + RT);

  const ParamTable& T = a.T;
  const int tid = threadIdx.x, ty = tid >> 5, tx = tid & 31;
  const int s = blockIdx.x % T.S;
  const int worker = blockIdx.x / T.S, n_workers = gridDim.x / T.S;
  const float* p = a.params + (int64_t)s * T.stack_floats;
  const float* pt = a.params_t + (int64_t)s * T.stack_floats;
  float* part = a.partials + (int64_t)blockIdx.x * T.stack_floats;
  const int dx = T.d_x, act_kind = a.desc.activation, sc_kind = a.desc.input_scaling;
  const int planes = 1 + T.L;                     // see the forward kernel
  const float* ckpt = a.ckpt + (int64_t)s * a.total_slots * planes * RT * H;
  float* ext = hbuf + H * LDA;

  auto load_ckpt = [&](int64_t slot, Tile<NB>& h) { load_plane<NB>(ckpt + slot * planes * RT * H, h); };
  // readout backward at the hidden state in `h` (also stored to hbuf): g += d loss / d h
  auto out_backward = [&](const Tile<NB>& h, const float* __restrict__ gsrc, bool before, Tile<NB>& g) {
    __syncthreads();
    put_tile<NB>(hbuf, h, H);
    Tile<NB> y;
    net_forward<NB>(T, NET_OUT, p, pt, hbuf, zb, wbuf, act_kind, y);
    Tile<NB> d;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
      for (int j = 0; j < NB; ++j) d.v[i][j] = 0.0f;
      const int r = 4 * ty + i, u = m_u[r];
      if (tx < T.O && u >= 0 && (!before || (m_ke[r] & 1))) d.v[i][0] = gsrc[pred_index_rt(T, (int64_t)u + (before ? 1 : 0), s, tx)];
    }
    net_backward<NB>(T, NET_OUT, p, part, hbuf, zb, nullptr, dbuf, wbuf, act_kind, d);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < NB; ++j) g.v[i][j] += d.v[i][j];
  };

  for (int64_t tile = worker; tile < a.n_tiles; tile += n_workers) {
    const int64_t slot0 = a.tile_slot_off[tile];
    const int kmax = a.tile_kmax[tile];
    __syncthreads();
    if (tid < RT) {
      const int u = a.perm[tile * RT + tid];
      m_u[tid] = u;
      m_ke[tid] = u >= 0 ? a.kenc[u] : 0;
    }
    __syncthreads();
    Tile<NB> g, h;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < NB; ++j) g.v[i][j] = 0.0f;

    // ---- preds_before[u+1] = out(h_end): the state after the tile's last step (rows that stopped earlier kept theirs) ----
    load_ckpt(slot0 + kmax, h);
    out_backward(h, a.grad_preds_before, true, g);

    // ---- Euler steps, last to first ----
    __syncthreads();
    if (tid < RT) {
      const int u = m_u[tid];
      for (int e = 0; e < dx; ++e) ext[e * LDA + tid] = u >= 0 ? scale_fwd_rt(sc_kind, a.values[(int64_t)u * dx + e]) : 0.0f;
    }
    for (int k = kmax - 1; k >= 0; --k) {
      __syncthreads();
      if (tid < RT) {
        const float tc = a.knots[(slot0 + k) * RT + tid], tn = a.knots[(slot0 + k + 1) * RT + tid];
        const float delta = __fsub_rn(tn, tc);        // 0 for rows that took fewer than k+1 steps
        ext[dx * LDA + tid] = tc;
        ext[(dx + 1) * LDA + tid] = delta;
        m_delta[tid] = (k < (m_ke[tid] >> 1)) ? delta : 0.0f;
      }
      load_ckpt(slot0 + k, h);
      scale_tile<NB>(sc_kind, h);                     // h := s(h_k), the ODE net's vector input
      put_tile<NB>(hbuf, h, H);
      Tile<NB> d;
      for (int l = 0; l < T.L; ++l) {                 // hidden-layer outputs of this step: from the checkpoints
        load_plane<NB>(ckpt + ((slot0 + k) * planes + 1 + l) * RT * H, d);
        put_tile<NB>(zb[l], d, H);
      }
      __syncthreads();                                // m_delta / ext rows written by the first warp above
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float delta = m_delta[4 * ty + i];
#pragma unroll
        for (int j = 0; j < NB; ++j) d.v[i][j] = delta * g.v[i][j];        // h' = h + delta f(h): d f = delta g
      }
      net_backward<NB>(T, NET_ODE, p, part, hbuf, zb, nullptr, dbuf, wbuf, act_kind, d);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < NB; ++j) g.v[i][j] = fmaf(d.v[i][j], scale_grad_rt(sc_kind, h.v[i][j]), g.v[i][j]);
    }

    // ---- preds[u] = out(h0), then the jump net ----
    load_ckpt(slot0, h);
    out_backward(h, a.grad_preds, false, g);
    __syncthreads();
    if (tid < RT) {
      const int u = m_u[tid];
      for (int e = 0; e < dx; ++e) ext[e * LDA + tid] = u >= 0 ? a.values[(int64_t)u * dx + e] : 0.0f;
    }
    {
      Tile<NB> h0;
      net_forward<NB>(T, NET_JUMP, p, pt, hbuf, zb, wbuf, act_kind, h0);   // h0 again (bit-identical to the checkpoint)
      __syncthreads();
      put_tile<NB>(dbuf, h0, H);       // z_L of the jump net (post-activation): read by net_backward before it reuses dbuf
      // rows without a unit contribute nothing
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (m_u[4 * ty + i] < 0) {
#pragma unroll
          for (int j = 0; j < NB; ++j) g.v[i][j] = 0.0f;
        }
      net_backward<NB>(T, NET_JUMP, p, part, hbuf, zb, dbuf, dbuf, wbuf, act_kind, g);
    }
  }
}

template <int NB>
int launch_rowtile(const SweepArgs& a, cudaStream_t st, bool backward) {
  if (a.n_tiles == 0) return NJODE_OK;
  using L_ = Layout<NB>;
  if (backward) {
    NJODE_CUDA_OK(cudaFuncSetAttribute(k_rowtile_backward<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L_::bwd_bytes()));
    njode_timing_begin(2, st);
    k_rowtile_backward<NB><<<a.n_workers, NTH, L_::bwd_bytes(), st>>>(a);
    njode_timing_end(2, st);
    NJODE_LAUNCH_OK("k_rowtile_backward");
  } else {
    NJODE_CUDA_OK(cudaFuncSetAttribute(k_rowtile_forward<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L_::fwd_bytes()));
    njode_timing_begin(1, st);
    k_rowtile_forward<NB><<<a.n_workers, NTH, L_::fwd_bytes(), st>>>(a);
    njode_timing_end(1, st);
    NJODE_LAUNCH_OK("k_rowtile_forward");
  }
  return NJODE_OK;
}

int dispatch_rowtile(const SweepArgs& a, cudaStream_t st, bool backward) {
  switch (a.T.H / 32) {
    case 1: return launch_rowtile<1>(a, st, backward);
    case 2: return launch_rowtile<2>(a, st, backward);
    case 3: return launch_rowtile<3>(a, st, backward);
    default: return launch_rowtile<4>(a, st, backward);
  }
}

size_t bwd_smem(int nb) {
  switch (nb) {
    case 1: return Layout<1>::bwd_bytes();
    case 2: return Layout<2>::bwd_bytes();
    case 3: return Layout<3>::bwd_bytes();
    default: return Layout<4>::bwd_bytes();
  }
}

}  // namespace

int njode_rowtile_supported(const NjodeDesc* d) {
  const int O = d->shared_network ? d->d_y * d->num_moments : d->d_y;
  return d->hidden % 32 == 0 && d->hidden >= 32 && d->hidden <= 128 && d->n_hidden_layers <= LMAX_RT &&
         d->d_x + 2 <= EXT_MAX && O <= O_MAX;
}

// persistent CTAs: as many as fit per SM by shared memory (the reverse sweep is the bigger one), a multiple of S
int njode_rowtile_workers(const NjodeDesc* d, int64_t n_tiles) {
  const int S = d->shared_network ? 1 : d->num_moments;
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int per_sm = (int)((227 * 1024) / (bwd_smem(d->hidden / 32) + 1024));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 2) per_sm = 2;            // registers: 256 threads x ~120 allow two resident CTAs
  int64_t per_stack = (int64_t)sms * per_sm / S;
  if (per_stack < 1) per_stack = 1;
  if (per_stack > n_tiles) per_stack = n_tiles > 0 ? n_tiles : 1;
  return (int)(per_stack * S);
}

int njode_rowtile_forward(const SweepArgs& a, cudaStream_t st) { return dispatch_rowtile(a, st, false); }
int njode_rowtile_backward(const SweepArgs& a, cudaStream_t st) { return dispatch_rowtile(a, st, true); }
