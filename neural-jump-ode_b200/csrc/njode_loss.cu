// njode_loss.cu -- nj_ode_loss (jump_ode.py:235-383): value and closed-form gradient w.r.t.
// preds / preds_before in one pass, plus the flat-buffer Adam step.
//
// Per observation (X: d values, Y/Yb: d x M predictions after / before the jump):
//   a = sum_d (X-Y0)^2, c = sum_d (X-Yb0)^2 (c := 0 at a trajectory's first observation when
//   ignore_first_continuity, jump_ode.py:315-317), l0 = (sqrt(a+eps)+sqrt(c+eps))^2     (:320)
//   M>1, direct:        V=W^2, Z=(X-Y0.detach())^2, Zb=(X-Yb0.detach())^2                (:336-344)
//        second_moment: V=W,   Z=Zb=X^2                                                  (:349-353)
//        l1 = (sqrt(sum_d (Z-V)^2+eps) + sqrt(sum_d (Zb-Vb)^2+eps))^2                    (:362-373)
//   trajectory loss = w0*mean_i l0 + w1*mean_i l1 (:321-325, :374-378); moments >= 2 are ignored;
//   batch loss = mean over trajectories (:383).
#include "njode_common.cuh"

#define LOSS_TB 128
#define LOSS_TRAJ_PER_BLOCK (LOSS_TB / 32)

// One WARP per trajectory: lane l takes observations l, l + 32, ... (a thread per trajectory walked its ~10-100
// observations serially, one dependent load chain each: 11 us for 4096 trajectories on 32 SMs); the lanes' sums are
// combined by a fixed xor-shuffle tree, the block's 4 trajectories and then the blocks in double: deterministic.
__global__ void __launch_bounds__(LOSS_TB)
k_loss(NjodeLossDesc ld, const float* __restrict__ X, const float* __restrict__ Y, const float* __restrict__ Yb,
       const int64_t* __restrict__ off, int64_t B, int d, int M, float traj_scale,
       float* __restrict__ gY, float* __restrict__ gYb, double* __restrict__ block_part) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t b = (int64_t)blockIdx.x * LOSS_TRAJ_PER_BLOCK + w;
  float traj_loss = 0.0f;
  if (b < B) {
    const int64_t lo = off[b], hi = off[b + 1];
    const int64_t n = hi - lo;
    const float inv_n = 1.0f / (float)n;                 // n == 0 -> NaN like torch's mean of empty
    const float g0 = ld.w0 * traj_scale * inv_n;
    const float g1 = ld.w1 * traj_scale * inv_n;
    const bool direct = ld.variance_method == NJODE_VAR_DIRECT;
    float sum0 = 0.0f, sum1 = 0.0f;
    for (int64_t o = lo + lane; o < hi; o += 32) {
      const bool keep = !(ld.ignore_first_continuity && o == lo);
      const float* x = X + o * d;
      const float* y = Y + o * d * M;
      const float* yb = Yb + o * d * M;
      float a = 0.0f, c = 0.0f, va = 0.0f, vc = 0.0f;
      for (int k = 0; k < d; ++k) {
        const float e = x[k] - y[k * M], eb = x[k] - yb[k * M];
        a += e * e;
        c += eb * eb;
        if (M > 1) {
          const float w1 = y[k * M + 1], wb = yb[k * M + 1];
          const float z = direct ? e * e : x[k] * x[k];
          const float zb = direct ? eb * eb : x[k] * x[k];
          const float v = direct ? w1 * w1 : w1, vb = direct ? wb * wb : wb;
          va += (z - v) * (z - v);
          vc += (zb - vb) * (zb - vb);
        }
      }
      if (!keep) { c = 0.0f; vc = 0.0f; }
      const float sa = sqrtf(a + ld.eps), sc = sqrtf(c + ld.eps);
      sum0 += (sa + sc) * (sa + sc);
      float sva = 1.0f, svc = 1.0f;
      if (M > 1) {
        sva = sqrtf(va + ld.eps);
        svc = sqrtf(vc + ld.eps);
        sum1 += (sva + svc) * (sva + svc);
      }
      if (gY) {
        // d l/d a = (sa+sc)/sa ; d a/d Y0 = -2 (X-Y0)
        const float da = g0 * (sa + sc) / sa, dc = keep ? g0 * (sa + sc) / sc : 0.0f;
        const float dva = g1 * (sva + svc) / sva, dvc = keep ? g1 * (sva + svc) / svc : 0.0f;
        float* gy = gY + o * d * M;
        float* gyb = gYb + o * d * M;
        for (int k = 0; k < d; ++k) {
          const float e = x[k] - y[k * M], eb = x[k] - yb[k * M];
          gy[k * M] = -2.0f * e * da;
          gyb[k * M] = -2.0f * eb * dc;
          if (M > 1) {
            const float w1 = y[k * M + 1], wb = yb[k * M + 1];
            const float z = direct ? e * e : x[k] * x[k];
            const float zb = direct ? eb * eb : x[k] * x[k];
            const float v = direct ? w1 * w1 : w1, vb = direct ? wb * wb : wb;
            // d (z-v)^2 / d w = -2 (z-v) * dv/dw ;  dv/dw = 2w (direct) or 1
            gy[k * M + 1] = -2.0f * (z - v) * (direct ? 2.0f * w1 : 1.0f) * dva;
            gyb[k * M + 1] = -2.0f * (zb - vb) * (direct ? 2.0f * wb : 1.0f) * dvc;
            for (int m = 2; m < M; ++m) { gy[k * M + m] = 0.0f; gyb[k * M + m] = 0.0f; }
          }
        }
      }
    }
    for (int o = 16; o > 0; o >>= 1) {
      sum0 += __shfl_xor_sync(0xffffffffu, sum0, o);
      sum1 += __shfl_xor_sync(0xffffffffu, sum1, o);
    }
    traj_loss = ld.w0 * (sum0 * inv_n);
    if (M > 1) traj_loss += ld.w1 * (sum1 * inv_n);
  }
  // deterministic block sum in double
  __shared__ double sh[LOSS_TRAJ_PER_BLOCK];
  if (lane == 0) sh[w] = (double)traj_loss;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < LOSS_TRAJ_PER_BLOCK; ++i) s += sh[i];
    block_part[blockIdx.x] = s;
  }
}

__global__ void k_loss_final(const double* __restrict__ block_part, int64_t nblocks, float traj_scale,
                             float* __restrict__ loss_out) {
  __shared__ double sh[256];
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < nblocks; i += blockDim.x) s += block_part[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int k = 128; k > 0; k >>= 1) {
    if ((int)threadIdx.x < k) sh[threadIdx.x] += sh[threadIdx.x + k];
    __syncthreads();
  }
  if (threadIdx.x == 0) loss_out[0] = (float)(sh[0] * (double)traj_scale);
}

extern "C" size_t njode_loss_workspace_bytes(int64_t B) {
  return (size_t)((B + LOSS_TRAJ_PER_BLOCK - 1) / LOSS_TRAJ_PER_BLOCK + 1) * sizeof(double);
}

extern "C" int njode_loss(const NjodeLossDesc* ld, const float* values, const float* preds,
                          const float* preds_before, const int64_t* obs_offsets, int64_t B, int64_t N,
                          int32_t d, int32_t M, float traj_scale,
                          float* loss_out, float* grad_preds, float* grad_preds_before,
                          void* workspace, size_t workspace_bytes, void* stream) {
  if (!ld) NJODE_FAIL(NJODE_EINVAL, "njode_loss: null loss descriptor");
  if (ld->variance_method != NJODE_VAR_DIRECT && ld->variance_method != NJODE_VAR_SECOND_MOMENT)
    NJODE_FAIL(NJODE_EINVAL, "njode_loss: unknown variance_method code %d", ld->variance_method);
  if (B < 1 || N < 0 || d < 1 || M < 1) NJODE_FAIL(NJODE_EINVAL, "njode_loss: bad sizes (B=%lld N=%lld d=%d M=%d)",
                                                    (long long)B, (long long)N, d, M);
  if ((grad_preds == nullptr) != (grad_preds_before == nullptr))
    NJODE_FAIL(NJODE_EINVAL, "njode_loss: grad_preds and grad_preds_before must both be given or both be NULL");
  if (workspace_bytes < njode_loss_workspace_bytes(B)) NJODE_FAIL(NJODE_EWORKSPACE, "njode_loss: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t nblocks = (B + LOSS_TRAJ_PER_BLOCK - 1) / LOSS_TRAJ_PER_BLOCK;
  k_loss<<<(unsigned)nblocks, LOSS_TB, 0, st>>>(*ld, values, preds, preds_before, obs_offsets, B, d, M, traj_scale,
                                                grad_preds, grad_preds_before, (double*)workspace);
  NJODE_LAUNCH_OK("k_loss");
  k_loss_final<<<1, 256, 0, st>>>((const double*)workspace, nblocks, traj_scale, loss_out);
  NJODE_LAUNCH_OK("k_loss_final");
  return NJODE_OK;
}

// ------------------------------------------------------------------------------------------------
// Adam (torch.optim.Adam, amsgrad=False, weight_decay added to the gradient; training.py:396)
// ------------------------------------------------------------------------------------------------
__global__ void k_adam(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                       float* __restrict__ v, int64_t n, float lr, float b1, float b2, float eps, float wd,
                       float bc1, float bc2_sqrt, float grad_scale) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float gi = g[i] * grad_scale;
  const float pi = p[i];
  if (wd != 0.0f) gi = fmaf(wd, pi, gi);
  const float mi = m[i] + (1.0f - b1) * (gi - m[i]);          // lerp form used by torch
  const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
  m[i] = mi;
  v[i] = vi;
  const float denom = sqrtf(vi) / bc2_sqrt + eps;
  p[i] = pi - (lr / bc1) * (mi / denom);
}

extern "C" int njode_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                               float lr, float beta1, float beta2, float eps, float weight_decay,
                               int64_t step, float grad_scale, void* stream) {
  if (n < 0 || step < 1) NJODE_FAIL(NJODE_EINVAL, "njode_adam_step: bad n/step");
  if (n == 0) return NJODE_OK;
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  k_adam<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, (float)bc1, (float)sqrt(bc2), grad_scale);
  NJODE_LAUNCH_OK("k_adam");
  return NJODE_OK;
}
