/*
 * njode.h -- C-ABI of libnjode_b200.so: the Neural Jump ODE training hot path on B200 (sm_100a).
 *
 * Plain C, raw device pointers and sizes, no torch types.  Every entry point replaces a piece of
 * the reference's Python hot path (file:line relative to the reference repository):
 *
 *   njode_schedule_*   the float32 Euler step rule            neural_jump_ode/models/jump_ode.py:188-203
 *   njode_forward      NeuralJumpODE.forward / forward_single neural_jump_ode/models/jump_ode.py:142-233
 *                      (JumpNN :15-26, ODEFunc :29-63, OutputNN :66-77, euler_step :122-140)
 *   njode_loss         nj_ode_loss (value and d/dpreds)       neural_jump_ode/models/jump_ode.py:235-383
 *   njode_backward     what loss.backward() does for this path neural_jump_ode/utils/training.py:69, :97
 *   njode_adam_step    optimizer.step() on the flat buffer    neural_jump_ode/utils/training.py:98, :396
 *   njode_dense_forward  the model on a dense time grid       neural_jump_ode/utils/plotting.py:133-256
 *
 * Conventions
 *   - All pointers are DEVICE pointers unless the name ends in _host.  The caller (PyTorch) owns
 *     every buffer; the library allocates nothing persistent and keeps no thread-local state
 *     except the last error string.
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises
 *     except where stated.
 *   - Return value: 0 on success, negative NJODE_E* otherwise; njode_last_error() gives the text.
 *     Nothing throws across the ABI.  There is no CPU fallback.
 *   - "Observation unit" o in [0,N): one (trajectory, observation) pair of the packed batch.  Unit o
 *     computes h=jump(x_o), preds[o]=out(h) and, if o is not the last observation of its trajectory,
 *     integrates to t_{o+1} and writes preds_before[o+1]=out(h_end).  Units are independent
 *     initial-value problems because the jump resets h from x_o alone (jump_ode.py:169, :176).
 *
 * Packed batch layout (all float32, C-contiguous)
 *   times        (N)            observation times, trajectory after trajectory
 *   values       (N, d_x)
 *   obs_offsets  (B+1) int64    trajectory b owns observations [obs_offsets[b], obs_offsets[b+1])
 *   preds, preds_before (N, d_y, M); preds_before[first observation of a trajectory] == 0
 *
 * Flat parameter layout (float32): for stack s = 0..S-1 (S = 1 if shared_network else num_moments):
 *   jump net   layers i=0..L:  W (H x in_i) row-major, b (H);     in_0 = d_x,        in_i = H
 *   ode  net   layers i=0..L:  W (H x in_i) row-major, b (H);     in_0 = H+d_x+2,    in_i = H
 *                              columns of layer 0 = [h(0..H-1), x(0..d_x-1), t_cur, t_new-t_cur]
 *   out  net   layers i=0..L:  W (out_i x H) row-major, b(out_i); out_i = H, out_L = O
 *   O = d_y (separate networks) or d_y*M (shared; flat output index = d*M + m, jump_ode.py:172)
 *   i.e. exactly the reference's nn.Linear weights (state_dict keys *.net.{3i}.{weight,bias}).
 */
#ifndef NJODE_H_
#define NJODE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NJODE_ABI_VERSION 2

enum {
  NJODE_OK = 0,
  NJODE_EINVAL = -1,      /* bad argument / unsupported configuration */
  NJODE_ECUDA = -2,       /* CUDA runtime error */
  NJODE_EWORKSPACE = -3,  /* workspace too small */
  NJODE_ECAPACITY = -4    /* njode_forward_batch: arena / checkpoint buffer smaller than this batch needs (header_host is valid) */
};

/* jump_ode.py:6-13; unknown names map to RELU on the Python side (jump_ode.py:18) */
enum { NJODE_ACT_RELU = 0, NJODE_ACT_TANH = 1, NJODE_ACT_SIGMOID = 2, NJODE_ACT_ELU = 3,
       NJODE_ACT_LEAKY_RELU = 4, NJODE_ACT_SELU = 5 };
/* jump_ode.py:43-50 */
enum { NJODE_SCALE_IDENTITY = 0, NJODE_SCALE_TANH = 1, NJODE_SCALE_SIGMOID = 2 };
/* jump_ode.py:333 / :346 */
enum { NJODE_VAR_DIRECT = 0, NJODE_VAR_SECOND_MOMENT = 1 };
/* kernel flavour: AUTO picks TILED (tcgen05, hidden 32 / 1 layer), then WIDE (tcgen05, hidden 64 / 128, <= 3 layers),
 * then ROWTILE (FP32 FMA), then GENERIC */
enum { NJODE_IMPL_AUTO = 0, NJODE_IMPL_GENERIC = 1, NJODE_IMPL_TILED = 2, NJODE_IMPL_ROWTILE = 3, NJODE_IMPL_WIDE = 4 };

typedef struct NjodeDesc {
  int32_t d_x;              /* input_dim */
  int32_t d_y;              /* output_dim */
  int32_t hidden;           /* hidden_dim H */
  int32_t n_hidden_layers;  /* L >= 1 */
  int32_t num_moments;      /* M >= 1 */
  int32_t shared_network;   /* 0/1 */
  int32_t activation;       /* NJODE_ACT_* */
  int32_t input_scaling;    /* NJODE_SCALE_* */
  int32_t has_dt;           /* 0: dt_ode_step is None -> one Euler step per interval */
  float   dt;               /* dt_ode_step as float32 */
  int32_t impl;             /* NJODE_IMPL_* */
  int32_t reserved;
} NjodeDesc;

typedef struct NjodeLossDesc {
  int32_t ignore_first_continuity;
  int32_t variance_method;  /* NJODE_VAR_* */
  float   eps;              /* 1e-10 in the reference */
  float   w0, w1;           /* moment weights (1,1 when moment_weights is None) */
  int32_t reserved;
} NjodeLossDesc;

/* schedule header written by njode_schedule_build (device int64[8]) */
enum { NJODE_HDR_TOTAL_STEPS = 0,   /* sum over trajectories of Euler steps ("trajectory-ODE-steps") */
       NJODE_HDR_TOTAL_SLOTS = 1,   /* checkpoint slots: sum over tiles of (kmax_tile + 1 + flavour extras) */
       NJODE_HDR_NUM_TILES = 2,
       NJODE_HDR_KMAX = 3,
       NJODE_HDR_WORDS = 8 };

int32_t     njode_abi_version(void);
const char* njode_last_error(void);

/* number of float32 parameters of one stack / of the whole model; -1 on a bad descriptor */
int64_t njode_params_per_stack(const NjodeDesc* desc);
int64_t njode_param_count(const NjodeDesc* desc);
int32_t njode_num_stacks(const NjodeDesc* desc);

/* rows per tile of the kernel flavour that (desc) selects; the schedule is built for it */
int32_t njode_tile_rows(const NjodeDesc* desc);
/* the flavour (NJODE_IMPL_*, never AUTO) that runs for this descriptor; -1 on a bad / unsupported descriptor */
int32_t njode_selected_impl(const NjodeDesc* desc);

/* ---- step schedule (jump_ode.py:188-203, float32 accumulation reproduced bit for bit) ---------
 * build: kenc[o] = (K_o << 1) | has_next_o ; perm = units sorted by K descending, padded with -1
 *        to n_tiles*tile_rows ; tile_kmax ; tile_slot_off = exclusive scan of (kmax+1) ; header.
 * knots: knots[(slot_off[tile]+k)*tile_rows + r] = t_k of row r, k = 0..kmax_tile (float32);
 *        t_0 = t_o, t_K = t_{o+1}; entries beyond a row's own K repeat t_K.
 * The caller reads `header` back (one small D2H copy) to size knots / checkpoints. */
size_t njode_schedule_workspace_bytes(int64_t B, int64_t N, int32_t tile_rows);
int njode_schedule_build(const NjodeDesc* desc, const float* times, const int64_t* obs_offsets,
                         int64_t B, int64_t N, int32_t tile_rows,
                         int32_t* kenc, int32_t* perm, int32_t* tile_kmax, int64_t* tile_slot_off,
                         int64_t* header, void* workspace, size_t workspace_bytes, void* stream);
int njode_schedule_knots(const float* times, const int32_t* kenc, const int32_t* perm,
                         const int32_t* tile_kmax, const int64_t* tile_slot_off,
                         int64_t N, int64_t n_tiles, int32_t tile_rows, const NjodeDesc* desc,
                         float* knots, void* stream);

/* ---- forward sweep -----------------------------------------------------------------------------
 * ckpt: float32 [S][total_slots][tile_rows * Hc] per-slot checkpoints (Hc = njode_ckpt_row_floats): the hidden
 *       state before every Euler step and after the last one, plus -- tiled / row-tiled flavours -- the ODE
 *       net's hidden-layer outputs of that step; opaque to the caller, written by njode_forward and read by
 *       njode_backward (the WIDE flavour's reverse sweep also WRITES its half of the buffer: per-layer data
 *       gradients for the weight-gradient GEMM).  Pass NULL for inference (no checkpoints written).
 * workspace: njode_forward_workspace_bytes (re-laid-out weights). */
int64_t njode_ckpt_row_floats(const NjodeDesc* desc);
/* Number of tiles the schedule of N observation units has (tiles hold tile_rows rows, of which a flavour- and
 * size-dependent number carries units; the rest is padding).  Sizes perm (n_tiles * tile_rows), tile_kmax, tile_slot_off. */
int64_t njode_num_tiles(const NjodeDesc* desc, int64_t N);
size_t  njode_forward_workspace_bytes(const NjodeDesc* desc);
int njode_forward(const NjodeDesc* desc, const float* params, const float* times, const float* values,
                  const int64_t* obs_offsets, int64_t B, int64_t N,
                  const int32_t* kenc, const int32_t* perm, const int32_t* tile_kmax,
                  const int64_t* tile_slot_off, const float* knots,
                  int64_t n_tiles, int64_t total_slots, int32_t tile_rows,
                  float* preds, float* preds_before, float* ckpt,
                  void* workspace, size_t workspace_bytes, void* stream);

/* ---- un-cached batch in ONE call: schedule build -> header to the host -> knots -> forward sweep -------------
 * What NeuralJumpODE.forward (jump_ode.py:218-233) costs when every training step brings a new batch: the three
 * calls above need a host round trip in the middle (the header sizes knots / checkpoints), and whatever the host
 * does around that round trip is time the GPU idles.  Here the caller passes buffers sized from a guess:
 *   arena   device, persistent for the life of the schedule: kenc, perm, tile_kmax, tile_slot_off, header, knots at
 *           the byte offsets njode_batch_arena_bytes reports in layout[NJODE_ARENA_*] (all but knots are independent
 *           of total_slots; knots come last, so an arena that is too large is fine)
 *   ckpt    device, ckpt_floats >= S * total_slots * tile_rows * njode_ckpt_row_floats, or want_ckpt = 0
 *   scratch device, njode_batch_scratch_bytes, free again when the call's work on `stream` is done
 *   header_host  host int64[NJODE_HDR_WORDS] (pinned for a truly asynchronous copy); valid on return
 * The call synchronises `stream` once, after the schedule is built.  If arena or ckpt turn out too small it returns
 * NJODE_ECAPACITY before any sweep work: size them from header_host[NJODE_HDR_TOTAL_SLOTS] and call again. */
enum { NJODE_ARENA_KENC = 0, NJODE_ARENA_PERM, NJODE_ARENA_TILE_KMAX, NJODE_ARENA_TILE_SLOT_OFF, NJODE_ARENA_HEADER,
       NJODE_ARENA_KNOTS, NJODE_ARENA_WORDS = 8 };
size_t njode_batch_arena_bytes(const NjodeDesc* desc, int64_t B, int64_t N, int64_t total_slots, int64_t* layout);
size_t njode_batch_scratch_bytes(const NjodeDesc* desc, int64_t B, int64_t N);
int njode_forward_batch(const NjodeDesc* desc, const float* params, const float* times, const float* values,
                        const int64_t* obs_offsets, int64_t B, int64_t N,
                        void* arena, size_t arena_bytes, int32_t want_ckpt, float* ckpt, int64_t ckpt_floats,
                        void* scratch, size_t scratch_bytes, int64_t* header_host,
                        float* preds, float* preds_before, void* stream);
/* The same in two halves, for callers with host work of their own to do while the schedule is being built on the
 * device: _begin launches the schedule build and the header copy and returns at once (it needs neither the
 * parameters nor the output buffers); _finish synchronises `stream`, checks the capacities and launches knots +
 * forward sweep.  njode_forward_batch = _begin then _finish with the same buffers. */
int njode_forward_batch_begin(const NjodeDesc* desc, const float* times, const int64_t* obs_offsets,
                              int64_t B, int64_t N, void* arena, size_t arena_bytes,
                              void* scratch, size_t scratch_bytes, int64_t* header_host, void* stream);
int njode_forward_batch_finish(const NjodeDesc* desc, const float* params, const float* times, const float* values,
                               const int64_t* obs_offsets, int64_t B, int64_t N,
                               void* arena, size_t arena_bytes, int32_t want_ckpt, float* ckpt, int64_t ckpt_floats,
                               void* scratch, size_t scratch_bytes, int64_t* header_host,
                               float* preds, float* preds_before, void* stream);

/* ---- dense-grid inference: the prediction at every time of a grid (utils/plotting.py:133-256) -------------------------
 * dense: float32 (B, G, d_y, M) raw readouts (mean, W), overwritten.  grid: (G) float32 ascending, shared by all
 * trajectories.  The step rule is the plotting code's, not the training one: from the current time to each grid time,
 * n_sub = max(1, int((t - t_cur) / dt_ode_step)) equal Euler sub-steps (1 if dt_ode_step is None), float32 throughout;
 * a grid time equal to an observation time holds the post-jump value, except at a trajectory's LAST observation
 * (pre-jump, plotting.py:210), and grid times before the first observation are 0.  Forward only.
 * workspace: njode_dense_workspace_bytes. */
size_t njode_dense_workspace_bytes(const NjodeDesc* desc);
int njode_dense_forward(const NjodeDesc* desc, const float* params, const float* times, const float* values,
                        const int64_t* obs_offsets, int64_t B, int64_t N, const float* grid, int64_t G,
                        float* dense, void* workspace, size_t workspace_bytes, void* stream);

/* ---- loss: value and gradient w.r.t. preds / preds_before in one pass --------------------------
 * loss_out: device float[1].  grad_* may be NULL (value only).  traj_scale = 1/B_global so that
 * data-parallel ranks sum to the single-GPU loss (pass 1/B for one GPU).
 * workspace: njode_loss_workspace_bytes(B). */
size_t njode_loss_workspace_bytes(int64_t B);
int njode_loss(const NjodeLossDesc* ldesc, const float* values, const float* preds,
               const float* preds_before, const int64_t* obs_offsets, int64_t B, int64_t N,
               int32_t d, int32_t M, float traj_scale,
               float* loss_out, float* grad_preds, float* grad_preds_before,
               void* workspace, size_t workspace_bytes, void* stream);

/* ---- reverse sweep (discretise-then-optimise: exact adjoint of every Euler step) ----------------
 * grad_params: float32 [njode_param_count], overwritten (not accumulated).
 * workspace: njode_backward_workspace_bytes. */
size_t njode_backward_workspace_bytes(const NjodeDesc* desc, int64_t n_tiles);
int njode_backward(const NjodeDesc* desc, const float* params, const float* times, const float* values,
                   const int64_t* obs_offsets, int64_t B, int64_t N,
                   const int32_t* kenc, const int32_t* perm, const int32_t* tile_kmax,
                   const int64_t* tile_slot_off, const float* knots,
                   int64_t n_tiles, int64_t total_slots, int32_t tile_rows,
                   const float* grad_preds, const float* grad_preds_before, float* ckpt,
                   float* grad_params, void* workspace, size_t workspace_bytes, void* stream);

/* ---- Adam on the flat parameter buffer (torch.optim.Adam semantics, weight_decay as L2-in-grad) --
 * step is the 1-based step count AFTER this update. grad_scale multiplies grad first (e.g. 1.0). */
int njode_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                    float lr, float beta1, float beta2, float eps, float weight_decay,
                    int64_t step, float grad_scale, void* stream);


/* ---- measurement hooks (bench.py) ---------------------------------------------------------------
 * njode_set_kernel_timing: the next njode_forward (which=1) / njode_backward (which=2; which=3: the WIDE flavour's
 * weight-gradient GEMM) call records the two caller-owned cudaEvent_t around that kernel only (on the call's stream).  One-shot;
 * pass NULLs to clear.  njode_ffma_peak: runs an FFMA-only kernel on the current device and returns the
 * measured dense FP32 FMA throughput in TFLOP/s (the roofline denominator of the FP32 path; it is not
 * in MEASURED_PEAKS.json).  Synchronises the device. */
int njode_set_kernel_timing(int32_t which, void* ev_start, void* ev_stop);
/* sticky device-side diagnostic word of the tcgen05 kernels (0 = healthy; bit 0 / 1: a TILED forward / reverse
 * sweep CTA gave up waiting on an mbarrier; bit 2 / 3 / 4: WIDE forward / reverse / weight-gradient).  A kernel that
 * gives up TRAPS after setting its bit, so every later CUDA call on the context fails: a protocol failure can never
 * yield silently wrong numbers.  Synchronises the device. */
int njode_device_status(uint32_t* status_host);
/* bring-up detail of the WIDE kernels, uint32[4]: {sweep status, first sweep wait site that gave up, weight-gradient
 * status, its first site}; site = code | warp << 8 | block << 16.  With NJODE_NO_TRAP=1 in the environment a kernel
 * that gives up records the site and runs on without waiting (results are garbage) instead of trapping. */
int njode_device_status_detail(uint32_t* words_host);
/* bring-up aid: {SM cycles, chain GEMMs} of each of the first n_ctas CTAs of the most recent WIDE sweep launch (forward
 * or reverse), uint64[n_ctas][2], n_ctas <= 512.  Synchronises the device. */
int njode_debug_cta_cycles(unsigned long long* out_host, int n_ctas);
/* phase-accounting build only (`make phase`, tools/phase_wide.py): per-CTA cycles per phase of the WIDE kernels,
 * uint64[n_ctas][8]; which = 1: the last sweep launch, 3: the weight-gradient GEMM.  Zeros in the product build. */
int njode_debug_phase(int32_t which, unsigned long long* out_host, int32_t n_ctas);
/* Number of CUDA kernels this library has launched in this process (every launch site counts itself);
 * reset != 0 returns the count and sets it to zero.  Measurement aid for bench.py's `gpu_launches`. */
int64_t njode_kernel_launches(int32_t reset);
int njode_ffma_peak(float* tflops_host);

#ifdef __cplusplus
}
#endif
#endif /* NJODE_H_ */
