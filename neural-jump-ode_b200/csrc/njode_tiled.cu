// njode_tiled.cu -- tuned FP32-FMA tile kernels (placeholder: not yet enabled)
#include "njode_common.cuh"
int njode_tiled_supported(const NjodeDesc* d) { (void)d; return 0; }
int njode_tiled_workers(const NjodeDesc* d, int64_t n_tiles) { (void)d; (void)n_tiles; return 1; }
int njode_tiled_forward(const SweepArgs& a, cudaStream_t st) { (void)a; (void)st; NJODE_FAIL(NJODE_EINVAL, "tiled kernels not built"); }
int njode_tiled_backward(const SweepArgs& a, cudaStream_t st) { (void)a; (void)st; NJODE_FAIL(NJODE_EINVAL, "tiled kernels not built"); }
