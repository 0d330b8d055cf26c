// njode_umma.cuh -- thin wrappers over the Blackwell tcgen05 (UMMA) / TMEM / mbarrier PTX used by the
// tiled sweep kernels.  Every encoding here was validated on a B200 by tools/umma_probe*.cu:
//   * smem matrix descriptor: start>>4 | LBO>>4 <<16 | SBO>>4 <<32 | version 1 <<46 | layout <<61
//       K-major  fp32/tf32 tiles: layout 2 (SWIZZLE_128B), 128-byte rows, 16-byte chunk ^= row%8, SBO = 1024
//       MN-major fp32/tf32 tiles: layout 1 (SWIZZLE_128B_BASE32B, the only MN-major layout for tf32),
//                                 128-byte rows, 32-byte chunk ^= row%4, SBO = 512, LBO = next 32-wide block
//   * instruction descriptor (kind::tf32): c=F32 (1<<4), a=b=TF32 (2<<7, 2<<10), a_major<<15, b_major<<16,
//       N>>3 <<17, M>>4 <<24
//   * A operand from TMEM (lane = row, column = k) for the chain GEMMs; accumulators in TMEM;
//       an M=64 accumulator keeps row i in lane (i%16) + 32*(i/16)
//   * 3xTF32 split: x = hi + lo, hi = rna_tf32(x), lo = rna_tf32(x - hi);  A*B ~= Al*Bh + Ah*Bl + Ah*Bh
//       (measured 1.5e-7 .. 4.8e-7 relative error against fp64, tools/umma_probe.cu tests 2 and 4)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}
__device__ __forceinline__ uint64_t desc_k(uint32_t saddr) { return desc(saddr, 16, 1024, 2); }
__device__ __forceinline__ uint64_t desc_mn(uint32_t saddr, uint32_t lbo_bytes) { return desc(saddr, lbo_bytes, 512, 1); }

__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// float index of element (r, c) inside a [rows][32] tile
__host__ __device__ __forceinline__ int swz_k(int r, int c) { return r * 32 + (((c >> 2) ^ (r & 7)) << 2) + (c & 3); }
__host__ __device__ __forceinline__ int swz_mn(int r, int c) { return r * 32 + (((c >> 3) ^ (r & 3)) << 3) + (c & 7); }

__device__ __forceinline__ float tf32_hi(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// D[tmem_d] (+)= A[tmem_a] * B[smem desc]   (A from TMEM, K-major; one k-step of 8)
__device__ __forceinline__ void mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
               :: "r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
// D[tmem_d] (+)= A[smem desc] * B[smem desc]
__device__ __forceinline__ void mma_ss(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
               :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;"); }
// bounded wait: returns false on timeout (the caller records an error instead of hanging the GPU)
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
#pragma unroll 1
  for (uint32_t spin = 0; spin < (1u << 24); ++spin) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(done) : "r"(a), "r"(parity) : "memory");
    if (done) return true;
  }
  return false;
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {   // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(dst_smem)), "r"(ncols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
}
__device__ __forceinline__ void tmem_free(uint32_t taddr, uint32_t ncols) {        // the same warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(ncols));
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t u[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]) : "r"(taddr));
  wait_ld();
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(u[i]);
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               :: "r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
                 "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
                 "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])) : "memory");
}

// 16 / 8 consecutive columns of this thread's TMEM lane <-> registers (two warps share a lane quadrant and
// each owns one 16-column half of a 32-column operand / accumulator)
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, float (&v)[16]) {
  uint32_t u[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
                 "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]) : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(u[i]);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t u[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
                 "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]) : "r"(taddr));
  wait_ld();
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(u[i]);
}
__device__ __forceinline__ void tmem_st16_raw(uint32_t taddr, const uint32_t (&u)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
               :: "r"(taddr), "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]), "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7]),
                 "r"(u[8]), "r"(u[9]), "r"(u[10]), "r"(u[11]), "r"(u[12]), "r"(u[13]), "r"(u[14]), "r"(u[15]) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
  uint32_t u[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) u[i] = __float_as_uint(v[i]);
  tmem_st16_raw(taddr, u);
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
// one lane of a fully converged warp
__device__ __forceinline__ bool elect_one() {
  uint32_t e;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(e));
  return e != 0;
}
// named barrier among `count` threads (count a multiple of 32)
__device__ __forceinline__ void named_bar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(count) : "memory");
}

// split a row of 32 floats into tf32 hi / lo parts.  hi = round-to-nearest(ties away) to 10 explicit
// mantissa bits = (bits + 0x1000) & ~0x1fff (2 integer ops; cvt.rna.tf32.f32 costs 4 SASS ops because it
// also special-cases inf/NaN); lo = x - hi is exact in fp32 and is handed to the tensor core unrounded
// (the MMA reads the top 19 bits): dropped part <= 2^-21 |x|.  3 instructions per element instead of 9.
__device__ __forceinline__ void split1(float x, uint32_t& hi, uint32_t& lo) {
  hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}

// 8 / 2 columns per thread: four warps share a TMEM lane quadrant (warp % 4) and each owns an 8-column
// slice of a 32-column operand / accumulator (tools/umma_probe4.cu validated the shared-quadrant access)
__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, float (&v)[8]) {
  uint32_t u[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]) : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(u[i]);
}
__device__ __forceinline__ void tmem_st8_raw(uint32_t taddr, const uint32_t (&u)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               :: "r"(taddr), "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]), "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7]) : "memory");
}
__device__ __forceinline__ void tmem_ld2_nowait(uint32_t taddr, float (&v)[2]) {
  uint32_t u0, u1;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(u0), "=r"(u1) : "r"(taddr));
  v[0] = __uint_as_float(u0);
  v[1] = __uint_as_float(u1);
}
__device__ __forceinline__ void tmem_st2(uint32_t taddr, const float (&v)[2]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%2};"
               :: "r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])) : "memory");
}
__device__ __forceinline__ void split8(const float (&v)[8], uint32_t (&hi)[8], uint32_t (&lo)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) split1(v[i], hi[i], lo[i]);
}
// columns [8c, 8c+8) of row r -> MN-major tile (one 32-byte chunk, position c ^ r%4); `tile_saddr` is the
// shared-space address of the tile (st.shared: the generic-pointer form compiles to slower generic stores)
__device__ __forceinline__ void st_shared_v4(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" :: "r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void chunk_to_mn_tile(uint32_t tile_saddr, int r, int c, const uint32_t (&v)[8]) {
  // Rows r and r+4 of a quarter-warp share the chunk position, so writing "low half, then high half" in every
  // lane is a 2-way bank conflict on each st.shared.v4 (measured: 8 wavefronts instead of 4).  Odd (r/4) lanes
  // write the high half first: 8 SELs per chunk buy back half of the shared-memory pipe time of the tile writes.
  const bool sw = (r >> 2) & 1;
  const uint32_t addr = tile_saddr + (uint32_t)r * 128u + (uint32_t)((c ^ (r & 3)) * 32) + (sw ? 16u : 0u);
  st_shared_v4(addr, sw ? v[4] : v[0], sw ? v[5] : v[1], sw ? v[6] : v[2], sw ? v[7] : v[3]);
  st_shared_v4(addr ^ 16u, sw ? v[0] : v[4], sw ? v[1] : v[5], sw ? v[2] : v[6], sw ? v[3] : v[7]);
}

}  // namespace umma
