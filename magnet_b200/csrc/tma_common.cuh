// sm_100a primitives as inline PTX: mbarrier + TMA (cp.async.bulk[.tensor]) for cost_tma.cu / cost_mma.cu, tensor memory
// and tcgen05.mma (UMMA) for cost_mma.cu.
#pragma once
#include <cuda.h>   // CUtensorMap and its enums (types only — cuTensorMapEncodeTiled is resolved at run time)
#include <stdint.h>

namespace magnet {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t arrivals) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(arrivals) : "memory");
}
// make the barrier initialisation visible to the async proxy (TMA) before the first copy is issued
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// one arrival + the number of bytes the async copies of this phase will deliver
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(bar), "r"(parity) : "memory");
}

// TMA tiled load of one box of a rank-4 tensor map into shared memory (UTMALDG); out-of-range elements are
// zero-filled by the copy engine, coordinates are signed.
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tmap, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
      ::"r"(dst), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar) : "memory");
}

// rank-5 variant (cost_mma.cu: (channel, x, y, hi/lo plane, image))
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* tmap, uint32_t bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
      ::"r"(dst), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(bar) : "memory");
}

// order generic-proxy accesses of shared memory (ld/st.shared) before later async-proxy accesses (TMA writes, UMMA reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 1-D bulk copy global -> shared (UBLKCP); bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar) : "memory");
}

__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}

// ---- tensor memory (TMEM) as per-thread scratch ---------------------------------------------------------------
// 32x32b shape: thread i of warp w owns TMEM lane 32*(w%4)+i; a column holds one 32-bit word per lane.  The column
// index may be a run-time (warp-uniform) value, which lets the per-hypothesis loops stay ROLLED while their
// accumulators live outside the register file.  Loads are completed (wait::ld) inside the same asm statement, so the
// results cannot be consumed early.
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld2(uint32_t taddr, float& a, float& b) {
  uint32_t x, y;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];\ntcgen05.wait::ld.sync.aligned;"
               : "=r"(x), "=r"(y) : "r"(taddr) : "memory");
  a = __uint_as_float(x);
  b = __uint_as_float(y);
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float& a, float& b, float& c, float& d) {
  uint32_t x, y, z, w;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];\ntcgen05.wait::ld.sync.aligned;"
               : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "r"(taddr) : "memory");
  a = __uint_as_float(x);
  b = __uint_as_float(y);
  c = __uint_as_float(z);
  d = __uint_as_float(w);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\ntcgen05.wait::ld.sync.aligned;"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr) : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
// 16 consecutive accumulator columns of my TMEM lane (epilogue of cost_mma.cu)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- tcgen05.mma (UMMA), one CTA, fp16 operands from shared memory, fp32 accumulator in tensor memory ------------
// Shared-memory matrix descriptor of a K-major operand in the canonical 128-byte-swizzle layout (what TMA's
// CU_TENSOR_MAP_SWIZZLE_128B writes): rows of 128 bytes (64 fp16 of K), 8-row atoms of 1024 bytes, `sbo` bytes between
// consecutive atoms.  Bit fields as cute::UMMA::SmemDescriptor: start address >> 4 [0,14), leading byte offset >> 4
// [16,30) (unused for swizzled K-major, set to 1), stride byte offset >> 4 [32,46), version 1 [46,48), layout type
// SWIZZLE_128B = 2 [61,64).  A K step of 16 fp16 (32 bytes) inside the swizzle row is start address + 2.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// Instruction descriptor, kind::f16 (cute::UMMA::InstrDescriptor): D fp32 [4,6) = 1, A / B fp16 [7,10) / [10,13) = 0,
// both K-major [15] / [16] = 0, N >> 3 at [17,23), M >> 4 at [24,29).
__device__ __forceinline__ uint32_t umma_idesc_f16(uint32_t M, uint32_t N) {
  return (1u << 4) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]^T, issued by ONE thread
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on `bar` when all tcgen05.mma issued so far by this thread have completed (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_st2(uint32_t taddr, float a, float b) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "r"(__float_as_uint(a)),
               "r"(__float_as_uint(b)) : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, float a, float b, float c, float d) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(__float_as_uint(a)),
               "r"(__float_as_uint(b)), "r"(__float_as_uint(c)), "r"(__float_as_uint(d)) : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

}  // namespace magnet
