// sm_100a building blocks shared by the tcgen05 kernels: mbarrier / TMA / tcgen05 PTX wrappers,
// UMMA descriptors, and the host-side tensor-map encoder (driver entry point fetched at run
// time so that the library links against nothing but the static CUDA runtime).
#pragma once
#include "common.cuh"
#include <cuda.h>

namespace i2l {
namespace tc {

// ------------------------------------------------------------------ device: barriers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t a, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t a) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t a, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t a, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "W_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra D_%=;\n\t"
      "bra W_%=;\n\t"
      "D_%=:\n\t}" ::"r"(a), "r"(parity) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ------------------------------------------------------------------ device: bulk / TMA copies
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(tm), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst),
      "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3, int c4,
                                            uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(dst),
      "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(bar)
      : "memory");
}
// im2col-mode load of an NHWC tensor (tensor map from make_im2col_map): `pixels` consecutive output
// positions starting at base input pixel (w, h, n), filter tap offsets (off_w, off_h), channels from c
__device__ __forceinline__ void tma_load_im2col(uint32_t dst, const CUtensorMap* tm, int c, int w, int h, int n,
                                                uint32_t off_w, uint32_t off_h, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6], {%7, %8};" ::"r"(dst),
      "l"(tm), "r"(c), "r"(w), "r"(h), "r"(n), "r"(bar), "h"((unsigned short)off_w), "h"((unsigned short)off_h)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory");
}

// ------------------------------------------------------------------ device: tcgen05
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "n"(COLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t tmem) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(COLS) : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]
__device__ __forceinline__ void tc_mma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]   (A rows = TMEM lanes, two bf16 K-elements per 32-bit column)
__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tc_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tc_ld16_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_ld8_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major shared-memory matrix descriptor.  swizzle_bytes in {32, 64, 128}; rows are
// swizzle_bytes long, 8-row groups are 8*swizzle_bytes apart (stride byte offset).
__host__ __device__ constexpr uint64_t desc_hi(int swizzle_bytes) {
  return ((uint64_t)((8 * swizzle_bytes) >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)(swizzle_bytes == 128 ? 2 : (swizzle_bytes == 64 ? 4 : 6)) << 61);
}
__device__ __forceinline__ uint64_t desc_base(uint32_t smem_base, int swizzle_bytes) {
  return desc_hi(swizzle_bytes) | (uint64_t)(((smem_base >> 4) & 0x3FFFu) | (1u << 16));
}
// instruction descriptor: bf16 x bf16 -> fp32, both operands K-major
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// byte offset of element (row, 16-byte chunk) inside a K-major swizzled tile whose rows are
// `row_bytes` (= swizzle span) long
__host__ __device__ inline uint32_t swz_off(uint32_t row, uint32_t chunk, uint32_t row_bytes) {
  uint32_t mask = row_bytes / 16 - 1;     // 128B: 7, 64B: 3, 32B: 1
  uint32_t grp = row_bytes == 128 ? (row & 7) : (row_bytes == 64 ? ((row >> 1) & 3) : ((row >> 2) & 1));
  return row * row_bytes + ((chunk ^ grp) & mask) * 16;
}

}  // namespace tc

// ------------------------------------------------------------------ host: tensor maps
// dims / strides innermost first; strides[i] = byte stride of dim i+1 (rank-1 entries).
// elem_bytes: 1 = uint8, 2 = bf16, 4 = fp32.  swizzle_bytes 0 = no swizzle.
int make_tensor_map(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box, int swizzle_bytes, int elem_bytes);

// im2col-mode map of an NHWC bf16 tensor (dims C,W,H,N): base pixels run over the bounding box
// [lower, extent-1+upper] per spatial dim with traversal stride (stride_w, stride_h); one load
// fetches `pixels` consecutive base pixels x `channels` channels.  Semantics probed on sm_100a:
// tools/probes/im2col_probe.cu.
int make_im2col_map(CUtensorMap* out, const void* base, int C, int W, int H, int N, int lower_w, int lower_h,
                    int upper_w, int upper_h, int channels, int pixels, int stride_w, int stride_h, int swizzle_bytes);
// same with explicit byte strides of the W / H / N dims (padded rows, overlapping pixels)
int make_im2col_map_strided(CUtensorMap* out, const void* base, int C, int W, int H, int N, uint64_t w_bytes, uint64_t h_bytes,
                            uint64_t n_bytes, int lower_w, int lower_h, int upper_w, int upper_h, int channels, int pixels,
                            int stride_w, int stride_h, int swizzle_bytes);

}  // namespace i2l
