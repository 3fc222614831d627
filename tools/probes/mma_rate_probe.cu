// Probe (GPU box): issue-to-completion cost of back-to-back tcgen05.mma kind::f16 (bf16, M = 128, K = 16) as a
// function of N, with the A operand in shared memory (SS) or in tensor memory (TS).  One CTA per SM on every SM.
// nvcc -gencode arch=compute_100a,code=sm_100a -o mma_rate_probe mma_rate_probe.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc128(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
template <int N, bool TS, int ACC = 1>   // ACC: number of accumulators the MMAs rotate over (1 = one dependent chain)
__global__ void __launch_bounds__(128, 1) probe(long long* out, int L) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t tbase;
  const uint32_t sb = smem_u32(smem), b = smem_u32(&bar);
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b)); asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tbase)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tbase;
  constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
  if (threadIdx.x == 0) {
    const uint64_t ad = desc128(sb), bd = desc128(sb + 16384);
    long long t0 = clock64();
    for (int i = 0; i < L; ++i) {
      // 4 K-steps of one 64-wide K block, like the conv kernels
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        if (TS)
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem + (uint32_t)(((i * 4 + ks) % ACC) * N)), "r"(tmem + 256 + ks * 8), "l"(bd + (uint64_t)((ks * 32) >> 4)), "r"(IDESC), "r"(1u) : "memory");
        else
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(ad + (uint64_t)((ks * 32) >> 4)), "l"(bd + (uint64_t)((ks * 32) >> 4)), "r"(IDESC), "r"(1u) : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(b) : "memory");
    asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(b) : "memory");
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}
template <int N, bool TS, int ACC = 1>
void run(long long* d, const char* nm) {
  const int smem = 16384 + 32768 + 1024;
  cudaFuncSetAttribute(probe<N, TS, ACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int grid : {1, 148}) {
    long long r[2];
    for (int rep = 0; rep < 2; ++rep) {
      const int L = 2000;
      probe<N, TS, ACC><<<grid, 128, smem>>>(d, L);
      cudaError_t e = cudaDeviceSynchronize();
      if (e) { printf("%s: %s\n", nm, cudaGetErrorString(e)); return; }
      cudaMemcpy(&r[rep], d, 8, cudaMemcpyDeviceToHost);
    }
    printf("%-4s N=%3d acc=%d grid=%3d: %7.1f cycles per 128xNx16 MMA  (%6.0f FLOP/clk/SM)\n", nm, N, ACC, grid, r[1] / 8000.0, 2.0 * 128 * N * 16 / (r[1] / 8000.0));
  }
}
int main() {
  long long* d; cudaMalloc(&d, 64);
  run<32, false>(d, "SS"); run<64, false>(d, "SS"); run<128, false>(d, "SS"); run<256, false>(d, "SS");
  run<32, true>(d, "TS"); run<64, true>(d, "TS"); run<128, true>(d, "TS"); run<256, true>(d, "TS");
  // independent accumulators: does the ~45-cycle cost of a small-N MMA come from the accumulator dependency?
  run<32, true, 2>(d, "TS"); run<32, true, 4>(d, "TS"); run<64, true, 2>(d, "TS"); run<32, false, 2>(d, "SS"); run<32, false, 4>(d, "SS");
  return 0;
}
