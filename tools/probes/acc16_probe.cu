// Probe (GPU box): (1) tensor-memory layout of an FP16 accumulator (tcgen05.mma kind::f16 with D = f16): is it one value
// per 32-bit column or two?  (2) tcgen05.ld throughput: cycles for 4 / 8 warps to read 128 fp32 columns of all 128 lanes
// (64 KB) -- the figure conv1's epilogue is bounded by (conv1).
// nvcc -gencode arch=compute_100a,code=sm_100a -o acc16_probe acc16_probe.cu
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdio.h>
#include <stdint.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc128(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void ld16(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr) : "memory");
}
template <int CFMT>   // 0: f16 accumulator, 1: f32
__global__ void __launch_bounds__(256, 1) probe(uint32_t* out, long long* cyc) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t tbase;
  const uint32_t sb = smem_u32(smem), b = smem_u32(&bar);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // A [128][64] fp16 (only k = 0 non-zero: m + 1), B [32][64] fp16 (k = 0: (n + 1) / 16), K-major SWIZZLE_128B
  for (int i = tid; i < (16384 + 4096) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
  __syncthreads();
  if (tid < 128) *reinterpret_cast<__half*>(smem + tid * 128 + ((0 ^ (tid & 7)) << 4)) = __float2half((float)(tid + 1));
  if (tid < 32) *reinterpret_cast<__half*>(smem + 16384 + tid * 128 + ((0 ^ (tid & 7)) << 4)) = __float2half((float)(tid + 1) / 16.f);
  if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b)); asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tbase)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tbase;
  constexpr uint32_t IDESC = ((uint32_t)CFMT << 4) | ((uint32_t)(32 >> 3) << 17) | ((128u >> 4) << 24);   // f16 x f16, N = 32, M = 128
  if (tid == 0) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(desc128(sb)), "l"(desc128(sb + 16384)), "r"(IDESC), "r"(0u) : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(b) : "memory");
  }
  asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(b) : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp < 4) {
    uint32_t r[16];
    for (int c = 0; c < 2; ++c) {
      ld16(tmem + ((uint32_t)(32 * warp) << 16) + c * 16, r);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int i = 0; i < 16; ++i) out[(32 * warp + lane) * 32 + c * 16 + i] = r[i];
    }
  }
  __syncthreads();
  // ---- tcgen05.ld throughput: NW warps (quadrant = warp & 3) each read 128 columns R times
  for (int nw = 4; nw <= 8; nw += 4) {
    __syncthreads();
    long long t0 = clock64();
    if (warp < nw) {
      uint32_t r[16], acc = 0;
      for (int rep = 0; rep < 64; ++rep) {
#pragma unroll
        for (int c = 0; c < (nw == 4 ? 8 : 4); ++c) {
          ld16(tmem + ((uint32_t)(32 * (warp & 3)) << 16) + ((warp >> 2) * 64 + c * 16), r);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          acc ^= r[0] ^ r[15];
        }
      }
      if (acc == 0x12345678u) out[0] = acc;
    }
    __syncthreads();
    long long t1 = clock64();
    if (tid == 0) cyc[nw / 4 - 1] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}
template <int CFMT>
void run(const char* nm) {
  uint32_t* d; long long* c;
  cudaMalloc(&d, 128 * 32 * 4); cudaMalloc(&c, 16);
  cudaMemset(d, 0xff, 128 * 32 * 4);
  const int smem = 16384 + 4096 + 1024;
  cudaFuncSetAttribute(probe<CFMT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe<CFMT><<<1, 256, smem>>>(d, c);
  cudaError_t e = cudaDeviceSynchronize();
  if (e) { printf("%s: %s\n", nm, cudaGetErrorString(e)); return; }
  static uint32_t h[128 * 32]; long long hc[2];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost); cudaMemcpy(hc, c, 16, cudaMemcpyDeviceToHost);
  printf("== %s accumulator: expected D[m][n] = (m+1)(n+1)/16\n", nm);
  for (int m : {0, 1, 37}) {
    printf(" lane %3d raw words:", m);
    for (int i = 0; i < 8; ++i) printf(" %08x", h[m * 32 + i]);
    printf("\n   as fp32:");
    for (int i = 0; i < 6; ++i) printf(" %g", *reinterpret_cast<float*>(&h[m * 32 + i]));
    printf("\n   as half pairs (lo,hi):");
    for (int i = 0; i < 6; ++i) { __half2 v = *reinterpret_cast<__half2*>(&h[m * 32 + i]); printf(" (%g,%g)", __low2float(v), __high2float(v)); }
    printf("\n");
  }
  printf(" tcgen05.ld of 128 lanes x 128 fp32 columns (64 KB), 64 repetitions: 4 warps %.0f cycles / rep, 8 warps %.0f cycles / rep\n",
         hc[0] / 64.0, hc[1] / 64.0);
}
int main() { run<1>("fp32"); run<0>("fp16"); return 0; }
