// Probe (GPU box): semantics of cp.async.bulk.tensor.4d ... im2col on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -o im2col_probe im2col_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <vector>
typedef CUresult (*EncIm2col)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave,
                              CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void probe(const __grid_constant__ CUtensorMap tm, float* out, int c, int w, int h, int n, int offw, int offh, int npix) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bar;
  uint32_t sb = (uint32_t)__cvta_generic_to_shared(smem), b = (uint32_t)__cvta_generic_to_shared(&bar);
  for (int i = threadIdx.x; i < npix * 32; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = -7.f;
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(npix * 128) : "memory");
    unsigned short ow = (unsigned short)offw, oh = (unsigned short)offh;
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6], {%7, %8};" ::"r"(sb),
        "l"(&tm), "r"(c), "r"(w), "r"(h), "r"(n), "r"(b), "h"(ow), "h"(oh)
        : "memory");
    asm volatile(
        "{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(b) : "memory");
  }
  __syncthreads();
  for (int i = threadIdx.x; i < npix * 32; i += blockDim.x) out[i] = reinterpret_cast<float*>(smem)[i];
}
int main() {
  const int N = 3, H = 6, W = 10, C = 32;
  std::vector<float> x((size_t)N * H * W * C);
  for (int n = 0; n < N; ++n) for (int h = 0; h < H; ++h) for (int w = 0; w < W; ++w) for (int c = 0; c < C; ++c)
    x[(((size_t)n * H + h) * W + w) * C + c] = (float)((n + 1) * 1000000 + h * 10000 + w * 100 + c);
  float *dx, *dout; cudaMalloc(&dx, x.size() * 4); cudaMemcpy(dx, x.data(), x.size() * 4, cudaMemcpyHostToDevice);
  const int NPIX = 128; cudaMalloc(&dout, NPIX * 32 * 4);
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &p, cudaEnableDefault, &q);
  EncIm2col enc = (EncIm2col)p; if (!enc) { printf("no entry point\n"); return 1; }
  struct Case { const char* name; int lo[2], hi[2]; unsigned es[4]; int c, w, h, n, offw, offh; };
  Case cases[] = {
    {"3x3 s1 p1 tap(0,0) start(-1,-1,0)", {-1, -1}, {-1, -1}, {1, 1, 1, 1}, 0, -1, -1, 0, 0, 0},
    {"3x3 s1 p1 tap(kw=2,kh=1) start(-1,-1,0)", {-1, -1}, {-1, -1}, {1, 1, 1, 1}, 0, -1, -1, 0, 2, 1},
    {"3x3 s2 p1 tap(0,0) start(-1,-1,0)", {-1, -1}, {-1, -1}, {1, 2, 2, 1}, 0, -1, -1, 0, 0, 0},
    {"3x3 s2 p1 tap(1,1) start(w=3,h=1,n=1)", {-1, -1}, {-1, -1}, {1, 2, 2, 1}, 0, 3, 1, 1, 1, 1},
    {"1x3(kw=3) s1 pw=1 ph=0 lo{-1,0} hi{-1,0} tap(0,0) start(-1,0,0)", {-1, 0}, {-1, 0}, {1, 1, 1, 1}, 0, -1, 0, 0, 0, 0},
    {"1x1 s2 start(0,0,2) c=0", {0, 0}, {0, 0}, {1, 2, 2, 1}, 0, 0, 0, 2, 0, 0},
  };
  for (auto& cs : cases) {
    CUtensorMap tm;
    cuuint64_t gd[4] = {C, W, H, N}; cuuint64_t gs[3] = {C * 4ull, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, dx, gd, gs, cs.lo, cs.hi, 32, NPIX, cs.es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("== %s : encode rc=%d\n", cs.name, (int)r);
    if (r) continue;
    cudaMemset(dout, 0, NPIX * 32 * 4);
    probe<<<1, 128, NPIX * 128>>>(tm, dout, cs.c, cs.w, cs.h, cs.n, cs.offw, cs.offh, NPIX);
    cudaError_t e = cudaDeviceSynchronize();
    if (e) { printf("  kernel error %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<float> o(NPIX * 32); cudaMemcpy(o.data(), dout, o.size() * 4, cudaMemcpyDeviceToHost);
    for (int i = 0; i < NPIX; ++i) {
      int v = (int)o[i * 32 + 1];   // channel 1
      if (v == 0) printf(" .");
      else if (v < 0) printf(" ?");
      else printf(" n%dh%dw%d", v / 1000000 - 1, (v / 10000) % 100, (v / 100) % 100);
      if (i % 10 == 9) printf("\n");
    }
    printf("\n");
  }
  return 0;
}
