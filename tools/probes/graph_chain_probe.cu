// Probe (GPU box): what does one dependent kernel cost inside a replayed CUDA graph?  Chains of 300 launches:
//   trivial         1 CTA x 32 threads, no shared memory
//   smem128k        + 128 KB dynamic shared memory (carve-out change against a small kernel in between)
//   tmem            + tcgen05.alloc / dealloc of 128 columns
//   mixed           alternating smem128k / trivial (what the decode loop does: GEMM, select, GEMM, ...)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o graph_chain_probe graph_chain_probe.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

__global__ void k_trivial(int* p) { if (threadIdx.x == 0 && p[0] == 12345) p[1] = 1; }
__global__ void k_smem(int* p) {
  extern __shared__ int sm[];
  if (threadIdx.x == 0) { sm[0] = p[0]; if (sm[0] == 12345) p[1] = 1; }
}
__global__ void __launch_bounds__(128, 1) k_tmem(int* p) {
  extern __shared__ int sm[];
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(sm);
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(s) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  __syncthreads();
  const uint32_t t = sm[0];
  if (threadIdx.x == 0 && p[0] == 12345) p[1] = 1;
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(t) : "memory");
}

template <class F>
static void run(const char* name, F enqueue, int n, int grid_note) {
  cudaStream_t s; cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
  cudaGraph_t g; cudaGraphExec_t ge;
  cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
  for (int i = 0; i < n; ++i) enqueue(s, i);
  cudaStreamEndCapture(s, &g);
  cudaGraphInstantiate(&ge, g, 0);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int w = 0; w < 3; ++w) cudaGraphLaunch(ge, s);
  cudaStreamSynchronize(s);
  cudaEventRecord(e0, s);
  for (int r = 0; r < 10; ++r) cudaGraphLaunch(ge, s);
  cudaEventRecord(e1, s);
  cudaStreamSynchronize(s);
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  printf("%-28s graph : %7.2f us per kernel (grid %d)\n", name, ms / 10 / n * 1e3, grid_note);
  cudaEventRecord(e0, s);
  for (int r = 0; r < 3; ++r) for (int i = 0; i < n; ++i) enqueue(s, i);
  cudaEventRecord(e1, s);
  cudaStreamSynchronize(s);
  cudaEventElapsedTime(&ms, e0, e1);
  printf("%-28s stream: %7.2f us per kernel\n", name, ms / 3 / n * 1e3);
  printf("  last error: %s\n", cudaGetErrorString(cudaGetLastError()));
}

int main() {
  int* p; cudaMalloc(&p, 64); cudaMemset(p, 0, 64);
  cudaFuncSetAttribute(k_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072);
  cudaFuncSetAttribute(k_tmem, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072);
  const int n = 300;
  for (int grid : {1, 128}) {
    run("trivial", [&](cudaStream_t s, int) { k_trivial<<<grid, 32, 0, s>>>(p); }, n, grid);
    run("smem128k", [&](cudaStream_t s, int) { k_smem<<<grid, 128, 131072, s>>>(p); }, n, grid);
    run("tmem+smem128k", [&](cudaStream_t s, int) { k_tmem<<<grid, 128, 131072, s>>>(p); }, n, grid);
    run("tmem+smem1k", [&](cudaStream_t s, int) { k_tmem<<<grid, 128, 1024, s>>>(p); }, n, grid);
    run("mixed smem128k/trivial", [&](cudaStream_t s, int i) { if (i & 1) k_trivial<<<grid, 256, 0, s>>>(p); else k_smem<<<grid, 128, 131072, s>>>(p); }, n, grid);
    run("mixed tmem128k/trivial", [&](cudaStream_t s, int i) { if (i & 1) k_trivial<<<grid, 256, 0, s>>>(p); else k_tmem<<<grid, 128, 131072, s>>>(p); }, n, grid);
  }
  return 0;
}
