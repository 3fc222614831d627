// Cycle probe of the warp-level sampling selection (csrc/sample_select.cuh): 8 warps per SM (the occupancy it has inside
// the persistent kernel), each warp runs the routine REP times on its own 512 logits.  Prints cycles per call.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I hmer-img2latex_b200/csrc -I include -o tools/probes/select_probe tools/probes/select_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "i2l_b200.h"
#include "sample_select.cuh"
using namespace i2l;

__global__ void __launch_bounds__(256) probe(const float* logits, int V, float temperature, int top_k, float top_p, int do_sample,
                                             int rep, long long* cycles, int* sink) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* x = logits + (size_t)(blockIdx.x * 8 + warp) * 512;
  float lg[16];
  for (int i = 0; i < 16; ++i) lg[i] = x[16 * lane + i];
  int acc = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int r = 0; r < rep; ++r) {
    float u = philox_uniform(7, (unsigned long long)r * 4096 + blockIdx.x * 8 + warp);
    acc += warp_sample_select(lg, V, lane, temperature, top_k, top_p, do_sample, u, nullptr);
    lg[r & 15] += 1e-3f * (float)(acc & 3);
  }
  long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (lane == 0) sink[blockIdx.x * 8 + warp] = acc;
}

int main(int argc, char** argv) {
  const int nb = 148, rep = 50;
  float* h = (float*)malloc(nb * 8 * 512 * 4);
  srand(1);
  for (int i = 0; i < nb * 8 * 512; ++i) {
    float a = 0; for (int k = 0; k < 12; ++k) a += rand() / (float)RAND_MAX; h[i] = (a - 6.f) * (argc > 1 ? atof(argv[1]) : 3.f);
  }
  float* d; long long* cyc; int* sink;
  cudaMalloc(&d, nb * 8 * 512 * 4); cudaMalloc(&cyc, nb * 8); cudaMalloc(&sink, nb * 8 * 4);
  cudaMemcpy(d, h, nb * 8 * 512 * 4, cudaMemcpyHostToDevice);
  struct { float t; int k; float p; int s; const char* name; } cases[] = {
      {0.8f, 50, 0.9f, 1, "T=0.8 k=50 p=0.9"}, {1.0f, 0, 0.9f, 1, "k=0 p=0.9 (full sort)"}, {1.0f, 200, 0.0f, 1, "k=200 p=0"},
      {1.0f, 50, 0.0f, 1, "k=50 p=0"}, {0.7f, 0, 0.0f, 0, "argmax(probs)"},
      {1.0f, 0, 0.0f, 1, "softmax + draw only"}, {1.0f, 50, 0.0f, 0, "k=50 p=0, no draw"}, {1.0f, 200, 0.0f, 0, "k=200 p=0, no draw"},
      {1.0f, 50, 0.9f, 0, "k=50 p=0.9, no draw"}};
  for (auto& c : cases) {
    probe<<<nb, 256>>>(d, 512, c.t, c.k, c.p, c.s, rep, cyc, sink);
    probe<<<nb, 256>>>(d, 512, c.t, c.k, c.p, c.s, rep, cyc, sink);
    cudaError_t e = cudaDeviceSynchronize();
    long long hc[148];
    cudaMemcpy(hc, cyc, sizeof(hc), cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < nb; ++i) mx = hc[i] > mx ? hc[i] : mx;
    printf("%-24s %8.0f cycles per call (8 warps/SM)  [%s]\n", c.name, (double)mx / rep, cudaGetErrorString(e));
  }
  return 0;
}
