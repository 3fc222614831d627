// Probe (GPU box): the floor of a persistent wide-decoder step whose weights do NOT fit on chip.  Every CTA streams its
// share of the step's weights L2 -> shared memory with TMA (16 KB k-blocks of a 128-row tile, 4..6-stage ring) and
// issues 128 x N x 16 SS MMAs against a resident B operand (N sequences x 512 K bf16).  No epilogue, no exchange:
// pure streaming + tensor time per "step" = TILES tiles x 8 k-blocks, on `grid` CTAs at once (L2 contention included).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../hmer-img2latex_b200/csrc -o stream_mma_probe \
//        stream_mma_probe.cu ../../hmer-img2latex_b200/csrc/tc_host.cu ../../hmer-img2latex_b200/csrc/api.cu -lcuda
#include "tc_common.cuh"
#include <stdio.h>
#include <stdlib.h>
using namespace i2l;
using namespace i2l::tc;

constexpr int STAGES = 6, KB_BYTES = 128 * 64 * 2;

template <int N>
__global__ void __launch_bounds__(128, 1) probe(const __grid_constant__ CUtensorMap tmw, int tiles_per_step, int steps,
                                                int rows_total) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5;
  constexpr int OFF_B = STAGES * KB_BYTES;                 // B operand: 8 k-blocks x (N rows x 128 B)
  constexpr int OFF_BAR = OFF_B + 8 * N * 128;
  const uint32_t bar = sbase + OFF_BAR;
  uint32_t* misc = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 8 * (2 * STAGES + 2));
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(bar + 8 * s, 1); mbar_init(bar + 8 * (STAGES + s), 1); }
    mbar_init(bar + 8 * 2 * STAGES, 1);
    mbar_fence_init();
  }
  for (int i = threadIdx.x; i < 8 * N * 128 / 16; i += 128) reinterpret_cast<uint4*>(smem + OFF_B)[i] = make_uint4(0x3c003c00, 0x3c003c00, 0, 0);
  if (warp == 1) tmem_alloc<256>(smem_u32(misc));
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = misc[0];
  const int n_kb = steps * tiles_per_step * 8;
  const int row_tiles = rows_total / 128;
  if (warp == 0) {
    if (elect_one()) {
      int stage = 0; uint32_t ph = 0;
      for (int i = 0; i < n_kb; ++i) {
        mbar_wait(bar + 8 * (STAGES + stage), ph ^ 1);
        mbar_arrive_expect_tx(bar + 8 * stage, KB_BYTES);
        const int tile = ((i >> 3) * gridDim.x + blockIdx.x) % row_tiles;     // every CTA walks its own tiles of the weight set
        tma_load_2d(sbase + stage * KB_BYTES, &tmw, (i & 7) * 64, tile * 128, bar + 8 * stage);
        if (++stage == STAGES) { stage = 0; ph ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    constexpr uint32_t IDESC = idesc_bf16(128, N);
    const uint64_t d0 = desc_base(sbase, 128);
    int stage = 0; uint32_t ph = 0;
    for (int i = 0; i < n_kb; ++i) {
      mbar_wait(bar + 8 * stage, ph);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t ad = d0 + (uint64_t)((stage * KB_BYTES + ks * 32) >> 4);
          const uint64_t bd = d0 + (uint64_t)((OFF_B + (i & 7) * N * 128 + ks * 32) >> 4);
          tc_mma_ss(tmem + ((i >> 3) & 1) * N, ad, bd, IDESC, (i & 7) | ks ? 1u : 0u);
        }
        tc_commit(bar + 8 * (STAGES + stage));
        if (i == n_kb - 1) tc_commit(bar + 8 * 2 * STAGES);
      }
      __syncwarp();
      if (++stage == STAGES) { stage = 0; ph ^= 1; }
    }
    mbar_wait(bar + 8 * 2 * STAGES, 0);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<256>(tmem);
}

template <int N>
void run(const CUtensorMap& tm, int grid, int tiles, int steps, int rows_total) {
  const int smem = STAGES * KB_BYTES + 8 * N * 128 + 256;
  cudaFuncSetAttribute(probe<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  probe<N><<<grid, 128, smem>>>(tm, tiles, 10, rows_total);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  probe<N><<<grid, 128, smem>>>(tm, tiles, steps, rows_total);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  const double us = ms * 1e3 / steps;
  printf("N=%3d grid=%3d tiles/step=%d: %7.2f us/step  (%.1f MB/step from L2 = %.2f TB/s, %d MMAs/step/CTA = %.0f cycles each @1.965GHz)  %s\n",
         N, grid, tiles, us, grid * tiles * 131072.0 / 1e6, grid * tiles * 131072.0 / us / 1e6, tiles * 32,
         us * 1965.0 / (tiles * 32), cudaGetErrorString(cudaGetLastError()));
}

int main() {
  const int rows_total = 6656 * 4;                  // 4 x the 512/512/2 weight set (rows of 512 bf16): 27 MB, L2 resident
  void* w; cudaMalloc(&w, (size_t)rows_total * 512 * 2); cudaMemset(w, 0, (size_t)rows_total * 512 * 2);
  CUtensorMap tm;
  uint64_t dims[2] = {512, (uint64_t)rows_total}; uint64_t str[1] = {1024}; uint32_t box[2] = {64, 128};
  if (make_tensor_map(&tm, w, 2, dims, str, box, 128, 2) != 0) { printf("tensor map failed: %s\n", i2l_last_error()); return 1; }
  for (int grid : {1, 64, 128, 148}) {
    run<64>(tm, grid, 7, 100, rows_total);
    run<128>(tm, grid, 7, 100, rows_total);
  }
  run<64>(tm, 128, 13, 100, rows_total);
  run<32>(tm, 128, 7, 100, rows_total);
  return 0;
}
