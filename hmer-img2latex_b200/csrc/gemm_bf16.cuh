// bf16 tcgen05 GEMM with the fused epilogue of the decode-step products (gemm_bf16.cu).
#pragma once
#include "tc_common.cuh"

namespace i2l {

struct GemmBf16 {
  CUtensorMap tmA1, tmW1;          // A1 (M,K1) / W1 (N,K1) bf16, from gemm_bf16_a_map / gemm_bf16_w_map
  CUtensorMap tmA2, tmW2;          // optional second product (K2 > 0)
  int K1 = 0, K2 = 0;
  const float* bias = nullptr;
  const float* add_rows = nullptr; int ld_add = 0;
  const float* add_table = nullptr; int ld_tab = 0; const int64_t* tab_idx = nullptr;
  float* C = nullptr; int ldc = 0;
  int M = 0, N = 0;
  int relu = 0;
  const int* skip_flag = nullptr;  // device flag: != 0 => the kernel exits at once
};
int gemm_bf16_a_map(CUtensorMap* out, const void* a, int M, int K, int lda);
int gemm_bf16_w_map(CUtensorMap* out, const void* w, int N, int K, int ldw);
int gemm_bf16(const GemmBf16& g, cudaStream_t s);

}  // namespace i2l
