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
  // ---- fused LSTM cell epilogue (cell_c != nullptr): N = 4H, the weight rows of every 128-row N tile are ordered
  // [gate i | f | g | o] x [32 consecutive hidden units] (PackedDec::g16c_*), so one tile holds all four gates of its 32
  // units.  bias / add_rows / add_table keep PyTorch's gate-major columns (gate * H + unit).  C is not written: the
  // epilogue computes c' = sig(f) c + sig(i) tanh(g), h' = sig(o) tanh(c') (decoder.py:76-82 / nn.LSTM) and stores
  // c' in place, h' as fp32 and as bf16 (cell_hb: the A operand of the NEXT product; a different buffer than the one
  // this launch reads through tmA*, other CTAs are still reading that one).
  float* cell_c = nullptr; float* cell_h = nullptr; __nv_bfloat16* cell_hb = nullptr; int cell_H = 0;
};
int gemm_bf16_a_map(CUtensorMap* out, const void* a, int M, int K, int lda);
int gemm_bf16_w_map(CUtensorMap* out, const void* w, int N, int K, int ldw);
int gemm_bf16(const GemmBf16& g, cudaStream_t s);

}  // namespace i2l
