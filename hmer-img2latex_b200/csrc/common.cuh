// Shared helpers for the i2l_b200 library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/i2l_b200.h"

namespace i2l {

void set_error(const char* fmt, ...);
void count_launch();   // every kernel launch of the library is counted (bench.py "gpu_launches")
void count_launches(long long n);      // kernel nodes executed by one replay of a captured launch sequence
void set_launch_counting(bool on);     // thread-local: off while a sequence is being captured

// Optional per-kernel CUDA-event timing (bench.py roofline): enabled by i2l_prof_enable(1).
// Usage: { KernelTimer t("conv2", stream); kernel<<<...>>>(...); }
struct KernelTimer {
  int slot; cudaStream_t s;
  KernelTimer(const char* name, cudaStream_t stream);
  ~KernelTimer();
};

#define I2L_CUDA_OK(expr)                                                                  \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      i2l::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return I2L_ERR_CUDA;                                                                 \
    }                                                                                      \
  } while (0)

#define I2L_LAUNCH_OK()                                                                    \
  do {                                                                                     \
    i2l::count_launch();                                                                   \
    cudaError_t _e = cudaGetLastError();                                                   \
    if (_e != cudaSuccess) {                                                               \
      i2l::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return I2L_ERR_CUDA;                                                                 \
    }                                                                                      \
  } while (0)

#define I2L_REQUIRE(cond, ...)                                                             \
  do {                                                                                     \
    if (!(cond)) {                                                                         \
      i2l::set_error(__VA_ARGS__);                                                         \
      return I2L_ERR_INVALID;                                                              \
    }                                                                                      \
  } while (0)

#define I2L_TRY(expr)                                                                      \
  do {                                                                                     \
    int _s = (expr);                                                                       \
    if (_s != I2L_OK) return _s;                                                           \
  } while (0)

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// Bump allocator over a caller-owned buffer (workspace or packed weights).
struct Arena {
  char* base;
  size_t cap;
  size_t off;
  Arena(void* b, size_t c) : base(reinterpret_cast<char*>(b)), cap(c), off(0) {}
  template <typename T>
  T* take(size_t n) {
    off = align_up(off, 256);
    T* p = reinterpret_cast<T*>(base + off);
    off += n * sizeof(T);
    return p;
  }
  bool ok() const { return off <= cap; }
};

int device_check();
int num_sms();

// ---------------------------------------------------------------- fp32 GEMM
// C[M,N] = act( A1[M,K1] W1[N,K1]^T (+ A2[M,K2] W2[N,K2]^T) + bias[N]
//               + add_rows[M,N] + add_table[tab_idx[m], N] )
struct GemmF32 {
  const float* A1 = nullptr; int lda1 = 0; const float* W1 = nullptr; int ldw1 = 0; int K1 = 0;
  const float* A2 = nullptr; int lda2 = 0; const float* W2 = nullptr; int ldw2 = 0; int K2 = 0;
  const float* bias = nullptr;
  const float* add_rows = nullptr; int ld_add = 0;
  const float* add_table = nullptr; int ld_tab = 0; const int64_t* tab_idx = nullptr;
  float* C = nullptr; int ldc = 0;
  int M = 0, N = 0;
  int relu = 0;
  // deterministic split-K over K1 (A2 must be null): partial sums in `splitk_ws`
  int splitk = 1; float* splitk_ws = nullptr;
  // device flag: when non-null and *skip_flag != 0 the kernel exits at once (device-side loop exit)
  const int* skip_flag = nullptr;
};
int gemm_f32(const GemmF32& g, cudaStream_t s);
size_t gemm_f32_splitk_ws_bytes(int M, int N, int splitk);

// ---------------------------------------------------------------- fp32 conv (NCHW, implicit GEMM)
struct ConvF32 {
  const float* x = nullptr;      // (B,Ci,Hi,Wi)
  const float* w = nullptr;      // (Co,Ci,KH,KW)
  const float* bias = nullptr;   // (Co) or null
  const float* residual = nullptr;  // (B,Co,Ho,Wo) or null
  float* y = nullptr;            // (B,Co,Ho,Wo)
  int B = 0, Ci = 0, Hi = 0, Wi = 0, Co = 0, KH = 0, KW = 0, stride = 1, pad = 0;
  int relu = 0;
};
int conv2d_f32(const ConvF32& c, cudaStream_t s);
int maxpool2d_f32(const float* x, float* y, int B, int C, int Hi, int Wi, int k, int stride, int pad,
                  cudaStream_t s);
int global_avgpool_f32(const float* x, float* y, int B, int C, int HW, cudaStream_t s);

}  // namespace i2l
