// fp32 reference-precision kernels (CUDA-core FFMA): tiled GEMM and implicit-GEMM
// convolution.  These back precision=I2L_FP32 (the mode whose logits must agree with
// the reference within 1e-3 relative) and every shape the tcgen05 fast path does not
// cover.  64x64x16 tiles, 256 threads, 4x4 register micro-tile.
#include "common.cuh"

namespace i2l {

namespace {

constexpr int BM = 64, BN = 64, BK = 16, NT = 256;

struct ConvGeom {
  int Ci, Hi, Wi, Co, KH, KW, stride, pad, Ho, Wo;
};

template <bool CONV>
__global__ void __launch_bounds__(NT) tile_kernel(GemmF32 g, ConvGeom cg, const float* __restrict__ residual) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Ws[BK][BN + 4];
  if (g.skip_flag != nullptr && *g.skip_flag != 0) return;
  const int t = threadIdx.x;
  const int tx = t % 16, ty = t / 16;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  float acc[4][4] = {};

  // conv: this thread always loads A column mm = t % 64
  int cb = 0, chi0 = 0, cwi0 = 0;
  bool cm_ok = false;
  if (CONV) {
    int m = m0 + (t % BM);
    cm_ok = m < g.M;
    int wo = m % cg.Wo;
    int r = m / cg.Wo;
    int ho = r % cg.Ho;
    cb = r / cg.Ho;
    chi0 = ho * cg.stride - cg.pad;
    cwi0 = wo * cg.stride - cg.pad;
  }

  for (int pair = 0; pair < 2; ++pair) {
    const float* __restrict__ A = pair == 0 ? g.A1 : g.A2;
    const float* __restrict__ W = pair == 0 ? g.W1 : g.W2;
    const int lda = pair == 0 ? g.lda1 : g.lda2;
    const int ldw = pair == 0 ? g.ldw1 : g.ldw2;
    const int K = pair == 0 ? g.K1 : g.K2;
    if (K == 0 || W == nullptr) continue;
    int kbeg = 0, kend = K;
    if (g.splitk > 1) {
      int chunk = ((K + BK - 1) / BK + g.splitk - 1) / g.splitk * BK;
      kbeg = blockIdx.z * chunk;
      kend = min(K, kbeg + chunk);
    }
    for (int k0 = kbeg; k0 < kend; k0 += BK) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int e = t + NT * j;
        if (CONV) {
          int mm = e % BM, kk = e / BM;
          int k = k0 + kk;
          float v = 0.f;
          if (cm_ok && k < kend) {
            int ci = k / (cg.KH * cg.KW);
            int r = k - ci * (cg.KH * cg.KW);
            int kh = r / cg.KW, kw = r - kh * cg.KW;
            int hi = chi0 + kh, wi = cwi0 + kw;
            if (hi >= 0 && hi < cg.Hi && wi >= 0 && wi < cg.Wi)
              v = A[((size_t)(cb * cg.Ci + ci) * cg.Hi + hi) * cg.Wi + wi];
          }
          As[kk][mm] = v;
        } else {
          int kk = e % BK, mm = e / BK;
          int m = m0 + mm, k = k0 + kk;
          As[kk][mm] = (m < g.M && k < kend) ? A[(size_t)m * lda + k] : 0.f;
        }
        {
          int kk = e % BK, nn = e / BK;
          int n = n0 + nn, k = k0 + kk;
          Ws[kk][nn] = (n < g.N && k < kend) ? W[(size_t)n * ldw + k] : 0.f;
        }
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
        float4 b4 = *reinterpret_cast<const float4*>(&Ws[kk][tx * 4]);
        float a[4] = {a4.x, a4.y, a4.z, a4.w};
        float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + ty * 4 + i;
    if (m >= g.M) continue;
    size_t conv_base = 0;
    int HoWo = 0;
    if (CONV) {
      int wo = m % cg.Wo;
      int r = m / cg.Wo;
      int ho = r % cg.Ho;
      int b = r / cg.Ho;
      HoWo = cg.Ho * cg.Wo;
      conv_base = (size_t)b * cg.Co * HoWo + (size_t)ho * cg.Wo + wo;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n >= g.N) continue;
      float v = acc[i][j];
      if (g.splitk > 1) {
        g.splitk_ws[((size_t)blockIdx.z * g.M + m) * g.N + n] = v;
        continue;
      }
      if (g.bias) v += g.bias[n];
      if (CONV) {
        size_t o = conv_base + (size_t)n * HoWo;
        if (residual) v += residual[o];
        if (g.relu) v = fmaxf(v, 0.f);
        g.C[o] = v;
      } else {
        if (g.add_rows) v += g.add_rows[(size_t)m * g.ld_add + n];
        if (g.add_table) v += g.add_table[(size_t)g.tab_idx[m] * g.ld_tab + n];
        if (g.relu) v = fmaxf(v, 0.f);
        g.C[(size_t)m * g.ldc + n] = v;
      }
    }
  }
}

__global__ void splitk_reduce_kernel(GemmF32 g) {
  if (g.skip_flag != nullptr && *g.skip_flag != 0) return;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t total = (size_t)g.M * g.N;
  if (i >= total) return;
  int m = (int)(i / g.N), n = (int)(i % g.N);
  float v = 0.f;
  for (int z = 0; z < g.splitk; ++z) v += g.splitk_ws[(size_t)z * total + i];
  if (g.bias) v += g.bias[n];
  if (g.add_rows) v += g.add_rows[(size_t)m * g.ld_add + n];
  if (g.add_table) v += g.add_table[(size_t)g.tab_idx[m] * g.ld_tab + n];
  if (g.relu) v = fmaxf(v, 0.f);
  g.C[(size_t)m * g.ldc + n] = v;
}

__global__ void maxpool_kernel(const float* __restrict__ x, float* __restrict__ y, int C, int Hi, int Wi,
                               int Ho, int Wo, int k, int stride, int pad, size_t total) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int wo = (int)(i % Wo);
  size_t r = i / Wo;
  int ho = (int)(r % Ho);
  size_t bc = r / Ho;
  const float* xp = x + bc * (size_t)Hi * Wi;
  float m = -INFINITY;
  for (int a = 0; a < k; ++a) {
    int hi = ho * stride - pad + a;
    if (hi < 0 || hi >= Hi) continue;
    for (int b = 0; b < k; ++b) {
      int wi = wo * stride - pad + b;
      if (wi < 0 || wi >= Wi) continue;
      m = fmaxf(m, xp[(size_t)hi * Wi + wi]);
    }
  }
  y[i] = m;
}

__global__ void avgpool_kernel(const float* __restrict__ x, float* __restrict__ y, int HW, int total) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) / 32;
  int lane = threadIdx.x % 32;
  if (warp >= total) return;
  const float* xp = x + (size_t)warp * HW;
  float s = 0.f;
  for (int i = lane; i < HW; i += 32) s += xp[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) y[warp] = s / (float)HW;
}

}  // namespace

size_t gemm_f32_splitk_ws_bytes(int M, int N, int splitk) {
  return splitk > 1 ? (size_t)splitk * M * N * sizeof(float) : 0;
}

int gemm_f32(const GemmF32& g, cudaStream_t s) {
  if (g.M <= 0 || g.N <= 0) return I2L_OK;
  I2L_REQUIRE(g.splitk >= 1 && (g.splitk == 1 || (g.A2 == nullptr && g.splitk_ws != nullptr)),
              "gemm_f32: invalid split-K configuration");
  dim3 grid(cdiv(g.M, BM), cdiv(g.N, BN), g.splitk);
  ConvGeom cg{};
  tile_kernel<false><<<grid, NT, 0, s>>>(g, cg, nullptr);
  I2L_LAUNCH_OK();
  if (g.splitk > 1) {
    size_t total = (size_t)g.M * g.N;
    splitk_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(g);
    I2L_LAUNCH_OK();
  }
  return I2L_OK;
}

int conv2d_f32(const ConvF32& c, cudaStream_t s) {
  ConvGeom cg{c.Ci, c.Hi, c.Wi, c.Co, c.KH, c.KW, c.stride, c.pad,
              (c.Hi + 2 * c.pad - c.KH) / c.stride + 1, (c.Wi + 2 * c.pad - c.KW) / c.stride + 1};
  GemmF32 g;
  g.A1 = c.x; g.W1 = c.w; g.ldw1 = c.Ci * c.KH * c.KW; g.K1 = c.Ci * c.KH * c.KW;
  g.bias = c.bias; g.C = c.y; g.M = c.B * cg.Ho * cg.Wo; g.N = c.Co; g.relu = c.relu;
  if (g.M <= 0) return I2L_OK;
  dim3 grid(cdiv(g.M, BM), cdiv(g.N, BN), 1);
  tile_kernel<true><<<grid, NT, 0, s>>>(g, cg, c.residual);
  I2L_LAUNCH_OK();
  return I2L_OK;
}

int maxpool2d_f32(const float* x, float* y, int B, int C, int Hi, int Wi, int k, int stride, int pad,
                  cudaStream_t s) {
  int Ho = (Hi + 2 * pad - k) / stride + 1, Wo = (Wi + 2 * pad - k) / stride + 1;
  size_t total = (size_t)B * C * Ho * Wo;
  if (total == 0) return I2L_OK;
  maxpool_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(x, y, C, Hi, Wi, Ho, Wo, k, stride, pad, total);
  I2L_LAUNCH_OK();
  return I2L_OK;
}

int global_avgpool_f32(const float* x, float* y, int B, int C, int HW, cudaStream_t s) {
  int total = B * C;
  if (total == 0) return I2L_OK;
  avgpool_kernel<<<cdiv(total * 32, 256), 256, 0, s>>>(x, y, HW, total);
  I2L_LAUNCH_OK();
  return I2L_OK;
}

}  // namespace i2l
