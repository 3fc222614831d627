// Encoder entry points: CNNEncoder.forward (model/encoder.py:111-129) and
// ResNetEncoder.forward (model/encoder.py:231-249).  fp32 path = NCHW implicit-GEMM
// convs on CUDA cores; the bf16 CNN path (tcgen05 implicit GEMM, NHWC) is in cnn_bf16.cu.
#include "common.cuh"
#include "encoder_bf16.cuh"
#include "resnet_net.cuh"
#include "resnet_bf16.cuh"
#include <vector>

namespace i2l {
namespace {

// ---------------------------------------------------------------- CNN
struct CnnLayout {
  size_t conv_w[I2L_MAX_CONV], conv_b[I2L_MAX_CONV], fc_w, fc_b, end_f32;
  int ci[I2L_MAX_CONV], h[I2L_MAX_CONV + 1], w[I2L_MAX_CONV + 1];   // input geometry of layer i
  size_t flat;
  size_t bf16_section, total_bytes;
};

int cnn_check(const i2l_cnn_desc* d) {
  I2L_REQUIRE(d != nullptr, "cnn: null descriptor");
  I2L_REQUIRE(d->n_conv >= 1 && d->n_conv <= I2L_MAX_CONV, "cnn: n_conv out of range");
  I2L_REQUIRE(d->kernel_size >= 1 && (d->kernel_size & 1), "cnn: kernel_size must be odd");
  I2L_REQUIRE(d->pool_size >= 1 && d->img_height > 0 && d->img_width > 0 && d->channels > 0 && d->embedding_dim > 0,
              "cnn: invalid geometry");
  return I2L_OK;
}

CnnLayout cnn_layout(const i2l_cnn_desc& d) {
  CnnLayout L{};
  size_t o = 0;
  auto take = [&](size_t n) { size_t r = o; o += (n + 63) / 64 * 64; return r; };
  int ci = d.channels, h = d.img_height, w = d.img_width;
  for (int i = 0; i < d.n_conv; ++i) {
    L.ci[i] = ci; L.h[i] = h; L.w[i] = w;
    L.conv_w[i] = take((size_t)d.filters[i] * ci * d.kernel_size * d.kernel_size);
    L.conv_b[i] = take(d.filters[i]);
    ci = d.filters[i];
    h /= d.pool_size; w /= d.pool_size;            // MaxPool2d floor semantics (encoder.py:91)
  }
  L.h[d.n_conv] = h; L.w[d.n_conv] = w;
  L.flat = (size_t)ci * h * w;
  L.fc_w = take((size_t)d.embedding_dim * L.flat);
  L.fc_b = take(d.embedding_dim);
  L.end_f32 = o;
  size_t bytes = o * 4;
  L.bf16_section = 0;
  if (d.precision == I2L_BF16 && cnn_bf16_supported(d)) {
    bytes = align_up(bytes, 1024);
    L.bf16_section = bytes;
    bytes += cnn_bf16_packed_bytes(d);
  }
  L.total_bytes = bytes;
  return L;
}

int fc_splitk(int M, int N, int K) {
  int tiles = cdiv(M, 64) * cdiv(N, 64);
  int sk = cdiv(2 * num_sms(), tiles);
  int maxk = K / 512;
  if (sk > maxk) sk = maxk;
  if (sk > 64) sk = 64;
  if (sk < 1) sk = 1;
  return sk;
}

struct CnnWs { float* a; float* b[2]; float* sk; size_t bytes; int splitk; };
CnnWs cnn_carve(const i2l_cnn_desc& d, const CnnLayout& L, int B, void* ws) {
  Arena ar(ws, (size_t)-1);
  size_t max_conv = 0, max_pool = 0;
  for (int i = 0; i < d.n_conv; ++i) {
    max_conv = std::max(max_conv, (size_t)d.filters[i] * L.h[i] * L.w[i]);
    max_pool = std::max(max_pool, (size_t)d.filters[i] * L.h[i + 1] * L.w[i + 1]);
  }
  CnnWs w{};
  w.a = ar.take<float>((size_t)B * max_conv);
  w.b[0] = ar.take<float>((size_t)B * max_pool);
  w.b[1] = ar.take<float>((size_t)B * max_pool);
  w.splitk = fc_splitk(B, d.embedding_dim, (int)L.flat);
  w.sk = ar.take<float>(gemm_f32_splitk_ws_bytes(B, d.embedding_dim, w.splitk) / 4);
  w.bytes = align_up(ar.off, 256);
  return w;
}

// ---------------------------------------------------------------- ResNet (topology: resnet_net.cuh)
__global__ void fold_bn_kernel(const float* __restrict__ w, const float* __restrict__ g, const float* __restrict__ b,
                               const float* __restrict__ m, const float* __restrict__ v, float* __restrict__ wo,
                               float* __restrict__ bo, int co, int per) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)co * per) return;
  int c = (int)(i / per);
  float sc = g[c] / sqrtf(v[c] + 1e-5f);
  wo[i] = w[i] * sc;
  if (i % per == 0) bo[c] = b[c] - m[c] * sc;
}

}  // namespace
}  // namespace i2l

using namespace i2l;

// ====================================================================== CNN C ABI
extern "C" int i2l_cnn_tensor_core_path(const i2l_cnn_desc* d) {
  return (d != nullptr && d->precision == I2L_BF16 && i2l::cnn_bf16_supported(*d)) ? 1 : 0;
}

extern "C" size_t i2l_cnn_packed_bytes(const i2l_cnn_desc* d) {
  if (cnn_check(d) != I2L_OK) return 0;
  return cnn_layout(*d).total_bytes;
}

extern "C" int i2l_cnn_pack(const i2l_cnn_desc* d, const i2l_cnn_params* p, void* packed, size_t packed_bytes,
                            void* stream) {
  I2L_TRY(device_check());
  I2L_TRY(cnn_check(d));
  I2L_REQUIRE(p && packed, "i2l_cnn_pack: null argument");
  CnnLayout L = cnn_layout(*d);
  if (packed_bytes < L.total_bytes) { set_error("i2l_cnn_pack: packed buffer too small"); return I2L_ERR_WORKSPACE; }
  cudaStream_t s = (cudaStream_t)stream;
  float* pk = reinterpret_cast<float*>(packed);
  const int ks = d->kernel_size;
  for (int i = 0; i < d->n_conv; ++i) {
    I2L_REQUIRE(p->conv_w[i] && p->conv_b[i], "i2l_cnn_pack: missing conv layer %d", i);
    I2L_CUDA_OK(cudaMemcpyAsync(pk + L.conv_w[i], p->conv_w[i], (size_t)d->filters[i] * L.ci[i] * ks * ks * 4,
                                cudaMemcpyDeviceToDevice, s));
    I2L_CUDA_OK(cudaMemcpyAsync(pk + L.conv_b[i], p->conv_b[i], (size_t)d->filters[i] * 4, cudaMemcpyDeviceToDevice, s));
  }
  I2L_REQUIRE(p->fc_w && p->fc_b, "i2l_cnn_pack: missing embedding layer");
  I2L_CUDA_OK(cudaMemcpyAsync(pk + L.fc_w, p->fc_w, (size_t)d->embedding_dim * L.flat * 4, cudaMemcpyDeviceToDevice, s));
  I2L_CUDA_OK(cudaMemcpyAsync(pk + L.fc_b, p->fc_b, (size_t)d->embedding_dim * 4, cudaMemcpyDeviceToDevice, s));
  if (L.bf16_section) I2L_TRY(cnn_bf16_pack(*d, *p, reinterpret_cast<char*>(packed) + L.bf16_section, s));
  return I2L_OK;
}

extern "C" size_t i2l_cnn_workspace_bytes(const i2l_cnn_desc* d, int32_t batch) {
  if (cnn_check(d) != I2L_OK || batch <= 0) return 0;
  CnnLayout L = cnn_layout(*d);
  if (L.bf16_section) return cnn_bf16_workspace_bytes(*d, batch);
  return cnn_carve(*d, L, batch, nullptr).bytes;
}

extern "C" int i2l_cnn_encoder_fwd(const i2l_cnn_desc* d, const void* packed, const float* x, int32_t batch,
                                   float* out, void* workspace, size_t workspace_bytes, void* stream) {
  return i2l_cnn_encoder_fwd_in(d, packed, x, I2L_IN_F32, batch, out, workspace, workspace_bytes, stream);
}

namespace i2l {
namespace {
struct NormArgs { float mean[4], stdv[4]; int mode; };

template <typename OutT>
__device__ __forceinline__ OutT norm_px(unsigned v, int c, const NormArgs& a) {
  float t = (float)v / 255.0f;                                        // data/utils.py:68, predictor.py:444
  float y = a.mode == I2L_NORM_PM1 ? t * 2.0f - 1.0f                  // utils.py:74 / predictor.py:446
                                   : (t - a.mean[c]) / a.stdv[c];     // utils.py:77-79
  if constexpr (sizeof(OutT) == 4) return y; else return __float2bfloat16(y);
}

// NCHW -> NCHW, 16 pixels per thread (requires H*W % 16 == 0)
template <typename OutT>
__global__ void normalize_u8_vec_kernel(const uint4* __restrict__ src, OutT* __restrict__ dst, size_t n16, int hw16,
                                        int C, NormArgs a) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n16) return;
  const int c = (int)((i / hw16) % C);
  uint4 v = __ldg(src + i);
  uint32_t w[4] = {v.x, v.y, v.z, v.w};
  OutT o[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) o[k] = norm_px<OutT>((w[k >> 2] >> (8 * (k & 3))) & 0xffu, c, a);
  uint4* d4 = reinterpret_cast<uint4*>(dst + i * 16);
  const uint4* o4 = reinterpret_cast<const uint4*>(o);
#pragma unroll
  for (int k = 0; k < (int)(16 * sizeof(OutT) / 16); ++k) d4[k] = o4[k];
}

template <typename OutT>
__global__ void normalize_u8_kernel(const uint8_t* __restrict__ src, OutT* __restrict__ dst, size_t n, int C, int HW,
                                    int nhwc, NormArgs a) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;   // index into dst (B,C,H,W)
  if (i >= n) return;
  const int c = (int)((i / HW) % C);
  size_t si = i;
  if (nhwc) { size_t b = i / ((size_t)C * HW); size_t p = i % HW; si = (b * HW + p) * C + c; }
  dst[i] = norm_px<OutT>(src[si], c, a);
}
}  // namespace
}  // namespace i2l

extern "C" int i2l_normalize_u8(const uint8_t* src, int32_t src_layout, int32_t batch, int32_t channels,
                                int32_t height, int32_t width, int32_t mode, const float* mean, const float* stdv,
                                void* dst, int32_t dst_dtype, void* stream) {
  I2L_TRY(device_check());
  I2L_REQUIRE(batch >= 0 && channels >= 1 && channels <= 4 && height > 0 && width > 0, "i2l_normalize_u8: invalid shape");
  I2L_REQUIRE(src_layout == 0 || src_layout == 1, "i2l_normalize_u8: src_layout must be 0 (NCHW) or 1 (NHWC)");
  I2L_REQUIRE(mode == I2L_NORM_PM1 || mode == I2L_NORM_MEANSTD, "i2l_normalize_u8: invalid mode");
  I2L_REQUIRE(dst_dtype == I2L_IN_F32 || dst_dtype == I2L_IN_BF16, "i2l_normalize_u8: invalid dst dtype");
  I2L_REQUIRE(mode == I2L_NORM_PM1 || (mean && stdv), "i2l_normalize_u8: mean/std required");
  if (batch == 0) return I2L_OK;
  I2L_REQUIRE(src && dst, "i2l_normalize_u8: null buffer");
  NormArgs a{};
  a.mode = mode;
  for (int c = 0; c < channels; ++c) { a.mean[c] = mean ? mean[c] : 0.f; a.stdv[c] = stdv ? stdv[c] : 1.f; }
  cudaStream_t s = (cudaStream_t)stream;
  const int HW = height * width;
  const size_t n = (size_t)batch * channels * HW;
  const bool vec = src_layout == 0 && HW % 16 == 0 && ((uintptr_t)src % 16 == 0) && ((uintptr_t)dst % 16 == 0);
  KernelTimer kt("pre.normalize_u8", s);
  if (vec) {
    const size_t n16 = n / 16;
    const unsigned grid = (unsigned)((n16 + 255) / 256);
    if (dst_dtype == I2L_IN_F32)
      normalize_u8_vec_kernel<float><<<grid, 256, 0, s>>>(reinterpret_cast<const uint4*>(src), reinterpret_cast<float*>(dst), n16, HW / 16, channels, a);
    else
      normalize_u8_vec_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(reinterpret_cast<const uint4*>(src), reinterpret_cast<__nv_bfloat16*>(dst), n16, HW / 16, channels, a);
  } else {
    const unsigned grid = (unsigned)((n + 255) / 256);
    if (dst_dtype == I2L_IN_F32)
      normalize_u8_kernel<float><<<grid, 256, 0, s>>>(src, reinterpret_cast<float*>(dst), n, channels, HW, src_layout, a);
    else
      normalize_u8_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(src, reinterpret_cast<__nv_bfloat16*>(dst), n, channels, HW, src_layout, a);
  }
  I2L_LAUNCH_OK();
  return I2L_OK;
}

extern "C" int i2l_cnn_encoder_fwd_in(const i2l_cnn_desc* d, const void* packed, const void* xin, int32_t x_dtype,
                                      int32_t batch, float* out, void* workspace, size_t workspace_bytes,
                                      void* stream) {
  I2L_TRY(device_check());
  I2L_TRY(cnn_check(d));
  I2L_REQUIRE(packed && batch >= 0, "i2l_cnn_encoder_fwd: invalid argument");
  I2L_REQUIRE(x_dtype == I2L_IN_F32 || x_dtype == I2L_IN_BF16, "i2l_cnn_encoder_fwd: invalid input dtype");
  if (batch == 0) return I2L_OK;
  I2L_REQUIRE(xin && out && workspace, "i2l_cnn_encoder_fwd: null buffer");
  cudaStream_t s = (cudaStream_t)stream;
  CnnLayout L = cnn_layout(*d);
  if (L.bf16_section)
    return cnn_bf16_fwd(*d, reinterpret_cast<const char*>(packed) + L.bf16_section, xin, x_dtype, batch, out, workspace,
                        workspace_bytes, s);
  if (x_dtype != I2L_IN_F32) {
    set_error("i2l_cnn_encoder_fwd: bf16 input needs precision == I2L_BF16 and a tcgen05 configuration (i2l_cnn_tensor_core_path)");
    return I2L_ERR_UNSUPPORTED;
  }
  const float* x = reinterpret_cast<const float*>(xin);
  CnnWs w = cnn_carve(*d, L, batch, workspace);
  if (workspace_bytes < w.bytes) { set_error("i2l_cnn_encoder_fwd: workspace too small (%zu < %zu)", workspace_bytes, w.bytes); return I2L_ERR_WORKSPACE; }
  const float* pk = reinterpret_cast<const float*>(packed);
  const float* cur = x;
  for (int i = 0; i < d->n_conv; ++i) {
    ConvF32 c;
    c.x = cur; c.w = pk + L.conv_w[i]; c.bias = pk + L.conv_b[i]; c.y = w.a;
    c.B = batch; c.Ci = L.ci[i]; c.Hi = L.h[i]; c.Wi = L.w[i]; c.Co = d->filters[i];
    c.KH = c.KW = d->kernel_size; c.stride = 1; c.pad = d->kernel_size / 2; c.relu = 1;
    { char nm[32]; snprintf(nm, sizeof nm, "cnn.conv%d_f32", i + 1); KernelTimer kt(nm, s); I2L_TRY(conv2d_f32(c, s)); }
    // conv output -> w.a, pooled output -> w.b[i&1] (the next layer reads it while writing w.a)
    { char nm[32]; snprintf(nm, sizeof nm, "cnn.pool%d_f32", i + 1); KernelTimer kt(nm, s);
      I2L_TRY(maxpool2d_f32(w.a, w.b[i & 1], batch, d->filters[i], L.h[i], L.w[i], d->pool_size, d->pool_size, 0, s)); }
    cur = w.b[i & 1];
  }
  GemmF32 g;
  g.M = batch; g.N = d->embedding_dim; g.C = out; g.ldc = d->embedding_dim;
  g.A1 = cur; g.lda1 = (int)L.flat; g.W1 = pk + L.fc_w; g.ldw1 = (int)L.flat; g.K1 = (int)L.flat;
  g.bias = pk + L.fc_b; g.relu = 1; g.splitk = w.splitk; g.splitk_ws = w.sk;
  KernelTimer kt("cnn.fc_f32", s);
  return gemm_f32(g, s);
}

extern "C" int i2l_cnn_encoder_fwd_u8(const i2l_cnn_desc* d, const void* packed, const uint8_t* x, int32_t norm_mode,
                                      const float* mean, const float* stdv, int32_t batch, float* out, void* workspace,
                                      size_t workspace_bytes, void* stream) {
  I2L_TRY(device_check());
  I2L_TRY(cnn_check(d));
  I2L_REQUIRE(packed && batch >= 0, "i2l_cnn_encoder_fwd_u8: invalid argument");
  I2L_REQUIRE(norm_mode == I2L_NORM_PM1 || (norm_mode == I2L_NORM_MEANSTD && mean && stdv), "i2l_cnn_encoder_fwd_u8: invalid normalisation");
  if (batch == 0) return I2L_OK;
  I2L_REQUIRE(x && out && workspace, "i2l_cnn_encoder_fwd_u8: null buffer");
  CnnLayout L = cnn_layout(*d);
  if (!L.bf16_section) {
    set_error("i2l_cnn_encoder_fwd_u8: fused uint8 input needs precision == I2L_BF16 and a tcgen05 configuration (i2l_cnn_tensor_core_path: 32/64/128, "
              "E=256); use i2l_normalize_u8 + i2l_cnn_encoder_fwd otherwise");
    return I2L_ERR_UNSUPPORTED;
  }
  float a[3], b[3];
  for (int c = 0; c < 3; ++c) {
    if (norm_mode == I2L_NORM_PM1) { a[c] = 2.0f / 255.0f; b[c] = -1.0f; }                  // x/255*2-1
    else { a[c] = 1.0f / (255.0f * stdv[c]); b[c] = -mean[c] / stdv[c]; }                   // (x/255-mean)/std
  }
  return cnn_bf16_fwd(*d, reinterpret_cast<const char*>(packed) + L.bf16_section, x, I2L_IN_U8, batch, out, workspace,
                      workspace_bytes, (cudaStream_t)stream, a, b);
}

// ====================================================================== ResNet C ABI
extern "C" int32_t i2l_resnet_num_convs(int32_t depth) {
  RNet n = build_resnet(depth);
  return n.ok ? (int32_t)n.convs.size() : -1;
}

extern "C" size_t i2l_resnet_packed_bytes(const i2l_resnet_desc* d) {
  if (!d) return 0;
  RNet n = build_resnet(d->depth);
  if (!n.ok || d->embedding_dim <= 0) return 0;
  size_t bytes = resnet_layout(n, d->embedding_dim).end_f32 * 4;
  if (d->precision == I2L_BF16) bytes = align_up(bytes, 1024) + resnet_bf16_packed_bytes(n, d->embedding_dim);
  return bytes;
}

extern "C" int i2l_resnet_pack(const i2l_resnet_desc* d, const i2l_resnet_params* p, void* packed,
                               size_t packed_bytes, void* stream) {
  I2L_TRY(device_check());
  I2L_REQUIRE(d && p && packed, "i2l_resnet_pack: null argument");
  RNet n = build_resnet(d->depth);
  I2L_REQUIRE(n.ok, "Invalid ResNet model name: resnet%d", d->depth);      // encoder.py:195-196
  I2L_REQUIRE(p->n_convs == (int)n.convs.size(), "i2l_resnet_pack: expected %d convs, got %d", (int)n.convs.size(), p->n_convs);
  RLayout L = resnet_layout(n, d->embedding_dim);
  if (packed_bytes < i2l_resnet_packed_bytes(d)) { set_error("i2l_resnet_pack: packed buffer too small"); return I2L_ERR_WORKSPACE; }
  cudaStream_t s = (cudaStream_t)stream;
  float* pk = reinterpret_cast<float*>(packed);
  for (size_t i = 0; i < n.convs.size(); ++i) {
    const RConv& c = n.convs[i];
    I2L_REQUIRE(p->conv_w[i] && p->bn_weight[i] && p->bn_bias[i] && p->bn_mean[i] && p->bn_var[i],
                "i2l_resnet_pack: missing tensors for conv %d", (int)i);
    int per = c.ci * c.k * c.k;
    size_t tot = (size_t)c.co * per;
    fold_bn_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(p->conv_w[i], p->bn_weight[i], p->bn_bias[i],
                                                               p->bn_mean[i], p->bn_var[i], pk + c.w_off,
                                                               pk + c.b_off, c.co, per);
    I2L_LAUNCH_OK();
  }
  I2L_REQUIRE(p->fc_w && p->fc_b, "i2l_resnet_pack: missing embedding layer");
  I2L_CUDA_OK(cudaMemcpyAsync(pk + L.fc_w, p->fc_w, (size_t)d->embedding_dim * n.feat * 4, cudaMemcpyDeviceToDevice, s));
  I2L_CUDA_OK(cudaMemcpyAsync(pk + L.fc_b, p->fc_b, (size_t)d->embedding_dim * 4, cudaMemcpyDeviceToDevice, s));
  if (d->precision == I2L_BF16)
    I2L_TRY(resnet_bf16_pack(n, d->embedding_dim, pk, pk + L.fc_w, reinterpret_cast<char*>(packed) + align_up(L.end_f32 * 4, 1024), s));
  return I2L_OK;
}

extern "C" size_t i2l_resnet_workspace_bytes(const i2l_resnet_desc* d, int32_t batch, int32_t img_width) {
  if (!d || batch <= 0 || img_width <= 0) return 0;
  RNet n = build_resnet(d->depth);
  if (!n.ok) return 0;
  if (resnet_bf16_supported(*d, img_width)) return resnet_bf16_workspace_bytes(n, batch, d->img_height, img_width);
  size_t act = align_up(resnet_max_act(n, d->img_height, img_width) * (size_t)batch * 4, 256);
  return 5 * act + align_up((size_t)batch * n.feat * 4, 256);
}

extern "C" int i2l_resnet_encoder_fwd(const i2l_resnet_desc* d, const void* packed, const float* x, int32_t batch,
                                      int32_t img_width, float* out, void* workspace, size_t workspace_bytes,
                                      void* stream) {
  I2L_TRY(device_check());
  I2L_REQUIRE(d && packed && batch >= 0 && img_width > 0, "i2l_resnet_encoder_fwd: invalid argument");
  if (batch == 0) return I2L_OK;
  I2L_REQUIRE(x && out && workspace, "i2l_resnet_encoder_fwd: null buffer");
  RNet n = build_resnet(d->depth);
  I2L_REQUIRE(n.ok, "Invalid ResNet model name: resnet%d", d->depth);
  size_t need = i2l_resnet_workspace_bytes(d, batch, img_width);
  if (workspace_bytes < need) { set_error("i2l_resnet_encoder_fwd: workspace too small (%zu < %zu)", workspace_bytes, need); return I2L_ERR_WORKSPACE; }
  cudaStream_t s = (cudaStream_t)stream;
  RLayout L = resnet_layout(n, d->embedding_dim);
  const float* pk = reinterpret_cast<const float*>(packed);
  if (resnet_bf16_supported(*d, img_width))
    return resnet_bf16_fwd(n, *d, pk, reinterpret_cast<const char*>(packed) + align_up(L.end_f32 * 4, 1024), pk + L.fc_w, pk + L.fc_b,
                           x, batch, img_width, out, workspace, workspace_bytes, s);
  size_t act = align_up(resnet_max_act(n, d->img_height, img_width) * (size_t)batch * 4, 256);
  char* base = reinterpret_cast<char*>(workspace);
  float* buf[5];
  for (int i = 0; i < 5; ++i) buf[i] = reinterpret_cast<float*>(base + i * act);
  float* pooled = reinterpret_cast<float*>(base + 5 * act);
  auto osz = [](int v, int k, int st, int p) { return (v + 2 * p - k) / st + 1; };
  auto run = [&](int idx, const float* in, float* o, int h, int w, const float* res, int relu) {
    const RConv& c = n.convs[idx];
    ConvF32 cv;
    cv.x = in; cv.w = pk + c.w_off; cv.bias = pk + c.b_off; cv.residual = res; cv.y = o;
    cv.B = batch; cv.Ci = c.ci; cv.Hi = h; cv.Wi = w; cv.Co = c.co; cv.KH = cv.KW = c.k; cv.stride = c.stride;
    cv.pad = c.pad; cv.relu = relu;
    return conv2d_f32(cv, s);
  };
  int h = d->img_height, w = img_width;
  I2L_TRY(run(0, x, buf[1], h, w, nullptr, 1));                      // conv1 + bn1 + relu
  h = osz(h, 7, 2, 3); w = osz(w, 7, 2, 3);
  I2L_TRY(maxpool2d_f32(buf[1], buf[0], batch, 64, h, w, 3, 2, 1, s));
  h = osz(h, 3, 2, 1); w = osz(w, 3, 2, 1);
  float* cur = buf[0]; float* nxt = buf[1]; float* t1 = buf[2]; float* t2 = buf[3]; float* idt = buf[4];
  for (const RBlock& b : n.blocks) {
    const RConv& c1 = n.convs[b.c1];
    const RConv& c2 = n.convs[b.c2];
    int h1 = osz(h, c1.k, c1.stride, c1.pad), w1 = osz(w, c1.k, c1.stride, c1.pad);
    int h2 = osz(h1, c2.k, c2.stride, c2.pad), w2 = osz(w1, c2.k, c2.stride, c2.pad);
    const float* res = cur;
    if (b.ds >= 0) { I2L_TRY(run(b.ds, cur, idt, h, w, nullptr, 0)); res = idt; }
    I2L_TRY(run(b.c1, cur, t1, h, w, nullptr, 1));
    if (b.c3 < 0) {
      I2L_TRY(run(b.c2, t1, nxt, h1, w1, res, 1));
    } else {
      I2L_TRY(run(b.c2, t1, t2, h1, w1, nullptr, 1));
      I2L_TRY(run(b.c3, t2, nxt, h2, w2, res, 1));
    }
    std::swap(cur, nxt);
    h = h2; w = w2;
  }
  I2L_TRY(global_avgpool_f32(cur, pooled, batch, n.feat, h * w, s));   // AdaptiveAvgPool2d(1)
  GemmF32 g;
  g.M = batch; g.N = d->embedding_dim; g.C = out; g.ldc = d->embedding_dim;
  g.A1 = pooled; g.lda1 = n.feat; g.W1 = pk + L.fc_w; g.ldw1 = n.feat; g.K1 = n.feat;
  g.bias = pk + L.fc_b; g.relu = 1;
  return gemm_f32(g, s);
}
