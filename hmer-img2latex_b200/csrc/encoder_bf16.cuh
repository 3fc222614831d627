// bf16 tcgen05 CNN encoder path (cnn_bf16.cu).
#pragma once
#include "common.cuh"

namespace i2l {
bool cnn_bf16_supported(const i2l_cnn_desc& d);
size_t cnn_bf16_packed_bytes(const i2l_cnn_desc& d);
int cnn_bf16_pack(const i2l_cnn_desc& d, const i2l_cnn_params& p, void* section, cudaStream_t s);
size_t cnn_bf16_workspace_bytes(const i2l_cnn_desc& d, int batch);
// x: (B,3,64,320) NCHW, fp32 (I2L_IN_F32), bf16 (I2L_IN_BF16) or raw uint8 pixels (I2L_IN_U8: conv1 applies
// y = norm_a[c] * x + norm_b[c], HOST arrays of 3 floats)
int cnn_bf16_fwd(const i2l_cnn_desc& d, const void* section, const void* x, int in_dtype, int batch, float* out,
                 void* ws, size_t ws_bytes, cudaStream_t s, const float* norm_a = nullptr, const float* norm_b = nullptr);
}  // namespace i2l
