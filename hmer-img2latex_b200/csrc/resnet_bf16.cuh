// bf16 tcgen05 ResNet trunk (resnet_bf16.cu).
#pragma once
#include "resnet_net.cuh"

namespace i2l {
// precision == I2L_BF16 and an even image width (the stem works on pixel pairs)
bool resnet_bf16_supported(const i2l_resnet_desc& d, int img_width);
size_t resnet_bf16_packed_bytes(const RNet& n, int embedding_dim);
// folded: the fp32 packed region (BN-folded conv weights / biases at RConv::w_off / b_off)
int resnet_bf16_pack(const RNet& n, int embedding_dim, const float* folded, const float* fc_w, void* section, cudaStream_t s);
size_t resnet_bf16_workspace_bytes(const RNet& n, int batch, int H, int W);
int resnet_bf16_fwd(const RNet& n, const i2l_resnet_desc& d, const float* folded, const void* section, const float* fc_w,
                    const float* fc_b, const float* x, int B, int W, float* out, void* ws, size_t ws_bytes, cudaStream_t s);
}  // namespace i2l
