#include "encoder_bf16.cuh"
namespace i2l {
bool cnn_bf16_supported(const i2l_cnn_desc&) { return false; }
size_t cnn_bf16_packed_bytes(const i2l_cnn_desc&) { return 0; }
int cnn_bf16_pack(const i2l_cnn_desc&, const i2l_cnn_params&, void*, cudaStream_t) { return I2L_ERR_UNSUPPORTED; }
size_t cnn_bf16_workspace_bytes(const i2l_cnn_desc&, int) { return 0; }
int cnn_bf16_fwd(const i2l_cnn_desc&, const void*, const float*, int, float*, void*, size_t, cudaStream_t) { return I2L_ERR_UNSUPPORTED; }
}
