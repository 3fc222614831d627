#include "decode_kernels.cuh"
namespace i2l {
bool persistent_supported(const i2l_dec_desc&) { return false; }
size_t persistent_packed_bytes(const i2l_dec_desc&) { return 0; }
int persistent_pack(const i2l_dec_desc&, const i2l_dec_params&, void*, cudaStream_t) { return I2L_ERR_UNSUPPORTED; }
size_t persistent_workspace_bytes(const i2l_dec_desc&, int, int) { return 0; }
int persistent_greedy(const i2l_dec_desc&, const void*, const float*, const PackedDec&, const float*, int, int, int,
                      int, float, int, int64_t*, int32_t*, int32_t*, void*, size_t, cudaStream_t) { return I2L_ERR_UNSUPPORTED; }
}
