// Host-side TMA tensor-map encoder.  cuTensorMapEncodeTiled is a driver API entry point; it is
// fetched through the runtime (cudaGetDriverEntryPoint) so the library needs no libcuda link.
#include "tc_common.cuh"
#include <mutex>

namespace i2l {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_tensor_map(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box, int swizzle_bytes, int elem_bytes) {
  EncodeTiledFn fn = get_encode();
  if (!fn) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return I2L_ERR_CUDA; }
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                          : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = fn(out, elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : elem_bytes == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu %llu %llu %llu %llu)", (int)r, rank,
              (unsigned long long)gdim[0], (unsigned long long)(rank > 1 ? gdim[1] : 0),
              (unsigned long long)(rank > 2 ? gdim[2] : 0), (unsigned long long)(rank > 3 ? gdim[3] : 0),
              (unsigned long long)(rank > 4 ? gdim[4] : 0));
    return I2L_ERR_CUDA;
  }
  return I2L_OK;
}

typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeIm2colFn get_encode_im2col() {
  static EncodeIm2colFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeIm2colFn>(p);
  });
  return fn;
}

int make_im2col_map(CUtensorMap* out, const void* base, int C, int W, int H, int N, int lower_w, int lower_h,
                    int upper_w, int upper_h, int channels, int pixels, int stride_w, int stride_h, int swizzle_bytes) {
  return make_im2col_map_strided(out, base, C, W, H, N, (uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2, lower_w, lower_h,
                                 upper_w, upper_h, channels, pixels, stride_w, stride_h, swizzle_bytes);
}

int make_im2col_map_strided(CUtensorMap* out, const void* base, int C, int W, int H, int N, uint64_t w_bytes, uint64_t h_bytes,
                            uint64_t n_bytes, int lower_w, int lower_h, int upper_w, int upper_h, int channels, int pixels,
                            int stride_w, int stride_h, int swizzle_bytes) {
  EncodeIm2colFn fn = get_encode_im2col();
  if (!fn) { set_error("cuTensorMapEncodeIm2col is not available from the driver"); return I2L_ERR_CUDA; }
  cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t gstr[3] = {w_bytes, h_bytes, n_bytes};
  int lo[2] = {lower_w, lower_h}, hi[2] = {upper_w, upper_h};      // corner order is {W, H}
  cuuint32_t es[4] = {1, (cuuint32_t)stride_w, (cuuint32_t)stride_h, 1};
  CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                          : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstr, lo, hi, (cuuint32_t)channels,
                  (cuuint32_t)pixels, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeIm2col failed with CUresult %d (C %d W %d H %d N %d, box [%d,%d]x[%d,%d], stride %d,%d)", (int)r, C, W,
              H, N, lower_w, upper_w, lower_h, upper_h, stride_w, stride_h);
    return I2L_ERR_CUDA;
  }
  return I2L_OK;
}

}  // namespace i2l
