// Library-wide C-ABI pieces: version, thread-local error string, device check.
#include "common.cuh"
#include <string.h>

namespace i2l {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int device_check() {
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("no CUDA device available (%s); this library has no CPU fallback", cudaGetErrorString(e));
    return I2L_ERR_NO_DEVICE;
  }
  int major = 0;
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess || major != 10) {
    cudaGetLastError();
    set_error("device %d is compute capability %d.x; the kernels are built for sm_100a only", dev, major);
    return I2L_ERR_NO_DEVICE;
  }
  return I2L_OK;
}

int num_sms() {
  int dev = 0, n = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  return n;
}

}  // namespace i2l

extern "C" const char* i2l_version(void) { return "i2l_b200 0.1.0 (sm_100a)"; }
extern "C" const char* i2l_last_error(void) { return i2l::g_err; }
extern "C" int i2l_device_check(void) { return i2l::device_check(); }
