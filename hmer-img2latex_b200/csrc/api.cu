// Library-wide C-ABI pieces: version, thread-local error string, device check.
#include "common.cuh"
#include <string.h>
#include <atomic>
#include <mutex>
#include <vector>

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


// ---------------------------------------------------------------- launch counter + kernel timers
static std::atomic<long long> g_launches{0};
static thread_local bool g_counting = true;    // off while a launch sequence is being CAPTURED (nothing executes)
void count_launch() { if (g_counting) g_launches.fetch_add(1, std::memory_order_relaxed); }
void count_launches(long long n) { g_launches.fetch_add(n, std::memory_order_relaxed); }   // kernel nodes of a graph replay
void set_launch_counting(bool on) { g_counting = on; }

namespace {
constexpr int kMaxProf = 96, kMaxEv = 4096;
struct ProfEntry {
  char name[48];
  std::vector<cudaEvent_t> ev;   // start/stop pairs
  int used = 0;                  // events used
};
std::mutex g_prof_mu;
ProfEntry g_prof[kMaxProf];
int g_nprof = 0;
std::atomic<int> g_prof_on{0};
}  // namespace

KernelTimer::KernelTimer(const char* name, cudaStream_t stream) : slot(-1), s(stream) {
  if (!g_prof_on.load(std::memory_order_relaxed)) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  int i = 0;
  for (; i < g_nprof; ++i) if (strcmp(g_prof[i].name, name) == 0) break;
  if (i == g_nprof) {
    if (g_nprof == kMaxProf) return;
    strncpy(g_prof[i].name, name, sizeof(g_prof[i].name) - 1);
    ++g_nprof;
  }
  ProfEntry& e = g_prof[i];
  if (e.used + 2 > kMaxEv) return;
  while ((int)e.ev.size() < e.used + 2) {
    cudaEvent_t x;
    if (cudaEventCreate(&x) != cudaSuccess) return;
    e.ev.push_back(x);
  }
  cudaEventRecord(e.ev[e.used], s);
  slot = i;
}
KernelTimer::~KernelTimer() {
  if (slot < 0) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  ProfEntry& e = g_prof[slot];
  cudaEventRecord(e.ev[e.used + 1], s);
  e.used += 2;
}

}  // namespace i2l

extern "C" long long i2l_launch_count(void) { return i2l::g_launches.load(); }
extern "C" void i2l_prof_enable(int on) { i2l::g_prof_on.store(on); }
extern "C" void i2l_prof_reset(void) {
  std::lock_guard<std::mutex> lk(i2l::g_prof_mu);
  for (int i = 0; i < i2l::g_nprof; ++i) i2l::g_prof[i].used = 0;
}
extern "C" int i2l_prof_count(void) { return i2l::g_nprof; }
extern "C" int i2l_prof_get(int idx, char* name, int name_len, int* launches, float* total_ms) {
  std::lock_guard<std::mutex> lk(i2l::g_prof_mu);
  if (idx < 0 || idx >= i2l::g_nprof) return I2L_ERR_INVALID;
  i2l::ProfEntry& e = i2l::g_prof[idx];
  if (name && name_len > 0) { strncpy(name, e.name, name_len - 1); name[name_len - 1] = 0; }
  float tot = 0.f;
  for (int k = 0; k + 1 < e.used; k += 2) {
    if (cudaEventSynchronize(e.ev[k + 1]) != cudaSuccess) return I2L_ERR_CUDA;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, e.ev[k], e.ev[k + 1]) != cudaSuccess) return I2L_ERR_CUDA;
    tot += ms;
  }
  if (launches) *launches = e.used / 2;
  if (total_ms) *total_ms = tot;
  return I2L_OK;
}

extern "C" const char* i2l_version(void) { return "i2l_b200 0.1.0 (sm_100a)"; }
extern "C" const char* i2l_last_error(void) { return i2l::g_err; }
extern "C" int i2l_device_check(void) { return i2l::device_check(); }
