// Device-side image geometry of the reference's `load_image` / `Predictor._prepare_image`
// (img2latex/data/utils.py:37-48, img2latex/data/transforms.py:9-56, training/predictor.py:427-436):
// optional convert("L"), Pillow resize (LANCZOS or BICUBIC), white right-padding / centre crop, for a
// ragged batch of uint8 images in one pair of launches.
//
// The arithmetic is Pillow's 8-bit resampler (src/libImaging/Resample.c): per output pixel a window of
// 22-bit fixed-point weights, a horizontal pass then a vertical pass with a uint8 intermediate, int32
// accumulation from a rounding bias of 1 << 21, arithmetic shift by 22 and a clamp to [0, 255] -- integer
// work, bit-exact against Pillow.  The weights are computed ON THE HOST in double precision with the same
// libm calls Pillow makes (i2l_resize_plan_build); the caller copies the plan next to the pixels, so the
// device kernels never evaluate sin() and cannot round a weight differently.
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <thread>
#include <utility>
#include <vector>

#include "common.cuh"

namespace i2l {
namespace {

constexpr int kPrecisionBits = 32 - 8 - 2;   // Resample.c PRECISION_BITS
constexpr uint32_t kPlanMagic = 0x69326c52;  // "i2lR"

struct PlanHeader {
  uint32_t magic; int32_t n, src_channels, out_channels, to_gray, target_h, target_w, filter, mode;
  int32_t max_inter_pixels;          // largest ceil(h / 8) * kept_w over the batch (grid of the horizontal pass)
  uint64_t plan_bytes, workspace_bytes, images_off;
};

struct ImagePlan {
  int64_t src_offset;                // bytes into the packed source buffer
  uint64_t inter_off;                // bytes into the workspace: h rows of `pitch` bytes
  uint64_t kh_off, bh_off, kv_off, bv_off;   // plan offsets: transposed weights (ksize, out) int32, bounds (out) int2
  int32_t h, w;                      // source size
  int32_t new_w, new_h;              // resize target of this image
  int32_t left;                      // first resized column kept (centre crop)
  int32_t kept_w, kept_h;            // resized columns / rows that land in the output
  int32_t need_h, need_v;            // Resample.c ImagingResample: passes that are not the identity
  int32_t pitch;                     // bytes per intermediate row (kept_w * out_channels rounded up to 16)
  int32_t pad_;
};

double sinc_filter(double x) {
  if (x == 0.0) return 1.0;
  x = x * M_PI;
  return sin(x) / x;
}
double lanczos_filter(double x) {           // Resample.c: truncated sinc, support 3
  if (-3.0 <= x && x < 3.0) return sinc_filter(x) * sinc_filter(x / 3);
  return 0.0;
}
double bicubic_filter(double x) {           // Resample.c: Keys cubic, a = -0.5, support 2
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
  if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
  return 0.0;
}

int ksize_of(int in_size, int out_size, double fsupport) {
  double filterscale = (double)in_size / out_size;
  if (filterscale < 1.0) filterscale = 1.0;
  return (int)ceil(fsupport * filterscale) * 2 + 1;
}

// Resample.c precompute_coeffs + normalize_coeffs_8bpc for the full box; weights stored TRANSPOSED
// ((ksize, out_size): consecutive output pixels read consecutive words).
void build_table(int in_size, int out_size, int filter, int32_t* kk_t, int32_t* bounds) {
  const double fsupport = filter == 0 ? 3.0 : 2.0;
  double (*f)(double) = filter == 0 ? lanczos_filter : bicubic_filter;
  double scale = (double)in_size / out_size, filterscale = scale;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = fsupport * filterscale;
  const int ksize = (int)ceil(support) * 2 + 1;
  const double ss = 1.0 / filterscale;
  std::vector<double> k(ksize);
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = (xx + 0.5) * scale;
    double ww = 0.0;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    for (int x = 0; x < xmax; ++x) {
      double w = f((x + xmin - center + 0.5) * ss);
      k[x] = w;
      ww += w;
    }
    for (int x = 0; x < xmax; ++x)
      if (ww != 0.0) k[x] /= ww;
    for (int x = 0; x < ksize; ++x) {
      int32_t q = 0;
      if (x < xmax) q = k[x] < 0 ? (int)(-0.5 + k[x] * (1 << kPrecisionBits)) : (int)(0.5 + k[x] * (1 << kPrecisionBits));
      kk_t[(size_t)x * out_size + xx] = q;
    }
    bounds[2 * xx] = xmin;
    bounds[2 * xx + 1] = xmax;
  }
}

// transforms.py:31-35: int(round(target_height * (width / height))), Python's round-half-to-even on a double
bool aspect_width(int width, int height, int target_h, int* out) {
  double v = (double)target_h * ((double)width / (double)height);
  double r = nearbyint(v);                    // default rounding mode: to nearest, ties to even
  if (r < 1.0 || r > 1e8) return false;
  *out = (int)r;
  return true;
}

struct Geometry { int new_w, new_h, left, kept_w, kept_h; bool ok; };
Geometry geometry(int h, int w, int target_h, int target_w, int mode) {
  Geometry g{0, 0, 0, 0, 0, true};
  if (mode == 1) {                            // predictor.py:436: image.resize(img_size[::-1]) -- plain stretch
    if (h <= 0 || w <= 0) { g.ok = false; return g; }
    g.new_w = target_w; g.new_h = target_h; g.kept_w = target_w; g.kept_h = target_h;
    return g;
  }
  if (h == 0) return g;                       // transforms.py:28-29: all-white image
  if (w <= 0 || !aspect_width(w, h, target_h, &g.new_w)) { g.ok = false; return g; }   // Pillow: ValueError
  g.new_h = target_h; g.kept_h = target_h;
  if (g.new_w > target_w) { g.left = (g.new_w - target_w) / 2; g.kept_w = target_w; }  // transforms.py:51-56
  else g.kept_w = g.new_w;                                                             // transforms.py:44-50
  return g;
}

__device__ __forceinline__ int clip8(int acc) {
  int v = acc >> kPrecisionBits;              // arithmetic shift, then clip8_lookups
  return min(max(v, 0), 255);
}
// Convert.c rgb2l (L24 >> 16): ITU-R 601-2 luma in 16-bit fixed point
__device__ __forceinline__ int luma(int r, int g, int b) { return (r * 19595 + g * 38470 + b * 7471 + 0x8000) >> 16; }

// Pass 1 (ImagingResampleHorizontal_8bpc, restricted to the columns the crop keeps): source HWC uint8 ->
// intermediate (h, pitch) uint8 rows of kept_w * CO interleaved samples (pitch = 16-byte multiple).
// One thread per (output column, group of RG source rows): the window bounds and every filter weight are
// loaded once and reused for the RG rows, so a tap costs one byte load per row.  blockIdx.y = image.
constexpr int kRowGroup = 8;
constexpr int kCtasPerImage = 6;    // measured on B200, 1024 ragged images -> 64x800: 2 / 4 / 6 / 10 / uncapped = 0.157 / 0.141 / 0.139 / 0.144 / 0.193 ms
template <int CS, int CO>
__global__ void __launch_bounds__(256) resize_rows_kernel(const uint8_t* __restrict__ src, const char* __restrict__ plan,
                                                          uint8_t* __restrict__ ws) {
  const PlanHeader* hd = reinterpret_cast<const PlanHeader*>(plan);
  const ImagePlan ip = reinterpret_cast<const ImagePlan*>(plan + hd->images_off)[blockIdx.y];
  const int groups = (ip.h + kRowGroup - 1) / kRowGroup;
  const int total = groups * ip.kept_w;
  const uint8_t* s = src + ip.src_offset;
  uint8_t* o = ws + ip.inter_off;
  const int32_t* kk = reinterpret_cast<const int32_t*>(plan + ip.kh_off);
  const int2* bnd = reinterpret_cast<const int2*>(plan + ip.bh_off);
  const size_t row_bytes = (size_t)ip.w * CS;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int g = idx / ip.kept_w, xx = idx - g * ip.kept_w, gx = xx + ip.left;
    const int y0 = g * kRowGroup;
    const int rows = min(kRowGroup, ip.h - y0);
    int acc[kRowGroup][CO];
    if (!ip.need_h) {                           // width unchanged: the pass is skipped (converting copy)
#pragma unroll
      for (int r = 0; r < kRowGroup; ++r) {
        if (r < rows) {
          const uint8_t* p = s + (size_t)(y0 + r) * row_bytes + (size_t)gx * CS;
          if (CS == 3 && CO == 1) acc[r][0] = luma(p[0], p[1], p[2]);
          else {
#pragma unroll
            for (int c = 0; c < CO; ++c) acc[r][c] = p[c];
          }
        }
      }
    } else {
      const int2 b = bnd[gx];
#pragma unroll
      for (int r = 0; r < kRowGroup; ++r)
#pragma unroll
        for (int c = 0; c < CO; ++c) acc[r][c] = 1 << (kPrecisionBits - 1);
      const uint8_t* p0 = s + (size_t)y0 * row_bytes + (size_t)b.x * CS;
      if (rows == kRowGroup) {
#pragma unroll 4
        for (int x = 0; x < b.y; ++x, p0 += CS) {
          const int k = kk[(size_t)x * ip.new_w + gx];
#pragma unroll
          for (int r = 0; r < kRowGroup; ++r) {
            const uint8_t* p = p0 + r * row_bytes;
            if (CS == 3 && CO == 1) acc[r][0] += luma(p[0], p[1], p[2]) * k;
            else {
#pragma unroll
              for (int c = 0; c < CO; ++c) acc[r][c] += p[c] * k;
            }
          }
        }
      } else {
        for (int x = 0; x < b.y; ++x, p0 += CS) {
          const int k = kk[(size_t)x * ip.new_w + gx];
#pragma unroll
          for (int r = 0; r < kRowGroup; ++r) {
            if (r < rows) {
              const uint8_t* p = p0 + r * row_bytes;
              if (CS == 3 && CO == 1) acc[r][0] += luma(p[0], p[1], p[2]) * k;
              else {
#pragma unroll
                for (int c = 0; c < CO; ++c) acc[r][c] += p[c] * k;
              }
            }
          }
        }
      }
#pragma unroll
      for (int r = 0; r < kRowGroup; ++r)
#pragma unroll
        for (int c = 0; c < CO; ++c) acc[r][c] = clip8(acc[r][c]);
    }
#pragma unroll
    for (int r = 0; r < kRowGroup; ++r) {
      if (r < rows) {
#pragma unroll
        for (int c = 0; c < CO; ++c) o[(size_t)(y0 + r) * ip.pitch + (size_t)xx * CO + c] = (uint8_t)acc[r][c];
      }
    }
  }
}

// Pass 2 (ImagingResampleVertical_8bpc) + padding: intermediate -> dst (n, CO, target_h, target_w) planar.
// The vertical pass treats a row as a flat array of kept_w * CO samples, and the weight of a tap is the same
// for the whole output row: one thread produces 16 (mode L) or 4 (RGB) consecutive samples of one output row
// from one aligned 128- / 32-bit load per tap.  Samples beyond the resized image take the colour of `Image.new(mode, size, 255)`
// (transforms.py:29,45-47): white for mode L, but (255, 0, 0) for mode RGB -- Pillow reads the integer 255
// as 0x0000FF, so the reference pads RGB images with red.
template <int CO, int SPT>       // SPT = samples per thread: 16 (one 128-bit load per tap) or 4
__global__ void __launch_bounds__(256) resize_cols_kernel(const uint8_t* __restrict__ ws, const char* __restrict__ plan,
                                                          uint8_t* __restrict__ dst) {
  const PlanHeader* hd = reinterpret_cast<const PlanHeader*>(plan);
  const ImagePlan ip = reinterpret_cast<const ImagePlan*>(plan + hd->images_off)[blockIdx.y];
  const int TH = hd->target_h, TW = hd->target_w, plane = TH * TW;
  const int row_samples = TW * CO;
  const int chunks = (row_samples + SPT - 1) / SPT;      // groups of SPT samples per output row
  const int total = TH * chunks;
  const int valid = ip.kept_w * CO;                      // samples of a row that come from the image
  const uint8_t* in = ws + ip.inter_off;
  const int32_t* kk = reinterpret_cast<const int32_t*>(plan + ip.kv_off);
  const int2* bnd = reinterpret_cast<const int2*>(plan + ip.bv_off);
  uint8_t* o = dst + (size_t)blockIdx.y * CO * plane;
  constexpr int W = SPT / 4;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int yy = idx / chunks, q = idx - yy * chunks, s0 = q * SPT;
    int v[SPT];
    const bool inside = s0 < valid && yy < ip.kept_h;
    if (!inside) {
#pragma unroll
      for (int j = 0; j < SPT; ++j) v[j] = 0;
    } else if (!ip.need_v) {
      uint32_t w[W];
      if (SPT == 16) *reinterpret_cast<uint4*>(w) = *reinterpret_cast<const uint4*>(in + (size_t)yy * ip.pitch + s0);
      else w[0] = *reinterpret_cast<const uint32_t*>(in + (size_t)yy * ip.pitch + s0);
#pragma unroll
      for (int j = 0; j < SPT; ++j) v[j] = (int)__byte_perm(w[j >> 2], 0u, 0x4440u + (j & 3));   // byte j & 3, zero-extended: one PRMT
    } else {
      const int2 b = bnd[yy];
#pragma unroll
      for (int j = 0; j < SPT; ++j) v[j] = 1 << (kPrecisionBits - 1);
      const uint8_t* p = in + (size_t)b.x * ip.pitch + s0;
      for (int y = 0; y < b.y; ++y, p += ip.pitch) {
        const int k = kk[(size_t)y * ip.new_h + yy];
        uint32_t w[W];
        if (SPT == 16) *reinterpret_cast<uint4*>(w) = *reinterpret_cast<const uint4*>(p);
        else w[0] = *reinterpret_cast<const uint32_t*>(p);
#pragma unroll
        for (int j = 0; j < SPT; ++j) v[j] += (int)__byte_perm(w[j >> 2], 0u, 0x4440u + (j & 3)) * k;
      }
#pragma unroll
      for (int j = 0; j < SPT; ++j) v[j] = clip8(v[j]);
    }
    if (CO == 1) {
#pragma unroll
      for (int j = 0; j < SPT; ++j)
        if (!inside || s0 + j >= valid) v[j] = 255;
      uint8_t* op = o + (size_t)yy * TW + s0;
      if (s0 + SPT <= TW && (reinterpret_cast<uintptr_t>(op) & (SPT - 1)) == 0) {
        uint32_t w[W];
#pragma unroll
        for (int j = 0; j < W; ++j)
          w[j] = (uint32_t)v[4 * j] | ((uint32_t)v[4 * j + 1] << 8) | ((uint32_t)v[4 * j + 2] << 16) | ((uint32_t)v[4 * j + 3] << 24);
        if (SPT == 16) *reinterpret_cast<uint4*>(op) = *reinterpret_cast<uint4*>(w);
        else *reinterpret_cast<uint32_t*>(op) = w[0];
      } else {
#pragma unroll
        for (int j = 0; j < SPT; ++j)
          if (s0 + j < TW) op[j] = (uint8_t)v[j];
      }
    } else {
#pragma unroll
      for (int j = 0; j < SPT; ++j) {
        const int smp = s0 + j;
        if (smp < row_samples) {
          const int xx = smp / CO, c = smp - xx * CO;
          const int val = (!inside || smp >= valid) ? (c == 0 ? 255 : 0) : v[j];
          o[(size_t)c * plane + (size_t)yy * TW + xx] = (uint8_t)val;
        }
      }
    }
  }
}

struct TableKey { int in, out; bool operator<(const TableKey& o) const { return in != o.in ? in < o.in : out < o.out; } };

// Walks the batch and lays the plan out; with `blob` == nullptr only the sizes are computed.
int layout_plan(const i2l_image_desc* imgs, int n, int src_channels, int to_gray, int target_h, int target_w,
                int filter, int mode, char* blob, size_t blob_bytes, size_t* need_bytes) {
  I2L_REQUIRE(n >= 0 && (n == 0 || imgs != nullptr), "resize plan: null image list");
  I2L_REQUIRE(src_channels == 1 || src_channels == 3, "resize plan: src_channels must be 1 (L) or 3 (RGB)");
  I2L_REQUIRE(!to_gray || src_channels == 3, "resize plan: to_gray needs RGB sources");
  I2L_REQUIRE(target_h > 0 && target_w > 0, "resize plan: target size must be positive");
  I2L_REQUIRE(filter == I2L_FILTER_LANCZOS || filter == I2L_FILTER_BICUBIC, "resize plan: unknown filter");
  I2L_REQUIRE(mode == I2L_RESIZE_ASPECT_PAD_CROP || mode == I2L_RESIZE_STRETCH, "resize plan: unknown mode");
  I2L_REQUIRE(n <= 65535, "resize plan: at most 65535 images per call");
  const int co = to_gray ? 1 : src_channels;
  const double fsupport = filter == 0 ? 3.0 : 2.0;
  size_t off = align_up(sizeof(PlanHeader), 16);
  const size_t images_off = off;
  off = align_up(off + (size_t)n * sizeof(ImagePlan), 16);
  std::map<TableKey, std::pair<size_t, size_t>> tables;      // (in, out) -> (weights offset, bounds offset)
  struct TableJob { int in, out; size_t k_off, b_off; };
  std::vector<TableJob> jobs;
  size_t ws = 0;
  int max_inter = 0;
  ImagePlan* ips = blob ? reinterpret_cast<ImagePlan*>(blob + images_off) : nullptr;
  auto table = [&](int in, int out) -> std::pair<size_t, size_t> {
    TableKey key{in, out};
    auto it = tables.find(key);
    if (it != tables.end()) return it->second;
    const int ks = ksize_of(in, out, fsupport);
    size_t k_off = off;
    off = align_up(off + (size_t)ks * out * sizeof(int32_t), 16);
    size_t b_off = off;
    off = align_up(off + (size_t)out * 2 * sizeof(int32_t), 16);
    if (blob) jobs.push_back(TableJob{in, out, k_off, b_off});
    tables[key] = {k_off, b_off};
    return {k_off, b_off};
  };
  for (int i = 0; i < n; ++i) {
    const int h = imgs[i].height, w = imgs[i].width;
    I2L_REQUIRE(h >= 0 && w >= 0 && imgs[i].src_offset >= 0, "resize plan: image %d has a negative size / offset", i);
    I2L_REQUIRE((int64_t)h * w <= (1 << 28), "resize plan: image %d is too large", i);
    Geometry g = geometry(h, w, target_h, target_w, mode);
    I2L_REQUIRE(g.ok, "resize plan: image %d (%dx%d): height and width must be > 0 (Pillow raises ValueError)", i, w, h);
    ImagePlan ip{};
    ip.src_offset = imgs[i].src_offset;
    ip.h = h; ip.w = w; ip.new_w = g.new_w; ip.new_h = g.new_h; ip.left = g.left; ip.kept_w = g.kept_w; ip.kept_h = g.kept_h;
    ip.need_h = g.kept_w > 0 && g.new_w != w;
    ip.need_v = g.kept_w > 0 && g.new_h != h;
    ip.inter_off = ws;
    ip.pitch = (int)align_up((size_t)g.kept_w * co, 16);
    ws += (size_t)h * ip.pitch;
    const int groups = (h + kRowGroup - 1) / kRowGroup;
    if (g.kept_w > 0 && groups * g.kept_w > max_inter) max_inter = groups * g.kept_w;
    if (ip.need_h) { auto t = table(w, g.new_w); ip.kh_off = t.first; ip.bh_off = t.second; }
    if (ip.need_v) { auto t = table(h, g.new_h); ip.kv_off = t.first; ip.bv_off = t.second; }
    if (ips && (size_t)(images_off + (i + 1) * sizeof(ImagePlan)) <= blob_bytes) ips[i] = ip;
  }
  *need_bytes = off;
  if (blob) {
    if (off > blob_bytes) { set_error("i2l_resize_plan_build: plan buffer too small (%zu < %zu)", blob_bytes, off); return I2L_ERR_WORKSPACE; }
    // the double-precision weight tables are the host cost of a batch (two sin() per weight, as in Pillow):
    // independent jobs, spread over the host cores
    std::atomic<size_t> next{0};
    std::atomic<bool> failed{false};
    auto worker = [&]() {
      for (size_t j = next.fetch_add(1); j < jobs.size(); j = next.fetch_add(1)) {
        try {
          build_table(jobs[j].in, jobs[j].out, filter, reinterpret_cast<int32_t*>(blob + jobs[j].k_off),
                      reinterpret_cast<int32_t*>(blob + jobs[j].b_off));
        } catch (...) {
          failed.store(true);
        }
      }
    };
    unsigned nt = std::thread::hardware_concurrency();
    if (const char* e = getenv("I2L_PLAN_THREADS")) nt = (unsigned)atoi(e);     // 1 = build the tables on the calling thread
    if (nt > 32) nt = 32;
    if (nt > jobs.size() / 4) nt = (unsigned)(jobs.size() / 4);
    std::vector<std::thread> pool;
    try {                                               // nothing may throw across the C ABI: a thread that cannot be
      for (unsigned t = 1; t < nt; ++t) pool.emplace_back(worker);   // created just leaves its jobs to the others
    } catch (...) {
    }
    worker();
    for (auto& t : pool) t.join();
    if (failed.load()) { set_error("i2l_resize_plan_build: out of host memory while building a weight table"); return I2L_ERR_INVALID; }
    PlanHeader hd{};
    hd.magic = kPlanMagic; hd.n = n; hd.src_channels = src_channels; hd.out_channels = co; hd.to_gray = to_gray;
    hd.target_h = target_h; hd.target_w = target_w; hd.filter = filter; hd.mode = mode; hd.max_inter_pixels = max_inter;
    hd.plan_bytes = off; hd.workspace_bytes = align_up(ws, 256) + 256; hd.images_off = images_off;
    memcpy(blob, &hd, sizeof(hd));
  }
  return I2L_OK;
}

}  // namespace
}  // namespace i2l

using namespace i2l;

extern "C" size_t i2l_resize_plan_bytes(const i2l_image_desc* imgs, int32_t n, int32_t src_channels, int32_t to_gray,
                                        int32_t target_h, int32_t target_w, int32_t filter, int32_t mode) {
  size_t need = 0;
  try {
    if (layout_plan(imgs, n, src_channels, to_gray, target_h, target_w, filter, mode, nullptr, 0, &need) != I2L_OK) return 0;
  } catch (...) {                                      // std::bad_alloc of the host containers: never across the ABI
    set_error("i2l_resize_plan_bytes: out of host memory");
    return 0;
  }
  return need;
}

extern "C" int i2l_resize_plan_build(const i2l_image_desc* imgs, int32_t n, int32_t src_channels, int32_t to_gray,
                                     int32_t target_h, int32_t target_w, int32_t filter, int32_t mode,
                                     void* plan_host, size_t plan_bytes) {
  I2L_REQUIRE(plan_host != nullptr, "i2l_resize_plan_build: null plan buffer");
  size_t need = 0;
  try {
    return layout_plan(imgs, n, src_channels, to_gray, target_h, target_w, filter, mode, reinterpret_cast<char*>(plan_host),
                       plan_bytes, &need);
  } catch (...) {
    set_error("i2l_resize_plan_build: out of host memory");
    return I2L_ERR_INVALID;
  }
}

extern "C" int i2l_pack_images(const void* const* images, const i2l_image_desc* descs, int32_t n, int32_t channels,
                               void* dst_host) {
  // HOST: gathers n separately allocated HWC uint8 images into the packed source buffer (descs[i].src_offset), the
  // copies spread over the host cores (the Python-level loop of 1024 slice assignments costs more than the H2D copy)
  I2L_REQUIRE(n >= 0 && (channels == 1 || channels == 3), "i2l_pack_images: invalid arguments");
  if (n == 0) return I2L_OK;
  I2L_REQUIRE(images && descs && dst_host, "i2l_pack_images: null argument");
  for (int i = 0; i < n; ++i)
    I2L_REQUIRE(descs[i].height >= 0 && descs[i].width >= 0 && descs[i].src_offset >= 0 &&
                    ((int64_t)descs[i].height * descs[i].width == 0 || images[i] != nullptr),
                "i2l_pack_images: image %d is invalid", i);
  std::atomic<int> next{0};
  auto worker = [&]() {
    for (int i = next.fetch_add(1); i < n; i = next.fetch_add(1)) {
      const size_t bytes = (size_t)descs[i].height * descs[i].width * channels;
      if (bytes) memcpy(reinterpret_cast<char*>(dst_host) + descs[i].src_offset, images[i], bytes);
    }
  };
  unsigned nt = std::thread::hardware_concurrency();
  if (const char* e = getenv("I2L_PLAN_THREADS")) nt = (unsigned)atoi(e);
  if (nt > 8) nt = 8;                                    // memory-bound: a few threads saturate the copy
  if ((int)nt > n / 16) nt = (unsigned)(n / 16);
  std::vector<std::thread> pool;
  try {
    for (unsigned t = 1; t < nt; ++t) pool.emplace_back(worker);
  } catch (...) {
  }
  worker();
  for (auto& t : pool) t.join();
  return I2L_OK;
}

extern "C" size_t i2l_resize_workspace_bytes(const void* plan_host) {
  if (!plan_host) return 0;
  PlanHeader hd;
  memcpy(&hd, plan_host, sizeof(hd));
  return hd.magic == kPlanMagic ? (size_t)hd.workspace_bytes : 0;
}

extern "C" int i2l_resize_pad_u8(const uint8_t* src, const void* plan_host, const void* plan_dev, uint8_t* dst,
                                 void* workspace, size_t workspace_bytes, void* stream) {
  I2L_TRY(device_check());
  I2L_REQUIRE(plan_host && plan_dev, "i2l_resize_pad_u8: null plan");
  PlanHeader hd;
  memcpy(&hd, plan_host, sizeof(hd));
  I2L_REQUIRE(hd.magic == kPlanMagic, "i2l_resize_pad_u8: plan_host is not a plan built by i2l_resize_plan_build");
  if (hd.n == 0) return I2L_OK;
  I2L_REQUIRE(src && dst && workspace, "i2l_resize_pad_u8: null argument");
  if (workspace_bytes < hd.workspace_bytes) { set_error("i2l_resize_pad_u8: workspace too small (%zu < %zu)", workspace_bytes, (size_t)hd.workspace_bytes); return I2L_ERR_WORKSPACE; }
  cudaStream_t s = (cudaStream_t)stream;
  const char* plan = reinterpret_cast<const char*>(plan_dev);
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  if (hd.max_inter_pixels > 0) {
    KernelTimer kt("pre.resize_rows", s);
    // a few CTAs per image, each looping over its share of the (column, row group) items: sizing the grid for the
    // largest image of a ragged batch leaves half of the CTAs with nothing to do and the rest with ~350 issue cycles
    int per_image = cdiv(hd.max_inter_pixels, 256);
    if (per_image > kCtasPerImage) per_image = kCtasPerImage;
    dim3 grid((unsigned)per_image, (unsigned)hd.n);
    if (hd.src_channels == 1) resize_rows_kernel<1, 1><<<grid, 256, 0, s>>>(src, plan, ws);
    else if (hd.to_gray) resize_rows_kernel<3, 1><<<grid, 256, 0, s>>>(src, plan, ws);
    else resize_rows_kernel<3, 3><<<grid, 256, 0, s>>>(src, plan, ws);
    I2L_LAUNCH_OK();
  }
  {
    KernelTimer kt("pre.resize_cols_pad", s);
    const int cap = kCtasPerImage;
    if (hd.out_channels == 1) {
      dim3 grid((unsigned)min(cdiv(hd.target_h * cdiv(hd.target_w, 16), 256), cap), (unsigned)hd.n);
      resize_cols_kernel<1, 16><<<grid, 256, 0, s>>>(ws, plan, dst);
    } else {
      dim3 grid((unsigned)min(cdiv(hd.target_h * cdiv(hd.target_w * 3, 4), 256), 4 * cap), (unsigned)hd.n);
      resize_cols_kernel<3, 4><<<grid, 256, 0, s>>>(ws, plan, dst);
    }
    I2L_LAUNCH_OK();
  }
  return I2L_OK;
}
