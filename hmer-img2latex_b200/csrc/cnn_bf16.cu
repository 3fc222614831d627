// bf16 CNN encoder, CNNEncoder.forward (model/encoder.py:111-129), as tcgen05 implicit-GEMM kernels for the
// reference's conv stack (conv 32/64/128, 3x3 "same", 2x2 max-pool, FC -> 256) on C x H x W images with C in {1, 3},
// H % 64 == 0, W % 32 == 0: the headline 3x64x320 (BASELINE configs 1-3) and the reference's default / serving shape
// 1x64x800 (encoder.py:50-64, training/predictor.py:409-414) among them.  Comments quote the headline sizes.
//
//   conv1  fp32 NCHW input -> bf16.  K = 27 (padded to 32): the im2col rows are built in shared
//          memory by the CTA (K-major SWIZZLE_64B UMMA operand), 4 accumulators = the 4 pixels
//          of each 2x2 pooling window, so bias + ReLU + max-pool are a per-thread epilogue.
//   conv2/3  activations live in HBM as bf16 "parity planes" [ph%2,pw%2][H/2][B][W/2][C] written
//          by the previous layer's epilogue.  With that layout every (pool-quadrant, filter-tap)
//          operand tile is a DENSE 5-D TMA box (zero fill outside the image = conv padding), and
//          a +1 row shift is a 16-row offset of the shared-memory descriptor, so one tile needs
//          8 TMA loads for its 36 (quadrant, tap) MMA groups.  Filter taps stay resident in
//          shared memory; accumulators (4 x Cout columns) in TMEM; epilogue = max over the 4
//          accumulators + bias + ReLU, written straight in the next layer's layout.
//   fc     TMA + tcgen05 split-K GEMM (A = conv3 output viewed as [B][40960], weight columns
//          permuted to NHWC order at pack time), fp32 partials + bias/ReLU reduction.
#include "encoder_bf16.cuh"
#include "tc_common.cuh"
#include <cuda_fp16.h>
#include <stdlib.h>

namespace i2l {
namespace {

using namespace tc;

constexpr int C1 = 32, C2 = 64, C3 = 128, EMB = 256;

// run-time geometry of one encoder configuration
struct Geom {
  int H, W, C0;                 // input image
  int TW, TH;                   // conv1 tiles (16 x 32 conv pixels = 8 x 16 pooled pixels) per image
  int PH1, PW1, PH2, PW2, PH3, PW3;   // pooled map sizes after conv1 / conv2 / conv3
  int WW2, WW3;                 // pooled columns per conv2 / conv3 tile (x 16 / WW images x 8 rows = 128 pixels)
  int FLAT;                     // FC input features
  int splits;                   // FC split-K factor
};
Geom make_geom(const i2l_cnn_desc& d) {
  Geom g{};
  g.H = d.img_height; g.W = d.img_width; g.C0 = d.channels;
  g.TW = g.W / 32; g.TH = g.H / 16;
  g.PH1 = g.H / 2; g.PW1 = g.W / 2; g.PH2 = g.H / 4; g.PW2 = g.W / 4; g.PH3 = g.H / 8; g.PW3 = g.W / 8;
  g.WW2 = (g.PW2 % 16) == 0 ? 16 : 8;                                 // W % 32 == 0  =>  PW2 % 8 == 0
  g.WW3 = (g.PW3 % 16) == 0 ? 16 : ((g.PW3 % 8) == 0 ? 8 : 4);        //              =>  PW3 % 4 == 0
  g.FLAT = C3 * g.PH3 * g.PW3;
  const int kblocks = g.FLAT / 64;
  g.splits = 1;
  for (int sp = 16; sp >= 1; --sp) if (kblocks % sp == 0) { g.splits = sp; break; }
  return g;
}
constexpr int BPAD = 4;       // the batch is padded to a multiple of the largest images-per-tile count (16 / 4)

// ------------------------------------------------------------------ packed section layout
struct Sec {
  size_t w1, b1, w2, b2, w3, b3, wfc, bfc, w1h, total;
};
Sec sec_layout(const Geom& G) {
  Sec s{};
  size_t o = 0;
  auto take = [&](size_t n) { size_t r = o; o = align_up(o + n, 1024); return r; };
  s.w1 = take(32 * 64);                 // [32 co][32 k] bf16, SW64 image
  s.b1 = take(C1 * 4);
  s.w2 = take(9 * C2 * C1 * 2);         // 9 taps x [64][32] SW64
  s.b2 = take(C2 * 4);
  s.w3 = take(9 * C3 * C2 * 2);         // 9 taps x [128][64] SW128
  s.b3 = take(C3 * 4);
  s.wfc = take((size_t)EMB * G.FLAT * 2); // [256][40960] bf16, NHWC column order
  s.bfc = take(EMB * 4);
  s.w1h = take(32 * 64);                // conv1 weights once more as fp16 (conv1_u8_kernel), same SW64 image
  s.total = o;
  return s;
}

__global__ void pack_conv_w_kernel(const float* __restrict__ w, int co_n, int ci_n, int layer, unsigned char* __restrict__ dst) {
  // layer 1: dst [co][k = ci*9+kh*3+kw (27 -> 32)], 64 B rows; layers 2,3: dst [tap][co][ci]
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (layer == 1) {
    if (i >= 32 * 32) return;
    int co = i / 32, k = i % 32;
    float v = k < 9 * ci_n ? w[co * 9 * ci_n + k] : 0.f;
    *reinterpret_cast<__nv_bfloat16*>(dst + swz_off(co, k / 8, 64) + (k % 8) * 2) = __float2bfloat16(v);
  } else {
    if (i >= 9 * co_n * ci_n) return;
    int ci = i % ci_n, co = (i / ci_n) % co_n, tap = i / (ci_n * co_n);
    float v = w[((size_t)co * ci_n + ci) * 9 + tap];
    uint32_t rb = ci_n * 2;
    *reinterpret_cast<__nv_bfloat16*>(dst + (size_t)tap * co_n * rb + swz_off(co, ci / 8, rb) + (ci % 8) * 2) = __float2bfloat16(v);
  }
}

__global__ void pack_conv1_w_f16_kernel(const float* __restrict__ w, int ci_n, unsigned char* __restrict__ dst) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 32 * 32) return;
  int co = i / 32, k = i % 32;
  float v = k < 9 * ci_n ? w[co * 9 * ci_n + k] : 0.f;
  *reinterpret_cast<__half*>(dst + swz_off(co, k / 8, 64) + (k % 8) * 2) = __float2half_rn(v);
}

__global__ void pack_fc_w_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ dst, int FLAT, int HW) {
  // dst[e][(h*40+w)*128 + c] = w[e][c*320 + h*40 + w]    (nn.Flatten on NCHW, encoder.py:125); HW = 8 * 40
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)EMB * FLAT) return;
  int kk = (int)(i % FLAT), e = (int)(i / FLAT);
  int c = kk % C3, hw = kk / C3;
  dst[i] = __float2bfloat16(w[(size_t)e * FLAT + (size_t)c * HW + hw]);
}

// ------------------------------------------------------------------ conv1
// Persistent, warp-specialised: pooled tile 8 x 16 of one image = 16 x 32 conv pixels.
//   warp 0      TMA producer: the zero-padded fp32 input patch (3 x 18 x 36) of the tile is one
//               4-D box load straight from the NCHW image (out-of-image = conv padding = OOB fill)
//   warps 6-13  im2col builders: thread (pooled pixel, qh) reads 3 ci x 3 rows x 4 floats as
//               LDS.64 and emits the two K=(ci,kh,kw) rows (qw = 0,1) as packed bf16 straight into
//               TENSOR MEMORY (tcgen05.st): the im2col matrix is the TMEM A operand of the MMA, so
//               there is no shared-memory round trip and no generic->async proxy fence per tile
//   warp 1      MMA issuer: 4 quadrants x 2 k-steps of 128x32x16 (A from TMEM, W from smem)
//   warps 2-5   epilogue: max over the 2x2 window, + bias, ReLU, bf16, parity-plane store
constexpr int C1_THREADS = 448;
constexpr int C1_PS = 8;                               // patch ring depth
constexpr int PATCH_H = 18;
// Input element type of conv1: fp32 (the reference's tensor dtype) or bf16 (half the H2D / HBM
// bytes; the im2col operand is bf16 either way, so the two give bit-identical results whenever the
// fp32 values are bf16-representable).  The TMA box starts on a 16-byte boundary of the image row:
// X0 columns left of the first needed one (2*pw0 - 1).
template <typename T> struct C1In;
template <> struct C1In<float> { static constexpr int PATCH_W = 40, X0 = 3, ELT = 4; };
template <> struct C1In<__nv_bfloat16> { static constexpr int PATCH_W = 48, X0 = 7, ELT = 2; };
// raw uint8 pixels: y = a[c] * x + b[c] (the normalisation of load_image / _prepare_image, data/utils.py:68-80,
// training/predictor.py:441-446) is applied while the im2col rows are built; out-of-image taps stay 0 in
// NORMALISED space (the conv's zero padding), which the TMA zero fill of the raw pixels alone would not give
template <> struct C1In<uint8_t> { static constexpr int PATCH_W = 64, X0 = 15, ELT = 1; };
struct C1Norm { float a[3], b[3]; };
constexpr int C1_PATCH_STRIDE = 9216;                  // ring slot (fp32 patch = 8640 B, bf16 patch = 5184 B)
constexpr int C1_OFF_W = 0;                            // [32 co][32 k] bf16 SW64, 2 KB
constexpr int C1_TC_ACC = 0, C1_TC_A = 256;            // TMEM columns: 2 x (4 x 32) accumulators, 2 x (4 x 16) im2col A tiles
constexpr int C1_OFF_PATCH = C1_OFF_W + 2048;
constexpr int C1_OFF_BAR = C1_OFF_PATCH + C1_PS * C1_PATCH_STRIDE;
constexpr int C1_SMEM = C1_OFF_BAR + 256;

template <typename InT, int CIN0>
__global__ void __launch_bounds__(C1_THREADS, 1)
conv1_kernel(const __grid_constant__ CUtensorMap tmx, const unsigned char* __restrict__ w1img,
             const float* __restrict__ bias1, __nv_bfloat16* __restrict__ act1, int B, int n_tiles, int dbg, const C1Norm nrm,
             int TW, int TH) {
  constexpr int PATCH_W = C1In<InT>::PATCH_W, PATCH_X0 = C1In<InT>::X0;
  constexpr int C1_PATCH_BYTES = CIN0 * PATCH_H * PATCH_W * C1In<InT>::ELT;
  const int TPI = TW * TH;                      // tiles per image
  const int QH = 4 * TH, QW = 8 * TW;           // act1 plane rows / columns (H / 4, W / 4)
  static_assert(C1_PATCH_BYTES <= C1_PATCH_STRIDE, "patch ring slot too small");
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar = sbase + C1_OFF_BAR;
  auto PFULL = [&](int s) { return bar + 8u * s; };
  auto PEMPTY = [&](int s) { return bar + 8u * (C1_PS + s); };
  auto AFULL = [&](int g) { return bar + 8u * (2 * C1_PS + g); };
  auto AEMPTY = [&](int g) { return bar + 8u * (2 * C1_PS + 2 + g); };
  auto TFULL = [&](int g) { return bar + 8u * (2 * C1_PS + 4 + g); };
  auto TEMPTY = [&](int g) { return bar + 8u * (2 * C1_PS + 6 + g); };
  uint32_t* misc = reinterpret_cast<uint32_t*>(smem + C1_OFF_BAR + 8 * (2 * C1_PS + 8));
  if ((sbase & 1023u) != 0) __trap();
  if (tid == 0) {
    for (int i = 0; i < C1_PS; ++i) { mbar_init(PFULL(i), 1); mbar_init(PEMPTY(i), 8); }
    for (int g = 0; g < 2; ++g) { mbar_init(AFULL(g), 8); mbar_init(AEMPTY(g), 1); mbar_init(TFULL(g), 1); mbar_init(TEMPTY(g), 4); }
    mbar_fence_init();
    tma_prefetch_desc(&tmx);
  }
  if (warp == 1) tmem_alloc<512>(smem_u32(misc));
  for (int i = tid; i < 2048 / 16; i += C1_THREADS) reinterpret_cast<uint4*>(smem + C1_OFF_W)[i] = reinterpret_cast<const uint4*>(w1img)[i];
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = misc[0];
  // contiguous tile range per CTA: neighbouring tiles share halo rows and DRAM pages
  const int per = (n_tiles + gridDim.x - 1) / gridDim.x;
  const int tile_beg = blockIdx.x * per, tile_end = min(n_tiles, tile_beg + per);

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int it = 0;
      for (int tile = tile_beg; tile < tile_end; ++tile, ++it) {
        const int tw = tile % TW, th = (tile / TW) % TH, b = tile / TPI;
        const int s = it % C1_PS;
        mbar_wait(PEMPTY(s), ((it / C1_PS) & 1) ^ 1);
        if (dbg & 1) { mbar_arrive(PFULL(s)); continue; }
        mbar_arrive_expect_tx(PFULL(s), C1_PATCH_BYTES);
        tma_load_4d(sbase + C1_OFF_PATCH + s * C1_PATCH_STRIDE, &tmx, 32 * tw - 1 - PATCH_X0, 16 * th - 1, 0, b, PFULL(s));
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint64_t dW = desc_base(sbase + C1_OFF_W, 64);
    constexpr uint32_t IDESC = idesc_bf16(128, 32);
    int it = 0;
    for (int tile = tile_beg; tile < tile_end; ++tile, ++it) {
      const int g = it & 1; const uint32_t par = (it >> 1) & 1;
      mbar_wait(TEMPTY(g), par ^ 1);
      mbar_wait(AFULL(g), par);
      tc_fence_after();
      if (elect_one()) {
        if (!(dbg & 2))
#pragma unroll
        for (int qd = 0; qd < 4; ++qd)
#pragma unroll
          for (int ks = 0; ks < 2; ++ks)
            tc_mma_ts(tmem + C1_TC_ACC + g * 128 + qd * 32, tmem + C1_TC_A + g * 64 + qd * 16 + ks * 8,
                      dW + (uint64_t)((ks * 32) >> 4), IDESC, ks);
        tc_commit(AEMPTY(g));
        tc_commit(TFULL(g));
      }
      __syncwarp();
    }
  } else if (warp < 6) {
    // ===================== epilogue (warps 2..5) =====================
    const int q = warp & 3;
    const int m = 32 * q + lane;
    float bias[32];
#pragma unroll
    for (int i = 0; i < 8; ++i) { float4 v = reinterpret_cast<const float4*>(bias1)[i]; bias[4 * i] = v.x; bias[4 * i + 1] = v.y; bias[4 * i + 2] = v.z; bias[4 * i + 3] = v.w; }
    int it = 0;
    for (int tile = tile_beg; tile < tile_end; ++tile, ++it) {
      const int tw = tile % TW, th = (tile / TW) % TH, b = tile / TPI;
      const int g = it & 1; const uint32_t par = (it >> 1) & 1;
      const int ph = th * 8 + (m >> 4), pw = tw * 16 + (m & 15);
      // act1 layout [plane = (ph&1)*2 + (pw&1)][16][B][80][32]
      const size_t pix = ((((size_t)((ph & 1) * 2 + (pw & 1)) * QH + (ph >> 1)) * B + b) * QW + (pw >> 1));
      uint4* dst = reinterpret_cast<uint4*>(act1 + pix * C1);
      mbar_wait(TFULL(g), par);
      tc_fence_after();
      const uint32_t ta = tmem + ((uint32_t)(32 * q) << 16) + g * 128;
      if (!(dbg & 4))
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t r0[16], r1[16], r2[16], r3[16];
        tc_ld16_nowait(ta + half * 16, r0); tc_ld16_nowait(ta + 32 + half * 16, r1);
        tc_ld16_nowait(ta + 64 + half * 16, r2); tc_ld16_nowait(ta + 96 + half * 16, r3);
        tc_wait_ld();
        uint32_t o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float a = fmaxf(fmaxf(__uint_as_float(r0[2 * i]), __uint_as_float(r1[2 * i])),
                          fmaxf(__uint_as_float(r2[2 * i]), __uint_as_float(r3[2 * i])));
          float c = fmaxf(fmaxf(__uint_as_float(r0[2 * i + 1]), __uint_as_float(r1[2 * i + 1])),
                          fmaxf(__uint_as_float(r2[2 * i + 1]), __uint_as_float(r3[2 * i + 1])));
          __nv_bfloat162 h2 = __floats2bfloat162_rn(fmaxf(a + bias[half * 16 + 2 * i], 0.f), fmaxf(c + bias[half * 16 + 2 * i + 1], 0.f));
          o[i] = *reinterpret_cast<uint32_t*>(&h2);
        }
        dst[half * 2] = make_uint4(o[0], o[1], o[2], o[3]);
        dst[half * 2 + 1] = make_uint4(o[4], o[5], o[6], o[7]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(TEMPTY(g));
    }
  } else {
    // ===================== im2col builders (warps 6..13) =====================
    // TMEM lanes are owned per warp quadrant: warp w writes rows 32*(w%4) .. +31; warps 6-9 build qh = 0, 10-13 qh = 1
    const int q = warp & 3, qh = (warp - 6) >> 2;
    const int m = 32 * q + lane;
    const uint32_t lane_addr = (uint32_t)(32 * q) << 16;
    const int pl = m >> 4, pwl = m & 15;
    int it = 0;
    for (int tile = tile_beg; tile < tile_end; ++tile, ++it) {
      const int s = it % C1_PS;
      const int g = it & 1; const uint32_t par = (it >> 1) & 1;
      mbar_wait(PFULL(s), (it / C1_PS) & 1);
      const unsigned char* patch = smem + C1_OFF_PATCH + s * C1_PATCH_STRIDE;
      // raw[ci][kh][j]: 32-bit words holding the 4 needed columns (qw + kw = 0..3) of patch row
      // (ci, 2*pl + qh + kh); fp32: one value per word, bf16: column e of the pair in half (e & 1)
      uint32_t raw[CIN0][3][4];
      if constexpr (sizeof(InT) == 1) {
        // needed patch bytes 15 + 2*pwl + j (j = 0..3): an unaligned 4-byte window over two words
        const int tw = tile % TW, th = (tile / TW) % TH;
        const bool top = th == 0 && pl == 0 && qh == 0, bot = th == TH - 1 && pl == 7 && qh == 1;  // kh = 0 / kh = 2 outside
        const bool left = tw == 0 && pwl == 0, right = tw == TW - 1 && pwl == 15;                  // j = 0 / j = 3 outside
        const int b0 = C1In<uint8_t>::X0 + 2 * pwl;
        const uint32_t* p0 = reinterpret_cast<const uint32_t*>(patch) + (2 * pl + qh) * (PATCH_W / 4) + (b0 >> 2);
        const uint32_t sh = (uint32_t)(b0 & 3) * 8;
#pragma unroll
        for (int ci = 0; ci < CIN0; ++ci)
#pragma unroll
          for (int kh = 0; kh < 3; ++kh) {
            const uint32_t* pp = p0 + (ci * PATCH_H + kh) * (PATCH_W / 4);
            const uint32_t win = __funnelshift_r(pp[0], pp[1], sh);
            const bool rz = (kh == 0 && top) || (kh == 2 && bot);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              // byte j -> exact fp32 integer via the 2^23 magic constant, then the affine normalisation
              float v = __uint_as_float(__byte_perm(win, 0x4B000000u, 0x7650u + j)) - 8388608.0f;
              v = fmaf(v, nrm.a[ci], nrm.b[ci]);
              if (rz || (j == 0 && left) || (j == 3 && right)) v = 0.f;
              raw[ci][kh][j] = __float_as_uint(v);
            }
          }
      } else if constexpr (sizeof(InT) == 4) {
        const float* p0 = reinterpret_cast<const float*>(patch) + (2 * pl + qh) * PATCH_W + 2 * pwl;
#pragma unroll
        for (int ci = 0; ci < CIN0; ++ci)
#pragma unroll
          for (int kh = 0; kh < 3; ++kh) {
            // needed columns 2*pwl + 3 .. + 6 of the patch row: three aligned LDS.64 (cols 2*pwl + 2 .. + 7)
            const float2* pp = reinterpret_cast<const float2*>(p0 + (ci * PATCH_H + kh) * PATCH_W + 2);
            float2 a = pp[0], c = pp[1], e = pp[2];
            raw[ci][kh][0] = __float_as_uint(a.y); raw[ci][kh][1] = __float_as_uint(c.x);
            raw[ci][kh][2] = __float_as_uint(c.y); raw[ci][kh][3] = __float_as_uint(e.x);
          }
      } else {
        // needed columns 2*pwl + 7 .. + 10: words pwl + 3 (hi), pwl + 4 (lo, hi), pwl + 5 (lo)
        const uint32_t* p0 = reinterpret_cast<const uint32_t*>(patch) + (2 * pl + qh) * (PATCH_W / 2) + pwl + 3;
#pragma unroll
        for (int ci = 0; ci < CIN0; ++ci)
#pragma unroll
          for (int kh = 0; kh < 3; ++kh) {
            const uint32_t* pp = p0 + (ci * PATCH_H + kh) * (PATCH_W / 2);
            uint32_t a = pp[0], c = pp[1], e = pp[2];
            raw[ci][kh][0] = a; raw[ci][kh][1] = c; raw[ci][kh][2] = c; raw[ci][kh][3] = e;
          }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(PEMPTY(s));       // patch values are in registers
      mbar_wait(AEMPTY(g), par ^ 1);
#pragma unroll
      for (int qw = 0; qw < 2; ++qw) {
        uint32_t pk[16];
#pragma unroll
        for (int k2 = 0; k2 < 16; ++k2) {
          const int ka = 2 * k2, kb = 2 * k2 + 1;
          const int ja = qw + ka % 3, jb = qw + kb % 3;       // column index 0..3 within raw[][][]
          constexpr int KREAL = 9 * CIN0;                     // 27 (RGB) or 9 (grey) real K entries, zero beyond
          if constexpr (sizeof(InT) != 2) {
            float va = ka < KREAL ? __uint_as_float(raw[ka / 9 % CIN0][(ka % 9) / 3][ja]) : 0.f;
            float vb = kb < KREAL ? __uint_as_float(raw[kb / 9 % CIN0][(kb % 9) / 3][jb]) : 0.f;
            __nv_bfloat162 h2 = __floats2bfloat162_rn(va, vb);
            pk[k2] = *reinterpret_cast<uint32_t*>(&h2);
          } else {
            // columns j = 0, 2 sit in the high half of their word, j = 1, 3 in the low half
            const uint32_t wa = ka < KREAL ? raw[ka / 9 % CIN0][(ka % 9) / 3][ja] : 0u;
            const uint32_t wb = kb < KREAL ? raw[kb / 9 % CIN0][(kb % 9) / 3][jb] : 0u;
            const uint32_t sel = ((ja & 1) ? 0x10u : 0x32u) | (((jb & 1) ? 0x54u : 0x76u) << 8);
            pk[k2] = __byte_perm(wa, wb, sel);
          }
        }
        tc_st16(tmem + lane_addr + C1_TC_A + g * 64 + (qh * 2 + qw) * 16, pk);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(AFULL(g));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
}

// ------------------------------------------------------------------ conv1, raw uint8 pixels (the bench / serving input)
// Same tile, patch ring, TMEM A operand and accumulators as conv1_kernel, restructured around the two costs ncu showed
// for the uint8 instantiation of that kernel (3100 warp-instructions per tile, half of them byte -> fp32 -> affine ->
// bf16 conversions repeated for every filter row a pixel takes part in):
//   * the im2col operand is FP16, not bf16: 0x6400 | byte IS the fp16 number 1024 + byte, so one PRMT places two pixels
//     in a half2, one HADD2 removes the offset exactly and one HFMA2 applies x * a[c] + b[c] (load_image / _prepare_image)
//     to both -- 3 instructions per pixel PAIR instead of 4 per pixel; fp16 carries 11 significant bits, so the operand
//     is closer to the reference's fp32 pixels than the bf16 one.  conv1's weights are packed as fp16 for this kernel.
//   * a builder thread owns a whole 2x2 pooling window (4 patch rows x 4 columns, every pixel converted once per thread
//     instead of once per filter row) and writes the four im2col rows of its pooled pixel; 4 warps build a tile, and
//     C1U_NG such groups work on consecutive tiles (one TMEM A buffer each) so that their LDS / STTM latencies overlap
//   * 8 epilogue warps (16 channels each), max-pool + bias + ReLU as two 3-input maxima and one add per channel
//   * tile coordinates advance incrementally (the generalised geometry made % TW, / TH run-time divisions)
constexpr int C1U_EPI = 8;                                  // epilogue warps
constexpr int C1U_B0 = 2 + C1U_EPI;                         // first builder warp
__host__ __device__ constexpr int c1u_threads(int ng) { return 32 * (C1U_B0 + 4 * ng); }   // ng builder groups = TMEM A buffers: 22 warps for 3
constexpr int C1U_PATCH_W = 64, C1U_X0 = 15;                // box starts 16 bytes left of image column 32 tw
constexpr int C1U_PATCH_STRIDE = 3584;                      // 3 x 18 x 64 = 3456 bytes per patch
constexpr int C1U_OFF_W = 0;
constexpr int C1U_OFF_PATCH = 2048;
constexpr int C1U_OFF_BAR = C1U_OFF_PATCH + C1_PS * C1U_PATCH_STRIDE;
constexpr int C1U_SMEM = C1U_OFF_BAR + 256;
constexpr int C1U_TC_A = 256;                               // TMEM: 2 x 128 accumulator columns, then C1U_NG x 64 A columns


struct TileIt {   // (tw, th) of a tile index that advances by a fixed step; the image index is not needed by every role
  int tw, th, b, TW, TH;
  __device__ TileIt(int tile, int TW_, int TH_) : TW(TW_), TH(TH_) { tw = tile % TW; th = (tile / TW) % TH; b = tile / (TW * TH); }
  __device__ void advance(int n) {
    tw += n;
    while (tw >= TW) { tw -= TW; if (++th == TH) { th = 0; ++b; } }
  }
};

template <int CIN0, int C1U_NG>
__global__ void __launch_bounds__(c1u_threads(C1U_NG), 1)
conv1_u8_kernel(const __grid_constant__ CUtensorMap tmx, const unsigned char* __restrict__ w1img_f16,
                const float* __restrict__ bias1, __nv_bfloat16* __restrict__ act1, int B, int n_tiles, const C1Norm nrm,
                int TW, int TH) {
  constexpr int PATCH_BYTES = CIN0 * PATCH_H * C1U_PATCH_W;
  constexpr int PW4 = C1U_PATCH_W / 4;
  constexpr int C1U_THREADS = c1u_threads(C1U_NG);
  static_assert(C1U_TC_A + 64 * C1U_NG <= 512, "TMEM columns");
  static_assert(PATCH_BYTES <= C1U_PATCH_STRIDE, "patch ring slot too small");
  const int QH = 4 * TH, QW = 8 * TW;           // act1 plane rows / columns (H / 4, W / 4)
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar = sbase + C1U_OFF_BAR;
  auto PFULL = [&](int s) { return bar + 8u * s; };
  auto PEMPTY = [&](int s) { return bar + 8u * (C1_PS + s); };
  auto AFULL = [&](int a) { return bar + 8u * (2 * C1_PS + a); };
  auto AEMPTY = [&](int a) { return bar + 8u * (2 * C1_PS + 4 + a); };
  auto TFULL = [&](int g) { return bar + 8u * (2 * C1_PS + 8 + g); };
  auto TEMPTY = [&](int g) { return bar + 8u * (2 * C1_PS + 10 + g); };
  uint32_t* misc = reinterpret_cast<uint32_t*>(smem + C1U_OFF_BAR + 8 * (2 * C1_PS + 12));
  if ((sbase & 1023u) != 0) __trap();
  if (tid == 0) {
    for (int i = 0; i < C1_PS; ++i) { mbar_init(PFULL(i), 1); mbar_init(PEMPTY(i), 4); }
    for (int a = 0; a < C1U_NG; ++a) { mbar_init(AFULL(a), 4); mbar_init(AEMPTY(a), 1); }
    for (int g = 0; g < 2; ++g) { mbar_init(TFULL(g), 1); mbar_init(TEMPTY(g), C1U_EPI); }
    mbar_fence_init();
    tma_prefetch_desc(&tmx);
  }
  if (warp == 1) tmem_alloc<512>(smem_u32(misc));
  for (int i = tid; i < 2048 / 16; i += C1U_THREADS) reinterpret_cast<uint4*>(smem + C1U_OFF_W)[i] = reinterpret_cast<const uint4*>(w1img_f16)[i];
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = misc[0];
  const int per = (n_tiles + gridDim.x - 1) / gridDim.x;
  const int tile_beg = blockIdx.x * per, tile_end = min(n_tiles, tile_beg + per);
  const int nt = max(tile_end - tile_beg, 0);

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one() && nt > 0) {
      TileIt t(tile_beg, TW, TH);
      for (int it = 0; it < nt; ++it, t.advance(1)) {
        const int s = it % C1_PS;
        mbar_wait(PEMPTY(s), ((it / C1_PS) & 1) ^ 1);
        mbar_arrive_expect_tx(PFULL(s), PATCH_BYTES);
        tma_load_4d(sbase + C1U_OFF_PATCH + s * C1U_PATCH_STRIDE, &tmx, 32 * t.tw - 1 - C1U_X0, 16 * t.th - 1, 0, t.b, PFULL(s));
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint64_t dW = desc_base(sbase + C1U_OFF_W, 64);
    constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // f16 x f16 -> f32
    int a = 0; uint32_t apar = 0;
    for (int it = 0; it < nt; ++it) {
      const int g = it & 1; const uint32_t par = (it >> 1) & 1;
      mbar_wait(TEMPTY(g), par ^ 1);
      mbar_wait(AFULL(a), apar);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int qd = 0; qd < 4; ++qd)
#pragma unroll
          for (int ks = 0; ks < 2; ++ks)
            tc_mma_ts(tmem + g * 128 + qd * 32, tmem + C1U_TC_A + a * 64 + qd * 16 + ks * 8, dW + (uint64_t)((ks * 32) >> 4), IDESC, ks);
        tc_commit(AEMPTY(a));
        tc_commit(TFULL(g));
      }
      __syncwarp();
      if (++a == C1U_NG) { a = 0; apar ^= 1; }
    }
  } else if (warp < C1U_B0) {
    // ===================== epilogue (8 warps: TMEM lane quadrant x channel half) =====================
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int m = 32 * q + lane;
    float bias[16];
#pragma unroll
    for (int i = 0; i < 4; ++i) { float4 v = reinterpret_cast<const float4*>(bias1)[half * 4 + i]; bias[4 * i] = v.x; bias[4 * i + 1] = v.y; bias[4 * i + 2] = v.z; bias[4 * i + 3] = v.w; }
    // act1 layout [plane = (ph&1)*2 + (pw&1)][H/4][B][W/4][32] with ph = 8 th + (m >> 4), pw = 16 tw + (m & 15): the pixel
    // index splits into a per-thread constant and a per-tile part (32-bit: B H W / 4 pixels)
    const int rowpix = B * QW;                                  // pixels of one plane row
    const int pix_thread = ((((m >> 4) & 1) * 2 + (m & 1)) * QH + (m >> 5)) * rowpix + ((m & 15) >> 1);
    TileIt t(tile_beg, TW, TH);
    for (int it = 0; it < nt; ++it, t.advance(1)) {
      const int g = it & 1; const uint32_t par = (it >> 1) & 1;
      const int pix = pix_thread + (t.th * 4) * rowpix + t.b * QW + t.tw * 8;
      uint4* dst = reinterpret_cast<uint4*>(act1 + (size_t)pix * C1 + half * 16);
      mbar_wait(TFULL(g), par);
      tc_fence_after();
      const uint32_t ta = tmem + ((uint32_t)(32 * q) << 16) + g * 128 + half * 16;
#pragma unroll
      for (int pass = 0; pass < 2; ++pass) {
        uint32_t r0[8], r1[8], r2[8], r3[8];
        tc_ld8_nowait(ta + pass * 8, r0); tc_ld8_nowait(ta + 32 + pass * 8, r1);
        tc_ld8_nowait(ta + 64 + pass * 8, r2); tc_ld8_nowait(ta + 96 + pass * 8, r3);
        tc_wait_ld();
        if (pass == 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(TEMPTY(g));     // the accumulator values are in registers
        }
        uint32_t o[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          // relu(max4 + b) = max(max4, -b) + b: two 3-input maxima and one add per channel
          const float b0 = bias[pass * 8 + 2 * i], b1 = bias[pass * 8 + 2 * i + 1];
          const float x0 = fmaxf(fmaxf(__uint_as_float(r0[2 * i]), __uint_as_float(r1[2 * i])), __uint_as_float(r2[2 * i]));
          const float x1 = fmaxf(fmaxf(__uint_as_float(r0[2 * i + 1]), __uint_as_float(r1[2 * i + 1])), __uint_as_float(r2[2 * i + 1]));
          const float y0 = fmaxf(fmaxf(x0, __uint_as_float(r3[2 * i])), -b0) + b0;
          const float y1 = fmaxf(fmaxf(x1, __uint_as_float(r3[2 * i + 1])), -b1) + b1;
          __nv_bfloat162 h2 = __floats2bfloat162_rn(y0, y1);
          o[i] = *reinterpret_cast<uint32_t*>(&h2);
        }
        dst[pass] = make_uint4(o[0], o[1], o[2], o[3]);
      }
    }
  } else {
    // ===================== im2col builders: group gi takes tiles it = gi, gi + NG, ... =====================
    const int q = warp & 3, gi = (warp - C1U_B0) >> 2;
    const int m = 32 * q + lane;
    const uint32_t lane_addr = (uint32_t)(32 * q) << 16;
    const int pl = m >> 4, pwl = m & 15;
    const int b0 = C1U_X0 + 2 * pwl;                               // patch byte of image column 2 pw - 1
    const uint32_t sh = (uint32_t)(b0 & 3) * 8;
    const int word0 = (2 * pl) * PW4 + (b0 >> 2);
    __half2 na[CIN0], nb[CIN0];
#pragma unroll
    for (int ci = 0; ci < CIN0; ++ci) { na[ci] = __float2half2_rn(nrm.a[ci]); nb[ci] = __float2half2_rn(nrm.b[ci]); }
    const __half2 k1024 = __float2half2_rn(1024.f);
    TileIt t(tile_beg + gi, TW, TH);
    uint32_t apar = 0;
    for (int it = gi; it < nt; it += C1U_NG, t.advance(C1U_NG), apar ^= 1) {
      const int s = it % C1_PS;
      // out-of-image taps are zero in NORMALISED space (the conv's padding): masks per patch row / column pair
      const uint32_t mtop = (t.th == 0 && pl == 0) ? 0u : 0xFFFFFFFFu, mbot = (t.th == TH - 1 && pl == 7) ? 0u : 0xFFFFFFFFu;
      const uint32_t m01 = (t.tw == 0 && pwl == 0) ? 0xFFFF0000u : 0xFFFFFFFFu, m23 = (t.tw == TW - 1 && pwl == 15) ? 0x0000FFFFu : 0xFFFFFFFFu;
      mbar_wait(PFULL(s), (it / C1_PS) & 1);
      const uint32_t* p0 = reinterpret_cast<const uint32_t*>(smem + C1U_OFF_PATCH + s * C1U_PATCH_STRIDE) + word0;
      uint32_t v01[CIN0][4], v23[CIN0][4];     // normalised fp16 pixel pairs (columns 0,1 / 2,3 of the 4 x 4 window)
#pragma unroll
      for (int ci = 0; ci < CIN0; ++ci)
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const uint32_t* pp = p0 + (ci * PATCH_H + r) * PW4;
          const uint32_t win = __funnelshift_r(pp[0], pp[1], sh);
          uint32_t u01 = __byte_perm(win, 0x64006400u, 0x5150), u23 = __byte_perm(win, 0x64006400u, 0x5352);
          __half2 h01 = __hfma2(__hsub2(*reinterpret_cast<__half2*>(&u01), k1024), na[ci], nb[ci]);
          __half2 h23 = __hfma2(__hsub2(*reinterpret_cast<__half2*>(&u23), k1024), na[ci], nb[ci]);
          const uint32_t mr = r == 0 ? mtop : (r == 3 ? mbot : 0xFFFFFFFFu);
          v01[ci][r] = *reinterpret_cast<uint32_t*>(&h01) & (m01 & mr);
          v23[ci][r] = *reinterpret_cast<uint32_t*>(&h23) & (m23 & mr);
        }
      __syncwarp();
      if (lane == 0) mbar_arrive(PEMPTY(s));       // patch values are in registers
      mbar_wait(AEMPTY(gi), apar ^ 1);
      constexpr int KREAL = 9 * CIN0;
#pragma unroll
      for (int qh = 0; qh < 2; ++qh)
#pragma unroll
        for (int qw = 0; qw < 2; ++qw) {
          uint32_t pk[16];
#pragma unroll
          for (int k2 = 0; k2 < 16; ++k2) {
            const int ka = 2 * k2, kb = 2 * k2 + 1;
            // K index k = (ci, kh, kw): window element (row qh + kh, column qw + kw); columns 0,2 = low halves
            const int ja = qw + ka % 3, jb = qw + kb % 3;
            const uint32_t wa = ka < KREAL ? (ja < 2 ? v01[ka / 9 % CIN0][qh + (ka % 9) / 3] : v23[ka / 9 % CIN0][qh + (ka % 9) / 3]) : 0u;
            const uint32_t wb = kb < KREAL ? (jb < 2 ? v01[kb / 9 % CIN0][qh + (kb % 9) / 3] : v23[kb / 9 % CIN0][qh + (kb % 9) / 3]) : 0u;
            if (ka >= KREAL) pk[k2] = 0u;
            else if (kb >= KREAL) pk[k2] = (ja & 1) ? (wa >> 16) : (wa & 0xFFFFu);
            else if (!(ja & 1) && (jb & 1) && ka / 3 == kb / 3) pk[k2] = wa;             // an aligned pair of one word
            else pk[k2] = __byte_perm(wa, wb, ((ja & 1) ? 0x32u : 0x10u) | (((jb & 1) ? 0x76u : 0x54u) << 8));
          }
          tc_st16(tmem + lane_addr + C1U_TC_A + gi * 64 + (qh * 2 + qw) * 16, pk);
        }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(AFULL(gi));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
}

// ------------------------------------------------------------------ conv2 / conv3
template <int CIN, int COUT, int WW, int NIMG, int STAGES, int NACC>
struct ConvCfg {
  static constexpr int HH = 8;
  static constexpr int ROWB = CIN * 2;                               // bytes per operand row = swizzle span
  static constexpr int A_ROWS = (HH + 1) * NIMG * WW;                // 144
  static constexpr int STAGE_BYTES = A_ROWS * ROWB;
  static constexpr int SHIFT_BYTES = NIMG * WW * ROWB;               // one pooled row down
  static constexpr int W_TAP_BYTES = COUT * ROWB;
  static constexpr int OFF_W = 0;
  static constexpr int OFF_A = 9 * W_TAP_BYTES;
  static constexpr int OFF_BAR = OFF_A + STAGES * STAGE_BYTES;
  static constexpr int SMEM = OFF_BAR + 256;
  static constexpr int KSTEPS = CIN / 16;
  static constexpr int TMEM_COLS = 4 * COUT * NACC;
  static_assert(NIMG * WW == 16 && HH * NIMG * WW == 128, "tile must be 128 pooled pixels with 16-row h shifts");
  static_assert(STAGE_BYTES % 1024 == 0 && SHIFT_BYTES % 1024 == 0, "alignment");
  static_assert(TMEM_COLS <= 512 && SMEM <= 232448, "budget");
};

// OUT_PARITY: write [plane][PH/2][B][PW/2][COUT] (next conv's input) else plain [B][PH][PW][COUT]
template <class Cfg, int CIN, int COUT, int WW, int NIMG, int STAGES, int NACC, bool OUT_PARITY>
__global__ void __launch_bounds__(192, 1)
conv_pool_kernel(const __grid_constant__ CUtensorMap tmap, const unsigned char* __restrict__ wimg,
                 const float* __restrict__ bias, __nv_bfloat16* __restrict__ out, int B, int PH, int PW, int n_tiles) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar = sbase + Cfg::OFF_BAR;
  auto FULL = [&](int s) { return bar + 8u * s; };
  auto EMPTY = [&](int s) { return bar + 8u * (STAGES + s); };
  auto TFULL = [&](int a) { return bar + 8u * (2 * STAGES + a); };
  auto TEMPTY = [&](int a) { return bar + 8u * (2 * STAGES + NACC + a); };
  const uint32_t WBAR = bar + 8u * (2 * STAGES + 2 * NACC);
  uint32_t* misc = reinterpret_cast<uint32_t*>(smem + Cfg::OFF_BAR + 8 * (2 * STAGES + 2 * NACC + 1));
  if ((sbase & 1023u) != 0) __trap();
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(FULL(s), 1); mbar_init(EMPTY(s), 1); }
    for (int a = 0; a < NACC; ++a) { mbar_init(TFULL(a), 1); mbar_init(TEMPTY(a), 128); }
    mbar_init(WBAR, 1);
    mbar_fence_init();
    tma_prefetch_desc(&tmap);
  }
  if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(smem_u32(misc));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = misc[0];
  const int tiles_w = PW / WW, tiles_h = PH / Cfg::HH;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      mbar_arrive_expect_tx(WBAR, 9 * Cfg::W_TAP_BYTES);
      for (int o = 0; o < 9 * Cfg::W_TAP_BYTES; o += 4096) bulk_g2s(sbase + Cfg::OFF_W + o, wimg + o, 4096, WBAR);
      int stage = 0; uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int tw = tile % tiles_w, th = (tile / tiles_w) % tiles_h, tn = tile / (tiles_w * tiles_h);
        const int ph0 = th * Cfg::HH, pw0 = tw * WW, n0 = tn * NIMG;
#pragma unroll 1
        for (int l = 0; l < 8; ++l) {
          const int rp = l < 4 ? 1 : 0, c = (0x3021 >> (4 * (l & 3))) & 3;   // plane_h; column case in the order 1, 2, 0, 3
          const int plane = rp * 2 + ((c - 1) & 1);
          // source pixel y = 2*ph + r - 1 lives in plane (y & 1) at row ph + floor((r-1)/2); same for x
          const int w2 = pw0 + (c == 0 ? -1 : (c == 3 ? 1 : 0));
          const int h2 = ph0 + (rp ? -1 : 0);
          mbar_wait(EMPTY(stage), ph ^ 1);
          mbar_arrive_expect_tx(FULL(stage), Cfg::STAGE_BYTES);
          tma_load_5d(sbase + Cfg::OFF_A + stage * Cfg::STAGE_BYTES, &tmap, 0, w2, n0, h2, plane, FULL(stage));
          if (++stage == STAGES) { stage = 0; ph ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    mbar_wait(WBAR, 0);
    tc_fence_after();
    // Accumulator columns of a tile: quadrant (qh, qw) at (2 qh + 1 - qw) * COUT.  A window (row r, column case c)
    // of the loaded tile feeds quadrant (qh, qw) through tap (r - qh, c - qw); for c = 1, 2 both qw are valid and
    // their taps (kw = c - 1 | c) are neighbours in the weight image, so ONE N = 2*COUT MMA serves both quadrants:
    // 48 instead of 72 MMAs per tile, a third less A-operand traffic on the shared-memory port.
    // The loads are visited with c in the order 1, 2, 0, 3, so every quadrant is first written by a merged MMA of
    // load 0, and every offset below is a compile-time constant (the stage of load l is l % STAGES).
    const uint64_t dA0 = desc_base(sbase + Cfg::OFF_A, Cfg::ROWB), dW0 = desc_base(sbase + Cfg::OFF_W, Cfg::ROWB);
    constexpr uint32_t IDESC1 = idesc_bf16(128, COUT), IDESC2 = idesc_bf16(128, 2 * COUT);
    static_assert(8 % STAGES == 0, "the stage of a load must not depend on the tile");
    uint32_t phs = 0;                              // bit s = parity of the next FULL(s) wait
    int acc = 0; uint32_t aph = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      mbar_wait(TEMPTY(acc), aph ^ 1);
      tc_fence_after();
      const uint32_t dacc = tmem + (uint32_t)(acc * 4 * COUT);
#pragma unroll
      for (int l = 0; l < 8; ++l) {
        constexpr int CORD[4] = {1, 2, 0, 3};
        const int rp = l < 4 ? 1 : 0, c = CORD[l & 3];
        const int stage = l % STAGES;
        mbar_wait(FULL(stage), (phs >> stage) & 1u);
        phs ^= 1u << stage;
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int off = 0; off < 2; ++off) {
            const int r = rp ? (off ? 2 : 0) : (off ? 3 : 1);
#pragma unroll
            for (int qh = 0; qh < 2; ++qh) {
              const int kh = r - qh;
              if (kh < 0 || kh > 2) continue;
              const bool first = l == 0 && ((r == 0 && qh == 0) || (r == 2 && qh == 1));   // first MMA into these quadrants
              const bool both = c == 1 || c == 2;
              const int kw_lo = both ? c - 1 : (c == 0 ? 0 : 2);                       // c = 0: qw = 0, kw = 0;  c = 3: qw = 1, kw = 2
              const int col = both ? 2 * qh : (c == 0 ? 2 * qh + 1 : 2 * qh);          // first accumulator block written
              const uint32_t d = dacc + (uint32_t)(col * COUT);
#pragma unroll
              for (int ks = 0; ks < Cfg::KSTEPS; ++ks) {
                const uint64_t ad = dA0 + (uint64_t)((stage * Cfg::STAGE_BYTES + off * Cfg::SHIFT_BYTES + ks * 32) >> 4);
                const uint64_t bd = dW0 + (uint64_t)(((kh * 3 + kw_lo) * Cfg::W_TAP_BYTES + ks * 32) >> 4);
                tc_mma_ss(d, ad, bd, both ? IDESC2 : IDESC1, (ks > 0 || !first) ? 1u : 0u);
              }
            }
          }
          tc_commit(EMPTY(stage));
          if (l == 7) tc_commit(TFULL(acc));
        }
        __syncwarp();
      }
      if (++acc == NACC) { acc = 0; aph ^= 1; }
    }
  } else {
    // ===================== epilogue warps 2..5 =====================
    const int q = warp & 3;
    const int m = 32 * q + lane;
    const int w = m % WW, n = (m / WW) % NIMG, h = m / (WW * NIMG);
    const uint32_t lane_addr = (uint32_t)(32 * q) << 16;
    int acc = 0; uint32_t aph = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int tw = tile % tiles_w, th = (tile / tiles_w) % tiles_h, tn = tile / (tiles_w * tiles_h);
      const int ph = th * Cfg::HH + h, pw = tw * WW + w, img = tn * NIMG + n;
      size_t pix;
      if (OUT_PARITY) pix = ((((size_t)((ph & 1) * 2 + (pw & 1)) * (PH >> 1) + (ph >> 1)) * B + img) * (PW >> 1) + (pw >> 1));
      else pix = ((size_t)img * PH + ph) * PW + pw;
      __nv_bfloat16* dst = out + pix * COUT;
      mbar_wait(TFULL(acc), aph);
      tc_fence_after();
      const uint32_t ta = tmem + lane_addr + (uint32_t)(acc * 4 * COUT);
#pragma unroll 1
      for (int cc = 0; cc < COUT / 16; ++cc) {
        uint32_t r0[16], r1[16], r2[16], r3[16];
        tc_ld16_nowait(ta + cc * 16, r0); tc_ld16_nowait(ta + COUT + cc * 16, r1);
        tc_ld16_nowait(ta + 2 * COUT + cc * 16, r2); tc_ld16_nowait(ta + 3 * COUT + cc * 16, r3);
        tc_wait_ld();
        uint32_t o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float a = fmaxf(fmaxf(__uint_as_float(r0[2 * i]), __uint_as_float(r1[2 * i])),
                          fmaxf(__uint_as_float(r2[2 * i]), __uint_as_float(r3[2 * i])));
          float c = fmaxf(fmaxf(__uint_as_float(r0[2 * i + 1]), __uint_as_float(r1[2 * i + 1])),
                          fmaxf(__uint_as_float(r2[2 * i + 1]), __uint_as_float(r3[2 * i + 1])));
          __nv_bfloat162 h2 = __floats2bfloat162_rn(fmaxf(a + __ldg(bias + cc * 16 + 2 * i), 0.f),
                                                    fmaxf(c + __ldg(bias + cc * 16 + 2 * i + 1), 0.f));
          o[i] = *reinterpret_cast<uint32_t*>(&h2);
        }
        if (img < B) {
          uint4* d4 = reinterpret_cast<uint4*>(dst + cc * 16);
          d4[0] = make_uint4(o[0], o[1], o[2], o[3]);
          d4[1] = make_uint4(o[4], o[5], o[6], o[7]);
        }
      }
      tc_fence_before();
      mbar_arrive(TEMPTY(acc));
      if (++acc == NACC) { acc = 0; aph ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<Cfg::TMEM_COLS>(tmem);
}

// ------------------------------------------------------------------ FC (split-K GEMM)
constexpr int FC_BM = 128, FC_BN = 256, FC_BK = 64, FC_STAGES = 4;
constexpr int FC_A_BYTES = FC_BM * FC_BK * 2, FC_B_BYTES = FC_BN * FC_BK * 2;
constexpr int FC_STAGE = FC_A_BYTES + FC_B_BYTES;
constexpr int FC_OFF_BAR = FC_STAGES * FC_STAGE;
constexpr int FC_SMEM = FC_OFF_BAR + 128;

__global__ void __launch_bounds__(192, 1)
fc_splitk_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, float* __restrict__ partial,
                 int M, int kblocks_per_split) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar = sbase + FC_OFF_BAR;
  auto FULL = [&](int s) { return bar + 8u * s; };
  auto EMPTY = [&](int s) { return bar + 8u * (FC_STAGES + s); };
  const uint32_t TFULL = bar + 8u * (2 * FC_STAGES);
  uint32_t* misc = reinterpret_cast<uint32_t*>(smem + FC_OFF_BAR + 8 * (2 * FC_STAGES + 1));
  if (tid == 0) {
    for (int s = 0; s < FC_STAGES; ++s) { mbar_init(FULL(s), 1); mbar_init(EMPTY(s), 1); }
    mbar_init(TFULL, 1);
    mbar_fence_init();
    tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmW);
  }
  if (warp == 1) tmem_alloc<256>(smem_u32(misc));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = misc[0];
  const int m0 = blockIdx.x * FC_BM, split = blockIdx.y;
  const int kb0 = split * kblocks_per_split;
  if (warp == 0) {
    if (elect_one()) {
      int stage = 0; uint32_t ph = 0;
      for (int kb = 0; kb < kblocks_per_split; ++kb) {
        mbar_wait(EMPTY(stage), ph ^ 1);
        mbar_arrive_expect_tx(FULL(stage), FC_STAGE);
        const uint32_t a = sbase + stage * FC_STAGE;
        tma_load_2d(a, &tmA, (kb0 + kb) * FC_BK, m0, FULL(stage));
        tma_load_2d(a + FC_A_BYTES, &tmW, (kb0 + kb) * FC_BK, 0, FULL(stage));
        if (++stage == FC_STAGES) { stage = 0; ph ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    constexpr uint32_t IDESC = idesc_bf16(128, 256);
    const uint64_t d0 = desc_base(sbase, 128);
    int stage = 0; uint32_t ph = 0;
    for (int kb = 0; kb < kblocks_per_split; ++kb) {
      mbar_wait(FULL(stage), ph);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          uint64_t ad = d0 + (uint64_t)((stage * FC_STAGE + ks * 32) >> 4);
          uint64_t bd = d0 + (uint64_t)((stage * FC_STAGE + FC_A_BYTES + ks * 32) >> 4);
          tc_mma_ss(tmem, ad, bd, IDESC, (kb | ks) ? 1u : 0u);
        }
        tc_commit(EMPTY(stage));
        if (kb == kblocks_per_split - 1) tc_commit(TFULL);
      }
      __syncwarp();
      if (++stage == FC_STAGES) { stage = 0; ph ^= 1; }
    }
  } else {
    const int q = warp & 3;
    const int m = m0 + 32 * q + lane;
    mbar_wait(TFULL, 0);
    tc_fence_after();
    float* dst = partial + ((size_t)split * M + m) * FC_BN;
#pragma unroll 1
    for (int cc = 0; cc < FC_BN / 16; ++cc) {
      uint32_t r[16];
      tc_ld16_nowait(tmem + ((uint32_t)(32 * q) << 16) + cc * 16, r);
      tc_wait_ld();
      if (m < M) {
#pragma unroll
        for (int i = 0; i < 4; ++i) reinterpret_cast<uint4*>(dst + cc * 16)[i] = make_uint4(r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<256>(tmem);
}

__global__ void fc_reduce_kernel(const float* __restrict__ partial, const float* __restrict__ bias, float* __restrict__ out,
                                 int M, int splits) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)M * EMB) return;
  float v = bias[i % EMB];
  for (int s = 0; s < splits; ++s) v += partial[(size_t)s * M * EMB + i];
  out[i] = fmaxf(v, 0.f);   // encoder.py:126-127
}

// ------------------------------------------------------------------ workspace
struct Ws { __nv_bfloat16 *act1, *act2, *act3; float* partial; size_t bytes; int splits; };
Ws carve(const Geom& G, int B, void* ws) {
  Arena a(ws, (size_t)-1);
  Ws w{};
  int Bp = (B + BPAD - 1) / BPAD * BPAD;   // conv2 / conv3 tiles cover groups of up to 4 images
  w.act1 = a.take<__nv_bfloat16>((size_t)Bp * G.PH1 * G.PW1 * C1);
  w.act2 = a.take<__nv_bfloat16>((size_t)Bp * G.PH2 * G.PW2 * C2);
  w.act3 = a.take<__nv_bfloat16>((size_t)Bp * G.PH3 * G.PW3 * C3);
  w.splits = G.splits;
  w.partial = a.take<float>((size_t)w.splits * B * EMB);
  w.bytes = align_up(a.off, 256);
  return w;
}

// conv2 / conv3 launch for one (pooled columns per tile) instantiation: WW x (16 / WW) images x 8 pooled rows
template <int CIN, int COUT, int WW, int STAGES, int NACC, bool OUT_PARITY>
int launch_conv_pool(const __nv_bfloat16* in, const unsigned char* wimg, const float* bias, __nv_bfloat16* out, int Bp,
                     int B_store, int PH, int PW, int sms, const char* name, cudaStream_t s) {
  constexpr int NIMG = 16 / WW;
  using Cfg = ConvCfg<CIN, COUT, WW, NIMG, STAGES, NACC>;
  // input planes [4][PH][Bp][PW][CIN] (PH x PW = this layer's POOLED output size = half the input map)
  CUtensorMap tm;
  uint64_t dims[5] = {(uint64_t)CIN, (uint64_t)PW, (uint64_t)Bp, (uint64_t)PH, 4};
  uint64_t str[4] = {(uint64_t)CIN * 2, (uint64_t)PW * CIN * 2, (uint64_t)Bp * PW * CIN * 2, (uint64_t)PH * Bp * PW * CIN * 2};
  uint32_t box[5] = {(uint32_t)CIN, (uint32_t)WW, (uint32_t)NIMG, 9, 1};
  I2L_TRY(make_tensor_map(&tm, in, 5, dims, str, box, CIN * 2, 2));
  auto kern = conv_pool_kernel<Cfg, CIN, COUT, WW, NIMG, STAGES, NACC, OUT_PARITY>;
  I2L_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
  const int n_tiles = (Bp / NIMG) * (PH / 8) * (PW / WW);
  KernelTimer kt(name, s);
  kern<<<min(n_tiles, sms), 192, Cfg::SMEM, s>>>(tm, wimg, bias, out, B_store, PH, PW, n_tiles);
  I2L_LAUNCH_OK();
  return I2L_OK;
}

}  // namespace

bool cnn_bf16_supported(const i2l_cnn_desc& d) {
  return (d.channels == 1 || d.channels == 3) && d.img_height >= 64 && (d.img_height % 64) == 0 && d.img_width >= 32 &&
         (d.img_width % 32) == 0 && d.img_height <= 1024 && d.img_width <= 8192 && d.n_conv == 3 && d.filters[0] == C1 &&
         d.filters[1] == C2 && d.filters[2] == C3 && d.kernel_size == 3 && d.pool_size == 2 && d.embedding_dim == EMB;
}
size_t cnn_bf16_packed_bytes(const i2l_cnn_desc& d) { return sec_layout(make_geom(d)).total; }

int cnn_bf16_pack(const i2l_cnn_desc& d, const i2l_cnn_params& p, void* section, cudaStream_t s) {
  const Geom G = make_geom(d);
  Sec L = sec_layout(G);
  unsigned char* sec = reinterpret_cast<unsigned char*>(section);
  pack_conv_w_kernel<<<4, 256, 0, s>>>(p.conv_w[0], C1, G.C0, 1, sec + L.w1);
  I2L_LAUNCH_OK();
  pack_conv1_w_f16_kernel<<<4, 256, 0, s>>>(p.conv_w[0], G.C0, sec + L.w1h);
  I2L_LAUNCH_OK();
  pack_conv_w_kernel<<<cdiv(9 * C2 * C1, 256), 256, 0, s>>>(p.conv_w[1], C2, C1, 2, sec + L.w2);
  I2L_LAUNCH_OK();
  pack_conv_w_kernel<<<cdiv(9 * C3 * C2, 256), 256, 0, s>>>(p.conv_w[2], C3, C2, 3, sec + L.w3);
  I2L_LAUNCH_OK();
  I2L_CUDA_OK(cudaMemcpyAsync(sec + L.b1, p.conv_b[0], C1 * 4, cudaMemcpyDeviceToDevice, s));
  I2L_CUDA_OK(cudaMemcpyAsync(sec + L.b2, p.conv_b[1], C2 * 4, cudaMemcpyDeviceToDevice, s));
  I2L_CUDA_OK(cudaMemcpyAsync(sec + L.b3, p.conv_b[2], C3 * 4, cudaMemcpyDeviceToDevice, s));
  I2L_CUDA_OK(cudaMemcpyAsync(sec + L.bfc, p.fc_b, EMB * 4, cudaMemcpyDeviceToDevice, s));
  size_t n = (size_t)EMB * G.FLAT;
  pack_fc_w_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(p.fc_w, reinterpret_cast<__nv_bfloat16*>(sec + L.wfc), G.FLAT,
                                                               G.PH3 * G.PW3);
  I2L_LAUNCH_OK();
  return I2L_OK;
}

size_t cnn_bf16_workspace_bytes(const i2l_cnn_desc& d, int batch) { return carve(make_geom(d), batch, nullptr).bytes; }

// I2L_DEBUG_SYNC=1 (diagnostics build): synchronise after every encoder kernel so that a device fault is attributed
static int dbg_sync(const char* what, cudaStream_t s) {
#ifdef I2L_DIAG
  static const bool on = getenv("I2L_DEBUG_SYNC") != nullptr;
#else
  constexpr bool on = false;   // production library: never synchronises the host
#endif
  if (!on) return I2L_OK;
  cudaError_t e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) { set_error("%s: %s", what, cudaGetErrorString(e)); return I2L_ERR_CUDA; }
  return I2L_OK;
}

template <typename InT, int CIN0>
static int launch_conv1(const CUtensorMap& tm, const unsigned char* w1, const float* b1, __nv_bfloat16* act1, int Bp, int n_tiles,
                        int dbg, const C1Norm& nrm, const Geom& G, int sms, cudaStream_t s) {
  auto kern = conv1_kernel<InT, CIN0>;
  I2L_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C1_SMEM));
  kern<<<min(n_tiles, sms), C1_THREADS, C1_SMEM, s>>>(tm, w1, b1, act1, Bp, n_tiles, dbg, nrm, G.TW, G.TH);
  I2L_LAUNCH_OK();
  return I2L_OK;
}

template <int CIN0, int NG>
static int launch_conv1_u8_ng(const CUtensorMap& tm, const unsigned char* w1h, const float* b1, __nv_bfloat16* act1, int Bp, int n_tiles,
                              const C1Norm& nrm, const Geom& G, int sms, cudaStream_t s) {
  auto kern = conv1_u8_kernel<CIN0, NG>;
  I2L_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C1U_SMEM));
  kern<<<min(n_tiles, sms), c1u_threads(NG), C1U_SMEM, s>>>(tm, w1h, b1, act1, Bp, n_tiles, nrm, G.TW, G.TH);
  I2L_LAUNCH_OK();
  return I2L_OK;
}
template <int CIN0>
static int launch_conv1_u8(const CUtensorMap& tm, const unsigned char* w1h, const float* b1, __nv_bfloat16* act1, int Bp, int n_tiles,
                           const C1Norm& nrm, const Geom& G, int sms, cudaStream_t s) {
#ifdef I2L_DIAG
  const int ng = getenv("I2L_CONV1_NG") ? atoi(getenv("I2L_CONV1_NG")) : 3;     // A-B: builder groups
  if (ng == 2) return launch_conv1_u8_ng<CIN0, 2>(tm, w1h, b1, act1, Bp, n_tiles, nrm, G, sms, s);
  if (ng == 4) return launch_conv1_u8_ng<CIN0, 4>(tm, w1h, b1, act1, Bp, n_tiles, nrm, G, sms, s);
#endif
  return launch_conv1_u8_ng<CIN0, 3>(tm, w1h, b1, act1, Bp, n_tiles, nrm, G, sms, s);
}

int cnn_bf16_fwd(const i2l_cnn_desc& d, const void* section, const void* x, int in_dtype, int B, float* out, void* ws,
                 size_t ws_bytes, cudaStream_t s, const float* norm_a, const float* norm_b) {
  const Geom G = make_geom(d);
  Ws w = carve(G, B, ws);
  if (ws_bytes < w.bytes) { set_error("cnn_bf16_fwd: workspace too small (%zu < %zu)", ws_bytes, w.bytes); return I2L_ERR_WORKSPACE; }
  Sec L = sec_layout(G);
  const unsigned char* sec = reinterpret_cast<const unsigned char*>(section);
  const int sms = num_sms();
  const int Bp = (B + BPAD - 1) / BPAD * BPAD;
  if (Bp != B) {   // the padding images of the multi-image conv3 tiles read zeros (their outputs are never stored)
    I2L_CUDA_OK(cudaMemsetAsync(w.act2, 0, (size_t)Bp * G.PH2 * G.PW2 * C2 * 2, s));
  }
  // ---- conv1: input x (B,C,H,W) fp32 / bf16 / uint8 NCHW read through a 4-D tensor map
  {
    const bool in_bf16 = in_dtype == I2L_IN_BF16, in_u8 = in_dtype == I2L_IN_U8;
    const uint64_t el = in_u8 ? 1 : (in_bf16 ? 2 : 4);
    C1Norm nrm{};
    if (in_u8) for (int c = 0; c < G.C0; ++c) { nrm.a[c] = norm_a[c]; nrm.b[c] = norm_b[c]; }
    CUtensorMap tm;
    uint64_t dims[4] = {(uint64_t)G.W, (uint64_t)G.H, (uint64_t)G.C0, (uint64_t)B};
    uint64_t str[3] = {G.W * el, (uint64_t)G.W * G.H * el, (uint64_t)G.W * G.H * G.C0 * el};
    uint32_t box[4] = {(uint32_t)(in_u8 ? C1In<uint8_t>::PATCH_W : in_bf16 ? C1In<__nv_bfloat16>::PATCH_W : C1In<float>::PATCH_W), PATCH_H,
                       (uint32_t)G.C0, 1};
    I2L_TRY(make_tensor_map(&tm, x, 4, dims, str, box, 0, (int)el));
    const int n_tiles = B * G.TW * G.TH;
#ifdef I2L_DIAG
    const int dbg = getenv("I2L_CONV1_DBG") ? atoi(getenv("I2L_CONV1_DBG")) : 0;
    const bool u8_legacy = getenv("I2L_CONV1_U8_LEGACY") != nullptr;   // A-B: the bf16-operand builder kernel
#else
    const int dbg = 0;
    constexpr bool u8_legacy = false;
#endif
    const unsigned char* w1 = sec + L.w1;
    const float* b1 = reinterpret_cast<const float*>(sec + L.b1);
    KernelTimer kt(in_u8 ? "cnn.conv1_u8in" : in_bf16 ? "cnn.conv1_bf16in" : "cnn.conv1_bf16", s);
    if (G.C0 == 3) {
      if (in_u8 && !u8_legacy) I2L_TRY((launch_conv1_u8<3>(tm, sec + L.w1h, b1, w.act1, Bp, n_tiles, nrm, G, sms, s)));
      else if (in_u8) I2L_TRY((launch_conv1<uint8_t, 3>(tm, w1, b1, w.act1, Bp, n_tiles, dbg, nrm, G, sms, s)));
      else if (in_bf16) I2L_TRY((launch_conv1<__nv_bfloat16, 3>(tm, w1, b1, w.act1, Bp, n_tiles, dbg, nrm, G, sms, s)));
      else I2L_TRY((launch_conv1<float, 3>(tm, w1, b1, w.act1, Bp, n_tiles, dbg, nrm, G, sms, s)));
    } else {
      if (in_u8 && !u8_legacy) I2L_TRY((launch_conv1_u8<1>(tm, sec + L.w1h, b1, w.act1, Bp, n_tiles, nrm, G, sms, s)));
      else if (in_u8) I2L_TRY((launch_conv1<uint8_t, 1>(tm, w1, b1, w.act1, Bp, n_tiles, dbg, nrm, G, sms, s)));
      else if (in_bf16) I2L_TRY((launch_conv1<__nv_bfloat16, 1>(tm, w1, b1, w.act1, Bp, n_tiles, dbg, nrm, G, sms, s)));
      else I2L_TRY((launch_conv1<float, 1>(tm, w1, b1, w.act1, Bp, n_tiles, dbg, nrm, G, sms, s)));
    }
  }
  I2L_TRY(dbg_sync("conv1", s));
  // ---- conv2: input planes [4][16][Bp][80][32] -> act2 planes [4][8][Bp][40][64]
  {
    const float* b2 = reinterpret_cast<const float*>(sec + L.b2);
    if (G.WW2 == 16) I2L_TRY((launch_conv_pool<C1, C2, 16, 8, 2, true>(w.act1, sec + L.w2, b2, w.act2, Bp, Bp, G.PH2, G.PW2, sms, "cnn.conv2_bf16", s)));
    else I2L_TRY((launch_conv_pool<C1, C2, 8, 8, 2, true>(w.act1, sec + L.w2, b2, w.act2, Bp, Bp, G.PH2, G.PW2, sms, "cnn.conv2_bf16", s)));
  }
  I2L_TRY(dbg_sync("conv2", s));
  // ---- conv3: input planes [4][8][Bp][40][64] -> act3 [B][8][40][128]
  {
    const float* b3 = reinterpret_cast<const float*>(sec + L.b3);
    if (G.WW3 == 16) I2L_TRY((launch_conv_pool<C2, C3, 16, 4, 1, false>(w.act2, sec + L.w3, b3, w.act3, Bp, B, G.PH3, G.PW3, sms, "cnn.conv3_bf16", s)));
    else if (G.WW3 == 8) I2L_TRY((launch_conv_pool<C2, C3, 8, 4, 1, false>(w.act2, sec + L.w3, b3, w.act3, Bp, B, G.PH3, G.PW3, sms, "cnn.conv3_bf16", s)));
    else I2L_TRY((launch_conv_pool<C2, C3, 4, 4, 1, false>(w.act2, sec + L.w3, b3, w.act3, Bp, B, G.PH3, G.PW3, sms, "cnn.conv3_bf16", s)));
  }
  I2L_TRY(dbg_sync("conv3", s));
  // ---- fc
  {
    CUtensorMap tmA, tmW;
    const uint64_t FLAT = (uint64_t)G.FLAT;
    uint64_t dA[2] = {FLAT, (uint64_t)B}; uint64_t sA[1] = {FLAT * 2ull}; uint32_t bA[2] = {FC_BK, FC_BM};
    uint64_t dW[2] = {FLAT, EMB}; uint32_t bW[2] = {FC_BK, FC_BN};
    I2L_TRY(make_tensor_map(&tmA, w.act3, 2, dA, sA, bA, 128, 2));
    I2L_TRY(make_tensor_map(&tmW, sec + L.wfc, 2, dW, sA, bW, 128, 2));
    I2L_CUDA_OK(cudaFuncSetAttribute(fc_splitk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FC_SMEM));
    const int kbs = G.FLAT / FC_BK / w.splits;
    KernelTimer kt("cnn.fc_bf16", s);
    fc_splitk_kernel<<<dim3(cdiv(B, FC_BM), w.splits), 192, FC_SMEM, s>>>(tmA, tmW, w.partial, B, kbs);
    I2L_LAUNCH_OK();
    size_t n = (size_t)B * EMB;
    fc_reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(w.partial, reinterpret_cast<const float*>(sec + L.bfc), out, B, w.splits);
    I2L_LAUNCH_OK();
  }
  return I2L_OK;
}

}  // namespace i2l
