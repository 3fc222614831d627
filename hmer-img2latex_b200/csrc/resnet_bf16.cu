// bf16 ResNet trunk (ResNetEncoder.forward, model/encoder.py:231-249; torchvision resnet18/34/50/101/152
// minus fc, eval-mode BatchNorm folded into the convs) as tcgen05 implicit-GEMM kernels.
//
//   activations  bf16 NHWC [B][H][W][C] in HBM
//   conv         one persistent warp-specialised kernel for every 1x1 / 3x3, stride 1 / 2 conv:
//                M tile = 128 consecutive output pixels (linear over n, ho, wo), N tile = 64 / 128
//                output channels, K loop over (kh, kw, 64-channel block).  The A operand tile of a
//                (tap, channel block) is ONE im2col-mode TMA load (cp.async.bulk.tensor...im2col):
//                the TMA unit walks the 128 output positions across row / image boundaries, applies
//                the conv stride and zero-fills the padding, and writes a K-major SWIZZLE_128B tile.
//                Weights are pre-packed as shared-memory images (BN folded, bf16, swizzled), one
//                bulk copy per K block.  Accumulators are double buffered in TMEM; the epilogue adds
//                bias (+ residual), applies ReLU and stores bf16 NHWC.
//   stem         7x7 stride-2 conv on 3 channels: the image is re-laid-out as NHWC4 bf16 with zero columns
//                physically padded left / right and viewed as pixel PAIRS (8 elements, 16 B).  Output
//                column wo reads the 4 consecutive pairs wo-2..wo+1 = 64 contiguous bytes, so the tensor
//                map describes OVERLAPPING 32-"channel" pixels (W stride 16 B): one im2col load per kh
//                fetches a K = 32 operand block (7 x 32 = 224, 147 useful), SWIZZLE_64B.  Same kernel.
//   maxpool 3x3/2, global average pool: bandwidth kernels; embedding Linear + ReLU: fp32 GEMM.
#include "resnet_bf16.cuh"
#include "gemm_bf16.cuh"
#include <stdlib.h>

namespace i2l {
namespace {

using namespace tc;

constexpr int BM = 128;                       // output pixels per tile
constexpr int STEM_PADL = 2, STEM_PADR = 2;   // zero pixel pairs stored left / right of every image row

// N tile: the tcgen05 pipe runs a 128 x N x 16 MMA in roughly 50 + 0.75 N cycles (measured: N = 64 -> ~100, N = 128 -> ~150),
// so wide tiles are what approaches the tensor peak
__host__ __device__ constexpr int conv_bn(int co) { return co == 64 ? 64 : (co % 256 == 0 ? 256 : 128); }

// ------------------------------------------------------------------ packing
// generic conv: image [n_nt][n_kb][BN rows][64 k] bf16, rows 128 B, SWIZZLE_128B; k block kb = tap * (Ci/64) + cb
__global__ void pack_conv_kernel(const float* __restrict__ w /*(Co,Ci,KH,KW) folded*/, int Co, int Ci, int KH, int KW, int BN,
                                 unsigned char* __restrict__ dst) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t tot = (size_t)Co * Ci * KH * KW;
  if (i >= tot) return;
  const int c = (int)(i % Ci);
  const int tap = (int)((i / Ci) % (KH * KW));
  const int co = (int)(i / ((size_t)Ci * KH * KW));
  const int nkb = KH * KW * (Ci / 64);
  const int kb = tap * (Ci / 64) + c / 64, kc = c % 64;
  const int nt = co / BN, r = co % BN;
  const float v = w[((size_t)co * Ci + c) * KH * KW + tap];
  const size_t off = ((size_t)nt * nkb + kb) * BN * 128 + swz_off(r, kc / 8, 128) + (kc % 8) * 2;
  *reinterpret_cast<__nv_bfloat16*>(dst + off) = __float2bfloat16(v);
}

// stem: image [7 kh][64 co][32 k] bf16, rows 64 B, SWIZZLE_64B; k = 8 j + 4 p + c with j = pair tap, p = pixel of
// the pair, c = channel of 4; input column w = 2 (wo - 2 + j) + p  ->  kw = 2 j + p - 1
__global__ void pack_stem_kernel(const float* __restrict__ w /*(64,3,7,7) folded*/, unsigned char* __restrict__ dst) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 7 * 64 * 32) return;
  const int k = i & 31, co = (i >> 5) & 63, kh = i >> 11;
  const int j = k >> 3, p = (k >> 2) & 1, c = k & 3, kw = 2 * j + p - 1;
  float v = 0.f;
  if (c < 3 && kw >= 0 && kw < 7) v = w[((co * 3 + c) * 7 + kh) * 7 + kw];
  *reinterpret_cast<__nv_bfloat16*>(dst + (size_t)kh * (64 * 64) + swz_off(co, k / 8, 64) + (k % 8) * 2) = __float2bfloat16(v);
}

// ------------------------------------------------------------------ the conv kernel
struct ConvArgs {
  const unsigned char* wimg;
  const float* bias;            // [Co] folded BN shift
  const __nv_bfloat16* res;     // residual (same shape as out) or null
  __nv_bfloat16* out;           // [P][Co]
  int P, HoWo, Wo, Co;
  int KH, KW, cblocks;          // taps and KB-channel blocks (stem: KW = 1, cblocks = 1, KB = 32)
  int stride_w, stride_h, pad_w, pad_h;
  int relu, n_mt, n_nt;
};

constexpr int WRES_MAX_KB = 9;                  // resident-weight variant: up to 9 K blocks (3x3, 64 channels) of a 64-wide N tile

// KB = K elements per pipeline stage: 64 (generic, SWIZZLE_128B rows) or 32 (stem, SWIZZLE_64B).
// WRES: the whole weight image of the (single) N tile stays in shared memory for the lifetime of the CTA and a
// stage carries only the A tile -- for the Cout = 64 layers the weights are a third of the L2 -> SM traffic, and
// those layers sit at the chip's L2 throughput cap.
template <int BN, int KB, bool WRES>
struct IgCfg {
  static constexpr int ROWB = KB * 2;
  static constexpr int A_ST = BM * ROWB;
  static constexpr int B_ST = BN * ROWB;
  static constexpr int STAGE = WRES ? A_ST : A_ST + B_ST;
  static constexpr int STAGES = KB == 32 ? 12 : (BN == 256 ? 4 : (BN == 128 ? 6 : 8));
  static constexpr int OFF_W = STAGES * STAGE;                       // resident weights (WRES)
  static constexpr int OFF_BAR = OFF_W + (WRES ? WRES_MAX_KB * B_ST : 0);
  static constexpr int SMEM = OFF_BAR + 256;
  static constexpr int TMEM_COLS = 2 * BN;
  static_assert(STAGE % 1024 == 0, "stage alignment");
  static_assert(SMEM <= 232448, "shared memory budget");
};

template <int BN, int KB, bool WRES>
__global__ void __launch_bounds__(192, 1) conv_igemm_kernel(const __grid_constant__ CUtensorMap tmA, const ConvArgs a) {
  using Cfg = IgCfg<BN, KB, WRES>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar = sbase + Cfg::OFF_BAR;
  auto FULL = [&](int s) { return bar + 8u * s; };
  auto EMPTY = [&](int s) { return bar + 8u * (STAGES + s); };
  auto TFULL = [&](int i) { return bar + 8u * (2 * STAGES + i); };
  auto TEMPTY = [&](int i) { return bar + 8u * (2 * STAGES + 2 + i); };
  const uint32_t WBAR = bar + 8u * (2 * STAGES + 4);
  uint32_t* misc = reinterpret_cast<uint32_t*>(smem + Cfg::OFF_BAR + 8 * (2 * STAGES + 5));
  if ((sbase & 1023u) != 0) __trap();
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(FULL(s), 1); mbar_init(EMPTY(s), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(TFULL(i), 1); mbar_init(TEMPTY(i), 128); }
    mbar_init(WBAR, 1);
    mbar_fence_init();
    tma_prefetch_desc(&tmA);
  }
  if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(smem_u32(misc));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = misc[0];
  const int n_tiles = a.n_mt * a.n_nt;
  const int nkb = a.KH * a.KW * a.cblocks;              // K blocks (= pipeline stages) per tile

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      if constexpr (WRES) {                                  // n_nt == 1: one weight image for every tile of this CTA
        mbar_arrive_expect_tx(WBAR, (uint32_t)(nkb * Cfg::B_ST));
        for (int kb = 0; kb < nkb; ++kb) bulk_g2s(sbase + Cfg::OFF_W + kb * Cfg::B_ST, a.wimg + (size_t)kb * Cfg::B_ST, Cfg::B_ST, WBAR);
      }
      int stage = 0; uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int mt = tile / a.n_nt, nt = tile % a.n_nt;
        const int p0 = mt * BM;
        const int n0 = p0 / a.HoWo, rem = p0 % a.HoWo;
        const int w0 = (rem % a.Wo) * a.stride_w - a.pad_w, h0 = (rem / a.Wo) * a.stride_h - a.pad_h;
        {
          const unsigned char* wt = a.wimg + (size_t)nt * nkb * Cfg::B_ST;
          int kb = 0;
          for (int kh = 0; kh < a.KH; ++kh)
            for (int kw = 0; kw < a.KW; ++kw)
              for (int cb = 0; cb < a.cblocks; ++cb, ++kb) {
                mbar_wait(EMPTY(stage), ph ^ 1);
                mbar_arrive_expect_tx(FULL(stage), Cfg::STAGE);
                const uint32_t dst = sbase + stage * Cfg::STAGE;
                tma_load_im2col(dst, &tmA, cb * KB, w0, h0, n0, kw, kh, FULL(stage));
                if constexpr (!WRES) bulk_g2s(dst + Cfg::A_ST, wt + (size_t)kb * Cfg::B_ST, Cfg::B_ST, FULL(stage));
                if (++stage == STAGES) { stage = 0; ph ^= 1; }
              }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t IDESC = idesc_bf16(128, BN);
    const int nstages = nkb;
    if constexpr (WRES) { mbar_wait(WBAR, 0); tc_fence_after(); }
    int stage = 0; uint32_t ph = 0; int acc = 0; uint32_t aph = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      mbar_wait(TEMPTY(acc), aph ^ 1);
      tc_fence_after();
      const uint32_t d = tmem + (uint32_t)(acc * BN);
      for (int kb = 0; kb < nstages; ++kb) {
        mbar_wait(FULL(stage), ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = sbase + stage * Cfg::STAGE, sb = WRES ? sbase + Cfg::OFF_W + kb * Cfg::B_ST : sa + Cfg::A_ST;
          {
            const uint64_t ad = desc_base(sa, Cfg::ROWB), bd = desc_base(sb, Cfg::ROWB);
#pragma unroll
            for (int ks = 0; ks < KB / 16; ++ks) tc_mma_ss(d, ad + (uint64_t)((ks * 32) >> 4), bd + (uint64_t)((ks * 32) >> 4), IDESC, (kb | ks) ? 1u : 0u);
          }
          tc_commit(EMPTY(stage));
          if (kb == nstages - 1) tc_commit(TFULL(acc));
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; ph ^= 1; }
      }
      if (++acc == 2) { acc = 0; aph ^= 1; }
    }
  } else {
    // ===================== epilogue warps 2..5 =====================
    const int q = warp & 3;
    const int m = 32 * q + lane;
    const uint32_t lane_addr = (uint32_t)(32 * q) << 16;
    int acc = 0; uint32_t aph = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int mt = tile / a.n_nt, nt = tile % a.n_nt;
      const int p = mt * BM + m;
      const bool valid = p < a.P;
      const size_t o = (size_t)p * a.Co + (size_t)nt * BN;
      const float* bias = a.bias + nt * BN;
      mbar_wait(TFULL(acc), aph);
      tc_fence_after();
      const uint32_t ta = tmem + lane_addr + (uint32_t)(acc * BN);
#pragma unroll 1
      for (int cc = 0; cc < BN / 32; ++cc) {
        uint32_t r[32];
        tc_ld16_nowait(ta + cc * 32, r);
        tc_ld16_nowait(ta + cc * 32 + 16, r + 16);
        uint4 rs[4] = {};
        if (a.res != nullptr && valid) {
          const uint4* rp = reinterpret_cast<const uint4*>(a.res + o + cc * 32);
#pragma unroll
          for (int i = 0; i < 4; ++i) rs[i] = __ldg(rp + i);
        }
        tc_wait_ld();
        const uint32_t* rw = reinterpret_cast<const uint32_t*>(rs);
        uint32_t ov[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float v0 = __uint_as_float(r[2 * i]) + __ldg(bias + cc * 32 + 2 * i);
          float v1 = __uint_as_float(r[2 * i + 1]) + __ldg(bias + cc * 32 + 2 * i + 1);
          if (a.res != nullptr) {
            __nv_bfloat162 rr = *reinterpret_cast<const __nv_bfloat162*>(&rw[i]);
            v0 += __low2float(rr); v1 += __high2float(rr);
          }
          if (a.relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
          __nv_bfloat162 h2 = __floats2bfloat162_rn(v0, v1);
          ov[i] = *reinterpret_cast<uint32_t*>(&h2);
        }
        if (valid) {
          uint4* d4 = reinterpret_cast<uint4*>(a.out + o + cc * 32);
#pragma unroll
          for (int i = 0; i < 4; ++i) d4[i] = make_uint4(ov[4 * i], ov[4 * i + 1], ov[4 * i + 2], ov[4 * i + 3]);
        }
      }
      tc_fence_before();
      mbar_arrive(TEMPTY(acc));
      if (++acc == 2) { acc = 0; aph ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<Cfg::TMEM_COLS>(tmem);
}

// ------------------------------------------------------------------ stem, row-window variant
// The 7x7/2 stem WITHOUT an im2col: for output row ho and filter row kh, the 4 pixel pairs that output column wo
// needs (wo-2 .. wo+1, physical wo .. wo+3) are 64 contiguous bytes of ONE image row, and column wo+1 needs the
// same bytes shifted by 16.  A no-swizzle K-major UMMA descriptor with a 16-byte ROW pitch (SBO = 128) and a
// 16-byte K-chunk pitch (LBO = 16) reads exactly those overlapping windows: row m, chunk c -> byte 16 (m + c).
// So the A operand of (tile, kh) is just a 1 KB slice of the image row, fetched with one plain bulk copy; a
// tile is 2 output rows x 61 columns (M = 128: rows 0-63 / 64-127 are the two slices, 1 KB apart; columns 61-63
// of each half read past their slice and are masked).  14 bulk copies of 1 KB and 14 MMAs (128x64x16) per tile,
// instead of 7 im2col loads of 128 unaligned 64-byte windows (the TMA unit processes ~0.2 windows per cycle).
constexpr int SR_COLS = 61;                      // valid output columns per half tile
constexpr int SR_STAGE = 2048, SR_STAGES = 28;   // one stage = (kh, both output rows)
constexpr int SR_OFF_W = SR_STAGES * SR_STAGE;   // resident weights: 7 x [64 co][32 k] SWIZZLE_64B
constexpr int SR_OFF_BAR = SR_OFF_W + 7 * 4096;
constexpr int SR_SMEM = SR_OFF_BAR + 512;

struct StemArgs {
  const unsigned char* x4;       // [B][H][Wpp pairs][8] bf16 (zero pairs stored left / right)
  const unsigned char* zero_row; // >= 1 KB of zeros: source of out-of-image rows
  const unsigned char* wimg;
  const float* bias;
  __nv_bfloat16* out;            // [B][Ho][Wo][64]
  int B, H, Wpp, Ho, Wo, tiles_w, n_tiles;
};

__device__ __forceinline__ uint64_t desc_nosw(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) |
         ((uint64_t)1 << 46);
}

__global__ void __launch_bounds__(256, 1) stem_rows_kernel(const StemArgs a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar = sbase + SR_OFF_BAR;
  auto FULL = [&](int s) { return bar + 8u * s; };
  auto EMPTY = [&](int s) { return bar + 8u * (SR_STAGES + s); };
  auto TFULL = [&](int i) { return bar + 8u * (2 * SR_STAGES + i); };
  auto TEMPTY = [&](int i) { return bar + 8u * (2 * SR_STAGES + 2 + i); };
  const uint32_t WBAR = bar + 8u * (2 * SR_STAGES + 4);
  uint32_t* misc = reinterpret_cast<uint32_t*>(smem + SR_OFF_BAR + 8 * (2 * SR_STAGES + 5));
  if ((sbase & 1023u) != 0) __trap();
  if (tid == 0) {
    for (int s = 0; s < SR_STAGES; ++s) { mbar_init(FULL(s), 2); mbar_init(EMPTY(s), 1); }     // two producers per stage
    for (int i = 0; i < 2; ++i) { mbar_init(TFULL(i), 1); mbar_init(TEMPTY(i), 128); }
    mbar_init(WBAR, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<128>(smem_u32(misc));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = misc[0];
  const int hop = a.Ho / 2;                                    // output row pairs per image

  if (warp == 0 || warp == 6 || warp == 7) {
    // ===================== producers: plain bulk copies of image-row slices =====================
    // a bulk copy costs ~100+ cycles of issue in its thread: warp 0 feeds output row 0 of every tile, warp 6 row 1
    // (warp 7 only pads the block to whole warp quads and idles)
    const int seg = warp == 0 ? 0 : 1;
    if (warp != 7 && elect_one()) {
      if (warp == 0) {
        mbar_arrive_expect_tx(WBAR, 7 * 4096);
        for (int kh = 0; kh < 7; ++kh) bulk_g2s(sbase + SR_OFF_W + kh * 4096, a.wimg + kh * 4096, 4096, WBAR);
      }
      int stage = 0; uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
        const int tw = tile % a.tiles_w, rp = (tile / a.tiles_w) % hop, n = tile / (a.tiles_w * hop);
        const int wo0 = tw * SR_COLS, ho0 = 2 * rp;
        for (int kh = 0; kh < 7; ++kh) {
          mbar_wait(EMPTY(stage), ph ^ 1);
          mbar_arrive_expect_tx(FULL(stage), SR_STAGE / 2);
          const uint32_t dst = sbase + stage * SR_STAGE;
          const int h = 2 * (ho0 + seg) - 3 + kh;
          const unsigned char* src = (h >= 0 && h < a.H) ? a.x4 + (((size_t)n * a.H + h) * a.Wpp + wo0) * 16 : a.zero_row;
          bulk_g2s(dst + seg * 1024, src, 1024, FULL(stage));
          if (++stage == SR_STAGES) { stage = 0; ph ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t IDESC = idesc_bf16(128, 64);
    mbar_wait(WBAR, 0);
    tc_fence_after();
    int stage = 0; uint32_t ph = 0; int acc = 0; uint32_t aph = 0;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
      mbar_wait(TEMPTY(acc), aph ^ 1);
      tc_fence_after();
      const uint32_t d = tmem + (uint32_t)(acc * 64);
      for (int kh = 0; kh < 7; ++kh) {
        mbar_wait(FULL(stage), ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = sbase + stage * SR_STAGE;
          const uint64_t bd = desc_base(sbase + SR_OFF_W + kh * 4096, 64);
#pragma unroll
          for (int ks = 0; ks < 2; ++ks)
            tc_mma_ss(d, desc_nosw(sa + ks * 32, 16, 128), bd + (uint64_t)((ks * 32) >> 4), IDESC, (kh | ks) ? 1u : 0u);
          tc_commit(EMPTY(stage));
          if (kh == 6) tc_commit(TFULL(acc));
        }
        __syncwarp();
        if (++stage == SR_STAGES) { stage = 0; ph ^= 1; }
      }
      if (++acc == 2) { acc = 0; aph ^= 1; }
    }
  } else if (warp < 6) {
    // ===================== epilogue warps 2..5: BN shift + ReLU -> bf16 NHWC =====================
    const int q = warp & 3;
    const int m = 32 * q + lane;
    const int seg = m >> 6, i = m & 63;
    const uint32_t lane_addr = (uint32_t)(32 * q) << 16;
    int acc = 0; uint32_t aph = 0;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
      const int tw = tile % a.tiles_w, rp = (tile / a.tiles_w) % hop, n = tile / (a.tiles_w * hop);
      const int wo = tw * SR_COLS + i, ho = 2 * rp + seg;
      const bool valid = i < SR_COLS && wo < a.Wo;
      __nv_bfloat16* dst = a.out + (((size_t)n * a.Ho + ho) * a.Wo + wo) * 64;
      mbar_wait(TFULL(acc), aph);
      tc_fence_after();
      const uint32_t ta = tmem + lane_addr + (uint32_t)(acc * 64);
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        uint32_t r[32];
        tc_ld16_nowait(ta + cc * 32, r);
        tc_ld16_nowait(ta + cc * 32 + 16, r + 16);
        tc_wait_ld();
        uint32_t ov[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const float v0 = fmaxf(__uint_as_float(r[2 * k]) + __ldg(a.bias + cc * 32 + 2 * k), 0.f);
          const float v1 = fmaxf(__uint_as_float(r[2 * k + 1]) + __ldg(a.bias + cc * 32 + 2 * k + 1), 0.f);
          __nv_bfloat162 h2 = __floats2bfloat162_rn(v0, v1);
          ov[k] = *reinterpret_cast<uint32_t*>(&h2);
        }
        if (valid) {
          uint4* d4 = reinterpret_cast<uint4*>(dst + cc * 32);
#pragma unroll
          for (int k = 0; k < 4; ++k) d4[k] = make_uint4(ov[4 * k], ov[4 * k + 1], ov[4 * k + 2], ov[4 * k + 3]);
        }
      }
      tc_fence_before();
      mbar_arrive(TEMPTY(acc));
      if (++acc == 2) { acc = 0; aph ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<128>(tmem);
}

// ------------------------------------------------------------------ bandwidth kernels
// (B,3,H,W) fp32 NCHW -> [B][H][Wp][4] bf16, Wp = W + 2 (PADL + PADR) pixels per row (zero columns), channel 3 = 0
__global__ void nchw_to_nhwc4_kernel(const float* __restrict__ x, uint2* __restrict__ y, int H, int W, int Wp, size_t total /*B*H*Wp*/) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int wp = (int)(i % Wp), w = wp - 2 * STEM_PADL;
  const size_t bh = i / Wp, b = bh / H, h = bh % H;
  uint2 o = make_uint2(0u, 0u);
  if (w >= 0 && w < W) {
    const size_t HW = (size_t)H * W;
    const float* s = x + b * 3 * HW + h * W + w;
    __nv_bfloat162 lo = __floats2bfloat162_rn(__ldg(s), __ldg(s + HW));
    __nv_bfloat162 hi = __floats2bfloat162_rn(__ldg(s + 2 * HW), 0.f);
    o = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
  }
  y[i] = o;
}

// MaxPool2d(3, stride 2, padding 1) on NHWC bf16; one thread = one output pixel x 8 channels
__global__ void maxpool3s2_nhwc_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int Hi, int Wi, int Ho,
                                       int Wo, int C, size_t total /*B*Ho*Wo*C/8*/) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c8 = C / 8;
  const int cg = (int)(i % c8);
  size_t pix = i / c8;
  const int wo = (int)(pix % Wo), ho = (int)((pix / Wo) % Ho);
  const size_t b = pix / ((size_t)Wo * Ho);
  float m[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) m[k] = -INFINITY;
#pragma unroll
  for (int dh = 0; dh < 3; ++dh) {
    const int h = 2 * ho - 1 + dh;
    if (h < 0 || h >= Hi) continue;
#pragma unroll
    for (int dw = 0; dw < 3; ++dw) {
      const int w = 2 * wo - 1 + dw;
      if (w < 0 || w >= Wi) continue;
      uint4 v = __ldg(reinterpret_cast<const uint4*>(x + ((b * Hi + h) * Wi + w) * C) + cg);
      const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
      for (int k = 0; k < 4; ++k) { m[2 * k] = fmaxf(m[2 * k], __low2float(h2[k])); m[2 * k + 1] = fmaxf(m[2 * k + 1], __high2float(h2[k])); }
    }
  }
  uint32_t o[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) { __nv_bfloat162 h2 = __floats2bfloat162_rn(m[2 * k], m[2 * k + 1]); o[k] = *reinterpret_cast<uint32_t*>(&h2); }
  reinterpret_cast<uint4*>(y + pix * C)[cg] = make_uint4(o[0], o[1], o[2], o[3]);
}

// AdaptiveAvgPool2d(1) on NHWC bf16 -> bf16 [B][C] (fp32 sums; the A operand of the embedding GEMM)
__global__ void avgpool_nhwc_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int HW, int C) {
  const int b = blockIdx.x;
  for (int c2 = threadIdx.x; c2 < C / 2; c2 += blockDim.x) {
    float s0 = 0.f, s1 = 0.f;
    const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(x + (size_t)b * HW * C) + c2;
    for (int i = 0; i < HW; ++i) { __nv_bfloat162 v = p[(size_t)i * (C / 2)]; s0 += __low2float(v); s1 += __high2float(v); }
    reinterpret_cast<__nv_bfloat162*>(y + (size_t)b * C)[c2] = __floats2bfloat162_rn(s0 / (float)HW, s1 / (float)HW);
  }
}

// ------------------------------------------------------------------ layout of the bf16 section
size_t conv_img_bytes(const RConv& c) { return (size_t)c.co * c.ci * c.k * c.k * 2; }   // generic convs: no padding elements
constexpr size_t STEM_IMG_BYTES = 7 * 64 * 64;

struct Sec { std::vector<size_t> off; size_t fc; size_t total; };
Sec sec_layout(const RNet& n, int E) {
  Sec s; size_t o = 0;
  for (size_t i = 0; i < n.convs.size(); ++i) {
    s.off.push_back(o);
    o = align_up(o + (i == 0 ? STEM_IMG_BYTES : conv_img_bytes(n.convs[i])), 1024);
  }
  s.fc = o;                                              // embedding_layer.weight (E, feat) bf16 row-major
  o = align_up(o + (size_t)E * n.feat * 2, 1024);
  s.total = o;
  return s;
}

__global__ void f32_to_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = __float2bfloat16(src[i]);
}

int out_dim(int v, int k, int st, int p) { return (v + 2 * p - k) / st + 1; }

struct Ws { __nv_bfloat16* x4; unsigned char* zero_row; __nv_bfloat16* buf[5]; __nv_bfloat16* pooled; size_t bytes; };
Ws carve(const RNet& n, int B, int H, int W, void* ws) {
  Arena a(ws, (size_t)-1);
  Ws w{};
  w.x4 = a.take<__nv_bfloat16>((size_t)B * H * (W + 2 * (STEM_PADL + STEM_PADR)) * 4 + 1024);   // + slack: the last row slice over-reads
  w.zero_row = a.take<unsigned char>(2048);
  const size_t act = resnet_max_act(n, H, W) * (size_t)B;
  for (int i = 0; i < 5; ++i) w.buf[i] = a.take<__nv_bfloat16>(act);
  w.pooled = a.take<__nv_bfloat16>((size_t)B * n.feat);
  w.bytes = align_up(a.off, 256);
  return w;
}

template <int BN, int KB, bool WRES = false>
int launch_conv(const CUtensorMap& tm, const ConvArgs& a, cudaStream_t s) {
  using Cfg = IgCfg<BN, KB, WRES>;
  auto kern = conv_igemm_kernel<BN, KB, WRES>;
  static thread_local int attr_dev = -1;      // the attribute is per device; set it once per (thread, device)
  int dev = 0;
  I2L_CUDA_OK(cudaGetDevice(&dev));
  if (attr_dev != dev) {
    I2L_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    attr_dev = dev;
  }
  const int n_tiles = a.n_mt * a.n_nt;
  kern<<<std::min(n_tiles, num_sms()), 192, Cfg::SMEM, s>>>(tm, a);
  I2L_LAUNCH_OK();
  return I2L_OK;
}

}  // namespace

bool resnet_bf16_supported(const i2l_resnet_desc& d, int img_width) {
  return d.precision == I2L_BF16 && d.img_height >= 2 && img_width >= 2 && (img_width % 2) == 0;
}

size_t resnet_bf16_packed_bytes(const RNet& n, int E) { return sec_layout(n, E).total; }

int resnet_bf16_pack(const RNet& n, int E, const float* folded /* fp32 packed region */, const float* fc_w, void* section,
                     cudaStream_t s) {
  Sec L = sec_layout(n, E);
  {
    const size_t tot = (size_t)E * n.feat;
    f32_to_bf16_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(fc_w, reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<unsigned char*>(section) + L.fc), tot);
    I2L_LAUNCH_OK();
  }
  unsigned char* sec = reinterpret_cast<unsigned char*>(section);
  pack_stem_kernel<<<cdiv(7 * 64 * 32, 256), 256, 0, s>>>(folded + n.convs[0].w_off, sec + L.off[0]);
  I2L_LAUNCH_OK();
  for (size_t i = 1; i < n.convs.size(); ++i) {
    const RConv& c = n.convs[i];
    I2L_REQUIRE(c.ci % 64 == 0 && c.co % 64 == 0, "resnet_bf16_pack: channel counts must be multiples of 64");
    const size_t tot = (size_t)c.co * c.ci * c.k * c.k;
    pack_conv_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(folded + c.w_off, c.co, c.ci, c.k, c.k, conv_bn(c.co), sec + L.off[i]);
    I2L_LAUNCH_OK();
  }
  return I2L_OK;
}

size_t resnet_bf16_workspace_bytes(const RNet& n, int batch, int H, int W) { return carve(n, batch, H, W, nullptr).bytes; }

int resnet_bf16_fwd(const RNet& n, const i2l_resnet_desc& d, const float* folded, const void* section, const float* fc_w,
                    const float* fc_b, const float* x, int B, int W, float* out, void* ws, size_t ws_bytes, cudaStream_t s) {
  const int H = d.img_height;
  Ws w = carve(n, B, H, W, ws);
  if (ws_bytes < w.bytes) { set_error("resnet_bf16_fwd: workspace too small (%zu < %zu)", ws_bytes, w.bytes); return I2L_ERR_WORKSPACE; }
  Sec L = sec_layout(n, d.embedding_dim);
  const unsigned char* sec = reinterpret_cast<const unsigned char*>(section);
  char tag[40];
  snprintf(tag, sizeof tag, "resnet%d", d.depth);
  // ---- input re-layout
  {
    const int Wp = W + 2 * (STEM_PADL + STEM_PADR);
    const size_t tot = (size_t)B * H * Wp;
    KernelTimer kt("rn.nchw_to_nhwc4", s);
    nchw_to_nhwc4_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(x, reinterpret_cast<uint2*>(w.x4), H, W, Wp, tot);
    I2L_LAUNCH_OK();
  }
  // ---- stem: conv 7x7/2 + bn + relu in pixel-pair space
  int h = out_dim(H, 7, 2, 3), wd = out_dim(W, 7, 2, 3);    // = H/2 (ceil), W/2
  {
    CUtensorMap tm;
    // overlapping view: "pixel" wo = the 32 elements (4 pairs) starting at stored pair wo = logical pair wo - PADL
    const int Wpp = W / 2 + STEM_PADL + STEM_PADR;      // stored pairs per row
    I2L_TRY(make_im2col_map_strided(&tm, w.x4, 32, W / 2, H, B, 16, (uint64_t)Wpp * 16, (uint64_t)H * Wpp * 16, 0, -3, 0, -3, 32, BM, 1, 2, 64));
    ConvArgs a{};
    a.wimg = sec + L.off[0]; a.bias = folded + n.convs[0].b_off; a.res = nullptr; a.out = w.buf[1];
    a.P = B * h * wd; a.HoWo = h * wd; a.Wo = wd; a.Co = 64; a.KH = 7; a.KW = 1; a.cblocks = 1;
    a.stride_w = 1; a.stride_h = 2; a.pad_w = 0; a.pad_h = 3; a.relu = 1; a.n_mt = cdiv(a.P, BM); a.n_nt = 1;
    KernelTimer kt("rn.stem_conv7x7", s);
#ifdef I2L_DIAG
    const bool stem_rows = (h % 2) == 0 && getenv("I2L_STEM_IM2COL") == nullptr;   // A/B switch of the diagnostics build
#else
    const bool stem_rows = (h % 2) == 0;
#endif
    if (stem_rows) {
      // row-window stem (stem_rows_kernel): needs an even number of output rows
      I2L_CUDA_OK(cudaMemsetAsync(w.zero_row, 0, 2048, s));
      StemArgs sa{};
      sa.x4 = reinterpret_cast<const unsigned char*>(w.x4); sa.zero_row = w.zero_row; sa.wimg = sec + L.off[0];
      sa.bias = folded + n.convs[0].b_off; sa.out = w.buf[1];
      sa.B = B; sa.H = H; sa.Wpp = Wpp; sa.Ho = h; sa.Wo = wd; sa.tiles_w = cdiv(wd, SR_COLS); sa.n_tiles = B * (h / 2) * sa.tiles_w;
      I2L_CUDA_OK(cudaFuncSetAttribute(stem_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SR_SMEM));
      stem_rows_kernel<<<std::min(sa.n_tiles, num_sms()), 256, SR_SMEM, s>>>(sa);
      I2L_LAUNCH_OK();
    } else {
      I2L_TRY((launch_conv<64, 32>(tm, a, s)));
    }
  }
  // ---- maxpool 3x3/2
  {
    const int ho = out_dim(h, 3, 2, 1), wo = out_dim(wd, 3, 2, 1);
    const size_t tot = (size_t)B * ho * wo * (64 / 8);
    KernelTimer kt("rn.maxpool", s);
    maxpool3s2_nhwc_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(w.buf[1], w.buf[0], h, wd, ho, wo, 64, tot);
    I2L_LAUNCH_OK();
    h = ho; wd = wo;
  }
  auto run = [&](int idx, const __nv_bfloat16* in, __nv_bfloat16* o, int hi, int wi, const __nv_bfloat16* res, int relu) -> int {
    const RConv& c = n.convs[idx];
    const int ho = out_dim(hi, c.k, c.stride, c.pad), wo = out_dim(wi, c.k, c.stride, c.pad);
    CUtensorMap tm;
    I2L_TRY(make_im2col_map(&tm, in, c.ci, wi, hi, B, -c.pad, -c.pad, c.pad - (c.k - 1), c.pad - (c.k - 1), 64, BM, c.stride, c.stride, 128));
    ConvArgs a{};
    a.wimg = sec + L.off[idx]; a.bias = folded + c.b_off; a.res = res; a.out = o;
    a.P = B * ho * wo; a.HoWo = ho * wo; a.Wo = wo; a.Co = c.co; a.KH = c.k; a.KW = c.k; a.cblocks = c.ci / 64;
    a.stride_w = a.stride_h = c.stride; a.pad_w = a.pad_h = c.pad; a.relu = relu; a.n_mt = cdiv(a.P, BM);
    const int bn = conv_bn(c.co);
    a.n_nt = c.co / bn;
    char nm[48];
    snprintf(nm, sizeof nm, "rn.conv%dx%d_c%d", c.k, c.k, c.co);
    KernelTimer kt(nm, s);
    if (bn == 64 && a.n_nt == 1 && a.KH * a.KW * a.cblocks <= WRES_MAX_KB) return launch_conv<64, 64, true>(tm, a, s);
    if (bn == 64) return launch_conv<64, 64>(tm, a, s);
    if (bn == 256) return launch_conv<256, 64>(tm, a, s);
    return launch_conv<128, 64>(tm, a, s);
  };
  __nv_bfloat16 *cur = w.buf[0], *nxt = w.buf[1], *t1 = w.buf[2], *t2 = w.buf[3], *idt = w.buf[4];
  for (const RBlock& b : n.blocks) {
    const RConv& c1 = n.convs[b.c1];
    const RConv& c2 = n.convs[b.c2];
    const int h1 = out_dim(h, c1.k, c1.stride, c1.pad), w1 = out_dim(wd, c1.k, c1.stride, c1.pad);
    const int h2 = out_dim(h1, c2.k, c2.stride, c2.pad), w2 = out_dim(w1, c2.k, c2.stride, c2.pad);
    const __nv_bfloat16* res = cur;
    if (b.ds >= 0) { I2L_TRY(run(b.ds, cur, idt, h, wd, nullptr, 0)); res = idt; }
    I2L_TRY(run(b.c1, cur, t1, h, wd, nullptr, 1));
    if (b.c3 < 0) {
      I2L_TRY(run(b.c2, t1, nxt, h1, w1, res, 1));
    } else {
      I2L_TRY(run(b.c2, t1, t2, h1, w1, nullptr, 1));
      I2L_TRY(run(b.c3, t2, nxt, h2, w2, res, 1));
    }
    std::swap(cur, nxt);
    h = h2; wd = w2;
  }
  {
    KernelTimer kt("rn.avgpool", s);
    avgpool_nhwc_kernel<<<B, 256, 0, s>>>(cur, w.pooled, h * wd, n.feat);
    I2L_LAUNCH_OK();
  }
  // embedding Linear + ReLU (encoder.py:245-247) on the tensor cores
  (void)fc_w;
  GemmBf16 g;
  g.M = B; g.N = d.embedding_dim; g.C = out; g.ldc = d.embedding_dim;
  I2L_TRY(gemm_bf16_a_map(&g.tmA1, w.pooled, B, n.feat, n.feat));
  I2L_TRY(gemm_bf16_w_map(&g.tmW1, sec + L.fc, d.embedding_dim, n.feat, n.feat));
  g.K1 = n.feat;
  g.bias = fc_b; g.relu = 1;
  KernelTimer kt("rn.fc_bf16", s);
  return gemm_bf16(g, s);
}

}  // namespace i2l
