// bf16 tensor-core GEMM for the per-step products of the GENERAL decode path (any V / E / H / L):
//   C[M,N] (fp32) = A1[M,K1] W1[N,K1]^T (+ A2[M,K2] W2[N,K2]^T) + bias[N] + add_rows[M,N] + add_table[tab_idx[m], N]
// A and W are bf16, K-major; TMA (SWIZZLE_128B) -> shared memory -> tcgen05.mma, fp32 accumulator in TMEM.
// One 128 x 128 output tile per CTA, 4-stage pipeline: warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-5 =
// epilogue.  K and N tails are zero-filled by the tensor maps; M / N tails are masked in the epilogue.
// Same epilogue terms as gemm_f32.cu (the gate GEMM of LSTMDecoder.decode_step, model/decoder.py:268-279, with
// W_ih [emb ; ctx] hoisted into the token table / per-row constant).
#include "gemm_bf16.cuh"

namespace i2l {
namespace {

using namespace tc;

__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

constexpr int GB_BM = 128, GB_BN = 128, GB_BK = 64, GB_STAGES = 4;
constexpr int GB_A_BYTES = GB_BM * GB_BK * 2, GB_B_BYTES = GB_BN * GB_BK * 2;
constexpr int GB_STAGE = GB_A_BYTES + GB_B_BYTES;
constexpr int GB_OFF_BAR = GB_STAGES * GB_STAGE;
constexpr int GB_SMEM = GB_OFF_BAR + 128;

__global__ void __launch_bounds__(192, 1) gemm_bf16_kernel(const __grid_constant__ GemmBf16 g) {
  if (g.skip_flag != nullptr && *g.skip_flag != 0) return;          // device-side loop exit
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar = sbase + GB_OFF_BAR;
  auto FULL = [&](int s) { return bar + 8u * s; };
  auto EMPTY = [&](int s) { return bar + 8u * (GB_STAGES + s); };
  const uint32_t TFULL = bar + 8u * (2 * GB_STAGES);
  uint32_t* misc = reinterpret_cast<uint32_t*>(smem + GB_OFF_BAR + 8 * (2 * GB_STAGES + 1));
  if (tid == 0) {
    for (int s = 0; s < GB_STAGES; ++s) { mbar_init(FULL(s), 1); mbar_init(EMPTY(s), 1); }
    mbar_init(TFULL, 1);
    mbar_fence_init();
    tma_prefetch_desc(&g.tmA1); tma_prefetch_desc(&g.tmW1);
  }
  if (warp == 1) tmem_alloc<GB_BN>(smem_u32(misc));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = misc[0];
  const int m0 = blockIdx.x * GB_BM, n0 = blockIdx.y * GB_BN;
  const int kb1 = (g.K1 + GB_BK - 1) / GB_BK, kb2 = (g.K2 + GB_BK - 1) / GB_BK;
  const int nkb = kb1 + kb2;
  if (warp == 0) {
    if (elect_one()) {
      int stage = 0; uint32_t ph = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(EMPTY(stage), ph ^ 1);
        mbar_arrive_expect_tx(FULL(stage), GB_STAGE);
        const uint32_t a = sbase + stage * GB_STAGE;
        const bool second = kb >= kb1;
        const int k0 = (second ? kb - kb1 : kb) * GB_BK;
        tma_load_2d(a, second ? &g.tmA2 : &g.tmA1, k0, m0, FULL(stage));
        tma_load_2d(a + GB_A_BYTES, second ? &g.tmW2 : &g.tmW1, k0, n0, FULL(stage));
        if (++stage == GB_STAGES) { stage = 0; ph ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    constexpr uint32_t IDESC = idesc_bf16(GB_BM, GB_BN);
    const uint64_t d0 = desc_base(sbase, 128);
    int stage = 0; uint32_t ph = 0;
    for (int kb = 0; kb < nkb; ++kb) {
      mbar_wait(FULL(stage), ph);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < GB_BK / 16; ++ks) {
          const uint64_t ad = d0 + (uint64_t)((stage * GB_STAGE + ks * 32) >> 4);
          const uint64_t bd = d0 + (uint64_t)((stage * GB_STAGE + GB_A_BYTES + ks * 32) >> 4);
          tc_mma_ss(tmem, ad, bd, IDESC, (kb | ks) ? 1u : 0u);
        }
        tc_commit(EMPTY(stage));
        if (kb == nkb - 1) tc_commit(TFULL);
      }
      __syncwarp();
      if (++stage == GB_STAGES) { stage = 0; ph ^= 1; }
    }
  } else {
    const int q = warp & 3;
    const int m = m0 + 32 * q + lane;
    const bool mv = m < g.M;
    const float* arow = (g.add_rows != nullptr && mv) ? g.add_rows + (size_t)m * g.ld_add : nullptr;
    const float* trow = (g.add_table != nullptr && mv) ? g.add_table + (size_t)g.tab_idx[m] * g.ld_tab : nullptr;
    mbar_wait(TFULL, 0);
    tc_fence_after();
    if (g.cell_c != nullptr) {
      // ---- fused LSTM cell: tile columns [32 gate + ul] = gate (i,f,g,o) of hidden unit u0 + ul; 16 units per pass
      const int Hh = g.cell_H, u0 = blockIdx.y * 32;
      const uint32_t ta = tmem + ((uint32_t)(32 * q) << 16);
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        uint32_t r[4][16];
#pragma unroll
        for (int gt = 0; gt < 4; ++gt) tc_ld16_nowait(ta + gt * 32 + half * 16, r[gt]);
        tc_wait_ld();
        if (mv) {
        const int ub = u0 + half * 16;                      // first of the 16 units of this pass
        float x[4][16];
#pragma unroll
        for (int gt = 0; gt < 4; ++gt) {
          const int col = gt * Hh + ub;                     // PyTorch gate-major column of the epilogue terms
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
            if (g.bias != nullptr) { const float4 t = __ldg(reinterpret_cast<const float4*>(g.bias + col + 4 * i)); a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w; }
            if (arow != nullptr) { const float4 t = *reinterpret_cast<const float4*>(arow + col + 4 * i); a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w; }
            if (trow != nullptr) { const float4 t = __ldg(reinterpret_cast<const float4*>(trow + col + 4 * i)); a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w; }
            x[gt][4 * i] = __uint_as_float(r[gt][4 * i]) + a.x;
            x[gt][4 * i + 1] = __uint_as_float(r[gt][4 * i + 1]) + a.y;
            x[gt][4 * i + 2] = __uint_as_float(r[gt][4 * i + 2]) + a.z;
            x[gt][4 * i + 3] = __uint_as_float(r[gt][4 * i + 3]) + a.w;
          }
        }
        float* cp = g.cell_c + (size_t)m * Hh + ub;
        float* hp = g.cell_h + (size_t)m * Hh + ub;
        float hn[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float4 cv = *reinterpret_cast<const float4*>(cp + 4 * i);
          float cc[4] = {cv.x, cv.y, cv.z, cv.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int k = 4 * i + e;
            // MUFU.TANH activations (sigmoid(x) = 0.5 tanh(x / 2) + 0.5), as in the persistent kernels: the precise
            // expf / tanhf / division sequences cost ~140 instructions per unit on the 4 epilogue warps of a CTA
            const float ig = fmaf(tanh_fast(0.5f * x[0][k]), 0.5f, 0.5f), fg = fmaf(tanh_fast(0.5f * x[1][k]), 0.5f, 0.5f);
            const float gg = tanh_fast(x[2][k]), og = fmaf(tanh_fast(0.5f * x[3][k]), 0.5f, 0.5f);
            const float cn = fmaf(fg, cc[e], ig * gg);
            cc[e] = cn;
            hn[k] = og * tanh_fast(cn);
          }
          *reinterpret_cast<float4*>(cp + 4 * i) = make_float4(cc[0], cc[1], cc[2], cc[3]);
          *reinterpret_cast<float4*>(hp + 4 * i) = make_float4(hn[4 * i], hn[4 * i + 1], hn[4 * i + 2], hn[4 * i + 3]);
        }
        uint32_t pk[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          __nv_bfloat162 h2 = __floats2bfloat162_rn(hn[2 * i], hn[2 * i + 1]);
          pk[i] = *reinterpret_cast<uint32_t*>(&h2);
        }
        uint4* hb = reinterpret_cast<uint4*>(g.cell_hb + (size_t)m * Hh + ub);
        hb[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        hb[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
        __syncwarp();
      }
    } else {
    // 16 consecutive columns per tcgen05.ld: 64-byte row segments of C / add_rows / the token's table row, moved as
    // float4 when the segment is complete and 16-byte aligned (always, for N % 16 == 0 and 4-float-aligned rows)
    const bool vec_ok = (g.ldc % 4) == 0 && (g.ld_add % 4) == 0 && (g.ld_tab % 4) == 0 && (reinterpret_cast<uintptr_t>(g.C) % 16) == 0 &&
                        (reinterpret_cast<uintptr_t>(g.add_rows) % 16) == 0 && (reinterpret_cast<uintptr_t>(g.add_table) % 16) == 0 &&
                        (reinterpret_cast<uintptr_t>(g.bias) % 16) == 0;
#pragma unroll 2
    for (int cc = 0; cc < GB_BN / 16; ++cc) {
      uint32_t r[16];
      tc_ld16_nowait(tmem + ((uint32_t)(32 * q) << 16) + cc * 16, r);
      const int nb = n0 + cc * 16;
      const bool full = vec_ok && mv && nb + 16 <= g.N;
      float4 ar[4] = {}, tr[4] = {}, bs[4] = {};
      if (full) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (arow != nullptr) ar[i] = *reinterpret_cast<const float4*>(arow + nb + 4 * i);
          if (trow != nullptr) tr[i] = __ldg(reinterpret_cast<const float4*>(trow + nb + 4 * i));
          if (g.bias != nullptr) bs[i] = __ldg(reinterpret_cast<const float4*>(g.bias + nb + 4 * i));
        }
      }
      tc_wait_ld();
      if (!mv || nb >= g.N) continue;
      float* dst = g.C + (size_t)m * g.ldc + nb;
      if (full) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float4 o;
          o.x = __uint_as_float(r[4 * i]) + bs[i].x + ar[i].x + tr[i].x;
          o.y = __uint_as_float(r[4 * i + 1]) + bs[i].y + ar[i].y + tr[i].y;
          o.z = __uint_as_float(r[4 * i + 2]) + bs[i].z + ar[i].z + tr[i].z;
          o.w = __uint_as_float(r[4 * i + 3]) + bs[i].w + ar[i].w + tr[i].w;
          if (g.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
          reinterpret_cast<float4*>(dst)[i] = o;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          if (nb + i < g.N) {
            float v = __uint_as_float(r[i]);
            if (g.bias != nullptr) v += __ldg(g.bias + nb + i);
            if (arow != nullptr) v += arow[nb + i];
            if (trow != nullptr) v += __ldg(trow + nb + i);
            dst[i] = g.relu ? fmaxf(v, 0.f) : v;
          }
        }
      }
    }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<GB_BN>(tmem);
}

}  // namespace

int gemm_bf16_operand_map(CUtensorMap* out, const void* base, int rows, int K, int ld_elems, int box_rows) {
  I2L_REQUIRE((ld_elems % 8) == 0 && (reinterpret_cast<uintptr_t>(base) % 16) == 0,
              "gemm_bf16: operand rows must be 16-byte aligned (leading dimension %d)", ld_elems);
  uint64_t dims[2] = {(uint64_t)K, (uint64_t)rows};
  uint64_t str[1] = {(uint64_t)ld_elems * 2};
  uint32_t box[2] = {GB_BK, (uint32_t)box_rows};
  return make_tensor_map(out, base, 2, dims, str, box, 128, 2);
}
int gemm_bf16_a_map(CUtensorMap* out, const void* a, int M, int K, int lda) { return gemm_bf16_operand_map(out, a, M, K, lda, GB_BM); }
int gemm_bf16_w_map(CUtensorMap* out, const void* w, int N, int K, int ldw) { return gemm_bf16_operand_map(out, w, N, K, ldw, GB_BN); }

int gemm_bf16(const GemmBf16& g, cudaStream_t s) {
  if (g.M <= 0 || g.N <= 0) return I2L_OK;
  I2L_REQUIRE(g.K1 > 0 && (g.C != nullptr || g.cell_c != nullptr), "gemm_bf16: invalid arguments");
  if (g.cell_c != nullptr)
    I2L_REQUIRE(g.cell_h && g.cell_hb && g.cell_H > 0 && g.N == 4 * g.cell_H && (g.cell_H % 32) == 0 &&
                (g.ld_add % 4) == 0 && (g.ld_tab % 4) == 0, "gemm_bf16: invalid fused-cell arguments");
  static thread_local int attr_dev = -1;
  int dev = 0;
  I2L_CUDA_OK(cudaGetDevice(&dev));
  if (attr_dev != dev) {
    I2L_CUDA_OK(cudaFuncSetAttribute(gemm_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GB_SMEM));
    attr_dev = dev;
  }
  gemm_bf16_kernel<<<dim3(cdiv(g.M, GB_BM), cdiv(g.N, GB_BN)), 192, GB_SMEM, s>>>(g);
  I2L_LAUNCH_OK();
  return I2L_OK;
}

}  // namespace i2l
