// Persistent beam search for the headline shape (E = H = 256, one LSTM layer, V <= 512, bf16):
// Seq2SeqModel._beam_search (model/seq2seq.py:234-298) run independently per image, all
// max_length steps inside ONE kernel, on the same resident-weight cluster design as the greedy
// kernel (decode_persistent.cu): a cluster of 4 CTAs owns 32 beam rows = 2 column groups of
// floor(16/K) images x K beams; W_hh / W_out slices live in tensor memory as the MMA A operand.
//
//   per step:   Epi-G   gates of row j are read from the accumulator column of its PARENT beam
//                       (tcgen05.ld of one column at a run-time address) and c follows the parent
//                       through a K-way register select: the reference's clone of new_hidden per
//                       candidate (seq2seq.py:272) costs no data movement at all
//               MMA-L / MMA-G(s+1) as in the greedy kernel
//               Epi-L   logits -> transposed tile in shared memory -> 8 threads per row keep a
//                       sorted top-K (value, index) of 16 vocabulary entries each, merged by
//                       bitonic half-cleaners over 3 shuffle rounds; running (max, sum exp) for
//                       log_softmax; one st.async record per (CTA, row) to all 4 CTAs
//               merge   (warp 0 of every CTA, redundantly) 4-way merge of the CTA records ->
//                       torch.topk(log_softmax(logits), K) per row (seq2seq.py:266-267); per image
//                       the K x K candidates are merged in stable descending order of the fp64
//                       score (sorted(..., reverse=True), 279-280), finished beams retire to
//                       `completed` (258-260), first-wins max (286-290); parents / tokens / scores
//                       are appended to the trace that the finalize kernel walks back
// Selection ranks by logit (ties -> lower index); ranking by the rounded fp32 log-prob can differ
// only where two distinct logits round to the same log-prob (documented near tie).
#include "decode_persistent_common.cuh"

namespace i2l {
namespace {

template <int K>
struct Geo {
  static constexpr int IPG = 16 / K;                       // images per 16-column group
  static constexpr int IPC = 2 * IPG;                      // images per cluster
  static constexpr int USED = IPG * K;                     // live columns per group
  static constexpr int XW = (2 * K + 2 + 3) / 4 * 4;       // 32-bit words per exchange record (16-byte multiple)
  // shared memory map (bytes)
  static constexpr int OFF_H = 0;                          // 2 h buffers (B operand)
  static constexpr int OFF_LT = OFF_H + 2 * HB_BYTES;      // logits tile, transposed: [32 rows][128 vocab] fp32
  static constexpr int OFF_XCHG = OFF_LT + NB * 128 * 4;   // [4 ctas][32 rows] records
  static constexpr int OFF_CAND = OFF_XCHG + CL * NB * XW * 4;   // [32 rows][K] (double score, int tok, pad)
  static constexpr int OFF_ROW = OFF_CAND + NB * K * 16;   // [32] (double score, int live, pad)
  static constexpr int OFF_NEW = OFF_ROW + NB * 16;        // [32] (double score, int parent slot, int tok)
  static constexpr int OFF_PUB = OFF_NEW + NB * 16;        // [32] (int parent column in group, int tok)
  static constexpr int OFF_BAR = OFF_PUB + NB * 8;
  static constexpr int OFF_MISC = OFF_BAR + 8 * 8;
  static constexpr int SMEM = OFF_MISC + 16;
  static_assert(K >= 1 && K <= 16 && SMEM <= 232448, "beam geometry");
};

enum { BAR_HFULL0 = 1, BAR_HFULL1 = 2, BAR_LDONE = 3, BAR_GDONE = 4, BAR_TOK = 5, BAR_FINAL = 6 };

__device__ __forceinline__ void st_async_v4(uint32_t cluster_addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d,
                                            uint32_t cluster_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(cluster_addr),
               "r"(a), "r"(b), "r"(c), "r"(d), "r"(cluster_bar)
               : "memory");
}
__device__ __forceinline__ uint32_t tc_ld1_nowait(uint32_t taddr) {
  uint32_t r;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
  return r;
}
// (value, index) total order of torch.topk / argmax: larger value first, lower index among equals
__device__ __forceinline__ bool before(float va, int ia, float vb, int ib) { return va > vb || (va == vb && ia < ib); }

struct BeamParams {
  const unsigned char* wimg;
  const float* gtok;
  const float* bias;
  const float* gctx;             // [B][1024] fp32 per IMAGE (PyTorch gate order)
  int* tr_parent; int* tr_token; double* tr_score;   // (T,B,K); pre-filled with -1 / NaN; tr_score may be null
  BeamState* bstate;             // [B]
  double* score;                 // [B*K] final beam scores
  int B, T, start_id, end_id;
  int* dbg_ctok; float* dbg_clogp;   // optional (T,B,K,K) dump of every live beam's top-K (token, log-prob)
};

template <int K>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(THREADS, 1) persistent_beam_kernel(BeamParams P) {
  using G = Geo<K>;
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster = (int)cluster_id_x();
  uint32_t* misc = reinterpret_cast<uint32_t*>(smem + G::OFF_MISC);
  const uint32_t bar = sbase + G::OFF_BAR;
  auto BAR = [&](int i) { return bar + 8u * i; };

  if ((sbase & 1023u) != 0) __trap();
  if (tid == 0) {
    mbar_init(BAR(BAR_HFULL0), 1);
    mbar_init(BAR(BAR_HFULL1), 1);
    mbar_init(BAR(BAR_LDONE), 1);
    mbar_init(BAR(BAR_GDONE), 1);
    mbar_init(BAR(BAR_TOK), 1);
    mbar_init(BAR(BAR_FINAL), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    misc[1] = 0;
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(sbase + G::OFF_MISC) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = tid; i < 2 * HB_BYTES / 16; i += THREADS) reinterpret_cast<uint4*>(smem + G::OFF_H)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = misc[0];
  if (warp < 8) {   // resident weights -> tensor memory (same image as the greedy kernel)
    const int p = 32 * (warp & 3) + lane;
    const uint32_t lane_addr = (uint32_t)(32 * (warp & 3)) << 16;
    const int half = warp >> 2;
#pragma unroll 1
    for (int t = 0; t < 3; ++t) {
      const uint4* src = reinterpret_cast<const uint4*>(P.wimg + (size_t)rank * WROW_BYTES + ((size_t)t * 128 + p) * (H * 2)) + half * 16;
      const uint32_t tcol = t == 0 ? TC_WG0 : (t == 1 ? TC_WG1 : TC_WO);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t r[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 v = __ldg(src + c * 4 + i);
          r[4 * i] = v.x; r[4 * i + 1] = v.y; r[4 * i + 2] = v.z; r[4 * i + 3] = v.w;
        }
        tc_st16(tmem + lane_addr + tcol + half * 64 + c * 16, r);
      }
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    tc_fence_before();
  }
  __syncthreads();
  tc_fence_after();
  cluster_sync_all();

  const uint32_t TM_L = tmem + TC_L, TM_G0 = tmem + TC_G0, TM_G1 = tmem + TC_G1;

  if (warp == 8) {
    // =========================== MMA issuer warp (identical to the greedy kernel) ===========================
    const uint64_t dbase = DESC_HI | (uint64_t)(((sbase >> 4) & 0x3FFFu) | (1u << 16));
    auto issue_tile = [&](uint32_t d_tmem, uint32_t a_col, uint32_t h_off) {
#pragma unroll
      for (int kb = 0; kb < 4; ++kb) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          uint64_t bd = dbase + (uint64_t)((h_off + kb * HSLICE_BYTES + k * 32) >> 4);
          tc_mma_ts(d_tmem, tmem + a_col + (kb * 4 + k) * 8, bd, IDESC, (kb | k) ? 1u : 0u);
        }
      }
    };
    tc_fence_after();
    if (elect_one()) {
      issue_tile(TM_G0, TC_WG0, G::OFF_H);
      issue_tile(TM_G1, TC_WG1, G::OFF_H);
      tc_commit(BAR(BAR_GDONE));
    }
    __syncwarp();
    for (int s = 0; s < P.T; ++s) {
      const int nb = (s + 1) & 1;
      mbar_wait(BAR(BAR_HFULL0 + nb), (uint32_t)((s >> 1) & 1));
      if (*reinterpret_cast<volatile uint32_t*>(&misc[1])) break;
      tc_fence_after();
      const uint32_t hb = G::OFF_H + nb * HB_BYTES;
      if (elect_one()) {
        issue_tile(TM_L, TC_WO, hb);
        tc_commit(BAR(BAR_LDONE));
        if (s + 1 < P.T) {
          issue_tile(TM_G0, TC_WG0, hb);
          issue_tile(TM_G1, TC_WG1, hb);
          tc_commit(BAR(BAR_GDONE));
        }
      }
      __syncwarp();
    }
    if (elect_one()) tc_commit(BAR(BAR_FINAL));
    __syncwarp();
    mbar_wait(BAR(BAR_FINAL), 0);
  } else {
    // =========================== epilogue warps (256 threads) ===========================
    const int q = warp & 3, cg = warp >> 2;
    const int p = 32 * q + lane;                          // accumulator row (TMEM lane) = CTA-local gate row / vocab row
    const int u = 16 * q + (lane & 15);                   // CTA-local hidden unit
    const bool hi = lane >= 16;
    const uint32_t lane_addr = (uint32_t)(32 * q) << 16;
    const int col0 = 16 * cg;
    float gctx0[16], gctx1[16], c[16];
    int tok[16], par[16];                                 // token / parent column (within the group) per column
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int img = cluster * G::IPC + cg * G::IPG + j / K;
      c[j] = 0.f; tok[j] = P.start_id; par[j] = j;
      if (j < G::USED && img < P.B) {
        const float* g = P.gctx + (size_t)img * 1024 + 64 * rank + u;
        gctx0[j] = g[(hi ? 1 : 0) * 256];
        gctx1[j] = g[(hi ? 3 : 2) * 256];
      } else {
        gctx0[j] = 0.f; gctx1[j] = 0.f;
      }
    }
    const float bias = P.bias[128 * rank + p];
    const float s1 = hi ? 0.5f : 1.0f, m1 = hi ? 0.5f : 1.0f, b1 = hi ? 0.5f : 0.0f;
    const float* gt_base = P.gtok + (size_t)rank * 256 + p;
    float* LT = reinterpret_cast<float*>(smem + G::OFF_LT);

    // ---- row / image state of the merge warp (warp 0: lane n <-> beam row n of the cluster) ----
    const int n = lane;                                   // meaningful in warp 0 only
    const int n_w = n & 15, n_img_l = n_w / K, n_slot = n_w - n_img_l * K;
    const int n_img = cluster * G::IPC + (n >> 4) * G::IPG + n_img_l;
    const bool n_valid = n_w < G::USED && n_img < P.B;
    const int n_leader = n - n_slot;                      // lane of beam slot 0 of the image
    const bool is_leader = n_valid && n_slot == 0;
    double base = 0.0;                                    // score of the beam in this row
    int curtok = P.start_id;
    // image state (leader lanes)
    int st_alive = n_valid ? 1 : 0, st_nbeams = 1, st_has = 0, st_best_step = -1, st_best_slot = -1, st_last = -1;
    double st_best = 0.0;

    int s = 0;
    for (; s < P.T; ++s) {
      // ---------------- Epi-G(s): gates (read from the parent's column) -> c_{s+1}, h_{s+1} ----------------
      float gt0[16], gt1[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float* g = gt_base + (size_t)tok[j] * 1024;
        gt0[j] = __ldg(g);
        gt1[j] = __ldg(g + 128);
      }
      mbar_wait(BAR(BAR_GDONE), s & 1);
      tc_fence_after();
      uint32_t r0[16], r1[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        r0[j] = tc_ld1_nowait(TM_G0 + lane_addr + col0 + par[j]);
        r1[j] = tc_ld1_nowait(TM_G1 + lane_addr + col0 + par[j]);
      }
      tc_wait_ld();
      {   // c follows the parent beam: K-way select inside each image's K columns
        float cn[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if (j < G::USED) {
            const int i0 = (j / K) * K;
            float v = c[i0];
#pragma unroll
            for (int t = 1; t < K; ++t) v = (par[j] == i0 + t) ? c[i0 + t] : v;
            cn[j] = v;
          } else {
            cn[j] = c[j];
          }
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) c[j] = cn[j];
      }
      const int nb = (s + 1) & 1;
      unsigned char* hdst = smem + G::OFF_H + nb * HB_BYTES + rank * HSLICE_BYTES;
      float y0[16], y1[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float x0 = __uint_as_float(r0[j]) + gt0[j] + gctx0[j];
        float x1 = __uint_as_float(r1[j]) + gt1[j] + gctx1[j];
        y0[j] = fmaf(tanh_approx(0.5f * x0), 0.5f, 0.5f);
        y1[j] = fmaf(tanh_approx(s1 * x1), m1, b1);
      }
      float pg[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) pg[j] = __shfl_xor_sync(0xffffffffu, y0[j] * y1[j], 16);
      float hn[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float cnv = fmaf(y0[j], c[j], pg[j]);
        c[j] = cnv;
        hn[j] = y1[j] * tanh_approx(cnv);
      }
      if (hi) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int nn = col0 + j;
          const int chunk = (u >> 3) ^ (nn & 7);
          *reinterpret_cast<__nv_bfloat16*>(hdst + nn * 128 + chunk * 16 + (u & 7) * 2) = __float2bfloat16(hn[j]);
        }
      }
      fence_proxy_async();
      tc_fence_before();
      epi_bar_sync();
      if (tid == 0) {
        const uint32_t src = sbase + G::OFF_H + nb * HB_BYTES + rank * HSLICE_BYTES;
        mbar_arrive_expect_tx(BAR(BAR_HFULL0 + nb), (CL - 1) * HSLICE_BYTES);
#pragma unroll
        for (uint32_t d = 1; d < CL; ++d) {
          uint32_t peer = (rank + d) & (CL - 1);
          bulk_s2peer(mapa(src, peer), src, HSLICE_BYTES, mapa(BAR(BAR_HFULL0 + nb), peer));
        }
      }
      // ---------------- Epi-L(s): logits -> per-row top-K + log-sum-exp partials ----------------
      mbar_wait(BAR(BAR_LDONE), s & 1);
      tc_fence_after();
      {
        float lg[16];
        tc_ld16(TM_L + lane_addr + col0, lg);
        tc_fence_before();
#pragma unroll
        for (int j = 0; j < 16; ++j) LT[(col0 + j) * 128 + p] = lg[j] + bias;   // lanes -> consecutive words
      }
      epi_bar_sync();
      {
        const int row = tid >> 3, part = tid & 7;           // 8 threads per beam row, 16 vocabulary entries each
        float v[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float4 t4 = reinterpret_cast<const float4*>(LT)[row * 32 + i * 8 + part];
          v[4 * i] = t4.x; v[4 * i + 1] = t4.y; v[4 * i + 2] = t4.z; v[4 * i + 3] = t4.w;
        }
        float lv[K]; int li[K];
#pragma unroll
        for (int k = 0; k < K; ++k) { lv[k] = -INFINITY; li[k] = 0x7fffffff; }
#pragma unroll
        for (int e = 0; e < 16; ++e) {                      // ascending index: strict > keeps the lower index first
          float x = v[e]; int ix = 4 * ((e >> 2) * 8 + part) + (e & 3);
#pragma unroll
          for (int k = 0; k < K; ++k) {
            const bool gt = x > lv[k];
            const float tv = gt ? lv[k] : x; const int ti = gt ? li[k] : ix;
            lv[k] = gt ? x : lv[k]; li[k] = gt ? ix : li[k];
            x = tv; ix = ti;
          }
        }
#pragma unroll
        for (int m = 1; m < 8; m <<= 1) {                   // merge with the partner's list (both end up identical)
          // bitonic half-cleaner against the reversed partner list (staged copy: entries k and K-1-k cross)
          float av[K]; int ai[K];
#pragma unroll
          for (int k = 0; k < K; ++k) { av[k] = lv[k]; ai[k] = li[k]; }
#pragma unroll
          for (int k = 0; k < K; ++k) {
            const float ov = __shfl_xor_sync(0xffffffffu, av[K - 1 - k], m);
            const int oi = __shfl_xor_sync(0xffffffffu, ai[K - 1 - k], m);
            const bool take = before(ov, oi, av[k], ai[k]);
            lv[k] = take ? ov : av[k]; li[k] = take ? oi : ai[k];
          }
          // odd-even transposition sort of the K survivors (descending, index-ascending among equals)
#pragma unroll
          for (int r = 0; r < K; ++r) {
#pragma unroll
            for (int k = r & 1; k + 1 < K; k += 2) {
              const bool sw = before(lv[k + 1], li[k + 1], lv[k], li[k]);
              const float tv = lv[k]; const int ti = li[k];
              lv[k] = sw ? lv[k + 1] : lv[k]; li[k] = sw ? li[k + 1] : li[k];
              lv[k + 1] = sw ? tv : lv[k + 1]; li[k + 1] = sw ? ti : li[k + 1];
            }
          }
        }
        const float mref = lv[0] == -INFINITY ? 0.f : lv[0];
        float se = 0.f;
#pragma unroll
        for (int e = 0; e < 16; ++e) se += __expf(v[e] - mref);
        se += __shfl_xor_sync(0xffffffffu, se, 1);
        se += __shfl_xor_sync(0xffffffffu, se, 2);
        se += __shfl_xor_sync(0xffffffffu, se, 4);
        if (part == 0) {
          uint32_t w[G::XW];
#pragma unroll
          for (int i = 0; i < G::XW; ++i) w[i] = 0;
#pragma unroll
          for (int k = 0; k < K; ++k) { w[2 * k] = __float_as_uint(lv[k]); w[2 * k + 1] = (uint32_t)(li[k] == 0x7fffffff ? 0x7fffffff : li[k] + 128 * (int)rank); }
          w[2 * K] = __float_as_uint(mref); w[2 * K + 1] = __float_as_uint(se);
          const uint32_t slot = sbase + G::OFF_XCHG + (rank * NB + row) * (G::XW * 4);
#pragma unroll
          for (uint32_t d = 0; d < CL; ++d) {
            const uint32_t dst = mapa(slot, d), dbar = mapa(BAR(BAR_TOK), d);
#pragma unroll
            for (int i = 0; i < G::XW / 4; ++i) st_async_v4(dst + 16 * i, w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3], dbar);
          }
        }
      }
      // ---------------- merge (warp 0): topk(log_softmax) per row, beam update per image ----------------
      if (warp == 0) {
        if (lane == 0) mbar_arrive_expect_tx(BAR(BAR_TOK), CL * NB * G::XW * 4);
        mbar_wait_cluster(BAR(BAR_TOK), s & 1);
        const float* X = reinterpret_cast<const float*>(smem + G::OFF_XCHG);
        const int* Xi = reinterpret_cast<const int*>(smem + G::OFF_XCHG);
        // log-sum-exp over the 4 CTA partials (torch.log_softmax, seq2seq.py:266)
        float M = -INFINITY;
#pragma unroll
        for (int r = 0; r < CL; ++r) M = fmaxf(M, X[(r * NB + n) * G::XW + 2 * K]);
        float S = 0.f;
#pragma unroll
        for (int r = 0; r < CL; ++r) S += X[(r * NB + n) * G::XW + 2 * K + 1] * expf(X[(r * NB + n) * G::XW + 2 * K] - M);
        const float lse = logf(S);
        // 4-way merge of the sorted CTA lists (lower rank = lower vocabulary index wins ties)
        int pos0 = 0, pos1 = 0, pos2 = 0, pos3 = 0;
        float tv[K]; int ti[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
          float bv = -INFINITY; int bi = 0x7fffffff, br = 0;
#pragma unroll
          for (int r = 0; r < CL; ++r) {
            const int ps = r == 0 ? pos0 : (r == 1 ? pos1 : (r == 2 ? pos2 : pos3));
            if (ps < K) {
              const float hv = X[(r * NB + n) * G::XW + 2 * ps];
              const int hidx = Xi[(r * NB + n) * G::XW + 2 * ps + 1];
              if (r == 0 || before(hv, hidx, bv, bi)) { bv = hv; bi = hidx; br = r; }
            }
          }
          pos0 += br == 0; pos1 += br == 1; pos2 += br == 2; pos3 += br == 3;
          tv[k] = (bv - M) - lse;                          // log-prob of the k-th best token
          ti[k] = bi;
        }
        // ---- beam bookkeeping (seq2seq.py:254-284) ----
        const int alive_i = __shfl_sync(0xffffffffu, st_alive, n_leader);
        const int nbeams_i = __shfl_sync(0xffffffffu, st_nbeams, n_leader);
        const bool live = n_valid && alive_i && n_slot < nbeams_i && curtok != P.end_id;   // 258-260
        double* cand_sc = reinterpret_cast<double*>(smem + G::OFF_CAND);
        int* cand_tk = reinterpret_cast<int*>(smem + G::OFF_CAND);
        double* row_sc = reinterpret_cast<double*>(smem + G::OFF_ROW);
        int* row_i = reinterpret_cast<int*>(smem + G::OFF_ROW);
        double* new_sc = reinterpret_cast<double*>(smem + G::OFF_NEW);
        int* new_i = reinterpret_cast<int*>(smem + G::OFF_NEW);
        row_sc[n * 2] = base; row_i[n * 4 + 2] = live ? 1 : 0;
        if (P.dbg_ctok != nullptr && rank == 0 && live) {
#pragma unroll
          for (int k = 0; k < K; ++k) {
            const size_t o = ((((size_t)s * P.B + n_img) * K + n_slot) * K) + k;
            P.dbg_ctok[o] = ti[k]; P.dbg_clogp[o] = tv[k];
          }
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
          cand_sc[(n * K + k) * 2] = base + (double)tv[k];                                   // 268-275
          cand_tk[(n * K + k) * 4 + 2] = ti[k];
        }
        __syncwarp();
        if (is_leader && st_alive) {
          for (int k = 0; k < st_nbeams; ++k) {             // finished beams retire to `completed` (258-260)
            if (!row_i[(n + k) * 4 + 2]) {
              const double sc = row_sc[(n + k) * 2];
              if (!st_has || sc > st_best) { st_has = 1; st_best = sc; st_best_step = s - 1; st_best_slot = k; }
            }
          }
          // stable descending merge of the live beams' sorted candidate lists (sorted(..., reverse=True), 279-280)
          int hp[K];
#pragma unroll
          for (int k = 0; k < K; ++k) hp[k] = 0;
          int nkeep = 0;
          bool all_end = true;
#pragma unroll 1
          for (int r = 0; r < K; ++r) {
            int best = -1; double bsc = 0.0;
#pragma unroll
            for (int k = 0; k < K; ++k) {
              if (k < st_nbeams && row_i[(n + k) * 4 + 2] && hp[k] < K) {
                const double sc = cand_sc[((n + k) * K + hp[k]) * 2];
                if (best < 0 || sc > bsc) { best = k; bsc = sc; }
              }
            }
            if (best < 0) break;
            int bh = 0;
#pragma unroll
            for (int k = 0; k < K; ++k) { if (k == best) { bh = hp[k]; hp[k] += 1; } }
            const int tk = cand_tk[((n + best) * K + bh) * 4 + 2];
            new_sc[(n + r) * 2] = bsc; new_i[(n + r) * 4 + 2] = best; new_i[(n + r) * 4 + 3] = tk;
            all_end = all_end && tk == P.end_id;
            ++nkeep;
          }
          for (int k = nkeep; k < K; ++k) { new_sc[(n + k) * 2] = 0.0; new_i[(n + k) * 4 + 2] = -1; new_i[(n + k) * 4 + 3] = -1; }
          if (nkeep == 0) {                                 // `if not candidates: break` (276-277)
            st_alive = 0;
          } else {
            st_nbeams = nkeep; st_last = s;
            if (all_end) {                                  // 282-284
              for (int k = 0; k < nkeep; ++k) {
                const double sc = new_sc[(n + k) * 2];
                if (!st_has || sc > st_best) { st_has = 1; st_best = sc; st_best_step = s; st_best_slot = k; }
              }
              st_alive = 0;
            }
          }
        }
        __syncwarp();
        int* pub = reinterpret_cast<int*>(smem + G::OFF_PUB);
        if (n_valid && alive_i) {
          const int ps = new_i[n * 4 + 2], tk = new_i[n * 4 + 3];
          const size_t tro = ((size_t)s * P.B + n_img) * K + n_slot;
          if (rank == 0) {
            P.tr_parent[tro] = ps; P.tr_token[tro] = tk;
            if (P.tr_score) P.tr_score[tro] = ps >= 0 ? new_sc[n * 2] : nan("");
          }
          base = ps >= 0 ? new_sc[n * 2] : 0.0;
          curtok = ps >= 0 ? tk : P.end_id;
          pub[n * 2] = ps >= 0 ? (n_w - n_slot + ps) : n_w;
          pub[n * 2 + 1] = curtok;
        } else {
          pub[n * 2] = n_w;
          pub[n * 2 + 1] = n_valid ? curtok : P.start_id;
        }
        const int alive_now = __shfl_sync(0xffffffffu, st_alive, n_leader);
        const bool cluster_done = __all_sync(0xffffffffu, !n_valid || !alive_now);
        if (lane == 0 && cluster_done) misc[1] = 1;
      }
      epi_bar_sync();
      {
        const int* pub = reinterpret_cast<const int*>(smem + G::OFF_PUB);
#pragma unroll
        for (int j = 0; j < 16; ++j) { par[j] = pub[(col0 + j) * 2]; tok[j] = pub[(col0 + j) * 2 + 1]; }
      }
      if (*reinterpret_cast<volatile uint32_t*>(&misc[1])) { ++s; break; }
    }
    if (warp == 0 && rank == 0 && n_valid) {
      P.score[(size_t)n_img * K + n_slot] = base;
      if (is_leader) {
        BeamState st;
        st.alive = st_alive; st.nbeams = st_nbeams; st.has_completed = st_has; st.best_step = st_best_step;
        st.best_slot = st_best_slot; st.best_score = st_best; st.last_step = st_last;
        P.bstate[n_img] = st;
      }
    }
    if (*reinterpret_cast<volatile uint32_t*>(&misc[1]) && tid == 0) {
      mbar_arrive(BAR(BAR_HFULL0 + ((s + 1) & 1)));       // release the MMA warp (early exit)
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 8) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

template <int K>
int launch_beam(const BeamParams& P, int batch, cudaStream_t s) {
  using G = Geo<K>;
  const int ncl = cdiv(batch, G::IPC);
  I2L_CUDA_OK(cudaFuncSetAttribute(persistent_beam_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM));
  persistent_beam_kernel<K><<<ncl * CL, THREADS, G::SMEM, s>>>(P);
  I2L_LAUNCH_OK();
  return I2L_OK;
}

}  // namespace

static int* g_dbg_ctok = nullptr;     // tests only: see i2l_debug_set_beam_trace
static float* g_dbg_clogp = nullptr;
int persistent_beam_set_debug(int* cand_tok, float* cand_logp) {
  g_dbg_ctok = cand_tok; g_dbg_clogp = cand_logp;
  return I2L_OK;
}

bool persistent_beam_supported(const i2l_dec_desc& d, int beam_size) {
  if (!persistent_supported(d)) return false;
  return beam_size >= 1 && beam_size <= 8 && beam_size <= d.vocab_size;
}

int persistent_beam(const i2l_dec_desc& d, const void* section, const float* gctx_img, int batch, int beam_size,
                    int start_id, int end_id, int max_length, BeamState* bstate, double* score, int* tr_parent,
                    int* tr_token, double* tr_score, cudaStream_t s) {
  I2L_REQUIRE(start_id >= 0 && start_id < d.vocab_size, "decode_beam: start token out of range");
  PSection ps = psection(d.vocab_size);
  const unsigned char* sec = reinterpret_cast<const unsigned char*>(section);
  BeamParams P{};
  P.wimg = sec + ps.wimg; P.gtok = reinterpret_cast<const float*>(sec + ps.gtok);
  P.bias = reinterpret_cast<const float*>(sec + ps.bias);
  P.gctx = gctx_img; P.tr_parent = tr_parent; P.tr_token = tr_token; P.tr_score = tr_score;
  P.bstate = bstate; P.score = score;
  P.B = batch; P.T = max_length; P.start_id = start_id; P.end_id = end_id;
  P.dbg_ctok = g_dbg_ctok; P.dbg_clogp = g_dbg_clogp;
  // traces default to "empty slot" (-1 / NaN): the kernel writes only the steps an image is alive in
  const size_t n = (size_t)max_length * batch * beam_size;
  I2L_CUDA_OK(cudaMemsetAsync(tr_parent, 0xff, n * sizeof(int), s));
  I2L_CUDA_OK(cudaMemsetAsync(tr_token, 0xff, n * sizeof(int), s));
  if (tr_score) I2L_CUDA_OK(cudaMemsetAsync(tr_score, 0xff, n * sizeof(double), s));
  KernelTimer kt("dec.beam_persistent", s);
  switch (beam_size) {
    case 1: return launch_beam<1>(P, batch, s);
    case 2: return launch_beam<2>(P, batch, s);
    case 3: return launch_beam<3>(P, batch, s);
    case 4: return launch_beam<4>(P, batch, s);
    case 5: return launch_beam<5>(P, batch, s);
    case 6: return launch_beam<6>(P, batch, s);
    case 7: return launch_beam<7>(P, batch, s);
    case 8: return launch_beam<8>(P, batch, s);
  }
  set_error("persistent_beam: unsupported beam size %d", beam_size);
  return I2L_ERR_UNSUPPORTED;
}

}  // namespace i2l
