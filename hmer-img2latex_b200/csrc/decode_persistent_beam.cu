// Persistent beam search for the headline shape (E = H = 256, one LSTM layer, V <= 512, bf16):
// Seq2SeqModel._beam_search (model/seq2seq.py:234-298) run independently per image, all
// max_length steps inside ONE kernel, on the same resident-weight cluster design as the greedy
// kernel (decode_persistent.cu): a cluster of 4 CTAs owns BNB = 48 beam rows = 3 column groups of
// floor(16/K) images x K beams (12 epilogue warps = 4 TMEM lane quadrants x 3 column groups); W_hh / W_out
// slices live in tensor memory as the MMA A operand.  With K = 5 that is 9 images per cluster: the 512 images
// of BASELINE configs[2] are 57 clusters = 228 CTAs = 2 waves on 148 SMs (the 32-row version needed 86
// clusters = 3 waves).  Tensor memory: 3 x 128 weight columns + 2 x 48 accumulator columns -- the logits
// accumulator ALIASES the second gate accumulator: the gates of step s are consumed before MMA-L(s) is issued,
// and MMA-G1(s+1) is issued only after every warp has copied the logits of step s out of tensor memory.
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

#ifndef I2L_TANH_F16X2
#define I2L_TANH_F16X2 0     // see decode_persistent.cu: two MUFU.TANH.F16 per f16x2 on sm_100a, no gain
#endif

namespace i2l {
namespace {

constexpr int NG = 3;                          // 16-column groups per cluster
constexpr int BNB = 16 * NG;                   // beam rows per cluster = N of every MMA
constexpr int BHSLICE = BNB * 128;             // one K-block of the h operand (one CTA's 64 units), bytes
constexpr int BHB = 4 * BHSLICE;               // one h buffer
constexpr int MW = (BNB + 31) / 32;            // merge warps: lane <-> beam row 32 w + lane
// tensor-memory columns: fp32 accumulators (the logits accumulator aliases the second gate accumulator, see above)
constexpr int BTC_G0 = 0, BTC_G1 = BNB, BTC_L = BNB;
static_assert(2 * BNB <= 128, "accumulators must stay below the first weight tile (column 128)");
constexpr uint32_t BIDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BNB >> 3) << 17) | ((128u >> 4) << 24);

template <int K>
struct Geo {
  static constexpr int IPG = 16 / K;                       // images per 16-column group
  static constexpr int IPC = NG * IPG;                     // images per cluster
  static constexpr int USED = IPG * K;                     // live columns per group
  static constexpr int XW = (2 * K + 2 + 3) / 4 * 4;       // 32-bit words per exchange record (16-byte multiple)
  // shared memory map (bytes)
  static constexpr int OFF_H = 0;                          // 2 h buffers (B operand)
  static constexpr int OFF_LT = OFF_H + 2 * BHB;           // logits tile, transposed: [BNB rows][128 vocab] fp32
  static constexpr int OFF_XCHG = OFF_LT + BNB * 128 * 4;  // [4 ctas][BNB rows] records
  static constexpr int OFF_CAND = OFF_XCHG + CL * BNB * XW * 4;   // [BNB rows][K] (double score, int tok, pad)
  static constexpr int OFF_ROW = OFF_CAND + BNB * K * 16;  // [BNB] (double score, int live, pad)
  static constexpr int OFF_NEW = OFF_ROW + BNB * 16;       // [BNB] (double score, int parent slot, int tok)
  static constexpr int OFF_PUB = OFF_NEW + BNB * 16;       // [BNB] (int parent column in group, int tok)
  static constexpr int OFF_BAR = OFF_PUB + BNB * 8;
  static constexpr int OFF_MISC = OFF_BAR + 16 * 8;
  static constexpr int SMEM = OFF_MISC + 32;               // tmem base, -, -, -, per-merge-warp "all images done"
  static_assert(K >= 1 && K <= 16 && SMEM <= 232448, "beam geometry");
};

// 4 x NG warps, no dedicated MMA warp: warp 0 issues the MMAs at the point where it would otherwise just wait for
// them.  12 warps = 3 per SM sub-partition = at most 168 registers per thread.
constexpr int BT = 4 * NG * 32;
__device__ __forceinline__ void beam_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(BT) : "memory"); }

// BAR_HS + 4 * buf + d: K-block of h buffer `buf` written by the CTA at cluster distance d (rank - d; d = 0: this CTA)
enum { BAR_LDONE = 3, BAR_GDONE = 4, BAR_TOK = 5, BAR_FINAL = 6, BAR_HS = 8 };

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

// float -> int with the same ordering under signed compare (-0.0 < +0.0, NaNs at the extremes)
__device__ __forceinline__ int ordered_int(float f) {
  const int b = __float_as_int(f);
  return b ^ ((b >> 31) & 0x7fffffff);
}
// fp64 score <-> int64 key with the same ordering (signed compare); the map is an involution
__device__ __forceinline__ long long score_key(double d) {
  const long long b = __double_as_longlong(d);
  return b ^ ((b >> 63) & 0x7fffffffffffffffLL);
}
__device__ __forceinline__ double key_score(long long k) { return __longlong_as_double(k ^ ((k >> 63) & 0x7fffffffffffffffLL)); }

// exact float -> double widening with integer ops only (the FP64 pipe of sm_100 has a ~100-cycle
// latency: F2F.F64 / DSETP stay off the critical path, DADD is the one FP64 instruction left).
// Denormal inputs flush to zero (a log-prob is 0, < -1e-7 or -inf, never denormal).
__device__ __forceinline__ double widen_f32(float f) {
  const uint32_t b = __float_as_uint(f);
  const uint32_t e = (b >> 23) & 0xffu, m = b & 0x7fffffu;
  const uint32_t ex = e == 0xffu ? 0x7ffu : e + 896u;      // inf / nan keep the all-ones exponent
  const uint32_t hi = (b & 0x80000000u) | (e == 0u ? 0u : ((ex << 20) | (m >> 3)));
  const uint32_t lo = e == 0u ? 0u : (m << 29);
  return __hiloint2double((int)hi, (int)lo);
}

struct BeamParams {
  const unsigned char* wimg;
  const float* gtok;
  const float* bias;
  const float* gctx;             // [B][1024] fp32 per IMAGE (PyTorch gate order)
  int* tr_parent; int* tr_token; double* tr_score;   // (T,B,K); pre-filled with -1 / NaN; tr_score may be null
  BeamState* bstate;             // [B]
  double* score;                 // [B*K] final beam scores
  int B, T, start_id, end_id;
  int* cand_tok; float* cand_logp;   // optional (T,B,K,K) audit trail of every live beam's top-K (token, log-prob)
  long long* dbg_ts; int dbg_step;   // diagnostics build: clock64() stamps of one step of cluster 0 / rank 0 / thread 0
};

#ifdef I2L_DIAG
#define BEAM_TS(slot)                                              \
  do {                                                             \
    asm volatile("" ::: "memory");                                 \
    if (ts_on && tid == 0) P.dbg_ts[slot] = clock64();             \
    asm volatile("" ::: "memory");                                 \
  } while (0)
#else
#define BEAM_TS(slot) do { } while (0)
#endif

template <int K>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(BT, 1) persistent_beam_kernel(BeamParams P) {
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
    for (int i = 0; i < 8; ++i) mbar_init(BAR(BAR_HS + i), 1);
    mbar_init(BAR(BAR_LDONE), 1);
    mbar_init(BAR(BAR_GDONE), 1);
    mbar_init(BAR(BAR_TOK), 1);
    mbar_init(BAR(BAR_FINAL), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    misc[1] = 0;
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(sbase + G::OFF_MISC) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = tid; i < 2 * BHB / 16; i += BT) reinterpret_cast<uint4*>(smem + G::OFF_H)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = misc[0];
  if (warp < 8) {   // resident weights -> tensor memory (same image as the greedy kernel), warps 0-7
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

  const uint32_t TM_L = tmem + BTC_L, TM_G0 = tmem + BTC_G0, TM_G1 = tmem + BTC_G1;

  // MMA issue (warp 0, one elected lane): D[gate/vocab row, beam row] = W (TMEM A operand) x h^T (smem B operand)
  const uint64_t dbase = DESC_HI | (uint64_t)(((sbase >> 4) & 0x3FFFu) | (1u << 16));
  auto issue_tile = [&](uint32_t d_tmem, uint32_t a_col, uint32_t h_off) {
#pragma unroll
    for (int kb = 0; kb < 4; ++kb) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        uint64_t bd = dbase + (uint64_t)((h_off + kb * BHSLICE + k * 32) >> 4);
        tc_mma_ts(d_tmem, tmem + a_col + (kb * 4 + k) * 8, bd, BIDESC, (kb | k) ? 1u : 0u);
      }
    }
  };
  if (warp == 0) {   // gates of step 0 from h_0 = 0 (buffer 0)
    tc_fence_after();
    if (elect_one()) {
      issue_tile(TM_G0, TC_WG0, G::OFF_H);
      issue_tile(TM_G1, TC_WG1, G::OFF_H);
      tc_commit(BAR(BAR_GDONE));
    }
    __syncwarp();
  }
  {
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
    const float2* gt_base = reinterpret_cast<const float2*>(P.gtok + (size_t)rank * 256) + p;   // [V][rank][128 rows][tile 0,1]
    float* LT = reinterpret_cast<float*>(smem + G::OFF_LT);
    // word of vocabulary row p inside a row of the transposed logits tile: 16-byte chunk c = p / 4 is stored at
    // c ^ ((c >> 3) & 3), so that the top-K threads (8 per row, 16 CONTIGUOUS vocabulary entries each = chunks
    // 4*part .. 4*part+3) read conflict-free float4s
    const int pw = 32 * q + ((((lane >> 2) ^ q) << 2) | (lane & 3));

    // ---- row / image state of the merge warps (warp w < MW: lane <-> beam row n = 32 w + lane of the cluster) ----
    const bool mw = warp < MW;
    const int n = 32 * warp + lane;                       // meaningful in the merge warps only
    const int n_w = n & 15, n_img_l = n_w / K, n_slot = n_w - n_img_l * K;
    const int n_img = cluster * G::IPC + (n >> 4) * G::IPG + n_img_l;
    const bool n_valid = mw && n < BNB && n_w < G::USED && n_img < P.B;
    const int n_leader = lane - n_slot;                   // lane of beam slot 0 of the image (same warp: groups are 16-aligned)
    const bool is_leader = n_valid && n_slot == 0;
    double base = 0.0;                                    // score of the beam in this row
    long long basekey = 0;                                // score_key(base)
    int curtok = P.start_id;
    // image state (leader lanes)
    int st_alive = n_valid ? 1 : 0, st_nbeams = 1, st_has = 0, st_best_step = -1, st_best_slot = -1, st_last = -1;
    long long st_best = 0;                                // best completed score as an order-preserving key

    if (mw && n < BNB) {   // row record of step 0: only slot 0 (the START beam) is live
      reinterpret_cast<double*>(smem + G::OFF_ROW)[n * 2] = 0.0;
      reinterpret_cast<int*>(smem + G::OFF_ROW)[n * 4 + 2] = (n_valid && n_slot == 0) ? 1 : 0;
    }
    int s = 0;
    for (; s < P.T; ++s) {
      // ---------------- Epi-G(s): gates (read from the parent's column) -> c_{s+1}, h_{s+1} ----------------
#ifdef I2L_DIAG
      const bool ts_on = P.dbg_ts != nullptr && cluster == 0 && rank == 0 && s == P.dbg_step;
#endif
      BEAM_TS(0);
      float gt0[16], gt1[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float2 g = __ldg(gt_base + (size_t)tok[j] * 512);
        gt0[j] = g.x;
        gt1[j] = g.y;
      }
      BEAM_TS(1);
      mbar_wait(BAR(BAR_GDONE), s & 1);
      tc_fence_after();
      BEAM_TS(2);
      uint32_t r0[16], r1[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        r0[j] = tc_ld1_nowait(TM_G0 + lane_addr + col0 + par[j]);
        r1[j] = tc_ld1_nowait(TM_G1 + lane_addr + col0 + par[j]);
      }
      tc_wait_ld();
      BEAM_TS(3);
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
      unsigned char* hdst = smem + G::OFF_H + nb * BHB + rank * BHSLICE;
      float y0[16], y1[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float x0 = __uint_as_float(r0[j]) + gt0[j] + gctx0[j];
        float x1 = __uint_as_float(r1[j]) + gt1[j] + gctx1[j];
#if I2L_TANH_F16X2
        float t0, t1;
        tanh2_f16(0.5f * x0, s1 * x1, t0, t1);
        y0[j] = fmaf(t0, 0.5f, 0.5f);
        y1[j] = fmaf(t1, m1, b1);
#else
        y0[j] = fmaf(tanh_approx(0.5f * x0), 0.5f, 0.5f);
        y1[j] = fmaf(tanh_approx(s1 * x1), m1, b1);
#endif
      }
      float pg[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) pg[j] = __shfl_xor_sync(0xffffffffu, y0[j] * y1[j], 16);
      float hn[16];
#if I2L_TANH_F16X2
#pragma unroll
      for (int j = 0; j < 16; j += 2) {
        const float ca = fmaf(y0[j], c[j], pg[j]), cb = fmaf(y0[j + 1], c[j + 1], pg[j + 1]);
        c[j] = ca; c[j + 1] = cb;
        float ta, tb;
        tanh2_f16(ca, cb, ta, tb);
        hn[j] = y1[j] * ta; hn[j + 1] = y1[j + 1] * tb;
      }
#else
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float cnv = fmaf(y0[j], c[j], pg[j]);
        c[j] = cnv;
        hn[j] = y1[j] * tanh_approx(cnv);
      }
#endif
      if (hi) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int nn = col0 + j;
          const int chunk = (u >> 3) ^ (nn & 7);
          *reinterpret_cast<__nv_bfloat16*>(hdst + nn * 128 + chunk * 16 + (u & 7) * 2) = __float2bfloat16(hn[j]);
        }
      }
      BEAM_TS(4);
      fence_proxy_async();
      tc_fence_before();
      beam_bar_sync();
      if (tid == 0) {
        const uint32_t src = sbase + G::OFF_H + nb * BHB + rank * BHSLICE;
        mbar_arrive(BAR(BAR_HS + 4 * nb));                          // own K-block is in place
#pragma unroll
        for (uint32_t d = 1; d < CL; ++d) {
          // arm the barrier of the block that arrives from distance d, send ours to the CTA at distance d
          mbar_arrive_expect_tx(BAR(BAR_HS + 4 * nb + d), BHSLICE);
          const uint32_t peer = (rank + d) & (CL - 1);
          bulk_s2peer(mapa(src, peer), src, BHSLICE, mapa(BAR(BAR_HS + 4 * nb + d), peer));
        }
      }
      if (warp == 0) {
        // MMA-L(s): the 4 MMAs of a K-block as soon as THAT block of h arrived (own block first, then the peers' in
        // the order their copies were sent; a 128x32x16 MMA costs ~45 cycles of the tensor pipe), then MMA-G(s+1)
        const uint32_t hb = G::OFF_H + nb * BHB;
#pragma unroll
        for (uint32_t d = 0; d < CL; ++d) {
          mbar_wait(BAR(BAR_HS + 4 * nb + d), (uint32_t)((s >> 1) & 1));
          tc_fence_after();
          const uint32_t kb = (rank - d) & (CL - 1);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t bd = dbase + (uint64_t)((hb + kb * BHSLICE + k * 32) >> 4);
              tc_mma_ts(TM_L, tmem + TC_WO + (kb * 4 + k) * 8, bd, BIDESC, (d | (uint32_t)k) ? 1u : 0u);
            }
            if (d == CL - 1) tc_commit(BAR(BAR_LDONE));
          }
          __syncwarp();
        }
        if (s + 1 < P.T && elect_one()) issue_tile(TM_G0, TC_WG0, hb);   // first gate tile of step s+1 (its accumulator is free)
        __syncwarp();
      }
      // ---------------- Epi-L(s): logits -> per-row top-K + log-sum-exp partials ----------------
      BEAM_TS(5);
      mbar_wait(BAR(BAR_LDONE), s & 1);
      tc_fence_after();
      BEAM_TS(6);
      {
        float lg[16];
        tc_ld16(TM_L + lane_addr + col0, lg);
        tc_fence_before();
#pragma unroll
        for (int j = 0; j < 16; ++j) LT[(col0 + j) * 128 + pw] = lg[j] + bias;  // lanes -> distinct banks
      }
      BEAM_TS(7);
      beam_bar_sync();
      if (warp == 0) {
        // every warp has copied the logits out of tensor memory (tcgen05.ld + fence::before_thread_sync above): the
        // accumulator the logits shared with the second gate tile is free -> MMA-G1(s+1), hidden under the top-K work
        tc_fence_after();
        if (s + 1 < P.T && elect_one()) {
          issue_tile(TM_G1, TC_WG1, G::OFF_H + ((s + 1) & 1) * BHB);
          tc_commit(BAR(BAR_GDONE));
        }
        __syncwarp();
      }
      BEAM_TS(8);
      {
        const int row = tid >> 3, part = tid & 7;           // 8 threads per beam row, vocabulary entries [16 part, 16 part + 16)
        float v[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float4 t4 = reinterpret_cast<const float4*>(LT)[row * 32 + 4 * part + (i ^ (part >> 1))];
          v[4 * i] = t4.x; v[4 * i + 1] = t4.y; v[4 * i + 2] = t4.z; v[4 * i + 3] = t4.w;
        }
        float lv[K]; int li[K];                             // sorted top-K of the 16 entries: (value, entry 0..15)
#pragma unroll
        for (int k = 0; k < K; ++k) { lv[k] = -INFINITY; li[k] = 0; }
#pragma unroll
        for (int e = 0; e < 16; ++e) {                      // ascending index: strict > keeps the lower index first
          float x = v[e]; int ix = e;
#pragma unroll
          for (int k = 0; k < K; ++k) {
            const bool gt = x > lv[k];
            const float tv = gt ? lv[k] : x; const int ti = gt ? li[k] : ix;
            lv[k] = gt ? x : lv[k]; li[k] = gt ? ix : li[k];
            x = tv; ix = ti;
          }
        }
        BEAM_TS(9);
        // sum of exp against the thread's own maximum: independent of the merge below, so the MUFU work
        // overlaps the shuffle chain; rescaled to the row maximum afterwards
        const float mloc = lv[0] == -INFINITY ? 0.f : lv[0];
        float se = 0.f;
#pragma unroll
        for (int e = 0; e < 16; ++e) se += __expf(v[e] - mloc);
        // merge of the 8 sorted lists of the row: K rounds of "best head over the 8 lanes".  The vocabulary index
        // grows with the lane, so among equal heads the lowest lane wins (ballot + ffs) and only values travel
        // through the max butterfly; the winner pops its list.  Every lane ends with the row's top-K.
        {
          float wv[K]; int wi[K];
          const int gsh = lane & 24;                        // first lane of the row's 8-lane group
#pragma unroll
          for (int r = 0; r < K; ++r) {
            float m = lv[0];
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4));
            const uint32_t grp = (__ballot_sync(0xffffffffu, lv[0] == m) >> gsh) & 0xffu;
            const int wl = (__ffs(grp) - 1) & 7;
            const int we = __shfl_sync(0xffffffffu, li[0], gsh + wl);
            wv[r] = m; wi[r] = 16 * wl + we;
            const bool won = part == wl;
#pragma unroll
            for (int k = 0; k + 1 < K; ++k) { lv[k] = won ? lv[k + 1] : lv[k]; li[k] = won ? li[k + 1] : li[k]; }
            lv[K - 1] = won ? -INFINITY : lv[K - 1];
          }
#pragma unroll
          for (int k = 0; k < K; ++k) { lv[k] = wv[k]; li[k] = wi[k]; }
        }
        BEAM_TS(10);
        const float mref = lv[0] == -INFINITY ? 0.f : lv[0];
        se *= __expf(mloc - mref);
        se += __shfl_xor_sync(0xffffffffu, se, 1);
        se += __shfl_xor_sync(0xffffffffu, se, 2);
        se += __shfl_xor_sync(0xffffffffu, se, 4);
        if (part == 0) {                                    // record -> own slot of the local exchange buffer
          uint32_t w[G::XW];
#pragma unroll
          for (int i = 0; i < G::XW; ++i) w[i] = 0;
#pragma unroll
          for (int k = 0; k < K; ++k) { w[2 * k] = __float_as_uint(lv[k]); w[2 * k + 1] = (uint32_t)(li[k] + 128 * (int)rank); }
          w[2 * K] = __float_as_uint(mref); w[2 * K + 1] = __float_as_uint(se);
          uint4* dst = reinterpret_cast<uint4*>(smem + G::OFF_XCHG + (rank * BNB + row) * (G::XW * 4));
#pragma unroll
          for (int i = 0; i < G::XW / 4; ++i) dst[i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
          fence_proxy_async();
        }
      }
      if (mw && n < BNB) {                                  // slots of the new beams start empty
        double* new_sc = reinterpret_cast<double*>(smem + G::OFF_NEW);
        int* new_i = reinterpret_cast<int*>(smem + G::OFF_NEW);
        new_sc[n * 2] = 0.0; new_i[n * 4 + 2] = -1; new_i[n * 4 + 3] = -1;
      }
      BEAM_TS(11);
      beam_bar_sync();
      BEAM_TS(23);
      if (tid == 0) {                                       // the CTA's 32 records -> the three peers (one bulk copy each)
        const uint32_t src = sbase + G::OFF_XCHG + rank * BNB * (G::XW * 4);
        mbar_arrive_expect_tx(BAR(BAR_TOK), (CL - 1) * BNB * G::XW * 4);
#pragma unroll
        for (uint32_t d = 1; d < CL; ++d) {
          const uint32_t peer = (rank + d) & (CL - 1);
          bulk_s2peer(mapa(src, peer), src, BNB * G::XW * 4, mapa(BAR(BAR_TOK), peer));
        }
      }
      BEAM_TS(24);
      if (mw && s > 0) {
        // finished beams retire to `completed` when next visited (258-260): depends on the previous step
        // only, so it runs while the records are in flight.  First-wins max in slot order.
        const int alive_i = __shfl_sync(0xffffffffu, st_alive, n_leader);
        const int nbeams_i = __shfl_sync(0xffffffffu, st_nbeams, n_leader);
        const bool ret = n_valid && alive_i && n_slot < nbeams_i && curtok == P.end_id;
#pragma unroll
        for (int k = 0; k < K; ++k) {
          const int src = (n_leader + k) & 31;              // lane of slot k of the image
          const int rk = __shfl_sync(0xffffffffu, ret ? 1 : 0, src);
          const long long sk = __shfl_sync(0xffffffffu, basekey, src);
          if (rk && (!st_has || sk > st_best)) { st_has = 1; st_best = sk; st_best_step = s - 1; st_best_slot = k; }
        }
      }
      BEAM_TS(12);
      mbar_wait_cluster(BAR(BAR_TOK), s & 1);
      BEAM_TS(18);
      if (tid < 4 * BNB) {
        // ---- torch.topk(log_softmax(logits), K) per row (seq2seq.py:266-267) on the first 4 BNB threads: 4 lanes per row take
        // one CTA record each; 4-way merge = K rounds of best-head butterfly (lower rank = lower vocabulary
        // index wins ties); log-sum-exp combined over the 4 partials
        const int row = tid >> 2, part = tid & 3;
        const float* X = reinterpret_cast<const float*>(smem + G::OFF_XCHG) + (part * BNB + row) * G::XW;
        float lv[K]; int li[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
          const float2 e = reinterpret_cast<const float2*>(X)[k];
          lv[k] = e.x; li[k] = __float_as_int(e.y);
        }
        const float2 ms = reinterpret_cast<const float2*>(X)[K];
        float M = ms.x;
        M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, 1));
        M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, 2));
        float S = ms.y * __expf(ms.x - M);
        S += __shfl_xor_sync(0xffffffffu, S, 1);
        S += __shfl_xor_sync(0xffffffffu, S, 2);
        const float lse = logf(S);
        const double* row_sc = reinterpret_cast<const double*>(smem + G::OFF_ROW);
        const int* row_i = reinterpret_cast<const int*>(smem + G::OFF_ROW);
        const double rbase = row_sc[row * 2];
        const bool rlive = row_i[row * 4 + 2] != 0;
        long long* cand_key = reinterpret_cast<long long*>(smem + G::OFF_CAND);
        int* cand_tk = reinterpret_cast<int*>(smem + G::OFF_CAND);
        // rank of every own entry among the 4 x K entries of the row by counting (no dependent rounds): the
        // own list is sorted, an entry of CTA record `q` precedes an equal value iff q < part (lower index)
        // (values as order-preserving ints: one compare + one add per pair, the tie rule is folded into the threshold)
        int rk[K], vi[K];
#pragma unroll
        for (int k = 0; k < K; ++k) { rk[k] = k; vi[k] = ordered_int(lv[k] + 0.0f); }   // + 0.0f: -0.0 and +0.0 compare equal
#pragma unroll
        for (int d = 1; d < 4; ++d) {
          const int lower = ((part ^ d) < part) ? 1 : 0;    // o precedes v  <=>  o > v, or o == v and lower  <=>  o > v - lower
          int thr[K];
#pragma unroll
          for (int k = 0; k < K; ++k) thr[k] = vi[k] - lower;
#pragma unroll
          for (int k2 = 0; k2 < K; ++k2) {
            const int o = __shfl_xor_sync(0xffffffffu, vi[k2], d);
#pragma unroll
            for (int k = 0; k < K; ++k) rk[k] += o > thr[k] ? 1 : 0;
          }
        }
        BEAM_TS(25);
        // candidates of the row (268-275), written at their rank: fp64 score as an order-preserving int64
        // key; dead rows sort last
        {
          long long ck[K]; float clp[K];
#pragma unroll
          for (int k = 0; k < K; ++k) {                     // K independent DADDs (the FP64 pipe is slow: keep them in flight together)
            clp[k] = (lv[k] - M) - lse;                     // log-prob of the rk-th best token
            ck[k] = score_key(rlive ? rbase + widen_f32(clp[k]) : -INFINITY);
          }
#pragma unroll
          for (int k = 0; k < K; ++k) {
            if (rk[k] < K) {
              cand_key[(row * K + rk[k]) * 2] = ck[k];
              cand_tk[(row * K + rk[k]) * 4 + 2] = li[k];
            }
          }
          if (P.cand_tok != nullptr && rank == 0 && rlive) {
            const int w16 = row & 15, img = cluster * G::IPC + (row >> 4) * G::IPG + w16 / K;
#pragma unroll
            for (int k = 0; k < K; ++k) {
              if (rk[k] < K) {
                const size_t o = ((((size_t)s * P.B + img) * K + w16 % K) * K) + rk[k];
                P.cand_tok[o] = li[k]; P.cand_logp[o] = clp[k];
              }
            }
          }
        }
        BEAM_TS(13);
      }
      BEAM_TS(14);
      beam_bar_sync();
      BEAM_TS(19);
      {
        // stable descending order of the image's K x K candidates (sorted(..., reverse=True)[:K], 279-280) by
        // rank counting: thread (row, r) counts the candidates that precede candidate r of beam `row`
        // (higher score, or equal score and earlier in (beam, rank) order); ranks < K are the new beams
        const int row = tid >> 3, r = tid & 7;
        const int w = row & 15, slot = w % K, row0i = row - slot;
        const long long* cand_key = reinterpret_cast<const long long*>(smem + G::OFF_CAND);
        const int* cand_tk = reinterpret_cast<const int*>(smem + G::OFF_CAND);
        const int* row_i = reinterpret_cast<const int*>(smem + G::OFF_ROW);
        if (r < K && w < G::USED && row_i[row * 4 + 2]) {
          const long long key = cand_key[(row * K + r) * 2];
          int rk = r;                                        // the beam's own list is sorted: r earlier entries
#pragma unroll
          for (int k2 = 0; k2 < K; ++k2) {
            // o precedes key  <=>  o > key, or o == key and k2 < slot  <=>  o > key - (k2 < slot)
            const long long thr = key - (k2 < slot ? 1 : 0);
            const bool other = k2 != slot;
#pragma unroll
            for (int r2 = 0; r2 < K; ++r2) {
              const long long o = cand_key[((row0i + k2) * K + r2) * 2];
              rk += (other && o > thr) ? 1 : 0;
            }
          }
          if (rk < K) {
            long long* new_key = reinterpret_cast<long long*>(smem + G::OFF_NEW);
            int* new_i = reinterpret_cast<int*>(smem + G::OFF_NEW);
            new_key[(row0i + rk) * 2] = key; new_i[(row0i + rk) * 4 + 2] = slot; new_i[(row0i + rk) * 4 + 3] = cand_tk[(row * K + r) * 4 + 2];
          }
        }
      }
      BEAM_TS(20);
      beam_bar_sync();
      BEAM_TS(15);
      if (mw) {
        const int nn_ = n < BNB ? n : 0;                    // lanes beyond the last row read row 0 and write nothing
        const long long* new_key = reinterpret_cast<const long long*>(smem + G::OFF_NEW);
        const int* new_i = reinterpret_cast<const int*>(smem + G::OFF_NEW);
        const int alive_i = __shfl_sync(0xffffffffu, st_alive, n_leader);
        const int ps = new_i[nn_ * 4 + 2], tk = new_i[nn_ * 4 + 3];
        const long long nkey = new_key[nn_ * 2];
        const double nsc = key_score(nkey);
        const uint32_t img_mask = ((1u << K) - 1u) << (n_leader & 31);
        const uint32_t kept = __ballot_sync(0xffffffffu, n_valid && ps >= 0) & img_mask;
        const uint32_t notend = __ballot_sync(0xffffffffu, n_valid && ps >= 0 && tk != P.end_id) & img_mask;
        if (is_leader && st_alive) {
          const int nkeep = __popc(kept);
          if (nkeep == 0) {                                 // `if not candidates: break` (276-277)
            st_alive = 0;
          } else {
            st_nbeams = nkeep; st_last = s;
            if (notend == 0) {                              // all new beams end with END (282-284): completed.extend
              // the new beams are sorted by score, max() keeps the first maximum: only slot 0 can displace
              if (!st_has || nkey > st_best) { st_has = 1; st_best = nkey; st_best_step = s; st_best_slot = 0; }
              st_alive = 0;
            }
          }
        }
        int* pub = reinterpret_cast<int*>(smem + G::OFF_PUB);
        if (n >= BNB) {
          // no row
        } else if (n_valid && alive_i) {
          const size_t tro = ((size_t)s * P.B + n_img) * K + n_slot;
          if (rank == 0) {
            P.tr_parent[tro] = ps; P.tr_token[tro] = tk;
            if (P.tr_score) P.tr_score[tro] = ps >= 0 ? nsc : nan("");
          }
          base = ps >= 0 ? nsc : 0.0;
          basekey = score_key(base);
          curtok = ps >= 0 ? tk : P.end_id;
          pub[n * 2] = ps >= 0 ? (n_w - n_slot + ps) : n_w;
          pub[n * 2 + 1] = curtok;
        } else {
          pub[n * 2] = n_w;
          pub[n * 2 + 1] = n_valid ? curtok : P.start_id;
        }
        const int alive_now = __shfl_sync(0xffffffffu, st_alive, n_leader);
        const int nbeams_now = __shfl_sync(0xffffffffu, st_nbeams, n_leader);
        // row record of the next step: (score, live) -- a beam whose last token is END is not expanded (258-260)
        double* row_sc = reinterpret_cast<double*>(smem + G::OFF_ROW);
        int* row_i = reinterpret_cast<int*>(smem + G::OFF_ROW);
        if (n < BNB) {
          row_sc[n * 2] = base;
          row_i[n * 4 + 2] = (n_valid && alive_now && n_slot < nbeams_now && curtok != P.end_id) ? 1 : 0;
        }
        const bool warp_done = __all_sync(0xffffffffu, !n_valid || !alive_now);
        if (lane == 0) misc[4 + warp] = warp_done ? 1u : 0u;
        BEAM_TS(16);
      }
      beam_bar_sync();
      bool all_done = true;                                 // uniform: flags written before the barrier
#pragma unroll
      for (int w = 0; w < MW; ++w) all_done = all_done && reinterpret_cast<volatile uint32_t*>(misc)[4 + w] != 0;
      BEAM_TS(17);
      {
        const int* pub = reinterpret_cast<const int*>(smem + G::OFF_PUB);
#pragma unroll
        for (int j = 0; j < 16; ++j) { par[j] = pub[(col0 + j) * 2]; tok[j] = pub[(col0 + j) * 2 + 1]; }
      }
      if (all_done) { ++s; break; }
    }
    if (rank == 0 && n_valid) {
      P.score[(size_t)n_img * K + n_slot] = base;
      if (is_leader) {
        BeamState st;
        st.alive = st_alive; st.nbeams = st_nbeams; st.has_completed = st_has; st.best_step = st_best_step;
        st.best_slot = st_best_slot; st.best_score = key_score(st_best); st.last_step = st_last;
        P.bstate[n_img] = st;
      }
    }
    if (warp == 0) {   // every MMA issued above has completed before tensor memory is released
      if (elect_one()) tc_commit(BAR(BAR_FINAL));
      __syncwarp();
      mbar_wait(BAR(BAR_FINAL), 0);
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

template <int K>
int launch_beam(const BeamParams& P, int batch, cudaStream_t s) {
  using G = Geo<K>;
  const int ncl = cdiv(batch, G::IPC);
  I2L_CUDA_OK(cudaFuncSetAttribute(persistent_beam_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM));
  persistent_beam_kernel<K><<<ncl * CL, BT, G::SMEM, s>>>(P);
  I2L_LAUNCH_OK();
  return I2L_OK;
}

}  // namespace

#ifdef I2L_DIAG
static long long* g_dbg_ts = nullptr;
static int g_dbg_step = 0;
int persistent_beam_set_ts(long long* ts, int step) { g_dbg_ts = ts; g_dbg_step = step; return I2L_OK; }
#endif

bool persistent_beam_supported(const i2l_dec_desc& d, int beam_size) {
  if (!persistent_supported(d)) return false;
  return beam_size >= 1 && beam_size <= 8 && beam_size <= d.vocab_size;
}

int persistent_beam(const i2l_dec_desc& d, const void* section, const float* gctx_img, int batch, int beam_size,
                    int start_id, int end_id, int max_length, BeamState* bstate, double* score, int* tr_parent,
                    int* tr_token, double* tr_score, int* cand_tok, float* cand_logp, cudaStream_t s) {
  I2L_REQUIRE(start_id >= 0 && start_id < d.vocab_size, "decode_beam: start token out of range");
  PSection ps = psection(d.vocab_size);
  const unsigned char* sec = reinterpret_cast<const unsigned char*>(section);
  BeamParams P{};
  P.wimg = sec + ps.wimg; P.gtok = reinterpret_cast<const float*>(sec + ps.gtok);
  P.bias = reinterpret_cast<const float*>(sec + ps.bias);
  P.gctx = gctx_img; P.tr_parent = tr_parent; P.tr_token = tr_token; P.tr_score = tr_score;
  P.bstate = bstate; P.score = score;
  P.B = batch; P.T = max_length; P.start_id = start_id; P.end_id = end_id;
  P.cand_tok = cand_tok; P.cand_logp = cand_logp;
#ifdef I2L_DIAG
  P.dbg_ts = g_dbg_ts; P.dbg_step = g_dbg_step;
#endif
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
