// Persistent greedy decode for the headline shape (E = H = 256, one LSTM layer, V <= 512, bf16):
// ONE kernel runs all max_length steps.  A cluster of 4 CTAs owns a slice of 32 sequences
// and keeps the whole decoder in shared memory for the lifetime of the kernel:
//
//   CTA rank r holds hidden units [64r, 64r+64): the 256 gate rows (i,f,g,o) of W_hh for
//   those units (two M=128 tcgen05 tiles) and vocab rows [128r, 128r+128) of W_out (one
//   tile).  All 192 KB of bf16 weights live in TENSOR MEMORY for the whole kernel (3 x 128
//   columns, written once with tcgen05.st) and feed the MMAs as the TMEM A operand, so a
//   step re-reads only the 1 KB/MMA B operand from shared memory instead of 4 KB/MMA of
//   weights (with N = 32 the SMEM operand fetch, not the tensor pipe, was the MMA bound).
//
//   The MMAs are "swapped": D[gate row, sequence] = W[gate row, :] . h[sequence, :], so the
//   weights are the (resident) A operand with M = 128 and the 32 sequences are the N = 32
//   B operand (h as bf16, 16 KB, double buffered).  Accumulators live in TMEM (3 x 32 cols).
//
//   prologue:   Gctx = W_ih0[:, E:2E] enc^T + b_ih0 + b_hh0 as the kernel's first two MMA tiles (the W_ctx tiles borrow
//               the tensor-memory columns of W_hh, the bf16 copy of the cluster's encoder rows borrows h buffer 0):
//               every epilogue thread ends up with the 2 x 16 per-sequence constants it adds each step
//   per step:   MMA-G  gates  = W_hh h_s            (tcgen05, issued by one thread)
//               Epi-G  + Gtok[tok_s] + Gctx, sigmoid/tanh (MUFU.TANH), c/h update in
//                      registers; i/g and f/o of a unit sit in lanes l and l+16 of one warp
//                      so sigma(i)tanh(g) moves by one shuffle; h_{s+1} slice (4 KB, one
//                      K-block of the B operand) is written locally and pushed to the three
//                      peers with cp.async.bulk shared::cta -> shared::cluster (DSMEM),
//                      completing on the peer's mbarrier -- no cluster barrier per step
//               MMA-L  logits = W_out h_{s+1}       (overlaps nothing: it is the critical path)
//               MMA-G  for step s+1 is issued right behind it and overlaps Epi-L
//               Epi-L  + bias, per-warp argmax with redux.sync.max.f32 + ballot, CTA
//                      partials exchanged through DSMEM stores + remote mbarrier arrives
//   Token append, EOS flags and both reference stop rules are handled on the device.
//
// Reference semantics: LSTMDecoder.decode_step (model/decoder.py:197-284) inside
// Seq2SeqModel._greedy_search (model/seq2seq.py:192-232); attention with src_len == 1 is the
// identity and W_ih [emb ; ctx] is hoisted into the Gtok table / Gctx (F4).
#include "decode_persistent_common.cuh"
#include "sample_select.cuh"
#include <limits.h>

// Timing ablations (tools/ablate_greedy.sh builds one library per mask; results are WRONG by design):
//  1 gtok gather from a fixed row   2 no MUFU in the cell update   4 no cluster token exchange
//  8 no cluster h exchange          16 no argmax                   32 no logits MMA wait (use stale accumulator)
#ifndef I2L_ABL
#define I2L_ABL 0
#endif
#ifndef I2L_TANH_F16X2
#define I2L_TANH_F16X2 0     // 1 = cell-update activations as tanh.approx.f16x2.  NOT a win on sm_100a: ptxas emits TWO
                             // MUFU.TANH.F16 (.H0 / .H1) per f16x2, so the XU-pipe op count is unchanged (cuobjdump -sass)
#endif

namespace i2l {

namespace {

// shared memory map (bytes)
constexpr int OFF_H = 0;                            // 2 h buffers (B operand, K-major SWIZZLE_128B)
constexpr int OFF_PART = OFF_H + 2 * HB_BYTES;      // [8 warps][16 cols] (float,int)
constexpr int OFF_XCHG = OFF_PART + 8 * 16 * 8;     // [4 ctas][32 cols] (float,int)
constexpr int OFF_TOK = OFF_XCHG + CL * NB * 8;     // [32] int tokens of the current step
constexpr int OFF_BAR = OFF_TOK + NB * 4;           // mbarriers
constexpr int OFF_MISC = OFF_BAR + 16 * 8;           // tmem base, exit flag
constexpr int SMEM_BYTES = OFF_MISC + 16;
// sampling mode only: the logits of a step are regrouped so that ONE warp holds a whole row
constexpr int LG_BLOCK_BYTES = (NB / CL) * 128 * 4;  // 8 sequences x 128 vocabulary rows, fp32 = 4 KB
constexpr int OFF_STAGE = (SMEM_BYTES + 127) / 128 * 128;          // [32 sequences][128 own vocab rows] fp32 (bulk-copy source)
constexpr int OFF_RX = OFF_STAGE + CL * LG_BLOCK_BYTES;            // [4 source CTAs][8 own sequences][128] fp32
constexpr int SMEM_BYTES_SAMPLE = OFF_RX + CL * LG_BLOCK_BYTES;
static_assert(SMEM_BYTES_SAMPLE <= 232448, "shared memory budget exceeded");

// BAR_HS + 4 * buf + d: K-block of h buffer `buf` written by the CTA at cluster distance d (rank - d; d = 0: this CTA)
enum { BAR_W = 0, BAR_CTX = 1, BAR_LDONE = 3, BAR_GDONE = 4, BAR_TOK = 5, BAR_FINAL = 6, BAR_LG = 7, BAR_HS = 8 };

// debug build of the kernel only (tools/debug_persistent.py): per-phase clock64() stamps of one
// step of cluster 0 / rank 0, and optional dumps of gates / h / logits of that step
#define I2L_TS(slot)                                                                       \
  do {                                                                                     \
    if (DBG && dbg_ts) reinterpret_cast<long long*>(P.dbg + 16 + 300000)[slot] = clock64(); \
  } while (0)

#if I2L_ABL & 2
#define TANH(x) ((x) * 0.25f)
#else
#define TANH(x) tanh_approx(x)
#endif

struct Params {
  const unsigned char* wimg;     // per-rank weight images
  const float* gtok;             // [V][4][128][2]
  const float* bias;             // [512]
  const unsigned char* wctx;     // per-rank W_ctx images (2 tiles, the gate tiles' row order)
  const float* enc;              // [B][E] fp32 encoder output
  const float* bsum0;            // [1024] b_ih0 + b_hh0 (PyTorch gate order)
  int64_t* tokens;               // [B][T+1]
  int* first_end;                // [B]
  unsigned char* allend;         // [n_clusters][T]
  int* cluster_steps;            // [n_clusters]
  int B, T, start_id, end_id, stop_rule;
  float temperature;
  float* dbg;                    // debug kernel only
  int dbg_step, dbg_dump;
  // sampling mode (Predictor.predict_batch, training/predictor.py:295-335)
  int V, top_k, do_sample;
  float top_p;
  unsigned long long seed, offset;
  const float* uniforms;         // [T][B] or null (Philox)
  float* probs_trace;            // [T][B][V] or null
  // teacher-forced mode (LSTMDecoder.forward, model/decoder.py:100-195)
  const int64_t* tok_t;          // [T][B] given input tokens (already range-checked)
  float* logits_out;             // [B][T][V]
  float* h_out; float* c_out;    // [B][H] final state, or null
};

// DBG: 0 = production, 1 = clock stamps, 2 = stamps + value dumps.  MODE: 0 = greedy argmax (partials combined across
// the cluster), 1 = temperature / top-k / top-p sampling: every CTA ships its 128-row logits slice of 8 sequences to the
// CTA that owns them (3 x 4 KB DSMEM bulk copies, the h-exchange pattern), each of the 32 epilogue warps of the cluster
// then runs the warp-level selection (sample_select.cuh) for ONE sequence and broadcasts its token with st.async;
// 2 = teacher forcing: the input token of every step is given, the logits of every step are written out, and there is
// no token exchange at all.
template <int DBG, int MODE>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(THREADS, 1) persistent_greedy_kernel(Params P) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster = (int)cluster_id_x();
  const int row0 = cluster * NB;
  uint32_t* misc = reinterpret_cast<uint32_t*>(smem + OFF_MISC);
  const uint32_t bar = sbase + OFF_BAR;
  auto BAR = [&](int i) { return bar + 8u * i; };

  if ((sbase & 1023u) != 0) __trap();

  if (tid == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(BAR(BAR_HS + i), 1);
    mbar_init(BAR(BAR_LDONE), 1);
    mbar_init(BAR(BAR_GDONE), 1);
    mbar_init(BAR(BAR_TOK), 1);
    mbar_init(BAR(BAR_FINAL), 1);
    mbar_init(BAR(BAR_LG), 1);
    mbar_init(BAR(BAR_CTX), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    misc[1] = 0;
  }
  if (warp == 8) {   // all 512 TMEM columns: 3 accumulators + 3 resident weight tiles
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(sbase + OFF_MISC) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // zero both h buffers (h_0 = 0, decoder.py:253-266)
  for (int i = tid; i < 2 * HB_BYTES / 16; i += THREADS) reinterpret_cast<uint4*>(smem + OFF_H)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = misc[0];
  const uint32_t TM_L = tmem + TC_L, TM_G0 = tmem + TC_G0, TM_G1 = tmem + TC_G1;
  const uint64_t dbase = DESC_HI | (uint64_t)(((sbase >> 4) & 0x3FFFu) | (1u << 16));
  // one 128-row tile of gates / logits: D[row, sequence] += W (TMEM A operand, 16 K-elements = 8 columns per MMA) h^T
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
  // 128 x 256 bf16 weight tile -> tensor memory: thread (quadrant q, lane) owns row 32q + lane; warps 0-3 write K
  // columns [0,64), warps 4-7 columns [64,128) (two bf16 per column)
  auto tile_to_tmem = [&](const unsigned char* tile, uint32_t tcol) {
    const int prow = 32 * (warp & 3) + lane;
    const int halfk = warp >> 2;
    const uint4* src = reinterpret_cast<const uint4*>(tile + (size_t)prow * (H * 2)) + halfk * 16;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t r[16];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 v = __ldg(src + c * 4 + i);
        r[4 * i] = v.x; r[4 * i + 1] = v.y; r[4 * i + 2] = v.z; r[4 * i + 3] = v.w;
      }
      tc_st16(tmem + ((uint32_t)(32 * (warp & 3)) << 16) + tcol + halfk * 64 + c * 16, r);
    }
  };
  // ---- context term of the gates, once per sequence (attention over src_len == 1 is the identity):
  // Gctx = W_ih0[:, E:2E] enc^T + b_ih0 + b_hh0 as the kernel's first two MMA tiles.  The W_ctx tiles borrow the tensor-
  // memory columns of the W_hh tiles, the bf16 copy of the cluster's 32 encoder rows borrows h buffer 0; every epilogue
  // thread ends up with exactly the 2 x 16 values it adds every step (the separate f32->bf16 + GEMM launches and the
  // 4 MB round trip through HBM of round 1 are gone).
  float gctx0[16], gctx1[16];
  {
    if (warp < 8) {
      tile_to_tmem(P.wctx + ((size_t)rank * 2 + 0) * 128 * (H * 2), TC_WG0);
      tile_to_tmem(P.wctx + ((size_t)rank * 2 + 1) * 128 * (H * 2), TC_WG1);
      tile_to_tmem(P.wimg + (size_t)rank * WROW_BYTES + (size_t)2 * 128 * (H * 2), TC_WO);     // W_out is not borrowed
    }
    // enc rows -> bf16 B operand (K-major SWIZZLE_128B, the layout the cell update writes h in)
    for (int i = tid; i < NB * (E / 8); i += THREADS) {
      const int n = i / (E / 8), ch = i % (E / 8);          // sequence, 8-element chunk
      const int row = row0 + n;
      uint4 o = make_uint4(0, 0, 0, 0);
      if (row < P.B) {
        const float4* src = reinterpret_cast<const float4*>(P.enc + (size_t)row * E + ch * 8);
        const float4 a = __ldg(src), b = __ldg(src + 1);
        __nv_bfloat162 h0 = __floats2bfloat162_rn(a.x, a.y), h1 = __floats2bfloat162_rn(a.z, a.w);
        __nv_bfloat162 h2 = __floats2bfloat162_rn(b.x, b.y), h3 = __floats2bfloat162_rn(b.z, b.w);
        o = make_uint4(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1), *reinterpret_cast<uint32_t*>(&h2),
                       *reinterpret_cast<uint32_t*>(&h3));
      }
      const int kb = ch >> 3, c8 = ch & 7;
      *reinterpret_cast<uint4*>(smem + OFF_H + kb * HSLICE_BYTES + n * 128 + ((c8 ^ (n & 7)) << 4)) = o;
    }
    if (warp < 8) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 8) {
      if (elect_one()) {
        issue_tile(TM_G0, TC_WG0, OFF_H);
        issue_tile(TM_G1, TC_WG1, OFF_H);
        tc_commit(BAR(BAR_CTX));
      }
      __syncwarp();
    }
    mbar_wait(BAR(BAR_CTX), 0);
    tc_fence_after();
    if (warp < 8) {
      const int q_ = warp & 3, cg_ = warp >> 2;
      const int u_ = 16 * q_ + (lane & 15);
      const bool hi_ = lane >= 16;
      uint32_t r0[16], r1[16];
      tc_ld16_nowait(TM_G0 + ((uint32_t)(32 * q_) << 16) + 16 * cg_, r0);
      tc_ld16_nowait(TM_G1 + ((uint32_t)(32 * q_) << 16) + 16 * cg_, r1);
      tc_wait_ld();
      const float b0 = P.bsum0[(hi_ ? 1 : 0) * 256 + 64 * rank + u_], b1 = P.bsum0[(hi_ ? 3 : 2) * 256 + 64 * rank + u_];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const bool live = row0 + 16 * cg_ + j < P.B;
        gctx0[j] = live ? __uint_as_float(r0[j]) + b0 : 0.f;
        gctx1[j] = live ? __uint_as_float(r1[j]) + b1 : 0.f;
      }
    }
    tc_fence_before();
    __syncthreads();                 // the context MMAs have completed and been read: their operands can be replaced
    tc_fence_after();
  }
  if (warp < 8) {
    // resident W_hh tiles -> tensor memory; h_0 = 0 back in buffer 0 (decoder.py:253-266)
    tile_to_tmem(P.wimg + (size_t)rank * WROW_BYTES + (size_t)0 * 128 * (H * 2), TC_WG0);
    tile_to_tmem(P.wimg + (size_t)rank * WROW_BYTES + (size_t)1 * 128 * (H * 2), TC_WG1);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    tc_fence_before();
  }
  for (int i = tid; i < HB_BYTES / 16; i += THREADS) reinterpret_cast<uint4*>(smem + OFF_H)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  __syncthreads();
  tc_fence_after();
  cluster_sync_all();   // every CTA's barriers are initialised before any remote traffic


  if (warp == 8) {
    // =========================== MMA issuer warp ===========================
    // the whole warp stays converged (waits are executed by all lanes); one elected lane issues
    tc_fence_after();
    if (elect_one()) {   // gates of step 0 from h_0 = 0 (buffer 0)
      issue_tile(TM_G0, TC_WG0, OFF_H);
      issue_tile(TM_G1, TC_WG1, OFF_H);
      tc_commit(BAR(BAR_GDONE));
    }
    __syncwarp();
    // The 16 logits MMAs cost ~45 cycles each on the tensor pipe (tools/probes/mma_rate_probe.cu: that is the floor of
    // a 128xNx16 MMA for N <= 64): the 4 of a K-block are issued as soon as THAT block of h has arrived -- the CTA's
    // own block first, then the peers' in the order their bulk copies were sent -- so most of the chain runs under
    // the DSMEM exchange instead of after it.
    bool stop = false;
    for (int s = 0; s < P.T && !stop; ++s) {
      const bool dbg_ts = DBG && cluster == 0 && rank == 0 && s == P.dbg_step && lane == 0;
      const int nb = (s + 1) & 1;                       // buffer holding h_{s+1}
      const uint32_t hb = OFF_H + nb * HB_BYTES;
      I2L_TS(16);
#pragma unroll
      for (uint32_t d = 0; d < CL; ++d) {
        mbar_wait(BAR(BAR_HS + 4 * nb + d), (uint32_t)((s >> 1) & 1));     // h_j (j = s+1) is use (j-1)/2 of its buffer
        if (d == 0 && *reinterpret_cast<volatile uint32_t*>(&misc[1])) { stop = true; break; }
        tc_fence_after();
        const uint32_t kb = (rank - d) & (CL - 1);      // K-block = rank of the CTA that produced it
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t bd = dbase + (uint64_t)((hb + kb * HSLICE_BYTES + k * 32) >> 4);
            tc_mma_ts(TM_L, tmem + TC_WO + (kb * 4 + k) * 8, bd, IDESC, (d | (uint32_t)k) ? 1u : 0u);
          }
          if (d == CL - 1) tc_commit(BAR(BAR_LDONE));
        }
        __syncwarp();
      }
      if (stop) break;
      I2L_TS(17);
      if (s + 1 < P.T && elect_one()) {
        issue_tile(TM_G0, TC_WG0, hb);                  // gates of step s+1
        issue_tile(TM_G1, TC_WG1, hb);
        tc_commit(BAR(BAR_GDONE));
      }
      __syncwarp();
      I2L_TS(18);
    }
    if (elect_one()) tc_commit(BAR(BAR_FINAL));         // every MMA issued above has completed before TMEM is released
    __syncwarp();
    mbar_wait(BAR(BAR_FINAL), 0);
  } else {
    // =========================== epilogue warps (256 threads) ===========================
    const int q = warp & 3, cg = warp >> 2;               // TMEM lane quadrant, column group
    const int p = 32 * q + lane;                          // accumulator row (TMEM lane)
    const int u = 16 * q + (lane & 15);                   // CTA-local hidden unit
    const bool hi = lane >= 16;                           // lanes 16-31 hold f / o and the cell state
    const uint32_t lane_addr = (uint32_t)(32 * q) << 16;
    const int col0 = 16 * cg;
    float c[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) c[j] = 0.f;
    const float bias = P.bias[128 * rank + p];
    const float s1 = hi ? 0.5f : 1.0f, m1 = hi ? 0.5f : 1.0f, b1 = hi ? 0.5f : 0.0f;
    int tok[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) tok[j] = P.start_id;
    constexpr int XW = 1;                                 // epilogue warp that runs the token exchange (not warp 0:
    const int xt = tid - 32 * XW;                         //  thread 0 issues the h bulk copies, warp 8 shares its scheduler)
    int fe = -1;                                          // first END position (exchange warp of rank 0)
    bool finished = false;
    float* part = reinterpret_cast<float*>(smem + OFF_PART);
    int* tok_s = reinterpret_cast<int*>(smem + OFF_TOK);
    const float2* gt_base = reinterpret_cast<const float2*>(P.gtok + (size_t)rank * 256) + p;   // [V][rank][128 rows][tile 0,1]
    int s = 0;
    if (MODE != 2 && rank == 0 && xt >= 0 && xt < NB && row0 + xt < P.B) P.tokens[(size_t)(row0 + xt) * (P.T + 1)] = P.start_id;
    float hlast[16];                                      // MODE 2: h of the last step (fp32, before the bf16 rounding)
    if (MODE == 2) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int row = row0 + col0 + j;
        tok[j] = row < P.B ? (int)P.tok_t[row] : 0;
        hlast[j] = 0.f;
      }
    }
    for (; s < P.T; ++s) {
      const bool dbg_ts = DBG && cluster == 0 && rank == 0 && s == P.dbg_step;
      const bool dbg_dump = DBG == 2 && s == P.dbg_step;
      // ---------------- Epi-G(s): gates -> c_{s+1}, h_{s+1} ----------------
      if (tid == 0) I2L_TS(0);
      int tok_n[16];                                      // MODE 2: next step's tokens, fetched a step ahead
      if (MODE == 2 && s + 1 < P.T) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int row = row0 + col0 + j;
          tok_n[j] = row < P.B ? (int)P.tok_t[(size_t)(s + 1) * P.B + row] : 0;
        }
      }
      float gt0[16], gt1[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {                      // token -> gate table rows (L2 resident)
        const float2 g = __ldg(gt_base + (size_t)((I2L_ABL & 1) ? 1 : tok[j]) * 512);
        gt0[j] = g.x;
        gt1[j] = g.y;
      }
      if (tid == 0) I2L_TS(1);
      mbar_wait(BAR(BAR_GDONE), s & 1);
      tc_fence_after();
      if (tid == 0) I2L_TS(2);
      uint32_t r0[16], r1[16];
      tc_ld16_nowait(TM_G0 + lane_addr + col0, r0);
      tc_ld16_nowait(TM_G1 + lane_addr + col0, r1);
      tc_wait_ld();
      if (tid == 0) I2L_TS(3);
      const int nb = (s + 1) & 1;
      unsigned char* hdst = smem + OFF_H + nb * HB_BYTES + rank * HSLICE_BYTES;
      // phase 1: gate activations for all 16 sequences (32 independent MUFU.TANH)
      float y0[16], y1[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float x0 = __uint_as_float(r0[j]) + gt0[j] + gctx0[j];
        float x1 = __uint_as_float(r1[j]) + gt1[j] + gctx1[j];
        if (DBG == 2 && dbg_dump) {
          float* da = P.dbg + 16 + (size_t)((cluster * 4 + rank) * 2) * 4096;
          da[p * 32 + col0 + j] = x0; da[4096 + p * 32 + col0 + j] = x1;
          float* dm = P.dbg + 16 + 200000 + (size_t)((cluster * 4 + rank) * 2) * 4096;
          dm[p * 32 + col0 + j] = __uint_as_float(r0[j]); dm[4096 + p * 32 + col0 + j] = __uint_as_float(r1[j]);
        }
#if I2L_TANH_F16X2
        float t0, t1;
        tanh2_f16(0.5f * x0, s1 * x1, t0, t1);                     // one MUFU for both gates of the row
        y0[j] = fmaf(t0, 0.5f, 0.5f);                              // sigmoid(i) | sigmoid(f)
        y1[j] = fmaf(t1, m1, b1);                                  // tanh(g)    | sigmoid(o)
#else
        y0[j] = fmaf(TANH(0.5f * x0), 0.5f, 0.5f);                 // sigmoid(i) | sigmoid(f)
        y1[j] = fmaf(TANH(s1 * x1), m1, b1);                       // tanh(g)    | sigmoid(o)
#endif
      }
      // phase 2: sigma(i) tanh(g) moves from lanes 0..15 to the lanes 16..31 that own c
      float pg[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) pg[j] = __shfl_xor_sync(0xffffffffu, y0[j] * y1[j], 16);
      // phase 3: cell / hidden update (lanes 16..31 are the owners; 0..15 compute don't-cares)
      float hn[16];
#if I2L_TANH_F16X2
#pragma unroll
      for (int j = 0; j < 16; j += 2) {
        const float ca = fmaf(y0[j], c[j], pg[j]), cb = fmaf(y0[j + 1], c[j + 1], pg[j + 1]);
        c[j] = ca; c[j + 1] = cb;
        float ta, tb;
        tanh2_f16(ca, cb, ta, tb);
        hn[j] = y1[j] * ta; hn[j + 1] = y1[j + 1] * tb;
        if (MODE == 2) { hlast[j] = hn[j]; hlast[j + 1] = hn[j + 1]; }
      }
#else
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float cn = fmaf(y0[j], c[j], pg[j]);
        c[j] = cn;
        hn[j] = y1[j] * TANH(cn);
        if (MODE == 2) hlast[j] = hn[j];
      }
#endif
      if (hi) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int n = col0 + j;                                  // B-operand row (sequence)
          const int chunk = (u >> 3) ^ (n & 7);                    // SWIZZLE_128B
          *reinterpret_cast<__nv_bfloat16*>(hdst + n * 128 + chunk * 16 + (u & 7) * 2) = __float2bfloat16(hn[j]);
          if (DBG == 2 && dbg_dump) P.dbg[16 + 65536 + (cluster * 32 + n) * 256 + 64 * rank + u] = hn[j];
        }
      }
      if (tid == 0) I2L_TS(4);
      fence_proxy_async();           // generic-proxy h writes -> visible to tcgen05.mma and bulk copies
      tc_fence_before();
      epi_bar_sync();
      if (tid == 0) {
        const uint32_t src = sbase + OFF_H + nb * HB_BYTES + rank * HSLICE_BYTES;
        mbar_arrive(BAR(BAR_HS + 4 * nb));                          // own K-block is in place
#pragma unroll
        for (uint32_t d = 1; d < CL; ++d) {
          // arm the barrier of the block that arrives from distance d, send ours to the CTA at distance d
          if (I2L_ABL & 8) { mbar_arrive(BAR(BAR_HS + 4 * nb + d)); continue; }
          mbar_arrive_expect_tx(BAR(BAR_HS + 4 * nb + d), HSLICE_BYTES);
          const uint32_t peer = (rank + d) & (CL - 1);
          bulk_s2peer(mapa(src, peer), src, HSLICE_BYTES, mapa(BAR(BAR_HS + 4 * nb + d), peer));
        }
      }
      // ---------------- Epi-L(s): logits -> tok_{s+1} ----------------
      if (tid == 0) I2L_TS(5);
      if (!(I2L_ABL & 32)) mbar_wait(BAR(BAR_LDONE), s & 1);
      tc_fence_after();
      if (tid == 0) I2L_TS(6);
      float lg[16];
      tc_ld16(TM_L + lane_addr + col0, lg);
      tc_fence_before();
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        lg[j] += bias;
        if (DBG == 2 && dbg_dump) P.dbg[16 + 131072 + ((cluster * 4 + rank) * 128 + p) * 32 + col0 + j] = lg[j];
        if (MODE == 0 && P.temperature != 1.0f) lg[j] = lg[j] / P.temperature;   // seq2seq.py:213-214
      }
      if (MODE == 2) {
        // ---- teacher forcing: logits of this step straight to (B,T,V); lanes = consecutive vocabulary rows
        const int v = 128 * (int)rank + p;
        if (v < P.V) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int row = row0 + col0 + j;
            if (row < P.B) P.logits_out[((size_t)row * P.T + s) * P.V + v] = lg[j];
          }
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) tok[j] = tok_n[j];
        continue;
      }
      if (MODE == 1) {
        // ---- regroup: this CTA's [128 vocab rows] x [32 sequences] slice -> whole rows at the owning warps
        float* stage = reinterpret_cast<float*>(smem + OFF_STAGE);
        const float* rx = reinterpret_cast<const float*>(smem + OFF_RX);
#pragma unroll
        for (int j = 0; j < 16; ++j) stage[(col0 + j) * 128 + p] = lg[j];
        if (xt == 0) mbar_arrive_expect_tx(BAR(BAR_TOK), NB * 4);     // 32 tokens will be stored here this step
        fence_proxy_async();
        epi_bar_sync();
        if (tid == 0) {
          mbar_arrive_expect_tx(BAR(BAR_LG), (CL - 1) * LG_BLOCK_BYTES);
#pragma unroll
          for (uint32_t d = 1; d < CL; ++d) {
            const uint32_t peer = (rank + d) & (CL - 1);              // owner of sequences [8 peer, 8 peer + 8)
            bulk_s2peer(mapa(sbase + OFF_RX + rank * LG_BLOCK_BYTES, peer), sbase + OFF_STAGE + peer * LG_BLOCK_BYTES,
                        LG_BLOCK_BYTES, mapa(BAR(BAR_LG), peer));
          }
        }
        mbar_wait(BAR(BAR_LG), s & 1);
        // warp w owns sequence n = 8 rank + w; vocabulary entry 16 lane + i lives in source CTA lane / 8, row 16 (lane % 8) + i
        const int n = (NB / CL) * (int)rank + warp;
        const int srcr = lane >> 3;
        const float* src = (srcr == (int)rank) ? stage + n * 128 : rx + (srcr * (NB / CL) + warp) * 128;
        float x[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 v4 = *reinterpret_cast<const float4*>(src + 16 * (lane & 7) + 4 * i);
          x[4 * i] = v4.x; x[4 * i + 1] = v4.y; x[4 * i + 2] = v4.z; x[4 * i + 3] = v4.w;
        }
        const int row = row0 + n;
        float u = 0.f;
        if (P.do_sample && row < P.B)
          u = P.uniforms ? P.uniforms[(size_t)s * P.B + row] : philox_uniform(P.seed, P.offset + (uint64_t)s * P.B + row);
        float* pout = (P.probs_trace && row < P.B) ? P.probs_trace + ((size_t)s * P.B + row) * P.V : nullptr;
        const int chosen = warp_sample_select(x, P.V, lane, P.temperature, P.top_k, P.top_p, P.do_sample, u, pout);
        if (lane == 0) {
#pragma unroll
          for (uint32_t d = 0; d < CL; ++d)
            st_async_b32(mapa(sbase + OFF_TOK + n * 4, d), (uint32_t)chosen, mapa(BAR(BAR_TOK), d));
        }
      } else
      // per-column argmax over the warp's 32 vocab rows: order-preserving integer keys, one
      // redux.sync.max.s32 + one ballot per column (uniform results, ~6 instructions per column).
      // The ballot keeps every row holding the maximum; the lowest one wins later (torch.argmax:
      // first index).  Lane 0 publishes the 16 (key, ballot) pairs of the warp.
      {
        int kmax[16]; unsigned bal[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          int key = __float_as_int(lg[j] + 0.0f);                    // -0 -> +0 (equal under torch's compare)
          key ^= (key >> 31) & 0x7fffffff;                           // monotone: float order == signed int order
          kmax[j] = (I2L_ABL & 16) ? key : redux_max_s32(key);
          bal[j] = (I2L_ABL & 16) ? 1u : __ballot_sync(0xffffffffu, key == kmax[j]);
        }
        if (tid == 0) I2L_TS(7);
        if (lane == 0) {
          uint4* dst = reinterpret_cast<uint4*>(part) + warp * 8;
#pragma unroll
          for (int j = 0; j < 8; ++j) dst[j] = make_uint4((uint32_t)kmax[2 * j], bal[2 * j], (uint32_t)kmax[2 * j + 1], bal[2 * j + 1]);
        }
      }
      if (MODE == 0) epi_bar_sync();
      if (tid == 0) I2L_TS(8);
      if (xt >= 0 && xt < NB) {
        int bk = INT_MIN, bi = 0;
        if (MODE == 1) {
          mbar_wait_cluster(BAR(BAR_TOK), s & 1);
          bi = *reinterpret_cast<volatile int*>(tok_s + xt);
        } else {
        const int cgrp = xt >> 4, j = xt & 15;
        const int2* pp = reinterpret_cast<const int2*>(part);
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) {                             // ascending row order: strict > keeps the first maximum
          const int2 e = pp[(cgrp * 4 + qq) * 16 + j];
          if (qq == 0 || e.x > bk) { bk = e.x; bi = 128 * (int)rank + 32 * qq + __ffs(e.y) - 1; }
        }
        // CTA partial -> slot [rank][column] of every CTA of the cluster; the store itself
        // signals the destination's mbarrier (st.async + complete_tx), no fence / arrive round trip
        const uint32_t slot = sbase + OFF_XCHG + (rank * NB + xt) * 8;
        if (xt == 0) I2L_TS(13);
        if (xt == 0) mbar_arrive_expect_tx(BAR(BAR_TOK), ((I2L_ABL & 4) ? 1 : CL) * NB * 8);
        if (xt == 0) I2L_TS(14);
#pragma unroll
        for (uint32_t d = 0; d < CL; ++d) {
          if ((I2L_ABL & 4) && d != rank) continue;
          st_async_v2(mapa(slot, d), (uint32_t)bk, (uint32_t)bi, mapa(BAR(BAR_TOK), d));
          if (d == 0 && xt == 0) I2L_TS(15);
        }
        if (xt == 0) I2L_TS(9);
        mbar_wait_cluster(BAR(BAR_TOK), s & 1);
        if (xt == 0) I2L_TS(10);
        const int2* xe = reinterpret_cast<const int2*>(smem + OFF_XCHG);
#pragma unroll
        for (int r = 0; r < CL; ++r) {                               // ascending rank = ascending vocab index
          const int2 e = xe[r * NB + xt];
          if (r == 0 || e.x > bk) { bk = e.x; bi = e.y; }
        }
        tok_s[xt] = bi;
        }
        // token append + EOS bookkeeping (seq2seq.py:216-221 / predictor.py:338-347)
        const int row = row0 + xt;
        const bool valid = row < P.B;
        const bool is_end = bi == P.end_id;
        if (valid && is_end && fe < 0) fe = s + 1;
        finished = finished || is_end;
        if (rank == 0 && valid) P.tokens[(size_t)row * (P.T + 1) + s + 1] = bi;
        const bool all_end = __all_sync(0xffffffffu, !valid || is_end);
        const bool all_fin = __all_sync(0xffffffffu, !valid || finished);
        if (xt == 0) {
          if (rank == 0) P.allend[(size_t)cluster * P.T + s] = all_end ? 1 : 0;
          if (P.stop_rule == I2L_STOP_ALL_FINISHED_STICKY && all_fin) misc[1] = 1;   // this cluster is done
        }
      }
      if (tid == 0) I2L_TS(11);
      epi_bar_sync();
      if (tid == 0) I2L_TS(12);
#pragma unroll
      for (int j = 0; j < 16; ++j) tok[j] = tok_s[col0 + j];
      if (*reinterpret_cast<volatile uint32_t*>(&misc[1])) { ++s; break; }
    }
    if (MODE == 2) {
      if (P.h_out != nullptr && hi) {                      // lanes 16..31 own (h, c) of unit u for 16 sequences
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int row = row0 + col0 + j;
          if (row < P.B) {
            P.h_out[(size_t)row * H + 64 * rank + u] = hlast[j];
            P.c_out[(size_t)row * H + 64 * rank + u] = c[j];
          }
        }
      }
    } else if (rank == 0 && xt >= 0 && xt < NB) {
      const int row = row0 + xt;
      if (row < P.B) {
        P.first_end[row] = fe;
        // positions past the cluster's last step keep the general path's "never written" value (the separate init
        // kernel of round 1 wrote -1 everywhere first)
        for (int t = s + 1; t <= P.T; ++t) P.tokens[(size_t)row * (P.T + 1) + t] = -1;
      }
      if (xt == 0) P.cluster_steps[cluster] = s;
    }
    if (*reinterpret_cast<volatile uint32_t*>(&misc[1]) && tid == 0) {
      // early exit: release the MMA thread that is waiting for an h buffer that will never fill
      mbar_arrive(BAR(BAR_HS + 4 * ((s + 1) & 1)));
    }
  }
  // ---- teardown: peers may still be reading / writing our shared memory ----
  tc_fence_before();
  cluster_sync_all();
  if (warp == 8) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

// ---- packing ---------------------------------------------------------------------------
__global__ void pack_weights_kernel(const float* __restrict__ w_hh, const float* __restrict__ out_w, int V,
                                    unsigned char* __restrict__ wimg) {
  // one thread per (rank, tile(0..2), row p, k): tiles 0,1 = gates, tile 2 = vocab rows
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)CL * 3 * 128 * H) return;
  int k = (int)(i % H);
  int p = (int)((i / H) % 128);
  int t = (int)((i / (H * 128)) % 3);
  int r = (int)(i / ((size_t)H * 128 * 3));
  float v;
  if (t < 2) {
    int qd = p >> 5, l = p & 31;
    int unit = 64 * r + 16 * qd + (l & 15);
    int gate = 2 * t + (l >= 16 ? 1 : 0);               // PyTorch gate order i,f,g,o
    v = w_hh[(size_t)(gate * H + unit) * H + k];
  } else {
    int vr = 128 * r + p;
    v = vr < V ? out_w[(size_t)vr * H + k] : 0.f;
  }
  size_t off = (size_t)r * WROW_BYTES + (((size_t)t * 128 + p) * H + k) * 2;
  *reinterpret_cast<__nv_bfloat16*>(wimg + off) = __float2bfloat16(v);
}

__global__ void pack_wctx_kernel(const float* __restrict__ w_ih0, unsigned char* __restrict__ img) {
  // one thread per (rank, tile(0..1), row p, k): W_ih0[:, E:2E] in the row order of the gate tiles (pack_weights_kernel)
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)CL * 2 * 128 * E) return;
  int k = (int)(i % E);
  int p = (int)((i / E) % 128);
  int t = (int)((i / (E * 128)) % 2);
  int r = (int)(i / ((size_t)E * 128 * 2));
  int qd = p >> 5, l = p & 31;
  int unit = 64 * r + 16 * qd + (l & 15);
  int gate = 2 * t + (l >= 16 ? 1 : 0);
  reinterpret_cast<__nv_bfloat16*>(img)[i] = __float2bfloat16(w_ih0[(size_t)(gate * H + unit) * (2 * E) + E + k]);
}

__global__ void pack_gtok_kernel(const float* __restrict__ gtok, int V, float* __restrict__ out) {
  // out[v][r][p][t] = gtok[v][gate(t,p)*256 + 64 r + unit(p)]   (both tiles of a row in one 8-byte load)
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)V * 1024) return;
  int p = (int)(i & 127), t = (int)((i >> 7) & 1), r = (int)((i >> 8) & 3), v = (int)(i >> 10);
  int qd = p >> 5, l = p & 31;
  int unit = 64 * r + 16 * qd + (l & 15);
  int gate = 2 * t + (l >= 16 ? 1 : 0);
  out[(size_t)v * 1024 + r * 256 + p * 2 + t] = gtok[(size_t)v * 1024 + gate * 256 + unit];
}

__global__ void pack_bias_kernel(const float* __restrict__ out_b, int V, float* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < VMAX) out[i] = i < V ? out_b[i] : -INFINITY;
}

__global__ void persistent_init_kernel(int64_t* tokens, int T1, int B, int start_id) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (size_t)B * T1) tokens[i] = (i % T1 == 0) ? start_id : -1;
}

__global__ void persistent_finalize_kernel(const int* first_end, const unsigned char* allend, const int* cluster_steps,
                                           int n_clusters, int B, int T, int stop_rule, int32_t* lengths,
                                           int32_t* steps_out) {
  // one block: steps_run = first step at which every cluster reported "all rows emitted END"
  // (ALL_END_SAME_STEP), or the last cluster to finish (sticky rule); then lengths
  __shared__ int steps_sh;
  if (threadIdx.x == 0) steps_sh = T;
  __syncthreads();
  if (stop_rule == I2L_STOP_ALL_END_SAME_STEP) {
    for (int s = threadIdx.x; s < T; s += blockDim.x) {
      bool all = true;
      for (int cidx = 0; cidx < n_clusters && all; ++cidx) all = allend[(size_t)cidx * T + s] != 0;
      if (all) atomicMin(&steps_sh, s + 1);
    }
  } else if (stop_rule == I2L_STOP_ALL_FINISHED_STICKY) {
    if (threadIdx.x == 0) {
      int steps = 0;
      for (int cidx = 0; cidx < n_clusters; ++cidx) steps = max(steps, cluster_steps[cidx]);
      steps_sh = steps;
    }
  }
  __syncthreads();
  const int steps = steps_sh;
  if (threadIdx.x == 0 && steps_out) *steps_out = steps;
  if (lengths)
    for (int i = threadIdx.x; i < B; i += blockDim.x) {
      int fe = first_end[i];
      lengths[i] = (fe >= 0 && fe <= steps) ? fe : steps + 1;
    }
}

struct PWs { float* gctx; __nv_bfloat16* encb; int* first_end; unsigned char* allend; int* cluster_steps; size_t bytes; };
PWs pcarve(int rows, int T, void* ws) {
  Arena a(ws, (size_t)-1);
  PWs w{};
  int ncl = cdiv(rows, NB);
  w.gctx = a.take<float>((size_t)rows * 1024);
  w.encb = a.take<__nv_bfloat16>((size_t)rows * E);
  w.first_end = a.take<int>(rows);
  w.allend = a.take<unsigned char>((size_t)ncl * (T > 0 ? T : 1));
  w.cluster_steps = a.take<int>(ncl);
  w.bytes = align_up(a.off, 256);
  return w;
}

}  // namespace

#ifdef I2L_DIAG   // diagnostics library only: the DBG instantiations (clock stamps / value dumps of one step) and their host sync
static float* g_dbg_buf = nullptr;   // host-side: when set, the DBG instantiation of the kernel is launched
int persistent_set_debug(float* buf) {
  g_dbg_buf = buf;
  return I2L_OK;
}
#endif

bool persistent_supported(const i2l_dec_desc& d) {
  return d.hidden_dim == H && d.embedding_dim == E && d.lstm_layers == 1 && d.vocab_size <= VMAX && d.vocab_size >= 1;
}
size_t persistent_packed_bytes(const i2l_dec_desc& d) { return psection(d.vocab_size).total; }

int persistent_pack(const i2l_dec_desc& d, const i2l_dec_params& p, const float* gtok_f32, void* section,
                    cudaStream_t s) {
  // gtok_f32: the fp32 token->gate table (V,4H) already produced on this stream by i2l_dec_pack
  PSection ps = psection(d.vocab_size);
  {
    size_t n = (size_t)d.vocab_size * 1024;
    pack_gtok_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(
        gtok_f32, d.vocab_size, reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(section) + ps.gtok));
    I2L_LAUNCH_OK();
  }
  unsigned char* sec = reinterpret_cast<unsigned char*>(section);
  size_t n = (size_t)CL * 3 * 128 * H;
  pack_weights_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(p.w_hh[0], p.out_w, d.vocab_size, sec + ps.wimg);
  I2L_LAUNCH_OK();
  pack_bias_kernel<<<2, 256, 0, s>>>(p.out_b, d.vocab_size, reinterpret_cast<float*>(sec + ps.bias));
  I2L_LAUNCH_OK();
  n = (size_t)CL * 2 * 128 * E;
  pack_wctx_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(p.w_ih[0], sec + ps.wctx);
  I2L_LAUNCH_OK();
  return I2L_OK;
}

size_t persistent_workspace_bytes(const i2l_dec_desc&, int rows, int max_length) { return pcarve(rows, max_length, nullptr).bytes; }

int persistent_greedy(const i2l_dec_desc& d, const void* section, const float* packed_f32, const PackedDec& lay,
                      const float* enc, int batch, int start_id, int end_id, int max_length, float temperature,
                      int stop_rule, int64_t* tokens, int32_t* lengths, int32_t* steps_run, void* ws, size_t ws_bytes,
                      cudaStream_t s, const PersistentSampleArgs* sample) {
  I2L_REQUIRE(start_id >= 0 && start_id < d.vocab_size, "decode loop: start token out of range");
  PWs w = pcarve(batch, max_length, ws);
  if (ws_bytes < w.bytes) { set_error("persistent_greedy: workspace too small"); return I2L_ERR_WORKSPACE; }
  PSection ps = psection(d.vocab_size);
  const unsigned char* sec = reinterpret_cast<const unsigned char*>(section);
  const int T1 = max_length + 1;
  if (max_length <= 0) {            // no kernel: the START column alone
    persistent_init_kernel<<<(unsigned)((batch * (size_t)T1 + 255) / 256), 256, 0, s>>>(tokens, T1, batch, start_id);
    I2L_LAUNCH_OK();
  }
  // (round 1 launched a token-init kernel, an f32 -> bf16 copy of enc and the context-term GEMM here; all three are now
  // part of the persistent kernel's prologue / tail)
  const int ncl = cdiv(batch, NB);
  if (max_length > 0) {
    Params P{};
    P.wimg = sec + ps.wimg; P.gtok = reinterpret_cast<const float*>(sec + ps.gtok);
    P.bias = reinterpret_cast<const float*>(sec + ps.bias);
    P.wctx = sec + ps.wctx; P.enc = enc; P.bsum0 = packed_f32 + lay.bsum[0];
    P.tokens = tokens; P.first_end = w.first_end; P.allend = w.allend; P.cluster_steps = w.cluster_steps;
    P.B = batch; P.T = max_length; P.start_id = start_id; P.end_id = end_id; P.stop_rule = stop_rule;
    P.temperature = temperature;
    P.V = d.vocab_size;
    if (sample) {
      P.top_k = sample->top_k; P.top_p = sample->top_p; P.do_sample = sample->do_sample;
      P.seed = sample->seed; P.offset = sample->offset; P.uniforms = sample->uniforms; P.probs_trace = sample->probs_trace;
    }
    KernelTimer kt(sample ? "dec.sample_persistent" : "dec.greedy_persistent", s);
    if (sample) {
      I2L_CUDA_OK(cudaFuncSetAttribute(persistent_greedy_kernel<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES_SAMPLE));
      persistent_greedy_kernel<0, 1><<<ncl * CL, THREADS, SMEM_BYTES_SAMPLE, s>>>(P);
#ifdef I2L_DIAG
    } else if (g_dbg_buf != nullptr) {
      float hdr[2];
      I2L_CUDA_OK(cudaMemcpyAsync(hdr, g_dbg_buf, sizeof(hdr), cudaMemcpyDeviceToHost, s));
      I2L_CUDA_OK(cudaStreamSynchronize(s));
      P.dbg = g_dbg_buf; P.dbg_step = (int)hdr[0]; P.dbg_dump = (int)hdr[1];
      if (P.dbg_dump) {
        I2L_CUDA_OK(cudaFuncSetAttribute(persistent_greedy_kernel<2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        persistent_greedy_kernel<2, 0><<<ncl * CL, THREADS, SMEM_BYTES, s>>>(P);
      } else {
        I2L_CUDA_OK(cudaFuncSetAttribute(persistent_greedy_kernel<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        persistent_greedy_kernel<1, 0><<<ncl * CL, THREADS, SMEM_BYTES, s>>>(P);
      }
#endif
    } else {
      I2L_CUDA_OK(cudaFuncSetAttribute(persistent_greedy_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
      persistent_greedy_kernel<0, 0><<<ncl * CL, THREADS, SMEM_BYTES, s>>>(P);
    }
    I2L_LAUNCH_OK();
  } else {
    I2L_CUDA_OK(cudaMemsetAsync(w.first_end, 0xff, (size_t)batch * 4, s));
    I2L_CUDA_OK(cudaMemsetAsync(w.cluster_steps, 0, (size_t)ncl * 4, s));
  }
  persistent_finalize_kernel<<<1, 256, 0, s>>>(w.first_end, w.allend, w.cluster_steps, ncl, batch, max_length, stop_rule,
                                               lengths, steps_run);
  I2L_LAUNCH_OK();
  return I2L_OK;
}

// Teacher-forced pass (LSTMDecoder.forward in eval mode) on the persistent kernel: zero initial state only.
int persistent_forward(const i2l_dec_desc& d, const void* section, const float* packed_f32, const PackedDec& lay,
                       const float* enc, const int64_t* tok_t, int batch, int seq_len, float* logits, float* h_out,
                       float* c_out, void* ws, size_t ws_bytes, cudaStream_t s) {
  PWs w = pcarve(batch, seq_len, ws);
  if (ws_bytes < w.bytes) { set_error("persistent_forward: workspace too small"); return I2L_ERR_WORKSPACE; }
  PSection ps = psection(d.vocab_size);
  const unsigned char* sec = reinterpret_cast<const unsigned char*>(section);
  Params P{};
  P.wimg = sec + ps.wimg; P.gtok = reinterpret_cast<const float*>(sec + ps.gtok);
  P.bias = reinterpret_cast<const float*>(sec + ps.bias);
  P.wctx = sec + ps.wctx; P.enc = enc; P.bsum0 = packed_f32 + lay.bsum[0];
  P.B = batch; P.T = seq_len; P.V = d.vocab_size; P.temperature = 1.0f;
  P.tok_t = tok_t; P.logits_out = logits; P.h_out = h_out; P.c_out = c_out;
  KernelTimer kt("dec.forward_persistent", s);
  I2L_CUDA_OK(cudaFuncSetAttribute(persistent_greedy_kernel<0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  persistent_greedy_kernel<0, 2><<<cdiv(batch, NB) * CL, THREADS, SMEM_BYTES, s>>>(P);
  I2L_LAUNCH_OK();
  return I2L_OK;
}

}  // namespace i2l
