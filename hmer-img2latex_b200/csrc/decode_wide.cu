// Persistent greedy decode for the decoders that do not fit the 4-CTA cluster kernel (decode_persistent.cu): any
// number of LSTM layers L <= 4, H % 64 == 0 -- the reference's shipped configurations E = H = 512, L = 2
// (img2latex/configs/config.yaml:45-50) and E = H = 1024, L = 3 (configs/resnet_lstm.yaml:45-50) among them.
// Reference semantics: LSTMDecoder.decode_step (model/decoder.py:197-284, nn.LSTM stack decoder.py:76-82) inside
// Seq2SeqModel._greedy_search (model/seq2seq.py:192-232); attention over src_len == 1 is the identity and
// W_ih0 [emb ; ctx] is hoisted into the Gtok table / Gctx (F4).
//
// 6.5 MB of bf16 weights (512 / 512 / 2) cannot stay on one cluster, and a stream-ordered step is L + 2 launches of
// ~10 us each.  Here ONE cooperative kernel runs all max_length steps on the whole GPU.  A step is a chain of L + 1
// products  gates_l = [h_{l-1}(s) ; h_l(s-1)] W_l^T  (+ cell)  and  logits = h_{L-1}(s) W_out^T  (+ argmax); the
// sequences are independent, so the only dependence is between the tiles of ONE 128-sequence block:
//
//   tile      128 sequences x 128 gate columns (= 4 gates x 32 hidden units, weight rows pre-ordered: PackedDec::g16c_*)
//             or 128 sequences x 32 vocabulary columns; tile t of a phase belongs to CTA t % gridDim.x for the whole
//             loop (static ownership: no scheduler, the processing order (step, phase, tile) is a global order in which
//             every dependence points backwards, so the flag waits cannot deadlock on co-resident CTAs)
//   sync      one monotone counter per 128-sequence block, incremented by every finished tile of that block; a tile
//             waits for the count that says "all tiles of the phase I read have finished" (ld.acquire + proxy fence,
//             then TMA) -- no grid barrier, blocks drift apart freely
//   warp 0    TMA producer: weight tiles are requested BEFORE the flag wait (they do not depend on it), the h tiles
//             after it; layer l >= 1 runs its h_l(s-1) W_hh half first, which is a whole step old
//   warp 1    MMA issuer (tcgen05, fp32 accumulators in TMEM, two accumulator slots)
//   warps 2-9 epilogue (TMEM lane quadrant x 16 of the tile's 32 units): LSTM cell (MUFU.TANH activations as in the other
//             decode kernels) -> c (fp32) and h (bf16, double buffered by step parity) in L2-resident global memory.
//             A logits tile folds its columns of a row into a 64-bit atomicMax key (order-preserving logit bits, then
//             ~index: the lowest index wins ties like torch.argmax); the layer-0 epilogue of step s reads ONE key per
//             row = the token of step s-1 (every tile of the block does so redundantly; tile column 0 appends it and
//             keeps the EOS bookkeeping), then gathers that token's Gtok row.  Its MMAs do not wait for the token, only
//             its epilogue does.  A tile is published by ONE release (CTA barrier, then fence + atomic of one thread).
//   sampling  MODE 1 (Predictor.predict_batch, V <= 512): the logits tiles store whole rows, a fourth phase spreads the
//             128 rows of a block over its layer-tile owners, one epilogue warp per row runs sample_select.cuh
//   stop      STICKY: a block whose sequences have all finished publishes stop_at[block]; its later tiles are "dead"
//             (processed without waiting, nothing written).  ALL_END_SAME_STEP: per-step flags, resolved by the
//             finalize kernel, as in decode_persistent.cu.  Every spin is bounded (~2 s): on expiry an abort flag makes
//             all remaining tiles dead, the kernel ends and the finalize kernel reports steps_run = -1 (an exception on the host side).
#include "decode_kernels.cuh"
#include "tc_common.cuh"
#include "sample_select.cuh"
#include <limits.h>
#include <stdlib.h>

namespace i2l {

int gemm_bf16_operand_map(CUtensorMap* out, const void* base, int rows, int K, int ld_elems, int box_rows);   // gemm_bf16.cu

namespace {

using namespace tc;

constexpr int WD_BM = 128, WD_BN = 128, WD_BK = 64, WD_LN = 32;
constexpr int WD_STAGES = 6;
constexpr int WD_A_BYTES = WD_BM * WD_BK * 2, WD_W_BYTES = WD_BN * WD_BK * 2, WD_WL_BYTES = WD_LN * WD_BK * 2;
constexpr int WD_STAGE = WD_A_BYTES + WD_W_BYTES;
constexpr int WD_OFF_BAR = WD_STAGES * WD_STAGE;
constexpr int WD_MAXT = 4;                                  // layer tiles per CTA and phase (per-tile row state in smem)
constexpr int WD_OFF_STATE = WD_OFF_BAR + 256;              // [WD_MAXT][128] finished flags
constexpr int WD_OFF_MISC = WD_OFF_STATE + WD_MAXT * 128;   // epilogue scratch
constexpr int WD_SMEM = WD_OFF_MISC + 64;
constexpr int WD_MAXL = 4;
constexpr int WD_THREADS = 320;                             // producer + MMA issuer + 8 epilogue warps
constexpr long long WD_SPIN_CYCLES = 4000000000LL;          // ~2 s at 1.9 GHz

struct WideParams {
  CUtensorMap tmH;                  // h, bf16 [(parity * L + l) * Bp + row][H], box 64 x 128
  CUtensorMap tmW[2 * WD_MAXL];     // [2 l] = W_hh[l], [2 l + 1] = W_ih[l] (l >= 1): (4H, H) bf16, cell-row order
  CUtensorMap tmO;                  // W_out (V, H) bf16, box 64 x 32
  const float* gctx;                // [B][4H]  enc W_ctx^T + b_ih0 + b_hh0 (PyTorch gate order)
  const float* gtok;                // [V][4H]  W_ih0[:, :E] emb[v]
  const float* bsum[WD_MAXL];       // [4H]     b_ih + b_hh of layer l >= 1
  const float* out_b;               // [V]
  __nv_bfloat16* hbuf;
  float* c;                         // [L][Bp][H]
  unsigned long long* best;         // [2 step parities][MB][128]: max over a step's logits tiles of (ordered logit, ~index)
  unsigned* cnt;                    // [MB] finished tiles
  int* stop_at;                     // [MB] step at which the block stopped (INT_MAX-ish: running)
  int* abort_flag;
  int64_t* tokens;                  // [B][T+1]
  int* first_end;                   // [B]
  unsigned char* allend;            // [MB][T]
  int* block_steps;                 // [MB]
  int B, Bp, H, L, V, T, NT, NV, MB, KB;
  int start_id, end_id, stop_rule;
  float temperature;
  long long* dbg_ts;                // diagnostics build: clock64 stamps of CTA 0, [step][8]
  // sampling mode (Predictor.predict_batch, training/predictor.py:295-335): the logits tiles write whole rows, a
  // selection phase (one epilogue warp per row, sample_select.cuh: the routine of the other sampling paths) draws
  float* logits;                    // [Bp][VP] fp32, VP = 32 NV
  int* tokbuf;                      // [2 step parities][Bp]
  int VP, top_k, do_sample;
  float top_p;
  unsigned long long seed, offset;
  const float* uniforms;            // [T][B] or null (Philox)
  float* probs_trace;               // [T][B][V] or null
};

#ifdef I2L_DIAG
#define WD_TS(step, slot) do { if (P.dbg_ts != nullptr && blockIdx.x == 0) P.dbg_ts[(size_t)(step) * 8 + (slot)] = clock64(); } while (0)
#else
#define WD_TS(step, slot) do { } while (0)
#endif

__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ int ld_volatile_s32(const int* p) { return *reinterpret_cast<const volatile int*>(p); }

// true: the count was reached (the tile is alive); false: the block stopped at or before step s, or the kernel aborted
__device__ __noinline__ bool wait_count(const WideParams& P, int mb, unsigned target, int s) {
  const long long t0 = clock64();
  for (;;) {
    if (ld_acquire_u32(P.cnt + mb) >= target) return true;
    if (ld_volatile_s32(P.stop_at + mb) <= s) return false;
    if (ld_volatile_s32(P.abort_flag) != 0) return false;
    if (clock64() - t0 > WD_SPIN_CYCLES) { atomicExch(P.abort_flag, 1); return false; }
  }
}
__device__ __forceinline__ bool block_dead(const WideParams& P, int mb, int s) {
  return ld_volatile_s32(P.stop_at + mb) <= s || ld_volatile_s32(P.abort_flag) != 0;
}
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// MODE 0: greedy argmax (64-bit atomicMax keys); MODE 1: temperature / top-k / top-p sampling (V <= 512)
template <int MODE>
__global__ void __launch_bounds__(WD_THREADS, 1) wide_loop_kernel(const __grid_constant__ WideParams P) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar = sbase + WD_OFF_BAR;
  auto FULL = [&](int s) { return bar + 8u * s; };
  auto EMPTY = [&](int s) { return bar + 8u * (WD_STAGES + s); };
  auto TFULL = [&](int a) { return bar + 8u * (2 * WD_STAGES + a); };
  auto TEMPTY = [&](int a) { return bar + 8u * (2 * WD_STAGES + 2 + a); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + WD_OFF_BAR + 8 * (2 * WD_STAGES + 4));
  unsigned char* fin = smem + WD_OFF_STATE;                    // [slot][row]: the row has emitted END
  volatile int* misc = reinterpret_cast<volatile int*>(smem + WD_OFF_MISC);
  if ((sbase & 1023u) != 0) __trap();
  if (tid == 0) {
    for (int s = 0; s < WD_STAGES; ++s) { mbar_init(FULL(s), 1); mbar_init(EMPTY(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(TFULL(a), 1); mbar_init(TEMPTY(a), 1); }
    mbar_fence_init();
    tma_prefetch_desc(&P.tmH); tma_prefetch_desc(&P.tmO);
    for (int i = 0; i < 2 * P.L; ++i) if (i != 1) tma_prefetch_desc(&P.tmW[i]);
  }
  for (int i = tid; i < WD_MAXT * 128; i += WD_THREADS) fin[i] = 0;
  if (warp == 1) tmem_alloc<256>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int G = gridDim.x, cta = blockIdx.x;
  const int L = P.L, NT = P.NT, NV = P.NV, KB = P.KB;
  const unsigned PS = (unsigned)(L * NT + NV + (MODE == 1 ? NT : 0));   // tiles (+ selection slices) of one block per step
  const int n_lt = P.MB * NT, n_vt = P.MB * NV;                // tiles per layer phase / logits phase

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int stage = 0; uint32_t ph = 0;
      auto slot = [&](uint32_t bytes) {
        mbar_wait(EMPTY(stage), ph ^ 1);
        mbar_arrive_expect_tx(FULL(stage), bytes);
        return sbase + (uint32_t)stage * WD_STAGE;
      };
      auto next = [&]() { if (++stage == WD_STAGES) { stage = 0; ph ^= 1; } };
      for (int s = 0; s < P.T; ++s) {
        const int par = s & 1;
        for (int l = 0; l < L; ++l) {
          for (int t = cta; t < n_lt; t += G) {
            const int mb = t / NT, nt = t - mb * NT;
            const int m0 = mb * WD_BM, n0 = nt * WD_BN;
            // recurrent half: h_l(s-1) W_hh[l]^T  (a whole step old; h_l(-1) = 0 sits in parity 1)
            bool ok = s == 0;
            for (int kb = 0; kb < KB; ++kb) {
              const uint32_t dst = slot(WD_STAGE);
              tma_load_2d(dst + WD_A_BYTES, &P.tmW[2 * l], kb * WD_BK, n0, FULL(stage));
              if (!ok) { wait_count(P, mb, (unsigned)(s - 1) * PS + (unsigned)(l + 1) * NT, s); asm volatile("fence.proxy.async;" ::: "memory"); ok = true; }
              tma_load_2d(dst, &P.tmH, kb * WD_BK, ((par ^ 1) * L + l) * P.Bp + m0, FULL(stage));
              next();
            }
            if (l > 0) {
              // input half: h_{l-1}(s) W_ih[l]^T  (this step's output of the layer below)
              ok = false;
              for (int kb = 0; kb < KB; ++kb) {
                const uint32_t dst = slot(WD_STAGE);
                tma_load_2d(dst + WD_A_BYTES, &P.tmW[2 * l + 1], kb * WD_BK, n0, FULL(stage));
                if (!ok) { wait_count(P, mb, (unsigned)s * PS + (unsigned)l * NT, s); asm volatile("fence.proxy.async;" ::: "memory"); ok = true; if (l == 1) WD_TS(s, 2); }
                tma_load_2d(dst, &P.tmH, kb * WD_BK, (par * L + l - 1) * P.Bp + m0, FULL(stage));
                next();
              }
            }
          }
        }
        for (int t = cta; t < n_vt; t += G) {
          const int mb = t / NV, nv = t - mb * NV;
          const int m0 = mb * WD_BM, n0 = nv * WD_LN;
          bool ok = false;
          for (int kb = 0; kb < KB; ++kb) {
            const uint32_t dst = slot(WD_A_BYTES + WD_WL_BYTES);
            tma_load_2d(dst + WD_A_BYTES, &P.tmO, kb * WD_BK, n0, FULL(stage));
            if (!ok) { wait_count(P, mb, (unsigned)s * PS + (unsigned)L * NT, s); asm volatile("fence.proxy.async;" ::: "memory"); ok = true; WD_TS(s, 5); }
            tma_load_2d(dst, &P.tmH, kb * WD_BK, (par * L + L - 1) * P.Bp + m0, FULL(stage));
            next();
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t IDESC_G = idesc_bf16(WD_BM, WD_BN), IDESC_L = idesc_bf16(WD_BM, WD_LN);
    const uint64_t d0 = desc_base(sbase, 128);
    int stage = 0; uint32_t ph = 0;
    uint32_t it = 0;
    auto run_tile = [&](int nkb, uint32_t idesc) {
      const uint32_t a = it & 1, apar = (it >> 1) & 1;
      mbar_wait(TEMPTY(a), apar ^ 1);
      tc_fence_after();
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(FULL(stage), ph);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < WD_BK / 16; ++ks) {
            const uint64_t ad = d0 + (uint64_t)((stage * WD_STAGE + ks * 32) >> 4);
            const uint64_t bd = d0 + (uint64_t)((stage * WD_STAGE + WD_A_BYTES + ks * 32) >> 4);
            tc_mma_ss(tmem + a * 128, ad, bd, idesc, (kb | ks) ? 1u : 0u);
          }
          tc_commit(EMPTY(stage));
          if (kb == nkb - 1) tc_commit(TFULL(a));
        }
        __syncwarp();
        if (++stage == WD_STAGES) { stage = 0; ph ^= 1; }
      }
      ++it;
    };
    for (int s = 0; s < P.T; ++s) {
      for (int l = 0; l < L; ++l)
        for (int t = cta; t < n_lt; t += G) run_tile(l == 0 ? KB : 2 * KB, IDESC_G);
      for (int t = cta; t < n_vt; t += G) run_tile(KB, IDESC_L);
    }
  } else {
    // ===================== epilogue (warps 2..9: TMEM lane quadrant = warp & 3, column half = (warp - 2) / 4) =========
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int r = 32 * q + lane;                               // row of the tile
    const int et = tid - 64;                                   // 0..255
    const uint32_t lane_addr = (uint32_t)(32 * q) << 16;
    const int H = P.H, T1 = P.T + 1;
    const bool sticky = P.stop_rule == I2L_STOP_ALL_FINISHED_STICKY;
    uint32_t it = 0;
    // token of a row = low word of the block's 64-bit (logit, ~index) maximum of the previous step
    auto token_of = [&](int mb, int step) {
      if (MODE == 1) return __ldcg(P.tokbuf + (size_t)(step & 1) * P.Bp + mb * WD_BM + r);
      const unsigned long long k = __ldcg(P.best + ((size_t)(step & 1) * P.MB + mb) * 128 + r);
      return (int)(0xFFFFFFFFu - (uint32_t)k);
    };
    // EOS bookkeeping of token `tok` = the output of step s - 1 (position s), by the 128 threads of column half 0; the
    // warp votes land in misc[0..3] (bit 0: every row emitted END at this step, bit 1: every row has finished)
    auto account = [&](int mb, int slot_i, bool owner, int s, int tok) {
      const int row = mb * WD_BM + r;
      const bool valid = row < P.B;
      const bool is_end = tok == P.end_id;
      unsigned char f = fin[slot_i * 128 + r];
      if (owner && valid) {
        P.tokens[(size_t)row * T1 + s] = tok;
        if (is_end && !f) P.first_end[row] = s;
      }
      f = f || is_end;
      fin[slot_i * 128 + r] = f;
      const bool w_end = __all_sync(0xffffffffu, !valid || is_end), w_fin = __all_sync(0xffffffffu, !valid || f);
      if (lane == 0) misc[q] = (w_end ? 1 : 0) | (w_fin ? 2 : 0);
    };
    // publish the tile: the CTA barrier orders every thread's global writes before the one release increment
    auto publish = [&](uint32_t a, int mb, bool dead) {
      tc_fence_before();
      epi_bar();
      if (et == 0) {
        mbar_arrive(TEMPTY(a));
        if (!dead) { __threadfence(); atomicAdd(P.cnt + mb, 1u); }
      }
    };
    for (int s = 0; s < P.T; ++s) {
      const int par = s & 1;
      for (int l = 0; l < L; ++l) {
        int slot_i = 0;
        for (int t = cta; t < n_lt; t += G, ++slot_i) {
          const int mb = t / NT, nt = t - mb * NT;
          const int row = mb * WD_BM + r;
          const bool valid = row < P.B;
          const uint32_t a = it & 1, apar = (it >> 1) & 1;
          ++it;
          const int ub = nt * 32 + half * 16;                  // first of this thread's 16 hidden units
          // ---- operands that depend on neither the token nor the accumulator: requested first
          float4 cv[4], ev[4][4];
          float* cp = P.c + ((size_t)l * P.Bp + (valid ? row : 0)) * H + ub;
          {
            const float* erow = l == 0 ? P.gctx + (size_t)(valid ? row : 0) * 4 * H : P.bsum[l];
#pragma unroll
            for (int i = 0; i < 4; ++i) cv[i] = __ldcg(reinterpret_cast<const float4*>(cp + 4 * i));
#pragma unroll
            for (int gt = 0; gt < 4; ++gt)
#pragma unroll
              for (int i = 0; i < 4; ++i) ev[gt][i] = __ldg(reinterpret_cast<const float4*>(erow + gt * H + ub + 4 * i));
          }
          bool dead = false;
          int tok = P.start_id;
          if (l == 0 && s > 0) {
            // ---- the token chosen at the end of step s - 1: needs every logits tile of the block
            if (et == 0) misc[4] = wait_count(P, mb, (unsigned)s * PS, s) ? 0 : 1;
            epi_bar();
            dead = misc[4] != 0;
            if (et == 0) WD_TS(s, 0);
            if (!dead) {
              tok = token_of(mb, s - 1);
              if (half == 0) account(mb, slot_i, nt == 0, s, tok);
            }
            epi_bar();
            if (!dead) {
              const int m = misc[0] & misc[1] & misc[2] & misc[3];
              if (nt == 0 && et == 0) P.allend[(size_t)mb * P.T + s - 1] = (unsigned char)(m & 1);
              if (sticky && (m & 2) != 0) {
                dead = true;                                   // every row of the block has finished: it stops here
                if (nt == 0 && et == 0) {
                  P.block_steps[mb] = s;
                  __threadfence();
                  *reinterpret_cast<volatile int*>(P.stop_at + mb) = s;
                }
              }
            }
          }
          if (l == 0 && !dead && valid) {
            const float* trow = P.gtok + (size_t)tok * 4 * H + ub;
#pragma unroll
            for (int gt = 0; gt < 4; ++gt)
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(trow + gt * H + 4 * i));
                ev[gt][i].x += v.x; ev[gt][i].y += v.y; ev[gt][i].z += v.z; ev[gt][i].w += v.w;
              }
          }
          mbar_wait(TFULL(a), apar);
          tc_fence_after();
          if (et == 0 && l == 1) WD_TS(s, 3);
          if (l > 0 && sticky) dead = ld_volatile_s32(P.stop_at + mb) <= s;
          uint32_t acc[4][16];
#pragma unroll
          for (int gt = 0; gt < 4; ++gt) tc_ld16_nowait(tmem + lane_addr + a * 128 + gt * 32 + half * 16, acc[gt]);
          tc_wait_ld();
          if (!dead && valid) {
            float hn[16];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              float cc[4] = {cv[i].x, cv[i].y, cv[i].z, cv[i].w};
              const float e0[4] = {ev[0][i].x, ev[0][i].y, ev[0][i].z, ev[0][i].w}, e1[4] = {ev[1][i].x, ev[1][i].y, ev[1][i].z, ev[1][i].w};
              const float e2[4] = {ev[2][i].x, ev[2][i].y, ev[2][i].z, ev[2][i].w}, e3[4] = {ev[3][i].x, ev[3][i].y, ev[3][i].z, ev[3][i].w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int k = 4 * i + e;
                // gate pre-activations: accumulator + (bias | Gctx row + Gtok row); MUFU.TANH activations
                const float xi = __uint_as_float(acc[0][k]) + e0[e], xf = __uint_as_float(acc[1][k]) + e1[e];
                const float xg = __uint_as_float(acc[2][k]) + e2[e], xo = __uint_as_float(acc[3][k]) + e3[e];
                const float ig = fmaf(tanh_fast(0.5f * xi), 0.5f, 0.5f), fg = fmaf(tanh_fast(0.5f * xf), 0.5f, 0.5f);
                const float gg = tanh_fast(xg), og = fmaf(tanh_fast(0.5f * xo), 0.5f, 0.5f);
                const float cn = fmaf(fg, cc[e], ig * gg);
                cc[e] = cn;
                hn[k] = og * tanh_fast(cn);
              }
              __stcg(reinterpret_cast<float4*>(cp + 4 * i), make_float4(cc[0], cc[1], cc[2], cc[3]));
            }
            uint32_t pk[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              __nv_bfloat162 h2 = __floats2bfloat162_rn(hn[2 * i], hn[2 * i + 1]);
              pk[i] = *reinterpret_cast<uint32_t*>(&h2);
            }
            uint4* hb = reinterpret_cast<uint4*>(P.hbuf + (((size_t)par * L + l) * P.Bp + row) * H + ub);
            __stcg(hb, make_uint4(pk[0], pk[1], pk[2], pk[3]));
            __stcg(hb + 1, make_uint4(pk[4], pk[5], pk[6], pk[7]));
          }
          publish(a, mb, dead);
          if (et == 0 && l == 0) WD_TS(s, 1);
          if (et == 0 && l == 1) WD_TS(s, 4);
        }
      }
      for (int t = cta; t < n_vt; t += G) {
        const int mb = t / NV, nv = t - mb * NV;
        const uint32_t a = it & 1, apar = (it >> 1) & 1;
        ++it;
        float bv[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int v = nv * WD_LN + half * 16 + i;
          bv[i] = v < P.V ? __ldg(P.out_b + v) : 0.f;
        }
        mbar_wait(TFULL(a), apar);
        tc_fence_after();
        if (et == 0) WD_TS(s, 6);
        const bool dead = sticky && ld_volatile_s32(P.stop_at + mb) <= s;
        uint32_t acc[16];
        tc_ld16_nowait(tmem + lane_addr + a * 128 + half * 16, acc);
        tc_wait_ld();
        if (!dead && MODE == 1) {
          // sampling: the raw logits (+ bias) of this thread's 16 columns, whole rows for the selection phase
          float4* dst = reinterpret_cast<float4*>(P.logits + ((size_t)mb * WD_BM + r) * P.VP + nv * WD_LN + half * 16);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            __stcg(dst + i, make_float4(__uint_as_float(acc[4 * i]) + bv[4 * i], __uint_as_float(acc[4 * i + 1]) + bv[4 * i + 1],
                                        __uint_as_float(acc[4 * i + 2]) + bv[4 * i + 2], __uint_as_float(acc[4 * i + 3]) + bv[4 * i + 3]));
        } else if (!dead) {
          float best = 0.f; int bi = -1;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int v = nv * WD_LN + half * 16 + i;
            if (v < P.V) {
              float x = __uint_as_float(acc[i]) + bv[i];
              if (P.temperature != 1.0f) x = x / P.temperature;          // seq2seq.py:213-214
              if (bi < 0 || x > best) { best = x; bi = v; }              // ascending v, strict >: first maximum
            }
          }
          unsigned long long* slot_p = P.best + ((size_t)par * P.MB + mb) * 128 + r;
          if (bi >= 0) {
            // order-preserving key (-0 -> +0: equal under torch's compare); ties between tiles: the lower index wins
            uint32_t u = __float_as_uint(best + 0.0f);
            u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
            atomicMax(slot_p, ((unsigned long long)u << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)bi));
          }
          // the other parity was read by the layer-0 epilogues of this step (all finished: this tile's flag) and is
          // written again by the logits tiles of step s + 1, which start after this tile has been published
          if (nv == 0 && half == 0 && s > 0) __stcg(P.best + ((size_t)(par ^ 1) * P.MB + mb) * 128 + r, 0ull);
        }
        publish(a, mb, dead);
        if (et == 0) WD_TS(s, 7);
      }
      if (MODE == 1) {
        // ---- selection: the 128 rows of a block are spread over its NT layer-tile owners, one epilogue warp per row
        const int rp = (WD_BM + NT - 1) / NT;
        for (int t = cta; t < n_lt; t += G) {
          const int mb = t / NT, nt = t - mb * NT;
          if (et == 0) misc[4] = wait_count(P, mb, (unsigned)s * PS + (unsigned)(L * NT + NV), s) ? 0 : 1;
          epi_bar();
          const bool dead = misc[4] != 0;
          if (!dead) {
            const int r_end = min(WD_BM, (nt + 1) * rp);
            for (int rr = nt * rp + (warp - 2); rr < r_end; rr += 8) {
              const int row = mb * WD_BM + rr;
              if (row >= P.B) continue;
              const float* x = P.logits + (size_t)row * P.VP;
              float lg[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) lg[i] = (16 * lane + i) < P.V ? __ldcg(x + 16 * lane + i) : 0.f;
              float u = 0.f;
              if (P.do_sample) u = P.uniforms ? P.uniforms[(size_t)s * P.B + row] : philox_uniform(P.seed, P.offset + (uint64_t)s * P.B + row);
              const int chosen = warp_sample_select(lg, P.V, lane, P.temperature, P.top_k, P.top_p, P.do_sample, u,
                                                    P.probs_trace ? P.probs_trace + ((size_t)s * P.B + row) * P.V : nullptr);
              if (lane == 0) __stcg(P.tokbuf + (size_t)par * P.Bp + row, chosen);
            }
          }
          epi_bar();
          if (et == 0 && !dead) { __threadfence(); atomicAdd(P.cnt + mb, 1u); }
        }
      }
    }
    // ---- the token of the last step: appended by the owner tile of every block that is still running
    {
      int slot_i = 0;
      for (int t = cta; t < n_lt; t += G, ++slot_i) {
        const int mb = t / NT, nt = t - mb * NT;
        if (nt != 0) continue;
        if (et == 0) misc[4] = wait_count(P, mb, (unsigned)P.T * PS, P.T) ? 0 : 1;
        epi_bar();
        const bool dead = misc[4] != 0;
        if (!dead && half == 0) account(mb, slot_i, true, P.T, token_of(mb, P.T - 1));
        epi_bar();
        if (!dead && et == 0) {
          P.allend[(size_t)mb * P.T + P.T - 1] = (unsigned char)(misc[0] & misc[1] & misc[2] & misc[3] & 1);
          P.block_steps[mb] = P.T;
        }
        epi_bar();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<256>(tmem);
}

__global__ void wide_init_kernel(int64_t* tokens, int T1, int B, int start_id, int* first_end, unsigned* cnt, int* stop_at,
                                 int* block_steps, int* abort_flag, int MB) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (size_t)B * T1) tokens[i] = (i % T1 == 0) ? start_id : -1;
  if (i < (size_t)B) first_end[i] = -1;
  if (i < (size_t)MB) { cnt[i] = 0u; stop_at[i] = INT_MAX; block_steps[i] = 0; }
  if (i == 0) *abort_flag = 0;
}

__global__ void wide_finalize_kernel(const int* first_end, const unsigned char* allend, const int* block_steps, int MB, int B,
                                     int T, int stop_rule, int32_t* lengths, int32_t* steps_out, const int* abort_flag) {
  if (abort_flag != nullptr && *abort_flag != 0) {
    // a spin of the loop ran into its bound (a CTA that never became resident, a lost update): the tokens are garbage.
    // steps_run = -1 is what the host side turns into an exception at its one synchronisation point.
    if (threadIdx.x == 0 && steps_out) *steps_out = -1;
    if (lengths) for (int i = threadIdx.x; i < B; i += blockDim.x) lengths[i] = -1;
    return;
  }
  // one block: steps_run = first step at which every 128-sequence block reported "all rows emitted END"
  // (ALL_END_SAME_STEP, seq2seq.py:220), or the last block to finish (sticky rule, predictor.py:343-347); then lengths
  __shared__ int steps_sh;
  if (threadIdx.x == 0) steps_sh = T;
  __syncthreads();
  if (stop_rule == I2L_STOP_ALL_END_SAME_STEP) {
    for (int s = threadIdx.x; s < T; s += blockDim.x) {
      bool all = true;
      for (int b = 0; b < MB && all; ++b) all = allend[(size_t)b * T + s] != 0;
      if (all) atomicMin(&steps_sh, s + 1);
    }
  } else if (stop_rule == I2L_STOP_ALL_FINISHED_STICKY) {
    if (threadIdx.x == 0) {
      int steps = 0;
      for (int b = 0; b < MB; ++b) steps = max(steps, block_steps[b]);
      steps_sh = steps;
    }
  }
  __syncthreads();
  const int steps = steps_sh;
  if (threadIdx.x == 0 && steps_out) *steps_out = steps;
  if (lengths)
    for (int i = threadIdx.x; i < B; i += blockDim.x) {
      const int fe = first_end[i];
      lengths[i] = (fe >= 0 && fe <= steps) ? fe : steps + 1;
    }
}

struct WWs {
  float* gctx; __nv_bfloat16* encb; __nv_bfloat16* hbuf; float* c; unsigned long long* best; unsigned* cnt; int* stop_at;
  int* block_steps; int* abort_flag; int* first_end; unsigned char* allend; float* logits; int* tokbuf; size_t bytes;
};
WWs wcarve(const i2l_dec_desc& d, int rows, int T, void* ws) {
  Arena a(ws, (size_t)-1);
  WWs w{};
  const size_t H = d.hidden_dim, L = d.lstm_layers, E = d.embedding_dim;
  const size_t MB = cdiv(rows, WD_BM), Bp = MB * WD_BM;
  w.gctx = a.take<float>(Bp * 4 * H);
  w.encb = a.take<__nv_bfloat16>(Bp * E);
  w.hbuf = a.take<__nv_bfloat16>(2 * L * Bp * H);
  w.c = a.take<float>(L * Bp * H);
  w.best = a.take<unsigned long long>(2 * MB * 128);
  w.cnt = a.take<unsigned>(MB);
  w.stop_at = a.take<int>(MB);
  w.block_steps = a.take<int>(MB);
  w.abort_flag = a.take<int>(1);
  w.first_end = a.take<int>(rows);
  w.allend = a.take<unsigned char>(MB * (size_t)(T > 0 ? T : 1));
  w.logits = a.take<float>(Bp * (size_t)cdiv(d.vocab_size, WD_LN) * WD_LN);   // sampling mode only
  w.tokbuf = a.take<int>(2 * Bp);
  w.bytes = align_up(a.off, 256);
  return w;
}

}  // namespace

// Shapes of the persistent whole-GPU loop: bf16, the cell-fused weight order (H % 32), K blocks of 64, L <= 4
bool wide_supported(const i2l_dec_desc& d) {
  return d.precision == I2L_BF16 && general_bf16_supported(d) && (d.hidden_dim % 64) == 0 && d.hidden_dim >= 128 &&
         (d.embedding_dim % 8) == 0 && d.lstm_layers >= 1 && d.lstm_layers <= WD_MAXL && d.vocab_size >= 1 && !persistent_supported(d);
}
size_t wide_workspace_bytes(const i2l_dec_desc& d, int rows, int max_length) { return wcarve(d, rows, max_length, nullptr).bytes; }

// grid of the loop for `rows` sequences, or 0 when the tile count exceeds what the static ownership covers
static int wide_grid(const i2l_dec_desc& d, int rows) {
  const int MB = cdiv(rows, WD_BM), NT = d.hidden_dim / 32;
  const int n_lt = MB * NT, sms = num_sms();
  const int per = cdiv(n_lt, sms);
  if (per > WD_MAXT) return 0;
  return cdiv(n_lt, per);
}
bool wide_batch_supported(const i2l_dec_desc& d, int rows) { return wide_supported(d) && rows > 0 && wide_grid(d, rows) > 0; }

int wide_greedy(const i2l_dec_desc& d, const void* packed, const PackedDec& lay, const float* enc, int batch, int start_id,
                int end_id, int max_length, float temperature, int stop_rule, int64_t* tokens, int32_t* lengths,
                int32_t* steps_run, void* ws, size_t ws_bytes, cudaStream_t s, const PersistentSampleArgs* sample) {
  I2L_REQUIRE(sample == nullptr || d.vocab_size <= 512, "wide_greedy: the in-kernel selection covers V <= 512");
  I2L_REQUIRE(start_id >= 0 && start_id < d.vocab_size, "decode loop: start token %d outside [0, %d) (nn.Embedding raises IndexError)",
              start_id, d.vocab_size);
  I2L_REQUIRE(lay.g16c != 0, "wide_greedy: packed weights lack the cell-ordered bf16 section");
  WWs w = wcarve(d, batch, max_length, ws);
  if (ws_bytes < w.bytes) { set_error("wide_greedy: workspace too small (%zu < %zu)", ws_bytes, w.bytes); return I2L_ERR_WORKSPACE; }
  const int H = d.hidden_dim, L = d.lstm_layers, V = d.vocab_size, T1 = max_length + 1;
  const int MB = cdiv(batch, WD_BM), Bp = MB * WD_BM;
  const char* pb = reinterpret_cast<const char*>(packed);
  const float* pk = reinterpret_cast<const float*>(packed);
  {
    const size_t tot = (size_t)batch * T1;
    wide_init_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(tokens, T1, batch, start_id, w.first_end, w.cnt, w.stop_at,
                                                                   w.block_steps, w.abort_flag, MB);
    I2L_LAUNCH_OK();
  }
  I2L_CUDA_OK(cudaMemsetAsync(w.hbuf, 0, (size_t)2 * L * Bp * H * 2, s));
  I2L_CUDA_OK(cudaMemsetAsync(w.c, 0, (size_t)L * Bp * H * 4, s));
  I2L_CUDA_OK(cudaMemsetAsync(w.best, 0, (size_t)2 * MB * 128 * 8, s));
  {
    KernelTimer kt("dec.gctx_gemm", s);
    I2L_TRY(make_gctx_bf16(d, packed, lay, enc, batch, w.gctx, w.encb, s));
  }
  if (max_length > 0) {
    static thread_local WideParams P;                          // 1.5 KB of tensor maps: not on the stack
    P = WideParams{};
    I2L_TRY(gemm_bf16_operand_map(&P.tmH, w.hbuf, 2 * L * Bp, H, H, WD_BM));
    for (int l = 0; l < L; ++l) {
      I2L_TRY(gemm_bf16_operand_map(&P.tmW[2 * l], pb + lay.g16c_w_hh[l], 4 * H, H, H, WD_BN));
      if (l > 0) I2L_TRY(gemm_bf16_operand_map(&P.tmW[2 * l + 1], pb + lay.g16c_w_ih[l], 4 * H, H, H, WD_BN));
      P.bsum[l] = pk + lay.bsum[l];
    }
    I2L_TRY(gemm_bf16_operand_map(&P.tmO, pb + lay.g16_out_w, V, H, H, WD_LN));
    P.gctx = w.gctx; P.gtok = pk + lay.gtok; P.out_b = pk + lay.out_b;
    P.hbuf = w.hbuf; P.c = w.c; P.best = w.best; P.cnt = w.cnt; P.stop_at = w.stop_at; P.abort_flag = w.abort_flag;
    P.tokens = tokens; P.first_end = w.first_end; P.allend = w.allend; P.block_steps = w.block_steps;
    P.B = batch; P.Bp = Bp; P.H = H; P.L = L; P.V = V; P.T = max_length; P.NT = H / 32; P.NV = cdiv(V, WD_LN); P.MB = MB;
    P.KB = H / WD_BK;
    P.start_id = start_id; P.end_id = end_id; P.stop_rule = stop_rule; P.temperature = temperature;
    const int grid = wide_grid(d, batch);
    I2L_REQUIRE(grid > 0, "wide_greedy: %d sequences need more tiles per CTA than the loop covers", batch);
    P.logits = w.logits; P.tokbuf = w.tokbuf; P.VP = P.NV * WD_LN;
    if (sample) {
      P.top_k = sample->top_k; P.top_p = sample->top_p; P.do_sample = sample->do_sample;
      P.seed = sample->seed; P.offset = sample->offset; P.uniforms = sample->uniforms; P.probs_trace = sample->probs_trace;
    }
    const void* kern = sample ? (const void*)wide_loop_kernel<1> : (const void*)wide_loop_kernel<0>;
    I2L_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, WD_SMEM));
    int per_sm = 0;
    if (sample) I2L_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, wide_loop_kernel<1>, WD_THREADS, WD_SMEM));
    else I2L_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, wide_loop_kernel<0>, WD_THREADS, WD_SMEM));
    I2L_REQUIRE(per_sm >= 1 && grid <= per_sm * num_sms(), "wide_greedy: the cooperative grid (%d CTAs) is not co-resident", grid);
#ifdef I2L_DIAG
    static long long* ts_dev = nullptr;                        // I2L_WIDE_TS=1: per-phase clock stamps of CTA 0, printed after a sync
    const bool want_ts = getenv("I2L_WIDE_TS") != nullptr && max_length <= 256;
    if (want_ts && ts_dev == nullptr) I2L_CUDA_OK(cudaMalloc(&ts_dev, 256 * 8 * sizeof(long long)));
    if (want_ts) { I2L_CUDA_OK(cudaMemsetAsync(ts_dev, 0, 256 * 8 * sizeof(long long), s)); P.dbg_ts = ts_dev; }
#endif
    void* args[] = {(void*)&P};
    {
      KernelTimer kt(sample ? "dec.sample_wide" : "dec.greedy_wide", s);
      I2L_CUDA_OK(cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(WD_THREADS), args, WD_SMEM, s));
      count_launch();
    }
#ifdef I2L_DIAG
    if (want_ts) {
      static long long ts[256 * 8];
      I2L_CUDA_OK(cudaStreamSynchronize(s));
      I2L_CUDA_OK(cudaMemcpy(ts, ts_dev, sizeof(ts), cudaMemcpyDeviceToHost));
      const int s0 = max_length / 2;
      for (int st = s0; st < s0 + 3 && st + 1 < max_length; ++st) {
        const long long* t = ts + st * 8;
        fprintf(stderr, "wide step %d (cycles since tokens known): L0 epi done %lld | L1 flag %lld  acc ready %lld  epi done %lld | "
                        "LG flag %lld  acc ready %lld  epi done %lld | next tokens known %lld\n", st, t[1] - t[0], t[2] - t[0],
                t[3] - t[0], t[4] - t[0], t[5] - t[0], t[6] - t[0], t[7] - t[0], ts[(st + 1) * 8] - t[0]);
      }
    }
#endif
  } else {
    I2L_CUDA_OK(cudaMemsetAsync(w.block_steps, 0, (size_t)MB * 4, s));
  }
  wide_finalize_kernel<<<1, 256, 0, s>>>(w.first_end, w.allend, w.block_steps, MB, batch, max_length, stop_rule, lengths, steps_run,
                                         max_length > 0 ? w.abort_flag : nullptr);
  I2L_LAUNCH_OK();
  return I2L_OK;
}

// host-side check after a synchronisation (tests / debugging): 1 when a spin of the last loop ran into its bound
int wide_aborted(const void* ws, const i2l_dec_desc& d, int rows, int max_length, int* out) {
  WWs w = wcarve(d, rows, max_length, const_cast<void*>(ws));
  I2L_CUDA_OK(cudaMemcpy(out, w.abort_flag, sizeof(int), cudaMemcpyDeviceToHost));
  return I2L_OK;
}

}  // namespace i2l
