// Warp-level token selection of the sampling loop (Predictor.predict_batch, training/predictor.py:295-335) for
// V <= 512, shared by the stream-ordered select kernel (decode_general.cu) and the persistent sampling kernel
// (decode_persistent.cu): one warp owns one row, vocabulary index 16 * lane + i in registers.
#pragma once
#include "decode_kernels.cuh"

namespace i2l {
namespace {

__device__ __forceinline__ void warp_argmax(float& v, int& i) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, v, o);
    int oi = __shfl_xor_sync(0xffffffffu, i, o);
    if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
  }
}


__device__ __forceinline__ uint32_t mulhilo(uint32_t a, uint32_t b, uint32_t* hi) {
  uint64_t p = (uint64_t)a * b;
  *hi = (uint32_t)(p >> 32);
  return (uint32_t)p;
}
__device__ float philox_uniform(uint64_t seed, uint64_t ctr) {
  uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = 0, c3 = 0;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t h0, h1;
    uint32_t l0 = mulhilo(0xD2511F53u, c0, &h0), l1 = mulhilo(0xCD9E8D57u, c2, &h1);
    uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
    c0 = n0; c1 = l1; c2 = n2; c3 = l0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return (float)(c0 >> 8) * (1.0f / 16777216.0f);
}


// Same arithmetic as sample_select_kernel, with the row held in registers (vocab index 16*lane + i) and ONE
// warp-level bitonic sort instead of two block-wide sorts in shared memory: top-k masking and renormalisation
// keep the order of the surviving entries, so the order found once is also the order of the top-p pass
// (predictor.py:311-317 sorts again).  The kept set is a prefix of the sorted order; it is carried back to index
// order as a threshold VALUE plus a rank among equal values, not as a scatter.
__device__ __forceinline__ double warp_excl_scan(double v, int lane, double* total) {
  double inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { double n = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += n; }
  *total = __shfl_sync(0xffffffffu, inc, 31);
  return inc - v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// x / d for a divisor shared by the whole row: one IEEE reciprocal per divisor, then a multiply and one fused
// residual correction per element (q = x r; q += (x - q d) r): correctly rounded except in rare half-way cases, 3
// instructions and no slow-path branch where the compiler's IEEE division costs ~12 and a convergence barrier.  The
// divisors here (a softmax denominator >= 1, kept-probability sums >= 1/V) are far from the under/overflow range.
struct RowDiv { float d, r; };
__device__ __forceinline__ RowDiv row_div(float d) { return RowDiv{d, 1.0f / d}; }
__device__ __forceinline__ float div_by(float x, const RowDiv& u) {
  const float q = x * u.r;
  return fmaf(fmaf(-q, u.d, x), u.r, q);
}
// sorted element at (warp-uniform) position pos of the 512-entry array held as key[16] per lane
__device__ __forceinline__ uint32_t sorted_at32(const uint32_t (&key)[16], int pos) {
  uint32_t sel = key[0];
#pragma unroll
  for (int i = 1; i < 16; ++i) if ((pos & 15) == i) sel = key[i];
  return __shfl_sync(0xffffffffu, sel, pos >> 4);
}


// logit[i] = raw logit of vocabulary entry 16 * lane + i (ignored beyond V); u = the row's uniform draw of this step;
// probs_out: optional V floats (filtered distribution).
// Returns the chosen token (warp-uniform).
__device__ __forceinline__ int warp_sample_select(const float (&logit)[16], int V, int lane, float temperature, int top_k,
                                                  float top_p, int do_sample, float u, float* probs_out) {
  float po[16];                                             // probabilities in index order
  const RowDiv dt = row_div(temperature);
  // softmax(logits / T)                                               predictor.py:295-297
  // (exp through ex2.approx: relative error <= ~5e-6 for the arguments of a softmax, far inside the 1e-3 the
  // distributions are compared at; the libm-accurate expf costs 17 instructions per entry)
  float lm = -INFINITY;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int v = 16 * lane + i;
    float val = -INFINITY;
    if (v < V) { val = logit[i]; if (temperature != 1.0f) val = div_by(val, dt); }
    po[i] = val;
    lm = fmaxf(lm, val);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) lm = fmaxf(lm, __shfl_xor_sync(0xffffffffu, lm, o));
  float ls = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) { const float e = (16 * lane + i) < V ? __expf(po[i] - lm) : 0.f; po[i] = e; ls += e; }
  const float s = warp_sum(ls);
  const RowDiv ds = row_div(s);
#pragma unroll
  for (int i = 0; i < 16; ++i) po[i] = div_by(po[i], ds);

  if (top_k > 0 || top_p > 0.0f) {
    // ---- ONE descending sort of the 512 probabilities as 32-bit keys (p >= 0: float order == unsigned order),
    // position t = 16 lane + i.  No index payload: the k-th value, the sorted cumulative sums and the size R of the
    // kept prefix depend on the VALUES only; which of several equal values at the cut are kept (torch: lowest index
    // first) is settled afterwards in index order by a rank among equals.  The network runs on keys XOR-ed with a
    // per-lane mask so that every compare-exchange is "max to the lower position": 2 instructions per exchange.
    uint32_t key[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) key[i] = (16 * lane + i) < V ? __float_as_uint(po[i]) : 0u;
    uint32_t mask = 0u;                                          // 0: ascending-complement not applied
#pragma unroll
    for (int lk = 1; lk <= 9; ++lk) {
      const int k = 1 << lk;
      if (k >= 16) {                                             // direction of this lane's 16 positions: bit k of t = 16 lane + i
        const uint32_t nm = ((16 * lane) & k) ? 0xFFFFFFFFu : 0u;
        const uint32_t flip = nm ^ mask;
#pragma unroll
        for (int i = 0; i < 16; ++i) key[i] ^= flip;
        mask = nm;
      }
#pragma unroll
      for (int lj = 8; lj >= 0; --lj) {
        if (lj >= lk) continue;
        const int j = 1 << lj;
        if (j >= 16) {
          const int lx = j >> 4;
          const bool lower = (lane & lx) == 0;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const uint32_t o = __shfl_xor_sync(0xffffffffu, key[i], lx);
            key[i] = lower ? max(key[i], o) : min(key[i], o);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            if (i & j) continue;
            const uint32_t a = key[i], b = key[i | j];
            const bool desc = k >= 16 ? true : ((i & k) == 0);   // k < 16: static direction; above: folded into the mask
            key[i] = desc ? max(a, b) : min(a, b);
            key[i | j] = desc ? min(a, b) : max(a, b);
          }
        }
      }
    }
    // k = 512: bit 9 of t is 0 for every position => mask == 0, keys are plain and sorted descending over t
    int R = V;                                                   // kept entries = sorted positions [0, R)
    float s2 = 1.f, s3 = 1.f, kth = 0.f;
    bool renorm2 = false, renorm3 = false;
    RowDiv d2{1.f, 1.f}, d3{1.f, 1.f};
    if (top_k > 0) {                                            // predictor.py:299-309 (ties with the k-th value are kept)
      const int k = min(top_k, V);
      kth = __uint_as_float(sorted_at32(key, k - 1));
      float l2 = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) if (po[i] >= kth) l2 += po[i];
      s2 = warp_sum(l2);
      renorm2 = s2 > 0.f;
      if (renorm2) d2 = row_div(s2);
    }
    uint32_t vR = 0u;                                           // value at sorted position R - 1
    int m_eq = 0x7fffffff;                                      // how many entries equal to vR lie inside the kept prefix
    if (top_p > 0.0f) {                                         // predictor.py:311-327
      // cumulative sum over the sorted, top-k-masked and renormalised probabilities (fp64, ATen CPU cumsum);
      // sorted position t is removed iff t >= 1 and float(cum[t-1]) > top_p
      double c[16], acc = 0.0;
      float sp2[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float v = __uint_as_float(key[i]);
        if (top_k > 0) { v = v >= kth ? v : 0.f; if (renorm2) v = div_by(v, d2); }
        sp2[i] = v;
        acc += (double)v;
        c[i] = acc;
      }
      double tot;
      const double base = warp_excl_scan(acc, lane, &tot);
      int keep = 0; float l3 = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const double prev = i == 0 ? base : base + c[i - 1];
        const bool rem = (lane > 0 || i > 0) && (float)prev > top_p;
        if (!rem) { ++keep; l3 += sp2[i]; }
      }
      R = __reduce_add_sync(0xffffffffu, keep);
      s3 = warp_sum(l3);
      renorm3 = s3 > 0.f;
      if (renorm3) d3 = row_div(s3);
      vR = sorted_at32(key, max(R, 1) - 1);
      int gt = 0;
#pragma unroll
      for (int i = 0; i < 16; ++i) gt += key[i] > vR ? 1 : 0;
      m_eq = max(R, 1) - __reduce_add_sync(0xffffffffu, gt);
    }
    // ---- back to index order: entry v survives top-p iff its value exceeds vR, or equals it and is among the
    // first m_eq such entries in index order (= the stable descending order torch.sort / the oracle produce)
    int eq_before = 0;
    if (top_p > 0.0f) {
      int eq = 0;
#pragma unroll
      for (int i = 0; i < 16; ++i) eq += ((16 * lane + i) < V ? __float_as_uint(po[i]) : 0u) == vR ? 1 : 0;
      int incl = eq;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int n = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += n; }
      eq_before = incl - eq;
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      float v = po[i];
      const uint32_t bits = (16 * lane + i) < V ? __float_as_uint(po[i]) : 0u;
      if (top_k > 0) { v = v >= kth ? v : 0.f; if (renorm2) v = div_by(v, d2); }
      if (top_p > 0.0f) {
        bool keep = bits > vR;
        if (bits == vR) { keep = eq_before < m_eq; ++eq_before; }
        if (!keep) v = 0.f;
        if (renorm3) v = div_by(v, d3);
      }
      po[i] = v;
    }
  }
  if (probs_out) {
#pragma unroll
    for (int i = 0; i < 16; ++i) if (16 * lane + i < V) probs_out[16 * lane + i] = po[i];
  }
  int chosen;
  if (do_sample) {                                            // predictor.py:330-331 (restated inverse-CDF draw)
    double c[16], acc = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) { acc += (double)po[i]; c[i] = acc; }
    double tot;
    const double base = warp_excl_scan(acc, lane, &tot);
    const double tgt = (double)u * tot;
    int best = 0x7fffffff, lastpos = -1;
#pragma unroll
    for (int i = 15; i >= 0; --i) {
      if (base + c[i] > tgt && 16 * lane + i < V) best = 16 * lane + i;
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) if (po[i] > 0.f) lastpos = 16 * lane + i;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
      lastpos = max(lastpos, __shfl_xor_sync(0xffffffffu, lastpos, o));
    }
    chosen = best != 0x7fffffff ? best : max(lastpos, 0);
  } else {                                                    // predictor.py:333-335 argmax(probs)
    float bv = -INFINITY; int bi = 0x7fffffff;
#pragma unroll
    for (int i = 0; i < 16; ++i) if (16 * lane + i < V && po[i] > bv) { bv = po[i]; bi = 16 * lane + i; }
    warp_argmax(bv, bi);
    chosen = bi == 0x7fffffff ? 0 : bi;
  }
  return chosen;
}

}  // namespace
}  // namespace i2l
