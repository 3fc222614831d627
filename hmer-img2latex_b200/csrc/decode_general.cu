// General decode path: any (V,E,H,L), fp32 arithmetic, one stream-ordered sequence of
// launches per step, every piece of loop bookkeeping (token append, EOS masks, loop exit,
// beam merge / back-pointers / state reorder, sampling filter + draw) on the device.
// The C-ABI entry points of the decoder live here; for the headline shape in bf16 they
// route the greedy loop to the persistent cluster kernel (decode_persistent.cu).
#include "decode_kernels.cuh"
#include "sample_select.cuh"
#include "gemm_bf16.cuh"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>
#include <vector>

namespace i2l {

// ------------------------------------------------------------------ packed layout
PackedDec dec_layout(const i2l_dec_desc& d) {
  PackedDec L{};
  size_t o = 0;
  auto take = [&](size_t n) { size_t r = o; o += (n + 63) / 64 * 64; return r; };
  const size_t V = d.vocab_size, E = d.embedding_dim, H = d.hidden_dim;
  L.emb = take(V * E);
  L.gtok = take(V * 4 * H);
  L.w_ih0 = take(4 * H * 2 * E);
  for (int l = 0; l < d.lstm_layers; ++l) {
    L.bsum[l] = take(4 * H);
    L.w_hh[l] = take(4 * H * H);
    L.w_ih[l] = l == 0 ? L.w_ih0 : take(4 * H * H);
  }
  L.out_w = take(V * H);
  L.out_b = take(V);
  L.end_f32 = o;
  size_t bytes = o * sizeof(float);
  L.bf16_section = 0;
  if (d.precision == I2L_BF16 && persistent_supported(d)) {
    bytes = align_up(bytes, 1024);
    L.bf16_section = bytes;
    bytes += persistent_packed_bytes(d);
  }
  L.g16 = 0;
  if (general_bf16_supported(d)) {
    bytes = align_up(bytes, 1024);
    L.g16 = bytes;
    auto take16 = [&](size_t n) { size_t r = bytes; bytes = align_up(bytes + n * 2, 256); return r; };
    for (int l = 0; l < d.lstm_layers; ++l) {
      L.g16_w_hh[l] = take16(4 * H * H);
      L.g16_w_ih[l] = l == 0 ? 0 : take16(4 * H * H);
    }
    L.g16_out_w = take16(V * H);
    L.g16_w_ctx = take16(4 * H * E);
    L.g16c = 0;
    if ((H % 32) == 0) {
      L.g16c = bytes;
      for (int l = 0; l < d.lstm_layers; ++l) {
        L.g16c_w_hh[l] = take16(4 * H * H);
        L.g16c_w_ih[l] = l == 0 ? 0 : take16(4 * H * H);
      }
    }
  }
  L.total_bytes = bytes;
  return L;
}

bool general_bf16_supported(const i2l_dec_desc& d) {
  return d.precision == I2L_BF16 && (d.hidden_dim % 8) == 0 && (d.embedding_dim % 8) == 0;
}

namespace {

__global__ void f32_to_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = __float2bfloat16(src[i]);
}
// dst (rows, cols) contiguous <- src (rows, cols) with leading dimension ld
__global__ void f32_to_bf16_ld_kernel(const float* __restrict__ src, int ld, __nv_bfloat16* __restrict__ dst, int cols, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = __float2bfloat16(src[(i / cols) * ld + (i % cols)]);
}

// (B,T) -> (T,B) with the nn.Embedding range check: ids outside [0,V) set *bad and are read as 0
__global__ void transpose_tokens_kernel(const int64_t* __restrict__ src, int64_t* __restrict__ dst, int B, int T, int V,
                                        int* bad) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * T) return;
  int t = (int)(i / B), b = (int)(i % B);
  int64_t v = src[(size_t)b * T + t];
  if (v < 0 || v >= V) { if (bad) atomicExch(bad, 1); v = 0; }
  dst[i] = v;
}

// (4H, K) fp32 gate weights -> bf16 with the rows of every 128-row tile ordered [gate][32 units]:
// dst row 128 t + 32 g + ul  <-  src row g H + 32 t + ul
__global__ void pack_cell_rows_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int H, int K) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)4 * H * K) return;
  const int k = (int)(i % K), r = (int)(i / K);
  const int t = r >> 7, gt = (r >> 5) & 3, ul = r & 31;
  dst[i] = __float2bfloat16(src[((size_t)gt * H + 32 * t + ul) * K + k]);
}

__global__ void add_vec_kernel(const float* a, const float* b, float* o, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) o[i] = a[i] + b[i];
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

__global__ void lstm_cell_kernel(const float* __restrict__ gates, float* __restrict__ h, float* __restrict__ c,
                                 int rows, int H, const int* skip, __nv_bfloat16* __restrict__ hb,
                                 float* __restrict__ h_seq, __nv_bfloat16* __restrict__ hb_seq, size_t seq_ld) {
  if (skip != nullptr && *skip != 0) return;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)rows * H) return;
  int r = (int)(i / H), j = (int)(i % H);
  const float* g = gates + (size_t)r * 4 * H;
  float ig = sigmoidf_(g[j]), fg = sigmoidf_(g[H + j]), gg = tanhf(g[2 * H + j]), og = sigmoidf_(g[3 * H + j]);
  float cn = fg * c[i] + ig * gg;
  c[i] = cn;
  const float hn = og * tanhf(cn);
  h[i] = hn;
  if (hb != nullptr) hb[i] = __float2bfloat16(hn);
  // teacher-forced forward: the top layer's h of this time step, row r at r * seq_ld
  if (h_seq != nullptr) h_seq[(size_t)r * seq_ld + j] = hn;
  if (hb_seq != nullptr) hb_seq[(size_t)r * seq_ld + j] = __float2bfloat16(hn);
}

__global__ void loop_init_kernel(int64_t* tokens, int T1, int rows, int start_id, int64_t* tok_cur,
                                 int* first_end, LoopState* st) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) { st->done = 0; st->steps_run = 0; st->finished_count = 0; st->ticket = 0; st->end_count = 0; }
  if (i < (size_t)rows * T1) tokens[i] = (i % T1 == 0) ? start_id : -1;
  if (i < (size_t)rows) { tok_cur[i] = start_id; first_end[i] = -1; }
}

// Token append + EOS bookkeeping + loop-exit test shared by the select kernels.
// Called by one thread per row with the row's chosen token; `block_new_end` is a
// shared counter zeroed by the caller.
__device__ __forceinline__ void commit_token(int row, int tok, int step, int T1, int end_id, int stop_rule,
                                             int64_t* tokens, int64_t* tok_cur, int* first_end,
                                             int* block_counter) {
  tokens[(size_t)row * T1 + step + 1] = tok;
  tok_cur[row] = tok;
  bool is_end = tok == end_id;
  bool newly = false;
  if (is_end && first_end[row] < 0) { first_end[row] = step + 1; newly = true; }
  if (stop_rule == I2L_STOP_ALL_END_SAME_STEP) { if (is_end) atomicAdd(block_counter, 1); }
  else if (stop_rule == I2L_STOP_ALL_FINISHED_STICKY) { if (newly) atomicAdd(block_counter, 1); }
}

__device__ __forceinline__ void block_loop_exit(int block_counter, int rows, int step, int stop_rule,
                                                LoopState* st) {
  // thread 0 of the block, after __syncthreads()
  if (stop_rule == I2L_STOP_ALL_END_SAME_STEP && block_counter) atomicAdd(&st->end_count, block_counter);
  if (stop_rule == I2L_STOP_ALL_FINISHED_STICKY && block_counter) atomicAdd(&st->finished_count, block_counter);
  __threadfence();
  int t = atomicAdd(&st->ticket, 1);
  if (t == (int)gridDim.x - 1) {
    __threadfence();
    if (stop_rule == I2L_STOP_ALL_END_SAME_STEP) {
      if (atomicAdd(&st->end_count, 0) == rows) st->done = 1;
      st->end_count = 0;
    } else if (stop_rule == I2L_STOP_ALL_FINISHED_STICKY) {
      if (atomicAdd(&st->finished_count, 0) == rows) st->done = 1;
    }
    st->steps_run = step + 1;
    st->ticket = 0;
  }
}

// seq2seq.py:213-215 -- (logits / temperature).argmax, first index wins.
__global__ void greedy_select_kernel(const float* __restrict__ logits, int V, int rows, float temperature,
                                     int step, int T1, int end_id, int stop_rule, int64_t* tokens,
                                     int64_t* tok_cur, int* first_end, LoopState* st) {
  if (st->done) return;
  __shared__ int counter;
  if (threadIdx.x == 0) counter = 0;
  __syncthreads();
  int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  int row = blockIdx.x * (blockDim.x / 32) + warp;
  if (row < rows) {
    const float* x = logits + (size_t)row * V;
    float bv = -INFINITY; int bi = 0x7fffffff;
    for (int j = lane; j < V; j += 32) {
      float v = x[j];
      if (temperature != 1.0f) v = v / temperature;
      if (v > bv || (v == bv && j < bi)) { bv = v; bi = j; }
    }
    warp_argmax(bv, bi);
    if (bi == 0x7fffffff) bi = 0;
    if (lane == 0) commit_token(row, bi, step, T1, end_id, stop_rule, tokens, tok_cur, first_end, &counter);
  }
  __syncthreads();
  if (threadIdx.x == 0) block_loop_exit(counter, rows, step, stop_rule, st);
}

__global__ void loop_finalize_kernel(const int* first_end, int rows, int max_length, int32_t* lengths,
                                     int32_t* steps_out, LoopState* st) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int steps = st->steps_run;
  if (i == 0 && steps_out) *steps_out = steps;
  if (i < rows && lengths) {
    int fe = first_end[i];
    lengths[i] = (fe >= 0 && fe <= steps) ? fe : steps + 1;
  }
}

// ------------------------------------------------------------------ sampling (predictor.py:295-335)
__device__ __forceinline__ float block_reduce_max(float v, float* scratch) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if (threadIdx.x % 32 == 0) scratch[threadIdx.x / 32] = v;
  __syncthreads();
  float r = scratch[0];
  for (int w = 1; w < (int)blockDim.x / 32; ++w) r = fmaxf(r, scratch[w]);
  return r;
}
__device__ __forceinline__ float block_reduce_sum(float v, float* scratch) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (threadIdx.x % 32 == 0) scratch[threadIdx.x / 32] = v;
  __syncthreads();
  float r = 0.f;
  for (int w = 0; w < (int)blockDim.x / 32; ++w) r += scratch[w];
  return r;
}

__device__ __forceinline__ bool precedes(float ka, int ia, float kb, int ib) {
  return ka > kb || (ka == kb && ia < ib);
}

// descending stable sort of (key,id) in shared memory, n2 = power of two
__device__ void bitonic_sort_desc(float* key, int* id, int n2) {
  for (int k = 2; k <= n2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n2; i += blockDim.x) {
        int ixj = i ^ j;
        if (ixj > i) {
          bool up = (i & k) == 0;
          float ka = key[i], kb = key[ixj];
          int ia = id[i], ib = id[ixj];
          if (precedes(kb, ib, ka, ia) == up) { key[i] = kb; key[ixj] = ka; id[i] = ib; id[ixj] = ia; }
        }
      }
      __syncthreads();
    }
  }
}

// One block per row.  dynamic smem: p[V] float, key[n2] float, id[n2] int, cdf[V] double
__global__ void sample_select_kernel(const float* __restrict__ logits, int V, int n2, int rows,
                                     float temperature, int top_k, float top_p, int do_sample,
                                     uint64_t seed, uint64_t offset, const float* __restrict__ uniforms,
                                     float* __restrict__ probs_trace, int step, int T1, int end_id,
                                     int stop_rule, int64_t* tokens, int64_t* tok_cur, int* first_end,
                                     LoopState* st) {
  if (st->done) return;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* cdf = reinterpret_cast<double*>(smem_raw);
  float* p = reinterpret_cast<float*>(cdf + V);
  float* key = p + V;
  int* id = reinterpret_cast<int*>(key + n2);
  __shared__ float scratch[32];
  __shared__ int counter;
  __shared__ int chosen;
  const int row = blockIdx.x;
  const int tid = threadIdx.x, nt = blockDim.x;
  if (tid == 0) { counter = 0; chosen = -1; }
  const float* x = logits + (size_t)row * V;

  // softmax(logits / T)                                               predictor.py:295-297
  float lm = -INFINITY;
  for (int j = tid; j < V; j += nt) {
    float v = x[j];
    if (temperature != 1.0f) v = v / temperature;
    p[j] = v;
    lm = fmaxf(lm, v);
  }
  float m = block_reduce_max(lm, scratch);
  float ls = 0.f;
  for (int j = tid; j < V; j += nt) { float e = expf(p[j] - m); p[j] = e; ls += e; }
  float s = block_reduce_sum(ls, scratch);
  for (int j = tid; j < V; j += nt) p[j] = p[j] / s;
  __syncthreads();

  if (top_k > 0) {                                                  // predictor.py:299-309
    int k = min(top_k, V);
    for (int j = tid; j < n2; j += nt) { key[j] = j < V ? p[j] : -1.f; id[j] = j; }
    __syncthreads();
    bitonic_sort_desc(key, id, n2);
    float kth = key[k - 1];
    float l2 = 0.f;
    for (int j = tid; j < V; j += nt) { float v = p[j]; if (v < kth) v = 0.f; p[j] = v; l2 += v; }
    float s2 = block_reduce_sum(l2, scratch);
    if (s2 > 0.f) for (int j = tid; j < V; j += nt) p[j] = p[j] / s2;
    __syncthreads();
  }
  if (top_p > 0.0f) {                                               // predictor.py:311-327
    for (int j = tid; j < n2; j += nt) { key[j] = j < V ? p[j] : -1.f; id[j] = j; }
    __syncthreads();
    bitonic_sort_desc(key, id, n2);
    // cumulative sum over the sorted probabilities, accumulated in double and rounded to
    // float per element (ATen's CPU cumsum uses acc_type<float> = double).
    if (tid < 32) {
      double carry = 0.0;
      for (int base = 0; base < V; base += 32) {
        int j = base + tid;
        double v = j < V ? (double)key[j] : 0.0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { double n = __shfl_up_sync(0xffffffffu, v, o); if (tid >= o) v += n; }
        v += carry;
        if (j < V) cdf[j] = v;
        carry = __shfl_sync(0xffffffffu, v, 31);
      }
    }
    __syncthreads();
    // sorted position i is removed iff cum[i-1] > top_p (i >= 1); position 0 always kept
    for (int i = tid; i < V; i += nt) {
      bool rem = i > 0 && (float)cdf[i - 1] > top_p;
      if (rem) p[id[i]] = 0.f;
    }
    __syncthreads();
    float l3 = 0.f;
    for (int j = tid; j < V; j += nt) l3 += p[j];
    float s3 = block_reduce_sum(l3, scratch);
    if (s3 > 0.f) for (int j = tid; j < V; j += nt) p[j] = p[j] / s3;
    __syncthreads();
  }
  if (probs_trace) {
    float* o = probs_trace + ((size_t)step * rows + row) * V;
    for (int j = tid; j < V; j += nt) o[j] = p[j];
  }
  if (do_sample) {                                                  // predictor.py:330-331 (restated draw)
    if (tid < 32) {
      double carry = 0.0;
      for (int base = 0; base < V; base += 32) {
        int j = base + tid;
        double v = j < V ? (double)p[j] : 0.0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { double n = __shfl_up_sync(0xffffffffu, v, o); if (tid >= o) v += n; }
        v += carry;
        if (j < V) cdf[j] = v;
        carry = __shfl_sync(0xffffffffu, v, 31);
      }
    }
    __syncthreads();
    float u = uniforms ? uniforms[(size_t)step * rows + row]
                       : philox_uniform(seed, offset + (uint64_t)step * rows + row);
    double tgt = (double)u * cdf[V - 1];
    // first j with cdf[j] > tgt (cdf is non-decreasing): block-wide min over candidates
    int best = 0x7fffffff;
    for (int j = tid; j < V; j += nt) if (cdf[j] > tgt) { best = j; break; }
    int lastpos = -1;
    for (int j = tid; j < V; j += nt) if (p[j] > 0.f) lastpos = j;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
      lastpos = max(lastpos, __shfl_xor_sync(0xffffffffu, lastpos, o));
    }
    __shared__ int sb[32], sl[32];
    if (tid % 32 == 0) { sb[tid / 32] = best; sl[tid / 32] = lastpos; }
    __syncthreads();
    if (tid == 0) {
      int b = 0x7fffffff, l = -1;
      for (int w = 0; w < nt / 32; ++w) { b = min(b, sb[w]); l = max(l, sl[w]); }
      chosen = b != 0x7fffffff ? b : max(l, 0);
    }
  } else {                                                          // predictor.py:333-335 argmax(probs)
    float bv = -INFINITY; int bi = 0x7fffffff;
    for (int j = tid; j < V; j += nt) { float v = p[j]; if (v > bv || (v == bv && j < bi)) { bv = v; bi = j; } }
    warp_argmax(bv, bi);
    __shared__ float sv[32]; __shared__ int si[32];
    if (tid % 32 == 0) { sv[tid / 32] = bv; si[tid / 32] = bi; }
    __syncthreads();
    if (tid == 0) {
      float v = sv[0]; int i = si[0];
      for (int w = 1; w < nt / 32; ++w) if (sv[w] > v || (sv[w] == v && si[w] < i)) { v = sv[w]; i = si[w]; }
      chosen = i == 0x7fffffff ? 0 : i;
    }
  }
  __syncthreads();
  if (tid == 0) {
    commit_token(row, chosen, step, T1, end_id, stop_rule, tokens, tok_cur, first_end, &counter);
    block_loop_exit(counter, rows, step, stop_rule, st);
  }
}

// ------------------------------------------------------------------ sampling, V <= 512: one warp per row
// (arithmetic in sample_select.cuh: warp_sample_select)
__global__ void __launch_bounds__(256) sample_select_warp_kernel(const float* __restrict__ logits, int V, int rows, float temperature,
                                                                int top_k, float top_p, int do_sample, uint64_t seed,
                                                                uint64_t offset, const float* __restrict__ uniforms,
                                                                float* __restrict__ probs_trace, int step, int T1, int end_id,
                                                                int stop_rule, int64_t* tokens, int64_t* tok_cur,
                                                                int* first_end, LoopState* st) {
  if (st->done) return;
  __shared__ int counter;
  if (threadIdx.x == 0) counter = 0;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;
  if (row < rows) {
    const float* x = logits + (size_t)row * V;
    float lg[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) lg[i] = (16 * lane + i) < V ? x[16 * lane + i] : 0.f;
    float u = 0.f;
    if (do_sample) u = uniforms ? uniforms[(size_t)step * rows + row] : philox_uniform(seed, offset + (uint64_t)step * rows + row);
    const int chosen = warp_sample_select(lg, V, lane, temperature, top_k, top_p, do_sample, u,
                                          probs_trace ? probs_trace + ((size_t)step * rows + row) * V : nullptr);
    if (lane == 0) commit_token(row, chosen, step, T1, end_id, stop_rule, tokens, tok_cur, first_end, &counter);
  }
  __syncthreads();
  if (threadIdx.x == 0) block_loop_exit(counter, rows, step, stop_rule, st);
}

// ------------------------------------------------------------------ beam (seq2seq.py:234-298)

__global__ void beam_init_kernel(BeamState* bs, double* score, int64_t* tok_cur, int* live, int B, int K,
                                 int start_id) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) {
    BeamState s; s.alive = 1; s.nbeams = 1; s.has_completed = 0; s.best_step = -1; s.best_slot = -1;
    s.best_score = 0.0; s.last_step = -1;
    bs[i] = s;
  }
  if (i < B * K) { score[i] = 0.0; tok_cur[i] = start_id; live[i] = (i % K) == 0; }
}

// One block per image, one warp per beam slot (K warps).
__global__ void beam_select_kernel(const float* __restrict__ logits, int V, int B, int K, int step,
                                   int end_id, BeamState* bstate, double* score, int64_t* tok_cur,
                                   int* live, int* parent_out, int* tr_parent, int* tr_token,
                                   double* tr_score) {
  __shared__ int c_tok[I2L_MAX_BEAM][I2L_MAX_BEAM];
  __shared__ double c_score[I2L_MAX_BEAM][I2L_MAX_BEAM];
  __shared__ int is_live[I2L_MAX_BEAM];
  const int b = blockIdx.x;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  BeamState st = bstate[b];
  const size_t tro = ((size_t)step * B + b) * K;
  if (!st.alive) {
    if (threadIdx.x < K) { tr_parent[tro + threadIdx.x] = -1; tr_token[tro + threadIdx.x] = -1;
                           tr_score[tro + threadIdx.x] = nan(""); parent_out[b * K + threadIdx.x] = -1; }
    return;
  }
  const int row = b * K + warp;
  bool lv = warp < st.nbeams && (int)tok_cur[row] != end_id;     // seq2seq.py:258-260
  if (lane == 0) is_live[warp] = lv;
  if (lv) {
    const float* x = logits + (size_t)row * V;
    float m = -INFINITY;
    for (int j = lane; j < V; j += 32) m = fmaxf(m, x[j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float s = 0.f;
    for (int j = lane; j < V; j += 32) s += expf(x[j] - m);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    float lse = logf(s);
    float pv = INFINITY; int pi = -1;                             // previously selected (value, index)
    double base = score[row];
    for (int r = 0; r < K; ++r) {                                 // torch.topk(log_probs, K): seq2seq.py:267
      float bv = -INFINITY; int bi = 0x7fffffff;
      for (int j = lane; j < V; j += 32) {
        float v = (x[j] - m) - lse;
        bool after_prev = v < pv || (v == pv && j > pi);
        if (after_prev && (v > bv || (v == bv && j < bi))) { bv = v; bi = j; }
      }
      warp_argmax(bv, bi);
      if (lane == 0) { c_tok[warp][r] = bi == 0x7fffffff ? -1 : bi; c_score[warp][r] = base + (double)bv; }
      pv = bv; pi = bi;
    }
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  // --- serial per-image bookkeeping (<= K*K candidates) ---
  for (int k = 0; k < st.nbeams; ++k) {                           // retire finished beams (258-260)
    if (!is_live[k]) {
      double sc = score[b * K + k];
      if (!st.has_completed || sc > st.best_score) {              // max(): first wins (288)
        st.has_completed = 1; st.best_score = sc; st.best_step = step - 1; st.best_slot = k;
      }
    }
  }
  int ck[I2L_MAX_BEAM], cj[I2L_MAX_BEAM];                         // kept candidates, sorted
  int nkeep = 0;
  for (int k = 0; k < st.nbeams; ++k) {
    if (!is_live[k]) continue;
    for (int j = 0; j < K; ++j) {
      if (c_tok[k][j] < 0) continue;
      double sc = c_score[k][j];
      // stable insertion (sorted(..., reverse=True) keeps candidate order among equals: 279)
      int pos = nkeep;
      while (pos > 0 && c_score[ck[pos - 1]][cj[pos - 1]] < sc) --pos;
      if (pos >= K) continue;
      int last = min(nkeep, K - 1);
      for (int q = last; q > pos; --q) { ck[q] = ck[q - 1]; cj[q] = cj[q - 1]; }
      ck[pos] = k; cj[pos] = j;
      if (nkeep < K) ++nkeep;
    }
  }
  if (nkeep == 0) {                                               // `if not candidates: break` (276-277)
    st.alive = 0;
    for (int k = 0; k < K; ++k) { tr_parent[tro + k] = -1; tr_token[tro + k] = -1; tr_score[tro + k] = nan("");
                                  parent_out[b * K + k] = -1; }
    bstate[b] = st;
    return;
  }
  bool all_end = true;
  double nsc[I2L_MAX_BEAM]; int ntk[I2L_MAX_BEAM];
  for (int k = 0; k < nkeep; ++k) { nsc[k] = c_score[ck[k]][cj[k]]; ntk[k] = c_tok[ck[k]][cj[k]];
                                    all_end = all_end && ntk[k] == end_id; }
  for (int k = 0; k < K; ++k) {
    bool v = k < nkeep;
    tr_parent[tro + k] = v ? ck[k] : -1;
    tr_token[tro + k] = v ? ntk[k] : -1;
    tr_score[tro + k] = v ? nsc[k] : nan("");
    parent_out[b * K + k] = v ? ck[k] : -1;
    score[b * K + k] = v ? nsc[k] : 0.0;
    tok_cur[b * K + k] = v ? ntk[k] : end_id;
    live[b * K + k] = v;
  }
  st.nbeams = nkeep;
  st.last_step = step;
  if (all_end) {                                                  // 282-284
    for (int k = 0; k < nkeep; ++k)
      if (!st.has_completed || nsc[k] > st.best_score) {
        st.has_completed = 1; st.best_score = nsc[k]; st.best_step = step; st.best_slot = k;
      }
    st.alive = 0;
  }
  bstate[b] = st;
}

// h/c rows follow their parent beam (the reference clones new_hidden per candidate, 272).
__global__ void beam_reorder_kernel(const float* __restrict__ h_src, const float* __restrict__ c_src,
                                    float* __restrict__ h_dst, float* __restrict__ c_dst,
                                    const int* __restrict__ parent, int R, int K, int H, int L) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t per = (size_t)R * H;
  if (i >= per * L) return;
  int l = (int)(i / per);
  size_t rem = i % per;
  int row = (int)(rem / H), j = (int)(rem % H);
  int par = parent[row];
  if (par < 0) return;
  size_t src = (size_t)l * per + (size_t)((row / K) * K + par) * H + j;
  h_dst[i] = h_src[src];
  c_dst[i] = c_src[src];
}

__global__ void beam_finalize_kernel(const BeamState* bstate, const double* score, const int* tr_parent,
                                     const int* tr_token, int B, int K, int T, int end_id,
                                     int64_t* out_tokens, int32_t* out_len, double* out_score) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  BeamState st = bstate[b];
  int step, slot; double sc;
  if (st.has_completed) { step = st.best_step; slot = st.best_slot; sc = st.best_score; }   // 287-288
  else { step = st.last_step; slot = 0; sc = step >= 0 ? score[b * K] : 0.0; }              // 289-290
  int64_t* out = out_tokens + (size_t)b * T;
  for (int i = 0; i < T; ++i) out[i] = -1;
  int len = step + 1;                       // tokens after START
  for (int t = step, k = slot; t >= 0; --t) {
    size_t o = ((size_t)t * B + b) * K + k;
    out[t] = tr_token[o];
    k = tr_parent[o];
  }
  int n = len;
  for (int i = 0; i < len; ++i) if (out[i] == end_id) { n = i; break; }   // cut at END (295-297)
  for (int i = n; i < len; ++i) out[i] = -1;
  out_len[b] = n;
  out_score[b] = sc;
}

// ------------------------------------------------------------------ attention (decoder.py:312-343)
// p1 (B,H) = hidden W[:, :H]^T + b ; p2 (B*L,H) = enc W[:, H:]^T ; one block per batch row.
__global__ void attention_combine_kernel(const float* __restrict__ p1, const float* __restrict__ p2,
                                         const float* __restrict__ v, const float* __restrict__ enc, int L,
                                         int H, int E, float* __restrict__ ctx) {
  extern __shared__ float sc[];   // L scores
  __shared__ float scratch[32];
  const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const int warp = tid / 32, lane = tid % 32, nw = nt / 32;
  for (int l = warp; l < L; l += nw) {
    const float* q = p2 + ((size_t)b * L + l) * H;
    float a = 0.f;
    for (int j = lane; j < H; j += 32) a += v[j] * tanhf(p1[(size_t)b * H + j] + q[j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) sc[l] = a;
  }
  __syncthreads();
  float lm = -INFINITY;
  for (int l = tid; l < L; l += nt) lm = fmaxf(lm, sc[l]);
  float m = block_reduce_max(lm, scratch);
  float ls = 0.f;
  for (int l = tid; l < L; l += nt) { float e = expf(sc[l] - m); sc[l] = e; ls += e; }
  float s = block_reduce_sum(ls, scratch);
  __syncthreads();
  for (int j = tid; j < E; j += nt) {
    float a = 0.f;
    for (int l = 0; l < L; ++l) a += (sc[l] / s) * enc[((size_t)b * L + l) * E + j];
    ctx[(size_t)b * E + j] = a;
  }
}

// ------------------------------------------------------------------ workspace carving
struct DecWs {
  float *gctx, *gates, *logits, *h[2], *c[2];
  __nv_bfloat16* hb;             // bf16 copy of h[0] (L,R,H): A operand of the tcgen05 GEMMs (precision bf16)
  __nv_bfloat16* hb2;            // second copy: the fused gate-GEMM + cell kernels read one and write the other
  __nv_bfloat16* encb;           // bf16 copy of the encoder output: A operand of the context-term GEMM (precision bf16)
  int64_t* tok_cur;
  int* first_end;
  LoopState* st;
  // beam only
  BeamState* bstate; double* score; int* live; int* parent; int* tr_parent; int* tr_token; double* tr_score;
  // graph-replayed loops: the captured launch sequence reads / writes ONLY workspace and packed-weight addresses;
  // the caller's tensors are copied in / out around the replay (see run_loop)
  float* g_enc; int64_t* g_tokens; int32_t* g_lengths; int32_t* g_steps; float* g_uniforms;
  size_t bytes;
};

DecWs carve(const i2l_dec_desc& d, int rows, int max_length, void* ws) {
  Arena a(ws, (size_t)-1);
  DecWs w{};
  const size_t H = d.hidden_dim, V = d.vocab_size, L = d.lstm_layers, R = rows;
  w.gctx = a.take<float>(R * 4 * H);
  w.gates = a.take<float>(R * 4 * H);
  w.logits = a.take<float>(R * V);
  for (int i = 0; i < 2; ++i) { w.h[i] = a.take<float>(L * R * H); w.c[i] = a.take<float>(L * R * H); }
  w.hb = a.take<__nv_bfloat16>(L * R * H);
  w.hb2 = a.take<__nv_bfloat16>(L * R * H);
  w.encb = a.take<__nv_bfloat16>(R * (size_t)d.embedding_dim);
  w.tok_cur = a.take<int64_t>(R);
  w.first_end = a.take<int>(R);
  w.st = a.take<LoopState>(1);
  w.bstate = a.take<BeamState>(R);
  w.score = a.take<double>(R);
  w.live = a.take<int>(R);
  w.parent = a.take<int>(R);
  size_t T = max_length > 0 ? max_length : 1;
  w.tr_parent = a.take<int>(T * R);
  w.tr_token = a.take<int>(T * R);
  w.tr_score = a.take<double>(T * R);
  w.g_enc = a.take<float>(R * (size_t)d.embedding_dim);
  w.g_tokens = a.take<int64_t>(R * (T + 1));
  w.g_lengths = a.take<int32_t>(R);
  w.g_steps = a.take<int32_t>(1);
  w.g_uniforms = a.take<float>(T * R);
  w.bytes = align_up(a.off, 256);
  return w;
}

// I2L_BF16_STREAMED = bf16 arithmetic on the stream-ordered path only (no persistent cluster kernel): same packed
// layout and workspace as I2L_BF16, so one packed model serves both; the entry points work on the normalised copy.
struct DescN {
  i2l_dec_desc d;
  bool streamed;
  explicit DescN(const i2l_dec_desc* in) : d(in ? *in : i2l_dec_desc{}), streamed(in && in->precision == I2L_BF16_STREAMED) {
    if (streamed) d.precision = I2L_BF16;
  }
};

int check_common(const i2l_dec_desc* d, const void* packed) {
  I2L_TRY(device_check());
  I2L_REQUIRE(d != nullptr && packed != nullptr, "decoder: null descriptor / packed weights");
  I2L_REQUIRE(d->vocab_size > 0 && d->embedding_dim > 0 && d->hidden_dim > 0 && d->lstm_layers >= 1 &&
              d->lstm_layers <= I2L_MAX_LSTM_LAYERS, "decoder: invalid dimensions");
  return I2L_OK;
}

// One LSTM-stack step + vocab projection for `rows` rows (decoder.py:274-280, attention with
// src_len==1 is the identity: SURVEY F3).  h/c updated in place.
int step_rows(const i2l_dec_desc& d, const float* pk, const PackedDec& lay, const DecWs& w, float* h, float* c,
              int rows, const int* skip, cudaStream_t s) {
  const int H = d.hidden_dim, V = d.vocab_size;
  for (int l = 0; l < d.lstm_layers; ++l) {
    GemmF32 g;
    g.M = rows; g.N = 4 * H; g.C = w.gates; g.ldc = 4 * H; g.skip_flag = skip;
    float* hl = h + (size_t)l * rows * H;
    float* cl = c + (size_t)l * rows * H;
    if (l == 0) {
      g.A1 = hl; g.lda1 = H; g.W1 = pk + lay.w_hh[0]; g.ldw1 = H; g.K1 = H;
      g.add_rows = w.gctx; g.ld_add = 4 * H;
      g.add_table = pk + lay.gtok; g.ld_tab = 4 * H; g.tab_idx = w.tok_cur;
    } else {
      g.A1 = h + (size_t)(l - 1) * rows * H; g.lda1 = H; g.W1 = pk + lay.w_ih[l]; g.ldw1 = H; g.K1 = H;
      g.A2 = hl; g.lda2 = H; g.W2 = pk + lay.w_hh[l]; g.ldw2 = H; g.K2 = H;
      g.bias = pk + lay.bsum[l];
    }
    I2L_TRY(gemm_f32(g, s));
    I2L_TRY(lstm_cell_f32(w.gates, hl, cl, rows, H, skip, s));
  }
  GemmF32 g;
  g.M = rows; g.N = V; g.C = w.logits; g.ldc = V; g.skip_flag = skip;
  g.A1 = h + (size_t)(d.lstm_layers - 1) * rows * H; g.lda1 = H; g.W1 = pk + lay.out_w; g.ldw1 = H; g.K1 = H;
  g.bias = pk + lay.out_b;
  return gemm_f32(g, s);
}

// The same step with the products on tcgen05 (gemm_bf16.cu): bf16 weights (PackedDec::g16*) and the bf16 copy of h
// written by the cell kernel; gate pre-activations, cell state, logits and every epilogue term stay fp32.
struct StepBf16 { GemmBf16 gates[I2L_MAX_LSTM_LAYERS]; GemmBf16 logits; };

int make_step_bf16(const i2l_dec_desc& d, const void* packed, const PackedDec& lay, const DecWs& w, int rows,
                   const int* skip, StepBf16* st) {
  const int H = d.hidden_dim, V = d.vocab_size;
  const float* pk = reinterpret_cast<const float*>(packed);
  const char* pb = reinterpret_cast<const char*>(packed);
  for (int l = 0; l < d.lstm_layers; ++l) {
    GemmBf16& g = st->gates[l];
    g = GemmBf16{};
    g.M = rows; g.N = 4 * H; g.C = w.gates; g.ldc = 4 * H; g.skip_flag = skip;
    const __nv_bfloat16* hl = w.hb + (size_t)l * rows * H;
    if (l == 0) {
      I2L_TRY(gemm_bf16_a_map(&g.tmA1, hl, rows, H, H));
      I2L_TRY(gemm_bf16_w_map(&g.tmW1, pb + lay.g16_w_hh[0], 4 * H, H, H));
      g.K1 = H;
      g.add_rows = w.gctx; g.ld_add = 4 * H;
      g.add_table = pk + lay.gtok; g.ld_tab = 4 * H; g.tab_idx = w.tok_cur;
    } else {
      I2L_TRY(gemm_bf16_a_map(&g.tmA1, w.hb + (size_t)(l - 1) * rows * H, rows, H, H));
      I2L_TRY(gemm_bf16_w_map(&g.tmW1, pb + lay.g16_w_ih[l], 4 * H, H, H));
      I2L_TRY(gemm_bf16_a_map(&g.tmA2, hl, rows, H, H));
      I2L_TRY(gemm_bf16_w_map(&g.tmW2, pb + lay.g16_w_hh[l], 4 * H, H, H));
      g.K1 = H; g.K2 = H;
      g.bias = pk + lay.bsum[l];
    }
  }
  GemmBf16& g = st->logits;
  g = GemmBf16{};
  g.M = rows; g.N = V; g.C = w.logits; g.ldc = V; g.skip_flag = skip;
  I2L_TRY(gemm_bf16_a_map(&g.tmA1, w.hb + (size_t)(d.lstm_layers - 1) * rows * H, rows, H, H));
  I2L_TRY(gemm_bf16_w_map(&g.tmW1, pb + lay.g16_out_w, V, H, H));
  g.K1 = H;
  g.bias = pk + lay.out_b;
  return I2L_OK;
}

// Fused variant (H % 32 == 0): per layer ONE launch = gate GEMM + LSTM cell in the epilogue (no (rows,4H) fp32 gates
// round trip, no cell kernel).  h is double buffered in bf16: even steps read w.hb and write w.hb2, odd steps the
// reverse -- a launch must not overwrite the operand other CTAs are still reading.  par[k] = structures of a step
// that READS buffer k.
struct StepFused { GemmBf16 gates[2][I2L_MAX_LSTM_LAYERS]; GemmBf16 logits[2]; };

int make_step_fused(const i2l_dec_desc& d, const void* packed, const PackedDec& lay, const DecWs& w, float* h, float* c,
                    int rows, const int* skip, StepFused* st) {
  const int H = d.hidden_dim, V = d.vocab_size;
  const float* pk = reinterpret_cast<const float*>(packed);
  const char* pb = reinterpret_cast<const char*>(packed);
  for (int par = 0; par < 2; ++par) {
    __nv_bfloat16* rd = par == 0 ? w.hb : w.hb2;          // h of the previous step
    __nv_bfloat16* wr = par == 0 ? w.hb2 : w.hb;          // h of this step
    for (int l = 0; l < d.lstm_layers; ++l) {
      GemmBf16& g = st->gates[par][l];
      g = GemmBf16{};
      g.M = rows; g.N = 4 * H; g.skip_flag = skip;
      g.cell_c = c + (size_t)l * rows * H; g.cell_h = h + (size_t)l * rows * H; g.cell_hb = wr + (size_t)l * rows * H;
      g.cell_H = H;
      if (l == 0) {
        I2L_TRY(gemm_bf16_a_map(&g.tmA1, rd, rows, H, H));
        I2L_TRY(gemm_bf16_w_map(&g.tmW1, pb + lay.g16c_w_hh[0], 4 * H, H, H));
        g.K1 = H;
        g.add_rows = w.gctx; g.ld_add = 4 * H;
        g.add_table = pk + lay.gtok; g.ld_tab = 4 * H; g.tab_idx = w.tok_cur;
      } else {
        // recurrent product first, then this step's h of the layer below: the accumulation order of decode_wide.cu,
        // so that the two paths agree bit for bit
        I2L_TRY(gemm_bf16_a_map(&g.tmA1, rd + (size_t)l * rows * H, rows, H, H));
        I2L_TRY(gemm_bf16_w_map(&g.tmW1, pb + lay.g16c_w_hh[l], 4 * H, H, H));
        I2L_TRY(gemm_bf16_a_map(&g.tmA2, wr + (size_t)(l - 1) * rows * H, rows, H, H));
        I2L_TRY(gemm_bf16_w_map(&g.tmW2, pb + lay.g16c_w_ih[l], 4 * H, H, H));
        g.K1 = H; g.K2 = H;
        g.bias = pk + lay.bsum[l];
      }
    }
    GemmBf16& g = st->logits[par];
    g = GemmBf16{};
    g.M = rows; g.N = V; g.C = w.logits; g.ldc = V; g.skip_flag = skip;
    I2L_TRY(gemm_bf16_a_map(&g.tmA1, wr + (size_t)(d.lstm_layers - 1) * rows * H, rows, H, H));
    I2L_TRY(gemm_bf16_w_map(&g.tmW1, pb + lay.g16_out_w, V, H, H));
    g.K1 = H;
    g.bias = pk + lay.out_b;
  }
  return I2L_OK;
}

#ifdef I2L_DIAG   // diagnostics build: per-kernel CUDA-event timers (recorded as event nodes when the loop is captured)
#define I2L_DIAG_TIMER(name) KernelTimer kt_diag_(name, s)
#else
#define I2L_DIAG_TIMER(name) do { } while (0)
#endif

int step_rows_fused(const i2l_dec_desc& d, const StepFused& st, int step, cudaStream_t s) {
  const int par = step & 1;
  for (int l = 0; l < d.lstm_layers; ++l) {
    I2L_DIAG_TIMER(l == 0 ? "gen.gate_cell_l0" : "gen.gate_cell_l1plus");
    I2L_TRY(gemm_bf16(st.gates[par][l], s));
  }
  I2L_DIAG_TIMER("gen.logits_gemm");
  return gemm_bf16(st.logits[par], s);
}

int step_rows_bf16(const i2l_dec_desc& d, const StepBf16& st, const DecWs& w, float* h, float* c, int rows,
                   const int* skip, cudaStream_t s) {
  const int H = d.hidden_dim;
  for (int l = 0; l < d.lstm_layers; ++l) {
    I2L_TRY(gemm_bf16(st.gates[l], s));
    I2L_TRY(lstm_cell_f32(w.gates, h + (size_t)l * rows * H, c + (size_t)l * rows * H, rows, H, skip, s,
                          w.hb + (size_t)l * rows * H));
  }
  return gemm_bf16(st.logits, s);
}

// gctx (rows,4H) = enc W_ih0[:, E:2E]^T + b_ih0 + b_hh0 ; enc row index = row / rows_per_enc
int make_gctx(const i2l_dec_desc& d, const float* pk, const PackedDec& lay, const float* enc, int n_enc,
              float* gctx, cudaStream_t s) {
  GemmF32 g;
  g.M = n_enc; g.N = 4 * d.hidden_dim; g.C = gctx; g.ldc = 4 * d.hidden_dim;
  g.A1 = enc; g.lda1 = d.embedding_dim; g.W1 = pk + lay.w_ih0 + d.embedding_dim; g.ldw1 = 2 * d.embedding_dim;
  g.K1 = d.embedding_dim; g.bias = pk + lay.bsum[0];
  return gemm_f32(g, s);
}

__global__ void repeat_rows_kernel(const float* __restrict__ src, float* __restrict__ dst, int n_src, int rep,
                                   int width) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t total = (size_t)n_src * rep * width;
  if (i >= total) return;
  size_t row = i / width; int j = (int)(i % width);
  dst[i] = src[(row / rep) * width + j];
}

}  // namespace

int lstm_cell_f32(const float* gates, float* h, float* c, int rows, int H, const int* skip_flag,
                  cudaStream_t s, __nv_bfloat16* hb, float* h_seq, __nv_bfloat16* hb_seq, size_t seq_ld) {
  size_t total = (size_t)rows * H;
  if (total == 0) return I2L_OK;
  lstm_cell_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(gates, h, c, rows, H, skip_flag, hb, h_seq, hb_seq,
                                                                   seq_ld);
  I2L_LAUNCH_OK();
  return I2L_OK;
}

int make_gctx_bf16(const i2l_dec_desc& d, const void* packed, const PackedDec& lay, const float* enc, int n, float* gctx,
                   __nv_bfloat16* encb, cudaStream_t s) {
  const int E = d.embedding_dim, H = d.hidden_dim;
  const size_t tot = (size_t)n * E;
  if (tot == 0) return I2L_OK;
  f32_to_bf16_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(enc, encb, tot);
  I2L_LAUNCH_OK();
  GemmBf16 g;
  g.M = n; g.N = 4 * H; g.C = gctx; g.ldc = 4 * H;
  I2L_TRY(gemm_bf16_a_map(&g.tmA1, encb, n, E, E));
  I2L_TRY(gemm_bf16_w_map(&g.tmW1, reinterpret_cast<const char*>(packed) + lay.g16_w_ctx, 4 * H, E, E));
  g.K1 = E;
  g.bias = reinterpret_cast<const float*>(packed) + lay.bsum[0];
  return gemm_bf16(g, s);
}

}  // namespace i2l

using namespace i2l;

// ====================================================================== C ABI
extern "C" size_t i2l_dec_packed_bytes(const i2l_dec_desc* d) {
  if (!d) return 0;
  return dec_layout(DescN(d).d).total_bytes;
}

extern "C" int i2l_dec_pack(const i2l_dec_desc* d_in, const i2l_dec_params* p, void* packed, size_t packed_bytes,
                            void* stream) {
  const DescN dn_(d_in);
  const i2l_dec_desc* d = d_in ? &dn_.d : nullptr;
  I2L_TRY(check_common(d, packed));
  I2L_REQUIRE(p != nullptr, "i2l_dec_pack: null params");
  cudaStream_t s = (cudaStream_t)stream;
  PackedDec lay = dec_layout(*d);
  if (packed_bytes < lay.total_bytes) { set_error("i2l_dec_pack: packed buffer too small (%zu < %zu)", packed_bytes, lay.total_bytes); return I2L_ERR_WORKSPACE; }
  float* pk = reinterpret_cast<float*>(packed);
  const size_t V = d->vocab_size, E = d->embedding_dim, H = d->hidden_dim;
  I2L_REQUIRE(p->embedding && p->out_w && p->out_b, "i2l_dec_pack: missing tensors");
  I2L_CUDA_OK(cudaMemcpyAsync(pk + lay.emb, p->embedding, V * E * 4, cudaMemcpyDeviceToDevice, s));
  for (int l = 0; l < d->lstm_layers; ++l) {
    I2L_REQUIRE(p->w_ih[l] && p->w_hh[l] && p->b_ih[l] && p->b_hh[l], "i2l_dec_pack: missing LSTM layer %d", l);
    size_t in = l == 0 ? 2 * E : H;
    I2L_CUDA_OK(cudaMemcpyAsync(pk + lay.w_ih[l], p->w_ih[l], 4 * H * in * 4, cudaMemcpyDeviceToDevice, s));
    I2L_CUDA_OK(cudaMemcpyAsync(pk + lay.w_hh[l], p->w_hh[l], 4 * H * H * 4, cudaMemcpyDeviceToDevice, s));
    add_vec_kernel<<<cdiv(4 * (int)H, 256), 256, 0, s>>>(p->b_ih[l], p->b_hh[l], pk + lay.bsum[l], 4 * (int)H);
    I2L_LAUNCH_OK();
  }
  I2L_CUDA_OK(cudaMemcpyAsync(pk + lay.out_w, p->out_w, V * H * 4, cudaMemcpyDeviceToDevice, s));
  I2L_CUDA_OK(cudaMemcpyAsync(pk + lay.out_b, p->out_b, V * 4, cudaMemcpyDeviceToDevice, s));
  // token -> gate table: gtok[v] = W_ih0[:, :E] emb[v]   
  GemmF32 g;
  g.M = (int)V; g.N = 4 * (int)H; g.C = pk + lay.gtok; g.ldc = 4 * (int)H;
  g.A1 = pk + lay.emb; g.lda1 = (int)E; g.W1 = pk + lay.w_ih0; g.ldw1 = 2 * (int)E; g.K1 = (int)E;
  I2L_TRY(gemm_f32(g, s));
  if (lay.bf16_section) I2L_TRY(persistent_pack(*d, *p, pk + lay.gtok, reinterpret_cast<char*>(packed) + lay.bf16_section, s));
  if (lay.g16) {
    char* pb = reinterpret_cast<char*>(packed);
    auto conv = [&](const float* src, size_t off, size_t n) -> int {
      f32_to_bf16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(src, reinterpret_cast<__nv_bfloat16*>(pb + off), n);
      I2L_LAUNCH_OK();
      return I2L_OK;
    };
    for (int l = 0; l < d->lstm_layers; ++l) {
      I2L_TRY(conv(p->w_hh[l], lay.g16_w_hh[l], 4 * H * H));
      if (l > 0) I2L_TRY(conv(p->w_ih[l], lay.g16_w_ih[l], 4 * H * H));
    }
    I2L_TRY(conv(p->out_w, lay.g16_out_w, V * H));
    if (lay.g16c) {
      auto convc = [&](const float* src, size_t off) -> int {
        const size_t n = 4 * H * H;
        pack_cell_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(src, reinterpret_cast<__nv_bfloat16*>(pb + off), (int)H, (int)H);
        I2L_LAUNCH_OK();
        return I2L_OK;
      };
      for (int l = 0; l < d->lstm_layers; ++l) {
        I2L_TRY(convc(p->w_hh[l], lay.g16c_w_hh[l]));
        if (l > 0) I2L_TRY(convc(p->w_ih[l], lay.g16c_w_ih[l]));
      }
    }
    {
      const size_t n = 4 * H * E;
      f32_to_bf16_ld_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(p->w_ih[0] + E, 2 * (int)E, reinterpret_cast<__nv_bfloat16*>(pb + lay.g16_w_ctx), (int)E, n);
      I2L_LAUNCH_OK();
    }
  }
  return I2L_OK;
}

extern "C" size_t i2l_dec_workspace_bytes(const i2l_dec_desc* d_in, int32_t rows, int32_t max_length) {
  const DescN dn_(d_in);
  const i2l_dec_desc* d = d_in ? &dn_.d : nullptr;
  if (!d || rows <= 0) return 0;
  size_t b = carve(*d, rows, max_length, nullptr).bytes;
  if (d->precision == I2L_BF16 && persistent_supported(*d)) b += persistent_workspace_bytes(*d, rows, max_length);
  if (wide_supported(*d)) b += wide_workspace_bytes(*d, rows, max_length);
  return b;
}

extern "C" int i2l_decode_step(const i2l_dec_desc* d_in, const void* packed, const float* enc, const int64_t* tok,
                               int32_t batch, const float* h_in, const float* c_in, float* logits, float* h_out,
                               float* c_out, int32_t* bad_token_flag, void* workspace, size_t workspace_bytes, void* stream) {
  const DescN dn_(d_in);
  const i2l_dec_desc* d = d_in ? &dn_.d : nullptr;
  I2L_TRY(check_common(d, packed));
  I2L_REQUIRE(batch >= 0 && enc && tok && logits && h_out && c_out, "i2l_decode_step: null argument");
  I2L_REQUIRE((h_in == nullptr) == (c_in == nullptr), "i2l_decode_step: h_in and c_in must both be given or both NULL");
  if (batch == 0) return I2L_OK;
  cudaStream_t s = (cudaStream_t)stream;
  PackedDec lay = dec_layout(*d);
  DecWs w = carve(*d, batch, 1, workspace);
  if (workspace_bytes < w.bytes) { set_error("i2l_decode_step: workspace too small (%zu < %zu)", workspace_bytes, w.bytes); return I2L_ERR_WORKSPACE; }
  const float* pk = reinterpret_cast<const float*>(packed);
  size_t n = (size_t)d->lstm_layers * batch * d->hidden_dim * sizeof(float);
  if (h_in) {
    I2L_CUDA_OK(cudaMemcpyAsync(h_out, h_in, n, cudaMemcpyDeviceToDevice, s));
    I2L_CUDA_OK(cudaMemcpyAsync(c_out, c_in, n, cudaMemcpyDeviceToDevice, s));
  } else {                                                       // decoder.py:253-266
    I2L_CUDA_OK(cudaMemsetAsync(h_out, 0, n, s));
    I2L_CUDA_OK(cudaMemsetAsync(c_out, 0, n, s));
  }
  // ids outside [0, V) never reach the gate-table gather: read as 0, reported through the flag (nn.Embedding: IndexError)
  transpose_tokens_kernel<<<cdiv(batch, 256), 256, 0, s>>>(tok, w.tok_cur, batch, 1, d->vocab_size, bad_token_flag);
  I2L_LAUNCH_OK();
  I2L_TRY(make_gctx(*d, pk, lay, enc, batch, w.gctx, s));
  DecWs w2 = w; w2.logits = logits;
  return step_rows(*d, pk, lay, w2, h_out, c_out, batch, nullptr, s);
}

// ---------------------------------------------------------------------------------------------
// The stream-ordered decode loops as ONE CUDA graph.  A step of the general path is 2 L + 2 small launches (gate GEMM
// + cell per layer, logits GEMM, selection) whose kernels run 1-4 us each: issued one by one the loop is bound by
// the host's launch rate (~5-7 us per launch through the C-ABI, 40-50 us per step for L = 2).  All loop bookkeeping
// already lives on the device (token append, EOS masks, both stop rules, the `done` flag every kernel checks), so the
// whole max_length-step sequence is captured once per (weights, workspace, shape, loop arguments) and replayed with a
// single cudaGraphLaunch.  The captured kernels read and write only workspace / packed-weight addresses; the caller's
// enc / uniforms are copied in before the replay and tokens / lengths / steps are copied out after it, so fresh
// output tensors per call do not invalidate the graph.  Capture runs on a private non-blocking stream (the caller's
// may be the legacy default stream, which cannot be captured); nothing executes during capture.
struct LoopKey {
  const void* packed; const void* ws; i2l_dec_desc d;
  int batch, start_id, end_id, max_length, stop_rule, sampling, top_k, has_uniforms, dev;
  float temperature, top_p; unsigned long long seed, offset;
  bool operator==(const LoopKey& o) const { return memcmp(this, &o, sizeof(LoopKey)) == 0; }
};
struct LoopGraph { LoopKey key; cudaGraphExec_t exec; long long kernels; unsigned long long last_use; };
constexpr int kLoopGraphs = 16;
static std::mutex g_graph_mu;
static LoopGraph g_graphs[kLoopGraphs];
static int g_n_graphs = 0;
static unsigned long long g_graph_tick = 0;

static int enqueue_loop(const i2l_dec_desc* d, const void* packed, const PackedDec& lay, const DecWs& w, const float* enc,
                        int batch, int start_id, int end_id, int max_length, float temperature, int stop_rule,
                        bool sampling_path, int top_k, float top_p, uint64_t seed, uint64_t offset, const float* uniforms,
                        int64_t* tokens, int32_t* lengths, int32_t* steps_run, float* probs_trace, cudaStream_t s) {
  const float* pk = reinterpret_cast<const float*>(packed);
  const int T1 = max_length + 1, V = d->vocab_size;
  size_t tot = (size_t)batch * T1;
  loop_init_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(tokens, T1, batch, start_id, w.tok_cur, w.first_end, w.st);
  I2L_LAUNCH_OK();
  size_t n = (size_t)d->lstm_layers * batch * d->hidden_dim * sizeof(float);
  I2L_CUDA_OK(cudaMemsetAsync(w.h[0], 0, n, s));
  I2L_CUDA_OK(cudaMemsetAsync(w.c[0], 0, n, s));
  // precision bf16: the context term on the tensor cores too (bf16 operands, like decode_persistent.cu / decode_wide.cu)
  if (lay.g16 != 0) I2L_TRY(make_gctx_bf16(*d, packed, lay, enc, batch, w.gctx, w.encb, s));
  else I2L_TRY(make_gctx(*d, pk, lay, enc, batch, w.gctx, s));
  const int* skip = &w.st->done;
  int n2 = 1; while (n2 < V) n2 <<= 1;
  size_t smem = (size_t)V * 8 + (size_t)V * 4 + (size_t)n2 * 8;
  if (sampling_path) {
    I2L_REQUIRE(smem <= 200 * 1024, "decode_sample: vocabulary too large for the on-chip sort (%d)", V);
    I2L_CUDA_OK(cudaFuncSetAttribute(sample_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  const int do_sample = temperature > 0.f && (top_k > 0 || top_p > 0.0f);   // predictor.py:330
  const bool tc = lay.g16 != 0;                       // precision bf16: the per-step GEMMs run on tcgen05
  const bool fused = tc && lay.g16c != 0;             // ... with the LSTM cell fused into the gate GEMM's epilogue
  StepBf16 st16;
  static thread_local StepFused stf;                  // ~10 KB of tensor maps: not on the stack
  if (tc) {
    I2L_CUDA_OK(cudaMemsetAsync(w.hb, 0, (size_t)d->lstm_layers * batch * d->hidden_dim * 2, s));
    if (fused) I2L_TRY(make_step_fused(*d, packed, lay, w, w.h[0], w.c[0], batch, skip, &stf));
    else I2L_TRY(make_step_bf16(*d, packed, lay, w, batch, skip, &st16));
  }
  for (int step = 0; step < max_length; ++step) {
    if (fused) I2L_TRY(step_rows_fused(*d, stf, step, s));
    else if (tc) I2L_TRY(step_rows_bf16(*d, st16, w, w.h[0], w.c[0], batch, skip, s));
    else I2L_TRY(step_rows(*d, pk, lay, w, w.h[0], w.c[0], batch, skip, s));
    if (sampling_path && V <= 512) {
      sample_select_warp_kernel<<<cdiv(batch, 8), 256, 0, s>>>(w.logits, V, batch, temperature, top_k, top_p, do_sample,
                                                              seed, offset, uniforms, probs_trace, step, T1, end_id,
                                                              stop_rule, tokens, w.tok_cur, w.first_end, w.st);
    } else if (sampling_path) {
      sample_select_kernel<<<batch, 256, smem, s>>>(w.logits, V, n2, batch, temperature, top_k, top_p, do_sample,
                                                   seed, offset, uniforms, probs_trace, step, T1, end_id,
                                                   stop_rule, tokens, w.tok_cur, w.first_end, w.st);
    } else {
      I2L_DIAG_TIMER("gen.greedy_select");
      greedy_select_kernel<<<cdiv(batch, 8), 256, 0, s>>>(w.logits, V, batch, temperature, step, T1, end_id,
                                                          stop_rule, tokens, w.tok_cur, w.first_end, w.st);
    }
    I2L_LAUNCH_OK();
  }
  loop_finalize_kernel<<<cdiv(batch, 256), 256, 0, s>>>(w.first_end, batch, max_length, lengths, steps_run, w.st);
  I2L_LAUNCH_OK();
  return I2L_OK;
}

static int run_loop(const i2l_dec_desc* d, const void* packed, const float* enc, int batch, int start_id,
                    int end_id, int max_length, float temperature, int stop_rule, bool sampling_path,
                    int top_k, float top_p, uint64_t seed, uint64_t offset, const float* uniforms,
                    int64_t* tokens, int32_t* lengths, int32_t* steps_run, float* probs_trace, void* workspace,
                    size_t workspace_bytes, cudaStream_t s) {
  I2L_TRY(check_common(d, packed));
  I2L_REQUIRE(batch >= 0 && max_length >= 0 && tokens != nullptr, "decode loop: invalid arguments");
  I2L_REQUIRE(batch == 0 || enc != nullptr, "decode loop: null encoder output");
  I2L_REQUIRE(stop_rule >= 0 && stop_rule <= 2, "decode loop: invalid stop rule");
  I2L_REQUIRE(start_id >= 0 && start_id < d->vocab_size, "decode loop: start token %d outside [0, %d) (nn.Embedding raises IndexError)",
              start_id, d->vocab_size);
  if (batch == 0) return I2L_OK;
  PackedDec lay = dec_layout(*d);
  DecWs w = carve(*d, batch, max_length, workspace);
  if (workspace_bytes < w.bytes) { set_error("decode loop: workspace too small (%zu < %zu)", workspace_bytes, w.bytes); return I2L_ERR_WORKSPACE; }
  const char* timer_name = sampling_path ? "dec.sample_loop_general" : "dec.greedy_loop_general";
  // graph replay: loops long enough to amortise a capture, no per-step trace requested, caller not capturing itself
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(s, &cap) != cudaSuccess) { cudaGetLastError(); cap = cudaStreamCaptureStatusActive; }
  const bool use_graph = max_length >= 8 && probs_trace == nullptr && cap == cudaStreamCaptureStatusNone &&
                         lengths != nullptr && steps_run != nullptr;
  if (!use_graph) {
    KernelTimer kt(timer_name, s);
    return enqueue_loop(d, packed, lay, w, enc, batch, start_id, end_id, max_length, temperature, stop_rule, sampling_path,
                        top_k, top_p, seed, offset, uniforms, tokens, lengths, steps_run, probs_trace, s);
  }
  LoopKey key;
  memset(&key, 0, sizeof(key));
  key.packed = packed; key.ws = workspace; key.d = *d; key.batch = batch; key.start_id = start_id; key.end_id = end_id;
  key.max_length = max_length; key.stop_rule = stop_rule; key.sampling = sampling_path ? 1 : 0; key.top_k = top_k;
  key.has_uniforms = uniforms != nullptr; key.temperature = temperature; key.top_p = top_p;
  key.seed = sampling_path && !uniforms ? seed : 0; key.offset = sampling_path && !uniforms ? offset : 0;
  I2L_CUDA_OK(cudaGetDevice(&key.dev));
  cudaGraphExec_t exec = nullptr;
  long long kernels = 0;
  {
    std::lock_guard<std::mutex> lk(g_graph_mu);
    for (int i = 0; i < g_n_graphs; ++i)
      if (g_graphs[i].key == key) { exec = g_graphs[i].exec; kernels = g_graphs[i].kernels; g_graphs[i].last_use = ++g_graph_tick; break; }
    if (exec == nullptr) {
      static thread_local cudaStream_t cs = nullptr;          // capture stream (never executes anything)
      if (cs == nullptr) I2L_CUDA_OK(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
      I2L_CUDA_OK(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
      set_launch_counting(false);
      const int rc = enqueue_loop(d, packed, lay, w, w.g_enc, batch, start_id, end_id, max_length, temperature, stop_rule,
                                  sampling_path, top_k, top_p, seed, offset, uniforms ? w.g_uniforms : nullptr, w.g_tokens,
                                  w.g_lengths, w.g_steps, nullptr, cs);
      set_launch_counting(true);
      cudaGraph_t graph = nullptr;
      const cudaError_t ce = cudaStreamEndCapture(cs, &graph);
      if (rc != I2L_OK) { if (graph) cudaGraphDestroy(graph); cudaGetLastError(); return rc; }
      if (ce != cudaSuccess) { set_error("decode loop: graph capture failed: %s", cudaGetErrorString(ce)); cudaGetLastError(); return I2L_ERR_CUDA; }
      size_t n_nodes = 0;
      cudaGraphGetNodes(graph, nullptr, &n_nodes);
      std::vector<cudaGraphNode_t> nodes(n_nodes);
      if (n_nodes) cudaGraphGetNodes(graph, nodes.data(), &n_nodes);
      for (size_t i = 0; i < n_nodes; ++i) {
        cudaGraphNodeType t;
        if (cudaGraphNodeGetType(nodes[i], &t) == cudaSuccess && t == cudaGraphNodeTypeKernel) ++kernels;
      }
      const cudaError_t ie = cudaGraphInstantiate(&exec, graph, 0);
      cudaGraphDestroy(graph);
      if (ie != cudaSuccess) { set_error("decode loop: graph instantiation failed: %s", cudaGetErrorString(ie)); cudaGetLastError(); return I2L_ERR_CUDA; }
      int slot = g_n_graphs;
      if (g_n_graphs < kLoopGraphs) ++g_n_graphs;
      else {                                                    // evict the least recently used graph
        slot = 0;
        for (int i = 1; i < kLoopGraphs; ++i) if (g_graphs[i].last_use < g_graphs[slot].last_use) slot = i;
        cudaGraphExecDestroy(g_graphs[slot].exec);
      }
      g_graphs[slot] = LoopGraph{key, exec, kernels, ++g_graph_tick};
    }
  }
  I2L_CUDA_OK(cudaMemcpyAsync(w.g_enc, enc, (size_t)batch * d->embedding_dim * sizeof(float), cudaMemcpyDeviceToDevice, s));
  if (uniforms)
    I2L_CUDA_OK(cudaMemcpyAsync(w.g_uniforms, uniforms, (size_t)max_length * batch * sizeof(float), cudaMemcpyDeviceToDevice, s));
  {
    KernelTimer kt(timer_name, s);
    I2L_CUDA_OK(cudaGraphLaunch(exec, s));
  }
  count_launches(kernels);
  I2L_CUDA_OK(cudaMemcpyAsync(tokens, w.g_tokens, (size_t)batch * (max_length + 1) * sizeof(int64_t), cudaMemcpyDeviceToDevice, s));
  I2L_CUDA_OK(cudaMemcpyAsync(lengths, w.g_lengths, (size_t)batch * sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
  I2L_CUDA_OK(cudaMemcpyAsync(steps_run, w.g_steps, sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
  return I2L_OK;
}

extern "C" int i2l_decode_greedy(const i2l_dec_desc* d_in, const void* packed, const float* enc, int32_t batch,
                                 int32_t start_id, int32_t end_id, int32_t max_length, float temperature,
                                 int32_t stop_rule, int64_t* tokens, int32_t* lengths, int32_t* steps_run,
                                 void* workspace, size_t workspace_bytes, void* stream) {
  const DescN dn_(d_in);
  const i2l_dec_desc* d = d_in ? &dn_.d : nullptr;
  I2L_TRY(check_common(d, packed));
  cudaStream_t s = (cudaStream_t)stream;
  if (!dn_.streamed && d->precision == I2L_BF16 && persistent_supported(*d) && batch > 0 && temperature > 0.f) {
    PackedDec lay = dec_layout(*d);
    size_t gen = carve(*d, batch, max_length, nullptr).bytes;
    I2L_REQUIRE(workspace_bytes >= gen + persistent_workspace_bytes(*d, batch, max_length),
                "i2l_decode_greedy: workspace too small");
    return persistent_greedy(*d, reinterpret_cast<const char*>(packed) + lay.bf16_section,
                             reinterpret_cast<const float*>(packed), lay, enc, batch, start_id, end_id,
                             max_length, temperature, stop_rule, tokens, lengths, steps_run,
                             reinterpret_cast<char*>(workspace) + gen, workspace_bytes - gen, s);
  }
  if (!dn_.streamed && wide_batch_supported(*d, batch) && temperature > 0.f && max_length > 0 && lengths && steps_run) {
    // decoders beyond the cluster kernel (the reference's shipped 512 / 512 / 2, 1024 / 1024 / 3): one cooperative
    // kernel runs the whole loop on the whole GPU (decode_wide.cu)
    PackedDec lay = dec_layout(*d);
    size_t gen = carve(*d, batch, max_length, nullptr).bytes;
    I2L_REQUIRE(workspace_bytes >= gen + wide_workspace_bytes(*d, batch, max_length), "i2l_decode_greedy: workspace too small");
    return wide_greedy(*d, packed, lay, enc, batch, start_id, end_id, max_length, temperature, stop_rule, tokens, lengths,
                       steps_run, reinterpret_cast<char*>(workspace) + gen, workspace_bytes - gen, s);
  }
  return run_loop(d, packed, enc, batch, start_id, end_id, max_length, temperature, stop_rule, false, 0, 0.f, 0,
                  0, nullptr, tokens, lengths, steps_run, nullptr, workspace, workspace_bytes, s);
}

extern "C" int i2l_decode_sample(const i2l_dec_desc* d_in, const void* packed, const float* enc, int32_t batch,
                                 int32_t start_id, int32_t end_id, int32_t max_length, float temperature,
                                 int32_t top_k, float top_p, uint64_t seed, uint64_t offset,
                                 const float* uniforms, int64_t* tokens, int32_t* lengths, int32_t* steps_run,
                                 float* probs_trace, void* workspace, size_t workspace_bytes, void* stream) {
  const DescN dn_(d_in);
  const i2l_dec_desc* d = d_in ? &dn_.d : nullptr;
  I2L_TRY(check_common(d, packed));
  // predictor.py:295 divides the logits by the temperature: 0 gives +-inf / NaN "probabilities" in the reference; refused here
  I2L_REQUIRE(temperature != 0.f, "i2l_decode_sample: temperature must be non-zero (the logits are divided by it)");
  cudaStream_t s = (cudaStream_t)stream;
  const bool no_persistent = dn_.streamed;
  if (!no_persistent && d->precision == I2L_BF16 && persistent_supported(*d) && batch > 0 && max_length > 0 &&
      end_id >= 0 && end_id < d->vocab_size && tokens != nullptr) {
    // headline decoder shape in bf16: the whole sampling loop runs inside the persistent cluster kernel
    PackedDec lay = dec_layout(*d);
    size_t gen = carve(*d, batch, max_length, nullptr).bytes;
    I2L_REQUIRE(workspace_bytes >= gen + persistent_workspace_bytes(*d, batch, max_length),
                "i2l_decode_sample: workspace too small");
    PersistentSampleArgs sa{};
    sa.top_k = top_k; sa.top_p = top_p;
    sa.do_sample = temperature > 0.f && (top_k > 0 || top_p > 0.0f);               // predictor.py:330
    sa.seed = seed; sa.offset = offset; sa.uniforms = uniforms; sa.probs_trace = probs_trace;
    return persistent_greedy(*d, reinterpret_cast<const char*>(packed) + lay.bf16_section,
                             reinterpret_cast<const float*>(packed), lay, enc, batch, start_id, end_id, max_length,
                             temperature, I2L_STOP_ALL_FINISHED_STICKY, tokens, lengths, steps_run,
                             reinterpret_cast<char*>(workspace) + gen, workspace_bytes - gen, s, &sa);
  }
  if (!no_persistent && wide_batch_supported(*d, batch) && d->vocab_size <= 512 && max_length > 0 && end_id >= 0 &&
      end_id < d->vocab_size && tokens != nullptr && lengths != nullptr && steps_run != nullptr) {
    // wide / multi-layer decoders (the reference's shipped 512 / 512 / 2): the sampling loop inside decode_wide.cu
    PackedDec lay = dec_layout(*d);
    size_t gen = carve(*d, batch, max_length, nullptr).bytes;
    I2L_REQUIRE(workspace_bytes >= gen + wide_workspace_bytes(*d, batch, max_length), "i2l_decode_sample: workspace too small");
    PersistentSampleArgs sa{};
    sa.top_k = top_k; sa.top_p = top_p;
    sa.do_sample = temperature > 0.f && (top_k > 0 || top_p > 0.0f);               // predictor.py:330
    sa.seed = seed; sa.offset = offset; sa.uniforms = uniforms; sa.probs_trace = probs_trace;
    return wide_greedy(*d, packed, lay, enc, batch, start_id, end_id, max_length, temperature, I2L_STOP_ALL_FINISHED_STICKY,
                       tokens, lengths, steps_run, reinterpret_cast<char*>(workspace) + gen, workspace_bytes - gen, s, &sa);
  }
  return run_loop(d, packed, enc, batch, start_id, end_id, max_length, temperature,
                  I2L_STOP_ALL_FINISHED_STICKY, true, top_k, top_p, seed, offset, uniforms, tokens, lengths,
                  steps_run, probs_trace, workspace, workspace_bytes, s);
}

extern "C" int i2l_decode_beam(const i2l_dec_desc* d_in, const void* packed, const float* enc, int32_t batch,
                               int32_t beam_size, int32_t start_id, int32_t end_id, int32_t max_length,
                               int64_t* out_tokens, int32_t* out_len, double* out_score, int32_t* trace_parent,
                               int32_t* trace_token, double* trace_score, int32_t* cand_token, float* cand_logp,
                               void* workspace, size_t workspace_bytes, void* stream) {
  const DescN dn_(d_in);
  const i2l_dec_desc* d = d_in ? &dn_.d : nullptr;
  I2L_TRY(check_common(d, packed));
  I2L_REQUIRE(beam_size >= 1 && beam_size <= I2L_MAX_BEAM, "i2l_decode_beam: beam_size must be in [1,%d]", I2L_MAX_BEAM);
  I2L_REQUIRE(beam_size <= d->vocab_size, "i2l_decode_beam: beam_size exceeds the vocabulary (torch.topk would raise)");
  I2L_REQUIRE(batch >= 0 && max_length >= 1 && out_tokens && out_len && out_score, "i2l_decode_beam: invalid arguments");
  I2L_REQUIRE(start_id >= 0 && start_id < d->vocab_size, "i2l_decode_beam: start token %d outside [0, %d) (nn.Embedding raises IndexError)",
              start_id, d->vocab_size);
  if (batch == 0) return I2L_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const int K = beam_size, R = batch * K, H = d->hidden_dim, V = d->vocab_size;
  PackedDec lay = dec_layout(*d);
  DecWs w = carve(*d, R, max_length, workspace);
  if (workspace_bytes < w.bytes) { set_error("i2l_decode_beam: workspace too small (%zu < %zu)", workspace_bytes, w.bytes); return I2L_ERR_WORKSPACE; }
  const float* pk = reinterpret_cast<const float*>(packed);
  int* trp = trace_parent ? trace_parent : w.tr_parent;
  int* trt = trace_token ? trace_token : w.tr_token;
  double* trs = trace_score ? trace_score : w.tr_score;
  I2L_REQUIRE((cand_token == nullptr) == (cand_logp == nullptr), "i2l_decode_beam: cand_token and cand_logp go together");
  const bool persistent = !dn_.streamed && d->precision == I2L_BF16 && persistent_beam_supported(*d, K) && end_id >= 0 &&
                          end_id < V;
  if (cand_token != nullptr && !persistent) {
    set_error("i2l_decode_beam: the candidate audit trail needs the persistent kernel (bf16 headline decoder, beam <= 8)");
    return I2L_ERR_UNSUPPORTED;
  }
  if (persistent) {
    // headline shape in bf16: the whole search runs inside one persistent cluster kernel
    I2L_TRY(make_gctx(*d, pk, lay, enc, batch, w.gates, s));
    I2L_TRY(persistent_beam(*d, reinterpret_cast<const char*>(packed) + lay.bf16_section, w.gates, batch, K, start_id,
                            end_id, max_length, w.bstate, w.score, trp, trt, trace_score, cand_token, cand_logp, s));
    beam_finalize_kernel<<<cdiv(batch, 128), 128, 0, s>>>(w.bstate, w.score, trp, trt, batch, K, max_length, end_id,
                                                          out_tokens, out_len, out_score);
    I2L_LAUNCH_OK();
    return I2L_OK;
  }
  beam_init_kernel<<<cdiv(R, 256), 256, 0, s>>>(w.bstate, w.score, w.tok_cur, w.live, batch, K, start_id);
  I2L_LAUNCH_OK();
  size_t n = (size_t)d->lstm_layers * R * H * sizeof(float);
  for (int i = 0; i < 2; ++i) {
    I2L_CUDA_OK(cudaMemsetAsync(w.h[i], 0, n, s));
    I2L_CUDA_OK(cudaMemsetAsync(w.c[i], 0, n, s));
  }
  // per-image context gates, repeated for the K beam rows of the image
  I2L_TRY(make_gctx(*d, pk, lay, enc, batch, w.gates, s));
  {
    size_t tot = (size_t)R * 4 * H;
    repeat_rows_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(w.gates, w.gctx, batch, K, 4 * H);
    I2L_LAUNCH_OK();
  }
  int cur = 0;
  KernelTimer kt("dec.beam_loop_general", s);
  for (int step = 0; step < max_length; ++step) {
    I2L_TRY(step_rows(*d, pk, lay, w, w.h[cur], w.c[cur], R, nullptr, s));
    beam_select_kernel<<<batch, 32 * K, 0, s>>>(w.logits, V, batch, K, step, end_id, w.bstate, w.score, w.tok_cur,
                                                w.live, w.parent, trp, trt, trs);
    I2L_LAUNCH_OK();
    size_t tot = (size_t)d->lstm_layers * R * H;
    beam_reorder_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(w.h[cur], w.c[cur], w.h[cur ^ 1], w.c[cur ^ 1],
                                                                     w.parent, R, K, H, d->lstm_layers);
    I2L_LAUNCH_OK();
    cur ^= 1;
  }
  beam_finalize_kernel<<<cdiv(batch, 128), 128, 0, s>>>(w.bstate, w.score, trp, trt, batch, K, max_length, end_id,
                                                        out_tokens, out_len, out_score);
  I2L_LAUNCH_OK();
  return I2L_OK;
}

// ---------------------------------------------------------------------------------------------
// Teacher-forced pass over a known token sequence (LSTMDecoder.forward in eval mode,
// decoder.py:100-195; both of its branches are the recurrence of decode_step -- attention over the
// single encoder vector is the identity).  Per step: one gate GEMM per layer (token term
// gathered from Gtok, context term from gctx) + the cell kernel, which also records the top layer's
// h at [b, t, :]; the vocabulary projection of ALL steps is ONE (B*T, H) x (H, V) GEMM at the end.
namespace i2l {
namespace {
struct FwdWs { DecWs w; int64_t* tok_t; float* h_seq; __nv_bfloat16* hb_seq; char* pws; size_t pws_bytes; size_t bytes; };
FwdWs carve_fwd(const i2l_dec_desc& d, int batch, int seq_len, void* ws) {
  FwdWs f{};
  f.w = carve(d, batch, 1, ws);
  Arena a(ws, (size_t)-1);
  a.off = f.w.bytes;
  const size_t n = (size_t)batch * seq_len;
  f.tok_t = a.take<int64_t>(n);
  f.h_seq = a.take<float>(n * d.hidden_dim);
  f.hb_seq = a.take<__nv_bfloat16>(n * d.hidden_dim);
  f.pws_bytes = (d.precision == I2L_BF16 && persistent_supported(d)) ? persistent_workspace_bytes(d, batch, seq_len) : 0;
  f.pws = a.take<char>(f.pws_bytes);
  f.bytes = align_up(a.off, 256);
  return f;
}
// (B,T) -> (T,B); ids outside [0,V) raise the device flag (nn.Embedding would raise IndexError) and read row 0
}  // namespace
}  // namespace i2l

extern "C" size_t i2l_dec_forward_workspace_bytes(const i2l_dec_desc* d_in, int32_t batch, int32_t seq_len) {
  const DescN dn_(d_in);
  const i2l_dec_desc* d = d_in ? &dn_.d : nullptr;
  if (!d || batch <= 0 || seq_len <= 0) return 0;
  return carve_fwd(*d, batch, seq_len, nullptr).bytes;
}

extern "C" int i2l_decoder_forward(const i2l_dec_desc* d_in, const void* packed, const float* enc,
                                   const int64_t* target, int32_t batch, int32_t seq_len, const float* h_in,
                                   const float* c_in, float* logits, float* h_out, float* c_out,
                                   int32_t* bad_token_flag, void* workspace, size_t workspace_bytes, void* stream) {
  const DescN dn_(d_in);
  const i2l_dec_desc* d = d_in ? &dn_.d : nullptr;
  I2L_TRY(check_common(d, packed));
  I2L_REQUIRE(batch >= 0 && seq_len >= 0, "i2l_decoder_forward: negative sizes");
  I2L_REQUIRE((h_in == nullptr) == (c_in == nullptr), "i2l_decoder_forward: h_in and c_in must both be given or both NULL");
  I2L_REQUIRE((h_out == nullptr) == (c_out == nullptr), "i2l_decoder_forward: h_out and c_out must both be given or both NULL");
  if (batch == 0 || seq_len == 0) return I2L_OK;
  I2L_REQUIRE(enc && target && logits, "i2l_decoder_forward: null argument");
  cudaStream_t s = (cudaStream_t)stream;
  const int B = batch, T = seq_len, H = d->hidden_dim, V = d->vocab_size, L = d->lstm_layers;
  PackedDec lay = dec_layout(*d);
  FwdWs f = carve_fwd(*d, B, T, workspace);
  if (workspace_bytes < f.bytes) { set_error("i2l_decoder_forward: workspace too small (%zu < %zu)", workspace_bytes, f.bytes); return I2L_ERR_WORKSPACE; }
  const DecWs& w = f.w;
  const float* pk = reinterpret_cast<const float*>(packed);
  const char* pb = reinterpret_cast<const char*>(packed);
  const size_t n_state = (size_t)L * B * H * sizeof(float);
  float *h = w.h[0], *c = w.c[0];
  if (h_in) {
    I2L_CUDA_OK(cudaMemcpyAsync(h, h_in, n_state, cudaMemcpyDeviceToDevice, s));
    I2L_CUDA_OK(cudaMemcpyAsync(c, c_in, n_state, cudaMemcpyDeviceToDevice, s));
  } else {                                                         // decoder.py:145-158
    I2L_CUDA_OK(cudaMemsetAsync(h, 0, n_state, s));
    I2L_CUDA_OK(cudaMemsetAsync(c, 0, n_state, s));
  }
  {
    size_t tot = (size_t)B * T;
    transpose_tokens_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(target, f.tok_t, B, T, V, bad_token_flag);
    I2L_LAUNCH_OK();
  }
  const bool tc = lay.g16 != 0;
  if (f.pws_bytes != 0 && h_in == nullptr && !dn_.streamed) {
    // headline decoder shape in bf16, zero initial state: all T steps inside the persistent cluster kernel
    float* ho = h_out ? h_out : nullptr;
    return persistent_forward(*d, reinterpret_cast<const char*>(packed) + lay.bf16_section, pk, lay, enc, f.tok_t, B, T,
                              logits, ho, c_out, f.pws, f.pws_bytes, s);
  }
  KernelTimer kt("dec.forward_teacher", s);
  if (tc) {
    I2L_TRY(make_gctx(*d, pk, lay, enc, B, w.gctx, s));       // fp32, as in the general decode loops
    size_t tot = (size_t)L * B * H;
    f32_to_bf16_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(h, w.hb, tot);
    I2L_LAUNCH_OK();
    StepBf16 st;
    I2L_TRY(make_step_bf16(*d, packed, lay, w, B, nullptr, &st));
    for (int t = 0; t < T; ++t) {
      for (int l = 0; l < L; ++l) {
        GemmBf16 g = st.gates[l];
        if (l == 0) g.tab_idx = f.tok_t + (size_t)t * B;
        I2L_TRY(gemm_bf16(g, s));
        const bool top = l == L - 1;
        I2L_TRY(lstm_cell_f32(w.gates, h + (size_t)l * B * H, c + (size_t)l * B * H, B, H, nullptr, s,
                              w.hb + (size_t)l * B * H, nullptr, top ? f.hb_seq + (size_t)t * H : nullptr,
                              (size_t)T * H));
      }
    }
    GemmBf16 g;
    g.M = B * T; g.N = V; g.C = logits; g.ldc = V; g.K1 = H;
    I2L_TRY(gemm_bf16_a_map(&g.tmA1, f.hb_seq, B * T, H, H));
    I2L_TRY(gemm_bf16_w_map(&g.tmW1, pb + lay.g16_out_w, V, H, H));
    g.bias = pk + lay.out_b;
    I2L_TRY(gemm_bf16(g, s));
  } else {
    I2L_TRY(make_gctx(*d, pk, lay, enc, B, w.gctx, s));
    for (int t = 0; t < T; ++t) {
      for (int l = 0; l < L; ++l) {
        GemmF32 g;
        g.M = B; g.N = 4 * H; g.C = w.gates; g.ldc = 4 * H;
        float* hl = h + (size_t)l * B * H;
        if (l == 0) {
          g.A1 = hl; g.lda1 = H; g.W1 = pk + lay.w_hh[0]; g.ldw1 = H; g.K1 = H;
          g.add_rows = w.gctx; g.ld_add = 4 * H;
          g.add_table = pk + lay.gtok; g.ld_tab = 4 * H; g.tab_idx = f.tok_t + (size_t)t * B;
        } else {
          g.A1 = h + (size_t)(l - 1) * B * H; g.lda1 = H; g.W1 = pk + lay.w_ih[l]; g.ldw1 = H; g.K1 = H;
          g.A2 = hl; g.lda2 = H; g.W2 = pk + lay.w_hh[l]; g.ldw2 = H; g.K2 = H;
          g.bias = pk + lay.bsum[l];
        }
        I2L_TRY(gemm_f32(g, s));
        const bool top = l == L - 1;
        I2L_TRY(lstm_cell_f32(w.gates, hl, c + (size_t)l * B * H, B, H, nullptr, s, nullptr,
                              top ? f.h_seq + (size_t)t * H : nullptr, nullptr, (size_t)T * H));
      }
    }
    GemmF32 g;
    g.M = B * T; g.N = V; g.C = logits; g.ldc = V;
    g.A1 = f.h_seq; g.lda1 = H; g.W1 = pk + lay.out_w; g.ldw1 = H; g.K1 = H; g.bias = pk + lay.out_b;
    I2L_TRY(gemm_f32(g, s));
  }
  if (h_out) {
    I2L_CUDA_OK(cudaMemcpyAsync(h_out, h, n_state, cudaMemcpyDeviceToDevice, s));
    I2L_CUDA_OK(cudaMemcpyAsync(c_out, c, n_state, cudaMemcpyDeviceToDevice, s));
  }
  return I2L_OK;
}

extern "C" size_t i2l_attention_workspace_bytes(int32_t hidden_dim, int32_t encoder_dim, int32_t batch,
                                                int32_t src_len) {
  (void)encoder_dim;
  return align_up((size_t)batch * hidden_dim * 4, 256) + align_up((size_t)batch * src_len * hidden_dim * 4, 256);
}

extern "C" int i2l_attention_fwd(int32_t H, int32_t E, const float* attn_w, const float* attn_b, const float* v_w,
                                 const float* hidden, const float* encoder_outputs, int32_t batch, int32_t src_len,
                                 float* context, void* workspace, size_t workspace_bytes, void* stream) {
  I2L_TRY(device_check());
  I2L_REQUIRE(H > 0 && E > 0 && batch >= 0 && src_len >= 1, "i2l_attention_fwd: invalid dimensions");
  I2L_REQUIRE(attn_w && attn_b && v_w && hidden && encoder_outputs && context, "i2l_attention_fwd: null argument");
  if (batch == 0) return I2L_OK;
  cudaStream_t s = (cudaStream_t)stream;
  if (src_len == 1) {
    // softmax over a single position is exactly 1.0 and 1.0 * e == e bit for bit
    I2L_CUDA_OK(cudaMemcpyAsync(context, encoder_outputs, (size_t)batch * E * 4, cudaMemcpyDeviceToDevice, s));
    return I2L_OK;
  }
  if (workspace_bytes < i2l_attention_workspace_bytes(H, E, batch, src_len) || !workspace) {
    set_error("i2l_attention_fwd: workspace too small");
    return I2L_ERR_WORKSPACE;
  }
  I2L_REQUIRE((size_t)src_len * 4 <= 48 * 1024, "i2l_attention_fwd: src_len too large");
  Arena a(workspace, workspace_bytes);
  float* p1 = a.take<float>((size_t)batch * H);
  float* p2 = a.take<float>((size_t)batch * src_len * H);
  GemmF32 g1;
  g1.M = batch; g1.N = H; g1.C = p1; g1.ldc = H; g1.A1 = hidden; g1.lda1 = H; g1.W1 = attn_w; g1.ldw1 = H + E;
  g1.K1 = H; g1.bias = attn_b;
  I2L_TRY(gemm_f32(g1, s));
  GemmF32 g2;
  g2.M = batch * src_len; g2.N = H; g2.C = p2; g2.ldc = H; g2.A1 = encoder_outputs; g2.lda1 = E;
  g2.W1 = attn_w + H; g2.ldw1 = H + E; g2.K1 = E;
  I2L_TRY(gemm_f32(g2, s));
  attention_combine_kernel<<<batch, 256, (size_t)src_len * 4, s>>>(p1, p2, v_w, encoder_outputs, src_len, H, E, context);
  I2L_LAUNCH_OK();
  return I2L_OK;
}

#ifdef I2L_DIAG   // diagnostics library only (make diag -> libi2l_b200_diag.so); the production library carries none of this
// debug aid (tools/debug_persistent.py): dump intermediates of one step of the persistent kernel
namespace i2l { int persistent_set_debug(float* buf); }
extern "C" int i2l_debug_set_buffer(float* buf) { return i2l::persistent_set_debug(buf); }
// test aid: (T,B,K,K) device buffers that receive every live beam's top-K (token, log-prob) of the
// next persistent beam calls (NULL, NULL switches the dump off)
namespace i2l { int persistent_beam_set_ts(long long* ts, int step); }
extern "C" int i2l_debug_set_beam_ts(long long* ts, int32_t step) { return i2l::persistent_beam_set_ts(ts, step); }
#endif
