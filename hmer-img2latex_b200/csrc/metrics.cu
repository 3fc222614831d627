// Sequence metrics of `img2latex evaluate` on the device: the integer parts of
// `levenshtein_distance` (training/metrics.py:49-94: the (rows+1) x (cols+1) DP table) and `bleu_n_score`
// (training/metrics.py:97-179: clipped n-gram matches, n = 1..4) for a batch of (prediction, target) id
// sequences -- the pure-Python loops that dominate evaluation once decoding runs at ~1 M images/s.
// One CTA of 5 warps per pair, both sequences staged in shared memory; warp 0 fills the edit-distance table, warps
// 1..4 count the matches of gram sizes 1..4 (a batch of 1024 pairs is otherwise only ~7 warps per SM, and every
// loop below is a chain of dependent shared-memory reads):
//  * edit distance: anti-diagonal wavefront.  Lane l owns 8 consecutive columns of a 256-column block and at
//    step t fills row t - l of them; D[r][c-1] arrives from lane l-1 by one shuffle per step, D[r-1][c-1] is the
//    value received one step earlier, D[r-1][c] stays in registers.  Wider targets run block after block with
//    the block's right-most column parked in shared memory.
//  * clipped matches: position i of the prediction counts iff its gram occurred fewer times before i than
//    it occurs in the target: sum_i [rank_i < count_t(g_i)] = sum_g min(count_p(g), count_t(g)).
// Outputs are integers (bit-exact); the host turns them into the reference's floats (metrics.py:87-94,
// 160-179) with the same libm calls.
#include "common.cuh"

namespace i2l {
namespace {

constexpr int kCPL = 8;                       // columns per lane
constexpr int kBlockCols = 32 * kCPL;
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ bool gram_eq(const int64_t* p, const int64_t (&g)[4], int n) {
  if (p[0] != g[0]) return false;
  for (int q = 1; q < n; ++q)
    if (p[q] != g[q]) return false;
  return true;
}

__global__ void __launch_bounds__(160) sequence_metrics_kernel(const int64_t* __restrict__ pred, int ldp,
                                                               const int32_t* __restrict__ plen,
                                                               const int64_t* __restrict__ tgt, int ldt,
                                                               const int32_t* __restrict__ tlen, int B, int max_n,
                                                               int32_t* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = blockIdx.x;
  const int m = min(max(plen[pair], 0), ldp), n = min(max(tlen[pair], 0), ldt);
  int64_t* a = reinterpret_cast<int64_t*>(smem);
  int64_t* b = a + ldp;
  int* bound = reinterpret_cast<int*>(b + ldt);
  for (int i = threadIdx.x; i < m; i += blockDim.x) a[i] = pred[(size_t)pair * ldp + i];
  for (int i = threadIdx.x; i < n; i += blockDim.x) b[i] = tgt[(size_t)pair * ldt + i];
  for (int r = threadIdx.x; r <= m; r += blockDim.x) bound[r] = r;            // D[r][0] = r
  __syncthreads();
  int32_t* o = out + (size_t)pair * 8;

  if (warp == 0) {
  // ---- edit distance (metrics.py:60-85)
  int dist = max(m, n);
  if (m > 0 && n > 0) {
    for (int c0 = 0; c0 < n; c0 += kBlockCols) {
      const int width = min(kBlockCols, n - c0);
      const int nlanes = (width + kCPL - 1) / kCPL;
      const int ncols = min(max(width - lane * kCPL, 0), kCPL);
      const bool park = c0 + width < n;                          // another block follows: park the last column
      int64_t bv[kCPL];
      int prev[kCPL];
#pragma unroll
      for (int j = 0; j < kCPL; ++j) {
        const int col = c0 + lane * kCPL + j;
        bv[j] = col < n ? b[col] : 0;
        prev[j] = col + 1;                                       // D[0][col + 1]
      }
      int diag_in = c0 + lane * kCPL;                            // D[0][first own column - 1]
      int last_out = 0;
      const int steps = m + nlanes - 1;
      for (int t = 1; t <= steps; ++t) {
        const int from_left = __shfl_up_sync(kFull, last_out, 1);
        const int r = t - lane;
        if (r >= 1 && r <= m && ncols > 0) {
          int left, diag;
          if (lane == 0) { left = bound[r]; diag = bound[r - 1]; }
          else { left = from_left; diag = diag_in; }
          const int next_diag = left;                            // D[r][first - 1] is the diagonal of row r + 1
          const int64_t av = a[r - 1];
#pragma unroll
          for (int j = 0; j < kCPL; ++j) {
            if (j < ncols) {
              const int up = prev[j];
              const int cur = av == bv[j] ? diag : 1 + min(min(up, left), diag);
              diag = up;
              left = cur;
              prev[j] = cur;
            }
          }
          last_out = left;
          diag_in = next_diag;
          if (park && lane == nlanes - 1) bound[r] = left;       // read by lane 0 only in the next block
        }
      }
      if (!park) dist = __shfl_sync(kFull, last_out, nlanes - 1);
      __syncwarp();
      if (park && lane == 0) bound[0] = c0 + width;              // D[0][c0 + width]
      __syncwarp();
    }
  }

  if (lane == 0) { o[0] = dist; o[5] = m; o[6] = n; o[7] = 0; }
  } else {
  // ---- clipped n-gram matches (metrics.py:125-158), gram size = warp index
    const int gs = warp;
    int cnt = 0;
    if (gs <= max_n) {
      const int Lg = m - gs + 1, Lt = n - gs + 1;
      if (m > 0 && n > 0 && Lg > 0 && Lt > 0) {
        for (int i = lane; i < Lg; i += 32) {
          int64_t g[4] = {a[i], 0, 0, 0};
          for (int q = 1; q < gs; ++q) g[q] = a[i + q];
          int rank = 0;
          for (int j = 0; j < i; ++j) rank += gram_eq(a + j, g, gs) ? 1 : 0;
          int ct = 0;
          for (int k = 0; k < Lt && ct <= rank; ++k) ct += gram_eq(b + k, g, gs) ? 1 : 0;
          cnt += rank < ct ? 1 : 0;
        }
      }
#pragma unroll
      for (int q = 16; q > 0; q >>= 1) cnt += __shfl_xor_sync(kFull, cnt, q);
    }
    if (lane == 0) o[gs] = cnt;
  }
}

// Per-row stable compaction: out[b] = [x for x in ids[b, :len[b]] if x not in drop] -- the id-level effect of
// `tokenizer.decode` (skips the four specials, data/tokenizer.py:177-189) followed by `tokenizer.encode`
// (cli.py:476-479), and of the padding filter on the targets (cli.py:471-474).  One warp per row.
struct DropSet { int64_t v[8]; int n; };
__global__ void __launch_bounds__(256) filter_ids_kernel(const int64_t* __restrict__ ids, int ld, const int32_t* __restrict__ len,
                                                         int B, DropSet drop, int64_t* __restrict__ out, int ld_out,
                                                         int32_t* __restrict__ out_len) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= B) return;
  const int n = len ? min(max(len[row], 0), ld) : ld;
  int kept = 0;
  for (int base = 0; base < n; base += 32) {
    const int i = base + lane;
    int64_t v = 0;
    bool keep = i < n;
    if (keep) {
      v = ids[(size_t)row * ld + i];
      for (int q = 0; q < drop.n; ++q) keep = keep && v != drop.v[q];
    }
    const unsigned mask = __ballot_sync(kFull, keep);
    const int pos = kept + __popc(mask & ((1u << lane) - 1));
    if (keep && pos < ld_out) out[(size_t)row * ld_out + pos] = v;
    kept += __popc(mask);
  }
  if (lane == 0) out_len[row] = min(kept, ld_out);
}

// Validation loss / accuracy of the teacher-forced pass (training/trainer.py:111-115, 517-529; training/metrics.py:
// 226-238): nn.CrossEntropyLoss(ignore_index = pad, reduction = "mean", label_smoothing = eps) over the logits rows
// and masked_accuracy (argmax == target on non-pad positions).  One warp per (sequence, position) row: the row is read
// ONCE (HBM-bound: V * 4 bytes per row) for the log-sum-exp, the target logit, the sum of logits and the argmax.
//   loss = (1 - eps) * mean_i(lse_i - x_i[t_i]) + (eps / V) * mean_i(V * lse_i - sum_c x_i[c])     over rows with t_i != pad
// A CTA handles 64 rows and writes ONE fp64 partial; one block reduces the partials in a fixed order (deterministic).
struct XentPartial { double nll, smooth; int tok, cor, bad, pad; };
constexpr int kXentRowsPerWarp = 8;
__global__ void __launch_bounds__(256) xent_rows_kernel(const float* __restrict__ logits, const int64_t* __restrict__ targets,
                                                        int N, int V, int64_t ignore_index, XentPartial* __restrict__ part) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double w_nll = 0.0, w_sm = 0.0;
  int w_tok = 0, w_cor = 0, w_bad = 0;
  for (int rr = 0; rr < kXentRowsPerWarp; ++rr) {
  const int row = (blockIdx.x * 8 + warp) * kXentRowsPerWarp + rr;
  if (row >= N) break;
  const float* x = logits + (size_t)row * V;
  const int64_t t = targets[row];
  float mx = -INFINITY, sum = 0.f, se = 0.f;
  int am = 0x7fffffff;
  float bm; int bi;
  if (V <= 512 && (V & 3) == 0) {
    // the whole row in registers: 128-bit loads, lane holds entries 4 lane + 128 k + {0..3}; read from HBM once
    float4 r[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = 4 * lane + 128 * k;
      r[k] = c < V ? *reinterpret_cast<const float4*>(x + c) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
      const float v4[4] = {r[k].x, r[k].y, r[k].z, r[k].w};
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (c < V) { if (v4[q] > mx) { mx = v4[q]; am = c + q; } sum += v4[q]; }
    }
    bm = mx; bi = am;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(kFull, bm, o);
      const int oi = __shfl_xor_sync(kFull, bi, o);
      if (ov > bm || (ov == bm && oi < bi)) { bm = ov; bi = oi; }      // torch.argmax: first index wins
    }
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (4 * lane + 128 * k < V) se += expf(r[k].x - bm) + expf(r[k].y - bm) + expf(r[k].z - bm) + expf(r[k].w - bm);
  } else {
    for (int c = lane; c < V; c += 32) {
      const float v = x[c];
      if (v > mx) { mx = v; am = c; }                     // first maximum of this lane's ascending indices
      sum += v;
    }
    bm = mx; bi = am;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(kFull, bm, o);
      const int oi = __shfl_xor_sync(kFull, bi, o);
      if (ov > bm || (ov == bm && oi < bi)) { bm = ov; bi = oi; }      // torch.argmax: first index wins
    }
    for (int c = lane; c < V; c += 32) se += expf(x[c] - bm);          // second pass hits L1
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { se += __shfl_xor_sync(kFull, se, o); sum += __shfl_xor_sync(kFull, sum, o); }
  {
    const bool valid = t != ignore_index;
    const bool inb = t >= 0 && t < V;
    const float lse = bm + logf(se);
    const float xt = (valid && inb) ? x[t] : 0.f;
    if (valid) {                                           // per-row terms in fp32 (as torch computes them), sums in fp64
      w_nll += (double)(lse - xt);
      w_sm += (double)((float)V * lse - sum);
      w_tok += 1; w_cor += ((int64_t)bi == t) ? 1 : 0; w_bad += inb ? 0 : 1;
    }
  }
  }
  // CTA partial in a fixed order: warp 0..7 (every lane of a warp holds the same sums)
  __shared__ XentPartial sp[8];
  if (lane == 0) sp[warp] = XentPartial{w_nll, w_sm, w_tok, w_cor, w_bad, 0};
  __syncthreads();
  if (threadIdx.x == 0) {
    XentPartial a = sp[0];
    for (int w = 1; w < 8; ++w) { a.nll += sp[w].nll; a.smooth += sp[w].smooth; a.tok += sp[w].tok; a.cor += sp[w].cor; a.bad += sp[w].bad; }
    part[blockIdx.x] = a;
  }
}

__global__ void __launch_bounds__(256) xent_reduce_kernel(const XentPartial* __restrict__ part, int n_part, int V, float eps,
                                                          float* __restrict__ loss, int32_t* __restrict__ counts) {
  __shared__ double s_nll[256], s_sm[256];
  __shared__ int s_tok[256], s_cor[256], s_bad[256];
  double a = 0.0, b = 0.0; int tk = 0, co = 0, bad = 0;
  for (int i = threadIdx.x; i < n_part; i += 256) {        // fixed assignment + fixed tree below: deterministic
    const XentPartial p = part[i];
    a += p.nll; b += p.smooth; tk += p.tok; co += p.cor; bad += p.bad;
  }
  s_nll[threadIdx.x] = a; s_sm[threadIdx.x] = b; s_tok[threadIdx.x] = tk; s_cor[threadIdx.x] = co; s_bad[threadIdx.x] = bad;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      s_nll[threadIdx.x] += s_nll[threadIdx.x + o]; s_sm[threadIdx.x] += s_sm[threadIdx.x + o];
      s_tok[threadIdx.x] += s_tok[threadIdx.x + o]; s_cor[threadIdx.x] += s_cor[threadIdx.x + o];
      s_bad[threadIdx.x] += s_bad[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double n = (double)s_tok[0];                     // 0 non-pad tokens: 0 / 0 = NaN, like torch
    *loss = (float)((1.0 - (double)eps) * (s_nll[0] / n) + ((double)eps / V) * (s_sm[0] / n));
    counts[0] = s_cor[0]; counts[1] = s_tok[0]; counts[2] = s_bad[0]; counts[3] = 0;
  }
}

}  // namespace
}  // namespace i2l

using namespace i2l;

static int xent_ctas(int rows) { return cdiv(rows, 8 * kXentRowsPerWarp); }
extern "C" size_t i2l_xent_workspace_bytes(int32_t rows) {
  return rows > 0 ? align_up((size_t)xent_ctas(rows) * sizeof(XentPartial), 256) + 256 : 256;
}

extern "C" int i2l_xent_metrics(const float* logits, const int64_t* targets, int32_t rows, int32_t vocab, int64_t ignore_index,
                                float label_smoothing, float* loss, int32_t* counts, void* workspace, size_t workspace_bytes,
                                void* stream) {
  I2L_TRY(device_check());
  I2L_REQUIRE(rows >= 0 && vocab >= 1, "i2l_xent_metrics: invalid sizes");
  I2L_REQUIRE(label_smoothing >= 0.f && label_smoothing <= 1.f, "i2l_xent_metrics: label_smoothing must be in [0,1]");
  I2L_REQUIRE(loss && counts, "i2l_xent_metrics: null output");
  I2L_REQUIRE(rows == 0 || (logits && targets), "i2l_xent_metrics: null argument");
  I2L_REQUIRE(workspace != nullptr, "i2l_xent_metrics: null workspace");
  if (workspace_bytes < i2l_xent_workspace_bytes(rows)) { set_error("i2l_xent_metrics: workspace too small"); return I2L_ERR_WORKSPACE; }
  cudaStream_t s = (cudaStream_t)stream;
  XentPartial* part = reinterpret_cast<XentPartial*>(workspace);
  const int n_cta = rows > 0 ? xent_ctas(rows) : 0;
  KernelTimer kt("eval.xent_metrics", s);
  if (rows > 0) {
    xent_rows_kernel<<<n_cta, 256, 0, s>>>(logits, targets, rows, vocab, ignore_index, part);
    I2L_LAUNCH_OK();
  }
  xent_reduce_kernel<<<1, 256, 0, s>>>(part, n_cta, vocab, label_smoothing, loss, counts);
  I2L_LAUNCH_OK();
  return I2L_OK;
}

extern "C" int i2l_filter_ids(const int64_t* ids, int32_t ld, const int32_t* len, int32_t batch, const int64_t* drop_host,
                              int32_t n_drop, int64_t* out, int32_t ld_out, int32_t* out_len, void* stream) {
  I2L_TRY(device_check());
  I2L_REQUIRE(batch >= 0 && ld >= 0 && ld_out >= 0 && n_drop >= 0 && n_drop <= 8, "i2l_filter_ids: invalid sizes (at most 8 dropped ids)");
  if (batch == 0) return I2L_OK;
  I2L_REQUIRE((ids || ld == 0) && (out || ld_out == 0) && out_len && (drop_host || n_drop == 0), "i2l_filter_ids: null argument");
  DropSet d{};
  d.n = n_drop;
  for (int i = 0; i < n_drop; ++i) d.v[i] = drop_host[i];
  filter_ids_kernel<<<cdiv(batch, 8), 256, 0, (cudaStream_t)stream>>>(ids, ld, len, batch, d, out, ld_out, out_len);
  I2L_LAUNCH_OK();
  return I2L_OK;
}

extern "C" int i2l_sequence_metrics(const int64_t* pred, int32_t ld_pred, const int32_t* pred_len, const int64_t* tgt,
                                    int32_t ld_tgt, const int32_t* tgt_len, int32_t batch, int32_t max_n,
                                    int32_t* out, void* stream) {
  I2L_TRY(device_check());
  I2L_REQUIRE(batch >= 0 && ld_pred >= 0 && ld_tgt >= 0, "i2l_sequence_metrics: negative sizes");
  I2L_REQUIRE(max_n >= 1 && max_n <= 4, "i2l_sequence_metrics: max_n must be in [1,4]");
  if (batch == 0) return I2L_OK;
  I2L_REQUIRE(pred_len && tgt_len && out && (pred || ld_pred == 0) && (tgt || ld_tgt == 0),
              "i2l_sequence_metrics: null argument");
  const size_t smem = align_up((size_t)8 * ((size_t)ld_pred + ld_tgt) + 4 * ((size_t)ld_pred + 1), 16);
  if (smem > 200 * 1024) {
    set_error("i2l_sequence_metrics: sequences too long for the on-chip tables (%d + %d tokens)", ld_pred, ld_tgt);
    return I2L_ERR_UNSUPPORTED;
  }
  I2L_CUDA_OK(cudaFuncSetAttribute(sequence_metrics_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaStream_t s = (cudaStream_t)stream;
  KernelTimer kt("eval.sequence_metrics", s);
  sequence_metrics_kernel<<<batch, 5 * 32, smem, s>>>(pred, ld_pred, pred_len, tgt, ld_tgt, tgt_len, batch, max_n, out);
  I2L_LAUNCH_OK();
  return I2L_OK;
}
