// Device-side bookkeeping of the decode loops (shared by the general fp32 path and
// the persistent bf16 path): LSTM cell update, token selection (argmax / filtered
// sampling / beam expansion), EOS masking and loop-exit flags.  No host sync.
#pragma once
#include "common.cuh"

namespace i2l {

// Loop-global device state (one per decode call, lives in the workspace).
struct LoopState {
  int done;            // != 0 once the reference loop would have executed `break`
  int steps_run;       // loop iterations the reference executes
  int finished_count;  // sticky rule: rows that have emitted END
  int ticket;          // block completion ticket of the current select kernel
  int end_count;       // ALL_END_SAME_STEP rule: rows emitting END in the current step
};

// Per-image state of the beam search (seq2seq.py:234-298), shared by the general and the
// persistent beam kernels and read by the finalize kernel.
struct BeamState {
  int alive;                    // loop still running for this image
  int nbeams;                   // len(beams)
  int has_completed;
  int best_step, best_slot;     // best entry of `completed` (first-wins max)
  double best_score;
  int last_step;                // last executed iteration
};
#define I2L_MAX_BEAM 16

struct PackedDec {     // offsets (in floats) into the packed decoder buffer, fp32 section
  size_t emb, gtok, w_ih0, bsum[I2L_MAX_LSTM_LAYERS], w_hh[I2L_MAX_LSTM_LAYERS],
      w_ih[I2L_MAX_LSTM_LAYERS], out_w, out_b, end_f32;
  size_t bf16_section;  // byte offset of the persistent-kernel section (0 if absent)
  // bf16 row-major copies of the per-step GEMM weights for the general path (byte offsets; g16 == 0 if absent):
  // W_hh[l] (4H,H), W_ih[l >= 1] (4H,H), W_out (V,H) -- operands of gemm_bf16.cu
  size_t g16, g16_w_hh[I2L_MAX_LSTM_LAYERS], g16_w_ih[I2L_MAX_LSTM_LAYERS], g16_out_w;
  size_t g16_w_ctx;     // W_ih0[:, E:2E] (4H,E): the per-sequence constant gate term gctx = enc W_ctx^T + b_ih0 + b_hh0
  // H % 32 == 0: the same gate weights with the rows of every 128-row tile ordered [i | f | g | o] x 32 units, for the
  // gate GEMM with the LSTM cell fused into its epilogue (gemm_bf16.cu, GemmBf16::cell_*); g16c == 0 if absent
  size_t g16c, g16c_w_hh[I2L_MAX_LSTM_LAYERS], g16c_w_ih[I2L_MAX_LSTM_LAYERS];
  size_t total_bytes;
};
// precision == I2L_BF16 and TMA-addressable rows (H % 8 == 0, E % 8 == 0): the general loops run their GEMMs on tcgen05
bool general_bf16_supported(const i2l_dec_desc& d);
// gctx (n,4H) = enc (n,E) W_ih0[:, E:2E]^T + b_ih0 + b_hh0 on the tensor cores; encb: scratch for n*E bf16
int make_gctx_bf16(const i2l_dec_desc& d, const void* packed, const PackedDec& lay, const float* enc, int n, float* gctx,
                   __nv_bfloat16* encb, cudaStream_t s);
PackedDec dec_layout(const i2l_dec_desc& d);

// hb: optional bf16 copy of the new h (A operand of the next bf16 GEMM); h_seq / hb_seq: optional extra copies
// with row stride seq_ld (the (B,T,H) hidden-state record of the teacher-forced forward)
int lstm_cell_f32(const float* gates, float* h, float* c, int rows, int H, const int* skip_flag,
                  cudaStream_t s, __nv_bfloat16* hb = nullptr, float* h_seq = nullptr,
                  __nv_bfloat16* hb_seq = nullptr, size_t seq_ld = 0);

// persistent bf16 greedy decode (decode_persistent.cu); returns I2L_ERR_UNSUPPORTED for
// shapes it does not cover so that the caller can take the general path.
bool persistent_supported(const i2l_dec_desc& d);
size_t persistent_packed_bytes(const i2l_dec_desc& d);
int persistent_pack(const i2l_dec_desc& d, const i2l_dec_params& p, const float* gtok_f32, void* section,
                    cudaStream_t s);
size_t persistent_workspace_bytes(const i2l_dec_desc& d, int rows, int max_length);
// sample != nullptr: the same kernel in sampling mode (Predictor.predict_batch, training/predictor.py:295-335): warp-level
// temperature / top-k / top-p selection instead of the argmax, one warp of the cluster per sequence.
struct PersistentSampleArgs {
  int top_k; float top_p; int do_sample;
  unsigned long long seed, offset;
  const float* uniforms;      // (T,B) or null
  float* probs_trace;         // (T,B,V) or null
};
int persistent_greedy(const i2l_dec_desc& d, const void* section, const float* packed_f32,
                      const PackedDec& lay, const float* enc, int batch, int start_id, int end_id,
                      int max_length, float temperature, int stop_rule, int64_t* tokens, int32_t* lengths,
                      int32_t* steps_run, void* ws, size_t ws_bytes, cudaStream_t s,
                      const PersistentSampleArgs* sample = nullptr);

// teacher-forced pass on the same kernel (mode 2): tok_t (T,B) given tokens, logits (B,T,V), zero initial state
int persistent_forward(const i2l_dec_desc& d, const void* section, const float* packed_f32, const PackedDec& lay,
                       const float* enc, const int64_t* tok_t, int batch, int seq_len, float* logits, float* h_out,
                       float* c_out, void* ws, size_t ws_bytes, cudaStream_t s);

// persistent bf16 beam search (decode_persistent_beam.cu): same resident-weight cluster kernel with
// log-softmax + top-K, per-image candidate merge, back-pointers and state reorder on the device.
bool persistent_beam_supported(const i2l_dec_desc& d, int beam_size);
int persistent_beam(const i2l_dec_desc& d, const void* section, const float* gctx_img, int batch, int beam_size,
                    int start_id, int end_id, int max_length, BeamState* bstate, double* score, int* tr_parent,
                    int* tr_token, double* tr_score, int* cand_tok, float* cand_logp, cudaStream_t s);

// persistent whole-GPU greedy loop for the decoders the cluster kernel does not cover (decode_wide.cu): bf16, any
// L <= 4, H % 64 == 0 -- the reference's shipped 512 / 512 / 2 and 1024 / 1024 / 3 decoders
bool wide_supported(const i2l_dec_desc& d);
bool wide_batch_supported(const i2l_dec_desc& d, int rows);
size_t wide_workspace_bytes(const i2l_dec_desc& d, int rows, int max_length);
int wide_greedy(const i2l_dec_desc& d, const void* packed, const PackedDec& lay, const float* enc, int batch, int start_id,
                int end_id, int max_length, float temperature, int stop_rule, int64_t* tokens, int32_t* lengths,
                int32_t* steps_run, void* ws, size_t ws_bytes, cudaStream_t s, const PersistentSampleArgs* sample = nullptr);
int wide_aborted(const void* ws, const i2l_dec_desc& d, int rows, int max_length, int* out);

}  // namespace i2l
