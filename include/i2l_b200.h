/*
 * i2l_b200.h -- C-ABI of the B200-native (sm_100a) img2latex inference hot path.
 *
 * This is the drop-in boundary for the reference's batched inference path
 * (Jeremy-Cleland/hmer-img2latex).  The reference has no FFI of its own: its
 * boundary is the Python nn.Module surface.  Each entry point below names the
 * reference interface it replaces (file:line relative to the reference root).
 * The Python host side (hmer-img2latex_b200/) binds these with ctypes and keeps
 * the reference's class / method signatures; see INTEGRATION.md for the stub a
 * reference maintainer would add.
 *
 * Conventions
 *  - plain pointers + sizes only; all tensor pointers are DEVICE pointers unless
 *    a parameter is documented as host; `stream` is a cudaStream_t passed as void*.
 *  - the caller owns every buffer (weights, packed weights, workspace, outputs);
 *    the library allocates nothing on the hot path and never synchronises the
 *    host (all launches are asynchronous on `stream`).
 *  - return value: 0 = I2L_OK, negative = i2l_status; a message for the calling
 *    thread is available from i2l_last_error().  Nothing throws across the ABI.
 *  - CUDA only, sm_100 only: no CPU fallback.  i2l_device_check() != 0 => every
 *    compute entry point fails with I2L_ERR_NO_DEVICE.
 *  - weights arrive in the reference's own state_dict layouts (fp32) and are
 *    re-laid-out once by the *_pack calls into a caller-owned packed buffer.
 */
#ifndef I2L_B200_H_
#define I2L_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define I2L_MAX_CONV 8
#define I2L_MAX_LSTM_LAYERS 8
#define I2L_MAX_RESNET_CONVS 160

typedef enum {
  I2L_OK = 0,
  I2L_ERR_INVALID = -1,     /* bad argument / unsupported shape combination        */
  I2L_ERR_CUDA = -2,        /* a CUDA runtime / driver call failed                 */
  I2L_ERR_NO_DEVICE = -3,   /* no sm_100 device is current                         */
  I2L_ERR_UNSUPPORTED = -4, /* valid request the library does not implement        */
  I2L_ERR_WORKSPACE = -5    /* workspace / packed buffer too small                 */
} i2l_status;

/* I2L_BF16_STREAMED (decoder descriptors only): bf16 tensor-core arithmetic on the stream-ordered path (one group
 * of launches per step) even where a persistent cluster kernel exists -- same packed layout and workspace sizes
 * as I2L_BF16, so one packed model serves both; the A/B reference for the persistent kernels. */
typedef enum { I2L_FP32 = 0, I2L_BF16 = 1, I2L_BF16_STREAMED = 2 } i2l_precision;

/* Loop-exit rules of the reference's three decode loops. */
typedef enum {
  I2L_STOP_NONE = 0,
  I2L_STOP_ALL_END_SAME_STEP = 1,   /* Seq2SeqModel._greedy_search, model/seq2seq.py:220      */
  I2L_STOP_ALL_FINISHED_STICKY = 2  /* Predictor.predict_batch, training/predictor.py:343-347 */
} i2l_stop_rule;

const char* i2l_version(void);
const char* i2l_last_error(void);
/* 0 when the current CUDA device is compute capability 10.x, else I2L_ERR_NO_DEVICE. */
int i2l_device_check(void);

/* ------------------------------------------------------------------------- */
/* CNN encoder -- replaces CNNEncoder.forward, model/encoder.py:111-129        */
/* (layer stack built at model/encoder.py:74-107).                             */
/* ------------------------------------------------------------------------- */
typedef struct {
  int32_t img_height, img_width, channels;
  int32_t n_conv;                 /* len(conv_filters)                          */
  int32_t filters[I2L_MAX_CONV];  /* conv_filters                               */
  int32_t kernel_size;            /* odd; padding "same" = kernel_size/2        */
  int32_t pool_size;
  int32_t embedding_dim;
  int32_t precision;              /* i2l_precision                              */
} i2l_cnn_desc;

typedef struct {                         /* state_dict tensors, fp32, device     */
  const float* conv_w[I2L_MAX_CONV];     /* cnn_layers.{0,3,6..}.weight (Co,Ci,k,k) */
  const float* conv_b[I2L_MAX_CONV];     /* cnn_layers.{0,3,6..}.bias              */
  const float* fc_w;                     /* embedding_layer.weight (E, C*H*W) NCHW flatten order */
  const float* fc_b;                     /* embedding_layer.bias                   */
} i2l_cnn_params;

/* 1 when this configuration runs on the tcgen05 implicit-GEMM kernels (precision == I2L_BF16, conv filters
 * 32/64/128, 3x3, pool 2, embedding 256, channels 1 or 3, height % 64 == 0, width % 32 == 0 -- the reference's
 * default 1x64x800, model/encoder.py:50-64, and the 3x64x320 benchmark shape among them), 0 when it runs on the
 * fp32 CUDA-core kernels.  bf16 / uint8 image tensors are accepted on the tensor-core path only. */
int i2l_cnn_tensor_core_path(const i2l_cnn_desc* d);
size_t i2l_cnn_packed_bytes(const i2l_cnn_desc* d);
int i2l_cnn_pack(const i2l_cnn_desc* d, const i2l_cnn_params* p, void* packed, size_t packed_bytes,
                 void* stream);
size_t i2l_cnn_workspace_bytes(const i2l_cnn_desc* d, int32_t batch);
/* x: (B,C,H,W) fp32 NCHW;  out: (B,E) fp32. */
int i2l_cnn_encoder_fwd(const i2l_cnn_desc* d, const void* packed, const float* x, int32_t batch,
                        float* out, void* workspace, size_t workspace_bytes, void* stream);

/* Same call with the input element type stated: I2L_IN_F32 (the reference's tensor dtype)
 * or I2L_IN_BF16 (precision == I2L_BF16 only: halves the host->device and HBM bytes of the
 * image; the conv1 operand is bf16 either way).  SURVEY 8d: "fp32 master -> bf16 for the bf16 runs". */
typedef enum { I2L_IN_F32 = 0, I2L_IN_BF16 = 1, I2L_IN_U8 = 2 /* i2l_cnn_encoder_fwd_u8 only */ } i2l_input_dtype;
int i2l_cnn_encoder_fwd_in(const i2l_cnn_desc* d, const void* packed, const void* x, int32_t x_dtype,
                           int32_t batch, float* out, void* workspace, size_t workspace_bytes,
                           void* stream);

/* Device-side pixel normalisation -- replaces the arithmetic of Predictor._prepare_image /
 * load_image (training/predictor.py:441-446: x/255*2-1; data/utils.py:68-80: x/255 then
 * *2-1 (1 channel) or (x-mean)/std (RGB)), so that raw uint8 pixels cross PCIe instead of
 * fp32 tensors.  src: uint8, layout NCHW (0) or NHWC (1); dst: (B,C,H,W) fp32 or bf16
 * (i2l_input_dtype).  mode I2L_NORM_PM1: y = x/255*2-1; I2L_NORM_MEANSTD: y = (x/255-mean[c])/std[c]
 * (mean/std: HOST pointers to `channels` floats).  fp32 results are bit-identical to the
 * reference's float32 arithmetic. */
typedef enum { I2L_NORM_PM1 = 0, I2L_NORM_MEANSTD = 1 } i2l_norm_mode;
int i2l_normalize_u8(const uint8_t* src, int32_t src_layout, int32_t batch, int32_t channels,
                     int32_t height, int32_t width, int32_t mode, const float* mean, const float* stdev,
                     void* dst, int32_t dst_dtype, void* stream);

/* CNNEncoder.forward on RAW uint8 pixels (B,C,H,W) NCHW: the normalisation above is fused into the
 * first convolution (the im2col operand is built from the pixels as a[c]*x + b[c]; out-of-image taps
 * are zero in normalised space like the reference's padding), so the image crosses PCIe and HBM once,
 * at one byte per pixel.  precision == I2L_BF16 and the tcgen05 shape only (else I2L_ERR_UNSUPPORTED:
 * call i2l_normalize_u8 + i2l_cnn_encoder_fwd).  mean / stdev: HOST pointers (I2L_NORM_MEANSTD). */
int i2l_cnn_encoder_fwd_u8(const i2l_cnn_desc* d, const void* packed, const uint8_t* x, int32_t norm_mode,
                           const float* mean, const float* stdev, int32_t batch, float* out, void* workspace,
                           size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------- */
/* Image geometry on the device -- replaces the PIL calls of load_image           */
/* (data/utils.py:37-48) and Predictor._prepare_image (training/predictor.py:     */
/* 427-436) for a RAGGED batch of uint8 images: optional convert("L"), Pillow     */
/* resize (8-bit resampler, bit-exact), then ResizeWithAspectRatio's white right  */
/* padding / centre crop (data/transforms.py:26-56).  SURVEY 8f-1.                */
/* ------------------------------------------------------------------------- */
typedef struct {
  int64_t src_offset;      /* byte offset of the image inside the packed source buffer            */
  int32_t height, width;   /* source size; pixels are HWC interleaved uint8 (np.array(PIL image)) */
} i2l_image_desc;
typedef enum { I2L_FILTER_LANCZOS = 0, I2L_FILTER_BICUBIC = 1 } i2l_resize_filter;
typedef enum {
  I2L_RESIZE_ASPECT_PAD_CROP = 0, /* ResizeWithAspectRatio.__call__, data/transforms.py:26-56: new_w = round(Ht*w/h),
                                     resize to (new_w,Ht), pad right with Image.new(mode,size,255) (L: white; RGB: (255,0,0), as Pillow reads the int) or centre-crop to Wt; h == 0 => all padding */
  I2L_RESIZE_STRETCH = 1          /* image.resize((Wt,Ht)), training/predictor.py:436 (Pillow's default filter = BICUBIC) */
} i2l_resize_mode;

/* HOST side, no CUDA: the plan holds per-image geometry and the fixed-point filter weights, computed in
 * double precision exactly like Pillow's precompute_coeffs / normalize_coeffs_8bpc.  The caller copies the
 * plan to the device next to the pixels (one H2D), keeps the host copy for the launch call.
 * Errors: an image whose resized width would be 0 (Pillow: ValueError) => I2L_ERR_INVALID. */
size_t i2l_resize_plan_bytes(const i2l_image_desc* imgs, int32_t n, int32_t src_channels, int32_t to_gray,
                             int32_t target_h, int32_t target_w, int32_t filter, int32_t mode);   /* 0 on error */
int i2l_resize_plan_build(const i2l_image_desc* imgs, int32_t n, int32_t src_channels, int32_t to_gray,
                          int32_t target_h, int32_t target_w, int32_t filter, int32_t mode, void* plan_host,
                          size_t plan_bytes);
/* HOST: gather n separately allocated HWC uint8 images (images[i] -> descs[i].height x width x channels bytes) into
 * the packed source buffer at descs[i].src_offset; the copies are spread over a few host threads. */
int i2l_pack_images(const void* const* images, const i2l_image_desc* descs, int32_t n, int32_t channels, void* dst_host);
size_t i2l_resize_workspace_bytes(const void* plan_host);
/* src: packed uint8 source images (device); dst: (n, C_out, target_h, target_w) uint8 NCHW (device), the input
 * layout of i2l_normalize_u8 / i2l_cnn_encoder_fwd_u8.  Two launches for the whole batch, no host sync. */
int i2l_resize_pad_u8(const uint8_t* src, const void* plan_host, const void* plan_dev, uint8_t* dst,
                      void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------- */
/* ResNet encoder -- replaces ResNetEncoder.forward, model/encoder.py:231-249   */
/* (torchvision trunk minus fc, model/encoder.py:184-199; eval-mode BN folded). */
/* ------------------------------------------------------------------------- */
typedef struct {
  int32_t depth;                 /* 18, 34, 50, 101, 152  (model_name)          */
  int32_t img_height;            /* width is a call-time argument (width buckets) */
  int32_t embedding_dim;
  int32_t precision;
} i2l_resnet_desc;

typedef struct {
  int32_t n_convs;               /* convs in canonical order: stem, then per block conv1,conv2[,conv3][,downsample] */
  const float* conv_w[I2L_MAX_RESNET_CONVS];
  const float* bn_weight[I2L_MAX_RESNET_CONVS];
  const float* bn_bias[I2L_MAX_RESNET_CONVS];
  const float* bn_mean[I2L_MAX_RESNET_CONVS];
  const float* bn_var[I2L_MAX_RESNET_CONVS];
  const float* fc_w;             /* embedding_layer.weight (E, 512|2048)         */
  const float* fc_b;
} i2l_resnet_params;

int32_t i2l_resnet_num_convs(int32_t depth);   /* <0 if depth is invalid */
size_t i2l_resnet_packed_bytes(const i2l_resnet_desc* d);
int i2l_resnet_pack(const i2l_resnet_desc* d, const i2l_resnet_params* p, void* packed,
                    size_t packed_bytes, void* stream);
size_t i2l_resnet_workspace_bytes(const i2l_resnet_desc* d, int32_t batch, int32_t img_width);
/* x: (B,3,H,W) fp32 NCHW; out: (B,E) fp32. */
int i2l_resnet_encoder_fwd(const i2l_resnet_desc* d, const void* packed, const float* x, int32_t batch,
                           int32_t img_width, float* out, void* workspace, size_t workspace_bytes,
                           void* stream);

/* ------------------------------------------------------------------------- */
/* Attention -- replaces Attention.forward, model/decoder.py:312-343            */
/* hidden (B,H), encoder_outputs (B,L,E) -> context (B,E).  General L; the     */
/* decode path always has L == 1 where the result is encoder_outputs itself.   */
/* ------------------------------------------------------------------------- */
size_t i2l_attention_workspace_bytes(int32_t hidden_dim, int32_t encoder_dim, int32_t batch, int32_t src_len);
int i2l_attention_fwd(int32_t hidden_dim, int32_t encoder_dim, const float* attn_w /*(H,H+E)*/,
                      const float* attn_b /*(H)*/, const float* v_w /*(1,H)*/, const float* hidden,
                      const float* encoder_outputs, int32_t batch, int32_t src_len, float* context,
                      void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------- */
/* LSTM decoder                                                                */
/* ------------------------------------------------------------------------- */
typedef struct {
  int32_t vocab_size, embedding_dim, hidden_dim, lstm_layers;
  int32_t attention;             /* LSTMDecoder(attention=...) -- src_len is 1 on this path */
  int32_t precision;
} i2l_dec_desc;

typedef struct {                              /* state_dict tensors, fp32, device */
  const float* embedding;                     /* decoder.embedding.weight (V,E)           */
  const float* w_ih[I2L_MAX_LSTM_LAYERS];     /* decoder.lstm.weight_ih_l{k} (4H, 2E | H)   */
  const float* w_hh[I2L_MAX_LSTM_LAYERS];     /* decoder.lstm.weight_hh_l{k} (4H, H)        */
  const float* b_ih[I2L_MAX_LSTM_LAYERS];
  const float* b_hh[I2L_MAX_LSTM_LAYERS];
  const float* out_w;                         /* decoder.output_layer.weight (V,H)        */
  const float* out_b;
} i2l_dec_params;

size_t i2l_dec_packed_bytes(const i2l_dec_desc* d);
int i2l_dec_pack(const i2l_dec_desc* d, const i2l_dec_params* p, void* packed, size_t packed_bytes,
                 void* stream);
/* rows = batch (greedy / sample / step) or batch*beam (beam). */
size_t i2l_dec_workspace_bytes(const i2l_dec_desc* d, int32_t rows, int32_t max_length);

/* replaces LSTMDecoder.decode_step, model/decoder.py:197-284.
 * enc (B,E) fp32; tok (B) int64; h_in/c_in (L,B,H) fp32 or NULL (= zeros, decoder.py:253-266);
 * logits (B,V) fp32; h_out/c_out (L,B,H) fp32 (may not alias h_in/c_in).
 * bad_token_flag: optional device int32 set to 1 when an id lies outside [0, vocab_size) (such
 * ids are read as 0 by the table gather, never out of bounds; nn.Embedding raises IndexError). */
int i2l_decode_step(const i2l_dec_desc* d, const void* packed, const float* enc, const int64_t* tok,
                    int32_t batch, const float* h_in, const float* c_in, float* logits, float* h_out,
                    float* c_out, int32_t* bad_token_flag, void* workspace, size_t workspace_bytes, void* stream);

/* replaces LSTMDecoder.forward in eval mode (dropout = identity), model/decoder.py:100-195:
 * the teacher-forced pass over a known token sequence (validation loss / accuracy path,
 * training/trainer.py:508-533).  enc (B,E) fp32; target (B,T) int64;
 * h_in/c_in (L,B,H) or NULL (= zeros, decoder.py:145-158); logits (B,T,V) fp32;
 * h_out/c_out (L,B,H) optional (both or neither).  bad_token_flag: optional device int32 set
 * to 1 when a token id lies outside [0,V) (nn.Embedding raises IndexError there; the kernel
 * reads row 0 instead and the host wrapper raises). */
size_t i2l_dec_forward_workspace_bytes(const i2l_dec_desc* d, int32_t batch, int32_t seq_len);
int i2l_decoder_forward(const i2l_dec_desc* d, const void* packed, const float* enc, const int64_t* target,
                        int32_t batch, int32_t seq_len, const float* h_in, const float* c_in, float* logits,
                        float* h_out, float* c_out, int32_t* bad_token_flag, void* workspace,
                        size_t workspace_bytes, void* stream);

/* replaces Seq2SeqModel._greedy_search, model/seq2seq.py:192-232 (stop rule
 * ALL_END_SAME_STEP) -- the whole loop runs on the device, no host sync per token.
 * tokens (B, max_length+1) int64, column 0 = start; lengths (B) int32 = position
 * of the first END (or steps_run+1 when none); steps_run: device int32 scalar =
 * loop iterations the reference would have executed. */
int i2l_decode_greedy(const i2l_dec_desc* d, const void* packed, const float* enc, int32_t batch,
                      int32_t start_id, int32_t end_id, int32_t max_length, float temperature,
                      int32_t stop_rule, int64_t* tokens, int32_t* lengths, int32_t* steps_run,
                      void* workspace, size_t workspace_bytes, void* stream);

/* replaces the batched loop of Predictor.predict_batch, training/predictor.py:283-347
 * (temperature / top-k / top-p filter 295-327, draw 330-335, sticky finished 343-347).
 * uniforms: optional (max_length,B) fp32 in [0,1) used for the inverse-CDF draw;
 * NULL => Philox4x32-10 keyed by (seed, offset + step*B + row).  temperature == 0 is refused (I2L_ERR_INVALID):
 * the reference divides the logits by it.
 * probs_trace: optional (max_length,B,V) fp32 dump of the filtered distribution. */
int i2l_decode_sample(const i2l_dec_desc* d, const void* packed, const float* enc, int32_t batch,
                      int32_t start_id, int32_t end_id, int32_t max_length, float temperature,
                      int32_t top_k, float top_p, uint64_t seed, uint64_t offset, const float* uniforms,
                      int64_t* tokens, int32_t* lengths, int32_t* steps_run, float* probs_trace,
                      void* workspace, size_t workspace_bytes, void* stream);

/* replaces Seq2SeqModel._beam_search, model/seq2seq.py:234-298, run independently
 * per image (the reference handles B==1 only).  out_tokens (B,max_length) int64:
 * best sequence, START stripped, cut at END, padded with -1; out_len (B) int32;
 * out_score (B) fp64.  Optional traces (max_length,B,K): parent slot / token /
 * score of every kept beam per step (-1 / NaN where a slot is empty).
 * cand_token / cand_logp: optional (max_length,B,K,K) int32 / fp32 audit trail -- for every live
 * beam of every step the K (token, fp32 log-prob) pairs of torch.topk(log_softmax(logits), K)
 * (seq2seq.py:266-267) the bookkeeping was fed; entries of dead beams are left untouched.  Both
 * or neither; supported where the search runs inside the persistent kernel (bf16 headline
 * decoder), I2L_ERR_UNSUPPORTED otherwise. */
int i2l_decode_beam(const i2l_dec_desc* d, const void* packed, const float* enc, int32_t batch,
                    int32_t beam_size, int32_t start_id, int32_t end_id, int32_t max_length,
                    int64_t* out_tokens, int32_t* out_len, double* out_score, int32_t* trace_parent,
                    int32_t* trace_token, double* trace_score, int32_t* cand_token, float* cand_logp,
                    void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------- */
/* Multi-GPU token exchange: the one collective of the batch-sharded */
/* path -- every rank ends with the (n_total, T1) id matrix -- as direct peer     */
/* stores over NVLink.  The reference has no distributed code; this replaces the  */
/* `torch.distributed.all_gather_into_tensor` a data-parallel port would call     */
/* after Seq2SeqModel._greedy_search (model/seq2seq.py:192-232) and is checked    */
/* against that collective (hmer-img2latex_b200/dist.py: gather_tokens).          */
/* ------------------------------------------------------------------------- */
/* Bytes of one rank's receive buffer (symmetric across ranks; must start zeroed):
 * [2 parities][world][cap = ceil(n_total/world)][T1 + 1] int32 + flags. */
size_t i2l_token_exchange_buffer_bytes(int32_t world, int32_t n_total, int32_t T1);
/* Stores this rank's shard -- tokens (b,T1) int64, lengths (b) int32, steps (device scalar) --
 * into slot `rank`, parity seq & 1, of EVERY buffer in peer_buffers[world] (device pointers
 * valid on this device: the local buffer and the peer-mapped ones), then publishes sequence
 * number `seq` (non-zero, +1 per step) with a system-scope release.  Rows beyond b are -1. */
int i2l_token_exchange_write(const int64_t* tokens, const int32_t* lengths, const int32_t* steps, int32_t b,
                             int32_t T1, int32_t n_total, int32_t rank, int32_t world, void* const* peer_buffers,
                             uint32_t seq, void* stream);
/* Waits (on the device, bounded spin: *timeout_flag = 1 on expiry) until all `world` slots of
 * parity seq & 1 carry `seq`, then widens them into tokens (n_total,T1) int64 in global image
 * order (contiguous balanced shards), lengths (n_total) and steps = max over ranks.  Call
 * read(seq - 1) before write(seq) on the same stream: two parities are then race-free. */
int i2l_token_exchange_read(const void* local_buffer, int32_t world, int32_t n_total, int32_t T1, uint32_t seq,
                            int64_t* tokens, int32_t* lengths, int32_t* steps, int32_t* timeout_flag, void* stream);

/* ------------------------------------------------------------------------- */
/* Evaluation metrics -- the integer work of levenshtein_distance / bleu_n_score  */
/* (training/metrics.py:49-94, 97-179; calculate_metrics 182-223, cli.py:493-495) */
/* for a batch of (prediction, target) id sequences.  SURVEY 8f-3.                */
/* pred (B, ld_pred) / tgt (B, ld_tgt) int64 id matrices, pred_len / tgt_len (B)  */
/* int32 (clamped to [0, ld]).  out (B, 8) int32 per pair: [0] edit distance      */
/* D[rows][cols] of metrics.py:60-85, [1..4] clipped n-gram matches for n = 1..4  */
/* (metrics.py:131-152; 0 beyond max_n), [5] pred_len, [6] tgt_len, [7] 0.  The   */
/* float formulas (metrics.py:87-94, 160-179) are applied by the host wrapper.    */
/* ------------------------------------------------------------------------- */
int i2l_sequence_metrics(const int64_t* pred, int32_t ld_pred, const int32_t* pred_len, const int64_t* tgt,
                         int32_t ld_tgt, const int32_t* tgt_len, int32_t batch, int32_t max_n, int32_t* out,
                         void* stream);

/* Validation loss / accuracy over teacher-forced logits -- replaces nn.CrossEntropyLoss(ignore_index = pad,
 * reduction = "mean", label_smoothing) as configured in training/trainer.py:111-115 and applied at 517-522, and
 * masked_accuracy (training/metrics.py:226-238).  logits (rows, vocab) fp32 row-major (rows = B*T of the (B,T,V)
 * tensor i2l_decoder_forward writes), targets (rows) int64.  loss: device float (NaN when no row counts, like torch);
 * counts: device int32[4] = {correct, non-pad tokens, targets outside [0,vocab) (torch raises), 0}. */
size_t i2l_xent_workspace_bytes(int32_t rows);
int i2l_xent_metrics(const float* logits, const int64_t* targets, int32_t rows, int32_t vocab, int64_t ignore_index,
                     float label_smoothing, float* loss, int32_t* counts, void* workspace, size_t workspace_bytes,
                     void* stream);

/* Per-row stable compaction out[b] = [x for x in ids[b, :len[b]] if x not in drop]: the id-level effect of
 * tokenizer.decode (drops the four specials, data/tokenizer.py:177-189) + tokenizer.encode on the predictions
 * (cli.py:476-479) and of the padding filter on the targets (cli.py:471-474).  len: NULL => every row has ld
 * ids.  drop_host: HOST pointer to n_drop <= 8 ids.  out (B, ld_out) int64, out_len (B) int32. */
int i2l_filter_ids(const int64_t* ids, int32_t ld, const int32_t* len, int32_t batch, const int64_t* drop_host,
                   int32_t n_drop, int64_t* out, int32_t ld_out, int32_t* out_len, void* stream);

/* ------------------------------------------------------------------------- */
/* Measurement aids (bench.py): count of kernel launches issued by the library  */
/* and optional CUDA-event timing of its named kernels on the launching stream. */
/* ------------------------------------------------------------------------- */
long long i2l_launch_count(void);
void i2l_prof_enable(int on);
void i2l_prof_reset(void);
int i2l_prof_count(void);
/* synchronises the recorded events; total_ms = sum over launches since the last reset */
int i2l_prof_get(int idx, char* name, int name_len, int* launches, float* total_ms);

#ifdef __cplusplus
}
#endif
#endif /* I2L_B200_H_ */
