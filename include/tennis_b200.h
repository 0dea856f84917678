/* tennis_b200 — C ABI of the B200-native hot path of HaydenFaulkner/Tennis.
 *
 * Every entry point replaces an operator that the reference reaches through MXNet/Gluon (the reference
 * has no FFI of its own; its "plugin boundary" is the Gluon HybridBlock API).  The citation after each
 * declaration names the reference call site the function stands in for (paths are into the reference tree).
 *
 * Conventions
 *   - plain C types only; device pointers are raw addresses the caller owns (PyTorch is only a container);
 *   - every function returns 0 on success, a negative tn_status on failure; tn_last_error() has the text;
 *   - all device work is enqueued asynchronously on the caller's stream (a cudaStream_t passed as void*);
 *   - handles own only the repacked weights, are bound to one device and are not thread-safe;
 *   - there is NO CPU fallback: on a device that is not sm_100 every call returns TN_ERR_ARCH.
 */
#ifndef TENNIS_B200_H_
#define TENNIS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  TN_OK = 0,
  TN_ERR_INVALID = -1,   /* bad argument */
  TN_ERR_CUDA = -2,      /* CUDA runtime error (text in tn_last_error) */
  TN_ERR_ARCH = -3,      /* device is not compute capability 10.x */
  TN_ERR_WORKSPACE = -4, /* workspace too small */
  TN_ERR_NCCL = -5
} tn_status;

enum { TN_ARCH_DENSENET121 = 0, TN_ARCH_RESNET18_V2 = 1 };
enum { TN_FRAMES_F32_NCHW = 0, TN_FRAMES_U8_NHWC = 1 };
/* arithmetic of the CNN forward: TN_PRECISION_BF16 = bf16 operands and activations, fp32 accumulation (speed path, logits within
 * ~2e-2 of the fp32 reference); TN_PRECISION_SPLIT_BF16 = every tensor a (hi, lo) bf16 pair and every contraction three
 * tensor-core products hi*Wh + hi*Wl + lo*Wh accumulated in fp32 (fp32-grade: logits within 1e-3, DenseNet-121 only) */
enum { TN_PRECISION_BF16 = 0, TN_PRECISION_SPLIT_BF16 = 1 };
enum { TN_CELL_GRU = 0, TN_CELL_LSTM = 1 };
enum { TN_POOL_MAX = 0, TN_POOL_MEAN = 1 };

typedef struct tn_backbone tn_backbone_t;
typedef struct tn_birnn tn_birnn_t;
typedef struct tn_gnmt tn_gnmt_t;
typedef void* tn_stream_t; /* cudaStream_t */

int tn_version(void);
const char* tn_last_error(void); /* thread-local, valid until the next call on this thread */
int tn_device_check(int device);  /* TN_OK iff `device` is an sm_100 part */

/* Launch accounting for benchmarks: kernels launched by this library are always counted; with timing_on, every
 * launch is additionally bracketed by CUDA events on its stream.  tn_profile_read sums them (ms) per kernel family:
 * the tcgen05 conv-GEMM vs everything else. */
int tn_profile_enable(int timing_on);
int tn_profile_read(double* conv_gemm_ms, long long* conv_gemm_launches, double* other_ms, long long* other_launches,
                    int reset);

/* ------------------------------------------------------------------ per-frame CNN backbone
 * Replaces gluoncv.model_zoo.get_model(name).features as called at train.py:204, evaluate.py:125,
 * train_gnmt.py:150 and wrapped at models/vision/definitions.py:22,30 (FrameModel.backbone).
 *
 * `params` is ONE flat host fp32 array in the order Gluon's `features.collect_params()` enumerates:
 *   DenseNet-121: conv0.weight(64,3,7,7); bn0{gamma,beta,running_mean,running_var};
 *                 for block in 1..4: for layer: bn1{4}, conv1.weight(128,Cin,1,1), bn2{4}, conv2.weight(32,128,3,3);
 *                   after blocks 1..3: transition bn{4}, conv.weight(C/2,C,1,1);
 *                 final bn{4}.
 *   ResNet-18 v2: bn_data{4}; conv0.weight(64,3,7,7); bn0{4};
 *                 for stage in 1..4: for block in 1..2: bn1{4}, conv1.weight(C,Cin,3,3), bn2{4}, conv2.weight(C,C,3,3),
 *                   [downsample.weight(C,Cin,1,1) when Cin != C];
 *                 final bn{4}.
 */
size_t tn_backbone_param_count(int arch);
int tn_backbone_feature_dim(int arch, int h, int w); /* 1024 @224 / 4096 @512 (DenseNet), 512 (ResNet) */
int tn_backbone_create(tn_backbone_t** out, int arch, int device, const float* params, size_t n_params);
void tn_backbone_destroy(tn_backbone_t* bb);
/* Select the arithmetic of subsequent tn_backbone_forward calls (and of tn_backbone_workspace_bytes, which grows in the split
 * mode).  The reference computes in fp32 (train.py:204, evaluate.py:125); default TN_PRECISION_BF16. */
int tn_backbone_set_precision(tn_backbone_t* bb, int mode);
size_t tn_backbone_workspace_bytes(const tn_backbone_t* bb, int n_frames, int h, int w);
/* frames: device, (n,3,h,w) fp32 NCHW already normalised (dataset.py:214-217 / train.py:142-147), or
 *         (n,h,w,3) uint8 NHWC raw pixels (ToTensor+Normalize is then applied on the device);
 * feats : device fp32 (n, D); feats_bf16: optional device bf16 copy (n, D) for the RNN input projection. */
int tn_backbone_forward(tn_backbone_t* bb, const void* frames, int frames_dtype, int n_frames, int h, int w,
                        float* feats, void* feats_bf16, void* workspace, size_t workspace_bytes, tn_stream_t stream);

/* ------------------------------------------------------------------ single convolution (building block / test hook)
 * gluon nn.Conv2D(use_bias=False) with the pre-activation BatchNorm+ReLU of the consuming layer fused in front and
 * an optional folded BatchNorm(+ReLU) / residual behind, as DenseNet/ResNet-v2 use it ([UPSTREAM] gluoncv densenet /
 * resnetv2 blocks); also stands in for the Debug model's Conv2D (definitions.py:121).
 * mode: 0 = RxS conv, 1 = 1x1 conv on the 2x2-average of the activated input (DenseNet transition),
 *       2 = 7x7 stem on a channel-padded NHWC4 image.
 * weight: host fp32 OIHW; pro_/epi_ scale+shift: host fp32 per-channel arrays or NULL. */
typedef struct tn_conv tn_conv_t;
int tn_conv_create(tn_conv_t** out, int device, const float* weight, int Cout, int Cin, int R, int S, int mode,
                   const float* pro_scale, const float* pro_shift, const float* epi_scale, const float* epi_shift);
void tn_conv_destroy(tn_conv_t* c);
/* x: device NHWC bf16 (n,H,W,in_cstride); out: device NHWC bf16 or fp32 (n,Ho,Wo,out_cstride) written at out_coff;
 * residual: optional device NHWC bf16 at the output resolution. */
int tn_conv_forward(tn_conv_t* c, const void* x, int in_cstride, int n, int H, int W, int stride, int pad, int pro_relu,
                    int epi_relu, void* out, int out_cstride, int out_coff, int out_fp32, const void* residual,
                    int res_cstride, tn_stream_t stream);
/* (n,3,h,w) fp32 NCHW -> (n,h,w,4) bf16 NHWC4 (input layout of the stem). */
int tn_frames_to_nhwc4(const float* frames, void* out_bf16, int n, int h, int w, tn_stream_t stream);

/* ------------------------------------------------------------------ Dense / temporal pooling
 * gluon nn.Dense(num_classes, flatten=True): definitions.py:25,31-32 (FrameModel.classes),
 * :60,70-71 (TemporalPooling.classes), :101,108-109 (CNNRNN.classes).  y = x W^T + b, W is (out,in). */
int tn_dense_forward(const float* x, const float* weight, const float* bias, float* y, int rows, int in_dim,
                     int out_dim, tn_stream_t stream);
/* F.max / F.mean over axis 1: definitions.py:66-69 (TemporalPooling), :107 (CNNRNN). x (B,T,D) -> y (B,D) */
int tn_temporal_pool(const float* x, float* y, int B, int T, int D, int pool, tn_stream_t stream);

/* ------------------------------------------------------------------ fused (bi)directional RNN layer
 * Replaces mx.gluon.rnn.GRU/LSTM(hidden, layout='NTC', bidirectional=True) at definitions.py:93-96,106 and,
 * with valid_length, cell.unroll(...) of BidirectionalCell / GRUCell / LSTMCell at gnmt.py:143-145.
 * Host fp32 weights per direction in Gluon order: i2h_weight (G*H, D), h2h_weight (G*H, H), i2h_bias, h2h_bias
 * (G = 3 GRU [r,z,n], 4 LSTM [i,f,g,o]); pass the reverse direction's pointers as NULL for ndir = 1. */
int tn_birnn_create(tn_birnn_t** out, int device, int cell, int D, int H, int ndir, const float* const* i2h_weight,
                    const float* const* h2h_weight, const float* const* i2h_bias, const float* const* h2h_bias);
void tn_birnn_destroy(tn_birnn_t* r);
/* on=1: compute the input projection in split-bf16 (x = hi + lo, three tensor-core products): ~fp32-accurate gates for the
 * captioner's encoder, where token ids must match; default 0 (plain bf16 operands, fp32 accumulate). */
int tn_birnn_set_precise(tn_birnn_t* r, int on);
/* Training: refresh the handle's packed copies from the (updated) fp32 parameters ON THE DEVICE, asynchronously on `stream`
 * (same four arrays as tn_birnn_create, but device pointers) -- what gluon.Trainer.step leaves behind at train.py:424. */
int tn_birnn_update_weights(tn_birnn_t* r, const float* const* i2h_weight, const float* const* h2h_weight,
                            const float* const* i2h_bias, const float* const* h2h_bias, tn_stream_t stream);
size_t tn_birnn_workspace_bytes(const tn_birnn_t* r, int B, int T);
/* x: device (B,T,D) fp32, or bf16 when x_is_bf16; valid_len: device int32 (B) or NULL;
 * y (B,T,ndir*H) / ymax (B,ndir*H) / h_final, c_final (ndir,B,H): device fp32, each may be NULL. */
int tn_birnn_forward(tn_birnn_t* r, const void* x, int x_is_bf16, const int32_t* valid_len, int B, int T, float* y,
                     float* ymax, float* h_final, float* c_final, void* workspace, size_t workspace_bytes,
                     tn_stream_t stream);

/* ------------------------------------------------------------------ training of the temporal head (V7, head-only)
 * The published CNN-RNN (models/README.md id 0042) trains BiGRU(128)+max+Dense on frozen / pre-extracted features
 * (train.py:197-217,231-236).  These entry points implement that step: forward with saved activations, softmax
 * cross-entropy (gluon SoftmaxCrossEntropyLoss, train.py:324,419), backward (ag.backward, train.py:421) and the
 * optimiser updates of gluon.Trainer.step (train.py:298-299,424; train_gnmt.py:310,337).  The CNN backward is not built. */
int tn_birnn_forward_train(tn_birnn_t* r, const void* x, int x_is_bf16, int B, int T, float* y, float* ymax, float* gx,
                           float* cseq, void* workspace, size_t workspace_bytes, tn_stream_t stream);
/* On return the first B*T*ndir*G*H floats of `workspace` hold d(loss)/d(x W_i2h^T) for every (b,t) and direction: a caller that
 * trains the CNN end to end multiplies them by W_i2h (tn_sgemm) to get the gradient w.r.t. the features. */
size_t tn_birnn_backward_workspace_bytes(const tn_birnn_t* r, int B, int T);
int tn_birnn_backward(tn_birnn_t* r, const void* x, int x_is_bf16, int B, int T, const float* gx, const float* y,
                      const float* cseq, const float* ymax, const float* d_ymax, const float* dy, float* dW_ih, float* dW_hh,
                      float* db_ih, float* db_hh, void* workspace, size_t workspace_bytes, tn_stream_t stream);
/* loss (B) = -log_softmax(logits)[label]; dlogits (B,C) = softmax - onehot (head gradient 1 per sample); either may be NULL */
int tn_softmax_ce(const float* logits, const int32_t* labels, float* loss, float* dlogits, int B, int C, tn_stream_t stream);
/* gluonnlp MaskedSoftmaxCELoss forward (train_gnmt.py:256,282,332): pred (B,T,V), label (B,T) float ids, valid_len (B) float
 * -> loss (B) = sum_{t < valid_len} CE_t / T. */
int tn_masked_softmax_ce(const float* pred, const float* label, const float* valid_len, float* loss, int B, int T, int V,
                         tn_stream_t stream);
int tn_dense_backward(const float* x, const float* weight, const float* dy, float* dx, float* dweight, float* dbias, int rows,
                      int in_dim, int out_dim, tn_stream_t stream);
/* sgd: g' = rescale*g + wd*w; m = momentum*m - lr*g'; w += m.   adam (step t >= 1): bias-corrected, eps outside sqrt. */
int tn_sgd_mom_update(float* weight, const float* grad, float* mom, size_t n, float lr, float momentum, float wd,
                      float rescale_grad, tn_stream_t stream);
int tn_adam_update(float* weight, const float* grad, float* mean, float* var, size_t n, float lr, float beta1, float beta2,
                   float eps, float wd, float rescale_grad, int t, tn_stream_t stream);

/* ------------------------------------------------------------------ GNMT decoder + beam search
 * Replaces GNMTDecoder (models/captioning/gnmt.py:163-404) + gluonnlp NMTModel.decode_step/decode_seq glue
 * (tgt_embed, tgt_proj) and gluonnlp BeamSearchSampler/BeamSearchScorer as driven by
 * BeamSearchTranslator.translate (utils/translation.py:42-82).  The encoder (gnmt.py:136-160) is tn_birnn_forward
 * with valid_len.  Host fp32 weights, Gluon layouts: per decoder layer i2h_weight (G*H, in_l) with in_0 = E+H,
 * in_l = 2H; h2h_weight (G*H, H); biases (G*H); attention query projection (H,H) (scaled Luong: query projected and
 * divided by sqrt(H), keys/values are the raw encoder memory); tgt_embed (V,E); tgt_proj weight (V,H) + bias (V). */
int tn_gnmt_create(tn_gnmt_t** out, int device, int cell, int H, int E, int V, int num_layers, int use_residual,
                   const float* const* i2h_weight, const float* const* h2h_weight, const float* const* i2h_bias,
                   const float* const* h2h_bias, const float* query_weight, const float* embed_weight,
                   const float* proj_weight, const float* proj_bias);
void tn_gnmt_destroy(tn_gnmt_t* g);
size_t tn_gnmt_workspace_bytes(const tn_gnmt_t* g, int rows, int max_len);
/* One step for R rows (model.decode_step, translation.py:52): step_ids device float (R); h_in/c_in device (L,R,H);
 * att_in (R,H); mem (R/rows_per_mem, T, H); src_len device int32 or NULL; outputs logits (R,V), states. */
int tn_gnmt_decode_step(tn_gnmt_t* g, const float* step_ids, const float* h_in, const float* c_in, const float* att_in,
                        const float* mem, const int32_t* src_len, int rows_per_mem, int R, int T, float* logits,
                        float* h_out, float* c_out, float* att_out, void* workspace, size_t workspace_bytes,
                        tn_stream_t stream);
/* GNMTDecoder.__call__(step_input, states) (gnmt.py:306-404): the decoder BLOCK's step.  step_emb device (R,E) already
 * embedded inputs; out (R,H) = rnn_out of the last layer (no target projection); states as in tn_gnmt_decode_step. */
int tn_gnmt_decoder_step(tn_gnmt_t* g, const float* step_emb, const float* h_in, const float* c_in, const float* att_in,
                         const float* mem, const int32_t* src_len, int rows_per_mem, int R, int T, float* out, float* h_out,
                         float* c_out, float* att_out, void* workspace, size_t workspace_bytes, tn_stream_t stream);
/* Teacher-forced decode (model.decode_seq, gnmt.py:254-304; train_gnmt.py:280,331): tgt_ids device float (B,T_tgt),
 * tgt_valid_len device int32 (B) or NULL, h0/c0 (L,B,H), logits (B,T_tgt,V). */
int tn_gnmt_decode_seq(tn_gnmt_t* g, const float* tgt_ids, const int32_t* tgt_valid_len, const float* h0,
                       const float* c0, const float* mem, const int32_t* src_len, int B, int T_src, int T_tgt,
                       float* logits, void* workspace, size_t workspace_bytes, tn_stream_t stream);
/* Beam search (translator.translate after the encoder): samples device int32 (B,beam,max_len+2) of which the first
 * *out_len columns are the reference's result, scores device (B,beam) descending, valid_len device int32 (B,beam). */
int tn_gnmt_beam_search(tn_gnmt_t* g, const float* mem, const int32_t* src_len, const float* h0, const float* c0, int B,
                        int T, int beam, int max_len, float alpha, float K, int bos, int eos, int32_t* samples,
                        float* scores, int32_t* valid_len, int* out_len, void* workspace, size_t workspace_bytes,
                        tn_stream_t stream);

/* ------------------------------------------------------------------ captioner training (G9; train_gnmt.py:330-337)
 * Building blocks of `with autograd.record(): out, _ = model(src, tgt[:, :-1], src_vl, tgt_vl - 1); loss = ...; loss.backward()`:
 * the GRU/LSTM cells of GNMTEncoder/GNMTDecoder unrolled step by step with saved activations (gnmt.py:143-145,345-404),
 * scaled-Luong attention (gluonnlp DotProductAttentionCell), Embedding, MaskedSoftmaxCELoss and output Dropout, each with its
 * backward.  fp32 throughout; matrices are row-major with explicit row strides (a time step of a (B,T,C) tensor is addressed in
 * place).  The Python side (tennis_b200/models/captioning/train_graph.py) sequences them. */
/* C[M,N] = alpha * op(A) op(B) + beta * C; op(A) = A (M x K, lda) or A^T (A stored K x M); op(B) = B (K x N) or B^T (N x K). */
int tn_sgemm(int transA, int transB, int M, int N, int K, float alpha, const float* A, int lda, const float* B, int ldb,
             float beta, float* C, int ldc, tn_stream_t stream);
/* One cell step on B rows: gi = x W_i2h^T and gh = h_prev W_h2h^T without biases (row strides in floats), Gluon gate order.
 * Writes h (and optionally a second strided copy h_out2), c (LSTM) and the 4H saved values per row that the backward needs
 * (LSTM: i f g o; GRU: r z n and W_hn h + b_hn).  h_prev / c_prev NULL = zero state. */
int tn_rnn_cell_forward(int cell, int B, int H, const float* gi, long long gi_stride, const float* gh, long long gh_stride,
                        const float* bi, const float* bh, const float* h_prev, long long hp_stride, const float* c_prev,
                        long long cp_stride, float* h_out, long long ho_stride, float* h_out2, long long ho2_stride, float* c_out,
                        long long co_stride, float* save, long long sv_stride, tn_stream_t stream);
/* Backward of step t.  dh/dc (B,H): in = gradient w.r.t. this step's state from step t+1, out = the direct part of the gradient
 * w.r.t. the previous state (the caller adds dgh W_h2h).  dy/dy2: gradients w.r.t. this step's output(s).  With valid_len, rows
 * with t >= len produce zeros and at t == len-1 the final-state gradients dh_last/dc_last enter (MXNet unroll(valid_length),
 * SURVEY.md A.4).  dgi (and dgh for GRU, where they differ): gate pre-activation gradients, strided rows. */
int tn_rnn_cell_backward(int cell, int B, int H, int t, const int32_t* valid_len, const float* save, long long sv_stride,
                         const float* h_prev, long long hp_stride, const float* c_prev, long long cp_stride, const float* c_cur,
                         long long cc_stride, const float* dy, long long dy_stride, const float* dy2, long long dy2_stride,
                         const float* dh_last, const float* dc_last, float* dh, float* dc, float* dgi, long long dgi_stride,
                         float* dgh, long long dgh_stride, tn_stream_t stream);
/* Whole-sequence drivers of the two calls above (the per-step loop of cell.unroll, gnmt.py:143-145, in C instead of Python).
 * GI (B,T,G*H) = X W_i2h^T; Hb / Cb (B,T+1,H): slot 0 = initial state, slot t+1 = state after step t; S (B,T,4H); scratch (B,G*H).
 * Backward: dh / dc zero-initialised, on return the gradients w.r.t. the initial state; DGH = DGI for LSTM; dWh accumulated. */
int tn_rnn_unroll_forward(int cell, int B, int T, int H, const float* GI, const float* Wh, const float* bi, const float* bh,
                          float* Hb, float* Cb, float* S, float* scratch, tn_stream_t stream);
int tn_rnn_unroll_backward(int cell, int B, int T, int H, const int32_t* valid_len, const float* S, const float* Hb, const float* Cb,
                           const float* Wh, const float* dY, const float* dh_last, const float* dc_last, float* dh, float* dc,
                           float* DGI, float* DGH, float* dWh, tn_stream_t stream);
/* q (B,H) = projected query; w (B,T) = softmax((q/sqrt(H)) mem^T masked by src_len) * mask; ctx = w mem, written to one or two
 * strided destinations.  Backward accumulates into dmem (B,T,H) and writes dq. */
int tn_attention_forward(const float* q, long long q_stride, const float* mem, const int32_t* src_len, int B, int T, int H,
                         float* w, float* ctx1, long long c1_stride, float* ctx2, long long c2_stride, tn_stream_t stream);
int tn_attention_backward(const float* q, long long q_stride, const float* mem, const int32_t* src_len, int B, int T, int H,
                          const float* w, const float* dctx1, long long d1_stride, const float* dctx2, long long d2_stride,
                          float* dq, long long dq_stride, float* dmem, tn_stream_t stream);
/* dweight[ids[r]] += dy[r] (ids are float token ids, as the scripts pass them). */
int tn_embedding_backward(const float* ids, const float* dy, long long dy_stride, float* dweight, int N, int E, int V,
                          tn_stream_t stream);
/* MaskedSoftmaxCELoss with gradient: loss (B) as tn_masked_softmax_ce; dpred (B,T,V) = head_grad[b]/T (softmax - onehot) on valid
 * tokens, 0 elsewhere (dpred/head_grad may be NULL); workspace_bt: B*T floats. */
int tn_masked_softmax_ce_grad(const float* pred, const float* label, const float* valid_len, const float* head_grad,
                              float* loss, float* dpred, float* workspace_bt, int B, int T, int V, tn_stream_t stream);
/* Inverted-dropout mask in {0, 1/(1-p)} from a counter-based generator; y = x * mask with rows t >= seq_len[b] zeroed
 * (Dropout at gnmt.py:152,389 + SequenceMask at :157-159,298-301); either mask or seq_len may be NULL. */
int tn_dropout_mask(float* mask, size_t n, float p, unsigned long long seed, tn_stream_t stream);
int tn_mul_mask(const float* x, const float* mask, const int32_t* seq_len, float* y, int B, int T, int C, tn_stream_t stream);
int tn_axpy(float* y, const float* x, float a, size_t n, tn_stream_t stream);

/* ------------------------------------------------------------------ CNN training (V7 with a trainable backbone)
 * `with ag.record(): out = net(x); ...; ag.backward(losses)` (train.py:415-421) through gluoncv DenseNet-121 / ResNet-18 v2:
 * training-mode BatchNorm (batch statistics, biased variance, running = momentum*running + (1-momentum)*batch)
 * fused with ReLU, im2col / col2im for the strided and 7x7 convolutions, max / average pooling, each with its backward.  fp32
 * NHWC: a row is a pixel, `ld*` is the channel count of the buffer (DenseNet's concat stays a channel offset).  These are the
 * memory-bound kernels (SIMT); the contractions are tn_gemm_tc below (tn_sgemm in the fp32 anchor mode); sequenced by
 * tennis_b200/models/vision/train_graph.py. */
/* Tensor-core contractions of the same training path (csrc/tn_gemm_tc.cu): the convolutions' forward, data-gradient and
 * weight-gradient GEMMs (train.py:415-421, MXNet Convolution forward/backward in fp32) on tcgen05.
 * tn_split_bf16: fp32 matrix (rows x cols, row stride ld) -> bf16 planes hi (+ lo, NULL = plain bf16) with x = hi + lo;
 *   transpose = 0: plane[orow(r)][c], 1: plane[c][orow(r)]; pad_h/pad_w > 0: r = (n,y,x) over pad_h x pad_w frames is re-indexed
 *   to (n,y+1,x+1) of the zero-padded grid of (pad_h+2) rows of pad_pitch (0 = pad_w+2) pixels (the border zeros are written by
 *   the same pass); shift in {-1,0,1}: transposed padded planes only, every pixel lands `shift` positions later; out_ld % 8 == 0.
 * tn_gemm_tc: D[m,n] = sum_t sum_k A[m + taps[4t], taps[4t+1] + k] * B[n + taps[4t+2], taps[4t+3] + k], k < K, taps on the HOST
 *   (NULL = one tap, no offsets; out-of-range reads are zeros; contraction offsets must be multiples of 8), passes = 3: hi*hi + hi*lo + lo*hi (fp32-grade), 1: hi*hi;
 *   C[orow(m)*c_row_stride + n*c_col_stride] = alpha*D + beta*C, unpad_h/w > 0: m runs over the padded grid, border rows are
 *   dropped.  tile_taps != 0: the taps are ntaps INDEPENDENT products in one launch (a 3x3 weight gradient), tap t written at
 *   C + t*c_tap_stride.  Few output tiles + long K: split-K through `workspace` (deterministic, tn_gemm_tc_workspace_bytes(M, N,
 *   tile_taps ? ntaps : 0) is always enough); a strided column needs the workspace. */
long long tn_gemm_tc_workspace_bytes(int M, int N, int tile_taps);
int tn_split_bf16(const float* src, long long ld, long long rows, int cols, int transpose, int pad_h, int pad_w, int pad_pitch,
                  int shift, void* hi, void* lo, long long out_ld, tn_stream_t stream);
int tn_gemm_tc(int M, int N, int K, int ntaps, const int* taps, int passes, const void* a_hi, const void* a_lo, long long a_rows,
               long long a_kdim, long long a_ld, const void* b_hi, const void* b_lo, long long b_rows, long long b_kdim,
               long long b_ld, float alpha, float beta, float* C, long long c_row_stride, long long c_col_stride, int tile_taps,
               long long c_tap_stride, int unpad_h, int unpad_w, void* workspace, long long workspace_bytes, tn_stream_t stream);
int tn_im2col_nhwc(const float* x, long long ldx, int N, int H, int W, int C, int R, int S, int stride, int pad, float* col,
                   tn_stream_t stream);
int tn_col2im_nhwc(const float* dcol, int N, int H, int W, int C, int R, int S, int stride, int pad, float* dx, long long lddx,
                   tn_stream_t stream); /* dx += scatter(dcol) */
int tn_bn_train_forward(const float* x, long long ldx, long long M, int C, const float* gamma, const float* beta, float eps,
                        float momentum, float* running_mean, float* running_var, int relu, float* mean, float* var, float* y,
                        long long ldy, tn_stream_t stream);
int tn_bn_train_backward(const float* x, long long ldx, const float* y, long long ldy, const float* dy, long long lddy, long long M,
                         int C, const float* mean, const float* var, const float* gamma, float eps, int relu, float* dgamma,
                         float* dbeta, float* dx, long long lddx, int accumulate, tn_stream_t stream);
int tn_maxpool_nhwc_forward(const float* x, long long ldx, int N, int H, int W, int C, int k, int stride, int pad, float* y,
                            long long ldy, int32_t* idx, tn_stream_t stream);
int tn_maxpool_nhwc_backward(const float* dy, long long lddy, const int32_t* idx, long long rows, int C, float* dx, long long lddx,
                             tn_stream_t stream); /* dx += */
int tn_avgpool_nhwc_forward(const float* x, long long ldx, int N, int H, int W, int C, int kh, int kw, float* y, long long ldy,
                            tn_stream_t stream); /* window = stride = (kh,kw), floor */
int tn_avgpool_nhwc_backward(const float* dy, long long lddy, int N, int H, int W, int C, int kh, int kw, float* dx, long long lddx,
                             int accumulate, tn_stream_t stream);

/* ------------------------------------------------------------------ device-side metrics (metrics/vision.py:27-58, train.py:427-431)
 * logits device fp32 (N,C), labels device int32 (N).  conf device uint64 (C,C) indexed [label][argmax prediction] (first maximal
 * class, like numpy argmax); hits device uint64 [3] = {top-1 hits, top-k hits (stable descending order: ties -> lower class
 * index first), samples seen}.  Counters accumulate across calls; the host zeroes them to reset and reads them once per epoch. */
int tn_metrics_update(const float* logits, const int32_t* labels, int N, int C, int top_k, unsigned long long* conf,
                      unsigned long long* hits, tn_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* TENNIS_B200_H_ */
