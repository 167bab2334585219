/* avsr_b200.h - C ABI of the B200-native AVSR seq2seq hot path.
 *
 * The reference (georgesterpu/avsr-tf1) is pure Python on TensorFlow 1.13 and has
 * no FFI of its own; its hot path bottoms out in TF library calls.  Every entry
 * point below therefore names the TF call site in the reference that it replaces
 * (file:line into the reference tree).  The Python host in avsr_tf1_b200/ binds
 * these with ctypes (avsr_tf1_b200/_lib.py); INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host
 *   - all floating point is fp32; lengths / ids are int32
 *   - sequences are frame-major (time-major): [T, B, F], F contiguous
 *   - `stream` is a cudaStream_t passed as void*
 *   - return 0 on success; otherwise avsr_last_error() describes the failure
 *   - no allocation, no synchronisation: calls only enqueue kernels on `stream`
 *     (so a whole training step can be captured into one CUDA graph)
 */
#ifndef AVSR_B200_H_
#define AVSR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* avsr_stream_t;

const char* avsr_last_error(void);
int avsr_version(void);
/* kernels launched by this library since load (bench.py's gpu_launches) */
unsigned long long avsr_launch_count(void);
/* Device timing of the hot kernels (bench.py roofline): while enabled, every launch of a persistent
 * attention-LSTM / LSTM kernel and of the tcgen05 GEMM outside stream capture is bracketed by CUDA events on its
 * launching stream.  avsr_kernel_timing(enable) resets the record and returns the previous setting;
 * avsr_kernel_times fills the summed milliseconds and launch counts of the 5 classes
 * {0 attention-LSTM fwd, 1 attention-LSTM bwd, 2 LSTM fwd, 3 LSTM bwd, 4 GEMM} (synchronises on the events). */
int avsr_kernel_timing(int enable);
int avsr_kernel_times(float* ms_out5, int* launches_out5);
/* 1 if the tcgen05 tensor-core GEMM path is enabled (default), 0 = exact fp32 CUDA cores */
int avsr_set_tensor_cores(int enable);
int avsr_get_tensor_cores(void);
/* Precision modes.  tensor cores ON (default): big products run on tcgen05 kind::tf32 with fp32
 * accumulation; every operand is rounded to tf32 (round-to-nearest) by the kernel that produces it
 * (weights: avsr_round_tf32 / the copy avsr_adam_clip_step maintains), so the hardware's truncation is a
 * no-op.  OFF: exact fp32 everywhere on CUDA cores. */
int avsr_round_tf32(avsr_stream_t stream, const float* src, float* dst, long long n);

/* ---- dense products: tf.matmul / tf.layers.dense call sites ------------------
 * (LSTMCell kernel product cells.py:14; memory_layer attention.py:26; my_dense
 * decoder_unimodal.py:112; state projections encoder.py:134-137,
 * decoder_bimodal.py:480-492; and their tf.gradients counterparts seq2seq.py:222)
 * C[M,N](ldc) = beta*C + op(A) op(B) (+ bias[N]);  beta in {0,1}
 * transA=0: A is [M,K] row-major (lda); transA=1: A is stored [K,M].  Same for B.
 * round_out=1 (beta=0 only, tensor-core mode only): C is stored tf32-rounded because it is itself
 * the operand of a later tensor-core product (the attention vector). */
int avsr_gemm(avsr_stream_t stream, int transA, int transB, int M, int N, int K, const float* A, int lda,
              const float* B, int ldb, float* C, int ldc, float beta, const float* bias, int round_out);
/* out[N] += column sums of X[M,N] (ldx)  (bias gradients) */
int avsr_colsum(avsr_stream_t stream, const float* X, int M, int N, int ldx, float* out);

/* ---- input batch normalisation: tf.layers.batch_normalization, encoder.py:44-50
 * statistics over all rows (= B*T positions INCLUDING padding), per feature.
 * Split in stats / apply so data-parallel ranks can all-reduce `sums` between. */
int avsr_bn_stats(avsr_stream_t stream, const float* x, long long rows, int F, float* sums /*[2F]: sum, sumsq*/);
int avsr_bn_apply_train(avsr_stream_t stream, const float* x, long long rows, int F, const float* sums,
                        double count, const float* gamma, const float* beta, float eps, float momentum,
                        float* y, float* xhat, float* invstd /*[F]*/, float* moving_mean, float* moving_var);
/* same, for x stored [d0,d1,F] (batch-major, as the reference's tensors are: encoder.py:44-50 normalises
 * [B,T,F]) with y / xhat written [d1,d0,F] (frame-major device layout): the boundary transpose of
 * avsr_transpose01 fused into the normalisation */
int avsr_bn_apply_train_t(avsr_stream_t stream, const float* x, int d0, int d1, int F, const float* sums,
                          double count, const float* gamma, const float* beta, float eps, float momentum,
                          float* y, float* xhat, float* invstd /*[F]*/, float* moving_mean, float* moving_var);
/* y of both apply calls is tf32-rounded in tensor-core mode (it only feeds the layer-0 gate product) */
int avsr_bn_apply_eval(avsr_stream_t stream, const float* x, long long rows, int F, const float* gamma,
                       const float* beta, const float* moving_mean, const float* moving_var, float eps, float* y);
/* dgamma / dbeta of the input normalisation from what the layer-0 weight-gradient pass already formed (the
 * normalised features only feed z = y Wx): dbeta = Wx colsum(dZ), dgamma_f = sum_n Wx[f,n] (dWx[f,n] - beta_f
 * colsum(dZ)[n]) / gamma_f with dWx = y^T dZ.  Accumulates into dgamma / dbeta; a feature whose gamma is exactly 0
 * contributes 0 to dgamma.  Saves the [T*B,F] product dZ Wx^T and two passes over it. */
int avsr_bn_input_grads(avsr_stream_t stream, const float* Wx, int ldw, const float* dWx, int ldg,
                        const float* colsum_dZ /*[N]*/, const float* gamma, const float* beta, int F, int N,
                        float* dgamma, float* dbeta);
/* backward: sums2 = [sum dy, sum dy*xhat] (all-reducible), then dx */
int avsr_bn_bwd_stats(avsr_stream_t stream, const float* dy, const float* xhat, long long rows, int F,
                      float* sums2 /*[2F]*/);
int avsr_bn_bwd_apply(avsr_stream_t stream, const float* dy, const float* xhat, long long rows, int F,
                      const float* sums2, double count, const float* gamma, const float* invstd, float* dx,
                      float* dgamma, float* dbeta);

/* ---- tf.reverse_sequence used by bidirectional_dynamic_rnn (encoder.py:110) --- */
int avsr_reverse_sequence(avsr_stream_t stream, const float* x, float* y, int T, int B, int F, const int* len);

/* boundary layout change: y[d1,d0,F] = x[d0,d1,F] (reference tensors are batch-major,
 * device tensors frame-major) */
int avsr_transpose01(avsr_stream_t stream, const float* x, float* y, int d0, int d1, int F);

/* lip crops as stored pixels: y[i] = (x[i] + shift) * scale.  dataset_writer.py:537 writes (v - 128) / 128 of uint8 pixels
 * into the records; a host batch that keeps the bytes crosses PCIe at a quarter of the size and is expanded here (exact:
 * every k / 128 is representable).  Buffers 16-byte aligned. */
int avsr_u8_to_f32(avsr_stream_t stream, const uint8_t* x, long long n, float scale, float shift, float* y);

/* ---- recurrent sequence op --------------------------------------------------
 * One call = one tf.nn.dynamic_rnn / seq2seq.dynamic_decode loop over an LSTMCell
 * (cells.py:14-18: gate order i,j,f,o, forget bias 1, cell_clip 1), optionally
 * wrapped in a seq2seq.AttentionWrapper with 1 or 2 mechanisms
 * (attention.py:132-191; AV-Align encoder.py:265-290; LAS decoder
 * decoder_unimodal.py:299-352; WLAS decoder decoder_bimodal.py:227-277).
 * Length semantics of dynamic_rnn(sequence_length) / impute_finished: past
 * len[b] the output row is 0 and the cell state is carried.                  */
enum { AVSR_ATTN_LUONG = 0, AVSR_ATTN_SCALED_LUONG = 1, AVSR_ATTN_BAHDANAU = 2, AVSR_ATTN_NORMED_BAHDANAU = 3 };

typedef struct AvsrAttnMech {
  int kind, Tm, Dm, A;
  const float* values;  /* [Tm,B,Dm] memory, zero past mem_len (encoder outputs already are) */
  const float* keys;    /* [Tm,B,A]  = values @ memory_layer */
  const int* mem_len;   /* [B] */
  const float* Wl;      /* attention_layer kernel [(H+Dm), A] */
  const float* Wq;      /* query_layer kernel [H,A]           (Bahdanau family) */
  const float* v;       /* attention_v [A] (normed: the effective g*v/|v|) */
  const float* g;       /* attention_g [1]                    (scaled Luong) */
  const float* bias;    /* attention_b [A]                    (normed Bahdanau) */
  /* saved by fwd for bwd (and exposed as parity probes) */
  float* align;         /* [T,B,Tm] alignments                 */
  float* hc;            /* [T,B,H+Dm] = [cell output | context] */
  float* pq;            /* [T,B,A] processed query             (Bahdanau family) */
  /* backward outputs (accumulated: caller zeroes them) */
  float* dkeys;         /* [Tm,B,A]  */
  float* dvalues;       /* [Tm,B,Dm] gradient through the context only */
  float* dWl;           /* [(H+Dm),A] */
  float* dWq;           /* [H,A] */
  float* dv;            /* [A] wrt the effective v */
  float* dg;            /* [1] */
  float* dbias;         /* [A] */
  float* dpq;           /* [T,B,A] scratch */
  float* ds;            /* [T,B,Tm] scratch: d(score) of every step (dkeys is formed after the loop) */
  float* dhc;           /* [T,B,H+Dm] scratch: d[cell output | context] of every step */
  const float* values_op; /* [Tm,B,Dm] the memory as the operand of a tensor-core product (tf32-rounded copy) or NULL =
                             values; read by the persistent Bahdanau kernels' projection PV = values Wl_c */
} AvsrAttnMech;

/* ScheduledEmbeddingTrainingHelper (decoder_unimodal.py:304-309) inside a whole-sequence call: step t's output decides,
 * per row with probability p, the decoder input of step t + 1 (an id drawn from Categorical(logits_t), logits_t =
 * out_t Wd + bd).  The caller prepares the teacher-forced sequence as usual (x = dropped embeddings of true_ids, gates =
 * x Wx + bias) and pre-fills used_ids = true_ids, sample_ids = -1; the recurrence replaces the rows that are drawn:
 * used_ids[t+1,b], sample_ids[t,b], x[t+1,b,:] (the dropped embedding of the drawn id, mask of the whole-sequence
 * avsr_dropout call on x: stream drop_stream + 3) and the x-projection it uses for that step.  Same generator streams
 * and arithmetic as avsr_sched_sample.  No gradient flows through the draws. */
typedef struct AvsrSampling {
  const float* Wd;         /* output layer kernel [O, V] */
  const float* bd;         /* [V] */
  const float* embedding;  /* [V, E] */
  const float* Wx;         /* x rows of the cell kernel [E, 4H] */
  const float* bias;       /* cell bias [4H] */
  int* used_ids;           /* [T,B] in/out */
  int* sample_ids;         /* [T,B] in/out */
  float* x;                /* [T,B,E] in/out */
  int V, E;
  uint32_t stream;         /* +0 Bernoulli select (hi = t, lo = b), +1 draw */
  uint32_t thr_p;          /* p * 2^32 */
} AvsrSampling;

typedef struct AvsrRnnSeq {
  int T, B, H, n_mech, output_attention;
  const int* len;       /* [B] */
  float* gates;         /* [T,B,4H] in: x_t @ Wx + bias; out: activations i,j,f,o */
  const float* Wrec;    /* [(At+H),4H] rows of the cell kernel under the x rows: [attention ; h] */
  const float* c0;      /* [B,H] initial cell state or NULL (zeros) */
  float* S;             /* [(T+1),B,At+H] state rows [attention | h]; caller initialises S[0] */
  float* craw;          /* [T,B,H] pre-clip cell values */
  float* out;           /* [T,B,O], O = At if output_attention else H; zero past len */
  float* cT;            /* [B,H] final (clipped) cell state, or NULL */
  float* hT;            /* [B,H] final h, or NULL */
  AvsrAttnMech mech[2];
  /* backward */
  const float* dout;    /* [T,B,O] */
  const float* dcT;     /* [B,H] or NULL */
  const float* dhT;     /* [B,H] or NULL */
  float* dZ;            /* [T,B,4H] gradient wrt the gate pre-activations (overwritten) */
  float* dA;            /* [T,B,At] scratch (gradient wrt attention vectors) */
  float* dWrec;         /* [(At+H),4H] accumulated */
  float* dc0;           /* [B,H] or NULL */
  float* dh0;           /* [B,H] or NULL */
  float* dbias;         /* [4H] or NULL: += column sums of dZ (the gradient of the cell bias) */
  float* work;          /* scratch, >= avsr_rnn_work_floats() floats; backward must see what forward left */
  float grad_scale;     /* power of two ~ 1/|gradient scale| (e.g. the token count): fp16 tensor-core operand
                           scaling of the persistent attention backward kernel; 0 = 1 */
  /* DropoutWrapper(LSTMCell) inside the loop (cells.py:46-54; non-variational: fresh masks every step).  The masks
   * are not stored: forward and backward regenerate them from the counter-based generator of avsr_dropout
   * (element (t, b, column) of stream drop_stream + {0: attention part of the cell input, 1: recurrent state h,
   * 2: cell output}).  The x part of the cell input is dropped by the caller (avsr_dropout on the sequence) before
   * the x-projection.  thr = keep probability * 2^32 (0 = keep everything, i.e. that dropout is off). */
  const uint32_t* rng;  /* [2] device words {seed, step counter}; NULL = no dropout */
  uint32_t drop_stream;
  uint32_t thr_in, thr_state, thr_out;
  /* step range [t_begin, t_end) of this call; 0,0 = the whole sequence.  A range forces the step-wise kernels: it
   * is how ScheduledEmbeddingTrainingHelper (decoder_unimodal.py:304-309) interleaves sampling with the recurrence.
   * stepwise != 0 forces them for a whole-sequence call too (the backward of a ranged forward). */
  int t_begin, t_end, stepwise;
  /* scheduled sampling inside a whole-sequence forward call, or NULL.  Only the persistent two-product kernels
   * implement it: ask avsr_rnn_sampling_fused() first and advance in step ranges with avsr_sched_sample otherwise. */
  const AvsrSampling* samp;
} AvsrRnnSeq;

/* At = sum of mechanism A; maxHD = max(H+Dm); maxA = max A; maxTm = max memory length (0,0,0,0 without attention) */
size_t avsr_rnn_work_floats(int B, int H, int At, int maxHD, int maxA, int maxTm);
/* sizeof(AvsrAttnMech), sizeof(AvsrRnnSeq) as compiled (bindings check their struct layout against it) */
int avsr_struct_sizes(int* out2);
/* 1 if avsr_rnn_seq_fwd(r) with r->samp set would draw the samples inside the recurrence (tensor-core mode, one Luong-family
 * mechanism, H = A = Dm = 256, T > 1), else 0 */
int avsr_rnn_sampling_fused(const AvsrRnnSeq* r);
int avsr_rnn_seq_fwd(avsr_stream_t stream, const AvsrRnnSeq* r);
int avsr_rnn_seq_bwd(avsr_stream_t stream, const AvsrRnnSeq* r);

/* effective v of normed Bahdanau and its backward (attention.py:34-42) */
int avsr_normed_v_fwd(avsr_stream_t stream, const float* v, const float* g, int A, float* veff);
int avsr_normed_v_bwd(avsr_stream_t stream, const float* v, const float* g, const float* dveff, int A, float* dv,
                      float* dg);

/* ---- randomness of the training graph -----------------------------------------
 * TF's Philox streams cannot be reproduced, so the reference's masks and samples are not bit-reproducible by
 * anyone; what is kept is their distribution and where they act.  One counter-based generator serves all of it:
 * word(seed, step, stream, hi, lo) = two rounds of the murmur3 32-bit finaliser over the five counters
 * (common.cuh avsr_rand_u32; restated in oracle/avsr_oracle.py rand_u32 so masks are bit-equal in the parity tests).
 * rng = device words {seed, step}: the host bumps `step` once per training step (graph-replay safe). */
/* inverted dropout tf.nn.dropout (DropoutWrapper cells.py:46-54): y[i] = word(first + i) < thr ? x[i] * 2^32 / thr : 0
 * for i < n (first + n <= 2^32) of stream `stream_id` (hi = 0, lo = first + i); `first` lets a step-wise caller
 * process a slice of a sequence with the mask of the whole-sequence call; in place allowed; the same call maps
 * dy -> dx.  thr = keep_prob * 2^32 (the exact keep probability is thr / 2^32, so the scaling is unbiased);
 * round_out: y stored tf32-rounded (it feeds a tensor-core product). */
int avsr_dropout(avsr_stream_t stream, const float* x, long long n, long long first, const uint32_t* rng,
                 uint32_t stream_id, uint32_t thr, int round_out, float* y);
/* ScheduledEmbeddingTrainingHelper.sample + next_inputs (decoder_unimodal.py:304-309): for row b,
 * with probability p (thr_p = p * 2^32; stream_id, hi = t, lo = b) the next decoder input id is drawn from
 * Categorical(logits[b,:]) (inverse CDF of the fp32 softmax with the word of stream_id + 1), else it is
 * true_next[b].  sampled[b] = drawn id or -1 (the helper's sample_ids). */
int avsr_sched_sample(avsr_stream_t stream, const float* logits, int B, int V, const uint32_t* rng,
                      uint32_t stream_id, int t, uint32_t thr_p, const int* true_next, int* next_ids, int* sampled);

/* ---- embedding lookup and its gradient (decoder_unimodal.py:170) ------------- */
int avsr_embedding_fwd(avsr_stream_t stream, const float* table, int V, int E, const int* ids, long long n,
                       float* out);
int avsr_embedding_bwd(avsr_stream_t stream, const float* dout, const int* ids, long long n, int V, int E,
                       float* dtable);

/* ---- seq2seq.sequence_loss with sequence_mask weights (seq2seq.py:142-171) ---
 * logits [T,B,V] (rows past labels_len are treated as zero logits, impute_finished);
 * labels [B,ldl] EOS-terminated.  loss_sum[0] += sum xent*w; dlogits = (softmax-onehot)*w*inv_denom.
 * inv_denom_dev / lr_t_dev are DEVICE scalars so a captured CUDA graph can be replayed with new values.
 * label_smoothing > 0 (seq2seq.py:147-155, devel.py:54-61): targets (1 - eps) onehot + eps / V and - as the reference's
 * smoothed loss function hands sequence_loss a reduced scalar - the UNMASKED mean over all T*B positions: rows past the
 * label length (imputed zero logits) add log V to loss_sum and get no gradient; the caller passes 1 / (T*B) as inv_denom. */
int avsr_seq_loss(avsr_stream_t stream, const float* logits, int T, int B, int V, const int* labels, int ldl,
                  const int* labels_len, const float* inv_denom_dev, float label_smoothing, float* loss_sum,
                  float* dlogits);

/* devel.py's per-token losses as `softmax_loss_function` of seq2seq.sequence_loss (seq2seq.py:156-163, devel.py:12-52):
 * kind 1 = mc_loss, 2 = focal_loss (gamma, 2.0 in the reference), on p = clip(softmax(logits), 1e-7, 1 - 1e-7); masked by
 * the label lengths and scaled by inv_denom like avsr_seq_loss; dlogits receives the gradient. */
int avsr_seq_loss_devel(avsr_stream_t stream, const float* logits, int T, int B, int V, const int* labels, int ldl,
                        const int* labels_len, const float* inv_denom, int kind, float gamma, float* loss_sum,
                        float* dlogits);

/* ---- Action-Unit regression head of the video encoder (encoder.py:173-189, seq2seq.py:188-190) -------------------
 * z [T,B,2] = encoder outputs @ video/dense/kernel + bias (pre-sigmoid); aus [B,T,2] as the reader delivers them
 * (batch-major payload, io_utils.py:45-46).  tf.losses.mean_squared_error(sigmoid(z), clip(aus,0,3)/3, weights =
 * sequence_mask(len)): loss_sum[0] += sum over t < len[b] of (p - y)^2; dz = 2 (p - y) p (1 - p) * scale_dev[0] there,
 * 0 past the length (scale = au_loss_weight / number of non-zero weights, a device scalar: graph replay). */
int avsr_au_loss(avsr_stream_t stream, const float* z, int T, int B, const float* aus, const int* len,
                 const float* scale_dev, float* loss_sum, float* dz);

/* ---- visual front-end (avsr/video.py resnet_cnn :143-195, per frame via cnn_layers :224-248) --------
 * tf.layers.conv2d on NHWC activations = avsr_im2col + avsr_gemm (+ bias) with the kernel variable [kh, kw, Cin, Cout]
 * read as [kh*kw*Cin, Cout]; its gradients = avsr_gemm (transposed) + avsr_colsum + avsr_col2im.  pad_top / pad_left are
 * TF's SAME padding (the extra pixel of an odd total goes to the end) or 0 for VALID; rows of `cols` = (n, oy, ox),
 * columns = (ky, kx, c).  round_out: cols stored tf32-rounded (operand of a tensor-core product).  batch_norm_relu
 * (:4-15) = avsr_bn_stats / avsr_bn_apply_train over rows = N*H*W with eps 1e-5, momentum 0.98, then avsr_relu_fwd. */
int avsr_im2col(avsr_stream_t stream, const float* x, int N, int H, int W, int C, int kh, int kw, int stride,
                int pad_top, int pad_left, int Ho, int Wo, int round_out, float* cols);
/* dx[N,H,W,C] = transpose of im2col applied to dcols (overwrites dx; gather form, deterministic) */
int avsr_col2im(avsr_stream_t stream, const float* dcols, int N, int H, int W, int C, int kh, int kw, int stride,
                int pad_top, int pad_left, int Ho, int Wo, float* dx);
/* the narrow layers (Cout = 8 or 16: nearly all pixels of the front-end) without the im2col buffer: y[N,Ho,Wo,Co] =
 * conv(x[N,H,W,Ci], w[kh*kw*Ci, Co]) + bias (bias may be NULL), exact fp32; the gradient wrt a stride-1 SAME input is the
 * same call on dy with the kernel flipped and transposed.  avsr_conv2d_wgrad: dW[kh*kw*Ci, Co] += x-patches^T dy. */
int avsr_conv2d_direct(avsr_stream_t stream, const float* x, int N, int H, int W, int Ci, const float* w,
                       const float* bias, int kh, int kw, int stride, int pad_top, int pad_left, int Ho, int Wo, int Co,
                       float* y);
int avsr_conv2d_wgrad(avsr_stream_t stream, const float* x, const float* dy, int N, int H, int W, int Ci, int kh, int kw,
                      int stride, int pad_top, int pad_left, int Ho, int Wo, int Co, float* dW);
/* The same convolutions on tensor cores (csrc/conv_mma.cu: implicit GEMM on mma.sync TF32, whole frames staged once in
 * shared memory, no im2col buffer) for Ci <= 64, Co in {8, 16, 32 k}, square kernels of 1 or 3, stride 1 or 2
 * (avsr_conv2d_tc_supported).  y = conv(x', w) (+ bias) (+ residual'), where the batch_norm_relu layers around the
 * convolution (video.py:4-15, 57-92) never materialise:
 *   in_bn  [2 Ci] (scale, shift) or NULL: x' = relu(x * scale + shift), applied while the frames are staged;
 *   residual (+ res_bn [2 Co]) or NULL: the `tf.add` of residual_block video.py:92, of the raw tensor or of its BN-ReLU;
 *   stats [2 Co] or NULL: += per-channel (sum, sum of squares) of y: the statistics pass of the NEXT batch_norm_relu;
 *   mask_u [N,Ho,Wo,Co] + mask_bn [4 Co] (scale, shift, a, b) or NULL: this call is an input gradient and the tensor it
 *     differentiates was z = relu(u * scale + shift): the result is masked by z > 0 (= d) and stats += (sum d,
 *     sum d * xhat), xhat = u * a + b - relu_bwd and the statistics pass of the BN backward, fused into the producer.
 * in_dilation = 2: x is read zero-stuffed (pixel (i, j) at (2 i, 2 j)) - with the kernel flipped / transposed and
 * pad = k - 1 - pad_fwd this is the input gradient of a stride-2 convolution (replaces avsr_gemm + avsr_col2im); with
 * in_dilation = 1 it is that of a stride-1 convolution.  Operands are tf32-rounded while staged. */
int avsr_conv2d_tc_supported(int Ci, int Co, int kh, int kw, int stride);
int avsr_conv2d_tc(avsr_stream_t stream, const float* x, int N, int H, int W, int Ci, const float* w, const float* bias,
                   int kh, int kw, int stride, int pad_top, int pad_left, int Ho, int Wo, int Co, int in_dilation,
                   const float* in_bn, const float* residual, const float* res_bn, const float* mask_u,
                   const float* mask_bn, float* stats, float* y);
/* dW[kh*kw*Ci, Co] += x'-patches^T dy on tensor cores (M = kh*kw*Ci, N = Co, K = pixels); in_bn as above;
 * dbias [Co] += column sums of dy (the bias gradient, summed while dy is staged) or NULL */
int avsr_conv2d_wgrad_tc(avsr_stream_t stream, const float* x, const float* in_bn, const float* dy, int N, int H, int W,
                         int Ci, int kh, int kw, int stride, int pad_top, int pad_left, int Ho, int Wo, int Co, float* dW,
                         float* dbias);
/* per-channel coefficients of a batch_norm_relu from the fused statistics: coef [4 C] = (scale = gamma invstd, shift =
 * beta - mean scale, a = invstd, b = -mean invstd); updates the moving statistics (momentum as tf.layers: m = m mom +
 * batch (1 - mom), biased variance).  avsr_bn_coef_eval: the same from the moving statistics (inference). */
int avsr_bn_finalize(avsr_stream_t stream, const float* sums, double count, const float* gamma, const float* beta,
                     float eps, float momentum, int C, float* moving_mean, float* moving_var, float* coef);
int avsr_bn_coef_eval(avsr_stream_t stream, const float* gamma, const float* beta, const float* moving_mean,
                      const float* moving_var, float eps, int C, float* coef);
/* du = gamma invstd (d - sum_d / n - xhat sum_dxhat / n) (+ residual): the apply pass of the batch_norm_relu backward from
 * the masked gradient d and sums2 = (sum d, sum d xhat) that avsr_conv2d_tc(mask_u, mask_bn) produced; xhat recomputed from
 * the saved BN input u */
int avsr_bn_relu_bwd_apply(avsr_stream_t stream, const float* d, const float* u, const float* coef, const float* sums2,
                           double count, const float* residual, long long rows, int C, float* du);
int avsr_relu_fwd(avsr_stream_t stream, const float* x, long long n, float* y);               /* in place allowed */
int avsr_relu_bwd(avsr_stream_t stream, const float* y, const float* dy, long long n, float* dx); /* dx = dy [y > 0] */
/* tf.nn.selu of the optional dense stack in front of an encoder (encoder.py:148-171, Dense(units, activation=selu,
 * use_bias=False)): y = 1.0507 (x > 0 ? x : 1.6733 (e^x - 1)); backward from the saved y.  In place allowed. */
int avsr_selu_fwd(avsr_stream_t stream, const float* x, long long n, float* y);
int avsr_selu_bwd(avsr_stream_t stream, const float* y, const float* dy, long long n, float* dx);

/* tf.contrib.rnn.HighwayWrapper around encoder layers > 0 (cells.py:89-90: `highway_encoder`; coupled gates):
 * carry = sigmoid(pre) with pre = x Wc + bc formed by avsr_gemm, y = x * carry + out * (1 - carry); y_op (or NULL) receives
 * the product-operand copy of y (tf32-rounded in tensor-core mode).  Backward: dx = dy * carry (the share through the
 * carry product is dpre Wc^T, a GEMM), dout = dy * (1 - carry), dpre = dy (x - out) carry (1 - carry) (operand-rounded). */
int avsr_highway_fwd(avsr_stream_t stream, const float* x, const float* pre, const float* out, long long n, float* y,
                     float* y_op);
int avsr_highway_bwd(avsr_stream_t stream, const float* dy, const float* x, const float* pre, const float* out,
                     long long n, float* dx, float* dout, float* dpre);

/* ---- optimiser (seq2seq.py:175-178, 195-257) --------------------------------- */
/* out[0] += sum x^2 */
int avsr_sumsq(avsr_stream_t stream, const float* x, long long n, float* out);
/* y += a*x */
int avsr_axpy(avsr_stream_t stream, float a, const float* x, float* y, long long n);
/* clip_by_global_norm + TF-Adam in one pass over the flat buffers; sumsq_dev[0] is the
 * squared global norm (device scalar); lr_t already contains the bias correction. */
int avsr_adam_clip_step(avsr_stream_t stream, float* params, const float* grads, float* m, float* v, long long n,
                        const float* sumsq_dev, float clip_norm, const float* lr_t_dev, float beta1, float beta2,
                        float eps, float* params_tf32 /* tf32-rounded copy kept in sync, or NULL */);

/* the optimisers of seq2seq.py:195-219 in the same single pass.  ADAM: as above.  NADAM: tf.contrib.opt.NadamOptimizer
 * (apply_adam with use_nesterov: numerator beta1*m + (1-beta1)*g).  ADAMW: tf.contrib.opt.AdamWOptimizer - decoupled
 * decay var -= weight_decay * var (not scaled by the learning rate) before the Adam update.  MOMENTUM:
 * tf.train.MomentumOptimizer(momentum = beta1, no nesterov): m = beta1*m + g; var -= lr*m (lr_t_dev holds the plain
 * learning rate; v, beta2, eps unused). */
enum { AVSR_OPT_ADAM = 0, AVSR_OPT_NADAM = 1, AVSR_OPT_ADAMW = 2, AVSR_OPT_MOMENTUM = 3 };
int avsr_optim_clip_step(avsr_stream_t stream, int kind, float* params, const float* grads, float* m, float* v,
                         long long n, const float* sumsq_dev, float clip_norm, const float* lr_t_dev, float beta1,
                         float beta2, float eps, float weight_decay, float* params_tf32);

/* ---- inference helpers (decoder_unimodal.py:176-271) -------------------------- */
/* greedy: ids[b] = argmax_v logits[b,:] (lowest index on ties) unless finished[b]; updates finished */
int avsr_greedy_pick(avsr_stream_t stream, const float* logits, int B, int V, int eos, int* finished,
                     int* sample_out /*[B] 0 if already finished*/, int* next_ids);
/* one BeamSearchDecoder step: log_softmax, finished masking, length penalty, top-k */
int avsr_beam_step(avsr_stream_t stream, const float* logits /*[B*W,V]*/, int B, int W, int V, int eos,
                   float length_penalty, float* log_probs /*[B,W] in/out*/, int* finished /*[B,W] in/out*/,
                   int* lengths /*[B,W] in/out*/, int* word_out, int* parent_out, float* score_out);
/* dst[i,:] = src[idx[i],:] for rows of width F */
int avsr_gather_rows(avsr_stream_t stream, const float* src, const int* idx, long long n, int F, float* dst);

#ifdef __cplusplus
}
#endif
#endif /* AVSR_B200_H_ */
