// Host side of the persistent AttentionWrapper(LSTMCell) layers (reference attention.py:132-191; AV-Align audio layer
// encoder.py:265-290; LAS / AV-Align decoder decoder_unimodal.py:299-352; WLAS decoder decoder_bimodal.py:227-277): picks the
// kernel family for a layer, prepares what the kernels read (fp16 keys / values, projected values, the fused recurrent
// matrix) and forms, after the recurrence, everything that is batched over all steps (attention vectors, weight gradients,
// dkeys / dvalues).  The kernels themselves:
//   attn_persist4.cu   one Luong-family mechanism, no DropoutWrapper: attention layer folded into the recurrent matrix,
//                      clusters of 4 CTAs x 8 utterances
//   attn_persist4d.cu  one mechanism under the DropoutWrapper and / or with scheduled sampling inside the kernel (two
//                      dependent products per step), Luong and Bahdanau families, any memory depth
//   attn_persist8w.cu  two Luong-family mechanisms (WLAS), clusters of 8 CTAs x 16 utterances
// (The cluster-of-8 generation of the single-mechanism kernels that used to live in this file - round 1, 15 clusters per
// GPU, tf32 / fp16 weights in shared memory - was retired in round 2: the cluster-of-4 kernels cover every shape it did.)
//
// The fold:  with att_{t-1} = [h_{t-1} | ctx_{t-1}] Wl,
//     z_t = x_t Wx + att_{t-1} Wa + h_{t-1} Wh  =  x_t Wx + h_{t-1} (Wh + Wl_h Wa) + ctx_{t-1} (Wl_c Wa)
// so the recurrent operand is [h | ctx] (K = H + Dm) against the fused matrix W' built per call by two small products.
// att_t itself (layer output for the Luong family, and the operand of the backward pass) is formed AFTER the loop by one
// batched product.  Step 0 sees att_{-1} = 0 (AttentionWrapper zero state): its recurrent term h_0 Wh is added to the
// x-projection before the launch.
#include <cuda_fp16.h>
#include <stdlib.h>

#include "../../include/avsr_b200.h"
#include "common.cuh"

namespace avsr {
namespace ap {

constexpr int H = 256;
constexpr int DM = 256;
constexpr int MAX_TM = 384;

// fp16 copy of a memory (keys / values / projected values): 8 elements per thread (two 16-byte loads, one 16-byte store);
// every caller's count is a multiple of 256 (rows of H or DM floats, 16-byte aligned work buffers)
__global__ void to_half_kernel(const float* __restrict__ src, __half* __restrict__ dst, long long n8) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  const float4 a = __ldg(reinterpret_cast<const float4*>(src) + 2 * i);
  const float4 b = __ldg(reinterpret_cast<const float4*>(src) + 2 * i + 1);
  const __half2 h0 = __floats2half2_rn(a.x, a.y), h1 = __floats2half2_rn(a.z, a.w);
  const __half2 h2 = __floats2half2_rn(b.x, b.y), h3 = __floats2half2_rn(b.z, b.w);
  uint4 o;
  o.x = *reinterpret_cast<const uint32_t*>(&h0); o.y = *reinterpret_cast<const uint32_t*>(&h1);
  o.z = *reinterpret_cast<const uint32_t*>(&h2); o.w = *reinterpret_cast<const uint32_t*>(&h3);
  reinterpret_cast<uint4*>(dst)[i] = o;
}

// all steps at once: rows = T*Bt, step of a row = row / Bt
__global__ void emit_attention_all_kernel(int T, int Bt, int At, int SW, const int* __restrict__ len,
                                          const float* __restrict__ S1, float* __restrict__ out) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)T * Bt * At) return;
  long long row = idx / At;
  int a = (int)(idx - row * At);
  int t = (int)(row / Bt), b = (int)(row - (long long)t * Bt);
  out[idx] = t < len[b] ? S1[(size_t)row * SW + a] : 0.0f;
}

// dA[t,b,:] += (t < len[b]) ? dout[t,b,:] : 0
__global__ void masked_add_kernel(float* __restrict__ dA, const float* __restrict__ dout, const int* __restrict__ len,
                                  int T, int Bt, int At, int rnd) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)T * Bt * At) return;
  long long row = idx / At;
  int t = (int)(row / Bt), b = (int)(row - (long long)t * Bt);
  float v = dA[idx] + ((dout && t < len[b]) ? dout[idx] : 0.0f);
  dA[idx] = maybe_tf32(v, rnd);
}

}  // namespace ap

int attn_persist4_launch_fwd(cudaStream_t st, int T, int B, int Tm, int scaled, const int* len, const int* mem_len,
                             float* gates, const float* Wp, const void* keys_h, const void* values_h, const float* g,
                             const float* c0, float* S, int SW, int At, float* craw, float* out, float* hc, float* align,
                             float* cT, float* hT);  // attn_persist4.cu
int attn_persist4_launch_bwd(cudaStream_t st, int T, int B, int Tm, int scaled, float grad_scale, const int* len,
                             const int* mem_len, const float* gates, const float* craw, const float* c0, const float* Wp,
                             const void* keys_h, const void* values_h, const float* g, const float* hc, const float* align,
                             const float* douthc, const float* dcT, const float* dhT, float* dZ, float* ds, float* dhc,
                             float* dg, float* dc0, float* dh0, float* dbias);  // attn_persist4.cu

int attn_persist4d_launch_fwd(cudaStream_t st, const AvsrRnnSeq* r, const void* keys_h, const void* values_h);  // attn_persist4d.cu
int attn_persist4d_launch_bwd(cudaStream_t st, const AvsrRnnSeq* r, const void* keys_h, const void* values_h);  // attn_persist4d.cu

// DropoutWrapper of the wrapped cell on: the two-product kernels of attn_persist4d.cu (no fold of the attention layer)
static bool layer_dropout(const AvsrRnnSeq* r) { return r->rng && (r->thr_in | r->thr_state | r->thr_out); }
// the two-product kernels also carry the in-kernel scheduled sampling (AvsrSampling)
static bool unfused_fwd(const AvsrRnnSeq* r) { return layer_dropout(r) || r->samp != nullptr; }

// shapes the persistent attention kernels of this file / attn_persist4*.cu cover
static bool persist_shape_ok(const AvsrRnnSeq* r) {
  if (r->n_mech != 1 || r->T <= 1) return false;
  const AvsrAttnMech& m = r->mech[0];
  if (m.kind >= AVSR_ATTN_BAHDANAU)  // Bahdanau family: two-product kernels, any memory depth (projected values)
    return r->H == 256 && m.A == 256 && m.Tm <= 384 && !r->output_attention;
  return r->H == 256 && m.A == 256 && m.Dm == 256 && m.Tm <= 384;
}
int attn_context_all(cudaStream_t st, int T, int B, int Tm, int Dm, const int* seq_len, const int* mem_len,
                     const float* align, const float* values, float* ctx, int ldc);  // attention.cu
static bool wlas_shape_ok(const AvsrRnnSeq* r);
int rnn_sampling_fused(const AvsrRnnSeq* r) {
  if (r && r->rng && tensor_cores_enabled() && r->n_mech == 2 && !(r->t_begin || r->t_end || r->stepwise) &&
      !getenv("AVSR_NO_ATTN_PERSIST"))
    return wlas_shape_ok(r);  // dual-attention decoder: cluster-of-8 kernels (attn_persist8w.cu)
  if (!(r && r->rng && tensor_cores_enabled() && persist_shape_ok(r) &&
        !getenv("AVSR_NO_ATTN_PERSIST") && !(r->t_begin || r->t_end || r->stepwise)))
    return 0;
  // Luong family: the wrapper emits the attention vector; Bahdanau family (persist_shape_ok): the cell output
  return r->mech[0].kind >= AVSR_ATTN_BAHDANAU ? !getenv("AVSR_NO_BAHDANAU_PERSIST") : r->output_attention != 0;
}

size_t attn_persist_work_floats(int B, int H, int Dm, int Tm) {
  // fused weights [(H+Dm),4H] + product scratch [H,4H] + fp16 keys / values + (Bahdanau family) the projected memory
  // values PV = values Wl_c [Tm*B, H] in fp32 before their fp16 copy takes the place of the values
  return (size_t)(H + Dm) * 4 * H + (size_t)H * 4 * H + ((size_t)Tm * B * (H + Dm) + 1) / 2 + 64 + (size_t)Tm * B * H;
}
static float* pv_scratch(float* scratch, int B, int H, int Dm, int Tm) {
  return scratch + (size_t)(H + Dm) * 4 * H + (size_t)H * 4 * H + ((size_t)Tm * B * (H + Dm) + 1) / 2 + 64;
}

// Forward of a single-mechanism Bahdanau-family layer (any memory depth) on the two-product kernels
static int bahdanau_persist_fwd(cudaStream_t st, const AvsrRnnSeq* r, float* scratch) {
  using namespace ap;
  const AvsrAttnMech& m = r->mech[0];
  const int T = r->T, B = r->B, At = m.A, SW = At + H, HD = H + m.Dm;
  float* tmp = scratch + (size_t)HD * 4 * H;
  __half* keys_h = reinterpret_cast<__half*>(tmp + (size_t)H * 4 * H);
  __half* pv_h = keys_h + (size_t)m.Tm * B * H;
  float* pv = pv_scratch(scratch, B, H, m.Dm, m.Tm);
  // step 0: att_{-1} = 0, so only h_0 Wh enters
  AVSR_TRY(gemm(st, 0, 0, B, 4 * H, H, r->S + At, SW, r->Wrec + (size_t)At * 4 * H, 4 * H, r->gates, 4 * H, 1.0f, nullptr));
  // PV = values Wl_c: the context half of the attention layer, once per batch
  AVSR_TRY(gemm(st, 0, 0, m.Tm * B, At, m.Dm, m.values_op ? m.values_op : m.values, m.Dm, m.Wl + (size_t)H * At, At, pv, At,
                0.0f, nullptr));
  const long long nk = (long long)m.Tm * B * H;
  AVSR_LAUNCH(to_half_kernel, cdiv(nk / 8, 256), 256, 0, st, m.keys, keys_h, nk / 8);
  AVSR_LAUNCH(to_half_kernel, cdiv(nk / 8, 256), 256, 0, st, pv, pv_h, nk / 8);
  AVSR_TRY(attn_persist4d_launch_fwd(st, r, keys_h, pv_h));
  // the true contexts of every step (parity probe; operand of the attention-layer weight gradient)
  return attn_context_all(st, T, B, m.Tm, m.Dm, r->len, m.mem_len, m.align, m.values, m.hc + H, HD);
}

// Forward of a single-mechanism Luong-family attention layer with the persistent kernel.  Returns -1 when
// the shape is not supported (caller falls back to the per-step path).
int attn_persist_fwd(cudaStream_t st, const AvsrRnnSeq* r, float* scratch) {
  using namespace ap;
  if (r->n_mech != 1 || r->T <= 0) return -1;
  const AvsrAttnMech& m = r->mech[0];
  if (m.kind > AVSR_ATTN_SCALED_LUONG) {
    if (!persist_shape_ok(r) || getenv("AVSR_NO_BAHDANAU_PERSIST")) return -1;
    return bahdanau_persist_fwd(st, r, scratch);
  }
  if (r->H != H || m.A != H || m.Dm != DM || m.Tm > MAX_TM) return -1;
  if (unfused_fwd(r) && !r->output_attention) return -1;
  const int T = r->T, B = r->B, At = m.A, SW = At + H;
  float* Wp = scratch;
  float* tmp = Wp + (size_t)(H + DM) * 4 * H;
  __half* keys_h = reinterpret_cast<__half*>(tmp + (size_t)H * 4 * H);
  __half* values_h = keys_h + (size_t)m.Tm * B * H;
  // step 0: att_{-1} = 0, so only h_0 Wh enters (exact AttentionWrapper zero-state semantics)
  AVSR_TRY(gemm(st, 0, 0, B, 4 * H, H, r->S + At, SW, r->Wrec + (size_t)At * 4 * H, 4 * H, r->gates, 4 * H, 1.0f, nullptr));
  if (unfused_fwd(r)) {
    // DropoutWrapper on (or scheduled sampling inside the recurrence): no fused matrix; the kernel forms the attention vectors itself and writes `out` (attention
    // vectors, zero past the length) and the state rows [a (.) m_in | hs]
    const long long nk = (long long)m.Tm * B * H, nv = (long long)m.Tm * B * DM;
    AVSR_LAUNCH(to_half_kernel, cdiv(nk / 8, 256), 256, 0, st, m.keys, keys_h, nk / 8);
    AVSR_LAUNCH(to_half_kernel, cdiv(nv / 8, 256), 256, 0, st, m.values, values_h, nv / 8);
    return attn_persist4d_launch_fwd(st, r, keys_h, values_h);
  }
  // fused recurrent matrix W' = [Wh + Wl_h Wa ; Wl_c Wa]
  AVSR_CHECK_CUDA(cudaMemcpyAsync(Wp, r->Wrec + (size_t)At * 4 * H, (size_t)H * 4 * H * sizeof(float),
                                  cudaMemcpyDeviceToDevice, st));
  AVSR_TRY(gemm(st, 0, 0, H, 4 * H, At, m.Wl, m.A, r->Wrec, 4 * H, Wp, 4 * H, 1.0f, nullptr));
  AVSR_TRY(gemm(st, 0, 0, DM, 4 * H, At, m.Wl + (size_t)H * m.A, m.A, r->Wrec, 4 * H, Wp + (size_t)H * 4 * H, 4 * H, 0.0f,
                nullptr));
  const long long nk = (long long)m.Tm * B * H, nv = (long long)m.Tm * B * DM;
  AVSR_LAUNCH(to_half_kernel, cdiv(nk / 8, 256), 256, 0, st, m.keys, keys_h, nk / 8);
  AVSR_LAUNCH(to_half_kernel, cdiv(nv / 8, 256), 256, 0, st, m.values, values_h, nv / 8);
  AVSR_TRY(attn_persist4_launch_fwd(st, T, B, m.Tm, m.kind == AVSR_ATTN_SCALED_LUONG, r->len, m.mem_len, r->gates, Wp, keys_h,
                                    values_h, m.g, r->c0, r->S, SW, At, r->craw, r->out, m.hc, m.align, r->cT, r->hT));
  // attention vectors of all steps in one product: S[1:, :, :At] = [h | ctx] Wl (tf32-rounded operand rows)
  AVSR_TRY(gemm(st, 0, 0, T * B, At, H + DM, m.hc, H + DM, m.Wl, m.A, r->S + (size_t)B * SW, SW, 0.0f, nullptr, 1));
  if (r->output_attention)  // layer output = attention vectors, zero past the length
    AVSR_LAUNCH(emit_attention_all_kernel, cdiv((long long)T * B * At, 256), 256, 0, st, T, B, At, SW, r->len,
                r->S + (size_t)B * SW, r->out);
  return 0;
}

// Backward counterpart.  `scratch` must still hold what attn_persist_fwd left there (fused matrix, fp16
// keys / values): the same descriptor / work buffer, forward first.  Fills dZ, ds, dhc(ctx), dA; accumulates
// dWrec, dWl, dkeys, dvalues, dg.  Returns -1 when unsupported.
int attn_outer(cudaStream_t st, int T, int B, int Tm, int C, const int* seq_len, const float* w, const float* x,
               int ldx, const float* scale, float* out);  // attention.cu

int attn_persist4d_launch_bahd_bwd(cudaStream_t st, const AvsrRnnSeq* r, const void* keys_h, const void* pv_h);  // attn_persist4d.cu
int attn_bahdanau_post(cudaStream_t st, int T, int B, int Tm, int A, const int* seq_len, const int* mem_len,
                       const float* ds, const float* pq, const float* keys, const float* v, const float* bias,
                       float* dkeys, float* dv, float* dbias);  // attention.cu

// Backward of a single-mechanism Bahdanau-family layer on the two-product kernels.  Needs only the saved activations
// (gates, craw, S, hc, pq, align): the fp16 keys and the projected values are rebuilt here.
static int bahdanau_persist_bwd(cudaStream_t st, const AvsrRnnSeq* r, float* scratch) {
  using namespace ap;
  const AvsrAttnMech& m = r->mech[0];
  const int T = r->T, B = r->B, At = m.A, SW = At + H, HD = H + m.Dm;
  AVSR_REQUIRE(m.dpq && m.ds && m.dhc && r->dA, "rnn bwd: Bahdanau scratch (dpq / ds / dhc / dA) missing");
  float* tmp = scratch + (size_t)HD * 4 * H;
  __half* keys_h = reinterpret_cast<__half*>(tmp + (size_t)H * 4 * H);
  __half* pv_h = keys_h + (size_t)m.Tm * B * H;
  float* pv = pv_scratch(scratch, B, H, m.Dm, m.Tm);
  AVSR_TRY(gemm(st, 0, 0, m.Tm * B, At, m.Dm, m.values_op ? m.values_op : m.values, m.Dm, m.Wl + (size_t)H * At, At, pv, At,
                0.0f, nullptr));
  const long long nk = (long long)m.Tm * B * H;
  AVSR_LAUNCH(to_half_kernel, cdiv(nk / 8, 256), 256, 0, st, m.keys, keys_h, nk / 8);
  AVSR_LAUNCH(to_half_kernel, cdiv(nk / 8, 256), 256, 0, st, pv, pv_h, nk / 8);
  // rows of finished steps stay zero: dWq sums over all of them
  AVSR_CHECK_CUDA(cudaMemsetAsync(m.dpq, 0, (size_t)T * B * At * sizeof(float), st));
  AVSR_TRY(attn_persist4d_launch_bahd_bwd(st, r, keys_h, pv_h));
  if (r->dh0)  // dh_0 += dz_0 Wh^T (the kernel stops before the product of step 0)
    AVSR_TRY(gemm(st, 0, 1, B, H, 4 * H, r->dZ, 4 * H, r->Wrec + (size_t)At * 4 * H, 4 * H, r->dh0, H, 1.0f, nullptr));
  // parameter gradients and per-utterance accumulations, all batched
  AVSR_TRY(gemm(st, 1, 0, SW, 4 * H, T * B, r->S, SW, r->dZ, 4 * H, r->dWrec, 4 * H, 1.0f, nullptr));
  AVSR_TRY(gemm(st, 1, 0, HD, At, T * B, m.hc, HD, r->dA, At, m.dWl, At, 1.0f, nullptr));
  // dctx_t = da_t Wl_c^T for every step -> dvalues (through the contexts)
  AVSR_TRY(gemm(st, 0, 1, T * B, m.Dm, At, r->dA, At, m.Wl + (size_t)H * At, At, m.dhc + H, HD, 0.0f, nullptr));
  AVSR_TRY(attn_outer(st, T, B, m.Tm, m.Dm, r->len, m.align, m.dhc + H, HD, nullptr, m.dvalues));
  AVSR_TRY(gemm(st, 1, 0, H, At, T * B, m.hc, HD, m.dpq, At, m.dWq, At, 1.0f, nullptr));
  return attn_bahdanau_post(st, T, B, m.Tm, At, r->len, m.mem_len, m.ds, m.pq, m.keys, m.v, m.bias, m.dkeys, m.dv, m.dbias);
}

// ---------------------------------------------------------------------------------------------------------------------
// Dual-attention decoder (WLAS, decoder_bimodal.py:179-277) on the cluster-of-8 kernels of attn_persist8w.cu: two
// Luong-family mechanisms, H = A = 256, any memory depth (the values are projected through the context half of each
// attention layer once per batch), DropoutWrapper and in-kernel scheduled sampling included.
// ---------------------------------------------------------------------------------------------------------------------
int wlas_persist8_launch_fwd(cudaStream_t st, const AvsrRnnSeq* r, const void* const* keys_h, const void* const* pv_h);  // attn_persist8w.cu
int wlas_persist8_launch_bwd(cudaStream_t st, const AvsrRnnSeq* r, const void* const* keys_h, const void* const* pv_h);  // attn_persist8w.cu

static bool wlas_shape_ok(const AvsrRnnSeq* r) {
  if (r->n_mech != 2 || r->T <= 1 || r->H != ap::H || !r->output_attention) return false;
  for (int k = 0; k < 2; ++k) {
    const AvsrAttnMech& m = r->mech[k];
    if (m.kind > AVSR_ATTN_SCALED_LUONG || m.A != ap::H || m.Tm > ap::MAX_TM) return false;
  }
  return !getenv("AVSR_NO_WLAS_PERSIST");
}
// fp16 keys | fp16 projected values | fp32 projected values, per mechanism
size_t wlas_persist_work_floats(int B, int H, int Tm) { return 2 * ((size_t)2 * Tm * B * H + 64); }
struct WlasScratch {
  const void* keys_h[2];
  const void* pv_h[2];
};
// (re)builds the fp16 memories in `scratch`: forward and backward both call it, so the backward depends on nothing but
// the saved activations
static int wlas_prepare(cudaStream_t st, const AvsrRnnSeq* r, float* scratch, WlasScratch* ws) {
  using namespace ap;
  const int B = r->B;
  float* o = scratch;
  for (int k = 0; k < 2; ++k) {
    const AvsrAttnMech& m = r->mech[k];
    const long long nk = (long long)m.Tm * B * H;
    __half* keys_h = reinterpret_cast<__half*>(o);
    __half* pv_h = keys_h + nk;
    float* pv = o + nk + 32;
    o = pv + nk + 32;
    // PV_k = values_k Wl_c,k: the context half of the attention layer, once per batch
    AVSR_TRY(gemm(st, 0, 0, m.Tm * B, H, m.Dm, m.values_op ? m.values_op : m.values, m.Dm, m.Wl + (size_t)H * H, H, pv, H, 0.0f,
                  nullptr));
    AVSR_LAUNCH(to_half_kernel, cdiv(nk / 8, 256), 256, 0, st, m.keys, keys_h, nk / 8);
    AVSR_LAUNCH(to_half_kernel, cdiv(nk / 8, 256), 256, 0, st, pv, pv_h, nk / 8);
    ws->keys_h[k] = keys_h;
    ws->pv_h[k] = pv_h;
  }
  return 0;
}

int wlas_persist_fwd(cudaStream_t st, const AvsrRnnSeq* r, float* scratch) {
  using namespace ap;
  if (!wlas_shape_ok(r)) return -1;
  const int T = r->T, B = r->B, At = 2 * H, SW = At + H;
  WlasScratch ws;
  AVSR_TRY(wlas_prepare(st, r, scratch, &ws));
  // step 0: att_{-1} = 0, so only h_0 Wh enters
  AVSR_TRY(gemm(st, 0, 0, B, 4 * H, H, r->S + At, SW, r->Wrec + (size_t)At * 4 * H, 4 * H, r->gates, 4 * H, 1.0f, nullptr));
  AVSR_TRY(wlas_persist8_launch_fwd(st, r, ws.keys_h, ws.pv_h));
  // the true contexts of every step (parity probe; operand of the attention-layer weight gradients)
  for (int k = 0; k < 2; ++k) {
    const AvsrAttnMech& m = r->mech[k];
    AVSR_TRY(attn_context_all(st, T, B, m.Tm, m.Dm, r->len, m.mem_len, m.align, m.values, m.hc + H, H + m.Dm));
  }
  return 0;
}

int wlas_persist_bwd(cudaStream_t st, const AvsrRnnSeq* r, float* scratch) {
  using namespace ap;
  if (!wlas_shape_ok(r)) return -1;
  const int T = r->T, B = r->B, At = 2 * H, SW = At + H;
  WlasScratch ws;
  AVSR_TRY(wlas_prepare(st, r, scratch, &ws));
  AVSR_TRY(wlas_persist8_launch_bwd(st, r, ws.keys_h, ws.pv_h));
  if (r->dh0)  // dh_0 += dz_0 Wh^T (the kernel stops before the product of step 0)
    AVSR_TRY(gemm(st, 0, 1, B, H, 4 * H, r->dZ, 4 * H, r->Wrec + (size_t)At * 4 * H, 4 * H, r->dh0, H, 1.0f, nullptr));
  // parameter gradients and per-utterance accumulations, all batched
  AVSR_TRY(gemm(st, 1, 0, SW, 4 * H, T * B, r->S, SW, r->dZ, 4 * H, r->dWrec, 4 * H, 1.0f, nullptr));
  for (int k = 0; k < 2; ++k) {
    const AvsrAttnMech& m = r->mech[k];
    const int HD = H + m.Dm;
    AVSR_REQUIRE(m.dhc && m.ds, "rnn bwd: mechanism scratch ds / dhc missing");
    const float* dAk = r->dA + (size_t)k * H;
    AVSR_TRY(gemm(st, 1, 0, HD, H, T * B, m.hc, HD, dAk, At, m.dWl, H, 1.0f, nullptr));
    // dctx_t = da_t Wl_c^T for every step -> dvalues (through the contexts)
    AVSR_TRY(gemm(st, 0, 1, T * B, m.Dm, H, dAk, At, m.Wl + (size_t)H * H, H, m.dhc + H, HD, 0.0f, nullptr));
    AVSR_TRY(attn_outer(st, T, B, m.Tm, m.Dm, r->len, m.align, m.dhc + H, HD, nullptr, m.dvalues));
    AVSR_TRY(attn_outer(st, T, B, m.Tm, H, r->len, m.ds, m.hc, HD, m.kind == AVSR_ATTN_SCALED_LUONG ? m.g : nullptr, m.dkeys));
  }
  return 0;
}

int attn_persist_bwd(cudaStream_t st, const AvsrRnnSeq* r, float* scratch, bool unfused) {
  using namespace ap;
  if (r->n_mech != 1 || r->T <= 1) return -1;
  const AvsrAttnMech& m = r->mech[0];
  if (m.kind > AVSR_ATTN_SCALED_LUONG) {
    if (!persist_shape_ok(r) || getenv("AVSR_NO_BAHDANAU_PERSIST") ||
        getenv("AVSR_NO_BAHDANAU_PERSIST_BWD"))
      return -1;
    return bahdanau_persist_bwd(st, r, scratch);
  }
  if (r->H != H || m.A != H || m.Dm != DM || m.Tm > MAX_TM) return -1;
  unfused = unfused || unfused_fwd(r);
  if (unfused && !r->output_attention) return -1;
  const int T = r->T, B = r->B, At = m.A, SW = At + H, HD = H + DM;
  const bool oa = r->output_attention != 0;
  float* Wp = scratch;
  float* tmp = Wp + (size_t)HD * 4 * H;
  __half* keys_h = reinterpret_cast<__half*>(tmp + (size_t)H * 4 * H);
  __half* values_h = keys_h + (size_t)m.Tm * B * H;
  if (unfused) {
    // two-product kernel: forms dZ, ds, dhc (ctx columns), dA itself; the fp16 memories are rebuilt here so that the
    // backward does not depend on which forward path ran (a ranged step-wise forward leaves the same activations)
    AVSR_REQUIRE(r->dA != nullptr, "rnn bwd: dA scratch missing");
    const long long nk = (long long)m.Tm * B * H, nv = (long long)m.Tm * B * DM;
    AVSR_LAUNCH(to_half_kernel, cdiv(nk / 8, 256), 256, 0, st, m.keys, keys_h, nk / 8);
    AVSR_LAUNCH(to_half_kernel, cdiv(nv / 8, 256), 256, 0, st, m.values, values_h, nv / 8);
    AVSR_CHECK_CUDA(cudaMemsetAsync(m.dhc, 0, (size_t)T * B * HD * sizeof(float), st));
    AVSR_TRY(attn_persist4d_launch_bwd(st, r, keys_h, values_h));
    if (r->dh0)  // dh_0 += dz_0 Wh^T (the kernel stops before the product of step 0)
      AVSR_TRY(gemm(st, 0, 1, B, H, 4 * H, r->dZ, 4 * H, r->Wrec + (size_t)At * 4 * H, 4 * H, r->dh0, H, 1.0f, nullptr));
    AVSR_TRY(gemm(st, 1, 0, SW, 4 * H, T * B, r->S, SW, r->dZ, 4 * H, r->dWrec, 4 * H, 1.0f, nullptr));
    AVSR_TRY(gemm(st, 1, 0, HD, At, T * B, m.hc, HD, r->dA, At, m.dWl, m.A, 1.0f, nullptr));
    AVSR_TRY(attn_outer(st, T, B, m.Tm, DM, r->len, m.align, m.dhc + H, HD, nullptr, m.dvalues));
    AVSR_TRY(attn_outer(st, T, B, m.Tm, At, r->len, m.ds, m.hc, HD, m.kind == AVSR_ATTN_SCALED_LUONG ? m.g : nullptr, m.dkeys));
    return 0;
  }
  // (dout Wl^T) for every step: the part of d[h | ctx] that does not depend on the recurrence
  // It is written into m.dhc itself: the kernel reads an entry and then overwrites the ctx columns with the
  // total dctx_t (same thread, same address).
  const float* dhc_in = nullptr;
  if (oa && r->dout) {
    AVSR_TRY(gemm(st, 0, 1, T * B, HD, At, r->dout, At, m.Wl, m.A, m.dhc, HD, 0.0f, nullptr));
    dhc_in = m.dhc;
  } else {
    AVSR_CHECK_CUDA(cudaMemsetAsync(m.dhc, 0, (size_t)T * B * HD * sizeof(float), st));
  }
  const int scaled = m.kind == AVSR_ATTN_SCALED_LUONG;
  AVSR_TRY(attn_persist4_launch_bwd(st, T, B, m.Tm, scaled, r->grad_scale > 0.0f ? r->grad_scale : 1.0f, r->len, m.mem_len,
                                    r->gates, r->craw, r->c0, Wp, keys_h, values_h, m.g, m.hc, m.align, dhc_in, r->dcT, r->dhT,
                                    r->dZ, m.ds, m.dhc, m.dg, r->dc0, r->dh0, r->dbias));
  if (r->dh0)  // dh_0 += dz_0 Wh^T (un-fused: the zero attention state of step 0)
    AVSR_TRY(gemm(st, 0, 1, B, H, 4 * H, r->dZ, 4 * H, r->Wrec + (size_t)At * 4 * H, 4 * H, r->dh0, H, 1.0f, nullptr));
  // dA_t = dz_{t+1} Wa^T (+ dout_t, masked): gradient wrt the attention vectors, for dWl
  AVSR_CHECK_CUDA(cudaMemsetAsync(r->dA + (size_t)(T - 1) * B * At, 0, (size_t)B * At * sizeof(float), st));
  AVSR_TRY(gemm(st, 0, 1, (T - 1) * B, At, 4 * H, r->dZ + (size_t)B * 4 * H, 4 * H, r->Wrec, 4 * H, r->dA, At, 0.0f, nullptr));
  AVSR_LAUNCH(masked_add_kernel, cdiv((long long)T * B * At, 256), 256, 0, st, r->dA, oa ? r->dout : nullptr, r->len, T, B,
              At, tensor_cores_enabled());
  // parameter gradients and per-utterance accumulations, all batched
  AVSR_TRY(gemm(st, 1, 0, SW, 4 * H, T * B, r->S, SW, r->dZ, 4 * H, r->dWrec, 4 * H, 1.0f, nullptr));
  AVSR_TRY(gemm(st, 1, 0, HD, At, T * B, m.hc, HD, r->dA, At, m.dWl, m.A, 1.0f, nullptr));
  AVSR_TRY(attn_outer(st, T, B, m.Tm, DM, r->len, m.align, m.dhc + H, HD, nullptr, m.dvalues));
  AVSR_TRY(attn_outer(st, T, B, m.Tm, At, r->len, m.ds, m.hc, HD, scaled ? m.g : nullptr, m.dkeys));
  return 0;
}

}  // namespace avsr
