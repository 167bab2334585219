// Recurrent sequence op: dynamic_rnn / dynamic_decode over LSTMCell, optionally
// wrapped by an AttentionWrapper with 1-2 mechanisms.  See include/avsr_b200.h.
//
// Step structure (reference call sites: cells.py:14-18, attention.py:132-191,
// encoder.py:265-290, decoder_unimodal.py:299-352, decoder_bimodal.py:227-277):
//   rec    = [attention_{t-1} | h_{t-1}] @ Wrec                 (gemm)
//   i,j,f,o= act(gates_t + rec); c,h update with length masking  (lstm_point_fwd)
//   per mechanism: (pq = h @ Wq) ; score -> softmax -> context   (attn_fwd)
//                  attention_m = [h | ctx] @ Wl                  (gemm)
#include <stdlib.h>

#include "../../include/avsr_b200.h"
#include "common.cuh"

namespace avsr {

// --------------------------------------------------------------------------- //
// LSTM pointwise
// --------------------------------------------------------------------------- //
struct HcDst {
  float* p[2];
  int ld[2];
};

// DropoutWrapper masks applied inside the loop (see AvsrRnnSeq.rng): streams +0 attention part of the cell input,
// +1 recurrent state h, +2 cell output.  thr == 0 switches a mask off.
struct Drop {
  const uint32_t* rng;
  uint32_t stream, thr_in, thr_state, thr_out;
  float inv_in, inv_state, inv_out;
};
static Drop drop_of(const AvsrRnnSeq* r) {
  Drop d = {r->rng, r->drop_stream, 0u, 0u, 0u, 1.0f, 1.0f, 1.0f};
  if (r->rng) {
    d.thr_in = r->thr_in;
    d.thr_state = r->thr_state;
    d.thr_out = r->thr_out;
    d.inv_in = inv_keep_of(d.thr_in);
    d.inv_state = inv_keep_of(d.thr_state);
    d.inv_out = inv_keep_of(d.thr_out);
  }
  return d;
}
static bool has_dropout(const AvsrRnnSeq* r) { return r->rng && (r->thr_in | r->thr_state | r->thr_out); }
// Step ranges (scheduled sampling) advance one step per call; everything else tries the persistent kernels first (under
// a DropoutWrapper: lstm_persist4.cu DROP variants, attn_persist4d.cu two-product kernels).
static bool stepwise_only(const AvsrRnnSeq* r) { return r->stepwise || r->t_begin || r->t_end; }

__global__ void lstm_point_fwd_kernel(int t, int B, int H, float* __restrict__ gates_t, const float* __restrict__ rec,
                                      const int* __restrict__ len, const float* __restrict__ c_prev,
                                      const float* __restrict__ h_prev, int ldh_prev, float* __restrict__ c_next,
                                      float* __restrict__ h_next, int ldh_next, float* __restrict__ craw_t,
                                      float* __restrict__ out_t, HcDst hc, int rnd, Drop dr) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * H) return;
  int b = idx / H, u = idx - b * H;
  float cp = c_prev[idx];
  float hp = h_prev[(size_t)b * ldh_prev + u];
  float hn, cn, ho;  // hn: what recurs (state dropout), ho: what the cell emits (output dropout)
  if (t >= len[b]) {
    hn = hp;
    ho = hp;
    cn = cp;
    craw_t[idx] = cp;
    if (out_t) out_t[idx] = 0.0f;  // dynamic_rnn emits zeros past the sequence length
  } else {
    float* g = gates_t + (size_t)b * 4 * H;
    const float* r = rec + (size_t)b * 4 * H;
    float gi = sigmoidf_acc(g[u] + r[u]);
    float gj = tanhf_acc(g[H + u] + r[H + u]);
    float gf = sigmoidf_acc(g[2 * H + u] + r[2 * H + u] + 1.0f);  // forget_bias = 1.0
    float go = sigmoidf_acc(g[3 * H + u] + r[3 * H + u]);
    float cr = gf * cp + gi * gj;
    cn = fminf(fmaxf(cr, -1.0f), 1.0f);  // cell_clip = 1.0 (cells.py:16)
    const float h = go * tanhf_acc(cn);
    ho = h * drop_factor(dr.rng, dr.stream + 2u, dr.thr_out, dr.inv_out, (uint32_t)t, (uint32_t)idx);
    hn = h * drop_factor(dr.rng, dr.stream + 1u, dr.thr_state, dr.inv_state, (uint32_t)t, (uint32_t)idx);
    g[u] = gi;
    g[H + u] = gj;
    g[2 * H + u] = gf;
    g[3 * H + u] = go;
    craw_t[idx] = cr;
    if (out_t) out_t[idx] = ho;
  }
  c_next[idx] = cn;
  // the state row and [h | ctx] are tensor-core operands: stored tf32-rounded (out_t keeps the exact h)
  h_next[(size_t)b * ldh_next + u] = maybe_tf32(hn, rnd);
  const float hr = maybe_tf32(ho, rnd);
  if (hc.p[0]) hc.p[0][(size_t)b * hc.ld[0] + u] = hr;
  if (hc.p[1]) hc.p[1][(size_t)b * hc.ld[1] + u] = hr;
}

// attention part of the next cell input: S_next[b, :At] *= input-dropout factor of step t_next (in place, after the
// attention vectors have been emitted as the layer output)
__global__ void drop_attention_kernel(int t_next, int B, int At, int SW, float* __restrict__ S_next, int rnd, Drop dr) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * At) return;
  int b = idx / At, a = idx - b * At;
  float* p = S_next + (size_t)b * SW + a;
  *p = maybe_tf32(*p * drop_factor(dr.rng, dr.stream, dr.thr_in, dr.inv_in, (uint32_t)t_next, (uint32_t)idx), rnd);
}

struct DhSrc {
  const float* p[4];
  int ld[4];
};

// grid over B*H.  dS_cur/dS_next rows are [datt(At) | dh(H)].
__global__ void lstm_point_bwd_kernel(int t, int B, int H, int At, const float* __restrict__ gates_t,
                                      const float* __restrict__ craw_t, const float* __restrict__ craw_prev,
                                      const float* __restrict__ c0, const int* __restrict__ len,
                                      const float* __restrict__ dout_h_t, const float* __restrict__ dS_cur,
                                      const float* __restrict__ dc_cur, DhSrc extra, float* __restrict__ dZ_t,
                                      float* __restrict__ dS_next, float* __restrict__ dc_next, int rnd, Drop dr) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * H) return;
  int b = idx / H, u = idx - b * H;
  int SW = At + H;
  float dh_in = dS_cur[(size_t)b * SW + At + u];
  float dc_in = dc_cur[idx];
  float* dz = dZ_t + (size_t)b * 4 * H;
  if (t >= len[b]) {
    dz[u] = 0.f;
    dz[H + u] = 0.f;
    dz[2 * H + u] = 0.f;
    dz[3 * H + u] = 0.f;
    dS_next[(size_t)b * SW + At + u] = dh_in;
    dc_next[idx] = dc_in;
  } else {
    float dho = dout_h_t ? dout_h_t[idx] : 0.0f;  // wrt the emitted (output-dropped) h
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (extra.p[k]) dho += extra.p[k][(size_t)b * extra.ld[k] + u];
    const float dh = dh_in * drop_factor(dr.rng, dr.stream + 1u, dr.thr_state, dr.inv_state, (uint32_t)t, (uint32_t)idx) +
                     dho * drop_factor(dr.rng, dr.stream + 2u, dr.thr_out, dr.inv_out, (uint32_t)t, (uint32_t)idx);
    const float* g = gates_t + (size_t)b * 4 * H;
    float gi = g[u], gj = g[H + u], gf = g[2 * H + u], go = g[3 * H + u];
    float cr = craw_t[idx];
    float c = fminf(fmaxf(cr, -1.0f), 1.0f);
    float tc = tanhf_acc(c);
    float cp = craw_prev ? fminf(fmaxf(craw_prev[idx], -1.0f), 1.0f) : (c0 ? c0[idx] : 0.0f);
    float dct = dc_in + dh * go * (1.0f - tc * tc);
    float dcr = (cr >= -1.0f && cr <= 1.0f) ? dct : 0.0f;
    // dZ only feeds tensor-core products (dZ Wrec^T, S^T dZ, x^T dZ, dZ Wx^T): stored tf32-rounded
    dz[u] = maybe_tf32(dcr * gj * gi * (1.0f - gi), rnd);
    dz[H + u] = maybe_tf32(dcr * gi * (1.0f - gj * gj), rnd);
    dz[2 * H + u] = maybe_tf32(dcr * cp * gf * (1.0f - gf), rnd);
    dz[3 * H + u] = maybe_tf32(dh * tc * go * (1.0f - go), rnd);
    dS_next[(size_t)b * SW + At + u] = 0.0f;
    dc_next[idx] = dcr * gf;
  }
  // attention part of dS_next: zero (filled by the Wrec^T product afterwards)
  for (int a = u; a < At; a += H) dS_next[(size_t)b * SW + a] = 0.0f;
}

// dA_t[b,:] = mask * (dS_cur.att + (oa ? dout_t : 0))
__global__ void attn_bwd_prep_kernel(int t, int B, int At, int SW, const int* __restrict__ len,
                                     const float* __restrict__ dS_cur, const float* __restrict__ dout_att_t,
                                     float* __restrict__ dA_t, int rnd, Drop dr) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * At) return;
  int b = idx / At, a = idx - b * At;
  float v = 0.0f;
  if (t < len[b])  // dS_cur.att is wrt the dropped attention input of step t + 1
    v = dS_cur[(size_t)b * SW + a] * drop_factor(dr.rng, dr.stream, dr.thr_in, dr.inv_in, (uint32_t)(t + 1), (uint32_t)idx) +
        (dout_att_t ? dout_att_t[idx] : 0.0f);
  dA_t[idx] = maybe_tf32(v, rnd);
}

// out_t[b,:] = t < len[b] ? S_next[b, :At] : 0
__global__ void emit_attention_kernel(int t, int B, int At, int SW, const int* __restrict__ len,
                                      const float* __restrict__ S_next, float* __restrict__ out_t) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * At) return;
  int b = idx / At, a = idx - b * At;
  out_t[idx] = t < len[b] ? S_next[(size_t)b * SW + a] : 0.0f;
}

__global__ void copy2d_kernel(const float* __restrict__ src, int lds, float* __restrict__ dst, int ldd, int rows,
                              int cols) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)rows * cols) return;
  size_t r = i / cols, c = i - r * cols;
  dst[r * ldd + c] = src ? src[r * lds + c] : 0.0f;
}

// attention kernels live in attention.cu
size_t attn_fwd_smem(int Tm, int A);
int attn_fwd(cudaStream_t st, int kind, int Tm, int B, int Dm, int A, const float* q, int ldq, const float* keys,
             const float* values, const int* mem_len, const float* v, const float* g, const float* bias,
             float* align_t, float* ctx_out, int ldctx, int rnd);
int attn_bwd_step(cudaStream_t st, int kind, int t, const int* seq_len, int Tm, int B, int Dm, int A, const float* q,
                  int ldq, const float* keys, const float* values, const int* mem_len, const float* v, const float* g,
                  const float* bias, const float* align_t, const float* dctx, int lddctx, float* dq_out, int lddq,
                  float* ds_out, float* dg, int rnd);
int attn_outer(cudaStream_t st, int T, int B, int Tm, int C, const int* seq_len, const float* w, const float* x,
               int ldx, const float* scale, float* out);
int attn_bahdanau_post(cudaStream_t st, int T, int B, int Tm, int A, const int* seq_len, const int* mem_len,
                       const float* ds, const float* pq, const float* keys, const float* v, const float* bias,
                       float* dkeys, float* dv, float* dbias);

// --------------------------------------------------------------------------- //
// host orchestration
// --------------------------------------------------------------------------- //
size_t attn_persist_work_floats(int B, int H, int Dm, int Tm);                 // attn_persist.cu
size_t wlas_persist_work_floats(int B, int H, int Tm);                         // attn_persist.cu
int wlas_persist_fwd(cudaStream_t st, const AvsrRnnSeq* r, float* scratch);    // attn_persist.cu (dual attention, clusters of 8)
int wlas_persist_bwd(cudaStream_t st, const AvsrRnnSeq* r, float* scratch);    // attn_persist.cu
int attn_persist_fwd(cudaStream_t st, const AvsrRnnSeq* r, float* scratch);    // attn_persist.cu
int attn_persist_bwd(cudaStream_t st, const AvsrRnnSeq* r, float* scratch, bool unfused);  // attn_persist.cu

struct WorkLayout {
  size_t rec, cbuf, dS, dcbuf, dHC, dq, persist, total;
};
static WorkLayout work_layout(int B, int H, int At, int maxHD, int maxA, int maxTm) {
  WorkLayout w;
  size_t o = 0;
  auto take = [&](size_t n) { size_t r = o; o += (n + 3) & ~(size_t)3; return r; };
  w.rec = take((size_t)B * 4 * H);
  w.cbuf = take((size_t)2 * B * H);
  w.dS = take((size_t)2 * B * (At + H));
  w.dcbuf = take((size_t)2 * B * H);
  w.dHC = take((size_t)2 * B * maxHD);
  w.dq = take((size_t)2 * B * (maxA > H ? maxA : H));
  size_t pw = maxTm > 0 ? attn_persist_work_floats(B, H, maxHD - H, maxTm) : 0;
  if (maxTm > 0 && maxA > 0 && At >= 2 * maxA) {  // two mechanisms: the dual-attention kernels keep both memories in fp16
    const size_t ww = wlas_persist_work_floats(B, H, maxTm);
    pw = pw > ww ? pw : ww;
  }
  w.persist = take(pw);
  w.total = o;
  return w;
}

static int check_common(const AvsrRnnSeq* r, int* At_out, int* maxHD, int* maxA, int* maxTm) {
  AVSR_REQUIRE(r != nullptr, "rnn: null descriptor");
  AVSR_REQUIRE(r->T >= 0 && r->B > 0 && r->H > 0, "rnn: bad dims T=%d B=%d H=%d", r->T, r->B, r->H);
  AVSR_REQUIRE(r->n_mech >= 0 && r->n_mech <= 2, "rnn: n_mech must be 0..2");
  int At = 0, hd = 0, ma = 0, mt = 0;
  for (int k = 0; k < r->n_mech; ++k) {
    const AvsrAttnMech& m = r->mech[k];
    AVSR_REQUIRE(m.kind >= 0 && m.kind <= 3, "rnn: unknown attention mechanism %d", m.kind);
    AVSR_REQUIRE(m.Tm > 0 && m.Dm > 0 && m.A > 0, "rnn: bad mechanism dims");
    if (m.kind <= AVSR_ATTN_SCALED_LUONG)
      AVSR_REQUIRE(m.A == r->H, "luong attention needs num_units == query depth (A=%d, H=%d)", m.A, r->H);
    At += m.A;
    hd = hd > r->H + m.Dm ? hd : r->H + m.Dm;
    ma = ma > m.A ? ma : m.A;
    mt = mt > m.Tm ? mt : m.Tm;
  }
  *At_out = At;
  *maxHD = hd;
  *maxA = ma;
  *maxTm = mt;
  return 0;
}


int lstm_persist_fwd(cudaStream_t st, const AvsrRnnSeq* r);   // lstm_persist.cu (tries the clusters of 4 first)
int lstm_persist4_fwd(cudaStream_t st, const AvsrRnnSeq* r);  // lstm_persist4.cu (also under dropout)

int rnn_sampling_fused(const AvsrRnnSeq* r);  // attn_persist.cu

int rnn_seq_fwd(cudaStream_t st, const AvsrRnnSeq* r) {
  int At, maxHD, maxA, maxTm;
  AVSR_TRY(check_common(r, &At, &maxHD, &maxA, &maxTm));
  const bool stepwise = stepwise_only(r);
  AVSR_REQUIRE(!r->samp || rnn_sampling_fused(r),
               "rnn: scheduled sampling inside the recurrence is not available for this layer (see avsr_rnn_sampling_fused); "
               "advance in step ranges with avsr_sched_sample instead");
  if (r->n_mech == 0 && tensor_cores_enabled() && !stepwise) {  // persistent cluster kernel (tcgen05, weights resident)
    const int rc = has_dropout(r) ? lstm_persist4_fwd(st, r) : lstm_persist_fwd(st, r);
    if (rc >= 0) return rc;
  }
  const int T = r->T, B = r->B, H = r->H, SW = At + H;
  const int rnd = tensor_cores_enabled();
  WorkLayout wl = work_layout(B, H, At, maxHD, maxA, maxTm);
  if (r->n_mech == 1 && rnd && T > 1 && !stepwise && !getenv("AVSR_NO_ATTN_PERSIST")) {  // T == 1: step-wise decoding (carried attention)
    // persistent cluster kernel for the Luong-family attention layer (AV-Align top layer, LAS decoder)
    const int rc = attn_persist_fwd(st, r, r->work + wl.persist);  // also writes `out` when it is the attention
    if (rc >= 0) return rc;
  }
  if (r->n_mech == 2 && rnd && T > 1 && !stepwise && !getenv("AVSR_NO_ATTN_PERSIST")) {
    // dual-attention (WLAS) decoder: persistent cluster-of-8 kernel
    const int rc = wlas_persist_fwd(st, r, r->work + wl.persist);
    if (rc >= 0) return rc;
  }
  float* rec = r->work + wl.rec;
  float* cbuf[2] = {r->work + wl.cbuf, r->work + wl.cbuf + (size_t)B * H};
  const int pw_grid = cdiv((long long)B * H, 256);
  const Drop dr = drop_of(r);
  const int t0 = r->t_begin, t1 = (r->t_begin || r->t_end) ? r->t_end : T;
  AVSR_REQUIRE(0 <= t0 && t0 <= t1 && t1 <= T, "rnn: bad step range [%d, %d) of %d", t0, t1, T);
  if (t0 == 0) AVSR_LAUNCH(copy2d_kernel, pw_grid, 256, 0, st, r->c0, H, cbuf[0], H, B, H);
  int cur = t0 & 1;  // the cell state of step t sits in cbuf[t & 1] (ranged calls continue where the last one stopped)
  for (int t = t0; t < t1; ++t) {
    const float* S_t = r->S + (size_t)t * B * SW;
    float* S_n = r->S + (size_t)(t + 1) * B * SW;
    float* gates_t = r->gates + (size_t)t * B * 4 * H;
    AVSR_TRY(gemm(st, 0, 0, B, 4 * H, SW, S_t, SW, r->Wrec, 4 * H, rec, 4 * H, 0.0f, nullptr));
    HcDst hc = {{nullptr, nullptr}, {0, 0}};
    for (int k = 0; k < r->n_mech; ++k) {
      hc.p[k] = r->mech[k].hc + (size_t)t * B * (H + r->mech[k].Dm);
      hc.ld[k] = H + r->mech[k].Dm;
    }
    float* out_h = (r->output_attention && r->n_mech > 0) ? nullptr : r->out + (size_t)t * B * H;
    AVSR_LAUNCH(lstm_point_fwd_kernel, pw_grid, 256, 0, st, t, B, H, gates_t, rec, r->len, cbuf[cur], S_t + At, SW,
                cbuf[cur ^ 1], S_n + At, SW, r->craw + (size_t)t * B * H, out_h, hc, rnd, dr);
    cur ^= 1;
    int off = 0;
    for (int k = 0; k < r->n_mech; ++k) {
      const AvsrAttnMech& m = r->mech[k];
      const bool luong = m.kind <= AVSR_ATTN_SCALED_LUONG;
      const float* q = hc.p[k];  // the query is the cell OUTPUT (= the state h unless dropout separates them)
      int ldq = hc.ld[k];
      if (!luong) {
        float* pq_t = m.pq + (size_t)t * B * m.A;
        AVSR_TRY(gemm(st, 0, 0, B, m.A, H, hc.p[k], hc.ld[k], m.Wq, m.A, pq_t, m.A, 0.0f, nullptr));
        q = pq_t;
        ldq = m.A;
      }
      AVSR_TRY(attn_fwd(st, m.kind, m.Tm, B, m.Dm, m.A, q, ldq, m.keys, m.values, m.mem_len, m.v, m.g, m.bias,
                        m.align + (size_t)t * B * m.Tm, hc.p[k] + H, hc.ld[k], rnd));
      AVSR_TRY(gemm(st, 0, 0, B, m.A, H + m.Dm, hc.p[k], hc.ld[k], m.Wl, m.A, S_n + off, SW, 0.0f, nullptr, 1));
      off += m.A;
    }
    if (r->output_attention && r->n_mech > 0)
      AVSR_LAUNCH(emit_attention_kernel, cdiv((long long)B * At, 256), 256, 0, st, t, B, At, SW, r->len, S_n,
                  r->out + (size_t)t * B * At);
    if (dr.thr_in && At > 0)
      AVSR_LAUNCH(drop_attention_kernel, cdiv((long long)B * At, 256), 256, 0, st, t + 1, B, At, SW, S_n, rnd, dr);
  }
  if (t1 == T) {
    if (r->cT) AVSR_LAUNCH(copy2d_kernel, pw_grid, 256, 0, st, cbuf[cur], H, r->cT, H, B, H);
    if (r->hT) AVSR_LAUNCH(copy2d_kernel, pw_grid, 256, 0, st, r->S + (size_t)T * B * SW + At, SW, r->hT, H, B, H);
  }
  return 0;
}

int lstm_persist_bwd(cudaStream_t st, const AvsrRnnSeq* r);   // lstm_persist.cu (clusters of 8)
int lstm_persist4_bwd(cudaStream_t st, const AvsrRnnSeq* r);  // lstm_persist4.cu (clusters of 4; forms dbias itself)

int rnn_seq_bwd(cudaStream_t st, const AvsrRnnSeq* r) {
  int At, maxHD, maxA, maxTm;
  AVSR_TRY(check_common(r, &At, &maxHD, &maxA, &maxTm));
  const int T = r->T, B = r->B, H = r->H, SW = At + H;
  const bool stepwise = stepwise_only(r);
  AVSR_REQUIRE(!(r->t_begin || r->t_end), "rnn bwd: step ranges are a forward-only feature");
  if (r->n_mech == 0 && tensor_cores_enabled() && !stepwise) {
    // reverse-time recurrence in one persistent cluster kernel
    int rc = lstm_persist4_bwd(st, r);
    if (rc < 0 && !has_dropout(r)) {
      rc = lstm_persist_bwd(st, r);
      if (rc == 0 && r->dbias && T > 0) AVSR_TRY(avsr_colsum((avsr_stream_t)st, r->dZ, T * B, 4 * H, 4 * H, r->dbias));
    }
    if (rc >= 0) {
      if (rc == 0 && T > 0)  // dWrec += S[0:T]^T @ dZ
        AVSR_TRY(gemm(st, 1, 0, SW, 4 * H, T * B, r->S, SW, r->dZ, 4 * H, r->dWrec, 4 * H, 1.0f, nullptr));
      return rc;
    }
  }
  const bool oa = r->output_attention && r->n_mech > 0;
  const int rnd = tensor_cores_enabled();
  WorkLayout wl = work_layout(B, H, At, maxHD, maxA, maxTm);
  if (r->n_mech == 1 && rnd && T > 1 && !getenv("AVSR_NO_ATTN_PERSIST")) {
    // After a step-wise forward (r->stepwise: scheduled sampling advanced the recurrence in ranges) only the two-product
    // kernels apply: they need nothing but the saved activations, while the folded ones need the scratch their own
    // forward left (fused matrix).
    AVSR_REQUIRE(r->mech[0].ds && r->mech[0].dhc, "rnn bwd: mechanism scratch ds / dhc missing");
    const int rc = attn_persist_bwd(st, r, r->work + wl.persist, r->stepwise != 0);
    if (rc >= 0) return rc;
  }
  if (r->n_mech == 2 && rnd && T > 1 && !getenv("AVSR_NO_ATTN_PERSIST")) {
    // (also after a step-wise forward: the kernel needs nothing but the saved activations)
    const int rc = wlas_persist_bwd(st, r, r->work + wl.persist);
    if (rc >= 0) return rc;
  }
  float* dS[2] = {r->work + wl.dS, r->work + wl.dS + (size_t)B * SW};
  float* dcb[2] = {r->work + wl.dcbuf, r->work + wl.dcbuf + (size_t)B * H};
  const int qw = maxA > H ? maxA : H;
  float* dq[2] = {r->work + wl.dq, r->work + wl.dq + (size_t)B * qw};
  const int pw_grid = cdiv((long long)B * H, 256);
  const Drop dr = drop_of(r);
  AVSR_LAUNCH(copy2d_kernel, cdiv((long long)B * SW, 256), 256, 0, st, (const float*)nullptr, 0, dS[0], SW, B, SW);
  AVSR_LAUNCH(copy2d_kernel, pw_grid, 256, 0, st, r->dhT, H, dS[0] + At, SW, B, H);
  AVSR_LAUNCH(copy2d_kernel, pw_grid, 256, 0, st, r->dcT, H, dcb[0], H, B, H);
  for (int k = 0; k < r->n_mech; ++k)
    AVSR_REQUIRE(r->mech[k].ds && r->mech[k].dhc, "rnn bwd: mechanism scratch ds / dhc missing");
  int cur = 0;
  for (int t = T - 1; t >= 0; --t) {
    float* dZ_t = r->dZ + (size_t)t * B * 4 * H;
    DhSrc extra = {{nullptr, nullptr, nullptr, nullptr}, {0, 0, 0, 0}};
    if (r->n_mech > 0) {
      float* dA_t = r->dA + (size_t)t * B * At;
      AVSR_LAUNCH(attn_bwd_prep_kernel, cdiv((long long)B * At, 256), 256, 0, st, t, B, At, SW, r->len, dS[cur],
                  (oa && r->dout) ? r->dout + (size_t)t * B * At : nullptr, dA_t, rnd, dr);
      int off = 0;
      for (int k = 0; k < r->n_mech; ++k) {
        const AvsrAttnMech& m = r->mech[k];
        const bool luong = m.kind <= AVSR_ATTN_SCALED_LUONG;
        const int HD = H + m.Dm;
        // d[h | ctx] = dA_m @ Wl^T   (kept for all steps: dvalues is formed after the loop)
        float* dHC_t = m.dhc + (size_t)t * B * HD;
        AVSR_TRY(gemm(st, 0, 1, B, HD, m.A, dA_t + off, At, m.Wl, m.A, dHC_t, HD, 0.0f, nullptr));
        const float* q = luong ? m.hc + (size_t)t * B * HD : m.pq + (size_t)t * B * m.A;
        const int ldq = luong ? HD : m.A;
        float* dq_out = luong ? dq[k] : m.dpq + (size_t)t * B * m.A;
        AVSR_TRY(attn_bwd_step(st, m.kind, t, r->len, m.Tm, B, m.Dm, m.A, q, ldq, m.keys, m.values, m.mem_len, m.v,
                               m.g, m.bias, m.align + (size_t)t * B * m.Tm, dHC_t + H, HD, dq_out, m.A,
                               m.ds + (size_t)t * B * m.Tm, m.dg, rnd));
        if (!luong) AVSR_TRY(gemm(st, 0, 1, B, H, m.A, dq_out, m.A, m.Wq, m.A, dq[k], H, 0.0f, nullptr));
        extra.p[2 * k] = dHC_t;
        extra.ld[2 * k] = HD;
        extra.p[2 * k + 1] = dq[k];
        extra.ld[2 * k + 1] = luong ? m.A : H;
        off += m.A;
      }
    }
    AVSR_LAUNCH(lstm_point_bwd_kernel, pw_grid, 256, 0, st, t, B, H, At, r->gates + (size_t)t * B * 4 * H,
                r->craw + (size_t)t * B * H, t > 0 ? r->craw + (size_t)(t - 1) * B * H : nullptr, r->c0, r->len,
                (oa || !r->dout) ? nullptr : r->dout + (size_t)t * B * H, dS[cur], dcb[cur], extra, dZ_t, dS[cur ^ 1],
                dcb[cur ^ 1], rnd, dr);
    // [datt_{t-1} | dh_{t-1}] += dZ_t @ Wrec^T
    AVSR_TRY(gemm(st, 0, 1, B, SW, 4 * H, dZ_t, 4 * H, r->Wrec, 4 * H, dS[cur ^ 1], SW, 1.0f, nullptr));
    cur ^= 1;
  }
  if (r->dh0) AVSR_LAUNCH(copy2d_kernel, pw_grid, 256, 0, st, dS[cur] + At, SW, r->dh0, H, B, H);
  if (r->dc0) AVSR_LAUNCH(copy2d_kernel, pw_grid, 256, 0, st, dcb[cur], H, r->dc0, H, B, H);
  if (T > 0) {
    if (r->dbias) AVSR_TRY(avsr_colsum((avsr_stream_t)st, r->dZ, T * B, 4 * H, 4 * H, r->dbias));
    // dWrec += S[0:T]^T @ dZ
    AVSR_TRY(gemm(st, 1, 0, SW, 4 * H, T * B, r->S, SW, r->dZ, 4 * H, r->dWrec, 4 * H, 1.0f, nullptr));
    int off = 0;
    for (int k = 0; k < r->n_mech; ++k) {
      const AvsrAttnMech& m = r->mech[k];
      const int HD = H + m.Dm;
      AVSR_TRY(gemm(st, 1, 0, HD, m.A, T * B, m.hc, HD, r->dA + off, At, m.dWl, m.A, 1.0f, nullptr));
      // per-utterance accumulations taken off the sequential path
      AVSR_TRY(attn_outer(st, T, B, m.Tm, m.Dm, r->len, m.align, m.dhc + H, HD, nullptr, m.dvalues));
      if (m.kind >= AVSR_ATTN_BAHDANAU) {
        AVSR_TRY(gemm(st, 1, 0, H, m.A, T * B, m.hc, HD, m.dpq, m.A, m.dWq, m.A, 1.0f, nullptr));
        AVSR_TRY(attn_bahdanau_post(st, T, B, m.Tm, m.A, r->len, m.mem_len, m.ds, m.pq, m.keys, m.v, m.bias, m.dkeys,
                                    m.dv, m.dbias));
      } else {
        AVSR_TRY(attn_outer(st, T, B, m.Tm, m.A, r->len, m.ds, m.hc, HD,
                            m.kind == AVSR_ATTN_SCALED_LUONG ? m.g : nullptr, m.dkeys));
      }
      off += m.A;
    }
  }
  return 0;
}

size_t rnn_work_floats(int B, int H, int At, int maxHD, int maxA, int maxTm) {
  return work_layout(B, H, At, maxHD, maxA, maxTm).total;
}

}  // namespace avsr
