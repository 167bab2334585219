// Persistent AttentionWrapper(DropoutWrapper(LSTMCell)) layer on clusters of four CTAs: the reference's DEFAULT
// training graph (cells.py:46-54 wraps every cell in DropoutWrapper(input, state, output keep = 0.9); avsr.py:51-56;
// attention.py:132-191 wraps that in the AttentionWrapper; encoder.py:265-290 AV-Align layer,
// decoder_unimodal.py:299-352 decoder).  Luong / scaled-Luong scorer, one mechanism, H = A = Dm = 256.
//
// Why a second kernel family next to attn_persist4.cu: there the attention layer is folded into the recurrent matrix
// (W' = [Wh + Wl_h Wa ; Wl_c Wa] over the operand [h | ctx]).  The DropoutWrapper's INPUT mask acts on the fed-back
// attention vector, i.e. exactly between the two factors of that fold, so the step needs two dependent products:
//
//   a_t      = [ho_t | ctx_t] Wa                       ho_t = h_t (.) m_out(t)   (the cell output: query and emitted h)
//   z_{t+1}  = gx_{t+1} + [a_t (.) m_in(t+1) | hs_t] [Wl_att ; Wh]              hs_t = h_t (.) m_state(t)
//
// Same cluster geometry as attn_persist4.cu (4 CTAs x 8 utterances; a CTA owns 64 hidden units = 256 gate rows, 64
// attention units and the attention of 2 utterances).  Per CTA:
//   * recurrent matrix [Wl_att ; Wh] restricted to its gate rows, split BY K, not by tile: the h half (K = 256) of both
//     128-row tiles in shared memory (SS products, ~68 clocks per 128 x 16 x 16 MMA: the A operand is read from shared
//     memory), the attention half of both tiles in tensor memory (columns 256..511; TS products, 8 clocks).  The slow h
//     half of step t+1 is issued as soon as hs_t has been gathered and runs behind the attention sweeps; only the fast
//     attention half sits on the critical path after the a_t all-gather (profiles/r02_ap4d_step_trace.txt: the product
//     went from 1.2 us to 0.2 us of the step);
//   * its 64 x 512 slice of Wa in tensor memory columns 128..255 as ONE 128-row A operand with K = 256: lanes 0..63
//     hold the rows that multiply ho, lanes 64..127 the rows that multiply ctx; the B operand has ho in rows 0..7 and
//     ctx in rows 8..15, so D[u][b] (lanes 0..63, columns 0..7) + D[64+u][8+b] is a_t - half the MMAs of a K = 512
//     product and exactly the tensor memory that was left;
//   * 128 accumulator columns: four accumulators (tile x K parity) for the recurrent product, four for the attention
//     product (the early h half of step t+1 is in flight while the attention product of step t runs).
// Exchanges per step (DSMEM st.async + mbarrier complete_tx, all-gathers over the 4 CTAs): {hs_t, ho_t}, ctx_t, a_t.
// Masks come from the counter-based generator (common.cuh avsr_rand_u32), computed one step ahead of their use.
#include "ap4_common.cuh"

namespace avsr {
namespace ap4 {

constexpr int AT = 256;                   // attention units (Luong: = H)
constexpr int Q_BYTES = 4 * NP * 128;     // P2 operand: [rows 0..7 ho | rows 8..15 ctx] x K = 256
constexpr int APART_FLOATS = 4 * NB * UPC;  // partial a_t planes [K group (2)][part (2)][b][u]

struct DropCfg {
  const uint32_t* rng;  // {seed, step}
  uint32_t stream, thr_in, thr_state, thr_out;
  float inv_in, inv_state, inv_out;
};
__device__ __forceinline__ float dfac(uint32_t seed, uint32_t step, uint32_t stream, uint32_t thr, float inv, uint32_t hi,
                                      uint32_t lo) {
  return (thr == 0u || avsr_rand_u32(seed, step, stream, hi, lo) < thr) ? inv : 0.0f;
}

// Step-latency trace (tools/ap4d_trace.py; compiled in only with -DAP4D_TRACE: a separate library, never the product):
// thread 0 of CTA 0 stamps clock64 at the synchronisation points of steps TR_T0 .. TR_T0 + TR_NT - 1.
#ifdef AP4D_TRACE
constexpr int TR_T0 = 100, TR_NT = 16, TR_NP = 12;
__device__ unsigned long long g_ap4d_trace[TR_NT * TR_NP];
#define AP4D_STAMP(k)                                                                         \
  do {                                                                                        \
    if (tid == 0 && blockIdx.x == 0 && t >= TR_T0 && t < TR_T0 + TR_NT)                       \
      g_ap4d_trace[(t - TR_T0) * TR_NP + (k)] = (unsigned long long)clock64();                \
  } while (0)
#else
#define AP4D_STAMP(k) do { } while (0)
#endif

// =====================================================================================================
// forward
// =====================================================================================================
struct DParams {
  int T, B, Tm;
  int scaled;
  const int* len;
  const int* mem_len;
  float* gates;          // [T,B,4H] in: x-projection (+ h0 Wh at t = 0); out: activations
  const float* Wrec;     // [(AT+H), 4H] rows [attention ; h]
  const float* Wa;       // attention_layer kernel [(H+DM), AT]
  const __half* keys;    // [Tm,B,H]
  const __half* values;  // [Tm,B,DM]
  const float* g;        // attention_g [1] or null
  const float* c0;       // [B,H] or null
  float* S;              // [(T+1),B,AT+H]; S[0] by the caller; rows 1.. = [a (.) m_in | hs], tf32-rounded
  float* craw;           // [T,B,H]
  float* out;            // [T,B,AT] attention vectors (tf32-rounded), zero past the length
  float* hc;             // [T,B,ldhc]  [ho | ctx], tf32-rounded (BAHD: only the ho columns; ldhc = H + memory depth)
  int ldhc;
  float* align;          // [T,B,Tm]
  float* cT;             // [B,H] or null
  float* hT;             // [B,H] or null
  DropCfg d;
  // BAHD instantiation (Bahdanau family, attention.py:25-42): query layer, effective v, bias (normed) or null, the
  // processed queries kept for the backward pass; `values` then holds the PROJECTED memory values Wl_c (see below) and
  // `out` [T,B,H] receives the cell output ho (the wrapper returns the cell output for this family)
  const float* Wq;       // [H, AT]
  const float* v;        // [AT]
  const float* batt;     // [AT] or null
  float* pq;             // [T,B,AT]
  // SAMPLE instantiation: ScheduledEmbeddingTrainingHelper inside the kernel (AvsrSampling, include/avsr_b200.h)
  const float* Wd;       // [AT, V]
  const float* bd;       // [V]
  const float* emb;      // [V, E]
  const float* Wx;       // [E, 4H]
  const float* bias;     // [4H]
  int* used_ids;         // [T,B]
  int* sample_ids;       // [T,B]
  float* x;              // [T,B,E]
  int V, E;
  uint32_t ss_stream, thr_p;
};

constexpr int SAMP_VP = 32;                                 // padded alphabet (V <= 32)
constexpr int SAMP_EMAX = 256;                              // E <= 256
constexpr int SAMP_SMEM = NB * UPC * 4 + UPC * SAMP_VP * 4 + CL * NB * SAMP_VP * 4 + SAMP_EMAX * 4 + NB * 4;
constexpr int BAHD_SMEM = NU * AT * 4 + NB * UPC * 4;         // processed queries of the CTA's utterances | projected contexts
constexpr size_t DFWD_SMEM = (size_t)W_BYTES + 2 * OP_BYTES + Q_BYTES + 4 * NB * UPC * 4 + APART_FLOATS * 4 +
                             NU * MAX_TM * 4 + NU * 8 * 4 + 96 + SAMP_SMEM + BAHD_SMEM + 1024;

// SAMPLE: scheduled sampling (decoder_unimodal.py:304-309) inside the recurrence.  Which (step, utterance) pairs
// are replaced is a function of the generator alone, so every CTA of the cluster knows it; for those pairs the CTAs
// reduce partial logits a_t Wd over their attention units through DSMEM, every CTA draws the same id (inverse CDF of
// the fp32 softmax, as avsr_sched_sample) and forms the x-projection of the drawn embedding for its own gate rows.
//
// BAHD: the Bahdanau scorers.  Two more things enter the step: the processed query pq_t = ho_t Wq and
// score = sum_u v_u tanh(keys_u + pq_u).  The query layer shares the attention product: the A operand in tensor memory
// holds the CTA's 64 rows of Wl_h^T in lanes 0..63 and its 64 rows of Wq^T in lanes 64..127, both multiply ho (K = 256),
// issued as soon as ho_t has been gathered; the pq slices are exchanged (all-to-all) to the owners of the utterances
// before the sweep.  The context half of the attention layer is taken out of the recurrence: the host projects the
// memory once per batch, PV = values Wl_c (any memory depth Dm), the sweep forms ctx' = sum_t a_t PV_t (256-d) and
// a_t = ho_t Wl_h + ctx'_t.  The true contexts (parity probe, operand of dWl) are formed after the loop from the saved
// alignments (attn_context_all).
template <bool SAMPLE, bool BAHD>
__global__ void __launch_bounds__(THREADS, 1) attn_lstm_persist4d_fwd_kernel(const DParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sW = base;                          // h half (K = 256) of the recurrent matrix, both 128-row tiles
  const uint32_t sOp = sW + W_BYTES;                 // two operand buffers [a (.) m_in | hs], NP rows (rows >= NB zero)
  const uint32_t sQ = sOp + 2 * OP_BYTES;            // [rows 0..7 ho | rows 8..15 ctx] x 256
  const uint32_t sAct = sQ + Q_BYTES;                // [4][NB][UPC] floats; also [NU][4][DM] partial contexts
  const uint32_t sAp = sAct + 4 * NB * UPC * 4;      // [4][NB][UPC] partial attention vectors
  const uint32_t sSc = sAp + APART_FLOATS * 4;       // [NU][MAX_TM] scores / alignments
  const uint32_t sRed = sSc + NU * MAX_TM * 4;       // [NU][8]
  const uint32_t sBar = sRed + NU * 8 * 4;           // [0] mma1 [1] mma2 [2,3] h_full[buf] [4] ctx_full [5] a_full
  const uint32_t sTmem = sBar + 64;                  // [6] logits_full (SAMPLE) [7] pq_full (BAHD)
  const uint32_t sSamp = sBar + 80;                  // SAMPLE: a_t [NB][UPC] | Wd slice [UPC][32] | partial logits [CL][NB][32] | x [256] | picks [NB]
  const uint32_t sPqA = sSamp + SAMP_SMEM;           // BAHD: pq of the CTA's utterances [NU][AT]
  const uint32_t sCxA = sPqA + NU * AT * 4;          // BAHD: projected contexts of the CTA's attention units [NB][UPC]
  const uint32_t barPq = sBar + 56;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  float* act = reinterpret_cast<float*>(gen + (sAct - base));
  float* apart = reinterpret_cast<float*>(gen + (sAp - base));
  float* sAf = reinterpret_cast<float*>(gen + (sSamp - base));
  float* sWd = sAf + NB * UPC;
  float* sLg = sWd + UPC * SAMP_VP;
  float* sX = sLg + CL * NB * SAMP_VP;
  int* sPick = reinterpret_cast<int*>(sX + SAMP_EMAX);
  float* sPq = reinterpret_cast<float*>(gen + (sPqA - base));
  float* sCx = reinterpret_cast<float*>(gen + (sCxA - base));
  const uint32_t sLgAddr = sSamp + (NB * UPC + UPC * SAMP_VP) * 4, barL = sBar + 48;
  float* sc_all = reinterpret_cast<float*>(gen + (sSc - base));
  float* part_all = act;
  float* red_all = reinterpret_cast<float*>(gen + (sRed - base));
  const uint32_t barM1 = sBar, barM2 = sBar + 8, barCtx = sBar + 32, barA = sBar + 40;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int b0 = cluster_id_x() * NB;
  const int T = p.T, B = p.B, Tm = p.Tm;

  if (tid == 0) {
    mbar_init(barM1, THREADS / 64);  // one commit per issuing warp: four warps issue each product
    mbar_init(barM2, THREADS / 64);
    for (int i = 2; i < 8; ++i) mbar_init(sBar + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if constexpr (SAMPLE) {  // output-layer rows of the CTA's attention units
    for (int i = tid; i < UPC * SAMP_VP; i += THREADS) {
      const int u = i / SAMP_VP, v = i % SAMP_VP;
      sWd[i] = v < p.V ? p.Wd[(size_t)(UPC * rank + u) * p.V + v] : 0.0f;
    }
  }
  // tensor memory (all 512 columns): [0, 64) accumulators of the recurrent product, [64, 128) of the attention product;
  // [128, 256) Wa slice (paired layout); [256, 512) attention half of the recurrent matrix (tile m at 256 + 128 m)
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(sTmem) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // h half -> shared memory as fp16: tile mm, row r = gate*32 + u  <->  Wrec[AT + k][gate*H + 64*rank + 32*mm + u], k < H
  for (int seg = warp; seg < H * 8; seg += THREADS / 32) {
    const int k = seg >> 3, g = (seg >> 1) & 3, mm = seg & 1;
    const float w = p.Wrec[(size_t)(AT + k) * 4 * H + g * H + UPC * rank + 32 * mm + lane];
    *reinterpret_cast<__half*>(gen + (sW - base) + mm * (W_BYTES / 2) + sw128h_off(128, g * 32 + lane, k)) = __float2half_rn(w);
  }
  // operand buffers start as zeros (the padding rows stay zero; no product reads them before they are filled)
  for (int i = tid; i < (2 * OP_BYTES + Q_BYTES) / 4; i += THREADS) reinterpret_cast<uint32_t*>(gen + (sOp - base))[i] = 0u;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(sTmem));
  const uint32_t tWa = tmem_base + 128, tW1 = tmem_base + 256;
  {
    // attention half -> tensor memory: tile hh at columns 128 hh .. +127; lane r = gate*32 + u <-> unit 32 hh + u; column c
    // holds K elements 2c, 2c+1 (k < AT).  warp w fills lane quarter (w & 3) = gate of tile w >> 2
    const int q = warp & 3, hh = warp >> 2;
    const float* col = p.Wrec + q * H + UPC * rank + 32 * hh + lane;
#pragma unroll 1
    for (int c0 = 0; c0 < 128; c0 += 32) {
      uint32_t r[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        const int k = 2 * (c0 + c);
        r[c] = pack_h2(col[(size_t)k * 4 * H], col[(size_t)(k + 1) * 4 * H]);
      }
      tmem_st32(tW1 + 128 * hh + c0 + ((uint32_t)(32 * q) << 16), r);
    }
    // Wa slice: lane l = 32 q + lane: part = l >> 6 (0: rows multiplying ho, 1: rows multiplying ctx), attention unit
    // 64*rank + (l & 63); column c holds K elements 2c, 2c+1 of that part; warp w fills columns 64*(w >> 2) .. +63
    // (BAHD: lanes 64..127 hold the CTA's rows of Wq^T instead - both halves multiply ho)
    const int l = 32 * q + lane;
    const float* wcol = (BAHD && l >= 64) ? p.Wq + UPC * rank + (l & 63)
                                          : p.Wa + (size_t)((l >> 6) * H) * AT + UPC * rank + (l & 63);
#pragma unroll 1
    for (int c0 = 0; c0 < 64; c0 += 32) {
      uint32_t r[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        const int k = 2 * (64 * hh + c0 + c);
        r[c] = pack_h2(wcol[(size_t)k * AT], wcol[(size_t)(k + 1) * AT]);
      }
      tmem_st32(tWa + 64 * hh + c0 + ((uint32_t)(32 * q) << 16), r);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  cluster_sync_all();

  // gate-math role: warp <-> (gate g, tile m): the gate rows of units 32*m + lane for all NB utterances
  const int g = warp & 3, m = warp >> 2;
  // product-issue roles (one elected lane, warp-uniform operands: see elect_one).  Warps with (w & 3) < 2 issue the
  // recurrent product: warp (m, j) the K blocks j, j + 2 of each half of tile m into accumulator 2 m + j - first the h
  // half (SS, from shared memory) as soon as hs_t is gathered, later the attention half (TS) and the commit.  Warps with
  // (w & 3) >= 2 issue the attention product: four K steps (of 16) each into accumulators 4 .. 7.
  const int warp_u = (int)warp_uniform((uint32_t)warp);
  const uint32_t tmem_u = warp_uniform(tmem_base);
  const int jq = warp_u & 3, m_u = warp_u >> 2;
  const int ia = 2 * m_u + (jq & 1);  // index of the warp within its issue group (0..3)
  const uint32_t acc1 = tmem_u + ia * NP, acc2 = tmem_u + (4 + ia) * NP;
  const uint32_t tWa_u = tmem_u + 128, tW1_u = tmem_u + 256 + 128 * m_u;
  const uint64_t dWh = make_desc_k128(sW + m_u * (W_BYTES / 2)), dQ = make_desc_k128(sQ);
  const uint64_t dOp[2] = {make_desc_k128(sOp), make_desc_k128(sOp + OP_BYTES)};
  // h half of step t+1 (operand K blocks 4 .. 7; starts the accumulation): eight SS products per issuing warp.  The tensor
  // pipe takes ~68 clocks for each and queues only a few, so issuing them in one go blocks the warp (and with it its
  // attention group) for ~1 100 clocks: they are fed two at a time from inside the sweeps, where the warp waits for its
  // loads anyway (att_fwd_core's hook), and the rest is flushed after the sweeps.
  int h_next = 8;           // next product of the h half (8 = nothing to issue); warp-uniform
  uint32_t h_buf = 0u;
  auto issue_rec_h2 = [&]() {
    if (h_next < 8) {       // (only ever < 8 in the issuing warps)
      if (elect_one()) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int i = h_next + e, kb = jq + 2 * (i >> 2), k4 = i & 3;
          umma_ss(acc1, desc_at(dWh, kb * (128 * 128) + k4 * 32), desc_at(dOp[h_buf], (4 + kb) * (NP * 128) + k4 * 32), IDESC,
                  i ? 1u : 0u);
        }
      }
      __syncwarp();
      h_next += 2;
    }
  };
  auto rec_h_begin = [&](uint32_t nbuf) {
    if (jq < 2) {
      h_next = 0;
      h_buf = nbuf;
    }
  };
  auto rec_h_flush = [&]() {
    while (h_next < 8) issue_rec_h2();
  };
  auto issue_rec_a = [&](uint32_t nbuf) {  // attention half (operand K blocks 0 .. 3) on top, then the commit
    if (jq < 2) {
      if (elect_one()) {
#pragma unroll
        for (int i2 = 0; i2 < 2; ++i2) {
          const int kb = jq + 2 * i2;
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4)
            umma_ts(acc1, tW1_u + (kb * 4 + k4) * 8, desc_at(dOp[nbuf], kb * (NP * 128) + k4 * 32), IDESC, 1u);
        }
        umma_commit(barM1);
      }
      __syncwarp();
    }
  };
  auto issue_att = [&]() {
    if (jq >= 2) {
      if (elect_one()) {
#pragma unroll
        for (int i2 = 0; i2 < 4; ++i2) {
          const int s4 = 4 * ia + i2;  // K step of 16: K block s4 >> 2, 32-byte slice s4 & 3
          umma_ts(acc2, tWa_u + s4 * 8, desc_at(dQ, (s4 >> 2) * (NP * 128) + (s4 & 3) * 32), IDESC, i2 ? 1u : 0u);
        }
        umma_commit(barM2);
      }
      __syncwarp();
    }
  };
  const int unit_g = UPC * rank + 32 * m + lane;
  // combine role (threads 0..127): utterance bq, units (hidden and attention) 4*uq .. 4*uq+3 of the CTA
  const bool comb = tid < 4 * 32;
  const int uq = tid & 15, bq = (tid >> 4) & 7;
  float c_state[4], h_state[4];
  int len_c = 0;
#pragma unroll
  for (int e = 0; e < 4; ++e) c_state[e] = h_state[e] = 0.0f;
  if (comb) {
    const int b = b0 + bq;
    len_c = (b < B) ? p.len[b] : 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int u = UPC * rank + 4 * uq + e;
      c_state[e] = (b < B && p.c0) ? p.c0[(size_t)b * H + u] : 0.0f;
      h_state[e] = (b < B) ? p.S[(size_t)b * (AT + H) + AT + u] : 0.0f;
    }
  }
  const uint32_t seed = p.d.rng ? p.d.rng[0] : 0u, rstep = p.d.rng ? p.d.rng[1] : 0u;
  float f_state[4], f_out[4], f_in[4];
  auto drop_factors = [&](int t) {  // state / output masks of step t, input mask of step t + 1 (acts on a_t)
    if (comb) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const uint32_t col = (uint32_t)(UPC * (int)rank + 4 * uq + e);
        const uint32_t idx = (uint32_t)(b0 + bq) * (uint32_t)H + col;  // H == AT
        f_state[e] = dfac(seed, rstep, p.d.stream + 1u, p.d.thr_state, p.d.inv_state, (uint32_t)t, idx);
        f_out[e] = dfac(seed, rstep, p.d.stream + 2u, p.d.thr_out, p.d.inv_out, (uint32_t)t, idx);
        f_in[e] = dfac(seed, rstep, p.d.stream, p.d.thr_in, p.d.inv_in, (uint32_t)(t + 1), idx);
      }
    }
  };
#pragma unroll
  for (int e = 0; e < 4; ++e) f_state[e] = f_out[e] = f_in[e] = 1.0f;
  drop_factors(0);
  int len_a[NB];
#pragma unroll
  for (int b = 0; b < NB; ++b) len_a[b] = (b0 + b < B) ? p.len[b0 + b] : 0;
  float gx[NB];
  {
    const float* grow0 = p.gates + (size_t)b0 * 4 * H + g * H + unit_g;
#pragma unroll
    for (int b = 0; b < NB; ++b) gx[b] = (0 < len_a[b]) ? grow0[(size_t)b * 4 * H] : 0.0f;
  }
  // attention role: utterance jl of this CTA, warp w4 of its group of four
  const int jl = warp >> 2, w4 = warp & 3, gt = tid & 127;
  const int bl_att = NU * (int)rank + jl;      // row of the utterance in the operand buffers
  const int b_att = b0 + bl_att;
  const int len_q = (b_att < B) ? p.len[b_att] : 0;
  const int L = (b_att < B) ? min(p.mem_len[b_att], Tm) : 0;
  const float gs = p.scaled ? p.g[0] : 1.0f;
  float* sc = sc_all + jl * MAX_TM;
  float* part = part_all + jl * 4 * DM;
  float* red = red_all + jl * 8;
  const uint32_t att_bar_id = 2 + jl;          // named barrier of the 128 threads of this utterance
  const AttRole role = {p.keys, p.values, L, B, b_att, Tm, w4, gt, lane, gs, att_bar_id, sc, part, red};

  uint32_t lphase = 0u;  // SAMPLE: completed phases of the logits barrier
  float v8[8], b8[8];    // BAHD: effective v and bias of the lane's 8 attention units
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    v8[e] = BAHD ? p.v[8 * lane + e] : 0.0f;
    b8[e] = (BAHD && p.batt) ? p.batt[8 * lane + e] : 0.0f;
  }
  // the attention product's accumulators -> partial planes [K group][part]: warp (grp, q) sums accumulators 4 + 2 grp, + 1
  // of lane quarter q; part 0 = lanes 0..63, part 1 = lanes 64..127 (Luong: the ctx rows, B columns 8..15; BAHD: Wq^T)
  auto att_epilogue = [&]() {
    const int grp = warp >> 2, q = warp & 3;
    const uint32_t a0 = tmem_base + ((uint32_t)(32 * q) << 16) + (4 + 2 * grp) * NP + ((!BAHD && q >= 2) ? 8 : 0);
    uint32_t r0[8], r1[8];
    tmem_ld8(a0, r0);
    tmem_ld8(a0 + NP, r1);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    float* ap = apart + ((grp * 2 + (q >> 1)) * NB) * UPC + 32 * (q & 1) + lane;
#pragma unroll
    for (int b = 0; b < NB; ++b) ap[b * UPC] = __uint_as_float(r0[b]) + __uint_as_float(r1[b]);
  };
  for (int t = 0; t < T; ++t) {
#ifdef AP4D_TRACE
    if (tid == 0 && blockIdx.x == 0) g_att_trace_row = (t >= TR_T0 && t < TR_T0 + TR_NT) ? t - TR_T0 : -1;
#endif
    AP4D_STAMP(0);
    float* grow = p.gates + ((size_t)t * B + b0) * 4 * H + g * H + unit_g;
    uint32_t selmask = 0u;  // SAMPLE: utterances of the cluster whose next input is drawn from this step's logits
    if constexpr (SAMPLE) {
      if (t + 1 < T) {
#pragma unroll
        for (int b = 0; b < NB; ++b)
          if (b0 + b < B && avsr_rand_u32(seed, rstep, p.ss_stream, (uint32_t)t, (uint32_t)(b0 + b)) < p.thr_p) selmask |= 1u << b;
      }
    }
    uint32_t r[8];
    if (t > 0) {
      mbar_wait(barM1, (t - 1) & 1);
      AP4D_STAMP(1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t r1[8];
      tmem_ld8(tmem_base + ((uint32_t)(32 * g) << 16) + (2 * m + 0) * NP, r);
      tmem_ld8(tmem_base + ((uint32_t)(32 * g) << 16) + (2 * m + 1) * NP, r1);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
#pragma unroll
      for (int b = 0; b < 8; ++b) r[b] = __float_as_uint(__uint_as_float(r[b]) + __uint_as_float(r1[b]));
    } else {
#pragma unroll
      for (int b = 0; b < 8; ++b) r[b] = 0u;  // att_{-1} = 0; h_0 Wh is already in the x-projection
    }
    float av[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      const float z = __uint_as_float(r[b]) + gx[b];
      float a;
      if (g == 1) a = tanhf_acc(z);
      else a = sigmoidf_acc(g == 2 ? z + 1.0f : z);
      av[b] = a;
      act[(g * NB + b) * UPC + 32 * m + lane] = a;
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    AP4D_STAMP(2);
    const uint32_t nb = (t + 1) & 1;
    const uint32_t hbar_n = sBar + 16 + 8 * nb;
    float hs[4], ho[4], cr[4];
    if (comb) {
      const bool live = t < len_c;
      if (live) {
        const float4 ai = *reinterpret_cast<const float4*>(&act[(0 * NB + bq) * UPC + 4 * uq]);
        const float4 aj = *reinterpret_cast<const float4*>(&act[(1 * NB + bq) * UPC + 4 * uq]);
        const float4 af = *reinterpret_cast<const float4*>(&act[(2 * NB + bq) * UPC + 4 * uq]);
        const float4 ao = *reinterpret_cast<const float4*>(&act[(3 * NB + bq) * UPC + 4 * uq]);
        const float vi[4] = {ai.x, ai.y, ai.z, ai.w}, vj[4] = {aj.x, aj.y, aj.z, aj.w};
        const float vf[4] = {af.x, af.y, af.z, af.w}, vo[4] = {ao.x, ao.y, ao.z, ao.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          cr[e] = vf[e] * c_state[e] + vi[e] * vj[e];
          const float c = fminf(fmaxf(cr[e], -1.0f), 1.0f);
          const float h = vo[e] * tanhf_acc(c);
          c_state[e] = c;
          ho[e] = tf32_rn(h * f_out[e]);         // the cell output: query, attention-layer operand
          h_state[e] = tf32_rn(h * f_state[e]);  // what recurs
        }
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          cr[e] = c_state[e];
          ho[e] = h_state[e];  // the carried state stands in for the output of a finished row (nobody reads it)
        }
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) hs[e] = h_state[e];
      // all-gathers (fp16): hs_t -> h half of the recurrent operand of step t+1; ho_t -> rows 0..7 of the attention operand
      const int ucol = UPC * (int)rank + 4 * uq;
      const uint32_t s01 = pack_h2(hs[0], hs[1]), s23 = pack_h2(hs[2], hs[3]);
      const uint32_t o01 = pack_h2(ho[0], ho[1]), o23 = pack_h2(ho[2], ho[3]);
      const uint32_t dS = sOp + nb * OP_BYTES + sw128h_off(NP, bq, AT + ucol);
      const uint32_t dO = sQ + sw128h_off(NP, bq, ucol);
#pragma unroll
      for (uint32_t dst = 0; dst < (uint32_t)CL; ++dst) {
        const uint32_t bar = mapa(hbar_n, dst);
        st_async_v2(mapa(dS, dst), bar, s01, s23);
        st_async_v2(mapa(dO, dst), bar, o01, o23);
      }
    }
    AP4D_STAMP(3);
    // HBM side of this step + x-projection of the next (overlaps the all-gather)
#pragma unroll
    for (int b = 0; b < NB; ++b)
      if (t < len_a[b]) grow[(size_t)b * 4 * H] = av[b];
    if (comb && b0 + bq < B) {
      const size_t row = (size_t)t * B + b0 + bq;
      const int u0 = UPC * rank + 4 * uq;
      *reinterpret_cast<float4*>(p.craw + row * H + u0) = make_float4(cr[0], cr[1], cr[2], cr[3]);
      *reinterpret_cast<float4*>(p.S + (row + B) * (AT + H) + AT + u0) = make_float4(hs[0], hs[1], hs[2], hs[3]);
      *reinterpret_cast<float4*>(p.hc + row * p.ldhc + u0) = make_float4(ho[0], ho[1], ho[2], ho[3]);
      if constexpr (BAHD)  // the wrapper emits the cell output for the Bahdanau family (zero past the length)
        *reinterpret_cast<float4*>(p.out + row * H + u0) =
            t < len_c ? make_float4(ho[0], ho[1], ho[2], ho[3]) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
    if constexpr (SAMPLE && BAHD) {  // the output layer reads the emitted h
      if (comb && selmask) {
        const bool live = t < len_c;
        *reinterpret_cast<float4*>(&sAf[bq * UPC + 4 * uq]) =
            live ? make_float4(ho[0], ho[1], ho[2], ho[3]) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      }
    }
    if (t + 1 < T) {
      const float* gnext = grow + (size_t)B * 4 * H;
#pragma unroll
      for (int b = 0; b < NB; ++b) gx[b] = (t + 1 < len_a[b]) ? gnext[(size_t)b * 4 * H] : 0.0f;
    }
    // ---------------- attention of utterance b_att with query ho_t ----------------
    const bool live_q = t < len_q;    // masked steps (and padding utterances) skip the memory sweep
    uint4 ra[4], rb[4];
    if (live_q) att_prefetch(role, p.keys, ra, rb);
    if (tid == 0) mbar_expect_tx(hbar_n, 2 * NB * H * 2);
    AP4D_STAMP(4);
    mbar_wait(hbar_n, (t >> 1) & 1);  // every CTA's hs_t / ho_t slices have landed
    AP4D_STAMP(5);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // h half of the recurrent product of step t+1: behind the sweeps (BAHD: behind the query product, which the sweeps wait for)
    if (!BAHD && t + 1 < T) rec_h_begin(nb);

    if constexpr (BAHD) {
      // [ho Wl_h | ho Wq] for the CTA's units as soon as ho_t is there; the pq slices go to the owners of the utterances
      issue_att();
      mbar_wait(barM2, t & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (t + 1 < T) rec_h_begin(nb);
      att_epilogue();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (comb) {
        const float4 p1 = *reinterpret_cast<const float4*>(&apart[(1 * NB + bq) * UPC + 4 * uq]);
        const float4 p3 = *reinterpret_cast<const float4*>(&apart[(3 * NB + bq) * UPC + 4 * uq]);
        const uint32_t dst = (uint32_t)(bq / NU);
        st_async_v4f(mapa(sPqA + (uint32_t)(((bq % NU) * AT + UPC * (int)rank + 4 * uq) * 4), dst), mapa(barPq, dst),
                     p1.x + p3.x, p1.y + p3.y, p1.z + p3.z, p1.w + p3.w);
      }
      if (tid == 0) mbar_expect_tx(barPq, NU * AT * 4);
      mbar_wait(barPq, t & 1);
    }
    float ctxv[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) ctxv[e] = 0.0f;
    if (live_q) {
      float q[8];
      if constexpr (BAHD) {
        const float4 q0 = *reinterpret_cast<const float4*>(&sPq[jl * AT + 8 * lane]);
        const float4 q1 = *reinterpret_cast<const float4*>(&sPq[jl * AT + 8 * lane + 4]);
        if (w4 == 0) {  // kept for the backward pass (without the bias)
          float* dst = p.pq + ((size_t)t * B + b_att) * AT + 8 * lane;
          *reinterpret_cast<float4*>(dst) = q0;
          *reinterpret_cast<float4*>(dst + 4) = q1;
        }
        q[0] = q0.x + b8[0]; q[1] = q0.y + b8[1]; q[2] = q0.z + b8[2]; q[3] = q0.w + b8[3];
        q[4] = q1.x + b8[4]; q[5] = q1.y + b8[5]; q[6] = q1.z + b8[6]; q[7] = q1.w + b8[7];
      } else {
        // query: lane holds dims 8*lane .. 8*lane+7 (one swizzled 16-byte chunk of the operand row)
        const uint4 qraw = *reinterpret_cast<const uint4*>(gen + (sQ - base) + sw128h_off(NP, bl_att, 8 * lane));
        unpack_q(qraw, q);
      }
      att_fwd_core<BAHD, BAHD>(role, q, v8, ra, rb, p.align + ((size_t)t * B + b_att) * Tm, ctxv, issue_rec_h2);
    }
    rec_h_flush();
    AP4D_STAMP(6);
    if (BAHD && w4 == 0) {
      // projected context ctx' = sum_t a_t PV_t: slices to the owners of the attention units (fp32)
      if (b_att < B && !live_q) {
        float* arow = p.align + ((size_t)t * B + b_att) * Tm;
        for (int tm = lane; tm < Tm; tm += 32) arow[tm] = 0.0f;
      }
      const uint32_t dst = (uint32_t)(lane >> 3);
      const uint32_t a0 = mapa(sCxA + (uint32_t)((bl_att * UPC + ((8 * lane) & (UPC - 1))) * 4), dst);
      const uint32_t bar = mapa(barCtx, dst);
      st_async_v4f(a0, bar, ctxv[0], ctxv[1], ctxv[2], ctxv[3]);
      st_async_v4f(a0 + 16, bar, ctxv[4], ctxv[5], ctxv[6], ctxv[7]);
    }
    if (!BAHD && w4 == 0) {
      // ctx_t of this utterance: HBM (tf32-rounded fp32, for the backward pass) + all-gather (fp16, rows 8..15 of sQ)
      if (b_att < B) {
        float* dst = p.hc + ((size_t)t * B + b_att) * p.ldhc + H + 8 * lane;
        *reinterpret_cast<float4*>(dst) = make_float4(ctxv[0], ctxv[1], ctxv[2], ctxv[3]);
        *reinterpret_cast<float4*>(dst + 4) = make_float4(ctxv[4], ctxv[5], ctxv[6], ctxv[7]);
        if (!live_q) {
          float* arow = p.align + ((size_t)t * B + b_att) * Tm;
          for (int tm = lane; tm < Tm; tm += 32) arow[tm] = 0.0f;
        }
      }
      const uint32_t dbuf = sQ + sw128h_off(NP, NB + bl_att, 8 * lane);
      const uint32_t c0 = pack_h2(ctxv[0], ctxv[1]), c1 = pack_h2(ctxv[2], ctxv[3]);
      const uint32_t c2 = pack_h2(ctxv[4], ctxv[5]), c3 = pack_h2(ctxv[6], ctxv[7]);
#pragma unroll
      for (uint32_t dst = 0; dst < (uint32_t)CL; ++dst) st_async_v4(mapa(dbuf, dst), mapa(barCtx, dst), c0, c1, c2, c3);
    }
    // masks of the next step: off the critical chain (this step's a_t needs f_in, kept in `fi`)
    float fi[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) fi[e] = f_in[e];
    drop_factors(t + 1);
    // ---------------- a_t = [ho | ctx] Wa for the CTA's 64 attention units ----------------
    if (tid == 0) mbar_expect_tx(barCtx, BAHD ? NB * UPC * 4 : NB * DM * 2);
    mbar_wait(barCtx, t & 1);
    AP4D_STAMP(7);
    if constexpr (!BAHD) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      issue_att();
      mbar_wait(barM2, t & 1);
      AP4D_STAMP(8);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      att_epilogue();
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }
    AP4D_STAMP(9);
    if (comb) {
      float a[4];
      {
        const float4 p0 = *reinterpret_cast<const float4*>(&apart[(0 * NB + bq) * UPC + 4 * uq]);
        const float4 p2 = *reinterpret_cast<const float4*>(&apart[(2 * NB + bq) * UPC + 4 * uq]);
        // Luong: planes 1, 3 hold the ctx half of the product; BAHD: the projected context arrived through sCx
        const float4 p1 = BAHD ? *reinterpret_cast<const float4*>(&sCx[bq * UPC + 4 * uq])
                               : *reinterpret_cast<const float4*>(&apart[(1 * NB + bq) * UPC + 4 * uq]);
        const float4 p3 = BAHD ? make_float4(0.0f, 0.0f, 0.0f, 0.0f)
                               : *reinterpret_cast<const float4*>(&apart[(3 * NB + bq) * UPC + 4 * uq]);
        a[0] = (p0.x + p1.x) + (p2.x + p3.x); a[1] = (p0.y + p1.y) + (p2.y + p3.y);
        a[2] = (p0.z + p1.z) + (p2.z + p3.z); a[3] = (p0.w + p1.w) + (p2.w + p3.w);
      }
      float ad[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        a[e] = tf32_rn(a[e]);
        ad[e] = tf32_rn(a[e] * fi[e]);  // input dropout of step t+1 on the fed-back attention vector
      }
      const int ucol = UPC * (int)rank + 4 * uq;
      const uint32_t dA = sOp + nb * OP_BYTES + sw128h_off(NP, bq, ucol);
      const uint32_t a01 = pack_h2(ad[0], ad[1]), a23 = pack_h2(ad[2], ad[3]);
#pragma unroll
      for (uint32_t dst = 0; dst < (uint32_t)CL; ++dst) st_async_v2(mapa(dA, dst), mapa(barA, dst), a01, a23);
      if (b0 + bq < B) {
        const size_t row = (size_t)t * B + b0 + bq;
        const bool live = t < len_c;
        if constexpr (!BAHD)
          *reinterpret_cast<float4*>(p.out + row * AT + ucol) =
              live ? make_float4(a[0], a[1], a[2], a[3]) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        *reinterpret_cast<float4*>(p.S + (row + B) * (AT + H) + ucol) = make_float4(ad[0], ad[1], ad[2], ad[3]);
      }
      if constexpr (SAMPLE && !BAHD) {
        if (selmask) *reinterpret_cast<float4*>(&sAf[bq * UPC + 4 * uq]) = make_float4(a[0], a[1], a[2], a[3]);
      }
    }
    if constexpr (SAMPLE) {
      if (selmask) {  // uniform over the cluster
        asm volatile("bar.sync 1, 256;" ::: "memory");
        // partial logits of the selected utterances over this CTA's attention units: warp = utterance, lane = class
        if ((selmask >> warp) & 1u) {
          float sacc = 0.0f;
#pragma unroll 8
          for (int u = 0; u < UPC; ++u) sacc = fmaf(sAf[warp * UPC + u], sWd[u * SAMP_VP + lane], sacc);
          const uint32_t dstoff = sLgAddr + (uint32_t)(((int)rank * NB + warp) * SAMP_VP + lane) * 4;
#pragma unroll
          for (uint32_t dst = 0; dst < (uint32_t)CL; ++dst) st_async_f(mapa(dstoff, dst), mapa(barL, dst), sacc);
        }
        if (tid == 0) mbar_expect_tx(barL, (uint32_t)__popc(selmask) * SAMP_VP * 4 * CL);
        mbar_wait(barL, lphase & 1);
        ++lphase;
        if ((selmask >> warp) & 1u) {
          // inverse CDF of the fp32 softmax, summed in class order (the arithmetic of sched_sample_kernel, misc.cu): lane v
          // forms exp(z_v - max) of its class, lane 0 adds them in class order (same values, same order; sX is free here)
          const float zv = lane < p.V ? p.bd[lane] + ((sLg[(0 * NB + warp) * SAMP_VP + lane] + sLg[(1 * NB + warp) * SAMP_VP + lane]) +
                                                      (sLg[(2 * NB + warp) * SAMP_VP + lane] + sLg[(3 * NB + warp) * SAMP_VP + lane]))
                                      : -INFINITY;
          const float mx = warp_max(zv);
          float* sE = sX + warp * SAMP_VP;
          sE[lane] = lane < p.V ? expf(zv - mx) : 0.0f;
          __syncwarp();
          if (lane == 0) {
            float total = 0.0f;
            for (int v = 0; v < p.V; ++v) total += sE[v];
            const float u01 = (float)(avsr_rand_u32(seed, rstep, p.ss_stream + 1u, (uint32_t)t, (uint32_t)(b0 + warp)) >> 8) * (1.0f / 16777216.0f);
            const float target = u01 * total;
            float cum = 0.0f;
            int pick = p.V - 1;
            for (int v = 0; v < p.V; ++v) {
              cum += sE[v];
              if (cum > target) {
                pick = v;
                break;
              }
            }
            sPick[warp] = pick;
            if (rank == 0) {
              p.sample_ids[(size_t)t * B + b0 + warp] = pick;
              p.used_ids[(size_t)(t + 1) * B + b0 + warp] = pick;
            }
          }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        // x-projection of the drawn embeddings for this thread's gate row (g, unit_g): replaces the prefetched one
        const float* wxcol = p.Wx + g * H + unit_g;
        const float bias_row = p.bias[g * H + unit_g];
#pragma unroll 1
        for (int bb = 0; bb < NB; ++bb) {
          if (!((selmask >> bb) & 1u)) continue;
          const int id = sPick[bb];
          for (int e = tid; e < p.E; e += THREADS) {
            const uint32_t lo = (uint32_t)(((size_t)(t + 1) * B + b0 + bb) * p.E + e);
            const float xv = tf32_rn(p.emb[(size_t)id * p.E + e] * dfac(seed, rstep, p.d.stream + 3u, p.d.thr_in, p.d.inv_in, 0u, lo));
            sX[e] = xv;
            if (rank == 0) p.x[((size_t)(t + 1) * B + b0 + bb) * p.E + e] = xv;
          }
          asm volatile("bar.sync 1, 256;" ::: "memory");
          // (16 weight loads in flight per thread: the column walk is 4 KB-strided L2 traffic; same summation order)
          float acc = bias_row;
          int e0 = 0;
          for (; e0 + 16 <= p.E; e0 += 16) {
            float wv[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) wv[k] = __ldg(wxcol + (size_t)(e0 + k) * 4 * H);
#pragma unroll
            for (int k = 0; k < 16; ++k) acc = fmaf(sX[e0 + k], wv[k], acc);
          }
          for (; e0 < p.E; ++e0) acc = fmaf(sX[e0], wxcol[(size_t)e0 * 4 * H], acc);
#pragma unroll
          for (int b = 0; b < NB; ++b)
            if (b == bb) gx[b] = (t + 1 < len_a[b]) ? acc : 0.0f;
          asm volatile("bar.sync 1, 256;" ::: "memory");
        }
      }
    }
    // recurrent product of step t+1 once every CTA's a_t slice has landed (hs_t landed before the attention).  After
    // the last step the wait only drains the all-gather: no st.async may be in flight towards a CTA that exits.
    if (tid == 0) mbar_expect_tx(barA, NB * AT * 2);
    mbar_wait(barA, t & 1);
    AP4D_STAMP(10);
    if (t + 1 < T) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      issue_rec_a(nb);
    }
    AP4D_STAMP(11);
  }
  if (comb && b0 + bq < B) {
    const size_t o = (size_t)(b0 + bq) * H + UPC * rank + 4 * uq;
    if (p.cT) *reinterpret_cast<float4*>(p.cT + o) = make_float4(c_state[0], c_state[1], c_state[2], c_state[3]);
    if (p.hT) *reinterpret_cast<float4*>(p.hT + o) = make_float4(h_state[0], h_state[1], h_state[2], h_state[3]);
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  cluster_sync_all();
}

// =====================================================================================================
// backward.  Per step t (descending), with dS_t = [dSa_t | dSh_t] = dz_{t+1} [Wl_att ; Wh]^T from the previous iteration:
//   da_t        = dSa_t (.) m_in(t+1) + dout_t                      (gradient wrt the attention vector; kept for dWa)
//   d[ho | ctx] = da_t Wa^T                                         (product 2: K = the CTA's 64 attention units)
//   attention backward of the CTA's 2 utterances: dctx -> d(align) -> ds -> dq (wrt the query ho)
//   dh_t        = (d ho + dq) (.) m_out(t) + dSh_t (.) m_state(t)   -> gate gradients dz_t
//   dS_{t-1}    = dz_t [Wl_att ; Wh]^T                              (product 1: K = the CTA's 256 gate columns)
// Both products are K-split over the cluster; their partial tiles are reduce-scattered through DSMEM to the owners of
// the rows (attention units / hidden units: all 8 utterances; ctx dims: the owners of the utterances).
// =====================================================================================================
struct DBwdParams {
  int T, B, Tm, scaled;
  float grad_scale, inv_grad_scale;
  const int* len;
  const int* mem_len;
  const float* gates;    // [T,B,4H] activations
  const float* craw;     // [T,B,H]
  const float* c0;       // [B,H] or null
  const float* Wrec;     // [(AT+H),4H]
  const float* Wa;       // [(H+DM),AT]
  const __half* keys;    // [Tm,B,H]
  const __half* values;  // [Tm,B,DM]
  const float* g;        // [1] or null
  const float* align;    // [T,B,Tm]
  const float* dout;     // [T,B,AT] gradient wrt the emitted attention vectors, or null
  const float* dcT;      // [B,H] or null
  const float* dhT;      // [B,H] or null
  float* dZ;             // [T,B,4H]
  float* ds;             // [T,B,Tm]
  float* dhc;            // [T,B,H+DM]: the ctx columns receive dctx_t
  float* dA;             // [T,B,AT] da_t (tf32-rounded)
  float* dg;             // [1] or null
  float* dc0;            // [B,H] or null
  float* dh0;            // [B,H] or null
  float* dbias;          // [4H] or null: += column sums of dZ
  DropCfg d;
};

constexpr int BW_TILE_BYTES = 4 * 128 * 128;         // one 128-row tile of Wrec^T restricted to the CTA's 256 gate columns
constexpr int BW_DZ_BYTES = 4 * NP * 128;            // B operand of product 1: 4 K-blocks (gates) x [NP rows x 64 units]
constexpr int BW_DA_BYTES = NP * 128;                // B operand of product 2: [NP rows x 64 attention units]
constexpr int REDH_FLOATS = CL * NB * UPC;           // [src][b][u]
constexpr int REDC_FLOATS = CL * NU * DM;            // [src][utt][dim]; also the dq partial scratch [NU][4][DM]
constexpr int DQ_FLOATS = CL * NU * UPC;             // [src][utt][u]
constexpr size_t DBWD_SMEM = (size_t)2 * BW_TILE_BYTES + BW_DZ_BYTES + BW_DA_BYTES + 3 * REDH_FLOATS * 4 + REDC_FLOATS * 4 +
                             DQ_FLOATS * 4 + NU * DM * 4 + 2 * NU * MAX_TM * 4 + NU * 8 * 4 + 80 + 1024;
static_assert(NU * 4 * DM <= REDC_FLOATS, "dq partial scratch must fit the ctx reduce buffer");

template <bool SMALL>
__global__ void __launch_bounds__(THREADS, 1) attn_lstm_persist4d_bwd_kernel(const DBwdParams p) {
  constexpr int MAXB = SMALL ? SMALL_B : 1;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sW = base;                                  // tiles 0, 1 (attention rows) of Wrec^T
  const uint32_t sDz = sW + 2 * BW_TILE_BYTES;
  const uint32_t sDa = sDz + BW_DZ_BYTES;
  const uint32_t sRedA = sDa + BW_DA_BYTES;                  // partial dSa of the CTA's attention units
  const uint32_t sRedH = sRedA + REDH_FLOATS * 4;            // partial dSh of the CTA's hidden units
  const uint32_t sRedH2 = sRedH + REDH_FLOATS * 4;           // partial d ho of the CTA's hidden units
  const uint32_t sRedC = sRedH2 + REDH_FLOATS * 4;           // partial dctx of the CTA's utterances
  const uint32_t sDq = sRedC + REDC_FLOATS * 4;
  const uint32_t sCtx = sDq + DQ_FLOATS * 4;                 // [NU][DM] dctx of the CTA's utterances
  const uint32_t sSc = sCtx + NU * DM * 4;                   // [NU][MAX_TM] alignments (long memories only)
  const uint32_t sDs = sSc + NU * MAX_TM * 4;                // [NU][MAX_TM] d(align) / ds (long memories only)
  const uint32_t sRed = sDs + NU * MAX_TM * 4;               // [NU][8]
  const uint32_t sBar = sRed + NU * 8 * 4;
  const uint32_t sTmem = sBar + 64;
  const uint32_t barMma = sBar, barMma2 = sBar + 8, barDz = sBar + 16, barRedA = sBar + 24, barRedH = sBar + 32,
                 barRedH2 = sBar + 40, barRedC = sBar + 48, barDq = sBar + 56;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  float* redA = reinterpret_cast<float*>(gen + (sRedA - base));
  float* redH = reinterpret_cast<float*>(gen + (sRedH - base));
  float* redH2 = reinterpret_cast<float*>(gen + (sRedH2 - base));
  float* redC = reinterpret_cast<float*>(gen + (sRedC - base));
  float* dqb = reinterpret_cast<float*>(gen + (sDq - base));
  float* ctx_all = reinterpret_cast<float*>(gen + (sCtx - base));
  float* part_all = redC;
  float* red_all = reinterpret_cast<float*>(gen + (sRed - base));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int b0 = cluster_id_x() * NB;
  const int T = p.T, B = p.B, Tm = p.Tm;

  if (tid == 0) {
    mbar_init(barMma, THREADS / 32);   // one commit per issuing warp
    mbar_init(barMma2, THREADS / 32);
    mbar_init(barDz, THREADS);
    mbar_init(barRedA, 1);
    mbar_init(barRedH, 1);
    mbar_init(barRedH2, 1);
    mbar_init(barRedC, 1);
    mbar_init(barDq, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // tensor memory (all 512 columns): [0, 128) accumulators of both products: 128-row tile mt, K half hj at column
  // 16 (2 mt + hj); [128, 256) the four 128 x 64 tiles of Wa^T (32 columns each); [256, 512) tiles 2, 3 (h rows) of Wrec^T
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(sTmem) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // A[n][k = g*64 + u] = Wrec[n][g*H + 64*rank + u]; rows n < 256 (attention) -> shared memory (tile n >> 7)
  for (int seg = warp; seg < 256 * 8; seg += THREADS / 32) {
    const int n = seg >> 3, g = (seg >> 1) & 3, u = 32 * (seg & 1) + lane;
    const float w = p.Wrec[(size_t)n * 4 * H + g * H + UPC * rank + u];
    *reinterpret_cast<__half*>(gen + (sW - base) + (n >> 7) * BW_TILE_BYTES + sw128h_off(128, n & 127, g * 64 + u)) =
        __float2half_rn(w);
  }
  for (int i = tid; i < (BW_DZ_BYTES + BW_DA_BYTES) / 4; i += THREADS) reinterpret_cast<uint32_t*>(gen + (sDz - base))[i] = 0u;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(sTmem));
  const uint32_t tWaT = tmem_base + 128, tA = tmem_base + 256;
  {
    // rows n = 256 + 128*tt + 32*q + lane (h rows) -> tensor memory tile tt; column c holds k = 2c, 2c+1 (adjacent units)
    const int q = warp & 3, tt = warp >> 2;
    const float* row = p.Wrec + (size_t)(256 + 128 * tt + 32 * q + lane) * 4 * H + UPC * rank;
#pragma unroll 1
    for (int c0 = 0; c0 < 128; c0 += 32) {
      uint32_t r[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        const int k = 2 * (c0 + c), g = k >> 6, u = k & 63;
        const float2 w = *reinterpret_cast<const float2*>(row + g * H + u);
        r[c] = pack_h2(w.x, w.y);
      }
      tmem_st32(tA + 128 * tt + c0 + ((uint32_t)(32 * q) << 16), r);
    }
    // Wa^T: row n = 128*wt + 32*q + lane of [ho dims | ctx dims], K = the CTA's 64 attention units
#pragma unroll 1
    for (int i = 0; i < 2; ++i) {
      const int wt = 2 * tt + i;
      const float* wrow = p.Wa + (size_t)(128 * wt + 32 * q + lane) * AT + UPC * rank;
      uint32_t r[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        const float2 w = *reinterpret_cast<const float2*>(wrow + 2 * c);
        r[c] = pack_h2(w.x, w.y);
      }
      tmem_st32(tWaT + 32 * wt + ((uint32_t)(32 * q) << 16), r);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  cluster_sync_all();

  const uint64_t dWt = make_desc_k128(sW), dDz = make_desc_k128(sDz), dDa = make_desc_k128(sDa);
  // gate-gradient role: thread = (local unit ul - hidden and attention -, utterances 2*(warp >> 1) + j)
  constexpr int PB = 2;
  const int ul = 32 * (warp & 1) + lane;
  const int unit = UPC * rank + ul;
  float dc[PB], dh_carry[PB];
  int len_t[PB];
#pragma unroll
  for (int j = 0; j < PB; ++j) {
    const int b = b0 + (warp >> 1) * PB + j;
    len_t[j] = (b < B) ? p.len[b] : 0;
    dc[j] = (b < B && p.dcT) ? p.dcT[(size_t)b * H + unit] : 0.0f;
    dh_carry[j] = (b < B && p.dhT) ? p.dhT[(size_t)b * H + unit] : 0.0f;
  }
  const uint32_t seed = p.d.rng ? p.d.rng[0] : 0u, rstep = p.d.rng ? p.d.rng[1] : 0u;
  float gi[PB], gj[PB], gf[PB], go[PB], crw[PB], cpv[PB], dov[PB];
#pragma unroll
  for (int j = 0; j < PB; ++j) gi[j] = gj[j] = gf[j] = go[j] = crw[j] = cpv[j] = dov[j] = 0.0f;
  auto load_step = [&](int t) {
#pragma unroll
    for (int j = 0; j < PB; ++j) {
      const int b = b0 + (warp >> 1) * PB + j;
      if (t >= 0 && t < len_t[j]) {
        const float* g = p.gates + ((size_t)t * B + b) * 4 * H + unit;
        gi[j] = g[0]; gj[j] = g[H]; gf[j] = g[2 * H]; go[j] = g[3 * H];
        const size_t o = ((size_t)t * B + b) * H + unit;
        crw[j] = p.craw[o];
        cpv[j] = t > 0 ? p.craw[o - (size_t)B * H] : (p.c0 ? p.c0[(size_t)b * H + unit] : 0.0f);
        dov[j] = p.dout ? p.dout[((size_t)t * B + b) * AT + unit] : 0.0f;  // wrt the emitted attention unit `unit`
      }
    }
  };
  // attention role
  const int jl = warp >> 2, w4 = warp & 3, gt = tid & 127;
  const int bl_att = NU * (int)rank + jl;
  const int b_att = b0 + bl_att;
  const int len_q = (b_att < B) ? p.len[b_att] : 0;
  const int L = (b_att < B) ? min(p.mem_len[b_att], Tm) : 0;
  const float gs = p.scaled ? p.g[0] : 1.0f;
  float* dctx_s = ctx_all + jl * DM;
  float* a_s = reinterpret_cast<float*>(gen + (sSc - base)) + jl * MAX_TM;
  float* ds_s = reinterpret_cast<float*>(gen + (sDs - base)) + jl * MAX_TM;
  float* part = part_all + jl * 4 * DM;
  float* red = red_all + jl * 8;
  const uint32_t att_bar_id = 2 + jl;
  const AttRole role = {p.keys, p.values, L, B, b_att, Tm, w4, gt, lane, gs, att_bar_id, nullptr, part, red};
  const float vzero8[8] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
  // product-issue / reduce-scatter roles (issue: elected lane, warp-uniform operands - see elect_one)
  const int q = warp & 3;
  const int warp_u = (int)warp_uniform((uint32_t)warp);
  const uint32_t tmem_u = warp_uniform(tmem_base);
  const int mt_i = warp_u >> 1, hj_i = warp_u & 1;
  const uint32_t acc_i = tmem_u + (2 * mt_i + hj_i) * NP;
  const uint32_t tWaT_u = tmem_u + 128, tA_u = tmem_u + 256;

  float bsum[4] = {0.0f, 0.0f, 0.0f, 0.0f};  // bias gradient of this thread's unit
  load_step(T - 1);
  for (int it = 0; it < T; ++it) {
    const int t = T - 1 - it;
    const bool live_q = t < len_q;
    // first values and this step's alignments: requested before the partial sums of the previous iteration arrive
    uint4 ra[4], rb[4];
    float al[MAXB];
    const int jrow = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
#pragma unroll
    for (int i = 0; i < MAXB; ++i) al[i] = 0.0f;
    if (live_q) {
      att_prefetch(role, p.values, ra, rb);
      if constexpr (SMALL) {
#pragma unroll
        for (int i = 0; i < MAXB; ++i) {
          const int tm = w4 + 32 * i + 4 * jrow;
          if (tm < L) al[i] = p.align[((size_t)t * B + b_att) * Tm + tm];
        }
      } else {
        for (int tm = gt; tm < Tm; tm += 128) a_s[tm] = p.align[((size_t)t * B + b_att) * Tm + tm];
      }
    }
    // masks of this step for the thread's unit (attention unit: input mask of step t+1; hidden unit: state / output)
    float f_in[PB], f_st[PB], f_o[PB];
#pragma unroll
    for (int j = 0; j < PB; ++j) {
      const uint32_t idx = (uint32_t)(b0 + (warp >> 1) * PB + j) * (uint32_t)H + (uint32_t)unit;  // H == AT
      f_in[j] = dfac(seed, rstep, p.d.stream, p.d.thr_in, p.d.inv_in, (uint32_t)(t + 1), idx);
      f_st[j] = dfac(seed, rstep, p.d.stream + 1u, p.d.thr_state, p.d.inv_state, (uint32_t)t, idx);
      f_o[j] = dfac(seed, rstep, p.d.stream + 2u, p.d.thr_out, p.d.inv_out, (uint32_t)t, idx);
    }
    // ---- (A) dS_t pushed during the previous iteration -> da_t, dSh_t ------------------------------------------
    float dh_in[PB], sa[PB];
#pragma unroll
    for (int j = 0; j < PB; ++j) {
      dh_in[j] = dh_carry[j];
      sa[j] = 0.0f;
    }
    if (it > 0) {
      if (tid == 0) {
        mbar_expect_tx(barRedA, REDH_FLOATS * 4);
        mbar_expect_tx(barRedH, REDH_FLOATS * 4);
      }
      mbar_wait(barRedA, (it - 1) & 1);
#pragma unroll
      for (int j = 0; j < PB; ++j) {
        const int bl = (warp >> 1) * PB + j;
#pragma unroll
        for (int src = 0; src < CL; ++src) sa[j] += redA[(src * NB + bl) * UPC + ul];
      }
    }
#pragma unroll
    for (int j = 0; j < PB; ++j) {
      const int bl = (warp >> 1) * PB + j;
      const float da = (t < len_t[j]) ? tf32_rn(sa[j] * f_in[j] + dov[j]) : 0.0f;
      if (b0 + bl < B) p.dA[((size_t)t * B + b0 + bl) * AT + unit] = da;
      *reinterpret_cast<__half*>(gen + (sDa - base) + sw128h_off(NP, bl, ul)) = __float2half_rn(da * p.grad_scale);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    asm volatile("bar.sync 1, 256;" ::: "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // ---- (B) product 2: d[ho | ctx] partial from the CTA's 64 attention units ------------------------------------
    if (elect_one()) {
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        const int s4 = 2 * hj_i + kk;  // K step of 16 attention units
        umma_ts(acc_i, tWaT_u + 32 * mt_i + s4 * 8, desc_at(dDa, s4 * 32), IDESC, kk ? 1u : 0u);
      }
      umma_commit(barMma2);
    }
    __syncwarp();
    if (it > 0) {
      mbar_wait(barRedH, (it - 1) & 1);
#pragma unroll
      for (int j = 0; j < PB; ++j) {
        const int bl = (warp >> 1) * PB + j;
#pragma unroll
        for (int src = 0; src < CL; ++src) dh_in[j] += redH[(src * NB + bl) * UPC + ul];
      }
    }
    // ---- (C) partial d[ho | ctx] -> owners ------------------------------------------------------------------------
    mbar_wait(barMma2, it & 1);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
    for (int mi = 0; mi < 2; ++mi) {
      const int mt = 2 * (warp >> 2) + mi;
      uint32_t r[8], r1[8];
      tmem_ld8(tmem_base + ((uint32_t)(32 * q) << 16) + (2 * mt) * NP, r);
      tmem_ld8(tmem_base + ((uint32_t)(32 * q) << 16) + (2 * mt + 1) * NP, r1);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int c = 0; c < NB; ++c) r[c] = __float_as_uint(__uint_as_float(r[c]) + __uint_as_float(r1[c]));
      if (mt < 2) {  // ho dims 128*mt + 32*q + lane -> owner CTA 2*mt + (q >> 1), local unit 32*(q & 1) + lane
        const uint32_t dst = (uint32_t)(2 * mt + (q >> 1));
        const uint32_t a0 = mapa(sRedH2 + (uint32_t)((rank * NB) * UPC + 32 * (q & 1) + lane) * 4, dst);
        const uint32_t bar = mapa(barRedH2, dst);
#pragma unroll
        for (int c = 0; c < NB; ++c) st_async_f(a0 + c * UPC * 4, bar, __uint_as_float(r[c]) * p.inv_grad_scale);
      } else {       // ctx dims 128*(mt-2) + 32*q + lane; column c = utterance -> owner CTA c / NU
        const int dim = 128 * (mt - 2) + 32 * q + lane;
#pragma unroll
        for (uint32_t dst = 0; dst < (uint32_t)CL; ++dst) {
          const uint32_t a0 = mapa(sRedC + (uint32_t)((rank * NU) * DM + dim) * 4, dst);
          const uint32_t bar = mapa(barRedC, dst);
#pragma unroll
          for (int u2 = 0; u2 < NU; ++u2) st_async_f(a0 + u2 * DM * 4, bar, __uint_as_float(r[NU * dst + u2]) * p.inv_grad_scale);
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    // ---- (D) dctx of the CTA's utterances -------------------------------------------------------------------------
    if (tid == 0) mbar_expect_tx(barRedC, REDC_FLOATS * 4);
    mbar_wait(barRedC, it & 1);
    if (live_q) {
#pragma unroll
      for (int i = 0; i < DM / 128; ++i) {
        const int d = gt + 128 * i;
        float v = 0.0f;
#pragma unroll
        for (int src = 0; src < CL; ++src) v += redC[(src * NU + jl) * DM + d];
        dctx_s[d] = v;
        p.dhc[((size_t)t * B + b_att) * (H + DM) + H + d] = v;
      }
    }
    // redC is free from here on (a peer pushes its next partials only after it has received this CTA's dS partials of
    // this iteration): it doubles as the dq partial scratch
    asm volatile("bar.sync 1, 256;" ::: "memory");
    // ---- (E) attention backward of the CTA's utterances, dq all-to-all --------------------------------------------
    float dqv[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) dqv[e] = 0.0f;
    float ds_keep[MAXB];
#pragma unroll
    for (int i = 0; i < MAXB; ++i) ds_keep[i] = 0.0f;
    if (live_q)
      att_bwd_core<SMALL>(role, dctx_s, ra, rb, al, ds_keep, a_s, ds_s, p.ds + ((size_t)t * B + b_att) * Tm, p.scaled != 0, p.dg, dqv,
                          vzero8, vzero8);
    if (w4 == 0) {
      // dq dims 8*lane .. +7 belong to the CTA owning units (8*lane)/64
      const uint32_t dst = (uint32_t)(lane >> 3);
      const uint32_t a0 = mapa(sDq + (uint32_t)(((rank * NU + jl) * UPC + ((8 * lane) & (UPC - 1))) * 4), dst);
      const uint32_t bar = mapa(barDq, dst);
      st_async_v4f(a0, bar, dqv[0], dqv[1], dqv[2], dqv[3]);
      st_async_v4f(a0 + 16, bar, dqv[4], dqv[5], dqv[6], dqv[7]);
    }
    // ---- (F) d ho + dq of this CTA's units -> gate gradients ------------------------------------------------------
    if (tid == 0) {
      mbar_expect_tx(barDq, DQ_FLOATS * 4);
      mbar_expect_tx(barRedH2, REDH_FLOATS * 4);
    }
    mbar_wait(barRedH2, it & 1);
    mbar_wait(barDq, it & 1);
    float dz[4][PB];
#pragma unroll
    for (int j = 0; j < PB; ++j) {
      const int bl = (warp >> 1) * PB + j;
      if (t < len_t[j]) {
        float dho = dqb[bl * UPC + ul];
#pragma unroll
        for (int src = 0; src < CL; ++src) dho += redH2[(src * NB + bl) * UPC + ul];
        const float dh = dh_in[j] * f_st[j] + dho * f_o[j];  // state-dropped h recurs, output-dropped h is the query / operand
        const float c = fminf(fmaxf(crw[j], -1.0f), 1.0f);
        const float tc = tanhf_acc(c);
        const float cp = t > 0 ? fminf(fmaxf(cpv[j], -1.0f), 1.0f) : cpv[j];
        const float dct = dc[j] + dh * go[j] * (1.0f - tc * tc);
        const float dcr = (crw[j] >= -1.0f && crw[j] <= 1.0f) ? dct : 0.0f;
        dz[0][j] = dcr * gj[j] * gi[j] * (1.0f - gi[j]);
        dz[1][j] = dcr * gi[j] * (1.0f - gj[j] * gj[j]);
        dz[2][j] = dcr * cp * gf[j] * (1.0f - gf[j]);
        dz[3][j] = dh * tc * go[j] * (1.0f - go[j]);
        dc[j] = dcr * gf[j];
        dh_carry[j] = 0.0f;
      } else {
        dz[0][j] = dz[1][j] = dz[2][j] = dz[3][j] = 0.0f;
        dh_carry[j] = dh_in[j];
      }
#pragma unroll
      for (int g = 0; g < 4; ++g)
        *reinterpret_cast<__half*>(gen + (sDz - base) + sw128h_off(NP, bl, g * 64 + ul)) = __float2half_rn(dz[g][j] * p.grad_scale);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    mbar_arrive(barDz);
    {
      // product 1: partial [dSa | dSh](512) x NB from this CTA's 256 gate columns; warp (mt, hj) issues the K half hj
      // (gates 2 hj, 2 hj + 1) of the 128-row tile mt into its own accumulator
      mbar_wait(barDz, it & 1);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < 2; ++kk)
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            const int kb = 2 * hj_i + kk;
            const uint64_t db = desc_at(dDz, kb * (NP * 128) + k4 * 32);
            if (mt_i < 2)
              umma_ss(acc_i, desc_at(dWt, mt_i * BW_TILE_BYTES + kb * (128 * 128) + k4 * 32), db, IDESC, (kk | k4) ? 1u : 0u);
            else
              umma_ts(acc_i, tA_u + (mt_i - 2) * 128 + (kb * 4 + k4) * 8, db, IDESC, (kk | k4) ? 1u : 0u);
          }
        umma_commit(barMma);
      }
      __syncwarp();
    }
#pragma unroll
    for (int j = 0; j < PB; ++j) {
      const int b = b0 + (warp >> 1) * PB + j;
      if (b < B) {
        float* o = p.dZ + ((size_t)t * B + b) * 4 * H + unit;
        o[0] = tf32_rn(dz[0][j]); o[H] = tf32_rn(dz[1][j]); o[2 * H] = tf32_rn(dz[2][j]); o[3 * H] = tf32_rn(dz[3][j]);
#pragma unroll
        for (int g = 0; g < 4; ++g) bsum[g] += tf32_rn(dz[g][j]);
      }
    }
    load_step(t - 1);
    if constexpr (SMALL) {
      if (live_q) att_bwd_small_tail(role, p.ds + ((size_t)t * B + b_att) * Tm, al, ds_keep, p.scaled != 0, p.dg);
    }
    // ---- (G) partial dS_{t-1} -> owners of the attention / hidden units -------------------------------------------
    mbar_wait(barMma, it & 1);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (it + 1 < T) {
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) {
        const int mt = 2 * (warp >> 2) + mi;  // tiles 0, 1: attention units; 2, 3: hidden units
        uint32_t r[8], r1[8];
        tmem_ld8(tmem_base + ((uint32_t)(32 * q) << 16) + (2 * mt) * NP, r);
        tmem_ld8(tmem_base + ((uint32_t)(32 * q) << 16) + (2 * mt + 1) * NP, r1);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int c = 0; c < NB; ++c) r[c] = __float_as_uint(__uint_as_float(r[c]) + __uint_as_float(r1[c]));
        const uint32_t dst = (uint32_t)(2 * (mt & 1) + (q >> 1));
        const uint32_t buf = mt < 2 ? sRedA : sRedH;
        const uint32_t a0 = mapa(buf + (uint32_t)((rank * NB) * UPC + 32 * (q & 1) + lane) * 4, dst);
        const uint32_t bar = mapa(mt < 2 ? barRedA : barRedH, dst);
#pragma unroll
        for (int c = 0; c < NB; ++c) st_async_f(a0 + c * UPC * 4, bar, __uint_as_float(r[c]) * p.inv_grad_scale);
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
#pragma unroll
  for (int j = 0; j < PB; ++j) {
    const int b = b0 + (warp >> 1) * PB + j;
    if (b < B) {
      // dh_0 = dz_0 Wh^T is added by the host; here only what fully masked utterances carry through
      if (p.dh0) p.dh0[(size_t)b * H + unit] = dh_carry[j];
      if (p.dc0) p.dc0[(size_t)b * H + unit] = dc[j];
    }
  }
  if (p.dbias) {
#pragma unroll
    for (int g = 0; g < 4; ++g) atomicAdd(p.dbias + g * H + unit, bsum[g]);
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  cluster_sync_all();
}


// =====================================================================================================
// backward, Bahdanau family (see the BAHD notes at the forward kernel).  Per step t (descending):
//   da_t   = dSa_t (.) m_in(t+1)                                 (kept for dWl; sent to the owners of the utterances)
//   d(align) = PV . da_t -> softmax backward -> ds -> dpq_u = sum_tm ds[tm] v_u (1 - tanh^2(keys[tm]_u + pq_u + b_u))
//   d ho   = da_t Wl_h^T + dpq_t Wq^T                            (ONE product: K = [64 attention units | 64 query units] of
//                                                                  the CTA, the dpq half issued when the sweeps are done)
//   dh_t   = (d ho + dout_t) (.) m_out(t) + dSh_t (.) m_state(t) -> gate gradients -> dS_{t-1} = dz_t [Wl_att ; Wh]^T
// =====================================================================================================
struct DBahdBwdParams {
  int T, B, Tm;
  float grad_scale, inv_grad_scale;
  const int* len;
  const int* mem_len;
  const float* gates;    // [T,B,4H] activations
  const float* craw;     // [T,B,H]
  const float* c0;       // [B,H] or null
  const float* Wrec;     // [(AT+H),4H]
  const float* Wa;       // attention_layer kernel [(H+Dm),AT] (only its first H rows enter the recurrence)
  const float* Wq;       // [H,AT]
  const float* v;        // [AT] effective v
  const float* batt;     // [AT] or null
  const __half* keys;    // [Tm,B,AT]
  const __half* pv;      // [Tm,B,AT] projected values
  const float* pq;       // [T,B,AT] processed queries of the forward pass
  const float* align;    // [T,B,Tm]
  const float* dout;     // [T,B,H] gradient wrt the emitted cell output, or null
  const float* dcT;      // [B,H] or null
  const float* dhT;      // [B,H] or null
  float* dZ;             // [T,B,4H]
  float* ds;             // [T,B,Tm]
  float* dpq;            // [T,B,AT]
  float* dA;             // [T,B,AT]
  float* dc0;            // [B,H] or null
  float* dh0;            // [B,H] or null
  float* dbias;          // [4H] or null
  DropCfg d;
};
constexpr size_t DBAHD_BWD_SMEM = (size_t)2 * BW_TILE_BYTES + BW_DZ_BYTES + 2 * BW_DA_BYTES + 3 * REDH_FLOATS * 4 +
                                  NU * 4 * DM * 4 + DQ_FLOATS * 4 + NU * AT * 4 + 2 * NU * MAX_TM * 4 + NU * 8 * 4 + 96 + 1024;

template <bool SMALL>
__global__ void __launch_bounds__(THREADS, 1) attn_lstm_persist4d_bahd_bwd_kernel(const DBahdBwdParams p) {
  constexpr int MAXB = SMALL ? SMALL_B : 1;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sW = base;                                  // tiles 0, 1 (attention rows) of Wrec^T
  const uint32_t sDz = sW + 2 * BW_TILE_BYTES;
  const uint32_t sDa = sDz + BW_DZ_BYTES;                    // B operand, da half of product 2
  const uint32_t sDp = sDa + BW_DA_BYTES;                    // B operand, dpq half of product 2
  const uint32_t sRedA = sDp + BW_DA_BYTES;                  // partial dSa of the CTA's attention units
  const uint32_t sRedH = sRedA + REDH_FLOATS * 4;            // partial dSh of the CTA's hidden units
  const uint32_t sRedH2 = sRedH + REDH_FLOATS * 4;           // partial d ho of the CTA's hidden units
  const uint32_t sPart = sRedH2 + REDH_FLOATS * 4;           // [NU][4][DM] per-warp partial dpq
  const uint32_t sDq = sPart + NU * 4 * DM * 4;              // dpq of the CTA's query units [src][utt][u]
  const uint32_t sDaU = sDq + DQ_FLOATS * 4;                 // da of the CTA's utterances [NU][AT]
  const uint32_t sSc = sDaU + NU * AT * 4;                   // [NU][MAX_TM] alignments (long memories only)
  const uint32_t sDs = sSc + NU * MAX_TM * 4;                // [NU][MAX_TM] d(align) / ds (long memories only)
  const uint32_t sRed = sDs + NU * MAX_TM * 4;               // [NU][8]
  const uint32_t sBar = sRed + NU * 8 * 4;
  const uint32_t sTmem = sBar + 72;
  const uint32_t barMma = sBar, barMma2 = sBar + 8, barDz = sBar + 16, barRedA = sBar + 24, barRedH = sBar + 32,
                 barRedH2 = sBar + 40, barDaU = sBar + 48, barDq = sBar + 56;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  float* redA = reinterpret_cast<float*>(gen + (sRedA - base));
  float* redH = reinterpret_cast<float*>(gen + (sRedH - base));
  float* redH2 = reinterpret_cast<float*>(gen + (sRedH2 - base));
  float* dqb = reinterpret_cast<float*>(gen + (sDq - base));
  float* dau_all = reinterpret_cast<float*>(gen + (sDaU - base));
  float* part_all = reinterpret_cast<float*>(gen + (sPart - base));
  float* red_all = reinterpret_cast<float*>(gen + (sRed - base));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int b0 = cluster_id_x() * NB;
  const int T = p.T, B = p.B, Tm = p.Tm;

  if (tid == 0) {
    mbar_init(barMma, THREADS / 32);   // one commit per issuing warp
    mbar_init(barMma2, THREADS / 32);
    mbar_init(barDz, THREADS);
    mbar_init(barRedA, 1);
    mbar_init(barRedH, 1);
    mbar_init(barRedH2, 1);
    mbar_init(barDaU, 1);
    mbar_init(barDq, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // tensor memory (all 512 columns): [0, 128) accumulators of both products; [128, 256) the two 128-row tiles (ho dims) of
  // [Wl_h^T | Wq^T] restricted to the CTA's 64 attention / query units (K = 128: 64 columns each); [256, 512) tiles 2, 3
  // (h rows) of Wrec^T
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(sTmem) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // A[n][k = g*64 + u] = Wrec[n][g*H + 64*rank + u]; rows n < 256 (attention) -> shared memory (tile n >> 7)
  for (int seg = warp; seg < 256 * 8; seg += THREADS / 32) {
    const int n = seg >> 3, g = (seg >> 1) & 3, u = 32 * (seg & 1) + lane;
    const float w = p.Wrec[(size_t)n * 4 * H + g * H + UPC * rank + u];
    *reinterpret_cast<__half*>(gen + (sW - base) + (n >> 7) * BW_TILE_BYTES + sw128h_off(128, n & 127, g * 64 + u)) =
        __float2half_rn(w);
  }
  for (int i = tid; i < (BW_DZ_BYTES + 2 * BW_DA_BYTES) / 4; i += THREADS) reinterpret_cast<uint32_t*>(gen + (sDz - base))[i] = 0u;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(sTmem));
  const uint32_t tB2 = tmem_base + 128, tA = tmem_base + 256;
  {
    // rows n = 256 + 128*tt + 32*q + lane (h rows) -> tensor memory tile tt; column c holds k = 2c, 2c+1 (adjacent units)
    const int q = warp & 3, tt = warp >> 2;
    const float* row = p.Wrec + (size_t)(256 + 128 * tt + 32 * q + lane) * 4 * H + UPC * rank;
#pragma unroll 1
    for (int c0 = 0; c0 < 128; c0 += 32) {
      uint32_t r[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        const int k = 2 * (c0 + c), g = k >> 6, u = k & 63;
        const float2 w = *reinterpret_cast<const float2*>(row + g * H + u);
        r[c] = pack_h2(w.x, w.y);
      }
      tmem_st32(tA + 128 * tt + c0 + ((uint32_t)(32 * q) << 16), r);
    }
    // [Wl_h^T | Wq^T]: row n = 128*tt + 32*q + lane (ho dim), K = the CTA's 64 attention units, then its 64 query units
#pragma unroll 1
    for (int i = 0; i < 2; ++i) {
      const float* wrow = (i == 0 ? p.Wa : p.Wq) + (size_t)(128 * tt + 32 * q + lane) * AT + UPC * rank;
      uint32_t r[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        const float2 w = *reinterpret_cast<const float2*>(wrow + 2 * c);
        r[c] = pack_h2(w.x, w.y);
      }
      tmem_st32(tB2 + 64 * tt + 32 * i + ((uint32_t)(32 * q) << 16), r);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  cluster_sync_all();

  const uint64_t dWt = make_desc_k128(sW), dDz = make_desc_k128(sDz), dDa = make_desc_k128(sDa), dDp = make_desc_k128(sDp);
  // gate-gradient role: thread = (local unit ul - hidden, attention and query -, utterances 2*(warp >> 1) + j)
  constexpr int PB = 2;
  const int ul = 32 * (warp & 1) + lane;
  const int unit = UPC * rank + ul;
  float dc[PB], dh_carry[PB];
  int len_t[PB];
#pragma unroll
  for (int j = 0; j < PB; ++j) {
    const int b = b0 + (warp >> 1) * PB + j;
    len_t[j] = (b < B) ? p.len[b] : 0;
    dc[j] = (b < B && p.dcT) ? p.dcT[(size_t)b * H + unit] : 0.0f;
    dh_carry[j] = (b < B && p.dhT) ? p.dhT[(size_t)b * H + unit] : 0.0f;
  }
  const uint32_t seed = p.d.rng ? p.d.rng[0] : 0u, rstep = p.d.rng ? p.d.rng[1] : 0u;
  float gi[PB], gj[PB], gf[PB], go[PB], crw[PB], cpv[PB], dov[PB];
#pragma unroll
  for (int j = 0; j < PB; ++j) gi[j] = gj[j] = gf[j] = go[j] = crw[j] = cpv[j] = dov[j] = 0.0f;
  auto load_step = [&](int t) {
#pragma unroll
    for (int j = 0; j < PB; ++j) {
      const int b = b0 + (warp >> 1) * PB + j;
      if (t >= 0 && t < len_t[j]) {
        const float* g = p.gates + ((size_t)t * B + b) * 4 * H + unit;
        gi[j] = g[0]; gj[j] = g[H]; gf[j] = g[2 * H]; go[j] = g[3 * H];
        const size_t o = ((size_t)t * B + b) * H + unit;
        crw[j] = p.craw[o];
        cpv[j] = t > 0 ? p.craw[o - (size_t)B * H] : (p.c0 ? p.c0[(size_t)b * H + unit] : 0.0f);
        dov[j] = p.dout ? p.dout[o] : 0.0f;  // wrt the emitted cell output
      }
    }
  };
  // attention role
  const int jl = warp >> 2, w4 = warp & 3, gt = tid & 127;
  const int bl_att = NU * (int)rank + jl;
  const int b_att = b0 + bl_att;
  const int len_q = (b_att < B) ? p.len[b_att] : 0;
  const int L = (b_att < B) ? min(p.mem_len[b_att], Tm) : 0;
  float* dau_s = dau_all + jl * AT;
  float* a_s = reinterpret_cast<float*>(gen + (sSc - base)) + jl * MAX_TM;
  float* ds_s = reinterpret_cast<float*>(gen + (sDs - base)) + jl * MAX_TM;
  float* part = part_all + jl * 4 * DM;
  float* red = red_all + jl * 8;
  const uint32_t att_bar_id = 2 + jl;
  const AttRole role = {p.keys, p.pv, L, B, b_att, Tm, w4, gt, lane, 1.0f, att_bar_id, nullptr, part, red};
  float v8[8], b8[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    v8[e] = p.v[8 * lane + e];
    b8[e] = p.batt ? p.batt[8 * lane + e] : 0.0f;
  }
  // product roles: product 2: warp (mt_b, ks_b) issues K step ks_b of both halves of tile mt_b and reduce-scatters the tile's
  // lane quarter q; product 1: warp (mt_i, hj_i) as in the Luong kernel
  // (issue: elected lane, warp-uniform operands - see elect_one)
  const int q = warp & 3;
  const int warp_u = (int)warp_uniform((uint32_t)warp);
  const uint32_t tmem_u = warp_uniform(tmem_base);
  const int mt_b = warp_u >> 2, ks_b = warp_u & 3;
  const uint32_t acc_b = tmem_u + warp_u * NP;
  const int mt_i = warp_u >> 1, hj_i = warp_u & 1;
  const uint32_t acc_i = tmem_u + (2 * mt_i + hj_i) * NP;
  const uint32_t tB2_u = tmem_u + 128, tA_u = tmem_u + 256;

  float bsum[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  load_step(T - 1);
  for (int it = 0; it < T; ++it) {
    const int t = T - 1 - it;
    const bool live_q = t < len_q;
    uint4 ra[4], rb[4];
    float al[MAXB], q8[8];
    const int jrow = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
#pragma unroll
    for (int i = 0; i < MAXB; ++i) al[i] = 0.0f;
#pragma unroll
    for (int e = 0; e < 8; ++e) q8[e] = 0.0f;
    if (live_q) {
      att_prefetch(role, p.pv, ra, rb);
      const float* pqr = p.pq + ((size_t)t * B + b_att) * AT + 8 * lane;
      const float4 q0 = *reinterpret_cast<const float4*>(pqr), q1 = *reinterpret_cast<const float4*>(pqr + 4);
      q8[0] = q0.x + b8[0]; q8[1] = q0.y + b8[1]; q8[2] = q0.z + b8[2]; q8[3] = q0.w + b8[3];
      q8[4] = q1.x + b8[4]; q8[5] = q1.y + b8[5]; q8[6] = q1.z + b8[6]; q8[7] = q1.w + b8[7];
      if constexpr (SMALL) {
#pragma unroll
        for (int i = 0; i < MAXB; ++i) {
          const int tm = w4 + 32 * i + 4 * jrow;
          if (tm < L) al[i] = p.align[((size_t)t * B + b_att) * Tm + tm];
        }
      } else {
        for (int tm = gt; tm < Tm; tm += 128) a_s[tm] = p.align[((size_t)t * B + b_att) * Tm + tm];
      }
    }
    float f_in[PB], f_st[PB], f_o[PB];
#pragma unroll
    for (int j = 0; j < PB; ++j) {
      const uint32_t idx = (uint32_t)(b0 + (warp >> 1) * PB + j) * (uint32_t)H + (uint32_t)unit;  // H == AT
      f_in[j] = dfac(seed, rstep, p.d.stream, p.d.thr_in, p.d.inv_in, (uint32_t)(t + 1), idx);
      f_st[j] = dfac(seed, rstep, p.d.stream + 1u, p.d.thr_state, p.d.inv_state, (uint32_t)t, idx);
      f_o[j] = dfac(seed, rstep, p.d.stream + 2u, p.d.thr_out, p.d.inv_out, (uint32_t)t, idx);
    }
    // ---- (A) dS_t pushed during the previous iteration -> da_t (operand of product 2; to the utterances' owners) -----
    float dh_in[PB], sa[PB];
#pragma unroll
    for (int j = 0; j < PB; ++j) {
      dh_in[j] = dh_carry[j];
      sa[j] = 0.0f;
    }
    if (it > 0) {
      if (tid == 0) {
        mbar_expect_tx(barRedA, REDH_FLOATS * 4);
        mbar_expect_tx(barRedH, REDH_FLOATS * 4);
      }
      mbar_wait(barRedA, (it - 1) & 1);
#pragma unroll
      for (int j = 0; j < PB; ++j) {
        const int bl = (warp >> 1) * PB + j;
#pragma unroll
        for (int src = 0; src < CL; ++src) sa[j] += redA[(src * NB + bl) * UPC + ul];
      }
    }
#pragma unroll
    for (int j = 0; j < PB; ++j) {
      const int bl = (warp >> 1) * PB + j;
      const float da = (t < len_t[j]) ? tf32_rn(sa[j] * f_in[j]) : 0.0f;
      if (b0 + bl < B) p.dA[((size_t)t * B + b0 + bl) * AT + unit] = da;
      *reinterpret_cast<__half*>(gen + (sDa - base) + sw128h_off(NP, bl, ul)) = __float2half_rn(da * p.grad_scale);
      const uint32_t dst = (uint32_t)(bl / NU);
      st_async_f(mapa(sDaU + (uint32_t)(((bl % NU) * AT + unit) * 4), dst), mapa(barDaU, dst), da);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    asm volatile("bar.sync 1, 256;" ::: "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // ---- (B) product 2, da half: d ho partial from the CTA's 64 attention units (the dpq half follows the sweeps) -----
    if (elect_one()) umma_ts(acc_b, tB2_u + 64 * mt_b + ks_b * 8, desc_at(dDa, ks_b * 32), IDESC, 0u);
    __syncwarp();
    if (it > 0) {
      mbar_wait(barRedH, (it - 1) & 1);
#pragma unroll
      for (int j = 0; j < PB; ++j) {
        const int bl = (warp >> 1) * PB + j;
#pragma unroll
        for (int src = 0; src < CL; ++src) dh_in[j] += redH[(src * NB + bl) * UPC + ul];
      }
    }
    // ---- (E) attention backward of the CTA's utterances: da of all 256 attention units has been gathered ------------
    if (tid == 0) mbar_expect_tx(barDaU, NU * AT * 4);
    mbar_wait(barDaU, it & 1);
    float dqv[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) dqv[e] = 0.0f;
    float ds_keep[MAXB];
#pragma unroll
    for (int i = 0; i < MAXB; ++i) ds_keep[i] = 0.0f;
    if (live_q)
      att_bwd_core<SMALL, true>(role, dau_s, ra, rb, al, ds_keep, a_s, ds_s, p.ds + ((size_t)t * B + b_att) * Tm, false, nullptr,
                                dqv, q8, v8);
    if (w4 == 0) {
      if (live_q) {  // dpq of this step: kept for dWq and the post-loop dkeys / dv pass
        float* dst = p.dpq + ((size_t)t * B + b_att) * AT + 8 * lane;
        *reinterpret_cast<float4*>(dst) = make_float4(dqv[0], dqv[1], dqv[2], dqv[3]);
        *reinterpret_cast<float4*>(dst + 4) = make_float4(dqv[4], dqv[5], dqv[6], dqv[7]);
      }
      // dpq dims 8*lane .. +7 belong to the CTA owning query units (8*lane)/64
      const uint32_t dst = (uint32_t)(lane >> 3);
      const uint32_t a0 = mapa(sDq + (uint32_t)(((rank * NU + jl) * UPC + ((8 * lane) & (UPC - 1))) * 4), dst);
      const uint32_t bar = mapa(barDq, dst);
      st_async_v4f(a0, bar, dqv[0], dqv[1], dqv[2], dqv[3]);
      st_async_v4f(a0 + 16, bar, dqv[4], dqv[5], dqv[6], dqv[7]);
    }
    // ---- (F') dpq of this CTA's query units -> second half of product 2 -> d ho partials to the owners ---------------
    if (tid == 0) mbar_expect_tx(barDq, DQ_FLOATS * 4);
    mbar_wait(barDq, it & 1);
#pragma unroll
    for (int j = 0; j < PB; ++j) {
      const int bl = (warp >> 1) * PB + j;
      *reinterpret_cast<__half*>(gen + (sDp - base) + sw128h_off(NP, bl, ul)) = __float2half_rn(dqb[bl * UPC + ul] * p.grad_scale);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    asm volatile("bar.sync 1, 256;" ::: "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (elect_one()) {
      umma_ts(acc_b, tB2_u + 64 * mt_b + 32 + ks_b * 8, desc_at(dDp, ks_b * 32), IDESC, 1u);
      umma_commit(barMma2);
    }
    __syncwarp();
    mbar_wait(barMma2, it & 1);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    {
      // warp (mt_b, q): rows 128*mt_b + 32*q + lane of d ho, summed over the four K-step accumulators of the tile
      uint32_t r0[8], r1[8], r2[8], r3[8];
      const uint32_t a0 = tmem_base + ((uint32_t)(32 * q) << 16) + (4 * mt_b) * NP;
      tmem_ld8(a0, r0);
      tmem_ld8(a0 + NP, r1);
      tmem_ld8(a0 + 2 * NP, r2);
      tmem_ld8(a0 + 3 * NP, r3);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      const uint32_t dst = (uint32_t)(2 * mt_b + (q >> 1));
      const uint32_t d0 = mapa(sRedH2 + (uint32_t)((rank * NB) * UPC + 32 * (q & 1) + lane) * 4, dst);
      const uint32_t bar = mapa(barRedH2, dst);
#pragma unroll
      for (int c = 0; c < NB; ++c)
        st_async_f(d0 + c * UPC * 4, bar,
                   ((__uint_as_float(r0[c]) + __uint_as_float(r1[c])) + (__uint_as_float(r2[c]) + __uint_as_float(r3[c]))) * p.inv_grad_scale);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    // ---- (F) d ho of this CTA's units -> gate gradients --------------------------------------------------------------
    if (tid == 0) mbar_expect_tx(barRedH2, REDH_FLOATS * 4);
    mbar_wait(barRedH2, it & 1);
    float dz[4][PB];
#pragma unroll
    for (int j = 0; j < PB; ++j) {
      const int bl = (warp >> 1) * PB + j;
      if (t < len_t[j]) {
        float dho = dov[j];
#pragma unroll
        for (int src = 0; src < CL; ++src) dho += redH2[(src * NB + bl) * UPC + ul];
        const float dh = dh_in[j] * f_st[j] + dho * f_o[j];
        const float c = fminf(fmaxf(crw[j], -1.0f), 1.0f);
        const float tc = tanhf_acc(c);
        const float cp = t > 0 ? fminf(fmaxf(cpv[j], -1.0f), 1.0f) : cpv[j];
        const float dct = dc[j] + dh * go[j] * (1.0f - tc * tc);
        const float dcr = (crw[j] >= -1.0f && crw[j] <= 1.0f) ? dct : 0.0f;
        dz[0][j] = dcr * gj[j] * gi[j] * (1.0f - gi[j]);
        dz[1][j] = dcr * gi[j] * (1.0f - gj[j] * gj[j]);
        dz[2][j] = dcr * cp * gf[j] * (1.0f - gf[j]);
        dz[3][j] = dh * tc * go[j] * (1.0f - go[j]);
        dc[j] = dcr * gf[j];
        dh_carry[j] = 0.0f;
      } else {
        dz[0][j] = dz[1][j] = dz[2][j] = dz[3][j] = 0.0f;
        dh_carry[j] = dh_in[j];
      }
#pragma unroll
      for (int g = 0; g < 4; ++g)
        *reinterpret_cast<__half*>(gen + (sDz - base) + sw128h_off(NP, bl, g * 64 + ul)) = __float2half_rn(dz[g][j] * p.grad_scale);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    mbar_arrive(barDz);
    {
      mbar_wait(barDz, it & 1);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < 2; ++kk)
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            const int kb = 2 * hj_i + kk;
            const uint64_t db = desc_at(dDz, kb * (NP * 128) + k4 * 32);
            if (mt_i < 2)
              umma_ss(acc_i, desc_at(dWt, mt_i * BW_TILE_BYTES + kb * (128 * 128) + k4 * 32), db, IDESC, (kk | k4) ? 1u : 0u);
            else
              umma_ts(acc_i, tA_u + (mt_i - 2) * 128 + (kb * 4 + k4) * 8, db, IDESC, (kk | k4) ? 1u : 0u);
          }
        umma_commit(barMma);
      }
      __syncwarp();
    }
#pragma unroll
    for (int j = 0; j < PB; ++j) {
      const int b = b0 + (warp >> 1) * PB + j;
      if (b < B) {
        float* o = p.dZ + ((size_t)t * B + b) * 4 * H + unit;
        o[0] = tf32_rn(dz[0][j]); o[H] = tf32_rn(dz[1][j]); o[2 * H] = tf32_rn(dz[2][j]); o[3 * H] = tf32_rn(dz[3][j]);
#pragma unroll
        for (int g = 0; g < 4; ++g) bsum[g] += tf32_rn(dz[g][j]);
      }
    }
    load_step(t - 1);
    if constexpr (SMALL) {
      if (live_q) att_bwd_small_tail(role, p.ds + ((size_t)t * B + b_att) * Tm, al, ds_keep, false, nullptr);
    }
    // ---- (G) partial dS_{t-1} -> owners of the attention / hidden units -------------------------------------------
    mbar_wait(barMma, it & 1);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (it + 1 < T) {
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) {
        const int mt = 2 * (warp >> 2) + mi;  // tiles 0, 1: attention units; 2, 3: hidden units
        uint32_t r[8], r1[8];
        tmem_ld8(tmem_base + ((uint32_t)(32 * q) << 16) + (2 * mt) * NP, r);
        tmem_ld8(tmem_base + ((uint32_t)(32 * q) << 16) + (2 * mt + 1) * NP, r1);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int c = 0; c < NB; ++c) r[c] = __float_as_uint(__uint_as_float(r[c]) + __uint_as_float(r1[c]));
        const uint32_t dst = (uint32_t)(2 * (mt & 1) + (q >> 1));
        const uint32_t buf = mt < 2 ? sRedA : sRedH;
        const uint32_t a0 = mapa(buf + (uint32_t)((rank * NB) * UPC + 32 * (q & 1) + lane) * 4, dst);
        const uint32_t bar = mapa(mt < 2 ? barRedA : barRedH, dst);
#pragma unroll
        for (int c = 0; c < NB; ++c) st_async_f(a0 + c * UPC * 4, bar, __uint_as_float(r[c]) * p.inv_grad_scale);
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
#pragma unroll
  for (int j = 0; j < PB; ++j) {
    const int b = b0 + (warp >> 1) * PB + j;
    if (b < B) {
      if (p.dh0) p.dh0[(size_t)b * H + unit] = dh_carry[j];
      if (p.dc0) p.dc0[(size_t)b * H + unit] = dc[j];
    }
  }
  if (p.dbias) {
#pragma unroll
    for (int g = 0; g < 4; ++g) atomicAdd(p.dbias + g * H + unit, bsum[g]);
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  cluster_sync_all();
}

}  // namespace ap4

// ---------------------------------------------------------------------------------------------------------------------
// host side: called by attn_persist_fwd / attn_persist_bwd (attn_persist.cu) when the layer's DropoutWrapper is on
// ---------------------------------------------------------------------------------------------------------------------
static ap4::DropCfg drop_cfg(const AvsrRnnSeq* r) {
  ap4::DropCfg d;
  d.rng = r->rng;
  d.stream = r->drop_stream;
  d.thr_in = r->rng ? r->thr_in : 0u; d.thr_state = r->rng ? r->thr_state : 0u; d.thr_out = r->rng ? r->thr_out : 0u;
  d.inv_in = inv_keep_of(d.thr_in); d.inv_state = inv_keep_of(d.thr_state); d.inv_out = inv_keep_of(d.thr_out);
  return d;
}

template <bool BAHD>
static int launch_fwd_variant(cudaStream_t st, const AvsrRnnSeq* r, const ap4::DParams& p) {
  if (r->samp)
    return ap4::launch_cluster(st, ap4::attn_lstm_persist4d_fwd_kernel<true, BAHD>, r->B, ap4::DFWD_SMEM, p, AVSR_K_ATTN_FWD);
  return ap4::launch_cluster(st, ap4::attn_lstm_persist4d_fwd_kernel<false, BAHD>, r->B, ap4::DFWD_SMEM, p, AVSR_K_ATTN_FWD);
}

// values_h: fp16 memory values (Luong family) or the fp16 PROJECTED values PV = values Wl_c (Bahdanau family)
int attn_persist4d_launch_fwd(cudaStream_t st, const AvsrRnnSeq* r, const void* keys_h, const void* values_h) {
  const AvsrAttnMech& m = r->mech[0];
  const bool bahd = m.kind >= AVSR_ATTN_BAHDANAU;
  ap4::DParams p;
  p.T = r->T; p.B = r->B; p.Tm = m.Tm; p.scaled = m.kind == AVSR_ATTN_SCALED_LUONG;
  p.len = r->len; p.mem_len = m.mem_len; p.gates = r->gates; p.Wrec = r->Wrec; p.Wa = m.Wl;
  p.keys = reinterpret_cast<const __half*>(keys_h); p.values = reinterpret_cast<const __half*>(values_h);
  p.g = m.g; p.c0 = r->c0; p.S = r->S; p.craw = r->craw; p.out = r->out; p.hc = m.hc; p.ldhc = r->H + m.Dm;
  p.align = m.align;
  p.cT = r->cT; p.hT = r->hT;
  p.d = drop_cfg(r);
  p.Wq = m.Wq; p.v = m.v; p.batt = m.bias; p.pq = m.pq;
  if (bahd) AVSR_REQUIRE(m.Wq && m.v && m.pq, "rnn: Bahdanau mechanism without query layer / v / pq buffer");
  p.Wd = p.bd = p.emb = p.Wx = p.bias = nullptr;
  p.used_ids = p.sample_ids = nullptr; p.x = nullptr; p.V = p.E = 0; p.ss_stream = p.thr_p = 0u;
  if (r->samp) {
    const AvsrSampling& s = *r->samp;
    AVSR_REQUIRE(r->rng != nullptr, "rnn: scheduled sampling needs the generator words (rng)");
    AVSR_REQUIRE(s.V > 0 && s.V <= ap4::SAMP_VP && s.E > 0 && s.E <= ap4::SAMP_EMAX, "rnn: sampling alphabet / embedding too wide");
    p.Wd = s.Wd; p.bd = s.bd; p.emb = s.embedding; p.Wx = s.Wx; p.bias = s.bias;
    p.used_ids = s.used_ids; p.sample_ids = s.sample_ids; p.x = s.x; p.V = s.V; p.E = s.E;
    p.ss_stream = s.stream; p.thr_p = s.thr_p;
  }
  return bahd ? launch_fwd_variant<true>(st, r, p) : launch_fwd_variant<false>(st, r, p);
}

int attn_persist4d_launch_bwd(cudaStream_t st, const AvsrRnnSeq* r, const void* keys_h, const void* values_h) {
  const AvsrAttnMech& m = r->mech[0];
  ap4::DBwdParams p;
  p.T = r->T; p.B = r->B; p.Tm = m.Tm; p.scaled = m.kind == AVSR_ATTN_SCALED_LUONG;
  p.grad_scale = r->grad_scale > 0.0f ? r->grad_scale : 1.0f;
  p.inv_grad_scale = 1.0f / p.grad_scale;
  p.len = r->len; p.mem_len = m.mem_len; p.gates = r->gates; p.craw = r->craw; p.c0 = r->c0;
  p.Wrec = r->Wrec; p.Wa = m.Wl;
  p.keys = reinterpret_cast<const __half*>(keys_h); p.values = reinterpret_cast<const __half*>(values_h);
  p.g = m.g; p.align = m.align; p.dout = r->dout; p.dcT = r->dcT; p.dhT = r->dhT;
  p.dZ = r->dZ; p.ds = m.ds; p.dhc = m.dhc; p.dA = r->dA; p.dg = m.dg; p.dc0 = r->dc0; p.dh0 = r->dh0; p.dbias = r->dbias;
  p.d = drop_cfg(r);
  return m.Tm <= ap4::SMALL_TM
             ? ap4::launch_cluster(st, ap4::attn_lstm_persist4d_bwd_kernel<true>, r->B, ap4::DBWD_SMEM, p, AVSR_K_ATTN_BWD)
             : ap4::launch_cluster(st, ap4::attn_lstm_persist4d_bwd_kernel<false>, r->B, ap4::DBWD_SMEM, p, AVSR_K_ATTN_BWD);
}

// keys_h / pv_h: fp16 keys and projected values (see attn_persist4d_launch_fwd)
int attn_persist4d_launch_bahd_bwd(cudaStream_t st, const AvsrRnnSeq* r, const void* keys_h, const void* pv_h) {
  const AvsrAttnMech& m = r->mech[0];
  ap4::DBahdBwdParams p;
  p.T = r->T; p.B = r->B; p.Tm = m.Tm;
  p.grad_scale = r->grad_scale > 0.0f ? r->grad_scale : 1.0f;
  p.inv_grad_scale = 1.0f / p.grad_scale;
  p.len = r->len; p.mem_len = m.mem_len; p.gates = r->gates; p.craw = r->craw; p.c0 = r->c0;
  p.Wrec = r->Wrec; p.Wa = m.Wl; p.Wq = m.Wq; p.v = m.v; p.batt = m.bias;
  p.keys = reinterpret_cast<const __half*>(keys_h); p.pv = reinterpret_cast<const __half*>(pv_h);
  p.pq = m.pq; p.align = m.align; p.dout = r->dout; p.dcT = r->dcT; p.dhT = r->dhT;
  p.dZ = r->dZ; p.ds = m.ds; p.dpq = m.dpq; p.dA = r->dA; p.dc0 = r->dc0; p.dh0 = r->dh0; p.dbias = r->dbias;
  p.d = drop_cfg(r);
  AVSR_REQUIRE(m.Wq && m.v && m.pq && m.dpq && m.ds && r->dA, "rnn bwd: Bahdanau mechanism buffers missing");
  return m.Tm <= ap4::SMALL_TM
             ? ap4::launch_cluster(st, ap4::attn_lstm_persist4d_bahd_bwd_kernel<true>, r->B, ap4::DBAHD_BWD_SMEM, p, AVSR_K_ATTN_BWD)
             : ap4::launch_cluster(st, ap4::attn_lstm_persist4d_bahd_bwd_kernel<false>, r->B, ap4::DBAHD_BWD_SMEM, p, AVSR_K_ATTN_BWD);
}

}  // namespace avsr

#ifdef AP4D_TRACE
// debug accessor of the trace library (tools/ap4d_trace.py): clock64 stamps [TR_NT][TR_NP] of the last forward launch
extern "C" int avsr_debug_ap4d_trace(unsigned long long* out, int n) {
  const int total = avsr::ap4::TR_NT * avsr::ap4::TR_NP;
  if (n < total) return -1;
  if (cudaDeviceSynchronize() != cudaSuccess) return -2;
  if (cudaMemcpyFromSymbol(out, avsr::ap4::g_ap4d_trace, sizeof(unsigned long long) * total) != cudaSuccess) return -3;
  return total;
}
extern "C" int avsr_debug_att_trace(unsigned long long* out, int n) {
  if (n < 16 * 8) return -1;
  if (cudaDeviceSynchronize() != cudaSuccess) return -2;
  if (cudaMemcpyFromSymbol(out, avsr::ap4::g_att_trace, sizeof(unsigned long long) * 16 * 8) != cudaSuccess) return -3;
  return 16 * 8;
}
#endif
