// Persistent DUAL-ATTENTION decoder layer (WLAS, decoder_bimodal.py:179-277): AttentionWrapper(DropoutWrapper(LSTMCell))
// with TWO Luong-family mechanisms (video memory, audio memory; attention.py:55-72, 108-128 `linear_fusion`: one
// attention layer per mechanism, attention = concat(a_video, a_audio), 512-d) in one launch per direction, on clusters of
// EIGHT CTAs.
//
// Why eight: the cell input is [x | a_0 | a_1] and the recurrent operand [a_0 (.) m_in | a_1 (.) m_in | hs] has K = 768, so the
// recurrent matrix [Wl_att ; Wh] is 768 x 1024 fp16 = 1.5 MB - more than four CTAs can keep on chip (227 KB shared memory
// + 256 KB tensor memory each, next to operands and accumulators).  With eight CTAs every CTA owns 32 hidden units =
// ONE 128-row gate tile x K = 768 = 384 tensor-memory columns; the remaining 128 columns are the accumulators.
//
// Step t (forward), per cluster of 16 utterances:
//   z_t      = gx_t + [a_{t-1} (.) m_in(t) | hs_{t-1}] [Wl_att ; Wh]     product 1: A in tensor memory (TS), K split over the
//                                                                        8 warps (6 MMAs each into 8 accumulators)
//   gates -> c_t, h_t; ho_t = h_t (.) m_out(t) (query, operand), hs_t = h_t (.) m_state(t); all-gather {hs_t, ho_t} (DSMEM)
//   ha_k     = ho_t Wl_h,k for the CTA's 32 attention units of both mechanisms  product 2: A (64 live rows) in shared memory
//   per mechanism k, for the CTA's 2 utterances: score = g_k keys_k . ho_t -> masked softmax -> ctx'_k = sum_tm a_tm PV_k,tm
//            with PV_k = values_k Wl_c,k projected ONCE per batch by the host (any memory depth): the context half of the
//            attention layer leaves the recurrence; slices of ctx'_k go to the owners of the attention units
//   a_k,t    = tf32(ha_k + ctx'_k); emitted (output_attention), a_k,t (.) m_in(t+1) all-gathered into the next operand
// ScheduledEmbeddingTrainingHelper (decoder_bimodal.py:229-234) runs inside the kernel as in attn_persist4d.cu.
// The backward kernel mirrors it: da_t = dSa_t (.) m_in + dout_t; d ho = sum_k da_k Wl_h,k^T (product 2) + dq_k from the
// attention backward of both mechanisms (d(align) = PV . da, softmax backward, dq = g ds^T keys); gate gradients;
// dS_{t-1} = dz_t [Wl_att ; Wh]^T (product 1, A = six 128-row tiles in tensor memory), reduce-scattered over the cluster.
#include "ap4_common.cuh"

namespace avsr {
namespace ap8 {
using namespace ap4;  // device helpers only (st.async, mbarrier, tcgen05 wrappers, the attention sweeps); geometry below

constexpr int CL8 = 8;
constexpr int NM = 2;                       // mechanisms
constexpr int UP8 = H / CL8;                // 32 hidden units (and attention units per mechanism) per CTA
constexpr int NB8 = 16;                     // utterances per cluster = N of the products
constexpr int NU8 = NB8 / CL8;              // utterances whose attention a CTA owns
constexpr int ATT = NM * H;                 // 512 attention units
constexpr int K8 = ATT + H;                 // 768
constexpr int KB8 = K8 / 64;                // 12 K-blocks of 64 halves
constexpr int OP8_BYTES = KB8 * NB8 * 128;  // one recurrent operand buffer
constexpr int Q8_BYTES = 4 * NB8 * 128;     // ho operand of product 2
constexpr int WA8_BYTES = 4 * 128 * 128;    // A of product 2: rows [mech 0 units | mech 1 units | 64 unused] x K = 256
static_assert(NB8 == NP, "N of the products");
static_assert(NU8 == NU, "the attention helpers assume two utterances per CTA");

struct Drop8 {
  const uint32_t* rng;  // {seed, step}
  uint32_t stream, thr_in, thr_state, thr_out;
  float inv_in, inv_state, inv_out;
};
__device__ __forceinline__ float dfac8(uint32_t seed, uint32_t step, uint32_t stream, uint32_t thr, float inv, uint32_t hi,
                                       uint32_t lo) {
  return (thr == 0u || avsr_rand_u32(seed, step, stream, hi, lo) < thr) ? inv : 0.0f;
}

// =====================================================================================================
// forward
// =====================================================================================================
struct WFwdParams {
  int T, B;
  int Tm[NM], scaled[NM];
  const int* len;
  const int* mem_len[NM];
  float* gates;             // [T,B,4H] in: x-projection (+ h0 Wh at t = 0); out: activations
  const float* Wrec;        // [(ATT+H), 4H] rows [a_0 ; a_1 ; h]
  const float* Wa[NM];      // attention_layer kernels [(H+Dm_k), H]: only the first H rows enter the recurrence
  const __half* keys[NM];   // [Tm_k,B,H]
  const __half* pv[NM];     // [Tm_k,B,H] projected values
  const float* g[NM];       // attention_g [1] or null
  const float* c0;          // [B,H] or null
  float* S;                 // [(T+1),B,ATT+H]; S[0] by the caller; rows 1.. = [a_0 (.) m | a_1 (.) m | hs], tf32-rounded
  float* craw;              // [T,B,H]
  float* out;               // [T,B,ATT] attention vectors (tf32-rounded), zero past the length
  float* hc[NM];            // [T,B,ldhc_k]: the ho columns
  int ldhc[NM];
  float* align[NM];         // [T,B,Tm_k]
  float* cT;
  float* hT;
  Drop8 d;
  // scheduled sampling (AvsrSampling) or Wd == null
  const float* Wd;          // [ATT, V]
  const float* bd;
  const float* emb;
  const float* Wx;
  const float* bias;
  int* used_ids;
  int* sample_ids;
  float* x;
  int V, E;
  uint32_t ss_stream, thr_p;
};

constexpr int S8_VP = 32;     // padded alphabet
constexpr int S8_EMAX = 256;
constexpr int FW8_ACT = 4 * NB8 * UP8;          // floats
constexpr int FW8_AP = NM * NB8 * UP8;          // ha / ctx' planes [mech][b][u]
constexpr int FW8_SAMP = NB8 * NM * UP8 + NM * UP8 * S8_VP + CL8 * NB8 * S8_VP + S8_EMAX + NB8;  // floats
constexpr size_t WFWD_SMEM = (size_t)WA8_BYTES + 2 * OP8_BYTES + Q8_BYTES + FW8_ACT * 4 + 2 * FW8_AP * 4 + NU8 * MAX_TM * 4 +
                             NU8 * 4 * DM * 4 + NU8 * 8 * 4 + 128 + FW8_SAMP * 4 + 1024;

template <bool SAMPLE>
__global__ void __launch_bounds__(THREADS, 1) wlas_persist8_fwd_kernel(const WFwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sWa = base;                          // A of product 2
  const uint32_t sOp = sWa + WA8_BYTES;               // two recurrent operand buffers [a_0 | a_1 | hs] x 16 rows
  const uint32_t sQ = sOp + 2 * OP8_BYTES;            // ho x 16 rows
  const uint32_t sAct = sQ + Q8_BYTES;                // [4][NB8][UP8]
  const uint32_t sAp = sAct + FW8_ACT * 4;            // ha [NM][NB8][UP8]
  const uint32_t sCxA = sAp + FW8_AP * 4;             // ctx' of the CTA's attention units [NM][NB8][UP8] (st.async target)
  const uint32_t sSc = sCxA + FW8_AP * 4;             // [NU8][MAX_TM]
  const uint32_t sPart = sSc + NU8 * MAX_TM * 4;      // [NU8][4][DM]
  const uint32_t sRed = sPart + NU8 * 4 * DM * 4;     // [NU8][8]
  const uint32_t sBar = sRed + NU8 * 8 * 4;           // [0] mma1 [1] mma2 [2,3] h_full[buf] [4] ctx_full [5] a_full [6] logits
  const uint32_t sTmem = sBar + 64;
  const uint32_t sSamp = sBar + 128;                  // a_t [NB8][64] | Wd rows [64][32] | logits [CL8][NB8][32] | x | picks
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  float* act = reinterpret_cast<float*>(gen + (sAct - base));
  float* apart = reinterpret_cast<float*>(gen + (sAp - base));
  float* sCx = reinterpret_cast<float*>(gen + (sCxA - base));
  float* sc_all = reinterpret_cast<float*>(gen + (sSc - base));
  float* part_all = reinterpret_cast<float*>(gen + (sPart - base));
  float* red_all = reinterpret_cast<float*>(gen + (sRed - base));
  float* sAf = reinterpret_cast<float*>(gen + (sSamp - base));
  float* sWd = sAf + NB8 * NM * UP8;
  float* sLg = sWd + NM * UP8 * S8_VP;
  float* sX = sLg + CL8 * NB8 * S8_VP;
  int* sPick = reinterpret_cast<int*>(sX + S8_EMAX);
  const uint32_t sLgAddr = sSamp + (NB8 * NM * UP8 + NM * UP8 * S8_VP) * 4;
  const uint32_t barM1 = sBar, barM2 = sBar + 8, barCtx = sBar + 32, barA = sBar + 40, barL = sBar + 48;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int b0 = cluster_id_x() * NB8;
  const int T = p.T, B = p.B;

  if (tid == 0) {
    mbar_init(barM1, THREADS / 32);
    mbar_init(barM2, THREADS / 32);
    for (int i = 2; i < 7; ++i) mbar_init(sBar + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if constexpr (SAMPLE) {  // output-layer rows of the CTA's attention units (row j = mech * 32 + u)
    for (int i = tid; i < NM * UP8 * S8_VP; i += THREADS) {
      const int j = i / S8_VP, v = i % S8_VP;
      sWd[i] = v < p.V ? p.Wd[(size_t)((j >> 5) * H + UP8 * rank + (j & 31)) * p.V + v] : 0.0f;
    }
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(sTmem) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // A of product 2 -> shared memory (fp16): row r = mech * 32 + u <-> Wa_mech[k][32*rank + u], k < H; rows 64..127 zero
  for (int i = tid; i < WA8_BYTES / 4; i += THREADS) reinterpret_cast<uint32_t*>(gen + (sWa - base))[i] = 0u;
  for (int i = tid; i < (2 * OP8_BYTES + Q8_BYTES) / 4; i += THREADS) reinterpret_cast<uint32_t*>(gen + (sOp - base))[i] = 0u;
  __syncthreads();
  for (int seg = warp; seg < H * NM; seg += THREADS / 32) {
    const int k = seg >> 1, mech = seg & 1;
    const float w = p.Wa[mech][(size_t)k * H + UP8 * rank + lane];
    *reinterpret_cast<__half*>(gen + (sWa - base) + sw128h_off(128, mech * 32 + lane, k)) = __float2half_rn(w);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(sTmem));
  const uint32_t tW = tmem_base + 128;
  {
    // recurrent tile -> tensor memory: lane r = gate*32 + u <-> Wrec[k][gate*H + 32*rank + u]; column c holds K elements
    // 2c, 2c+1.  warp w fills lane quarter (w & 3) = gate, columns 192*(w >> 2) .. +191
    const int q = warp & 3, hh = warp >> 2;
    const float* col = p.Wrec + q * H + UP8 * rank + lane;
#pragma unroll 1
    for (int c0 = 0; c0 < 192; c0 += 32) {
      uint32_t r[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        const int k = 2 * (192 * hh + c0 + c);
        r[c] = pack_h2(col[(size_t)k * 4 * H], col[(size_t)(k + 1) * 4 * H]);
      }
      tmem_st32(tW + 192 * hh + c0 + ((uint32_t)(32 * q) << 16), r);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  cluster_sync_all();

  // gate-math role: gate g = lane quarter, unit = lane, utterances 8*hf .. 8*hf+7
  const int g = warp & 3, hf = warp >> 2;
  const int unit_g = UP8 * rank + lane;
  // (issue: elected lane, warp-uniform operands - ap4_common.cuh elect_one)
  const int warp_u = (int)ap4::warp_uniform((uint32_t)warp);
  const uint32_t tmem_u = ap4::warp_uniform(tmem_base);
  const uint32_t acc_w = tmem_u + warp_u * NP;  // every warp issues into its own accumulator
  const uint32_t tW_u = tmem_u + 128;
  const uint64_t dWa = make_desc_k128(sWa), dQ = make_desc_k128(sQ);
  const uint64_t dOp[2] = {make_desc_k128(sOp), make_desc_k128(sOp + OP8_BYTES)};
  auto issue_rec = [&](uint32_t nbuf) {  // K steps 6 w .. 6 w + 5 of 48
    if (ap4::elect_one()) {
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const int ks = 6 * warp_u + i;
        umma_ts(acc_w, tW_u + ks * 8, desc_at(dOp[nbuf], (ks >> 2) * (NB8 * 128) + (ks & 3) * 32), IDESC, i ? 1u : 0u);
      }
      umma_commit(barM1);
    }
    __syncwarp();
  };
  auto issue_att = [&]() {  // K steps 2 w, 2 w + 1 of 16
    if (ap4::elect_one()) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int ks = 2 * warp_u + i;
        umma_ss(acc_w, desc_at(dWa, (ks >> 2) * (128 * 128) + (ks & 3) * 32), desc_at(dQ, (ks >> 2) * (NB8 * 128) + (ks & 3) * 32),
                IDESC, i ? 1u : 0u);
      }
      umma_commit(barM2);
    }
    __syncwarp();
  };
  // combine role (threads 0..127): utterance bq, units 4*uq .. 4*uq+3 of the CTA (hidden, and attention of both mechanisms)
  const bool comb = tid < 128;
  const int uq = tid & 7, bq = (tid >> 3) & 15;
  float c_state[4], h_state[4];
  int len_c = 0;
#pragma unroll
  for (int e = 0; e < 4; ++e) c_state[e] = h_state[e] = 0.0f;
  if (comb) {
    const int b = b0 + bq;
    len_c = (b < B) ? p.len[b] : 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int u = UP8 * rank + 4 * uq + e;
      c_state[e] = (b < B && p.c0) ? p.c0[(size_t)b * H + u] : 0.0f;
      h_state[e] = (b < B) ? p.S[(size_t)b * K8 + ATT + u] : 0.0f;
    }
  }
  const uint32_t seed = p.d.rng ? p.d.rng[0] : 0u, rstep = p.d.rng ? p.d.rng[1] : 0u;
  float f_state[4], f_out[4], f_in[NM][4];
  auto drop_factors = [&](int t) {  // state / output masks of step t, input masks of step t + 1 (act on a_t)
    if (comb) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const uint32_t col = (uint32_t)(UP8 * (int)rank + 4 * uq + e);
        const uint32_t idx = (uint32_t)(b0 + bq) * (uint32_t)H + col;
        f_state[e] = dfac8(seed, rstep, p.d.stream + 1u, p.d.thr_state, p.d.inv_state, (uint32_t)t, idx);
        f_out[e] = dfac8(seed, rstep, p.d.stream + 2u, p.d.thr_out, p.d.inv_out, (uint32_t)t, idx);
#pragma unroll
        for (int k = 0; k < NM; ++k)
          f_in[k][e] = dfac8(seed, rstep, p.d.stream, p.d.thr_in, p.d.inv_in, (uint32_t)(t + 1),
                             (uint32_t)(b0 + bq) * (uint32_t)ATT + (uint32_t)(k * H) + col);
      }
    }
  };
#pragma unroll
  for (int e = 0; e < 4; ++e) f_state[e] = f_out[e] = f_in[0][e] = f_in[1][e] = 1.0f;
  drop_factors(0);
  int len_a[8];
#pragma unroll
  for (int b = 0; b < 8; ++b) len_a[b] = (b0 + 8 * hf + b < B) ? p.len[b0 + 8 * hf + b] : 0;
  float gx[8];
  {
    const float* grow0 = p.gates + (size_t)(b0 + 8 * hf) * 4 * H + g * H + unit_g;
#pragma unroll
    for (int b = 0; b < 8; ++b) gx[b] = (0 < len_a[b]) ? grow0[(size_t)b * 4 * H] : 0.0f;
  }
  // attention role: utterance jl of this CTA, warp w4 of its group of four
  const int jl = warp >> 2, w4 = warp & 3, gt = tid & 127;
  const int bl_att = NU8 * (int)rank + jl;
  const int b_att = b0 + bl_att;
  const int len_q = (b_att < B) ? p.len[b_att] : 0;
  float* sc = sc_all + jl * MAX_TM;
  float* part = part_all + jl * 4 * DM;
  float* red = red_all + jl * 8;
  const uint32_t att_bar_id = 2 + jl;
  const int L0 = (b_att < B) ? min(p.mem_len[0][b_att], p.Tm[0]) : 0;
  const int L1 = (b_att < B) ? min(p.mem_len[1][b_att], p.Tm[1]) : 0;
  const AttRole role0 = {p.keys[0], p.pv[0], L0, B, b_att, p.Tm[0], w4, gt, lane, p.scaled[0] ? p.g[0][0] : 1.0f, att_bar_id, sc, part, red};
  const AttRole role1 = {p.keys[1], p.pv[1], L1, B, b_att, p.Tm[1], w4, gt, lane, p.scaled[1] ? p.g[1][0] : 1.0f, att_bar_id, sc, part, red};
  const float vzero8[8] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
  uint32_t lphase = 0u;

  for (int t = 0; t < T; ++t) {
    float* grow = p.gates + ((size_t)t * B + b0 + 8 * hf) * 4 * H + g * H + unit_g;
    uint32_t selmask = 0u;  // SAMPLE: utterances of the cluster whose next input is drawn from this step's logits
    if constexpr (SAMPLE) {
      if (t + 1 < T) {
#pragma unroll
        for (int b = 0; b < NB8; ++b)
          if (b0 + b < B && avsr_rand_u32(seed, rstep, p.ss_stream, (uint32_t)t, (uint32_t)(b0 + b)) < p.thr_p) selmask |= 1u << b;
      }
    }
    float z[8];
#pragma unroll
    for (int b = 0; b < 8; ++b) z[b] = 0.0f;
    if (t > 0) {
      mbar_wait(barM1, (t - 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a0 = tmem_base + ((uint32_t)(32 * g) << 16) + 8 * hf;
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2) {
        uint32_t r0[8], r1[8], r2[8], r3[8];
        tmem_ld8(a0 + (4 * h2 + 0) * NP, r0);
        tmem_ld8(a0 + (4 * h2 + 1) * NP, r1);
        tmem_ld8(a0 + (4 * h2 + 2) * NP, r2);
        tmem_ld8(a0 + (4 * h2 + 3) * NP, r3);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int b = 0; b < 8; ++b)
          z[b] += (__uint_as_float(r0[b]) + __uint_as_float(r1[b])) + (__uint_as_float(r2[b]) + __uint_as_float(r3[b]));
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    float av[8];
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      const float zz = z[b] + gx[b];
      float a;
      if (g == 1) a = tanhf_acc(zz);
      else a = sigmoidf_acc(g == 2 ? zz + 1.0f : zz);
      av[b] = a;
      act[(g * NB8 + 8 * hf + b) * UP8 + lane] = a;
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const uint32_t nb = (t + 1) & 1;
    const uint32_t hbar_n = sBar + 16 + 8 * nb;
    float hs[4], ho[4], cr[4];
    if (comb) {
      const bool live = t < len_c;
      if (live) {
        const float4 ai = *reinterpret_cast<const float4*>(&act[(0 * NB8 + bq) * UP8 + 4 * uq]);
        const float4 aj = *reinterpret_cast<const float4*>(&act[(1 * NB8 + bq) * UP8 + 4 * uq]);
        const float4 af = *reinterpret_cast<const float4*>(&act[(2 * NB8 + bq) * UP8 + 4 * uq]);
        const float4 ao = *reinterpret_cast<const float4*>(&act[(3 * NB8 + bq) * UP8 + 4 * uq]);
        const float vi[4] = {ai.x, ai.y, ai.z, ai.w}, vj[4] = {aj.x, aj.y, aj.z, aj.w};
        const float vf[4] = {af.x, af.y, af.z, af.w}, vo[4] = {ao.x, ao.y, ao.z, ao.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          cr[e] = vf[e] * c_state[e] + vi[e] * vj[e];
          const float c = fminf(fmaxf(cr[e], -1.0f), 1.0f);
          const float h = vo[e] * tanhf_acc(c);
          c_state[e] = c;
          ho[e] = tf32_rn(h * f_out[e]);
          h_state[e] = tf32_rn(h * f_state[e]);
        }
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          cr[e] = c_state[e];
          ho[e] = h_state[e];
        }
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) hs[e] = h_state[e];
      const int ucol = UP8 * (int)rank + 4 * uq;
      const uint32_t s01 = pack_h2(hs[0], hs[1]), s23 = pack_h2(hs[2], hs[3]);
      const uint32_t o01 = pack_h2(ho[0], ho[1]), o23 = pack_h2(ho[2], ho[3]);
      const uint32_t dS = sOp + nb * OP8_BYTES + sw128h_off(NB8, bq, ATT + ucol);
      const uint32_t dO = sQ + sw128h_off(NB8, bq, ucol);
#pragma unroll
      for (uint32_t dst = 0; dst < (uint32_t)CL8; ++dst) {
        const uint32_t bar = mapa(hbar_n, dst);
        st_async_v2(mapa(dS, dst), bar, s01, s23);
        st_async_v2(mapa(dO, dst), bar, o01, o23);
      }
    }
    // HBM side of this step + x-projection of the next (overlaps the all-gather)
#pragma unroll
    for (int b = 0; b < 8; ++b)
      if (t < len_a[b]) grow[(size_t)b * 4 * H] = av[b];
    if (comb && b0 + bq < B) {
      const size_t row = (size_t)t * B + b0 + bq;
      const int u0 = UP8 * rank + 4 * uq;
      *reinterpret_cast<float4*>(p.craw + row * H + u0) = make_float4(cr[0], cr[1], cr[2], cr[3]);
      *reinterpret_cast<float4*>(p.S + (row + B) * K8 + ATT + u0) = make_float4(hs[0], hs[1], hs[2], hs[3]);
      *reinterpret_cast<float4*>(p.hc[0] + row * p.ldhc[0] + u0) = make_float4(ho[0], ho[1], ho[2], ho[3]);
      *reinterpret_cast<float4*>(p.hc[1] + row * p.ldhc[1] + u0) = make_float4(ho[0], ho[1], ho[2], ho[3]);
    }
    if (t + 1 < T) {
      const float* gnext = grow + (size_t)B * 4 * H;
#pragma unroll
      for (int b = 0; b < 8; ++b) gx[b] = (t + 1 < len_a[b]) ? gnext[(size_t)b * 4 * H] : 0.0f;
    }
    // ---------------- attention of utterance b_att with query ho_t, both mechanisms ----------------
    const bool live_q = t < len_q;
    uint4 ra[4], rb[4];
    if (live_q) att_prefetch(role0, p.keys[0], ra, rb);
    if (tid == 0) mbar_expect_tx(hbar_n, 2 * NB8 * H * 2);
    mbar_wait(hbar_n, (t >> 1) & 1);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    issue_att();  // ha_k = ho Wl_h,k: in flight during the sweeps
    float q[8];
    if (live_q) {
      const uint4 qraw = *reinterpret_cast<const uint4*>(gen + (sQ - base) + sw128h_off(NB8, bl_att, 8 * lane));
      unpack_q(qraw, q);
    }
#pragma unroll
    for (int k = 0; k < NM; ++k) {
      float ctxv[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) ctxv[e] = 0.0f;
      float* arow = p.align[k] + ((size_t)t * B + b_att) * p.Tm[k];
      if (live_q) att_fwd_core<false, true>(k == 0 ? role0 : role1, q, vzero8, ra, rb, arow, ctxv);
      if (k == 0 && live_q) att_prefetch(role1, p.keys[1], ra, rb);
      if (w4 == 0) {
        if (b_att < B && !live_q)
          for (int tm = lane; tm < p.Tm[k]; tm += 32) arow[tm] = 0.0f;
        // ctx'_k dims 8*lane .. +7 -> owner of attention units (8*lane)/32
        const uint32_t dst = (uint32_t)(lane >> 2);
        const uint32_t a0 = mapa(sCxA + (uint32_t)(((k * NB8 + bl_att) * UP8 + ((8 * lane) & (UP8 - 1))) * 4), dst);
        const uint32_t bar = mapa(barCtx, dst);
        st_async_v4f(a0, bar, ctxv[0], ctxv[1], ctxv[2], ctxv[3]);
        st_async_v4f(a0 + 16, bar, ctxv[4], ctxv[5], ctxv[6], ctxv[7]);
      }
    }
    // masks of the next step: off the critical chain (this step's a_t needs f_in, kept in `fi`)
    float fi[NM][4];
#pragma unroll
    for (int k = 0; k < NM; ++k)
#pragma unroll
      for (int e = 0; e < 4; ++e) fi[k][e] = f_in[k][e];
    drop_factors(t + 1);
    // ---------------- ha_k of the CTA's attention units: accumulators -> planes [mech][b][u] ----------------
    mbar_wait(barM2, t & 1);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (g < NM) {  // lane quarter = mechanism (rows 32 mech + u); this warp's utterances 8 hf .. +7
      const uint32_t a0 = tmem_base + ((uint32_t)(32 * g) << 16) + 8 * hf;
      float s[8];
#pragma unroll
      for (int b = 0; b < 8; ++b) s[b] = 0.0f;
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2) {
        uint32_t r0[8], r1[8], r2[8], r3[8];
        tmem_ld8(a0 + (4 * h2 + 0) * NP, r0);
        tmem_ld8(a0 + (4 * h2 + 1) * NP, r1);
        tmem_ld8(a0 + (4 * h2 + 2) * NP, r2);
        tmem_ld8(a0 + (4 * h2 + 3) * NP, r3);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int b = 0; b < 8; ++b)
          s[b] += (__uint_as_float(r0[b]) + __uint_as_float(r1[b])) + (__uint_as_float(r2[b]) + __uint_as_float(r3[b]));
      }
#pragma unroll
      for (int b = 0; b < 8; ++b) apart[(g * NB8 + 8 * hf + b) * UP8 + lane] = s[b];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (tid == 0) mbar_expect_tx(barCtx, NM * NB8 * UP8 * 4);
    mbar_wait(barCtx, t & 1);
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (comb) {
      const int ucol = UP8 * (int)rank + 4 * uq;
      const bool live = t < len_c;
#pragma unroll
      for (int k = 0; k < NM; ++k) {
        const float4 p0 = *reinterpret_cast<const float4*>(&apart[(k * NB8 + bq) * UP8 + 4 * uq]);
        const float4 p1 = *reinterpret_cast<const float4*>(&sCx[(k * NB8 + bq) * UP8 + 4 * uq]);
        float a[4] = {tf32_rn(p0.x + p1.x), tf32_rn(p0.y + p1.y), tf32_rn(p0.z + p1.z), tf32_rn(p0.w + p1.w)};
        float ad[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) ad[e] = tf32_rn(a[e] * fi[k][e]);  // input dropout of step t+1 on the fed-back attention
        const uint32_t dA = sOp + nb * OP8_BYTES + sw128h_off(NB8, bq, k * H + ucol);
        const uint32_t a01 = pack_h2(ad[0], ad[1]), a23 = pack_h2(ad[2], ad[3]);
#pragma unroll
        for (uint32_t dst = 0; dst < (uint32_t)CL8; ++dst) st_async_v2(mapa(dA, dst), mapa(barA, dst), a01, a23);
        if (b0 + bq < B) {
          const size_t row = (size_t)t * B + b0 + bq;
          *reinterpret_cast<float4*>(p.out + row * ATT + k * H + ucol) =
              live ? make_float4(a[0], a[1], a[2], a[3]) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
          *reinterpret_cast<float4*>(p.S + (row + B) * K8 + k * H + ucol) = make_float4(ad[0], ad[1], ad[2], ad[3]);
        }
        if constexpr (SAMPLE) {
          if (selmask) *reinterpret_cast<float4*>(&sAf[bq * (NM * UP8) + k * UP8 + 4 * uq]) = make_float4(a[0], a[1], a[2], a[3]);
        }
      }
    }
    if constexpr (SAMPLE) {
      if (selmask) {  // uniform over the cluster
        asm volatile("bar.sync 1, 256;" ::: "memory");
        // partial logits of the selected utterances over this CTA's 64 attention units: warp = utterances w, w + 8
#pragma unroll
        for (int rep = 0; rep < 2; ++rep) {
          const int bb = warp + 8 * rep;
          if ((selmask >> bb) & 1u) {
            float sacc = 0.0f;
#pragma unroll 8
            for (int u = 0; u < NM * UP8; ++u) sacc = fmaf(sAf[bb * (NM * UP8) + u], sWd[u * S8_VP + lane], sacc);
            const uint32_t dstoff = sLgAddr + (uint32_t)(((int)rank * NB8 + bb) * S8_VP + lane) * 4;
#pragma unroll
            for (uint32_t dst = 0; dst < (uint32_t)CL8; ++dst) st_async_f(mapa(dstoff, dst), mapa(barL, dst), sacc);
          }
        }
        if (tid == 0) mbar_expect_tx(barL, (uint32_t)__popc(selmask) * S8_VP * 4 * CL8);
        mbar_wait(barL, lphase & 1);
        ++lphase;
#pragma unroll 1
        for (int rep = 0; rep < 2; ++rep) {
          const int bb = warp + 8 * rep;
          if (((selmask >> bb) & 1u) && lane == 0) {
            // inverse CDF of the fp32 softmax, summed in class order (the arithmetic of sched_sample_kernel, misc.cu)
            float zl[S8_VP];
            float mx = -INFINITY;
            for (int v = 0; v < p.V; ++v) {
              float s = 0.0f;
              for (int src = 0; src < CL8; ++src) s += sLg[(src * NB8 + bb) * S8_VP + v];
              zl[v] = p.bd[v] + s;
              mx = fmaxf(mx, zl[v]);
            }
            float total = 0.0f;
            for (int v = 0; v < p.V; ++v) total += expf(zl[v] - mx);
            const float u01 = (float)(avsr_rand_u32(seed, rstep, p.ss_stream + 1u, (uint32_t)t, (uint32_t)(b0 + bb)) >> 8) * (1.0f / 16777216.0f);
            const float target = u01 * total;
            float cum = 0.0f;
            int pick = p.V - 1;
            for (int v = 0; v < p.V; ++v) {
              cum += expf(zl[v] - mx);
              if (cum > target) {
                pick = v;
                break;
              }
            }
            sPick[bb] = pick;
            if (rank == 0) {
              p.sample_ids[(size_t)t * B + b0 + bb] = pick;
              p.used_ids[(size_t)(t + 1) * B + b0 + bb] = pick;
            }
          }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        // x-projection of the drawn embeddings for this thread's gate row (g, unit_g): replaces the prefetched one
        const float* wxcol = p.Wx + g * H + unit_g;
        const float bias_row = p.bias[g * H + unit_g];
#pragma unroll 1
        for (int bb = 0; bb < NB8; ++bb) {
          if (!((selmask >> bb) & 1u)) continue;
          const int id = sPick[bb];
          for (int e = tid; e < p.E; e += THREADS) {
            const uint32_t lo = (uint32_t)(((size_t)(t + 1) * B + b0 + bb) * p.E + e);
            const float xv = tf32_rn(p.emb[(size_t)id * p.E + e] * dfac8(seed, rstep, p.d.stream + 3u, p.d.thr_in, p.d.inv_in, 0u, lo));
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
          if ((bb >> 3) == hf) {
#pragma unroll
            for (int b = 0; b < 8; ++b)
              if (b == (bb & 7)) gx[b] = (t + 1 < len_a[b]) ? acc : 0.0f;
          }
          asm volatile("bar.sync 1, 256;" ::: "memory");
        }
      }
    }
    // recurrent product of step t+1 once every CTA's a_t slices have landed (hs_t landed before the attention).  After
    // the last step the wait only drains the all-gather: no st.async may be in flight towards a CTA that exits.
    if (tid == 0) mbar_expect_tx(barA, NB8 * ATT * 2);
    mbar_wait(barA, t & 1);
    if (t + 1 < T) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      issue_rec(nb);
    }
  }
  if (comb && b0 + bq < B) {
    const size_t o = (size_t)(b0 + bq) * H + UP8 * rank + 4 * uq;
    if (p.cT) *reinterpret_cast<float4*>(p.cT + o) = make_float4(c_state[0], c_state[1], c_state[2], c_state[3]);
    if (p.hT) *reinterpret_cast<float4*>(p.hT + o) = make_float4(h_state[0], h_state[1], h_state[2], h_state[3]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  cluster_sync_all();
}

// =====================================================================================================
// backward.  Per step t (descending), with dS_t = [dSa_0 | dSa_1 | dSh] = dz_{t+1} [Wl_att ; Wh]^T from the previous
// iteration (K split over the cluster, reduce-scattered to the owners of the rows):
//   da_k,t  = dSa_k,t (.) m_in(t+1) + dout_k,t          kept for dWl_k / dctx_k (host, after the loop); to the owners of
//                                                        the utterances for the attention backward
//   d ho    = sum_k da_k Wl_h,k^T                        product 2 (K = the CTA's 64 attention units), reduce-scattered
//   per mechanism: d(align) = PV_k . da_k -> softmax backward -> ds_k -> dq_k = g_k ds_k^T keys_k (wrt the query ho)
//   dh_t    = (d ho + dq_0 + dq_1) (.) m_out(t) + dSh_t (.) m_state(t) -> gate gradients dz_t
//   dS_{t-1} partial = dz_t(CTA's 128 gate columns) [Wl_att ; Wh]^T     product 1: six 128-row tiles in tensor memory
// =====================================================================================================
struct WBwdParams {
  int T, B;
  int Tm[NM], scaled[NM];
  float grad_scale, inv_grad_scale;
  const int* len;
  const int* mem_len[NM];
  const float* gates;
  const float* craw;
  const float* c0;
  const float* Wrec;
  const float* Wa[NM];
  const __half* keys[NM];
  const __half* pv[NM];
  const float* g[NM];
  const float* align[NM];
  const float* dout;        // [T,B,ATT] or null
  const float* dcT;
  const float* dhT;
  float* dZ;                // [T,B,4H]
  float* ds[NM];            // [T,B,Tm_k]
  float* dA;                // [T,B,ATT]
  float* dg[NM];
  float* dc0;
  float* dh0;
  float* dbias;
  Drop8 d;
};

constexpr int BW8_WT_BYTES = 2 * 128 * 128;              // two 128-row tiles (ho dims) of [Wl_h,0^T | Wl_h,1^T], K = 64
constexpr int BW8_DZ_BYTES = 2 * NB8 * 128;              // B of product 1: K = 128 gate columns
constexpr int BW8_DA_BYTES = NB8 * 128;                  // B of product 2: K = 64 attention units
constexpr int RED8A_FLOATS = CL8 * NM * NB8 * UP8;       // [src][mech][b][u]
constexpr int RED8H_FLOATS = CL8 * NB8 * UP8;            // [src][b][u]
constexpr size_t WBWD_SMEM = (size_t)BW8_WT_BYTES + BW8_DZ_BYTES + BW8_DA_BYTES + RED8A_FLOATS * 4 + 2 * RED8H_FLOATS * 4 +
                             NB8 * UP8 * 4 + NU8 * ATT * 4 + NM * NU8 * MAX_TM * 4 + NU8 * MAX_TM * 4 + NU8 * 4 * DM * 4 +
                             NU8 * 8 * 4 + 128 + 1024;

__global__ void __launch_bounds__(THREADS, 1) wlas_persist8_bwd_kernel(const WBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sWt = base;
  const uint32_t sDz = sWt + BW8_WT_BYTES;
  const uint32_t sDa = sDz + BW8_DZ_BYTES;
  const uint32_t sRedA = sDa + BW8_DA_BYTES;                 // partial dSa of the CTA's attention units
  const uint32_t sRedH = sRedA + RED8A_FLOATS * 4;           // partial dSh of the CTA's hidden units
  const uint32_t sRedH2 = sRedH + RED8H_FLOATS * 4;          // partial d ho of the CTA's hidden units
  const uint32_t sDq = sRedH2 + RED8H_FLOATS * 4;            // dq of the CTA's hidden units [b][u]
  const uint32_t sDaU = sDq + NB8 * UP8 * 4;                 // da of the CTA's utterances [NU8][ATT]
  const uint32_t sSc = sDaU + NU8 * ATT * 4;                 // alignments [NM][NU8][MAX_TM]
  const uint32_t sDs = sSc + NM * NU8 * MAX_TM * 4;          // d(align) / ds [NU8][MAX_TM]
  const uint32_t sPart = sDs + NU8 * MAX_TM * 4;             // [NU8][4][DM]
  const uint32_t sRed = sPart + NU8 * 4 * DM * 4;            // [NU8][8]
  const uint32_t sBar = sRed + NU8 * 8 * 4;
  const uint32_t sTmem = sBar + 64;
  const uint32_t barMma = sBar, barMma2 = sBar + 8, barDz = sBar + 16, barRedA = sBar + 24, barRedH = sBar + 32,
                 barRedH2 = sBar + 40, barDaU = sBar + 48, barDq = sBar + 56;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  float* redA = reinterpret_cast<float*>(gen + (sRedA - base));
  float* redH = reinterpret_cast<float*>(gen + (sRedH - base));
  float* redH2 = reinterpret_cast<float*>(gen + (sRedH2 - base));
  float* dqb = reinterpret_cast<float*>(gen + (sDq - base));
  float* dau_all = reinterpret_cast<float*>(gen + (sDaU - base));
  float* as_all = reinterpret_cast<float*>(gen + (sSc - base));
  float* ds_all = reinterpret_cast<float*>(gen + (sDs - base));
  float* part_all = reinterpret_cast<float*>(gen + (sPart - base));
  float* red_all = reinterpret_cast<float*>(gen + (sRed - base));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int b0 = cluster_id_x() * NB8;
  const int T = p.T, B = p.B;

  if (tid == 0) {
    mbar_init(barMma, 6);               // warps 0..5 issue one tile each
    mbar_init(barMma2, THREADS / 32);
    mbar_init(barDz, THREADS);
    mbar_init(barRedA, 1);
    mbar_init(barRedH, 1);
    mbar_init(barRedH2, 1);
    mbar_init(barDaU, 1);
    mbar_init(barDq, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // tensor memory: [0, 128) accumulators of both products; [128, 512) the six 128-row tiles of Wrec^T restricted to the
  // CTA's 128 gate columns (K = 128: 64 columns each)
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(sTmem) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // A of product 2: row n (ho dim) of tile n >> 7, k = mech*32 + u <-> Wa_mech[n][32*rank + u]
  for (int seg = warp; seg < H * NM; seg += THREADS / 32) {
    const int n = seg >> 1, mech = seg & 1;
    const float w = p.Wa[mech][(size_t)n * H + UP8 * rank + lane];
    *reinterpret_cast<__half*>(gen + (sWt - base) + (n >> 7) * (128 * 128) + sw128h_off(128, n & 127, mech * 32 + lane)) =
        __float2half_rn(w);
  }
  for (int i = tid; i < (BW8_DZ_BYTES + BW8_DA_BYTES) / 4; i += THREADS) reinterpret_cast<uint32_t*>(gen + (sDz - base))[i] = 0u;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(sTmem));
  const uint32_t tA = tmem_base + 128;
  {
    // rows n = 128*mt + 32*q + lane of Wrec -> tile mt; column c holds k = 2c, 2c+1 (k = gate*32 + u: adjacent units)
    const int q = warp & 3, t3 = warp >> 2;
#pragma unroll 1
    for (int i = 0; i < 3; ++i) {
      const int mt = 3 * t3 + i;
      const float* row = p.Wrec + (size_t)(128 * mt + 32 * q + lane) * 4 * H + UP8 * rank;
#pragma unroll 1
      for (int c0 = 0; c0 < 64; c0 += 32) {
        uint32_t r[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const int k = 2 * (c0 + c), gg = k >> 5, u = k & 31;
          const float2 w = *reinterpret_cast<const float2*>(row + gg * H + u);
          r[c] = pack_h2(w.x, w.y);
        }
        tmem_st32(tA + 64 * mt + c0 + ((uint32_t)(32 * q) << 16), r);
      }
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  cluster_sync_all();

  const uint64_t dWt = make_desc_k128(sWt), dDz = make_desc_k128(sDz), dDa = make_desc_k128(sDa);
  // gate-gradient role: thread = (local unit ul - hidden, and attention of both mechanisms -, utterances 2*warp + j)
  constexpr int PB = 2;
  const int ul = lane;
  const int unit = UP8 * rank + ul;
  float dc[PB], dh_carry[PB];
  int len_t[PB];
#pragma unroll
  for (int j = 0; j < PB; ++j) {
    const int b = b0 + warp * PB + j;
    len_t[j] = (b < B) ? p.len[b] : 0;
    dc[j] = (b < B && p.dcT) ? p.dcT[(size_t)b * H + unit] : 0.0f;
    dh_carry[j] = (b < B && p.dhT) ? p.dhT[(size_t)b * H + unit] : 0.0f;
  }
  const uint32_t seed = p.d.rng ? p.d.rng[0] : 0u, rstep = p.d.rng ? p.d.rng[1] : 0u;
  float gi[PB], gj[PB], gf[PB], go[PB], crw[PB], cpv[PB], dov[NM][PB];
#pragma unroll
  for (int j = 0; j < PB; ++j) gi[j] = gj[j] = gf[j] = go[j] = crw[j] = cpv[j] = dov[0][j] = dov[1][j] = 0.0f;
  auto load_step = [&](int t) {
#pragma unroll
    for (int j = 0; j < PB; ++j) {
      const int b = b0 + warp * PB + j;
      if (t >= 0 && t < len_t[j]) {
        const float* gp = p.gates + ((size_t)t * B + b) * 4 * H + unit;
        gi[j] = gp[0]; gj[j] = gp[H]; gf[j] = gp[2 * H]; go[j] = gp[3 * H];
        const size_t o = ((size_t)t * B + b) * H + unit;
        crw[j] = p.craw[o];
        cpv[j] = t > 0 ? p.craw[o - (size_t)B * H] : (p.c0 ? p.c0[(size_t)b * H + unit] : 0.0f);
#pragma unroll
        for (int k = 0; k < NM; ++k) dov[k][j] = p.dout ? p.dout[((size_t)t * B + b) * ATT + k * H + unit] : 0.0f;
      }
    }
  };
  // attention role
  const int jl = warp >> 2, w4 = warp & 3, gt = tid & 127;
  const int bl_att = NU8 * (int)rank + jl;
  const int b_att = b0 + bl_att;
  const int len_q = (b_att < B) ? p.len[b_att] : 0;
  float* dau_s = dau_all + jl * ATT;
  float* ds_s = ds_all + jl * MAX_TM;
  float* part = part_all + jl * 4 * DM;
  float* red = red_all + jl * 8;
  const uint32_t att_bar_id = 2 + jl;
  const int L0 = (b_att < B) ? min(p.mem_len[0][b_att], p.Tm[0]) : 0;
  const int L1 = (b_att < B) ? min(p.mem_len[1][b_att], p.Tm[1]) : 0;
  // (`values` = what the first sweep multiplies with da: the projected values)
  const AttRole role0 = {p.keys[0], p.pv[0], L0, B, b_att, p.Tm[0], w4, gt, lane, p.scaled[0] ? p.g[0][0] : 1.0f, att_bar_id, nullptr, part, red};
  const AttRole role1 = {p.keys[1], p.pv[1], L1, B, b_att, p.Tm[1], w4, gt, lane, p.scaled[1] ? p.g[1][0] : 1.0f, att_bar_id, nullptr, part, red};
  const float vzero8[8] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
  const int q = warp & 3, hf = warp >> 2;
  const int warp_u = (int)ap4::warp_uniform((uint32_t)warp);  // (issue: elected lane, warp-uniform operands)
  const uint32_t tmem_u = ap4::warp_uniform(tmem_base), tA_u = tmem_u + 128;

  float bsum[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  load_step(T - 1);
  for (int it = 0; it < T; ++it) {
    const int t = T - 1 - it;
    const bool live_q = t < len_q;
    uint4 ra[4], rb[4];
    if (live_q) {
      att_prefetch(role0, p.pv[0], ra, rb);
#pragma unroll
      for (int k = 0; k < NM; ++k) {
        float* a_s = as_all + (k * NU8 + jl) * MAX_TM;
        const float* arow = p.align[k] + ((size_t)t * B + b_att) * p.Tm[k];
        for (int tm = gt; tm < p.Tm[k]; tm += 128) a_s[tm] = arow[tm];
      }
    }
    float f_in[NM][PB], f_st[PB], f_o[PB];
#pragma unroll
    for (int j = 0; j < PB; ++j) {
      const uint32_t bg = (uint32_t)(b0 + warp * PB + j);
      const uint32_t idx = bg * (uint32_t)H + (uint32_t)unit;
#pragma unroll
      for (int k = 0; k < NM; ++k)
        f_in[k][j] = dfac8(seed, rstep, p.d.stream, p.d.thr_in, p.d.inv_in, (uint32_t)(t + 1), bg * (uint32_t)ATT + (uint32_t)(k * H + unit));
      f_st[j] = dfac8(seed, rstep, p.d.stream + 1u, p.d.thr_state, p.d.inv_state, (uint32_t)t, idx);
      f_o[j] = dfac8(seed, rstep, p.d.stream + 2u, p.d.thr_out, p.d.inv_out, (uint32_t)t, idx);
    }
    // ---- (A) dS_t pushed during the previous iteration -> da_k,t (operand of product 2; to the utterances' owners) ----
    float dh_in[PB], sa[NM][PB];
#pragma unroll
    for (int j = 0; j < PB; ++j) {
      dh_in[j] = dh_carry[j];
      sa[0][j] = sa[1][j] = 0.0f;
    }
    if (it > 0) {
      if (tid == 0) {
        mbar_expect_tx(barRedA, RED8A_FLOATS * 4);
        mbar_expect_tx(barRedH, RED8H_FLOATS * 4);
      }
      mbar_wait(barRedA, (it - 1) & 1);
#pragma unroll
      for (int j = 0; j < PB; ++j) {
        const int bl = warp * PB + j;
#pragma unroll
        for (int k = 0; k < NM; ++k)
#pragma unroll
          for (int src = 0; src < CL8; ++src) sa[k][j] += redA[((src * NM + k) * NB8 + bl) * UP8 + ul];
      }
    }
#pragma unroll
    for (int j = 0; j < PB; ++j) {
      const int bl = warp * PB + j;
      const uint32_t dst = (uint32_t)(bl / NU8);
#pragma unroll
      for (int k = 0; k < NM; ++k) {
        const float da = (t < len_t[j]) ? tf32_rn(sa[k][j] * f_in[k][j] + dov[k][j]) : 0.0f;
        if (b0 + bl < B) p.dA[((size_t)t * B + b0 + bl) * ATT + k * H + unit] = da;
        *reinterpret_cast<__half*>(gen + (sDa - base) + sw128h_off(NB8, bl, k * UP8 + ul)) = __float2half_rn(da * p.grad_scale);
        st_async_f(mapa(sDaU + (uint32_t)(((bl % NU8) * ATT + k * H + unit) * 4), dst), mapa(barDaU, dst), da);
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    asm volatile("bar.sync 1, 256;" ::: "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // ---- (B) product 2: d ho partial from the CTA's 64 attention units; warp (tile hf, K step q) ----------------------
    if (ap4::elect_one()) {
      umma_ss(tmem_u + warp_u * NP, desc_at(dWt, (warp_u >> 2) * (128 * 128) + (warp_u & 3) * 32), desc_at(dDa, (warp_u & 3) * 32),
              IDESC, 0u);
      umma_commit(barMma2);
    }
    __syncwarp();
    if (it > 0) {
      mbar_wait(barRedH, (it - 1) & 1);
#pragma unroll
      for (int j = 0; j < PB; ++j) {
        const int bl = warp * PB + j;
#pragma unroll
        for (int src = 0; src < CL8; ++src) dh_in[j] += redH[(src * NB8 + bl) * UP8 + ul];
      }
    }
    // ---- (C) partial d ho -> owners of the hidden units: warp (tile hf, lane quarter q): dims 128 hf + 32 q + lane ------
    mbar_wait(barMma2, it & 1);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    {
      const uint32_t dst = (uint32_t)(4 * hf + q);
      const uint32_t d0 = mapa(sRedH2 + (uint32_t)((rank * NB8) * UP8 + lane) * 4, dst);
      const uint32_t bar = mapa(barRedH2, dst);
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2) {
        uint32_t r0[8], r1[8], r2[8], r3[8];
        const uint32_t a0 = tmem_base + ((uint32_t)(32 * q) << 16) + (4 * hf) * NP + 8 * h2;
        tmem_ld8(a0, r0);
        tmem_ld8(a0 + NP, r1);
        tmem_ld8(a0 + 2 * NP, r2);
        tmem_ld8(a0 + 3 * NP, r3);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int c = 0; c < 8; ++c)
          st_async_f(d0 + (8 * h2 + c) * UP8 * 4, bar,
                     ((__uint_as_float(r0[c]) + __uint_as_float(r1[c])) + (__uint_as_float(r2[c]) + __uint_as_float(r3[c]))) * p.inv_grad_scale);
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    // ---- (E) attention backward of the CTA's utterances, both mechanisms; dq all-to-all ---------------------------------
    if (tid == 0) mbar_expect_tx(barDaU, NU8 * ATT * 4);
    mbar_wait(barDaU, it & 1);
    float dq[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) dq[e] = 0.0f;
    if (live_q) {
#pragma unroll
      for (int k = 0; k < NM; ++k) {
        float dqv[8], al1[1] = {0.0f}, dk1[1] = {0.0f};
#pragma unroll
        for (int e = 0; e < 8; ++e) dqv[e] = 0.0f;
        att_bwd_core<false, false>(k == 0 ? role0 : role1, dau_s + k * H, ra, rb, al1, dk1, as_all + (k * NU8 + jl) * MAX_TM, ds_s,
                                   p.ds[k] + ((size_t)t * B + b_att) * p.Tm[k], p.scaled[k] != 0, p.dg[k], dqv, vzero8, vzero8);
        if (k == 0) att_prefetch(role1, p.pv[1], ra, rb);
#pragma unroll
        for (int e = 0; e < 8; ++e) dq[e] += dqv[e];
      }
    }
    if (w4 == 0) {
      // dq dims 8*lane .. +7 belong to the CTA owning hidden units (8*lane)/32
      const uint32_t dst = (uint32_t)(lane >> 2);
      const uint32_t a0 = mapa(sDq + (uint32_t)((bl_att * UP8 + ((8 * lane) & (UP8 - 1))) * 4), dst);
      const uint32_t bar = mapa(barDq, dst);
      st_async_v4f(a0, bar, dq[0], dq[1], dq[2], dq[3]);
      st_async_v4f(a0 + 16, bar, dq[4], dq[5], dq[6], dq[7]);
    }
    // ---- (F) d ho + dq of this CTA's units -> gate gradients ---------------------------------------------------------------
    if (tid == 0) {
      mbar_expect_tx(barDq, NB8 * UP8 * 4);
      mbar_expect_tx(barRedH2, RED8H_FLOATS * 4);
    }
    mbar_wait(barRedH2, it & 1);
    mbar_wait(barDq, it & 1);
    float dz[4][PB];
#pragma unroll
    for (int j = 0; j < PB; ++j) {
      const int bl = warp * PB + j;
      if (t < len_t[j]) {
        float dho = dqb[bl * UP8 + ul];
#pragma unroll
        for (int src = 0; src < CL8; ++src) dho += redH2[(src * NB8 + bl) * UP8 + ul];
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
      for (int gg = 0; gg < 4; ++gg)
        *reinterpret_cast<__half*>(gen + (sDz - base) + sw128h_off(NB8, bl, gg * UP8 + ul)) = __float2half_rn(dz[gg][j] * p.grad_scale);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    mbar_arrive(barDz);
    {
      // product 1: partial dS_{t-1} (768 x 16) from this CTA's 128 gate columns; warp mt < 6 issues the 8 K steps of tile mt
      mbar_wait(barDz, it & 1);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (warp_u < 6) {
        if (ap4::elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)
            umma_ts(tmem_u + warp_u * NP, tA_u + 64 * warp_u + ks * 8, desc_at(dDz, (ks >> 2) * (NB8 * 128) + (ks & 3) * 32), IDESC,
                    ks ? 1u : 0u);
          umma_commit(barMma);
        }
        __syncwarp();
      }
    }
#pragma unroll
    for (int j = 0; j < PB; ++j) {
      const int b = b0 + warp * PB + j;
      if (b < B) {
        float* o = p.dZ + ((size_t)t * B + b) * 4 * H + unit;
        o[0] = tf32_rn(dz[0][j]); o[H] = tf32_rn(dz[1][j]); o[2 * H] = tf32_rn(dz[2][j]); o[3 * H] = tf32_rn(dz[3][j]);
#pragma unroll
        for (int gg = 0; gg < 4; ++gg) bsum[gg] += tf32_rn(dz[gg][j]);
      }
    }
    load_step(t - 1);
    // ---- (G) partial dS_{t-1} -> owners of the attention / hidden units -----------------------------------------------------
    mbar_wait(barMma, it & 1);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (it + 1 < T) {
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int mt = hf + 2 * i;  // tiles 0, 1: mechanism 0; 2, 3: mechanism 1; 4, 5: hidden units
        const uint32_t dst = (uint32_t)(4 * (mt & 1) + q);
        const uint32_t buf = mt < 4 ? sRedA + (uint32_t)(((rank * NM + (mt >> 1)) * NB8) * UP8 + lane) * 4
                                    : sRedH + (uint32_t)((rank * NB8) * UP8 + lane) * 4;
        const uint32_t d0 = mapa(buf, dst);
        const uint32_t bar = mapa(mt < 4 ? barRedA : barRedH, dst);
        uint32_t r0[8], r1[8];
        const uint32_t a0 = tmem_base + ((uint32_t)(32 * q) << 16) + mt * NP;
        tmem_ld8(a0, r0);
        tmem_ld8(a0 + 8, r1);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          st_async_f(d0 + c * UP8 * 4, bar, __uint_as_float(r0[c]) * p.inv_grad_scale);
          st_async_f(d0 + (8 + c) * UP8 * 4, bar, __uint_as_float(r1[c]) * p.inv_grad_scale);
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
#pragma unroll
  for (int j = 0; j < PB; ++j) {
    const int b = b0 + warp * PB + j;
    if (b < B) {
      // dh_0 = dz_0 Wh^T is added by the host; here only what fully masked utterances carry through
      if (p.dh0) p.dh0[(size_t)b * H + unit] = dh_carry[j];
      if (p.dc0) p.dc0[(size_t)b * H + unit] = dc[j];
    }
  }
  if (p.dbias) {
#pragma unroll
    for (int gg = 0; gg < 4; ++gg) atomicAdd(p.dbias + gg * H + unit, bsum[gg]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  cluster_sync_all();
}

template <typename Kern, typename P>
static int launch_cluster8(cudaStream_t st, Kern kern, int B, size_t smem, const P& p, int klass) {
  AVSR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(cdiv(B, NB8) * CL8);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CL8;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  const int slot = kernel_timer_begin(st, klass);
  AVSR_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
  kernel_timer_end(st, slot);
  ++g_launch_count;
  return 0;
}

static Drop8 drop_cfg8(const AvsrRnnSeq* r) {
  Drop8 d;
  d.rng = r->rng;
  d.stream = r->drop_stream;
  d.thr_in = r->rng ? r->thr_in : 0u; d.thr_state = r->rng ? r->thr_state : 0u; d.thr_out = r->rng ? r->thr_out : 0u;
  d.inv_in = inv_keep_of(d.thr_in); d.inv_state = inv_keep_of(d.thr_state); d.inv_out = inv_keep_of(d.thr_out);
  return d;
}

}  // namespace ap8

// keys_h[k] / pv_h[k]: fp16 keys and fp16 projected values PV_k = values_k Wl_c,k of mechanism k
int wlas_persist8_launch_fwd(cudaStream_t st, const AvsrRnnSeq* r, const void* const* keys_h, const void* const* pv_h) {
  using namespace ap8;
  WFwdParams p;
  p.T = r->T; p.B = r->B;
  for (int k = 0; k < NM; ++k) {
    const AvsrAttnMech& m = r->mech[k];
    p.Tm[k] = m.Tm; p.scaled[k] = m.kind == AVSR_ATTN_SCALED_LUONG; p.mem_len[k] = m.mem_len; p.Wa[k] = m.Wl;
    p.keys[k] = reinterpret_cast<const __half*>(keys_h[k]); p.pv[k] = reinterpret_cast<const __half*>(pv_h[k]);
    p.g[k] = m.g; p.hc[k] = m.hc; p.ldhc[k] = r->H + m.Dm; p.align[k] = m.align;
    AVSR_REQUIRE(!p.scaled[k] || m.g, "rnn: scaled Luong mechanism without attention_g");
  }
  p.len = r->len; p.gates = r->gates; p.Wrec = r->Wrec; p.c0 = r->c0; p.S = r->S; p.craw = r->craw; p.out = r->out;
  p.cT = r->cT; p.hT = r->hT;
  p.d = drop_cfg8(r);
  p.Wd = p.bd = p.emb = p.Wx = p.bias = nullptr;
  p.used_ids = p.sample_ids = nullptr; p.x = nullptr; p.V = p.E = 0; p.ss_stream = p.thr_p = 0u;
  if (r->samp) {
    const AvsrSampling& s = *r->samp;
    AVSR_REQUIRE(r->rng != nullptr, "rnn: scheduled sampling needs the generator words (rng)");
    AVSR_REQUIRE(s.V > 0 && s.V <= S8_VP && s.E > 0 && s.E <= S8_EMAX, "rnn: sampling alphabet / embedding too wide");
    p.Wd = s.Wd; p.bd = s.bd; p.emb = s.embedding; p.Wx = s.Wx; p.bias = s.bias;
    p.used_ids = s.used_ids; p.sample_ids = s.sample_ids; p.x = s.x; p.V = s.V; p.E = s.E;
    p.ss_stream = s.stream; p.thr_p = s.thr_p;
    return launch_cluster8(st, wlas_persist8_fwd_kernel<true>, r->B, WFWD_SMEM, p, AVSR_K_ATTN_FWD);
  }
  return launch_cluster8(st, wlas_persist8_fwd_kernel<false>, r->B, WFWD_SMEM, p, AVSR_K_ATTN_FWD);
}

int wlas_persist8_launch_bwd(cudaStream_t st, const AvsrRnnSeq* r, const void* const* keys_h, const void* const* pv_h) {
  using namespace ap8;
  WBwdParams p;
  p.T = r->T; p.B = r->B;
  p.grad_scale = r->grad_scale > 0.0f ? r->grad_scale : 1.0f;
  p.inv_grad_scale = 1.0f / p.grad_scale;
  for (int k = 0; k < NM; ++k) {
    const AvsrAttnMech& m = r->mech[k];
    p.Tm[k] = m.Tm; p.scaled[k] = m.kind == AVSR_ATTN_SCALED_LUONG; p.mem_len[k] = m.mem_len; p.Wa[k] = m.Wl;
    p.keys[k] = reinterpret_cast<const __half*>(keys_h[k]); p.pv[k] = reinterpret_cast<const __half*>(pv_h[k]);
    p.g[k] = m.g; p.align[k] = m.align; p.ds[k] = m.ds; p.dg[k] = m.dg;
    AVSR_REQUIRE(m.ds != nullptr, "rnn bwd: mechanism scratch ds missing");
  }
  p.len = r->len; p.gates = r->gates; p.craw = r->craw; p.c0 = r->c0; p.Wrec = r->Wrec;
  p.dout = r->dout; p.dcT = r->dcT; p.dhT = r->dhT; p.dZ = r->dZ; p.dA = r->dA; p.dc0 = r->dc0; p.dh0 = r->dh0; p.dbias = r->dbias;
  p.d = drop_cfg8(r);
  AVSR_REQUIRE(r->dA != nullptr, "rnn bwd: dA scratch missing");
  return launch_cluster8(st, wlas_persist8_bwd_kernel, r->B, WBWD_SMEM, p, AVSR_K_ATTN_BWD);
}

}  // namespace avsr
