// Persistent fused AttentionWrapper(LSTMCell) layer on clusters of FOUR CTAs (forward and backward; Luong /
// scaled-Luong scorer, one mechanism): the AV-Align cross-modal audio layer (reference encoder.py:265-290) and
// the LAS / AV-Align decoder (decoder_unimodal.py:299-352).  Same algebra as attn_persist.cu (fused recurrent
// matrix W' over the operand [h | ctx], attention folded out of the recurrence) - see the header there.
//
// Why four: a B200 keeps only 15 clusters of 8 CTAs resident (GPC granularity) but 33 clusters of 4, so a batch
// of 256 utterances runs as ONE wave of 32 clusters x 8 utterances on 128 SMs, and every CTA owns the attention of
// only 2 utterances.  The memory sweeps (scores / context, d(align) / dq) are bound by what one SM can pull out of
// L2 and convert from fp16, so halving the utterances per SM is what shortens the step.
//
// A CTA now owns 64 hidden units = 256 gate rows = two 128-row M tiles of the per-step product, but only one
// tile of W' (128 x 512 fp16 = 128 KB) fits in shared memory.  The second tile lives in TENSOR MEMORY and enters
// tcgen05.mma as the A operand from TMEM (lane = row, 32-bit column c = K elements 2c, 2c+1; checked by
// tools/micro/ts_mma_test.cu).  The backward kernel holds four 128 x 256 tiles of W'^T the same way: two in shared
// memory, two in tensor memory.
//
// There is no separate MMA-issue warp: lane 0 of warp 0 issues the products at the points of the step where it
// waits for the same barriers anyway (8 warps, 2 per SM sub-partition).
#include "ap4_common.cuh"

namespace avsr {
namespace ap4 {


// =====================================================================================================
// forward
// =====================================================================================================
struct Params {
  int T, B, Tm;
  int scaled;            // scaled_luong: score *= g
  int out_h;             // 1: `out` receives the cell output h; 0: out is filled by the host
  const int* len;        // [B] query lengths
  const int* mem_len;    // [B]
  float* gates;          // [T,B,4H] in: x-projection (+ h0 Wh at t = 0); out: activations
  const float* Wp;       // fused recurrent matrix [(H+DM), 4H] fp32
  const __half* keys;    // [Tm,B,H] fp16 copy
  const __half* values;  // [Tm,B,DM] fp16 copy
  const float* g;        // attention_g [1] or null
  const float* c0;       // [B,H] or null
  float* S;              // [(T+1),B,At+H]; S[0] initialised by the caller; this kernel writes the h columns
  int SW, At;            // row width of S and offset of the h columns
  float* craw;           // [T,B,H]
  float* out;            // [T,B,H] (only if out_h)
  float* hc;             // [T,B,H+DM]  [h | ctx], tf32-rounded
  float* align;          // [T,B,Tm]
  float* cT;             // [B,H] or null
  float* hT;             // [B,H] or null
};

constexpr size_t FWD_SMEM = (size_t)W_BYTES + 2 * OP_BYTES + 4 * NB * UPC * 4 + NU * MAX_TM * 4 + NU * 8 * 4 + 64 + 1024;
static_assert(NU * 4 * DM == 4 * NB * UPC, "partial-context scratch aliases the activation buffer");

__global__ void __launch_bounds__(THREADS, 1) attn_lstm_persist4_fwd_kernel(const Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sW = base;                          // W' tile 0: gate rows of the CTA's units 0..31
  const uint32_t sOp = sW + W_BYTES;                 // two operand buffers [h | ctx], NP rows (rows >= NB stay zero)
  const uint32_t sAct = sOp + 2 * OP_BYTES;          // [4][NB][UPC] floats; also [NU][4][DM] partial contexts
  const uint32_t sSc = sAct + 4 * NB * UPC * 4;      // [NU][MAX_TM] scores / alignments
  const uint32_t sRed = sSc + NU * MAX_TM * 4;       // [NU][8] reduction scratch
  const uint32_t sBar = sRed + NU * 8 * 4;           // [0] mma_done [1,2] h_full[buf] [3,4] ctx_full[buf]
  const uint32_t sTmem = sBar + 40;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  float* act = reinterpret_cast<float*>(gen + (sAct - base));
  float* sc_all = reinterpret_cast<float*>(gen + (sSc - base));
  float* part_all = act;
  float* red_all = reinterpret_cast<float*>(gen + (sRed - base));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int b0 = cluster_id_x() * NB;
  const int T = p.T, B = p.B, Tm = p.Tm;

  if (tid == 0) {
    mbar_init(sBar, THREADS / 32);  // mma_done: one commit per issuing warp
    for (int i = 1; i < 5; ++i) mbar_init(sBar + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // tensor memory (all 512 columns): [0, 128) gate accumulators: tile m, K slice j at column 16 (4 m + j);
  // [256, 512) W' tile 1
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(sTmem) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // W' tile 0 -> shared memory as fp16: row r = gate*32 + u  <->  Wp[k][gate*H + 64*rank + u], u < 32
  for (int seg = warp; seg < KTOT * 4; seg += THREADS / 32) {
    const int k = seg >> 2, g = seg & 3;
    const float w = p.Wp[(size_t)k * 4 * H + g * H + UPC * rank + lane];
    *reinterpret_cast<__half*>(gen + (sW - base) + sw128h_off(128, g * 32 + lane, k)) = __float2half_rn(w);
  }
  // operand buffers start as zeros (no product is issued at t = 0; the padding rows stay zero)
  for (int i = tid; i < 2 * OP_BYTES / 4; i += THREADS) reinterpret_cast<uint32_t*>(gen + (sOp - base))[i] = 0u;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(sTmem));
  const uint32_t tW1 = tmem_base + 256;
  {
    // W' tile 1 -> tensor memory: lane r = gate*32 + u <-> unit 32 + u; column c holds K elements 2c, 2c+1.
    // warp w fills lane quarter (w & 3) = gate, columns 128*(w >> 2) .. +127
    const int q = warp & 3, hh = warp >> 2;
    const float* col = p.Wp + q * H + UPC * rank + 32 + lane;
#pragma unroll 1
    for (int c0 = 0; c0 < 128; c0 += 32) {
      uint32_t r[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        const int k = 2 * (128 * hh + c0 + c);
        r[c] = pack_h2(col[(size_t)k * 4 * H], col[(size_t)(k + 1) * 4 * H]);
      }
      tmem_st32(tW1 + 128 * hh + c0 + ((uint32_t)(32 * q) << 16), r);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  cluster_sync_all();

  // gate-math role: warp <-> (gate g, tile m): the gate rows of units 32*m + lane for all NB utterances
  const int g = warp & 3, m = warp >> 2;
  // product-issue role.  Issuing a tcgen05.mma costs ~80 clocks of one thread (descriptor -> uniform registers), far
  // more than a 128 x 16 x 16 product takes, and the issue of the ctx half sits on the critical path of the step.  So
  // EVERY warp issues: lane 0 of warp (m, j) issues K blocks j (h half) and 4 + j (ctx half) of tile m into its own
  // accumulator columns; the gate math adds the four partial accumulators of its tile.
  const int jq = warp & 3;
  const uint32_t acc_col = tmem_base + (4 * m + jq) * NP;
  const uint64_t dW0 = make_desc_k128(sW);
  const uint64_t dOp[2] = {make_desc_k128(sOp), make_desc_k128(sOp + OP_BYTES)};
  auto issue_block = [&](int kb, uint32_t nbuf, bool clears) {
    if (lane == 0) {
#pragma unroll
      for (int k4 = 0; k4 < 4; ++k4) {
        const uint64_t db = desc_at(dOp[nbuf], kb * (NP * 128) + k4 * 32);
        const uint32_t acc = (clears && k4 == 0) ? 0u : 1u;
        if (m == 0) umma_ss(acc_col, desc_at(dW0, kb * (128 * 128) + k4 * 32), db, IDESC, acc);
        else umma_ts(acc_col, tW1 + (kb * 4 + k4) * 8, db, IDESC, acc);
      }
    }
    __syncwarp();
  };
  const int unit_g = UPC * rank + 32 * m + lane;
  // combine role (threads 0..127): utterance bq, units 4*uq .. 4*uq+3 of the CTA
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
      h_state[e] = (b < B) ? p.S[(size_t)b * p.SW + p.At + u] : 0.0f;
    }
  }
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

  for (int t = 0; t < T; ++t) {
    float* grow = p.gates + ((size_t)t * B + b0) * 4 * H + g * H + unit_g;
    uint32_t r[8];
    if (t > 0) {
      mbar_wait(sBar, (t - 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t r1[8], r2[8], r3[8];
      tmem_ld8(tmem_base + ((uint32_t)(32 * g) << 16) + (4 * m + 0) * NP, r);
      tmem_ld8(tmem_base + ((uint32_t)(32 * g) << 16) + (4 * m + 1) * NP, r1);
      tmem_ld8(tmem_base + ((uint32_t)(32 * g) << 16) + (4 * m + 2) * NP, r2);
      tmem_ld8(tmem_base + ((uint32_t)(32 * g) << 16) + (4 * m + 3) * NP, r3);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
#pragma unroll
      for (int b = 0; b < 8; ++b)
        r[b] = __float_as_uint((__uint_as_float(r[b]) + __uint_as_float(r1[b])) + (__uint_as_float(r2[b]) + __uint_as_float(r3[b])));
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
    const uint32_t nb = (t + 1) & 1;
    const uint32_t hbar_n = sBar + 8 + 8 * nb, cbar_n = sBar + 24 + 8 * nb;
    float hv[4], ov[4], cr[4];
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
          ov[e] = h;
          h_state[e] = tf32_rn(h);
        }
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          cr[e] = c_state[e];
          ov[e] = 0.0f;
        }
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) hv[e] = h_state[e];
      // all-gather of h_t (fp16): operand of step t+1 and query of this step's attention
      const uint32_t off = sw128h_off(NP, bq, UPC * (int)rank + 4 * uq);
      const uint32_t u01 = pack_h2(hv[0], hv[1]), u23 = pack_h2(hv[2], hv[3]);
      const uint32_t dbuf = sOp + nb * OP_BYTES + off;
#pragma unroll
      for (uint32_t dst = 0; dst < (uint32_t)CL; ++dst) st_async_v2(mapa(dbuf, dst), mapa(hbar_n, dst), u01, u23);
    }
    // HBM side of this step + x-projection of the next (overlaps the all-gather)
#pragma unroll
    for (int b = 0; b < NB; ++b)
      if (t < len_a[b]) grow[(size_t)b * 4 * H] = av[b];
    if (comb && b0 + bq < B) {
      const size_t row = (size_t)t * B + b0 + bq;
      const int u0 = UPC * rank + 4 * uq;
      *reinterpret_cast<float4*>(p.craw + row * H + u0) = make_float4(cr[0], cr[1], cr[2], cr[3]);
      if (p.out_h) *reinterpret_cast<float4*>(p.out + row * H + u0) = make_float4(ov[0], ov[1], ov[2], ov[3]);
      *reinterpret_cast<float4*>(p.S + (row + B) * p.SW + p.At + u0) = make_float4(hv[0], hv[1], hv[2], hv[3]);
      *reinterpret_cast<float4*>(p.hc + row * (H + DM) + u0) = make_float4(hv[0], hv[1], hv[2], hv[3]);
    }
    if (t + 1 < T) {
      const float* gnext = grow + (size_t)B * 4 * H;
#pragma unroll
      for (int b = 0; b < NB; ++b) gx[b] = (t + 1 < len_a[b]) ? gnext[(size_t)b * 4 * H] : 0.0f;
    }
    // ---------------- attention of utterance b_att with query h_t ----------------
    const bool live_q = t < len_q;    // masked steps (and padding utterances) skip the memory sweep
    // software-pipelined sweeps (ap4_common.cuh att_fwd_core); the first keys are requested before h_t has landed
    uint4 ra[4], rb[4];
    if (live_q) att_prefetch(role, p.keys, ra, rb);
    if (tid == 0) mbar_expect_tx(hbar_n, NB * H * 2);
    mbar_wait(hbar_n, (t >> 1) & 1);  // every CTA's h_t slice has landed in buffer nb

    float ctxv[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) ctxv[e] = 0.0f;
    if (live_q) {
      // query: lane holds dims 8*lane .. 8*lane+7 (one swizzled 16-byte chunk of the operand row)
      const uint4 qraw = *reinterpret_cast<const uint4*>(gen + (sOp - base) + nb * OP_BYTES + sw128h_off(NP, bl_att, 8 * lane));
      float q[8];
      const float vnone[8] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
      unpack_q(qraw, q);
      att_fwd_core(role, q, vnone, ra, rb, p.align + ((size_t)t * B + b_att) * Tm, ctxv);
    }
    if (w4 == 0) {
      // ctx_t of this utterance: HBM (tf32-rounded fp32, for the backward pass) + all-gather (fp16 operand)
      if (b_att < B) {
        float* dst = p.hc + ((size_t)t * B + b_att) * (H + DM) + H + 8 * lane;
        *reinterpret_cast<float4*>(dst) = make_float4(ctxv[0], ctxv[1], ctxv[2], ctxv[3]);
        *reinterpret_cast<float4*>(dst + 4) = make_float4(ctxv[4], ctxv[5], ctxv[6], ctxv[7]);
        if (!live_q) {
          float* arow = p.align + ((size_t)t * B + b_att) * Tm;
          for (int tm = lane; tm < Tm; tm += 32) arow[tm] = 0.0f;
        }
      }
      const uint32_t off = sw128h_off(NP, bl_att, H + 8 * lane);
      const uint32_t dbuf = sOp + nb * OP_BYTES + off;
      const uint32_t c0 = pack_h2(ctxv[0], ctxv[1]), c1 = pack_h2(ctxv[2], ctxv[3]);
      const uint32_t c2 = pack_h2(ctxv[4], ctxv[5]), c3 = pack_h2(ctxv[6], ctxv[7]);
#pragma unroll
      for (uint32_t dst = 0; dst < (uint32_t)CL; ++dst) st_async_v4(mapa(dbuf, dst), mapa(cbar_n, dst), c0, c1, c2, c3);
    }
    {
      // Gate products of step t+1.  The h half is issued here, after this warp's share of the attention and while the
      // contexts of the other CTAs are still in flight (h_t has landed: every warp waited for it above; every warp
      // has read the accumulators of step t before its activations reached the barrier that precedes the h
      // all-gather); the ctx half once every context of this step has landed.  After the last step the wait only
      // drains the all-gathers: no st.async may be in flight towards this CTA when it exits.
      if (t + 1 < T) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        issue_block(jq, nb, true);
      }
      if (tid == 0) mbar_expect_tx(cbar_n, NB * DM * 2);
      mbar_wait(cbar_n, (t >> 1) & 1);
      if (t + 1 < T) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        issue_block(4 + jq, nb, false);
        if (lane == 0) umma_commit(sBar);
        __syncwarp();
      }
    }
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
// backward:  [dh_{t-1} | dctx_{t-1}] = dz_t W'^T + (dout_{t-1} Wl^T)
// A CTA owns the gate columns of its 64 units (K split of the product over the cluster): partial [h | ctx] rows are
// reduce-scattered through DSMEM (h rows to the owners of the units, ctx rows to the owners of the utterances);
// attention backward of its 2 utterances; dq all-to-all.  dz enters the tensor core as fp16 scaled by a power of two.
// =====================================================================================================
struct BwdParams {
  int T, B, Tm, scaled;
  float grad_scale, inv_grad_scale;
  const int* len;
  const int* mem_len;
  const float* gates;    // [T,B,4H] activations
  const float* craw;     // [T,B,H]
  const float* c0;       // [B,H] or null
  const float* Wp;       // fused recurrent matrix [(H+DM),4H]
  const __half* keys;    // [Tm,B,H]
  const __half* values;  // [Tm,B,DM]
  const float* g;        // [1] or null
  const float* hc;       // [T,B,H+DM] forward [h | ctx]
  const float* align;    // [T,B,Tm]
  const float* douthc;   // [T,B,H+DM] = dout Wl^T (unmasked) or null
  const float* dcT;      // [B,H] or null
  const float* dhT;      // [B,H] or null
  float* dZ;             // [T,B,4H]
  float* ds;             // [T,B,Tm]
  float* dhc;            // [T,B,H+DM]: the ctx columns receive dctx_t
  float* dg;             // [1] or null
  float* dc0;            // [B,H] or null
  float* dh0;            // [B,H] or null
  float* dbias;          // [4H] or null: += column sums of dZ
  long long* dbg;        // AVSR_AP_DEBUG: clock samples [64 iterations][12] of CTA 0, thread 0
};
#define AP4B_STAMP(slot)                                                                         \
  do {                                                                                           \
    if (p.dbg && blockIdx.x == 0 && tid == 0 && it < 64) p.dbg[it * 12 + (slot)] = clock64(); \
  } while (0)

constexpr int BW_TILE_BYTES = 4 * 128 * 128;         // one 128-row tile of W'^T restricted to the CTA's 256 gate columns
constexpr int BW_DZ_BYTES = 4 * NP * 128;            // B operand: 4 K-blocks (gates) x [NP rows x 64 units]
constexpr int REDH_FLOATS = CL * NB * UPC;           // [src][b][u]
constexpr int REDC_FLOATS = CL * NU * DM;            // [src][utt][dim]; also the dq partial scratch [NU][4][DM]
constexpr int DQ_FLOATS = CL * NU * UPC;             // [src][utt][u]
constexpr size_t BWD_SMEM = (size_t)2 * BW_TILE_BYTES + BW_DZ_BYTES + REDH_FLOATS * 4 + REDC_FLOATS * 4 + DQ_FLOATS * 4 +
                            NU * DM * 4 + 2 * NU * MAX_TM * 4 + NU * 8 * 4 + 64 + 1024;
static_assert(NU * 4 * DM <= REDC_FLOATS, "dq partial scratch must fit the ctx reduce buffer");

template <bool SMALL>
__global__ void __launch_bounds__(THREADS, 1) attn_lstm_persist4_bwd_kernel(const BwdParams p) {
  constexpr int MAXB = SMALL ? SMALL_B : 1;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sW = base;                                  // tiles 0, 1 (h rows) of W'^T
  const uint32_t sDz = sW + 2 * BW_TILE_BYTES;
  const uint32_t sRedH = sDz + BW_DZ_BYTES;
  const uint32_t sRedC = sRedH + REDH_FLOATS * 4;
  const uint32_t sDq = sRedC + REDC_FLOATS * 4;
  const uint32_t sCtx = sDq + DQ_FLOATS * 4;                 // [NU][DM] dctx of the CTA's utterances
  const uint32_t sSc = sCtx + NU * DM * 4;                   // [NU][MAX_TM] alignments (long memories only)
  const uint32_t sDs = sSc + NU * MAX_TM * 4;                // [NU][MAX_TM] d(align) / ds (long memories only)
  const uint32_t sRed = sDs + NU * MAX_TM * 4;               // [NU][8]
  const uint32_t sBar = sRed + NU * 8 * 4;  // [0] mma_done [1] dz_ready [2] redH_full [3] redC_full [4] dq_full
  const uint32_t sTmem = sBar + 40;
  const uint32_t barMma = sBar, barDz = sBar + 8, barRedH = sBar + 16, barRedC = sBar + 24, barDq = sBar + 32;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  float* redH = reinterpret_cast<float*>(gen + (sRedH - base));
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
    mbar_init(barMma, THREADS / 32);  // one commit per issuing warp
    mbar_init(barDz, THREADS);
    mbar_init(barRedH, 1);
    mbar_init(barRedC, 1);
    mbar_init(barDq, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // tensor memory (all 512 columns): [0, 128) accumulators: 128-row tile mt, K half hj at column 16 (2 mt + hj);
  // [256, 512) tiles 2, 3 (ctx dims) of W'^T as A operands (128 columns = 256 K each)
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(sTmem) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // A[n][k = g*64 + u] = Wp[n][g*H + 64*rank + u]; rows n < 256 -> shared memory (tile n >> 7)
  for (int seg = warp; seg < 256 * 8; seg += THREADS / 32) {
    const int n = seg >> 3, g = (seg >> 1) & 3, u = 32 * (seg & 1) + lane;
    const float w = p.Wp[(size_t)n * 4 * H + g * H + UPC * rank + u];
    *reinterpret_cast<__half*>(gen + (sW - base) + (n >> 7) * BW_TILE_BYTES + sw128h_off(128, n & 127, g * 64 + u)) =
        __float2half_rn(w);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(sTmem));
  const uint32_t tA = tmem_base + 256;
  {
    // rows n = 256 + 128*tt + 32*q + lane -> tensor memory tile tt; column c holds k = 2c, 2c+1 (adjacent units)
    const int q = warp & 3, tt = warp >> 2;
    const float* row = p.Wp + (size_t)(256 + 128 * tt + 32 * q + lane) * 4 * H + UPC * rank;
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
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  cluster_sync_all();

  const uint64_t dWt = make_desc_k128(sW), dDz = make_desc_k128(sDz);
  // gate-gradient role: thread = (local unit ul, utterances 2*(warp >> 1) + j)
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
        dov[j] = p.douthc ? p.douthc[((size_t)t * B + b) * (H + DM) + unit] : 0.0f;
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
  // reduce-scatter role after the product: warps 0-3 forward tiles 0, 1 (h rows), warps 4-7 tiles 2, 3 (ctx dims)
  const int q = warp & 3;

  float bsum[4] = {0.0f, 0.0f, 0.0f, 0.0f};  // bias gradient of this thread's unit: sum of dz over its utterances / steps
  load_step(T - 1);
  for (int it = 0; it < T; ++it) {
    const int t = T - 1 - it;
    const bool live_q = t < len_q;
    // ---- top: partial sums pushed during the previous iteration ------------------------------------------
    float dh_in[PB];
#pragma unroll
    for (int j = 0; j < PB; ++j) dh_in[j] = dh_carry[j];
    AP4B_STAMP(0);
    uint4 ra[4], rb[4];  // software-pipelined sweeps: the first values are requested before the partials arrive
    float dctx_pre[DM / 128];  // this step's dout part of dctx, likewise
    float al[MAXB];            // and its alignments: al[i] = a[w4 + 32*i + 4*jrow], the row whose batch total this lane gets
    const int jrow = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
#pragma unroll
    for (int i = 0; i < MAXB; ++i) al[i] = 0.0f;
    if (live_q) {
      att_prefetch(role, p.values, ra, rb);
#pragma unroll
      for (int i = 0; i < DM / 128; ++i)
        dctx_pre[i] = p.douthc ? p.douthc[((size_t)t * B + b_att) * (H + DM) + H + gt + 128 * i] : 0.0f;
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
    if (it > 0) {
      if (tid == 0) {
        mbar_expect_tx(barRedC, REDC_FLOATS * 4);
        mbar_expect_tx(barRedH, REDH_FLOATS * 4);
      }
      mbar_wait(barRedH, (it - 1) & 1);
#pragma unroll
      for (int j = 0; j < PB; ++j) {
        const int bl = (warp >> 1) * PB + j;
#pragma unroll
        for (int src = 0; src < CL; ++src) dh_in[j] += redH[(src * NB + bl) * UPC + ul];
      }
      mbar_wait(barRedC, (it - 1) & 1);
    }
    AP4B_STAMP(1);
    if (live_q) {
#pragma unroll
      for (int i = 0; i < DM / 128; ++i) {
        const int d = gt + 128 * i;
        float v = dctx_pre[i];
        if (it > 0) {
#pragma unroll
          for (int src = 0; src < CL; ++src) v += redC[(src * NU + jl) * DM + d];
        }
        dctx_s[d] = v;
        p.dhc[((size_t)t * B + b_att) * (H + DM) + H + d] = v;
      }
    }
    // from here on redH and redC are free again: a peer can only push the next partials after it has received this
    // CTA's dq, which leaves after this barrier (so one copy of each suffices, and redC doubles as dq scratch)
    asm volatile("bar.sync 1, 256;" ::: "memory");
    AP4B_STAMP(2);
    // ---- (A/B) attention backward of the CTA's utterances, dq all-to-all ---------------------------------
    float dqv[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) dqv[e] = 0.0f;
    float ds_keep[MAXB];  // d(score) of the rows this lane owns (written to HBM off the critical path)
#pragma unroll
    for (int i = 0; i < MAXB; ++i) ds_keep[i] = 0.0f;
    if (live_q)
      att_bwd_core<SMALL>(role, dctx_s, ra, rb, al, ds_keep, a_s, ds_s, p.ds + ((size_t)t * B + b_att) * Tm, p.scaled != 0, p.dg, dqv,
                          vzero8, vzero8);
    AP4B_STAMP(3);
    if (w4 == 0) {
      // dq dims 8*lane .. +7 belong to the CTA owning units (8*lane)/64
      const uint32_t dst = (uint32_t)(lane >> 3);
      const uint32_t a0 = mapa(sDq + (uint32_t)(((rank * NU + jl) * UPC + ((8 * lane) & (UPC - 1))) * 4), dst);
      const uint32_t bar = mapa(barDq, dst);
      st_async_v4f(a0, bar, dqv[0], dqv[1], dqv[2], dqv[3]);
      st_async_v4f(a0 + 16, bar, dqv[4], dqv[5], dqv[6], dqv[7]);
    }
    // ---- (C/D) dq of this CTA's units -> gate gradients --------------------------------------------------
    AP4B_STAMP(4);
    if (tid == 0) mbar_expect_tx(barDq, DQ_FLOATS * 4);
    mbar_wait(barDq, it & 1);
    AP4B_STAMP(5);
    float dz[4][PB];
#pragma unroll
    for (int j = 0; j < PB; ++j) {
      const int bl = (warp >> 1) * PB + j;
      float dh = dh_in[j];
      if (t < len_t[j]) {
        dh += dov[j] + dqb[bl * UPC + ul];
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
        dh_carry[j] = dh;
      }
#pragma unroll
      for (int g = 0; g < 4; ++g)
        *reinterpret_cast<__half*>(gen + (sDz - base) + sw128h_off(NP, bl, g * 64 + ul)) = __float2half_rn(dz[g][j] * p.grad_scale);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_arrive(barDz);
    {
      // partial [h | ctx](512) x NB from this CTA's 256 gate columns, once every warp's dz is in shared memory.  Every
      // warp issues (a tcgen05.mma costs ~80 clocks of its issuing thread): lane 0 of warp (mt, hj) the K half hj
      // (gates 2 hj, 2 hj + 1) of the 128-row tile mt into its own accumulator columns; the reduce-scatter adds the two.
      const int mt = warp >> 1, hj = warp & 1;
      mbar_wait(barDz, it & 1);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (lane == 0) {
#pragma unroll
        for (int kk = 0; kk < 2; ++kk)
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            const int kb = 2 * hj + kk;
            const uint64_t db = desc_at(dDz, kb * (NP * 128) + k4 * 32);
            if (mt < 2)
              umma_ss(tmem_base + (2 * mt + hj) * NP, desc_at(dWt, mt * BW_TILE_BYTES + kb * (128 * 128) + k4 * 32), db, IDESC,
                      (kk | k4) ? 1u : 0u);
            else
              umma_ts(tmem_base + (2 * mt + hj) * NP, tA + (mt - 2) * 128 + (kb * 4 + k4) * 8, db, IDESC, (kk | k4) ? 1u : 0u);
          }
        umma_commit(barMma);
      }
      __syncwarp();
    }
    AP4B_STAMP(6);
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
      // d(score) of this step (dkeys is formed after the loop) and d(attention_g), off the critical path
      if (live_q) att_bwd_small_tail(role, p.ds + ((size_t)t * B + b_att) * Tm, al, ds_keep, p.scaled != 0, p.dg);
    }
    AP4B_STAMP(7);
    // ---- (E) partial products -> owners -------------------------------------------------------------------
    mbar_wait(barMma, it & 1);
    AP4B_STAMP(8);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
    for (int mi = 0; mi < 2; ++mi) {
      const int mt = 2 * (warp >> 2) + mi;
      if (mt < 2) {  // h rows: global unit 128*mt + 32*q + lane -> owner CTA 2*mt + (q >> 1), local unit 32*(q & 1) + lane
        uint32_t r[8], r1[8];
        tmem_ld8(tmem_base + ((uint32_t)(32 * q) << 16) + (2 * mt) * NP, r);
        tmem_ld8(tmem_base + ((uint32_t)(32 * q) << 16) + (2 * mt + 1) * NP, r1);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int c = 0; c < NB; ++c) r[c] = __float_as_uint(__uint_as_float(r[c]) + __uint_as_float(r1[c]));
        const uint32_t dst = (uint32_t)(2 * mt + (q >> 1));
        const uint32_t a0 = mapa(sRedH + (uint32_t)((rank * NB) * UPC + 32 * (q & 1) + lane) * 4, dst);
        const uint32_t bar = mapa(barRedH, dst);
#pragma unroll
        for (int c = 0; c < NB; ++c) st_async_f(a0 + c * UPC * 4, bar, __uint_as_float(r[c]) * p.inv_grad_scale);
      } else if (it + 1 < T) {  // ctx dims 128*(mt-2) + 32*q + lane; column c = utterance -> owner CTA c / NU
        uint32_t r[8], r1[8];
        tmem_ld8(tmem_base + ((uint32_t)(32 * q) << 16) + (2 * mt) * NP, r);
        tmem_ld8(tmem_base + ((uint32_t)(32 * q) << 16) + (2 * mt + 1) * NP, r1);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int c = 0; c < NB; ++c) r[c] = __float_as_uint(__uint_as_float(r[c]) + __uint_as_float(r1[c]));
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
    AP4B_STAMP(9);
  }
  // drain the last reduce-scatter (nothing may be in flight towards this CTA when it exits).  Its content is NOT
  // dh_0: step 0 saw att_{-1} = 0, so dh_0 = dz_0 Wh^T with the un-fused Wh - added by the host; here only the
  // gradient carried through fully masked utterances is written.
  if (T > 0) {
    if (tid == 0) mbar_expect_tx(barRedH, REDH_FLOATS * 4);
    mbar_wait(barRedH, (T - 1) & 1);
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

// AVSR_AP_DEBUG (never during graph capture): average clocks CTA 0 / thread 0 spends in each phase of a step
static int print_dbg(cudaStream_t st, long long* dbg, const char* tag, int T, int Tm, int ns, const char* const* names) {
  AVSR_CHECK_CUDA(cudaStreamSynchronize(st));
  long long h[64 * 12];
  AVSR_CHECK_CUDA(cudaMemcpy(h, dbg, sizeof(h), cudaMemcpyDeviceToHost));
  cudaFree(dbg);
  const int n = T < 64 ? T : 64;
  double acc[12] = {0};
  for (int t = 3; t < n; ++t) {
    for (int k = 1; k < ns; ++k) acc[k] += (double)(h[t * 12 + k] - h[t * 12 + k - 1]);
    acc[0] += (double)(h[t * 12] - h[(t - 1) * 12 + ns - 1]);
  }
  fprintf(stderr, "%s T=%d Tm=%d clocks/step:", tag, T, Tm);
  double tot = 0;
  for (int k = 0; k < ns; ++k) {
    fprintf(stderr, " %s=%.0f", names[k], acc[k] / (n - 3));
    tot += acc[k] / (n - 3);
  }
  fprintf(stderr, " total=%.0f\n", tot);
  return 0;
}

}  // namespace ap4

// launched by attn_persist_fwd / attn_persist_bwd (attn_persist.cu), which prepare the fused matrix, the fp16 copies
// of keys / values and the batched products around the recurrent kernel
int attn_persist4_launch_fwd(cudaStream_t st, int T, int B, int Tm, int scaled, const int* len, const int* mem_len,
                             float* gates, const float* Wp, const void* keys_h, const void* values_h, const float* g,
                             const float* c0, float* S, int SW, int At, float* craw, float* out, float* hc, float* align,
                             float* cT, float* hT) {
  ap4::Params p;
  p.T = T; p.B = B; p.Tm = Tm; p.scaled = scaled; p.out_h = 0;
  p.len = len; p.mem_len = mem_len; p.gates = gates; p.Wp = Wp;
  p.keys = reinterpret_cast<const __half*>(keys_h); p.values = reinterpret_cast<const __half*>(values_h);
  p.g = g; p.c0 = c0; p.S = S; p.SW = SW; p.At = At; p.craw = craw; p.out = out; p.hc = hc; p.align = align;
  p.cT = cT; p.hT = hT;
  return ap4::launch_cluster(st, ap4::attn_lstm_persist4_fwd_kernel, B, ap4::FWD_SMEM, p, AVSR_K_ATTN_FWD);
}

int attn_persist4_launch_bwd(cudaStream_t st, int T, int B, int Tm, int scaled, float grad_scale, const int* len,
                             const int* mem_len, const float* gates, const float* craw, const float* c0, const float* Wp,
                             const void* keys_h, const void* values_h, const float* g, const float* hc, const float* align,
                             const float* douthc, const float* dcT, const float* dhT, float* dZ, float* ds, float* dhc,
                             float* dg, float* dc0, float* dh0, float* dbias) {
  ap4::BwdParams p;
  p.T = T; p.B = B; p.Tm = Tm; p.scaled = scaled;
  p.grad_scale = grad_scale; p.inv_grad_scale = 1.0f / grad_scale;
  p.len = len; p.mem_len = mem_len; p.gates = gates; p.craw = craw; p.c0 = c0; p.Wp = Wp;
  p.keys = reinterpret_cast<const __half*>(keys_h); p.values = reinterpret_cast<const __half*>(values_h);
  p.g = g; p.hc = hc; p.align = align; p.douthc = douthc; p.dcT = dcT; p.dhT = dhT; p.dZ = dZ; p.ds = ds; p.dhc = dhc;
  p.dg = dg; p.dc0 = dc0; p.dh0 = dh0; p.dbias = dbias;
  p.dbg = nullptr;
  if (getenv("AVSR_AP_DEBUG")) {
    static const char* names[10] = {"loop-top", "wait redH/redC+fold", "dctx+cta bar", "attention bwd", "reduce+send dq",
                                    "wait dq", "dz->smem+issue", "hbm st/ld", "wait MMA", "tmem ld+push"};
    AVSR_CHECK_CUDA(cudaMalloc(&p.dbg, 64 * 12 * sizeof(long long)));
    AVSR_CHECK_CUDA(cudaMemset(p.dbg, 0, 64 * 12 * sizeof(long long)));
    AVSR_TRY(Tm <= ap4::SMALL_TM ? ap4::launch_cluster(st, ap4::attn_lstm_persist4_bwd_kernel<true>, B, ap4::BWD_SMEM, p, AVSR_K_ATTN_BWD)
                                 : ap4::launch_cluster(st, ap4::attn_lstm_persist4_bwd_kernel<false>, B, ap4::BWD_SMEM, p, AVSR_K_ATTN_BWD));
    return ap4::print_dbg(st, p.dbg, "[ap4 bwd]", T, Tm, 10, names);
  }
  return Tm <= ap4::SMALL_TM ? ap4::launch_cluster(st, ap4::attn_lstm_persist4_bwd_kernel<true>, B, ap4::BWD_SMEM, p, AVSR_K_ATTN_BWD)
                             : ap4::launch_cluster(st, ap4::attn_lstm_persist4_bwd_kernel<false>, B, ap4::BWD_SMEM, p, AVSR_K_ATTN_BWD);
}

}  // namespace avsr
