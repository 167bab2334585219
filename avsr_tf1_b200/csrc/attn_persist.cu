// Persistent fused AttentionWrapper(LSTMCell) layer, forward (Luong / scaled-Luong scorer, one
// mechanism): the AV-Align cross-modal audio layer (reference encoder.py:265-290) and the LAS / AV-Align
// decoder (decoder_unimodal.py:299-352) in ONE launch per sequence.
//
// Same cluster decomposition as lstm_persist.cu: a cluster of 8 CTAs owns 16 utterances for all steps;
// CTA `rank` owns the gate columns of 32 hidden units and the attention of 2 utterances.
//
// The attention vector is folded out of the recurrence:  with att_{t-1} = [h_{t-1} | ctx_{t-1}] Wl,
//     z_t = x_t Wx + att_{t-1} Wa + h_{t-1} Wh  =  x_t Wx + h_{t-1} (Wh + Wl_h Wa) + ctx_{t-1} (Wl_c Wa)
// so the recurrent operand is [h | ctx] (K = H + Dm) against the fused matrix W' (built per call by two
// small products, kept resident in shared memory as fp16 - tf32-rounded values are exactly representable
// in fp16, so operands equal the tf32 operands of the rest of the path).  att_t itself (layer output for
// the Luong family, and the operand of the backward pass) is formed AFTER the loop by one batched product.
// Step 0 sees att_{-1} = 0 (AttentionWrapper zero state): its recurrent term h_0 Wh is added to the
// x-projection by the host before the launch and the kernel issues no product at t = 0.
//
// per step:   tcgen05.mma kind::f16: D[128 gate rows, 16] = W'^T[128, H+Dm] . [h|ctx]^T   (h half first:
//             it overlaps the attention of the previous step; ctx half when the contexts have landed)
//             gate math -> h_t -> st.async all-gather (operand of step t+1 and query of step t)
//             attention of the CTA's two utterances: scores = g * keys.h, masked softmax, context
//             (fp16 copies of keys / values streamed from L2, 4 rows in flight per warp)
//             ctx_t -> st.async all-gather
#include <cuda_fp16.h>
#include <stdlib.h>

#include "../../include/avsr_b200.h"
#include "common.cuh"

namespace avsr {
namespace ap {

constexpr int GM_WARPS = 8;
constexpr int THREADS = (GM_WARPS + 1) * 32;
constexpr int NB = 16;
constexpr int CL = 8;
constexpr int H = 256;
constexpr int DM = 256;
constexpr int KTOT = H + DM;            // 512
constexpr int KB = KTOT / 64;           // 8 K-blocks of 64 halves (128 B)
constexpr int W_BYTES = KB * 128 * 128; // 128 KB
constexpr int OP_BYTES = KB * NB * 128; // 16 KB per operand buffer
constexpr int MAX_TM = 384;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa(uint32_t local, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_async_v2(uint32_t addr, uint32_t mbar, uint32_t a, uint32_t b) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b32 [%0], {%2, %3}, [%1];" ::"r"(addr),
               "r"(mbar), "r"(a), "r"(b)
               : "memory");
}
__device__ __forceinline__ void st_async_v4(uint32_t addr, uint32_t mbar, uint32_t a, uint32_t b, uint32_t c,
                                            uint32_t d) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%2, %3, %4, %5}, [%1];" ::"r"(addr),
               "r"(mbar), "r"(a), "r"(b), "r"(c), "r"(d)
               : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "AP_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra AP_DONE;\n\t"
      "bra AP_WAIT;\n\t"
      "AP_DONE:\n\t"
      "}" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ uint64_t make_desc_k128(uint32_t saddr) {  // K-major, SWIZZLE_128B, SBO = 1024 B
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_h2(uint32_t u) { return __half22float2(*reinterpret_cast<__half2*>(&u)); }

// byte offset of half element (row, k) in a K-major SWIZZLE_128B operand with 64-half K blocks of `rows` rows
__device__ __forceinline__ uint32_t sw128h_off(int rows, int row, int k) {
  const int kb = k >> 6, kk = k & 63;
  return (uint32_t)(kb * rows * 128 + row * 128 + ((((kk >> 3) ^ (row & 7)) << 4)) + ((kk & 7) << 1));
}

// instruction descriptor: D = f32, A = B = f16, both K-major, N = NB, M = 128
constexpr uint32_t IDESC = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(NB >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

struct Params {
  int T, B, Tm;
  int scaled;            // scaled_luong: score *= g
  int out_h;             // 1: `out` receives the cell output h (Bahdanau family); 0: out is filled by the host
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

__global__ void __launch_bounds__(THREADS, 1) attn_lstm_persist_fwd_kernel(const Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sW = base;
  const uint32_t sOp = sW + W_BYTES;                 // two operand buffers [h | ctx]
  const uint32_t sAct = sOp + 2 * OP_BYTES;          // [4][NB][32] floats
  const uint32_t sSc = sAct + 4 * NB * 32 * 4;       // [2][MAX_TM] scores / alignments
  const uint32_t sPart = sSc + 2 * MAX_TM * 4;       // [2][4][DM] partial contexts
  const uint32_t sRed = sPart + 2 * 4 * DM * 4;      // [2][8] reduction scratch
  const uint32_t sBar = sRed + 64;                   // [0] mma_done [1,2] h_full[buf] [3,4] ctx_full[buf]
  const uint32_t sTmem = sBar + 40;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  float* act = reinterpret_cast<float*>(gen + (sAct - base));
  float* sc_all = reinterpret_cast<float*>(gen + (sSc - base));
  float* part_all = reinterpret_cast<float*>(gen + (sPart - base));
  float* red_all = reinterpret_cast<float*>(gen + (sRed - base));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int b0 = cluster_id_x() * NB;
  const int T = p.T, B = p.B, Tm = p.Tm;

  if (tid == 0) {
    for (int i = 0; i < 5; ++i) mbar_init(sBar + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == GM_WARPS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(sTmem) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // resident fused weights as fp16: row r = gate*32 + u <-> Wp[k][gate*H + 32*rank + u]
  for (int seg = warp; seg < KTOT * 4; seg += THREADS / 32) {
    const int k = seg >> 2, g = seg & 3;
    const float w = p.Wp[(size_t)k * 4 * H + g * H + 32 * rank + lane];
    *reinterpret_cast<__half*>(gen + (sW - base) + sw128h_off(128, g * 32 + lane, k)) = __float2half_rn(w);
  }
  // operand buffers start as zeros (no product is issued at t = 0)
  for (int i = tid; i < 2 * OP_BYTES / 4; i += THREADS) reinterpret_cast<uint32_t*>(gen + (sOp - base))[i] = 0u;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(sTmem));
  cluster_sync_all();

  if (warp == GM_WARPS) {
    // ================= MMA issuer =================
    // iteration t consumes [h_{t-1} | ctx_{t-1}] from buffer t&1; the extra iteration t = T only drains the
    // final all-gathers so that no st.async is in flight towards this CTA when it exits
    for (int t = 1; t <= T; ++t) {
      const uint32_t ob = sOp + (t & 1) * OP_BYTES;
      const uint32_t par = ((t - 1) >> 1) & 1;
      const uint32_t hbar = sBar + 8 + 8 * (t & 1), cbar = sBar + 24 + 8 * (t & 1);
      if (lane == 0) mbar_expect_tx(hbar, NB * H * 2);
      mbar_wait(hbar, par);
      if (t < T) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (lane == 0) {
#pragma unroll
          for (int kb = 0; kb < H / 64; ++kb)
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4)
              umma_f16(tmem_base, make_desc_k128(sW + kb * (128 * 128) + k4 * 32),
                       make_desc_k128(ob + kb * (NB * 128) + k4 * 32), IDESC, (kb | k4) ? 1u : 0u);
        }
        __syncwarp();
      }
      if (lane == 0) mbar_expect_tx(cbar, NB * DM * 2);
      mbar_wait(cbar, par);
      if (t < T) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (lane == 0) {
#pragma unroll
          for (int kb = H / 64; kb < KB; ++kb)
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4)
              umma_f16(tmem_base, make_desc_k128(sW + kb * (128 * 128) + k4 * 32),
                       make_desc_k128(ob + kb * (NB * 128) + k4 * 32), IDESC, 1u);
          umma_commit(sBar);
        }
        __syncwarp();
      }
    }
  } else {
    // ================= gate math + attention warps =================
    const int g = warp & 3, ch = warp >> 2;
    const int unit = 32 * rank + lane;
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
        const int u = 32 * rank + 4 * uq + e;
        c_state[e] = (b < B && p.c0) ? p.c0[(size_t)b * H + u] : 0.0f;
        h_state[e] = (b < B) ? p.S[(size_t)b * p.SW + p.At + u] : 0.0f;
      }
    }
    int len_a[8];
#pragma unroll
    for (int b = 0; b < 8; ++b) len_a[b] = (b0 + ch * 8 + b < B) ? p.len[b0 + ch * 8 + b] : 0;
    float gx[8];
    {
      const float* grow0 = p.gates + ((size_t)b0 + ch * 8) * 4 * H + g * H + unit;
#pragma unroll
      for (int b = 0; b < 8; ++b) gx[b] = (0 < len_a[b]) ? grow0[(size_t)b * 4 * H] : 0.0f;
    }
    // attention role: utterance jl of this CTA, warp w4 of its group of four
    const int jl = warp >> 2, w4 = warp & 3, gt = tid & 127;
    const int bl_att = 2 * (int)rank + jl;       // row of the utterance in the operand buffers
    const int b_att = b0 + bl_att;
    const int len_q = (b_att < B) ? p.len[b_att] : 0;
    const int L = (b_att < B) ? min(p.mem_len[b_att], Tm) : 0;
    const float gs = p.scaled ? p.g[0] : 1.0f;
    float* sc = sc_all + jl * MAX_TM;
    float* part = part_all + jl * 4 * DM;
    float* red = red_all + jl * 8;
    const uint32_t att_bar_id = 2 + jl;          // named barrier of the 128 threads of this utterance

    for (int t = 0; t < T; ++t) {
      float* grow = p.gates + ((size_t)t * B + b0 + ch * 8) * 4 * H + g * H + unit;
      uint32_t r[8];
      if (t > 0) {
        mbar_wait(sBar, (t - 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        tmem_ld8(tmem_base + ((uint32_t)(32 * g) << 16) + ch * 8, r);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      } else {
#pragma unroll
        for (int b = 0; b < 8; ++b) r[b] = 0u;  // att_{-1} = 0; h_0 Wh is already in the x-projection
      }
      float av[8];
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        const float z = __uint_as_float(r[b]) + gx[b];
        float a;
        if (g == 1) a = tanhf_acc(z);
        else a = sigmoidf_acc(g == 2 ? z + 1.0f : z);
        av[b] = a;
        act[(g * NB + ch * 8 + b) * 32 + lane] = a;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const uint32_t nb = (t + 1) & 1;
      const uint32_t hbar_n = sBar + 8 + 8 * nb, cbar_n = sBar + 24 + 8 * nb;
      float hv[4], ov[4], cr[4];
      if (comb) {
        const bool live = t < len_c;
        if (live) {
          const float4 ai = *reinterpret_cast<const float4*>(&act[(0 * NB + bq) * 32 + 4 * uq]);
          const float4 aj = *reinterpret_cast<const float4*>(&act[(1 * NB + bq) * 32 + 4 * uq]);
          const float4 af = *reinterpret_cast<const float4*>(&act[(2 * NB + bq) * 32 + 4 * uq]);
          const float4 ao = *reinterpret_cast<const float4*>(&act[(3 * NB + bq) * 32 + 4 * uq]);
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
        const int k0 = 32 * (int)rank + 4 * uq;  // first of the 4 units; 8 bytes inside one 16-byte chunk
        const uint32_t off = sw128h_off(NB, bq, k0);
        const uint32_t u01 = pack_h2(hv[0], hv[1]), u23 = pack_h2(hv[2], hv[3]);
        const uint32_t dbuf = sOp + nb * OP_BYTES + off;
#pragma unroll
        for (uint32_t dst = 0; dst < (uint32_t)CL; ++dst) st_async_v2(mapa(dbuf, dst), mapa(hbar_n, dst), u01, u23);
      }
      // HBM side of this step + x-projection of the next (overlaps the all-gather)
#pragma unroll
      for (int b = 0; b < 8; ++b)
        if (t < len_a[b]) grow[(size_t)b * 4 * H] = av[b];
      if (comb && b0 + bq < B) {
        const size_t row = (size_t)t * B + b0 + bq;
        const int u0 = 32 * rank + 4 * uq;
        *reinterpret_cast<float4*>(p.craw + row * H + u0) = make_float4(cr[0], cr[1], cr[2], cr[3]);
        if (p.out_h) *reinterpret_cast<float4*>(p.out + row * H + u0) = make_float4(ov[0], ov[1], ov[2], ov[3]);
        *reinterpret_cast<float4*>(p.S + (row + B) * p.SW + p.At + u0) = make_float4(hv[0], hv[1], hv[2], hv[3]);
        *reinterpret_cast<float4*>(p.hc + row * (H + DM) + u0) = make_float4(hv[0], hv[1], hv[2], hv[3]);
      }
      if (t + 1 < T) {
        const float* gnext = grow + (size_t)B * 4 * H;
#pragma unroll
        for (int b = 0; b < 8; ++b) gx[b] = (t + 1 < len_a[b]) ? gnext[(size_t)b * 4 * H] : 0.0f;
      }
      // ---------------- attention of utterance b_att with query h_t ----------------
      mbar_wait(hbar_n, (t >> 1) & 1);  // every CTA's h_t slice has landed in buffer nb
      const bool live_q = t < len_q;    // masked steps (and padding utterances) skip the memory sweep
      float ctxv[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) ctxv[e] = 0.0f;
      if (live_q) {
        // query: lane holds dims 8*lane .. 8*lane+7 (one swizzled 16-byte chunk of the operand row)
        const uint4 qraw = *reinterpret_cast<const uint4*>(gen + (sOp - base) + nb * OP_BYTES + sw128h_off(NB, bl_att, 8 * lane));
        float q[8];
        {
          float2 a = unpack_h2(qraw.x), b = unpack_h2(qraw.y), c = unpack_h2(qraw.z), d = unpack_h2(qraw.w);
          q[0] = a.x; q[1] = a.y; q[2] = b.x; q[3] = b.y; q[4] = c.x; q[5] = c.y; q[6] = d.x; q[7] = d.y;
        }
        // scores: rows tm = w4 + 4*i, four rows in flight
        for (int tm0 = w4; tm0 < L; tm0 += 16) {
          uint4 k[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int tm = tm0 + 4 * j;
            k[j] = tm < L ? __ldg(reinterpret_cast<const uint4*>(p.keys + ((size_t)tm * B + b_att) * H) + lane)
                          : make_uint4(0, 0, 0, 0);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float2 a = unpack_h2(k[j].x), b = unpack_h2(k[j].y), c = unpack_h2(k[j].z), d = unpack_h2(k[j].w);
            float s = a.x * q[0] + a.y * q[1] + b.x * q[2] + b.y * q[3] + c.x * q[4] + c.y * q[5] + d.x * q[6] + d.y * q[7];
            s = warp_sum(s);
            if (lane == 0 && tm0 + 4 * j < L) sc[tm0 + 4 * j] = gs * s;
          }
        }
        asm volatile("bar.sync %0, 128;" ::"r"(att_bar_id) : "memory");
        // masked softmax over the L scores (128 threads)
        float mx = -INFINITY;
        for (int tm = gt; tm < L; tm += 128) mx = fmaxf(mx, sc[tm]);
        mx = warp_max(mx);
        if (lane == 0) red[w4] = mx;
        asm volatile("bar.sync %0, 128;" ::"r"(att_bar_id) : "memory");
        mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
        float sum = 0.0f;
        for (int tm = gt; tm < L; tm += 128) {
          const float e = __expf(sc[tm] - mx);
          sc[tm] = e;
          sum += e;
        }
        sum = warp_sum(sum);
        if (lane == 0) red[4 + w4] = sum;
        asm volatile("bar.sync %0, 128;" ::"r"(att_bar_id) : "memory");
        const float inv = L > 0 ? 1.0f / ((red[4] + red[5]) + (red[6] + red[7])) : 0.0f;
        float* arow = p.align + ((size_t)t * B + b_att) * Tm;
        for (int tm = gt; tm < Tm; tm += 128) {
          const float a = tm < L ? sc[tm] * inv : 0.0f;
          if (tm < L) sc[tm] = a;
          arow[tm] = a;
        }
        asm volatile("bar.sync %0, 128;" ::"r"(att_bar_id) : "memory");
        // context: rows tm = w4 + 4*i, lane accumulates dims 8*lane .. +7
        for (int tm0 = w4; tm0 < L; tm0 += 16) {
          uint4 v[4];
          float a[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int tm = tm0 + 4 * j;
            a[j] = tm < L ? sc[tm] : 0.0f;
            v[j] = tm < L ? __ldg(reinterpret_cast<const uint4*>(p.values + ((size_t)tm * B + b_att) * DM) + lane)
                          : make_uint4(0, 0, 0, 0);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float2 x0 = unpack_h2(v[j].x), x1 = unpack_h2(v[j].y), x2 = unpack_h2(v[j].z), x3 = unpack_h2(v[j].w);
            ctxv[0] = fmaf(a[j], x0.x, ctxv[0]); ctxv[1] = fmaf(a[j], x0.y, ctxv[1]);
            ctxv[2] = fmaf(a[j], x1.x, ctxv[2]); ctxv[3] = fmaf(a[j], x1.y, ctxv[3]);
            ctxv[4] = fmaf(a[j], x2.x, ctxv[4]); ctxv[5] = fmaf(a[j], x2.y, ctxv[5]);
            ctxv[6] = fmaf(a[j], x3.x, ctxv[6]); ctxv[7] = fmaf(a[j], x3.y, ctxv[7]);
          }
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) part[w4 * DM + 8 * lane + e] = ctxv[e];
        asm volatile("bar.sync %0, 128;" ::"r"(att_bar_id) : "memory");
        if (w4 == 0) {
#pragma unroll
          for (int e = 0; e < 8; ++e)
            ctxv[e] = tf32_rn((part[8 * lane + e] + part[DM + 8 * lane + e]) + (part[2 * DM + 8 * lane + e] + part[3 * DM + 8 * lane + e]));
        }
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
        const uint32_t off = sw128h_off(NB, bl_att, H + 8 * lane);
        const uint32_t dbuf = sOp + nb * OP_BYTES + off;
        const uint32_t c0 = pack_h2(ctxv[0], ctxv[1]), c1 = pack_h2(ctxv[2], ctxv[3]);
        const uint32_t c2 = pack_h2(ctxv[4], ctxv[5]), c3 = pack_h2(ctxv[6], ctxv[7]);
#pragma unroll
        for (uint32_t dst = 0; dst < (uint32_t)CL; ++dst) st_async_v4(mapa(dbuf, dst), mapa(cbar_n, dst), c0, c1, c2, c3);
      }
    }
    if (comb && b0 + bq < B) {
      const size_t o = (size_t)(b0 + bq) * H + 32 * rank + 4 * uq;
      if (p.cT) *reinterpret_cast<float4*>(p.cT + o) = make_float4(c_state[0], c_state[1], c_state[2], c_state[3]);
      if (p.hT) *reinterpret_cast<float4*>(p.hT + o) = make_float4(h_state[0], h_state[1], h_state[2], h_state[3]);
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == GM_WARPS)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem_base) : "memory");
  cluster_sync_all();
}

constexpr size_t SMEM_BYTES = (size_t)W_BYTES + 2 * OP_BYTES + 4 * NB * 32 * 4 + 2 * MAX_TM * 4 + 2 * 4 * DM * 4 + 64 + 64 + 1024;

__global__ void to_half_kernel(const float* __restrict__ src, __half* __restrict__ dst, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = __float2half_rn(src[i]);
}

}  // namespace ap

size_t attn_persist_work_floats(int B, int H, int Dm, int Tm) {
  // fused weights [(H+Dm),4H] + product scratch [H,4H] + fp16 keys / values
  return (size_t)(H + Dm) * 4 * H + (size_t)H * 4 * H + ((size_t)Tm * B * (H + Dm) + 1) / 2 + 64;
}

// Forward of a single-mechanism Luong-family attention layer with the persistent kernel.  Returns -1 when
// the shape is not supported (caller falls back to the per-step path).
int attn_persist_fwd(cudaStream_t st, const AvsrRnnSeq* r, float* scratch) {
  using namespace ap;
  if (r->n_mech != 1 || r->T <= 0) return -1;
  const AvsrAttnMech& m = r->mech[0];
  if (m.kind > AVSR_ATTN_SCALED_LUONG) return -1;
  if (r->H != H || m.A != H || m.Dm != DM || m.Tm > MAX_TM) return -1;
  const int T = r->T, B = r->B, At = m.A, SW = At + H;
  float* Wp = scratch;
  float* tmp = Wp + (size_t)(H + DM) * 4 * H;
  __half* keys_h = reinterpret_cast<__half*>(tmp + (size_t)H * 4 * H);
  __half* values_h = keys_h + (size_t)m.Tm * B * H;
  // step 0: att_{-1} = 0, so only h_0 Wh enters (exact AttentionWrapper zero-state semantics)
  AVSR_TRY(gemm(st, 0, 0, B, 4 * H, H, r->S + At, SW, r->Wrec + (size_t)At * 4 * H, 4 * H, r->gates, 4 * H, 1.0f, nullptr));
  // fused recurrent matrix W' = [Wh + Wl_h Wa ; Wl_c Wa]
  AVSR_CHECK_CUDA(cudaMemcpyAsync(Wp, r->Wrec + (size_t)At * 4 * H, (size_t)H * 4 * H * sizeof(float),
                                  cudaMemcpyDeviceToDevice, st));
  AVSR_TRY(gemm(st, 0, 0, H, 4 * H, At, m.Wl, m.A, r->Wrec, 4 * H, Wp, 4 * H, 1.0f, nullptr));
  AVSR_TRY(gemm(st, 0, 0, DM, 4 * H, At, m.Wl + (size_t)H * m.A, m.A, r->Wrec, 4 * H, Wp + (size_t)H * 4 * H, 4 * H, 0.0f,
                nullptr));
  const long long nk = (long long)m.Tm * B * H, nv = (long long)m.Tm * B * DM;
  AVSR_LAUNCH(to_half_kernel, cdiv(nk, 256), 256, 0, st, m.keys, keys_h, nk);
  AVSR_LAUNCH(to_half_kernel, cdiv(nv, 256), 256, 0, st, m.values, values_h, nv);
  Params p;
  p.T = T; p.B = B; p.Tm = m.Tm;
  p.scaled = m.kind == AVSR_ATTN_SCALED_LUONG;
  p.out_h = 0;
  p.len = r->len; p.mem_len = m.mem_len; p.gates = r->gates; p.Wp = Wp; p.keys = keys_h; p.values = values_h;
  p.g = m.g; p.c0 = r->c0; p.S = r->S; p.SW = SW; p.At = At; p.craw = r->craw; p.out = r->out; p.hc = m.hc;
  p.align = m.align; p.cT = r->cT; p.hT = r->hT;
  static bool attr = false;
  if (!attr) {
    AVSR_CHECK_CUDA(cudaFuncSetAttribute(attn_lstm_persist_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)SMEM_BYTES));
    attr = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(cdiv(B, NB) * CL);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CL;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  AVSR_CHECK_CUDA(cudaLaunchKernelEx(&cfg, attn_lstm_persist_fwd_kernel, p));
  ++g_launch_count;
  // attention vectors of all steps in one product: S[1:, :, :At] = [h | ctx] Wl (tf32-rounded operand rows)
  AVSR_TRY(gemm(st, 0, 0, T * B, At, H + DM, m.hc, H + DM, m.Wl, m.A, r->S + (size_t)B * SW, SW, 0.0f, nullptr, 1));
  return 0;
}

}  // namespace avsr
