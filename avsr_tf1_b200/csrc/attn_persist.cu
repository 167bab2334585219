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

// NB = utterances per cluster: 16 (8 gate/attention warps per CTA) up to 240 utterances; 32 (16 warps) above,
// because a B200 keeps at most 15 clusters of 8 CTAs resident and a 16th cluster would run as a second wave.
constexpr int CL = 8;
constexpr int H = 256;
constexpr int DM = 256;
constexpr int KTOT = H + DM;            // 512
constexpr int KB = KTOT / 64;           // 8 K-blocks of 64 halves (128 B)
constexpr int W_BYTES = KB * 128 * 128; // 128 KB
constexpr int MAX_TM = 384;
constexpr int RIF = 8;                  // memory rows in flight per warp in the attention sweeps (L2 latency hiding)

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

// Sums each of the 8 per-lane values v[0..7] over the 32 lanes with 9 shuffles (instead of 8 x 5): after each
// exchange a lane keeps half of the values.  Returns the complete sum of v[j] in the lanes with
// j == ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1).
__device__ __forceinline__ float warp_reduce8(const float (&v)[8], int lane) {
  const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
  float k0 = (h16 ? v[4] : v[0]) + __shfl_xor_sync(0xffffffffu, h16 ? v[0] : v[4], 16);
  float k1 = (h16 ? v[5] : v[1]) + __shfl_xor_sync(0xffffffffu, h16 ? v[1] : v[5], 16);
  float k2 = (h16 ? v[6] : v[2]) + __shfl_xor_sync(0xffffffffu, h16 ? v[2] : v[6], 16);
  float k3 = (h16 ? v[7] : v[3]) + __shfl_xor_sync(0xffffffffu, h16 ? v[3] : v[7], 16);
  float m0 = (h8 ? k2 : k0) + __shfl_xor_sync(0xffffffffu, h8 ? k0 : k2, 8);
  float m1 = (h8 ? k3 : k1) + __shfl_xor_sync(0xffffffffu, h8 ? k1 : k3, 8);
  float n = (h4 ? m1 : m0) + __shfl_xor_sync(0xffffffffu, h4 ? m0 : m1, 4);
  n += __shfl_xor_sync(0xffffffffu, n, 2);
  n += __shfl_xor_sync(0xffffffffu, n, 1);
  return n;
}

// one memory row (256 halves) as 32 lanes x 16 bytes; rows at or past `L` read as zeros without touching memory
__device__ __forceinline__ uint4 ld_row(const __half* __restrict__ mat, int tm, int L, int B, int b, int lane) {
  return tm < L ? __ldg(reinterpret_cast<const uint4*>(mat + ((size_t)tm * B + b) * 256) + lane) : make_uint4(0, 0, 0, 0);
}
__device__ __forceinline__ float dot8(const uint4& r, const float (&q)[8]) {
  const float2 a = unpack_h2(r.x), b = unpack_h2(r.y), c = unpack_h2(r.z), d = unpack_h2(r.w);
  return a.x * q[0] + a.y * q[1] + b.x * q[2] + b.y * q[3] + c.x * q[4] + c.y * q[5] + d.x * q[6] + d.y * q[7];
}
__device__ __forceinline__ void axpy8(float w, const uint4& r, float (&acc)[8]) {
  const float2 a = unpack_h2(r.x), b = unpack_h2(r.y), c = unpack_h2(r.z), d = unpack_h2(r.w);
  acc[0] = fmaf(w, a.x, acc[0]); acc[1] = fmaf(w, a.y, acc[1]);
  acc[2] = fmaf(w, b.x, acc[2]); acc[3] = fmaf(w, b.y, acc[3]);
  acc[4] = fmaf(w, c.x, acc[4]); acc[5] = fmaf(w, c.y, acc[5]);
  acc[6] = fmaf(w, d.x, acc[6]); acc[7] = fmaf(w, d.y, acc[7]);
}

// byte offset of half element (row, k) in a K-major SWIZZLE_128B operand with 64-half K blocks of `rows` rows
__device__ __forceinline__ uint32_t sw128h_off(int rows, int row, int k) {
  const int kb = k >> 6, kk = k & 63;
  return (uint32_t)(kb * rows * 128 + row * 128 + ((((kk >> 3) ^ (row & 7)) << 4)) + ((kk & 7) << 1));
}

// instruction descriptor: D = f32, A = B = f16, both K-major, N = nb, M = 128
__host__ __device__ constexpr uint32_t idesc_for(int nb) {
  return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(nb >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

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
  long long* dbg;        // AVSR_AP_DEBUG: clock samples [64 steps][12] of CTA 0, thread 0
};
#define AP_STAMP(slot)                                                                          \
  do {                                                                                          \
    if (p.dbg && blockIdx.x == 0 && tid == 0 && t < 64) p.dbg[t * 12 + (slot)] = clock64();  \
  } while (0)

template <int NB>
struct FwdCfg {
  // gate-math / attention warps (4 per attended utterance).  There is no separate MMA-issue warp: the register file
  // is split per SM sub-partition (16384 each, warps dealt round-robin), so a 17th warp would cap every thread of
  // the 32-utterance variant at 96 registers (5 warps on one sub-partition) and the memory sweeps would spill - and
  // spills miss the ~28 KB of L1 left beside 220 KB of shared memory.  Lane 0 of warp 0 issues the products at
  // the two points of the step where it has to wait for the same barriers anyway.
  static constexpr int GMW = NB / 2;
  static constexpr int THREADS = GMW * 32;
  static constexpr int NU = NB / CL;                  // utterances whose attention this CTA owns
  static constexpr int OP_BYTES = KB * NB * 128;      // one [h | ctx] operand buffer
  // the partial-context scratch [NU][4][DM] aliases the activation exchange buffer [4][NB][32] (same size): the
  // activations of step t+1 are only written after every context of step t has been all-gathered
  static constexpr size_t SMEM = (size_t)W_BYTES + 2 * OP_BYTES + 4 * NB * 32 * 4 + NU * MAX_TM * 4 + NU * 8 * 4 + 64 + 1024;
  static_assert(NU * 4 * DM == 4 * NB * 32, "partial-context scratch must fit the activation buffer");
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
};

template <int NB>
__global__ void __launch_bounds__(FwdCfg<NB>::THREADS, 1) attn_lstm_persist_fwd_kernel(const Params p) {
  constexpr int GM_WARPS = FwdCfg<NB>::GMW, THREADS = FwdCfg<NB>::THREADS, NU = FwdCfg<NB>::NU;
  constexpr int OP_BYTES = FwdCfg<NB>::OP_BYTES;
  constexpr uint32_t IDESC = idesc_for(NB);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sW = base;
  const uint32_t sOp = sW + W_BYTES;                 // two operand buffers [h | ctx]
  const uint32_t sAct = sOp + 2 * OP_BYTES;          // [4][NB][32] floats; also [NU][4][DM] partial contexts
  const uint32_t sSc = sAct + 4 * NB * 32 * 4;       // [NU][MAX_TM] scores / alignments
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
    for (int i = 0; i < 5; ++i) mbar_init(sBar + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
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

  {
    // ================= gate math + attention warps (lane 0 of warp 0 also issues the products) =================
    const int g = warp & 3, ch = warp >> 2;
    const int unit = 32 * rank + lane;
    const bool comb = tid < 8 * NB;
    const int uq = tid & 7, bq = (tid >> 3) & (NB - 1);
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
    const int bl_att = NU * (int)rank + jl;      // row of the utterance in the operand buffers
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
      AP_STAMP(0);
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
      AP_STAMP(1);
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
      asm volatile("bar.sync 1, %0;" ::"n"(GM_WARPS * 32) : "memory");
      AP_STAMP(2);
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
      AP_STAMP(3);
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
      AP_STAMP(4);
      const bool live_q = t < len_q;    // masked steps (and padding utterances) skip the memory sweep
      // Memory sweeps are software-pipelined in half-batches of 4 rows (ra: rows j = 0..3 of a batch of 8, rb:
      // j = 4..7): the loads of the next half-batch are in flight while the current one is consumed, and the first
      // batch of the keys is requested before the h all-gather has landed (it does not depend on the query).
      uint4 ra[4], rb[4];
      if (live_q) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          ra[j] = ld_row(p.keys, w4 + 4 * j, L, B, b_att, lane);
          rb[j] = ld_row(p.keys, w4 + 16 + 4 * j, L, B, b_att, lane);
        }
      }
      if (tid == 0) mbar_expect_tx(hbar_n, NB * H * 2);
      mbar_wait(hbar_n, (t >> 1) & 1);  // every CTA's h_t slice has landed in buffer nb
      if (warp == 0 && t + 1 < T) {
        // h half of the gate product of step t+1 (it overlaps this step's attention).  Every warp has read the
        // accumulators of step t before its activations reached the barrier that precedes the h all-gather.
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (lane == 0) {
          const uint32_t ob = sOp + nb * OP_BYTES;
#pragma unroll
          for (int kb = 0; kb < H / 64; ++kb)
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4)
              umma_f16(tmem_base, make_desc_k128(sW + kb * (128 * 128) + k4 * 32),
                       make_desc_k128(ob + kb * (NB * 128) + k4 * 32), IDESC, (kb | k4) ? 1u : 0u);
        }
        __syncwarp();
      }
      AP_STAMP(5);
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
        // scores: rows tm = w4 + 4*i, RIF rows in flight per warp, one 9-shuffle reduction per batch of 8 rows
        static_assert(RIF == 8, "warp_reduce8 expects 8 rows per batch");
        const int jrow = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
        for (int tm0 = w4; tm0 < L; tm0 += 4 * RIF) {
          float sacc[RIF];
#pragma unroll
          for (int j = 0; j < 4; ++j) sacc[j] = dot8(ra[j], q);
#pragma unroll
          for (int j = 0; j < 4; ++j) ra[j] = ld_row(p.keys, tm0 + 32 + 4 * j, L, B, b_att, lane);
#pragma unroll
          for (int j = 0; j < 4; ++j) sacc[4 + j] = dot8(rb[j], q);
#pragma unroll
          for (int j = 0; j < 4; ++j) rb[j] = ld_row(p.keys, tm0 + 48 + 4 * j, L, B, b_att, lane);
          const float tot = warp_reduce8(sacc, lane);
          if ((lane & 3) == 0 && tm0 + 4 * jrow < L) sc[tm0 + 4 * jrow] = gs * tot;
        }
        // first batch of the values: in flight during the softmax
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          ra[j] = ld_row(p.values, w4 + 4 * j, L, B, b_att, lane);
          rb[j] = ld_row(p.values, w4 + 16 + 4 * j, L, B, b_att, lane);
        }
        asm volatile("bar.sync %0, 128;" ::"r"(att_bar_id) : "memory");
        AP_STAMP(6);
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
        AP_STAMP(7);
        // context: rows tm = w4 + 4*i, lane accumulates dims 8*lane .. +7
        for (int tm0 = w4; tm0 < L; tm0 += 4 * RIF) {
          float a[RIF];
#pragma unroll
          for (int j = 0; j < RIF; ++j) a[j] = tm0 + 4 * j < L ? sc[tm0 + 4 * j] : 0.0f;
#pragma unroll
          for (int j = 0; j < 4; ++j) axpy8(a[j], ra[j], ctxv);
#pragma unroll
          for (int j = 0; j < 4; ++j) ra[j] = ld_row(p.values, tm0 + 32 + 4 * j, L, B, b_att, lane);
#pragma unroll
          for (int j = 0; j < 4; ++j) axpy8(a[4 + j], rb[j], ctxv);
#pragma unroll
          for (int j = 0; j < 4; ++j) rb[j] = ld_row(p.values, tm0 + 48 + 4 * j, L, B, b_att, lane);
        }
        AP_STAMP(8);
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
      AP_STAMP(9);
      if (warp == 0) {
        // ctx half of the gate product of step t+1, once every context of this step has landed.  After the last
        // step the wait only drains the all-gathers: no st.async may be in flight towards this CTA when it exits.
        if (lane == 0) mbar_expect_tx(cbar_n, NB * DM * 2);
        mbar_wait(cbar_n, (t >> 1) & 1);
        if (t + 1 < T) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (lane == 0) {
            const uint32_t ob = sOp + nb * OP_BYTES;
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
    }
    if (comb && b0 + bq < B) {
      const size_t o = (size_t)(b0 + bq) * H + 32 * rank + 4 * uq;
      if (p.cT) *reinterpret_cast<float4*>(p.cT + o) = make_float4(c_state[0], c_state[1], c_state[2], c_state[3]);
      if (p.hT) *reinterpret_cast<float4*>(p.hT + o) = make_float4(h_state[0], h_state[1], h_state[2], h_state[3]);
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem_base) : "memory");
  cluster_sync_all();
}


__global__ void to_half_kernel(const float* __restrict__ src, __half* __restrict__ dst, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = __float2half_rn(src[i]);
}


// =====================================================================================================
// backward of the same layer, same fused algebra:
//     [dh_{t-1} | dctx_{t-1}] = dz_t W'^T + (dout_{t-1} Wl^T)          (second term precomputed for all t)
// CTA `rank` owns the gate columns of its 32 units (K split of the product, reduce-scattered through DSMEM:
// h rows to the owners of the units, ctx rows to the owners of the utterances) and the attention backward
// of 2 utterances:  d(align) = values.dctx -> softmax backward -> ds (saved; dkeys/dvalues are formed after
// the loop) -> dq = g * sum_tm ds keys  -> all-to-all to the owners of the units.
// dz enters the tensor core as fp16 scaled by a power of two (`grad_scale`, undone on the accumulators).
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
  long long* dbg;        // AVSR_AP_DEBUG: clock samples [64 iterations][12] of CTA 0, thread 0
};
#define APB_STAMP(slot)                                                                          \
  do {                                                                                           \
    if (p.dbg && blockIdx.x == 0 && tid == 0 && it < 64) p.dbg[it * 12 + (slot)] = clock64(); \
  } while (0)

constexpr int BW_W_BYTES = 2 * KTOT * 128;           // A operand: 2 K-blocks x [512 rows x 128 B]

template <int NB>
struct BwdCfg {
  static constexpr int GMW = NB / 2;
  static constexpr int THREADS = GMW * 32;             // no separate MMA-issue warp (see FwdCfg)
  static constexpr int NU = NB / CL;                   // utterances whose attention backward this CTA owns
  static constexpr int DZ_BYTES = 2 * NB * 128;        // B operand: 2 K-blocks x [NB rows x 128 B]
  static constexpr int REDH_FLOATS = CL * 32 * NB;     // [src][u][b]
  static constexpr int REDC_FLOATS = CL * NU * DM;     // [src][utt][dim]; also the dq partial scratch [NU][4][DM]
  static constexpr int DQ_FLOATS = CL * NU * 32;       // [src][utt][u]
  static constexpr int TCOLS = 4 * NB;                 // TMEM columns: four 128-row tiles of [h | ctx]
  static constexpr size_t SMEM = (size_t)BW_W_BYTES + DZ_BYTES + REDH_FLOATS * 4 + REDC_FLOATS * 4 + DQ_FLOATS * 4 +
                                 NU * DM * 4 + 2 * NU * MAX_TM * 4 + NU * 8 * 4 + 64 + 1024;
  static_assert(NU * 4 * DM <= REDC_FLOATS, "dq partial scratch must fit the ctx reduce buffer");
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
};

__device__ __forceinline__ void st_async_v4f(uint32_t addr, uint32_t mbar, float a, float b, float c, float d) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%2, %3, %4, %5}, [%1];" ::"r"(addr),
               "r"(mbar), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}
__device__ __forceinline__ void st_async_f(uint32_t addr, uint32_t mbar, float a) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f32 [%0], %2, [%1];" ::"r"(addr), "r"(mbar), "f"(a)
               : "memory");
}
template <int N>
__device__ __forceinline__ void tmem_ldn(uint32_t taddr, uint32_t (&r)[N]);
template <>
__device__ __forceinline__ void tmem_ldn<16>(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
template <>
__device__ __forceinline__ void tmem_ldn<32>(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}

// Per iteration (time step t = T-1-it) and CTA:
//   top   wait for the partial sums pushed during the previous iteration (h rows -> redH, ctx rows -> redC) and
//         fold them into registers / dctx_s; ONE CTA-wide barrier: from here on redH and redC are free again, so a
//         single copy of each suffices (a peer can only push the next partials after it has received this CTA's
//         dq, which leaves after the barrier) and redC doubles as the dq partial scratch
//   A/B   attention backward of the CTA's utterances, dq all-to-all
//   C/D   gate gradients dz_t -> shared-memory operand, HBM
//   E     partial products from TMEM -> owners (reduce-scatter through DSMEM)
template <int NB>
__global__ void __launch_bounds__(BwdCfg<NB>::THREADS, 1) attn_lstm_persist_bwd_kernel(const BwdParams p) {
  using C = BwdCfg<NB>;
  constexpr int GM_WARPS = C::GMW, THREADS = C::THREADS, NU = C::NU;
  constexpr int REDH_FLOATS = C::REDH_FLOATS, REDC_FLOATS = C::REDC_FLOATS, DQ_FLOATS = C::DQ_FLOATS;
  constexpr uint32_t IDESC = idesc_for(NB);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sW = base;
  const uint32_t sDz = sW + BW_W_BYTES;
  const uint32_t sRedH = sDz + C::DZ_BYTES;
  const uint32_t sRedC = sRedH + REDH_FLOATS * 4;
  const uint32_t sDq = sRedC + REDC_FLOATS * 4;
  const uint32_t sCtx = sDq + DQ_FLOATS * 4;                 // [NU][DM] dctx of the CTA's utterances
  const uint32_t sSc = sCtx + NU * DM * 4;                   // [NU][MAX_TM] alignments
  const uint32_t sDs = sSc + NU * MAX_TM * 4;                // [NU][MAX_TM] d(align) / ds
  const uint32_t sRed = sDs + NU * MAX_TM * 4;               // [NU][8]
  const uint32_t sBar = sRed + NU * 8 * 4;  // [0] mma_done [1] dz_ready [2] redH_full [3] redC_full [4] dq_full
  const uint32_t sTmem = sBar + 40;
  const uint32_t barMma = sBar, barDz = sBar + 8, barRedH = sBar + 16, barRedC = sBar + 24, barDq = sBar + 32;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  float* redH = reinterpret_cast<float*>(gen + (sRedH - base));
  float* redC = reinterpret_cast<float*>(gen + (sRedC - base));
  float* dqb = reinterpret_cast<float*>(gen + (sDq - base));
  float* ctx_all = reinterpret_cast<float*>(gen + (sCtx - base));
  float* sc_all = reinterpret_cast<float*>(gen + (sSc - base));
  float* ds_all = reinterpret_cast<float*>(gen + (sDs - base));
  float* part_all = redC;
  float* red_all = reinterpret_cast<float*>(gen + (sRed - base));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int b0 = cluster_id_x() * NB;
  const int T = p.T, B = p.B, Tm = p.Tm;

  if (tid == 0) {
    mbar_init(barMma, 1);
    mbar_init(barDz, GM_WARPS * 32);
    mbar_init(barRedH, 1);
    mbar_init(barRedC, 1);
    mbar_init(barDq, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sTmem), "n"(C::TCOLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // resident operand: A[n][g*32 + u] = Wp[n][g*H + 32*rank + u] as fp16 (rows n = [h | ctx] dims)
  for (int seg = warp; seg < KTOT * 4; seg += THREADS / 32) {
    const int n = seg >> 2, g = seg & 3;
    const float w = p.Wp[(size_t)n * 4 * H + g * H + 32 * rank + lane];
    *reinterpret_cast<__half*>(gen + (sW - base) + sw128h_off(KTOT, n, g * 32 + lane)) = __float2half_rn(w);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(sTmem));
  cluster_sync_all();

  {
    const int unit = 32 * rank + lane;
    constexpr int PB = 2;
    float dc[PB], dh_carry[PB];
    int len_t[PB];
#pragma unroll
    for (int j = 0; j < PB; ++j) {
      const int b = b0 + warp * PB + j;
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
        const int b = b0 + warp * PB + j;
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
    float* a_s = sc_all + jl * MAX_TM;
    float* ds_s = ds_all + jl * MAX_TM;
    float* part = part_all + jl * 4 * DM;
    float* red = red_all + jl * 8;
    const uint32_t att_bar_id = 2 + jl;
    // reduce-scatter role after the product: warp quad `jl` forwards the 128-row tiles TPW*jl .. TPW*jl+TPW-1
    // (tiles 0, 1: h rows; tiles 2, 3: ctx dims)
    constexpr int TPW = 4 / NU;
    const int q = warp & 3;

    load_step(T - 1);
    for (int it = 0; it < T; ++it) {
      const int t = T - 1 - it;
      const bool live_q = t < len_q;
      // ---- top: partial sums of the previous iteration ---------------------------------------------------
      float dh_in[PB];
#pragma unroll
      for (int j = 0; j < PB; ++j) dh_in[j] = dh_carry[j];
      APB_STAMP(0);
      // software-pipelined memory sweeps (see the forward kernel): the first batch of the values is requested before
      // the partial sums of the previous iteration have arrived
      uint4 ra[4], rb[4];
      if (live_q) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          ra[j] = ld_row(p.values, w4 + 4 * j, L, B, b_att, lane);
          rb[j] = ld_row(p.values, w4 + 16 + 4 * j, L, B, b_att, lane);
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
          const int bl = warp * PB + j;
#pragma unroll
          for (int src = 0; src < CL; ++src)
            dh_in[j] += redH[(src * 32 + lane) * NB + ((((bl >> 2) ^ (lane & (NB / 4 - 1))) << 2) | (bl & 3))];
        }
        mbar_wait(barRedC, (it - 1) & 1);
      }
      APB_STAMP(1);
      if (live_q) {
        for (int d = gt; d < DM; d += 128) {
          float v = p.douthc ? p.douthc[((size_t)t * B + b_att) * (H + DM) + H + d] : 0.0f;
          if (it > 0) {
#pragma unroll
            for (int src = 0; src < CL; ++src) v += redC[(src * NU + jl) * DM + d];
          }
          dctx_s[d] = v;
          p.dhc[((size_t)t * B + b_att) * (H + DM) + H + d] = v;
        }
        for (int tm = gt; tm < Tm; tm += 128) a_s[tm] = p.align[((size_t)t * B + b_att) * Tm + tm];
      }
      asm volatile("bar.sync 1, %0;" ::"n"(GM_WARPS * 32) : "memory");
      APB_STAMP(2);
      // ---- (A/B) attention backward of the CTA's utterances, dq all-to-all -------------------------------
      float dqv[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) dqv[e] = 0.0f;
      if (live_q) {
        float dcx[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) dcx[e] = dctx_s[8 * lane + e];
        // d(align)[tm] = values[tm] . dctx
        const int jrow = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
        for (int tm0 = w4; tm0 < L; tm0 += 4 * RIF) {
          float sacc[RIF];
#pragma unroll
          for (int j = 0; j < 4; ++j) sacc[j] = dot8(ra[j], dcx);
#pragma unroll
          for (int j = 0; j < 4; ++j) ra[j] = ld_row(p.values, tm0 + 32 + 4 * j, L, B, b_att, lane);
#pragma unroll
          for (int j = 0; j < 4; ++j) sacc[4 + j] = dot8(rb[j], dcx);
#pragma unroll
          for (int j = 0; j < 4; ++j) rb[j] = ld_row(p.values, tm0 + 48 + 4 * j, L, B, b_att, lane);
          const float tot = warp_reduce8(sacc, lane);
          if ((lane & 3) == 0 && tm0 + 4 * jrow < L) ds_s[tm0 + 4 * jrow] = tot;
        }
        // first batch of the keys: in flight during the softmax backward
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          ra[j] = ld_row(p.keys, w4 + 4 * j, L, B, b_att, lane);
          rb[j] = ld_row(p.keys, w4 + 16 + 4 * j, L, B, b_att, lane);
        }
        asm volatile("bar.sync %0, 128;" ::"r"(att_bar_id) : "memory");
        APB_STAMP(3);
        float dot = 0.0f;
        for (int tm = gt; tm < L; tm += 128) dot = fmaf(a_s[tm], ds_s[tm], dot);
        dot = warp_sum(dot);
        if (lane == 0) red[w4] = dot;
        asm volatile("bar.sync %0, 128;" ::"r"(att_bar_id) : "memory");
        dot = (red[0] + red[1]) + (red[2] + red[3]);
        // d(score) of this step, before the Luong scale (dkeys is formed after the loop).  d(attention_g) =
        // sum_tm ds[tm] (keys[tm].h) = (1/g) sum_tm ds[tm] log a[tm]: the scores are log a + log Z over g and
        // sum_tm ds[tm] = 0, so the normaliser drops out and the keys need not be multiplied by the query again.
        float* dsrow = p.ds + ((size_t)t * B + b_att) * Tm;
        float gacc = 0.0f;
        for (int tm = gt; tm < Tm; tm += 128) {
          const float a = tm < L ? a_s[tm] : 0.0f;
          const float d = a * (ds_s[tm] - dot);
          dsrow[tm] = tm < L ? d : 0.0f;
          if (tm < L) ds_s[tm] = d;
          if (a > 0.0f) gacc = fmaf(d, __logf(a), gacc);
        }
        if (p.scaled && p.dg && gs != 0.0f) {
          gacc = warp_sum(gacc);
          if (lane == 0) atomicAdd(p.dg, gacc / gs);
        }
        asm volatile("bar.sync %0, 128;" ::"r"(att_bar_id) : "memory");
        APB_STAMP(4);
        // keys sweep: dq += g * ds[tm] * keys[tm]
        for (int tm0 = w4; tm0 < L; tm0 += 4 * RIF) {
          float d[RIF];
#pragma unroll
          for (int j = 0; j < RIF; ++j) d[j] = tm0 + 4 * j < L ? ds_s[tm0 + 4 * j] : 0.0f;
#pragma unroll
          for (int j = 0; j < 4; ++j) axpy8(d[j], ra[j], dqv);
#pragma unroll
          for (int j = 0; j < 4; ++j) ra[j] = ld_row(p.keys, tm0 + 32 + 4 * j, L, B, b_att, lane);
#pragma unroll
          for (int j = 0; j < 4; ++j) axpy8(d[4 + j], rb[j], dqv);
#pragma unroll
          for (int j = 0; j < 4; ++j) rb[j] = ld_row(p.keys, tm0 + 48 + 4 * j, L, B, b_att, lane);
        }
        APB_STAMP(5);
#pragma unroll
        for (int e = 0; e < 8; ++e) part[w4 * DM + 8 * lane + e] = dqv[e];
        asm volatile("bar.sync %0, 128;" ::"r"(att_bar_id) : "memory");
        if (w4 == 0) {
#pragma unroll
          for (int e = 0; e < 8; ++e)
            dqv[e] = gs * ((part[8 * lane + e] + part[DM + 8 * lane + e]) + (part[2 * DM + 8 * lane + e] + part[3 * DM + 8 * lane + e]));
        }
      }
      if (w4 == 0) {
        // dq dims 8*lane .. +7 belong to the CTA owning units (8*lane)/32
        const uint32_t dst = (uint32_t)(lane >> 2);
        const uint32_t a0 = mapa(sDq + (uint32_t)(((rank * NU + jl) * 32 + ((8 * lane) & 31)) * 4), dst);
        const uint32_t bar = mapa(barDq, dst);
        st_async_v4f(a0, bar, dqv[0], dqv[1], dqv[2], dqv[3]);
        st_async_v4f(a0 + 16, bar, dqv[4], dqv[5], dqv[6], dqv[7]);
      }
      // ---- (C/D) dq of this CTA's units -> gate gradients ------------------------------------------------
      APB_STAMP(6);
      if (tid == 0) mbar_expect_tx(barDq, DQ_FLOATS * 4);
      mbar_wait(barDq, it & 1);
      APB_STAMP(7);
      float dz[4][PB];
#pragma unroll
      for (int j = 0; j < PB; ++j) {
        const int bl = warp * PB + j;
        float dh = dh_in[j];
        if (t < len_t[j]) {
          dh += dov[j] + dqb[bl * 32 + lane];
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
          *reinterpret_cast<__half*>(gen + (sDz - base) + sw128h_off(NB, bl, g * 32 + lane)) =
              __float2half_rn(dz[g][j] * p.grad_scale);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(barDz);
      if (warp == 0) {
        // partial [h | ctx](512) x NB from this CTA's 128 gate columns, once every warp's dz is in shared memory
        mbar_wait(barDz, it & 1);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (lane == 0) {
#pragma unroll
          for (int mt = 0; mt < KTOT / 128; ++mt)
#pragma unroll
            for (int kb = 0; kb < 2; ++kb)
#pragma unroll
              for (int k4 = 0; k4 < 4; ++k4)
                umma_f16(tmem_base + mt * NB, make_desc_k128(sW + kb * (KTOT * 128) + mt * (128 * 128) + k4 * 32),
                         make_desc_k128(sDz + kb * (NB * 128) + k4 * 32), IDESC, (kb | k4) ? 1u : 0u);
          umma_commit(barMma);
        }
        __syncwarp();
      }
      APB_STAMP(8);
#pragma unroll
      for (int j = 0; j < PB; ++j) {
        const int b = b0 + warp * PB + j;
        if (b < B) {
          float* o = p.dZ + ((size_t)t * B + b) * 4 * H + unit;
          o[0] = tf32_rn(dz[0][j]); o[H] = tf32_rn(dz[1][j]); o[2 * H] = tf32_rn(dz[2][j]); o[3 * H] = tf32_rn(dz[3][j]);
        }
      }
      load_step(t - 1);
      APB_STAMP(9);
      // ---- (E) partial products -> owners ---------------------------------------------------------------
      mbar_wait(barMma, it & 1);
      APB_STAMP(10);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int mi = 0; mi < TPW; ++mi) {
        const int mt = TPW * jl + mi;
        if (mt < 2) {  // h rows 128*mt + 32*q + lane -> owner CTA 4*mt + q
          uint32_t r[NB];
          tmem_ldn<NB>(tmem_base + ((uint32_t)(32 * q) << 16) + mt * NB, r);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          const uint32_t dst = (uint32_t)(4 * mt + q);
          const uint32_t a0 = mapa(sRedH + (uint32_t)((rank * 32 + lane) * NB) * 4, dst);
          const uint32_t bar = mapa(barRedH, dst);
#pragma unroll
          for (int v = 0; v < NB / 4; ++v)  // 16-byte chunks XOR-swizzled by the unit: conflict-light reads on the owner
            st_async_v4f(a0 + ((v ^ (lane & (NB / 4 - 1))) << 4), bar, __uint_as_float(r[4 * v]) * p.inv_grad_scale,
                         __uint_as_float(r[4 * v + 1]) * p.inv_grad_scale, __uint_as_float(r[4 * v + 2]) * p.inv_grad_scale,
                         __uint_as_float(r[4 * v + 3]) * p.inv_grad_scale);
        } else if (it + 1 < T) {  // ctx dims 128*(mt-2) + 32*q + lane; column c = utterance -> owner CTA c / NU
          uint32_t r[NB];
          tmem_ldn<NB>(tmem_base + ((uint32_t)(32 * q) << 16) + mt * NB, r);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          const int dim = 128 * (mt - 2) + 32 * q + lane;
#pragma unroll
          for (uint32_t dst = 0; dst < (uint32_t)CL; ++dst) {
            // redC[src = rank][utt 0..NU-1][dim]: the utterances of one owner are DM floats apart
            const uint32_t a0 = mapa(sRedC + (uint32_t)((rank * NU) * DM + dim) * 4, dst);
            const uint32_t bar = mapa(barRedC, dst);
#pragma unroll
            for (int ul = 0; ul < NU; ++ul)
              st_async_f(a0 + ul * DM * 4, bar, __uint_as_float(r[NU * dst + ul]) * p.inv_grad_scale);
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      APB_STAMP(11);
    }
    // drain the last reduce-scatter (nothing may be in flight towards this CTA when it exits).  Its content is
    // NOT dh_0: step 0 saw att_{-1} = 0, so dh_0 = dz_0 Wh^T with the un-fused Wh - added by the host; here only
    // the gradient carried through fully masked utterances is written.
    if (T > 0) {
      if (tid == 0) mbar_expect_tx(barRedH, REDH_FLOATS * 4);
      mbar_wait(barRedH, (T - 1) & 1);
    }
#pragma unroll
    for (int j = 0; j < PB; ++j) {
      const int b = b0 + warp * PB + j;
      if (b < B) {
        if (p.dh0) p.dh0[(size_t)b * H + unit] = dh_carry[j];
        if (p.dc0) p.dc0[(size_t)b * H + unit] = dc[j];
      }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::TCOLS) : "memory");
  cluster_sync_all();
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

// Utterances per cluster.  A B200 keeps at most 15 clusters of 8 CTAs resident, so more than 240 utterances
// take 32-utterance slices (one wave of <= 8 clusters per 256) instead of a second wave of 16-utterance ones.
// AVSR_AP_SLICE=16|32 overrides the choice (parity tests of the wide variant at small batches).
static int slice_width(int B) {
  if (const char* e = getenv("AVSR_AP_SLICE")) {
    const int v = atoi(e);
    if (v == 16 || v == 32) return v;
  }
  return B > 240 ? 32 : 16;
}

static void print_phase_clocks(const char* tag, const long long* h, int n, int ns, const char* const* names) {
  double acc[12] = {0};
  for (int t = 3; t < n; ++t) {
    for (int k = 1; k < ns; ++k) acc[k] += (double)(h[t * 12 + k] - h[t * 12 + k - 1]);
    acc[0] += (double)(h[t * 12] - h[(t - 1) * 12 + ns - 1]);
  }
  fprintf(stderr, "%s clocks/step:", tag);
  double tot = 0;
  for (int k = 0; k < ns; ++k) {
    fprintf(stderr, " %s=%.0f", names[k], acc[k] / (n - 3));
    tot += acc[k] / (n - 3);
  }
  fprintf(stderr, " total=%.0f\n", tot);
}

template <typename Kern, typename P>
static int launch_cluster(cudaStream_t st, Kern kern, int B, int nb, int threads, size_t smem, const P& p) {
  AVSR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(cdiv(B, nb) * CL);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CL;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  AVSR_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
  ++g_launch_count;
  return 0;
}

// Cluster width of the persistent attention kernels: 4 (attn_persist4.cu: 33 clusters fit a B200, 2 attended
// utterances per CTA) unless AVSR_AP_CLUSTER=8 selects the kernels of this file.
static int cluster_width() {
  if (const char* e = getenv("AVSR_AP_CLUSTER")) {
    if (atoi(e) == 8) return 8;
  }
  return 4;
}

}  // namespace ap
int cluster_width_ap() { return ap::cluster_width(); }

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
int cluster_width_ap();
static bool wlas_shape_ok(const AvsrRnnSeq* r);
int rnn_sampling_fused(const AvsrRnnSeq* r) {
  if (r && r->rng && tensor_cores_enabled() && r->n_mech == 2 && !(r->t_begin || r->t_end || r->stepwise) &&
      !getenv("AVSR_NO_ATTN_PERSIST"))
    return wlas_shape_ok(r);  // dual-attention decoder: cluster-of-8 kernels (attn_persist8w.cu)
  if (!(r && r->rng && tensor_cores_enabled() && persist_shape_ok(r) && cluster_width_ap() == 4 &&
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
  AVSR_LAUNCH(to_half_kernel, cdiv(nk, 256), 256, 0, st, m.keys, keys_h, nk);
  AVSR_LAUNCH(to_half_kernel, cdiv(nk, 256), 256, 0, st, pv, pv_h, nk);
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
    if (!persist_shape_ok(r) || cluster_width() != 4 || getenv("AVSR_NO_BAHDANAU_PERSIST")) return -1;
    return bahdanau_persist_fwd(st, r, scratch);
  }
  if (r->H != H || m.A != H || m.Dm != DM || m.Tm > MAX_TM) return -1;
  if (unfused_fwd(r) && (!r->output_attention || cluster_width() != 4)) return -1;
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
    AVSR_LAUNCH(to_half_kernel, cdiv(nk, 256), 256, 0, st, m.keys, keys_h, nk);
    AVSR_LAUNCH(to_half_kernel, cdiv(nv, 256), 256, 0, st, m.values, values_h, nv);
    return attn_persist4d_launch_fwd(st, r, keys_h, values_h);
  }
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
  p.dbg = nullptr;
  long long* dbg_dev = nullptr;
  if (getenv("AVSR_AP_DEBUG")) {
    AVSR_CHECK_CUDA(cudaMalloc(&dbg_dev, 64 * 12 * sizeof(long long)));
    AVSR_CHECK_CUDA(cudaMemset(dbg_dev, 0, 64 * 12 * sizeof(long long)));
    p.dbg = dbg_dev;
  }
  if (cluster_width() == 4) {
    AVSR_TRY(attn_persist4_launch_fwd(st, T, B, m.Tm, p.scaled, p.len, p.mem_len, p.gates, p.Wp, p.keys, p.values, p.g, p.c0, p.S,
                                      p.SW, p.At, p.craw, p.out, p.hc, p.align, p.cT, p.hT));
  } else {
    AVSR_TRY(slice_width(B) == 32 ? launch_cluster(st, attn_lstm_persist_fwd_kernel<32>, B, 32, FwdCfg<32>::THREADS, FwdCfg<32>::SMEM, p)
                                   : launch_cluster(st, attn_lstm_persist_fwd_kernel<16>, B, 16, FwdCfg<16>::THREADS, FwdCfg<16>::SMEM, p));
  }
  if (dbg_dev) {
    AVSR_CHECK_CUDA(cudaStreamSynchronize(st));
    long long h[64 * 12];
    AVSR_CHECK_CUDA(cudaMemcpy(h, dbg_dev, sizeof(h), cudaMemcpyDeviceToHost));
    cudaFree(dbg_dev);
    const char* names[10] = {"loop-top", "wait MMA+ld", "act+bar", "combine+send h", "hbm st/ld", "wait h gather",
                             "scores", "softmax", "ctx sweep", "reduce+send ctx"};
    char tag[96];
    snprintf(tag, sizeof tag, "[ap fwd T=%d B=%d Tm=%d]", T, B, m.Tm);
    print_phase_clocks(tag, h, T < 64 ? T : 64, 10, names);
  }
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
  AVSR_LAUNCH(to_half_kernel, cdiv(nk, 256), 256, 0, st, m.keys, keys_h, nk);
  AVSR_LAUNCH(to_half_kernel, cdiv(nk, 256), 256, 0, st, pv, pv_h, nk);
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
    AVSR_LAUNCH(to_half_kernel, cdiv(nk, 256), 256, 0, st, m.keys, keys_h, nk);
    AVSR_LAUNCH(to_half_kernel, cdiv(nk, 256), 256, 0, st, pv, pv_h, nk);
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
    if (!persist_shape_ok(r) || cluster_width() != 4 || getenv("AVSR_NO_BAHDANAU_PERSIST") ||
        getenv("AVSR_NO_BAHDANAU_PERSIST_BWD"))
      return -1;
    return bahdanau_persist_bwd(st, r, scratch);
  }
  if (r->H != H || m.A != H || m.Dm != DM || m.Tm > MAX_TM) return -1;
  unfused = unfused || unfused_fwd(r);
  if (unfused && (!r->output_attention || cluster_width() != 4)) return -1;
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
    AVSR_LAUNCH(to_half_kernel, cdiv(nk, 256), 256, 0, st, m.keys, keys_h, nk);
    AVSR_LAUNCH(to_half_kernel, cdiv(nv, 256), 256, 0, st, m.values, values_h, nv);
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
  BwdParams p;
  p.T = T; p.B = B; p.Tm = m.Tm; p.scaled = m.kind == AVSR_ATTN_SCALED_LUONG;
  p.grad_scale = r->grad_scale > 0.0f ? r->grad_scale : 1.0f;
  p.inv_grad_scale = 1.0f / p.grad_scale;
  p.len = r->len; p.mem_len = m.mem_len; p.gates = r->gates; p.craw = r->craw; p.c0 = r->c0; p.Wp = Wp;
  p.keys = keys_h; p.values = values_h; p.g = m.g; p.hc = m.hc; p.align = m.align; p.douthc = dhc_in;
  p.dcT = r->dcT; p.dhT = r->dhT; p.dZ = r->dZ; p.ds = m.ds; p.dhc = m.dhc; p.dg = m.dg; p.dc0 = r->dc0; p.dh0 = r->dh0;
  p.dbg = nullptr;
  long long* dbg_dev = nullptr;
  if (getenv("AVSR_AP_DEBUG")) {
    AVSR_CHECK_CUDA(cudaMalloc(&dbg_dev, 64 * 12 * sizeof(long long)));
    AVSR_CHECK_CUDA(cudaMemset(dbg_dev, 0, 64 * 12 * sizeof(long long)));
    p.dbg = dbg_dev;
  }
  if (cluster_width() == 4) {
    AVSR_TRY(attn_persist4_launch_bwd(st, T, B, m.Tm, p.scaled, p.grad_scale, p.len, p.mem_len, p.gates, p.craw, p.c0, p.Wp,
                                      p.keys, p.values, p.g, p.hc, p.align, p.douthc, p.dcT, p.dhT, p.dZ, p.ds, p.dhc, p.dg,
                                      p.dc0, p.dh0, r->dbias));
  } else {
    AVSR_TRY(slice_width(B) == 32 ? launch_cluster(st, attn_lstm_persist_bwd_kernel<32>, B, 32, BwdCfg<32>::THREADS, BwdCfg<32>::SMEM, p)
                                   : launch_cluster(st, attn_lstm_persist_bwd_kernel<16>, B, 16, BwdCfg<16>::THREADS, BwdCfg<16>::SMEM, p));
    if (r->dbias) AVSR_TRY(avsr_colsum((avsr_stream_t)st, r->dZ, T * B, 4 * H, 4 * H, r->dbias));
  }
  if (dbg_dev) {
    AVSR_CHECK_CUDA(cudaStreamSynchronize(st));
    long long h[64 * 12];
    AVSR_CHECK_CUDA(cudaMemcpy(h, dbg_dev, sizeof(h), cudaMemcpyDeviceToHost));
    cudaFree(dbg_dev);
    const char* names[12] = {"loop-top", "wait redH/redC+fold", "dctx+cta bar", "values sweep", "softmax bwd", "keys sweep",
                             "reduce+send dq", "wait dq", "dz->smem", "hbm st/ld", "wait MMA", "tmem ld+push"};
    char tag[96];
    snprintf(tag, sizeof tag, "[ap bwd T=%d B=%d Tm=%d]", T, B, m.Tm);
    print_phase_clocks(tag, h, T < 64 ? T : 64, 12, names);
  }
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
  AVSR_TRY(attn_outer(st, T, B, m.Tm, At, r->len, m.ds, m.hc, HD, p.scaled ? m.g : nullptr, m.dkeys));
  return 0;
}

}  // namespace avsr
