// Persistent fused LSTM layer (forward and backward) on clusters of FOUR CTAs, H = 256: the cluster-of-4
// counterpart of lstm_persist.cu (same reference semantics: tf.nn.dynamic_rnn over LSTMCell(cell_clip=1,
// forget_bias=1), cells.py:14-18, encoder.py:80, SURVEY.md A.1/A.2, and its tf.gradients).
//
// Why four: a B200 keeps 33 clusters of 4 CTAs resident but only 15 clusters of 8, so 256 utterances run as ONE
// wave of 32 clusters x 8 utterances on 128 SMs.  The step is latency bound (exchange -> product -> gate math), and
// every term of that chain shrinks with the slice: 8 activations per thread, 4 KB of h per CTA through DSMEM.
//
// A CTA owns 64 hidden units = 256 gate rows = two 128-row M tiles.  The recurrent kernel Wh enters tcgen05.mma as
// the A operand FROM TENSOR MEMORY (fp16: tf32-rounded values are exactly representable, so the operands equal the
// tf32 operands of the rest of the path; lane = row, 32-bit column c = K elements 2c, 2c+1; 128 columns per tile,
// checked by tools/micro/ts_mma_test.cu): the weights never occupy shared memory and the product does not wait for an
// A tile to be read from it.  The backward kernel holds its two 128 x 256 tiles of Wh^T the same way; dz enters as
// fp16 scaled by a power of two (`grad_scale`, undone on the accumulators).
#include <cuda_fp16.h>
#include <stdlib.h>

#include "../../include/avsr_b200.h"
#include "common.cuh"

namespace avsr {
namespace lp4 {

constexpr int CL = 4;
constexpr int H = 256;
constexpr int UPC = H / CL;             // 64 hidden units per CTA
constexpr int NB = 8;                   // utterances per cluster
constexpr int NP = 16;                  // N of the products (M = 128 needs N % 16 == 0): 8 utterances + 8 zero rows
constexpr int KBH = H / 64;             // 4 K-blocks of 64 halves (128 B)
constexpr int THREADS = 256;
constexpr int OP_BYTES = KBH * NP * 128;  // one h operand buffer (8 KB)

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
__device__ __forceinline__ void st_async_f(uint32_t addr, uint32_t mbar, float a) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f32 [%0], %2, [%1];" ::"r"(addr), "r"(mbar), "f"(a)
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
      "LP4_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra LP4_DONE;\n\t"
      "bra LP4_WAIT;\n\t"
      "LP4_DONE:\n\t"
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
__device__ __forceinline__ uint64_t desc_at(uint64_t d, uint32_t byte_off) { return d + (uint64_t)(byte_off >> 4); }
// A from tensor memory, B from a shared-memory descriptor
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
// One lane of a converged warp (elect.sync).  With the issuing lane elected and every operand derived from provably
// warp-uniform values (warp index / tensor-memory base passed through __shfl_sync) ptxas keeps the descriptors in uniform
// registers and emits the tcgen05.mma back to back; with `lane == 0` and per-thread operands every MMA sat in an
// ELECT / 4 x R2UR / branch loop of ~290 clocks (tools/ap4d_trace.py).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0u;
}
__device__ __forceinline__ uint32_t warp_uniform(uint32_t v) { return __shfl_sync(0xffffffffu, v, 0); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
      "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
      "r"(r[31])
      : "memory");
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
// byte offset of half element (row, k) in a K-major SWIZZLE_128B operand with 64-half K blocks of `rows` rows
__device__ __forceinline__ uint32_t sw128h_off(int rows, int row, int k) {
  const int kb = k >> 6, kk = k & 63;
  return (uint32_t)(kb * rows * 128 + row * 128 + ((((kk >> 3) ^ (row & 7)) << 4)) + ((kk & 7) << 1));
}
// instruction descriptor: D = f32, A = B = f16, both K-major, N = NP, M = 128
constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(NP >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

// =====================================================================================================
// forward
// =====================================================================================================
struct FwdParams {
  int T, B;
  const int* len;
  float* gates;       // [T,B,4H] in: x-projection + bias; out: activations
  const float* Wrec;  // [H,4H] (tf32-rounded operand copy)
  const float* c0;    // [B,H] or null
  float* S;           // [(T+1),B,H] ; S[0] = h0 (caller)
  float* craw;        // [T,B,H]
  float* out;         // [T,B,H]
  float* cT;          // [B,H] or null
  float* hT;          // [B,H] or null
  long long* dbg;     // AVSR_LP_DEBUG: clock samples [64 steps][8] of CTA 0 (thread 0: slots 0-5, thread 128+32: 6-7)
  // DropoutWrapper state / output masks (AvsrRnnSeq.rng), used by the DROP instantiation only
  const uint32_t* rng;
  uint32_t stream, thr_state, thr_out;
  float inv_state, inv_out;
};
#define LP4_STAMP(slot)                                                                        \
  do {                                                                                         \
    if (p.dbg && blockIdx.x == 0 && tid == 0 && t < 64) p.dbg[t * 8 + (slot)] = clock64();  \
  } while (0)

constexpr size_t FWD_SMEM = (size_t)2 * OP_BYTES + 4 * NB * UPC * 4 + 64 + 1024;

// DROP: the cell sits in a DropoutWrapper (cells.py:46-54): the emitted h carries the output mask, the recurrent h
// (operand of the next step, S rows, final state) the state mask; both are regenerated from the counter-based
// generator, computed one step ahead (off the exchange -> product -> gate-math chain).
template <bool DROP>
__global__ void __launch_bounds__(THREADS, 1) lstm_persist4_fwd_kernel(const FwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sOp = base;                         // two h operand buffers, NP rows (rows >= NB stay zero)
  const uint32_t sAct = sOp + 2 * OP_BYTES;          // [4][NB][UPC] floats
  const uint32_t sBar = sAct + 4 * NB * UPC * 4;     // [0,1] mma_done[tile] [2,3] h_full[buf]
  const uint32_t sTmem = sBar + 32;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  float* act = reinterpret_cast<float*>(gen + (sAct - base));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int b0 = cluster_id_x() * NB;
  const int T = p.T, B = p.B;

  if (tid == 0) {
    mbar_init(sBar, 4);       // mma_done[tile 0]: one commit per issuing warp
    mbar_init(sBar + 8, 4);   // mma_done[tile 1]
    mbar_init(sBar + 16, 1);  // h_full[buffer 0]
    mbar_init(sBar + 24, 1);  // h_full[buffer 1]
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // tensor memory (all 512 columns): [0, 128) accumulators: tile m, K quarter j at column 16 (4 m + j); [256, 512)
  // the two 128 x 256 tiles of Wh^T (tile m: gate rows of the CTA's units 32*m .. 32*m+31), 128 columns each
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(sTmem) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // initial h (S[0]) of this slice into operand buffer 0; everything else (padding rows, buffer 1) zero
  for (int i = tid; i < 2 * OP_BYTES / 4; i += THREADS) reinterpret_cast<uint32_t*>(gen + (sOp - base))[i] = 0u;
  __syncthreads();
  for (int i = tid; i < NB * H; i += THREADS) {
    const int b = i / H, k = i - b * H;
    const float v = (b0 + b < B) ? p.S[(size_t)(b0 + b) * H + k] : 0.0f;
    *reinterpret_cast<__half*>(gen + (sOp - base) + sw128h_off(NP, b, k)) = __float2half_rn(v);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(sTmem));
  const uint32_t tW = tmem_base + 256;
  {
    // lane r = gate*32 + u of tile m <-> Wrec[k][gate*H + 64*rank + 32*m + u]; column c holds K elements 2c, 2c+1.
    // warp w fills lane quarter (w & 3) = gate of tile (w >> 2)
    const int q = warp & 3, mt = warp >> 2;
    const float* col = p.Wrec + q * H + UPC * rank + 32 * mt + lane;
#pragma unroll 1
    for (int c0 = 0; c0 < 128; c0 += 32) {
      uint32_t r[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        const int k = 2 * (c0 + c);
        r[c] = pack_h2(col[(size_t)k * 4 * H], col[(size_t)(k + 1) * 4 * H]);
      }
      tmem_st32(tW + 128 * mt + c0 + ((uint32_t)(32 * q) << 16), r);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  cluster_sync_all();

  // gate-math role: warp <-> (gate g, tile m): the gate rows of units 32*m + lane for all NB utterances
  const int g = warp & 3, m = warp >> 2;
  const int unit_g = UPC * rank + 32 * m + lane;
  const uint32_t mbar_m = sBar + 8 * m;
  // product-issue role.  Issuing a tcgen05.mma costs ~80 clocks of one thread (descriptor -> uniform registers), far
  // more than the 128 x 16 x 16 product takes, and the 16 K steps of a tile sit on the critical path of the step.  So
  // EVERY warp issues: lane 0 of warp (m, j) issues the K quarter j (= K block j, four steps) of tile m into its own
  // accumulator columns, and the gate math adds the four partial accumulators of its tile.
  const int warp_u = (int)warp_uniform((uint32_t)warp);  // (uniform operands + elected lane: see elect_one)
  const uint32_t tmem_u = warp_uniform(tmem_base);
  const int jq = warp_u & 3, m_u = warp_u >> 2;
  const uint32_t acc_col = tmem_u + (4 * m_u + jq) * NP;
  const uint32_t tW_u = tmem_u + 256, mbar_mu = sBar + 8 * m_u;
  const uint64_t dOp[2] = {make_desc_k128(sOp), make_desc_k128(sOp + OP_BYTES)};
  auto issue_quarter = [&](uint32_t nbuf) {
    if (elect_one()) {
#pragma unroll
      for (int k4 = 0; k4 < 4; ++k4)
        umma_ts(acc_col, tW_u + 128 * m_u + (jq * 4 + k4) * 8, desc_at(dOp[nbuf], jq * (NP * 128) + k4 * 32), IDESC, k4 ? 1u : 0u);
      umma_commit(mbar_mu);
    }
    __syncwarp();
  };
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
      h_state[e] = (b < B) ? p.S[(size_t)b * H + u] : 0.0f;
    }
  }
  float f_state[4] = {1.0f, 1.0f, 1.0f, 1.0f}, f_out[4] = {1.0f, 1.0f, 1.0f, 1.0f};
  uint32_t rng_seed = 0u, rng_step = 0u;
  if (DROP) {
    rng_seed = p.rng[0];
    rng_step = p.rng[1];
  }
  auto drop_factors = [&](int t) {  // masks of step t for this thread's 4 units of utterance b0 + bq
    if (DROP && comb) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const uint32_t idx = (uint32_t)((b0 + bq) * H + UPC * (int)rank + 4 * uq + e);
        f_state[e] = (p.thr_state == 0u || avsr_rand_u32(rng_seed, rng_step, p.stream + 1u, (uint32_t)t, idx) < p.thr_state)
                         ? p.inv_state : 0.0f;
        f_out[e] = (p.thr_out == 0u || avsr_rand_u32(rng_seed, rng_step, p.stream + 2u, (uint32_t)t, idx) < p.thr_out)
                       ? p.inv_out : 0.0f;
      }
    }
  };
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
  // product of step 0: h_0 is in buffer 0
  if (T > 0) issue_quarter(0);
  // x-projection of step 1, two steps ahead of its use: the stream of gate rows comes from HBM (~1000+ clocks), more
  // than what is left of a step after the prefetch is issued
  float gx1[NB];
  {
    const float* grow1 = p.gates + ((size_t)B + b0) * 4 * H + g * H + unit_g;
#pragma unroll
    for (int b = 0; b < NB; ++b) gx1[b] = (1 < T && 1 < len_a[b]) ? grow1[(size_t)b * 4 * H] : 0.0f;
  }

  for (int t = 0; t < T; ++t) {
    float* grow = p.gates + ((size_t)t * B + b0) * 4 * H + g * H + unit_g;
    uint32_t r[8];
    LP4_STAMP(0);
    mbar_wait(mbar_m, t & 1);  // recurrent product of this step (tile m)
    LP4_STAMP(1);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t r1[8], r2[8], r3[8];
    tmem_ld8(tmem_base + ((uint32_t)(32 * g) << 16) + (4 * m + 0) * NP, r);
    tmem_ld8(tmem_base + ((uint32_t)(32 * g) << 16) + (4 * m + 1) * NP, r1);
    tmem_ld8(tmem_base + ((uint32_t)(32 * g) << 16) + (4 * m + 2) * NP, r2);
    tmem_ld8(tmem_base + ((uint32_t)(32 * g) << 16) + (4 * m + 3) * NP, r3);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    float av[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) {  // i, f, o: sigmoid (forget bias 1); j: tanh
      const float z = ((__uint_as_float(r[b]) + __uint_as_float(r1[b])) + (__uint_as_float(r2[b]) + __uint_as_float(r3[b]))) + gx[b];
      float a;
      if (g == 1) a = tanhf_acc(z);
      else a = sigmoidf_acc(g == 2 ? z + 1.0f : z);
      av[b] = a;
      act[(g * NB + b) * UPC + 32 * m + lane] = a;
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    LP4_STAMP(2);
    const uint32_t nb = (t + 1) & 1;
    const uint32_t hbar_n = sBar + 16 + 8 * nb;
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
          const float c = fminf(fmaxf(cr[e], -1.0f), 1.0f);  // cell_clip = 1.0 (cells.py:16)
          const float h = vo[e] * tanhf_acc(c);
          c_state[e] = c;
          if (DROP) {
            ov[e] = h * f_out[e];
            h_state[e] = tf32_rn(h * f_state[e]);
          } else {
            ov[e] = h;
            h_state[e] = tf32_rn(h);  // the recurrent operand / next layer's operand
          }
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
      if (t + 1 < T) {
        // all-gather of h_t (fp16): operand of step t+1 in every CTA of the cluster
        const uint32_t off = sw128h_off(NP, bq, UPC * (int)rank + 4 * uq);
        const uint32_t u01 = pack_h2(hv[0], hv[1]), u23 = pack_h2(hv[2], hv[3]);
        const uint32_t dbuf = sOp + nb * OP_BYTES + off;
#pragma unroll
        for (uint32_t dst = 0; dst < (uint32_t)CL; ++dst) st_async_v2(mapa(dbuf, dst), mapa(hbar_n, dst), u01, u23);
      }
    }
    LP4_STAMP(3);
    // HBM side of this step and the x-projection of step t+2, issued before the wait for the exchange
#pragma unroll
    for (int b = 0; b < NB; ++b)
      if (t < len_a[b]) grow[(size_t)b * 4 * H] = av[b];  // activations, kept for the backward pass
    if (comb && b0 + bq < B) {
      const size_t o = ((size_t)t * B + b0 + bq) * H + UPC * rank + 4 * uq;
      *reinterpret_cast<float4*>(p.craw + o) = make_float4(cr[0], cr[1], cr[2], cr[3]);
      *reinterpret_cast<float4*>(p.out + o) = make_float4(ov[0], ov[1], ov[2], ov[3]);
      *reinterpret_cast<float4*>(p.S + o + (size_t)B * H) = make_float4(hv[0], hv[1], hv[2], hv[3]);
    }
    if (DROP && t + 1 < T) drop_factors(t + 1);
#pragma unroll
    for (int b = 0; b < NB; ++b) gx[b] = gx1[b];
    if (t + 2 < T) {
      const float* gnext = grow + (size_t)2 * B * 4 * H;
#pragma unroll
      for (int b = 0; b < NB; ++b) gx1[b] = (t + 2 < len_a[b]) ? gnext[(size_t)b * 4 * H] : 0.0f;
    }
    LP4_STAMP(4);
    if (t + 1 < T) {
      // product of step t+1, as soon as every CTA's h_t slice has landed.  Every warp has read the accumulators of
      // step t before its activations reached the barrier above.
      if (tid == 0) mbar_expect_tx(hbar_n, NB * H * 2);
      mbar_wait(hbar_n, (t >> 1) & 1);
      LP4_STAMP(5);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      issue_quarter(nb);
    }
    LP4_STAMP(6);
  }
  if (comb && b0 + bq < B) {  // final states
    const size_t o = (size_t)(b0 + bq) * H + UPC * rank + 4 * uq;
    if (p.cT) *reinterpret_cast<float4*>(p.cT + o) = make_float4(c_state[0], c_state[1], c_state[2], c_state[3]);
    if (p.hT) *reinterpret_cast<float4*>(p.hT + o) = make_float4(h_state[0], h_state[1], h_state[2], h_state[3]);
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  cluster_sync_all();  // no CTA exits while a peer could still address its shared memory
}

// =====================================================================================================
// backward:  dh_{t-1} = dz_t Wh^T.  A CTA forms dz_t for its own 64 units, multiplies by its Wh[:, own gate
// columns] (K split over the cluster) and the partial dh_{t-1} [H, NB] tiles are reduce-scattered to the owners of
// the out-units with st.async, signalled through mbarriers.
// =====================================================================================================
struct BwdParams {
  int T, B;
  float grad_scale, inv_grad_scale;
  const int* len;
  const float* gates;  // [T,B,4H] activations i,j,f,o
  const float* Wrec;   // [H,4H]
  const float* c0;     // [B,H] or null
  const float* craw;   // [T,B,H]
  const float* dout;   // [T,B,H] or null
  const float* dcT;    // [B,H] or null
  const float* dhT;    // [B,H] or null
  float* dZ;           // [T,B,4H]
  float* dc0;          // [B,H] or null
  float* dh0;          // [B,H] or null
  float* dbias;        // [4H] or null: += column sums of dZ
  const uint32_t* rng;  // DropoutWrapper state / output masks (DROP instantiation only)
  uint32_t stream, thr_state, thr_out;
  float inv_state, inv_out;
};

constexpr int BW_DZ_BYTES = 4 * NP * 128;            // B operand: 4 K-blocks (gates) x [NP rows x 64 units]
constexpr int REDH_FLOATS = CL * NB * UPC;           // one parity of the reduce buffer [src][b][u]
constexpr size_t BWD_SMEM = (size_t)BW_DZ_BYTES + 2 * REDH_FLOATS * 4 + 64 + 1024;

template <bool DROP>
__global__ void __launch_bounds__(THREADS, 1) lstm_persist4_bwd_kernel(const BwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sDz = base;
  const uint32_t sRed = sDz + BW_DZ_BYTES;           // two parities
  const uint32_t sBar = sRed + 2 * REDH_FLOATS * 4;  // [0] mma_done [1] dz_ready [2],[3] red_full[parity]
  const uint32_t sTmem = sBar + 32;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  float* red = reinterpret_cast<float*>(gen + (sRed - base));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int b0 = cluster_id_x() * NB;
  const int T = p.T, B = p.B;

  if (tid == 0) {
    mbar_init(sBar, 8);  // one commit per issuing warp
    mbar_init(sBar + 8, THREADS);
    mbar_init(sBar + 16, 1);
    mbar_init(sBar + 24, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // tensor memory (all 512 columns): [0, 128) accumulators: 128-row tile mt of dh, K quarter j (= gate j) at column
  // 16 (4 mt + j); [256, 512) the two tiles of A[n][k = g*64 + u] = Wrec[n][g*H + 64*rank + u] (tile = n >> 7)
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(sTmem) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = tid; i < BW_DZ_BYTES / 4; i += THREADS) reinterpret_cast<uint32_t*>(gen + (sDz - base))[i] = 0u;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(sTmem));
  const uint32_t tA = tmem_base + 256;
  const int warp_u = (int)warp_uniform((uint32_t)warp);  // (uniform operands + elected lane: see elect_one)
  const uint32_t tmem_u = warp_uniform(tmem_base), tA_u = tmem_u + 256;
  {
    const int q = warp & 3, tt = warp >> 2;
    const float* row = p.Wrec + (size_t)(128 * tt + 32 * q + lane) * 4 * H + UPC * rank;
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

  const uint64_t dDz = make_desc_k128(sDz);
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
  // saved forward values of a step (activations, cell values, incoming gradient), loaded TWO iterations ahead of their
  // use: they stream from HBM and one iteration is shorter than that latency
  struct StepVals {
    float gi[PB], gj[PB], gf[PB], go[PB], crw[PB], cpv[PB], dov[PB];
  };
  StepVals cur, nxt;
#pragma unroll
  for (int j = 0; j < PB; ++j) {
    cur.gi[j] = cur.gj[j] = cur.gf[j] = cur.go[j] = cur.crw[j] = cur.cpv[j] = cur.dov[j] = 0.0f;
    nxt = cur;
  }
  auto load_step = [&](StepVals& v, int t) {
#pragma unroll
    for (int j = 0; j < PB; ++j) {
      const int b = b0 + (warp >> 1) * PB + j;
      if (t >= 0 && t < len_t[j]) {
        const float* g = p.gates + ((size_t)t * B + b) * 4 * H + unit;
        v.gi[j] = g[0]; v.gj[j] = g[H]; v.gf[j] = g[2 * H]; v.go[j] = g[3 * H];
        const size_t o = ((size_t)t * B + b) * H + unit;
        v.crw[j] = p.craw[o];
        v.cpv[j] = t > 0 ? p.craw[o - (size_t)B * H] : (p.c0 ? p.c0[(size_t)b * H + unit] : 0.0f);
        v.dov[j] = p.dout ? p.dout[o] : 0.0f;
      }
    }
  };
  // reduce-scatter role after the product: warps 0-3 forward tile 0, warps 4-7 tile 1; lane quarter q = warp & 3
  const int q = warp & 3, mt_push = warp >> 2;

  uint32_t rng_seed = 0u, rng_step = 0u;
  if (DROP) {
    rng_seed = p.rng[0];
    rng_step = p.rng[1];
  }
  float bsum[4] = {0.0f, 0.0f, 0.0f, 0.0f};  // bias gradient of this thread's unit: sum of dz over its utterances / steps
  load_step(cur, T - 1);
  load_step(nxt, T - 2);
  for (int it = 0; it < T; ++it) {
    const int t = T - 1 - it;
    // recurrent dh of this step: partial sums pushed by all CTAs during the previous iteration + carry
    const float* rbuf = red + (it & 1) * REDH_FLOATS;
    float f_state[PB], f_out[PB];  // DropoutWrapper masks of step t (regenerated; before the wait: independent of it)
#pragma unroll
    for (int j = 0; j < PB; ++j) {
      f_state[j] = f_out[j] = 1.0f;
      if (DROP) {
        const uint32_t idx = (uint32_t)((b0 + (warp >> 1) * PB + j) * H + unit);
        if (p.thr_state != 0u)
          f_state[j] = avsr_rand_u32(rng_seed, rng_step, p.stream + 1u, (uint32_t)t, idx) < p.thr_state ? p.inv_state : 0.0f;
        if (p.thr_out != 0u)
          f_out[j] = avsr_rand_u32(rng_seed, rng_step, p.stream + 2u, (uint32_t)t, idx) < p.thr_out ? p.inv_out : 0.0f;
      }
    }
    if (it > 0) {
      const uint32_t bar = sBar + 16 + 8 * (it & 1);
      if (tid == 0) mbar_expect_tx(bar, REDH_FLOATS * 4);
      mbar_wait(bar, ((it - 1) >> 1) & 1);
    }
    float dz[4][PB];
#pragma unroll
    for (int j = 0; j < PB; ++j) {
      const int bl = (warp >> 1) * PB + j;
      float dh = dh_carry[j];
      if (it > 0) {
#pragma unroll
        for (int src = 0; src < CL; ++src) dh += rbuf[(src * NB + bl) * UPC + ul];
      }
      if (t < len_t[j]) {
        if (DROP) dh = dh * f_state[j] + cur.dov[j] * f_out[j];  // state-dropped h recurs, output-dropped h is emitted
        else dh += cur.dov[j];
        const float c = fminf(fmaxf(cur.crw[j], -1.0f), 1.0f);
        const float tc = tanhf_acc(c);
        const float cp = t > 0 ? fminf(fmaxf(cur.cpv[j], -1.0f), 1.0f) : cur.cpv[j];
        const float dct = dc[j] + dh * cur.go[j] * (1.0f - tc * tc);
        const float dcr = (cur.crw[j] >= -1.0f && cur.crw[j] <= 1.0f) ? dct : 0.0f;  // gradient of the cell clip
        dz[0][j] = tf32_rn(dcr * cur.gj[j] * cur.gi[j] * (1.0f - cur.gi[j]));
        dz[1][j] = tf32_rn(dcr * cur.gi[j] * (1.0f - cur.gj[j] * cur.gj[j]));
        dz[2][j] = tf32_rn(dcr * cp * cur.gf[j] * (1.0f - cur.gf[j]));
        dz[3][j] = tf32_rn(dh * tc * cur.go[j] * (1.0f - cur.go[j]));
        dc[j] = dcr * cur.gf[j];
        dh_carry[j] = 0.0f;
        if (b0 + bl < B) {
#pragma unroll
          for (int g = 0; g < 4; ++g) bsum[g] += dz[g][j];
        }
      } else {
        dz[0][j] = dz[1][j] = dz[2][j] = dz[3][j] = 0.0f;
        dh_carry[j] = dh;  // state (and its gradient) is carried through masked steps
      }
#pragma unroll
      for (int g = 0; g < 4; ++g)
        *reinterpret_cast<__half*>(gen + (sDz - base) + sw128h_off(NP, bl, g * 64 + ul)) = __float2half_rn(dz[g][j] * p.grad_scale);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_arrive(sBar + 8);
    cur = nxt;
    load_step(nxt, t - 2);
    {
      // partial dh (256 rows) x NB from this CTA's 256 gate columns, once every warp's dz is in shared memory.  Every
      // warp issues (see the forward kernel): lane 0 of warp (mt, j) the K quarter j (gate j) of the 128-row tile mt
      // into its own accumulator columns; the reduce-scatter below adds the four partial accumulators.
      const int mt = warp_u >> 2, jq = warp_u & 3;
      mbar_wait(sBar + 8, it & 1);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4)
          umma_ts(tmem_u + (4 * mt + jq) * NP, tA_u + 128 * mt + (jq * 4 + k4) * 8, desc_at(dDz, jq * (NP * 128) + k4 * 32), IDESC,
                  k4 ? 1u : 0u);
        umma_commit(sBar);
      }
      __syncwarp();
    }
    // HBM side of this step while the product runs
#pragma unroll
    for (int j = 0; j < PB; ++j) {
      const int b = b0 + (warp >> 1) * PB + j;
      if (b < B) {
        float* o = p.dZ + ((size_t)t * B + b) * 4 * H + unit;
        o[0] = dz[0][j]; o[H] = dz[1][j]; o[2 * H] = dz[2][j]; o[3 * H] = dz[3][j];
      }
    }
    // partial dh_{t-1}[k, b] for the 128 out-units of tile mt_push -> pushed to the owners of k
    mbar_wait(sBar, it & 1);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    {
      uint32_t r[8], r1[8], r2[8], r3[8];
      tmem_ld8(tmem_base + ((uint32_t)(32 * q) << 16) + (4 * mt_push + 0) * NP, r);
      tmem_ld8(tmem_base + ((uint32_t)(32 * q) << 16) + (4 * mt_push + 1) * NP, r1);
      tmem_ld8(tmem_base + ((uint32_t)(32 * q) << 16) + (4 * mt_push + 2) * NP, r2);
      tmem_ld8(tmem_base + ((uint32_t)(32 * q) << 16) + (4 * mt_push + 3) * NP, r3);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int c = 0; c < NB; ++c)
        r[c] = __float_as_uint((__uint_as_float(r[c]) + __uint_as_float(r1[c])) + (__uint_as_float(r2[c]) + __uint_as_float(r3[c])));
      // global unit 128*mt + 32*q + lane -> owner CTA 2*mt + (q >> 1), local unit 32*(q & 1) + lane
      const uint32_t dst = (uint32_t)(2 * mt_push + (q >> 1));
      const uint32_t rnext = sRed + ((it + 1) & 1) * REDH_FLOATS * 4;
      const uint32_t a0 = mapa(rnext + (uint32_t)((rank * NB) * UPC + 32 * (q & 1) + lane) * 4, dst);
      const uint32_t bar = mapa(sBar + 16 + 8 * ((it + 1) & 1), dst);
#pragma unroll
      for (int c = 0; c < NB; ++c) st_async_f(a0 + c * UPC * 4, bar, __uint_as_float(r[c]) * p.inv_grad_scale);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  // gradient wrt the initial state: partials of the last iteration
  if (T > 0) {
    const uint32_t bar = sBar + 16 + 8 * (T & 1);
    if (tid == 0) mbar_expect_tx(bar, REDH_FLOATS * 4);
    mbar_wait(bar, ((T - 1) >> 1) & 1);
  }
  const float* rbuf = red + (T & 1) * REDH_FLOATS;
#pragma unroll
  for (int j = 0; j < PB; ++j) {
    const int bl = (warp >> 1) * PB + j;
    const int b = b0 + bl;
    float dh = dh_carry[j];
    if (T > 0) {
#pragma unroll
      for (int src = 0; src < CL; ++src) dh += rbuf[(src * NB + bl) * UPC + ul];
    }
    if (b < B) {
      if (p.dh0) p.dh0[(size_t)b * H + unit] = dh;
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

template <typename Kern, typename P>
static int launch_cluster(cudaStream_t st, Kern kern, int B, size_t smem, const P& p, int klass) {
  AVSR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(cdiv(B, NB) * CL);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CL;
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

}  // namespace lp4

// Returns -1 if this layer shape is not handled (H != 256, or AVSR_LP_CLUSTER=8 selects lstm_persist.cu).
static bool lp4_enabled() {
  const char* e = getenv("AVSR_LP_CLUSTER");
  return !(e && atoi(e) == 8);
}

int lstm_persist4_fwd(cudaStream_t st, const AvsrRnnSeq* r) {
  if (r->n_mech != 0 || r->T <= 0 || r->H != lp4::H || !lp4_enabled()) return -1;
  lp4::FwdParams p;
  p.T = r->T; p.B = r->B;
  p.len = r->len; p.gates = r->gates; p.Wrec = r->Wrec; p.c0 = r->c0; p.S = r->S; p.craw = r->craw; p.out = r->out;
  p.cT = r->cT; p.hT = r->hT;
  p.dbg = nullptr;
  p.rng = r->rng;
  p.stream = r->drop_stream;
  p.thr_state = r->rng ? r->thr_state : 0u;
  p.thr_out = r->rng ? r->thr_out : 0u;
  p.inv_state = inv_keep_of(p.thr_state);
  p.inv_out = inv_keep_of(p.thr_out);
  const bool drop = (p.thr_state | p.thr_out) != 0u;
  if (getenv("AVSR_LP4_DEBUG")) {
    AVSR_CHECK_CUDA(cudaMalloc(&p.dbg, 64 * 8 * sizeof(long long)));
    AVSR_CHECK_CUDA(cudaMemset(p.dbg, 0, 64 * 8 * sizeof(long long)));
    AVSR_TRY(lp4::launch_cluster(st, lp4::lstm_persist4_fwd_kernel<false>, r->B, lp4::FWD_SMEM, p, AVSR_K_LSTM_FWD));
    AVSR_CHECK_CUDA(cudaStreamSynchronize(st));
    long long h[64 * 8];
    AVSR_CHECK_CUDA(cudaMemcpy(h, p.dbg, sizeof(h), cudaMemcpyDeviceToHost));
    cudaFree(p.dbg);
    const int n = r->T < 64 ? r->T : 64;
    const char* names[7] = {"loop-top", "wait MMA", "ld+act+bar", "combine+send h", "hbm st/ld", "wait h", "issue"};
    double acc[7] = {0};
    for (int t = 3; t < n - 1; ++t) {
      for (int k = 1; k < 7; ++k) acc[k] += (double)(h[t * 8 + k] - h[t * 8 + k - 1]);
      acc[0] += (double)(h[t * 8] - h[(t - 1) * 8 + 6]);
    }
    fprintf(stderr, "[lp4 fwd T=%d B=%d] clocks/step:", r->T, r->B);
    double tot = 0;
    for (int k = 0; k < 7; ++k) {
      fprintf(stderr, " %s=%.0f", names[k], acc[k] / (n - 4));
      tot += acc[k] / (n - 4);
    }
    fprintf(stderr, " total=%.0f\n", tot);
    return 0;
  }
  if (drop)
    return lp4::launch_cluster(st, lp4::lstm_persist4_fwd_kernel<true>, r->B, lp4::FWD_SMEM, p, AVSR_K_LSTM_FWD);
  return lp4::launch_cluster(st, lp4::lstm_persist4_fwd_kernel<false>, r->B, lp4::FWD_SMEM, p, AVSR_K_LSTM_FWD);
}

int lstm_persist4_bwd(cudaStream_t st, const AvsrRnnSeq* r) {
  if (r->n_mech != 0 || r->T <= 0 || r->H != lp4::H || !lp4_enabled()) return -1;
  lp4::BwdParams p;
  p.T = r->T; p.B = r->B;
  p.grad_scale = r->grad_scale > 0.0f ? r->grad_scale : 1.0f;
  p.inv_grad_scale = 1.0f / p.grad_scale;
  p.len = r->len; p.gates = r->gates; p.Wrec = r->Wrec; p.c0 = r->c0; p.craw = r->craw; p.dout = r->dout;
  p.dcT = r->dcT; p.dhT = r->dhT; p.dZ = r->dZ; p.dc0 = r->dc0; p.dh0 = r->dh0; p.dbias = r->dbias;
  p.rng = r->rng;
  p.stream = r->drop_stream;
  p.thr_state = r->rng ? r->thr_state : 0u;
  p.thr_out = r->rng ? r->thr_out : 0u;
  p.inv_state = inv_keep_of(p.thr_state);
  p.inv_out = inv_keep_of(p.thr_out);
  if (p.thr_state | p.thr_out)
    return lp4::launch_cluster(st, lp4::lstm_persist4_bwd_kernel<true>, r->B, lp4::BWD_SMEM, p, AVSR_K_LSTM_BWD);
  return lp4::launch_cluster(st, lp4::lstm_persist4_bwd_kernel<false>, r->B, lp4::BWD_SMEM, p, AVSR_K_LSTM_BWD);
}

}  // namespace avsr
