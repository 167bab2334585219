// Persistent fused LSTM layer (forward and backward): one thread-block CLUSTER owns a slice of NB
// utterances for ALL time steps.  The recurrent kernel Wh [H,4H] stays resident in shared memory, split
// by hidden-unit ranges across the CL CTAs of the cluster (32 units = 128 gate columns per CTA, the four
// gates of a unit co-located).
//
// forward, per step and CTA:
//   tcgen05.mma kind::tf32 (swap-AB)   D[128 gate rows, NB] = Wh_slice^T[128, H] . h_{t-1}^T[H, NB]
//   tcgen05.ld -> + x-projection (prefetched from HBM one step ahead) -> sigmoid/tanh  (warp <-> gate)
//   shared-memory exchange -> c/h update, length masking, cell clip
//   h_t (tf32-rounded) is pushed into the B-operand buffers of all CL CTAs with st.async (DSMEM); the
//   stores complete_tx on the DESTINATION's mbarrier, so a consumer only ever waits for the bytes it
//   needs - there is no cluster-wide barrier on the step path.
//
// backward: same decomposition.  A CTA forms dz_t for its own 32 units, multiplies by its resident
// Wh[:, own gate columns] (K split over the gate columns) and the partial dh_{t-1} [H, NB] tiles are
// reduce-scattered to the owners of the out-units with st.async, again signalled through mbarriers.
//
// Reference semantics: tf.nn.dynamic_rnn over LSTMCell(cell_clip=1, forget_bias=1), cells.py:14-18,
// encoder.py:80 and its tf.gradients.  Requires H = 32*CL with CL in {4, 8}.
#include <stdlib.h>

#include "../../include/avsr_b200.h"
#include "common.cuh"

namespace avsr {
namespace lp {

constexpr int GM_WARPS = 8;                   // gate-math warps
constexpr int THREADS = (GM_WARPS + 1) * 32;  // + 1 MMA-issue warp
// NB = utterances per cluster (16 or 32).  A B200 can keep only 15 clusters of 8 CTAs resident (GPC granularity),
// so batches above 240 use 32-utterance slices (8 clusters for B = 256) instead of a second wave of clusters.

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
// asynchronous 16-byte store into a peer CTA's shared memory; completes 16 tx-bytes on the peer's mbarrier
__device__ __forceinline__ void st_async_v4(uint32_t addr, uint32_t mbar, float a, float b, float c, float d) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%2, %3, %4, %5}, [%1];" ::"r"(addr),
               "r"(mbar), "f"(a), "f"(b), "f"(c), "f"(d)
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
      "LP_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra LP_DONE;\n\t"
      "bra LP_WAIT;\n\t"
      "LP_DONE:\n\t"
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
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
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
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
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
template <int N>
__device__ __forceinline__ void tmem_ldn(uint32_t taddr, uint32_t (&r)[N]);
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]);
template <>
__device__ __forceinline__ void tmem_ldn<8>(uint32_t taddr, uint32_t (&r)[8]) { tmem_ld8(taddr, r); }
template <>
__device__ __forceinline__ void tmem_ldn<32>(uint32_t taddr, uint32_t (&r)[32]) { tmem_ld32(taddr, r); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}

template <>
__device__ __forceinline__ void tmem_ldn<16>(uint32_t taddr, uint32_t (&r)[16]) { tmem_ld16(taddr, r); }

// shared-memory offset of element (row, k) of a K-major SWIZZLE_128B operand whose 32-float K-blocks
// hold `rows` rows each
__device__ __forceinline__ uint32_t sw128_off(int rows, int row, int k) {
  const int kb = k >> 5, kk = k & 31;
  return (uint32_t)(kb * rows * 128 + row * 128 + ((((kk >> 2) ^ (row & 7)) << 4)) + ((kk & 3) << 2));
}

// instruction descriptor: D = f32, A = B = tf32, both K-major, N = NB, M = 128
__host__ __device__ constexpr uint32_t idesc_for(int nb) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(nb >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// =====================================================================================================
// forward
// =====================================================================================================
struct FwdParams {
  int T, B, H;
  const int* len;
  float* gates;       // [T,B,4H] in: x-projection + bias; out: activations
  const float* Wrec;  // [H,4H] (tf32-rounded operand copy)
  const float* c0;    // [B,H] or null
  float* S;           // [(T+1),B,H] ; S[0] = h0 (caller)
  float* craw;        // [T,B,H]
  float* out;         // [T,B,H]
  float* cT;          // [B,H] or null
  float* hT;          // [B,H] or null
  long long* dbg;     // optional per-phase clock samples [64 steps][8] of CTA 0 (AVSR_LP_DEBUG=1), else null
};
__device__ __forceinline__ long long globaltimer_ns() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define LP_STAMP(slot)                                                                        \
  do {                                                                                        \
    if (p.dbg && blockIdx.x == 0 && tid == 0 && t < 64) {                                     \
      p.dbg[t * 8 + (slot)] = clock64();                                                      \
      if ((slot) == 0) p.dbg[t * 8 + 7] = globaltimer_ns();                                   \
    }                                                                                         \
  } while (0)

template <int CL, int NB>
__global__ void __launch_bounds__(THREADS, 1) lstm_persist_fwd_kernel(const FwdParams p) {
  constexpr int H = 32 * CL;
  constexpr int NA = NB / 2;                  // utterances per activation warp
  constexpr uint32_t IDESC = idesc_for(NB);
  constexpr int KB = H / 32;                  // 32-float K blocks
  constexpr int W_BYTES = KB * 128 * 128;     // A operand: 128 gate rows x H
  constexpr int HB_BYTES = KB * NB * 128;     // B operand: NB utterances x H
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sW = base;
  const uint32_t sH0 = sW + W_BYTES;          // two B-operand buffers (ping-pong over steps)
  const uint32_t sAct = sH0 + 2 * HB_BYTES;   // [4 gates][NB][32 units]
  const uint32_t sBar = sAct + 4 * NB * 32 * 4;  // [0] mma_done, [1],[2] h_full[buffer]
  const uint32_t sTmem = sBar + 24;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));  // generic pointer to `base`
  float* act = reinterpret_cast<float*>(gen + (sAct - base));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int b0 = cluster_id_x() * NB;         // first utterance of this cluster's slice
  const int T = p.T, B = p.B;

  // ---- one-time setup -------------------------------------------------------------
  if (tid == 0) {
    mbar_init(sBar, 1);
    mbar_init(sBar + 8, 1);
    mbar_init(sBar + 16, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == GM_WARPS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(sTmem) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // resident weights: row r = gate*32 + u  <->  Wrec[k][gate*H + 32*rank + u]
  for (int seg = warp; seg < H * 4; seg += THREADS / 32) {
    const int k = seg >> 2, g = seg & 3;
    const float w = p.Wrec[(size_t)k * 4 * H + g * H + 32 * rank + lane];
    const int r = g * 32 + lane;
    *reinterpret_cast<float*>(gen + (sW - base) + sw128_off(128, r, k)) = w;
  }
  // initial h (S[0]) of this slice into B-operand buffer 0
  for (int i = tid; i < NB * H; i += THREADS) {
    const int b = i / H, k = i - b * H;
    const float v = (b0 + b < B) ? p.S[(size_t)(b0 + b) * H + k] : 0.0f;
    *reinterpret_cast<float*>(gen + (sH0 - base) + sw128_off(NB, b, k)) = v;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(sTmem));
  cluster_sync_all();  // every CTA of the cluster has its barriers / buffers initialised

  if (warp == GM_WARPS) {
    // ================= MMA issuer =================
    for (int t = 0; t < T; ++t) {
      const uint32_t hb = sH0 + (t & 1) * HB_BYTES;
      if (t > 0) {  // h_{t-1}: NB x H floats pushed by the CL CTAs of the cluster (st.async complete_tx)
        const uint32_t bar = sBar + 8 + 8 * (t & 1);
        if (lane == 0) mbar_expect_tx(bar, NB * H * 4);
        mbar_wait(bar, ((t - 1) >> 1) & 1);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (lane == 0) {
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            const uint64_t da = make_desc_k128(sW + kb * (128 * 128) + k4 * 32);
            const uint64_t db = make_desc_k128(hb + kb * (NB * 128) + k4 * 32);
            umma_tf32(tmem_base, da, db, IDESC, (kb | k4) ? 1u : 0u);
          }
        }
        umma_commit(sBar);
      }
      __syncwarp();
    }
  } else {
    // ================= gate math: warp w <-> gate (w & 3), utterances NA*(w >> 2) .. +NA-1 ============
    const int g = warp & 3, ch = warp >> 2;
    const int unit = 32 * rank + lane;     // global hidden unit of this lane's gate row
    // combine mapping (threads 0 .. 8*NB-1): utterance bq, units 4*uq .. 4*uq+3
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
        h_state[e] = (b < B) ? p.S[(size_t)b * H + u] : 0.0f;
      }
    }
    int len_a[NA];  // lengths of this warp's utterances (activation phase)
#pragma unroll
    for (int b = 0; b < NA; ++b) len_a[b] = (b0 + ch * NA + b < B) ? p.len[b0 + ch * NA + b] : 0;
    float gx[NA];   // x-projection; later steps are prefetched one step ahead
    {
      const float* grow0 = p.gates + ((size_t)b0 + ch * NA) * 4 * H + g * H + unit;
#pragma unroll
      for (int b = 0; b < NA; ++b) gx[b] = (0 < len_a[b]) ? grow0[(size_t)b * 4 * H] : 0.0f;
    }
    for (int t = 0; t < T; ++t) {
      float* grow = p.gates + ((size_t)t * B + b0 + ch * NA) * 4 * H + g * H + unit;
      LP_STAMP(0);
      mbar_wait(sBar, t & 1);  // recurrent product of this step
      LP_STAMP(1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t r[NA];
      tmem_ldn<NA>(tmem_base + ((uint32_t)(32 * g) << 16) + ch * NA, r);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      LP_STAMP(2);
      float av[NA];
#pragma unroll
      for (int b = 0; b < NA; ++b) {  // i, f, o: sigmoid (forget bias 1); j: tanh
        const float z = __uint_as_float(r[b]) + gx[b];
        float a;
        if (g == 1) a = tanhf_acc(z);
        else a = sigmoidf_acc(g == 2 ? z + 1.0f : z);
        av[b] = a;
        act[(g * NB + ch * NA + b) * 32 + lane] = a;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      LP_STAMP(3);
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
            const float c = fminf(fmaxf(cr[e], -1.0f), 1.0f);  // cell_clip = 1.0 (cells.py:16)
            const float h = vo[e] * tanhf_acc(c);
            c_state[e] = c;
            ov[e] = h;
            h_state[e] = tf32_rn(h);  // the recurrent operand / next layer's operand
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
          // B operand of step t+1 in every CTA: row = utterance bq, K block = rank, 16-byte chunk uq (swizzled)
          const uint32_t off = (uint32_t)(rank * (NB * 128) + bq * 128 + ((uq ^ (bq & 7)) << 4));
          const uint32_t dbuf = sH0 + ((t + 1) & 1) * HB_BYTES + off;
          const uint32_t dbar = sBar + 8 + 8 * ((t + 1) & 1);
#pragma unroll
          for (uint32_t dst = 0; dst < (uint32_t)CL; ++dst)
            st_async_v4(mapa(dbuf, dst), mapa(dbar, dst), hv[0], hv[1], hv[2], hv[3]);
        }
      }
      LP_STAMP(4);
      // HBM side of this step + x-projection of the next, off the recurrent critical path
#pragma unroll
      for (int b = 0; b < NA; ++b)
        if (t < len_a[b]) grow[(size_t)b * 4 * H] = av[b];  // activations, kept for the backward pass
      if (comb && b0 + bq < B) {
        const size_t o = ((size_t)t * B + b0 + bq) * H + 32 * rank + 4 * uq;
        *reinterpret_cast<float4*>(p.craw + o) = make_float4(cr[0], cr[1], cr[2], cr[3]);
        *reinterpret_cast<float4*>(p.out + o) = make_float4(ov[0], ov[1], ov[2], ov[3]);
        *reinterpret_cast<float4*>(p.S + o + (size_t)B * H) = make_float4(hv[0], hv[1], hv[2], hv[3]);
      }
      if (t + 1 < T) {
        const float* gnext = grow + (size_t)B * 4 * H;
#pragma unroll
        for (int b = 0; b < NA; ++b) gx[b] = (t + 1 < len_a[b]) ? gnext[(size_t)b * 4 * H] : 0.0f;
      }
      LP_STAMP(5);
    }
    if (comb && b0 + bq < B) {  // final states
      const size_t o = (size_t)(b0 + bq) * H + 32 * rank + 4 * uq;
      if (p.cT) *reinterpret_cast<float4*>(p.cT + o) = make_float4(c_state[0], c_state[1], c_state[2], c_state[3]);
      if (p.hT) *reinterpret_cast<float4*>(p.hT + o) = make_float4(h_state[0], h_state[1], h_state[2], h_state[3]);
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == GM_WARPS)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem_base) : "memory");
  cluster_sync_all();  // no CTA exits while a peer could still address its shared memory
}

template <int CL, int NB>
static int launch_fwd(cudaStream_t st, const FwdParams& p) {
  constexpr int H = 32 * CL;
  constexpr int KB = H / 32;
  const size_t smem = (size_t)KB * 128 * 128 + 2 * (size_t)KB * NB * 128 + 4 * NB * 32 * 4 + 64 + 1024;
  auto kern = lstm_persist_fwd_kernel<CL, NB>;
  static bool attr = false;
  if (!attr) {
    AVSR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(cdiv(p.B, NB) * CL);
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
  AVSR_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
  ++g_launch_count;
  return 0;
}

// =====================================================================================================
// backward
// =====================================================================================================
// reduce buffer [src][u][b]: element (src, u, b) with the 16-byte chunks of a row XOR-swizzled by u, so the
// reader (lanes = u, fixed b) is not a 32-way bank conflict while the writer keeps 16-byte stores
template <int NB>
__device__ __forceinline__ int red_off(int src, int u, int b) {
  return (src * 32 + u) * NB + ((((b >> 2) ^ (u & (NB / 4 - 1))) << 2) | (b & 3));
}
struct BwdParams {
  int T, B, H;
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
};

template <int CL, int NB>
__global__ void __launch_bounds__(THREADS, 1) lstm_persist_bwd_kernel(const BwdParams p) {
  constexpr int H = 32 * CL;
  constexpr uint32_t IDESC = idesc_for(NB);
  constexpr int MT = H / 128;                   // M tiles of the partial product (H out-units)
  constexpr int W_BYTES = 4 * H * 128;          // A operand: 4 gate blocks x [H rows x 32 cols]
  constexpr int DZ_BYTES = 4 * NB * 128;        // B operand: 4 gate blocks x [NB rows x 32 cols]
  constexpr int RED_FLOATS = CL * 32 * NB;      // one parity of the reduce buffer [src][u][b]
  constexpr int PB = NB / GM_WARPS;             // utterances per thread
  constexpr int TCOLS = (MT * NB) < 32 ? 32 : MT * NB;  // TMEM columns (power of two >= 32)
  constexpr int NH = NB / 2;                    // MT == 1: the two warps of a quadrant split the columns
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sW = base;
  const uint32_t sDz = sW + W_BYTES;
  const uint32_t sRed = sDz + DZ_BYTES;
  const uint32_t sBar = sRed + 2 * RED_FLOATS * 4;  // [0] mma_done  [1] dz_ready  [2],[3] red_full[parity]
  const uint32_t sTmem = sBar + 32;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  float* red = reinterpret_cast<float*>(gen + (sRed - base));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int b0 = cluster_id_x() * NB;
  const int T = p.T, B = p.B;

  if (tid == 0) {
    mbar_init(sBar, 1);
    mbar_init(sBar + 8, GM_WARPS * 32);
    mbar_init(sBar + 16, 1);
    mbar_init(sBar + 24, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == GM_WARPS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sTmem), "n"(TCOLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // resident weights: A[k][g*32 + u] = Wrec[k][g*H + 32*rank + u]   (K-major: gate block g, row k)
  for (int seg = warp; seg < H * 4; seg += THREADS / 32) {
    const int k = seg >> 2, g = seg & 3;
    const float w = p.Wrec[(size_t)k * 4 * H + g * H + 32 * rank + lane];
    const uint32_t off = (uint32_t)(g * (H * 128) + k * 128 + ((((lane >> 2) ^ (k & 7)) << 4)) + ((lane & 3) << 2));
    *reinterpret_cast<float*>(gen + (sW - base) + off) = w;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(sTmem));
  cluster_sync_all();

  if (warp == GM_WARPS) {
    // ================= MMA issuer =================
    for (int it = 0; it < T; ++it) {
      mbar_wait(sBar + 8, it & 1);  // dz of this step is in shared memory
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (lane == 0) {
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
              const uint64_t da = make_desc_k128(sW + g * (H * 128) + mt * (128 * 128) + k4 * 32);
              const uint64_t db = make_desc_k128(sDz + g * (NB * 128) + k4 * 32);
              umma_tf32(tmem_base + mt * NB, da, db, IDESC, (g | k4) ? 1u : 0u);
            }
          }
        }
        umma_commit(sBar);
      }
      __syncwarp();
    }
  } else {
    // ============ gate-gradient math: thread = (unit lane, utterances PB*warp .. PB*warp+PB-1) ==========
    const int unit = 32 * rank + lane;
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
          dov[j] = p.dout ? p.dout[o] : 0.0f;
        }
      }
    };
    // reduce-scatter role of this warp: TMEM tile / column range it forwards, and to which CTA
    const int q = warp & 3;
    const int mt = (MT == 2) ? (warp >> 2) : 0;
    const int c0col = (MT == 2) ? 0 : (warp >> 2) * NH;
    const uint32_t dst = (uint32_t)(mt * 4 + q);        // owner of out-units 128*mt + 32*q + lane
    load_step(T - 1);
    for (int it = 0; it < T; ++it) {
      const int t = T - 1 - it;
      // recurrent dh of this step: partial sums pushed by all CTAs during the previous iteration + carry
      const float* rbuf = red + (it & 1) * RED_FLOATS;
      if (it > 0) {
        const uint32_t bar = sBar + 16 + 8 * (it & 1);
        if (tid == 0) mbar_expect_tx(bar, RED_FLOATS * 4);
        mbar_wait(bar, ((it - 1) >> 1) & 1);
      }
      float dz[4][PB];
#pragma unroll
      for (int j = 0; j < PB; ++j) {
        float dh = dh_carry[j];
        if (it > 0) {
#pragma unroll
          for (int src = 0; src < CL; ++src) dh += rbuf[red_off<NB>(src, lane, warp * PB + j)];
        }
        if (t < len_t[j]) {
          dh += dov[j];
          const float c = fminf(fmaxf(crw[j], -1.0f), 1.0f);
          const float tc = tanhf_acc(c);
          const float cp = t > 0 ? fminf(fmaxf(cpv[j], -1.0f), 1.0f) : cpv[j];
          const float dct = dc[j] + dh * go[j] * (1.0f - tc * tc);
          const float dcr = (crw[j] >= -1.0f && crw[j] <= 1.0f) ? dct : 0.0f;  // gradient of the cell clip
          dz[0][j] = tf32_rn(dcr * gj[j] * gi[j] * (1.0f - gi[j]));
          dz[1][j] = tf32_rn(dcr * gi[j] * (1.0f - gj[j] * gj[j]));
          dz[2][j] = tf32_rn(dcr * cp * gf[j] * (1.0f - gf[j]));
          dz[3][j] = tf32_rn(dh * tc * go[j] * (1.0f - go[j]));
          dc[j] = dcr * gf[j];
          dh_carry[j] = 0.0f;
        } else {
          dz[0][j] = dz[1][j] = dz[2][j] = dz[3][j] = 0.0f;
          dh_carry[j] = dh;  // state (and its gradient) is carried through masked steps
        }
        // B operand: gate block g, row = utterance, element = unit lane (swizzled 16-byte chunks)
        const int bl = warp * PB + j;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const uint32_t off = (uint32_t)(g * (NB * 128) + bl * 128 + ((((lane >> 2) ^ (bl & 7)) << 4)) + ((lane & 3) << 2));
          *reinterpret_cast<float*>(gen + (sDz - base) + off) = dz[g][j];
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(sBar + 8);
      // HBM side of this step while the product runs
#pragma unroll
      for (int j = 0; j < PB; ++j) {
        const int b = b0 + warp * PB + j;
        if (b < B) {
          float* o = p.dZ + ((size_t)t * B + b) * 4 * H + unit;
          o[0] = dz[0][j]; o[H] = dz[1][j]; o[2 * H] = dz[2][j]; o[3 * H] = dz[3][j];
        }
      }
      load_step(t - 1);
      // partial dh_{t-1}[k, b] for all H out-units k -> pushed to the owners of k
      mbar_wait(sBar, it & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t rnext = sRed + ((it + 1) & 1) * RED_FLOATS * 4;
      const uint32_t rbar = mapa(sBar + 16 + 8 * ((it + 1) & 1), dst);
      if (MT == 2) {
        uint32_t r[NB];
        tmem_ldn<NB>(tmem_base + ((uint32_t)(32 * q) << 16) + mt * NB, r);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const uint32_t a0 = mapa(rnext + (uint32_t)((rank * 32 + lane) * NB) * 4, dst);
#pragma unroll
        for (int v = 0; v < NB / 4; ++v)
          st_async_v4(a0 + ((v ^ (lane & (NB / 4 - 1))) << 4), rbar, __uint_as_float(r[4 * v]), __uint_as_float(r[4 * v + 1]),
                      __uint_as_float(r[4 * v + 2]), __uint_as_float(r[4 * v + 3]));
      } else {
        uint32_t r[NH];
        tmem_ldn<NH>(tmem_base + ((uint32_t)(32 * q) << 16) + c0col, r);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const uint32_t a0 = mapa(rnext + (uint32_t)((rank * 32 + lane) * NB) * 4, dst);
#pragma unroll
        for (int v = 0; v < NH / 4; ++v)
          st_async_v4(a0 + (((c0col / 4 + v) ^ (lane & (NB / 4 - 1))) << 4), rbar, __uint_as_float(r[4 * v]), __uint_as_float(r[4 * v + 1]),
                      __uint_as_float(r[4 * v + 2]), __uint_as_float(r[4 * v + 3]));
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    // gradient wrt the initial state: partials of the last iteration
    if (T > 0) {
      const uint32_t bar = sBar + 16 + 8 * (T & 1);
      if (tid == 0) mbar_expect_tx(bar, RED_FLOATS * 4);
      mbar_wait(bar, ((T - 1) >> 1) & 1);
    }
    const float* rbuf = red + (T & 1) * RED_FLOATS;
#pragma unroll
    for (int j = 0; j < PB; ++j) {
      const int b = b0 + warp * PB + j;
      float dh = dh_carry[j];
      if (T > 0) {
#pragma unroll
        for (int src = 0; src < CL; ++src) dh += rbuf[red_off<NB>(src, lane, warp * PB + j)];
      }
      if (b < B) {
        if (p.dh0) p.dh0[(size_t)b * H + unit] = dh;
        if (p.dc0) p.dc0[(size_t)b * H + unit] = dc[j];
      }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == GM_WARPS)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TCOLS) : "memory");
  cluster_sync_all();
}

template <int CL, int NB>
static int launch_bwd(cudaStream_t st, const BwdParams& p) {
  constexpr int H = 32 * CL;
  const size_t smem = (size_t)4 * H * 128 + 4 * NB * 128 + 2 * (size_t)CL * 32 * NB * 4 + 64 + 1024;
  auto kern = lstm_persist_bwd_kernel<CL, NB>;
  static bool attr = false;
  if (!attr) {
    AVSR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(cdiv(p.B, NB) * CL);
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
  AVSR_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
  ++g_launch_count;
  return 0;
}

}  // namespace lp

// Debug aid (AVSR_LP_DEBUG=1, never during graph capture): prints the average clocks CTA 0 spends between
// the phases of a forward step.
static int lp_debug_fwd(cudaStream_t st, lp::FwdParams p, int H) {
  {  // how many clusters of 8 can be co-resident?  (GPC granularity can strand SMs)
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(128);
    cfg.blockDim = dim3(lp::THREADS);
    cfg.dynamicSmemBytes = (size_t)8 * 128 * 128 + 2 * (size_t)8 * 16 * 128 + 4 * 16 * 32 * 4 + 64 + 1024;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 8; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int ncl = -1;
    cudaFuncSetAttribute(lp::lstm_persist_fwd_kernel<8, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.dynamicSmemBytes);
    cudaError_t e = cudaOccupancyMaxActiveClusters(&ncl, lp::lstm_persist_fwd_kernel<8, 16>, &cfg);
    fprintf(stderr, "[lp] cudaOccupancyMaxActiveClusters(cluster 8, %zu B smem) = %d (%s)\n", cfg.dynamicSmemBytes, ncl,
            cudaGetErrorString(e));
    at[0].val.clusterDim.x = 4;
    e = cudaOccupancyMaxActiveClusters(&ncl, lp::lstm_persist_fwd_kernel<8, 16>, &cfg);
    fprintf(stderr, "[lp] ... cluster 4 = %d; ", ncl);
    at[0].val.clusterDim.x = 2;
    e = cudaOccupancyMaxActiveClusters(&ncl, lp::lstm_persist_fwd_kernel<8, 16>, &cfg);
    fprintf(stderr, "cluster 2 = %d\n", ncl);
  }
  long long* d = nullptr;
  AVSR_CHECK_CUDA(cudaMalloc(&d, 64 * 8 * sizeof(long long)));
  AVSR_CHECK_CUDA(cudaMemset(d, 0, 64 * 8 * sizeof(long long)));
  p.dbg = d;
  int rc = H == 256 ? (p.B > 240 ? lp::launch_fwd<8, 32>(st, p) : lp::launch_fwd<8, 16>(st, p)) : lp::launch_fwd<4, 16>(st, p);
  AVSR_CHECK_CUDA(cudaStreamSynchronize(st));
  long long h[64 * 8];
  AVSR_CHECK_CUDA(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
  cudaFree(d);
  const int n = p.T < 64 ? p.T : 64;
  const int NS = 6;
  double acc[NS] = {0};
  for (int t = 2; t < n; ++t) {
    for (int k = 1; k < NS; ++k) acc[k] += (double)(h[t * 8 + k] - h[t * 8 + k - 1]);
    acc[0] += (double)(h[t * 8] - h[(t - 1) * 8 + NS - 1]);
  }
  const char* names[NS] = {"loop-top", "wait MMA", "tmem ld", "act+bar", "combine+st.async", "hbm st/ld"};
  fprintf(stderr, "[lp fwd T=%d B=%d H=%d] clocks/step:", p.T, p.B, H);
  double tot = 0;
  for (int k = 0; k < NS; ++k) {
    fprintf(stderr, " %s=%.0f", names[k], acc[k] / (n - 2));
    tot += acc[k] / (n - 2);
  }
  const double ns = (double)(h[(n - 1) * 8 + 7] - h[2 * 8 + 7]) / (n - 3);
  fprintf(stderr, " total=%.0f  wall=%.0f ns/step -> SM clock %.0f MHz\n", tot, ns, tot / ns * 1e3);
  return rc;
}

int lstm_persist4_fwd(cudaStream_t st, const AvsrRnnSeq* r);  // lstm_persist4.cu (clusters of 4, H = 256)
int lstm_persist4_bwd(cudaStream_t st, const AvsrRnnSeq* r);

// Returns -1 if this layer shape is not handled by the persistent kernels.
int lstm_persist_fwd(cudaStream_t st, const AvsrRnnSeq* r) {
  if (r->n_mech != 0 || r->T <= 0) return -1;
  if (!getenv("AVSR_LP_DEBUG")) {
    const int rc = lstm_persist4_fwd(st, r);
    if (rc >= 0) return rc;
  }
  if (r->H != 128 && r->H != 256) return -1;
  lp::FwdParams p;
  p.T = r->T; p.B = r->B; p.H = r->H;
  p.len = r->len; p.gates = r->gates; p.Wrec = r->Wrec; p.c0 = r->c0; p.S = r->S; p.craw = r->craw; p.out = r->out;
  p.cT = r->cT; p.hT = r->hT;
  p.dbg = nullptr;
  if (getenv("AVSR_LP_DEBUG")) return lp_debug_fwd(st, p, r->H);
  // more than 15 clusters of 8 cannot be co-resident: wider slices instead of a second wave
  if (r->H == 256) return r->B > 240 ? lp::launch_fwd<8, 32>(st, p) : lp::launch_fwd<8, 16>(st, p);
  return r->B > 33 * 16 ? lp::launch_fwd<4, 32>(st, p) : lp::launch_fwd<4, 16>(st, p);
}

int lstm_persist_bwd(cudaStream_t st, const AvsrRnnSeq* r) {
  if (r->n_mech != 0 || r->T <= 0) return -1;
  if (r->H != 128 && r->H != 256) return -1;
  lp::BwdParams p;
  p.T = r->T; p.B = r->B; p.H = r->H;
  p.len = r->len; p.gates = r->gates; p.Wrec = r->Wrec; p.c0 = r->c0; p.craw = r->craw; p.dout = r->dout;
  p.dcT = r->dcT; p.dhT = r->dhT; p.dZ = r->dZ; p.dc0 = r->dc0; p.dh0 = r->dh0;
  if (r->H == 256) return r->B > 240 ? lp::launch_bwd<8, 32>(st, p) : lp::launch_bwd<8, 16>(st, p);
  return r->B > 33 * 16 ? lp::launch_bwd<4, 32>(st, p) : lp::launch_bwd<4, 16>(st, p);
}

}  // namespace avsr
