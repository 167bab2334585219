// Device helpers shared by the persistent AttentionWrapper(LSTMCell) kernels on clusters of four CTAs
// (attn_persist4.cu: attention layer folded into the recurrent matrix; attn_persist4d.cu: the DropoutWrapper
// variant with two dependent products per step).  See the headers of those files.
#pragma once
#include <cuda_fp16.h>
#include <stdlib.h>

#include "../../include/avsr_b200.h"
#include "common.cuh"

namespace avsr {
namespace ap4 {

constexpr int CL = 4;
constexpr int H = 256;
constexpr int DM = 256;
constexpr int KTOT = H + DM;            // 512
constexpr int KB = KTOT / 64;           // 8 K-blocks of 64 halves (128 B)
constexpr int UPC = H / CL;             // 64 hidden units per CTA
constexpr int NB = 8;                   // utterances per cluster
constexpr int NP = 16;                  // N of the products (M = 128 needs N % 16 == 0): 8 utterances + 8 zero rows
constexpr int NU = NB / CL;             // utterances whose attention a CTA owns
constexpr int THREADS = 256;
constexpr int W_BYTES = KB * 128 * 128; // one 128 x 512 fp16 tile
constexpr int OP_BYTES = KB * NP * 128; // one [h | ctx] operand buffer
constexpr int MAX_TM = 384;
constexpr int RIF = 8;                  // memory rows per batch of the attention sweeps

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
__device__ __forceinline__ void st_async_v4(uint32_t addr, uint32_t mbar, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%2, %3, %4, %5}, [%1];" ::"r"(addr),
               "r"(mbar), "r"(a), "r"(b), "r"(c), "r"(d)
               : "memory");
}
__device__ __forceinline__ void st_async_v4f(uint32_t addr, uint32_t mbar, float a, float b, float c, float d) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%2, %3, %4, %5}, [%1];" ::"r"(addr),
               "r"(mbar), "f"(a), "f"(b), "f"(c), "f"(d)
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
      "AP4_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra AP4_DONE;\n\t"
      "bra AP4_WAIT;\n\t"
      "AP4_DONE:\n\t"
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
// descriptor of the same operand `byte_off` bytes further (the start-address field counts 16-byte units)
__device__ __forceinline__ uint64_t desc_at(uint64_t d, uint32_t byte_off) { return d + (uint64_t)(byte_off >> 4); }
// A and B from shared-memory descriptors
__device__ __forceinline__ void umma_ss(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
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
// One lane of a converged warp (elect.sync): together with provably warp-uniform operands (warp index / tensor-memory
// base passed through __shfl_sync, see warp_uniform) ptxas keeps the descriptors in uniform registers and emits the
// tcgen05.mma directly - with `lane == 0` and per-thread operands every MMA sat in an ELECT / 4 x R2UR / branch loop
// (~290 clocks per instruction, measured with tools/ap4d_trace.py).
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
__device__ __forceinline__ float2 unpack_h2(uint32_t u) { return __half22float2(*reinterpret_cast<__half2*>(&u)); }

// Sums each of the 8 per-lane values v[0..7] over the 32 lanes with 9 shuffles.  Returns the complete sum of v[j] in
// the lanes with j == ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1).
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
// Packed fp32 FMA of sm_100 (FFMA2: two independent fp32 FMAs per issue slot).  The sweeps are issue bound as much as
// latency bound (ncu: issue slots 37 % busy with two warps per scheduler), so a row costs 8 conversions + 4 FFMA2
// instead of 8 + 8.
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  unsigned long long ua = *reinterpret_cast<unsigned long long*>(&a), ub = *reinterpret_cast<unsigned long long*>(&b),
                     uc = *reinterpret_cast<unsigned long long*>(&c), ud;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(ud) : "l"(ua), "l"(ub), "l"(uc));
  return *reinterpret_cast<float2*>(&ud);
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
// (forward sweeps only: in the backward kernels the packed form measured 2.6 % slower - register pairs)
__device__ __forceinline__ float dot8p(const uint4& r, const float (&q)[8]) {
  float2 s = fma2(unpack_h2(r.x), make_float2(q[0], q[1]), make_float2(0.0f, 0.0f));
  s = fma2(unpack_h2(r.y), make_float2(q[2], q[3]), s);
  s = fma2(unpack_h2(r.z), make_float2(q[4], q[5]), s);
  s = fma2(unpack_h2(r.w), make_float2(q[6], q[7]), s);
  return s.x + s.y;
}
__device__ __forceinline__ void axpy8p(float w, const uint4& r, float (&acc)[8]) {
  const float2 w2 = make_float2(w, w);
  const float2 a = fma2(w2, unpack_h2(r.x), make_float2(acc[0], acc[1]));
  const float2 b = fma2(w2, unpack_h2(r.y), make_float2(acc[2], acc[3]));
  const float2 c = fma2(w2, unpack_h2(r.z), make_float2(acc[4], acc[5]));
  const float2 d = fma2(w2, unpack_h2(r.w), make_float2(acc[6], acc[7]));
  acc[0] = a.x; acc[1] = a.y; acc[2] = b.x; acc[3] = b.y; acc[4] = c.x; acc[5] = c.y; acc[6] = d.x; acc[7] = d.y;
}
// lane that holds the total of row j of a batch after warp_reduce8 (its three neighbours hold copies)
__device__ __forceinline__ int reduce8_src(int j) { return ((j >> 2) & 1) * 16 + ((j >> 1) & 1) * 8 + (j & 1) * 4; }
// The SMALL variants of the kernels (memories of at most SMALL_TM rows, e.g. the 75 video frames of the cross-modal
// layer) keep the scores / alignments of a warp's rows in registers: SMALL_B batches of 8 rows per warp (rows
// w4 + 4*(8*i + j)), fully unrolled.  Longer memories (the decoder's 300 audio frames) use rolled loops and shared
// memory instead - unrolling 12 batches three times over blows the instruction cache.
constexpr int SMALL_TM = 96;
constexpr int SMALL_B = SMALL_TM / 32;

// byte offset of half element (row, k) in a K-major SWIZZLE_128B operand with 64-half K blocks of `rows` rows
__device__ __forceinline__ uint32_t sw128h_off(int rows, int row, int k) {
  const int kb = k >> 6, kk = k & 63;
  return (uint32_t)(kb * rows * 128 + row * 128 + ((((kk >> 3) ^ (row & 7)) << 4)) + ((kk & 7) << 1));
}
// instruction descriptor: D = f32, A = B = f16, both K-major, N = NP, M = 128
constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(NP >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

// ---------------------------------------------------------------------------------------------------------------------
// Attention of ONE utterance by its group of four warps (128 threads, named barrier `bar_id`): shared by the folded
// kernels (attn_persist4.cu) and the DropoutWrapper kernels (attn_persist4d.cu).  Rows tm = w4 + 4*i of the memory
// belong to warp w4; a lane holds dims 8*lane .. 8*lane+7 of a row (one 16-byte load).  The sweeps are software
// pipelined in half-batches of 4 rows (ra / rb): the caller requests the first batch (att_prefetch) before it waits
// for the query, the core requests the first batch of the second matrix before the softmax.
// ---------------------------------------------------------------------------------------------------------------------
struct AttRole {
  const __half* keys;
  const __half* values;
  int L, B, b_att, Tm, w4, gt, lane;
  float gs;            // Luong scale (1 when unscaled)
  uint32_t bar_id;
  float* sc;           // [MAX_TM] scores / alignments (shared memory)
  float* part;         // [4][DM] per-warp partial sums (shared memory)
  float* red;          // [8] reduction scratch (shared memory)
};
__device__ __forceinline__ void att_bar(uint32_t id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void att_prefetch(const AttRole& a, const __half* mat, uint4 (&ra)[4], uint4 (&rb)[4]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    ra[j] = ld_row(mat, a.w4 + 4 * j, a.L, a.B, a.b_att, a.lane);
    rb[j] = ld_row(mat, a.w4 + 16 + 4 * j, a.L, a.B, a.b_att, a.lane);
  }
}
__device__ __forceinline__ void unpack_q(const uint4& qraw, float (&q)[8]) {
  const float2 a = unpack_h2(qraw.x), b = unpack_h2(qraw.y), c = unpack_h2(qraw.z), d = unpack_h2(qraw.w);
  q[0] = a.x; q[1] = a.y; q[2] = b.x; q[3] = b.y; q[4] = c.x; q[5] = c.y; q[6] = d.x; q[7] = d.y;
}
// score of one memory row against the query: Luong keys.q (attention.py:55-72) or Bahdanau sum_u v_u tanh(keys_u + pq_u)
// (attention.py:25-42; q then holds the processed query plus the bias of the normed variant, v the effective v)
template <bool BAHD>
__device__ __forceinline__ float score8(const uint4& r, const float (&q)[8], const float (&v)[8]) {
  if constexpr (!BAHD) {
    return dot8p(r, q);
  } else {
    const float2 a = unpack_h2(r.x), b = unpack_h2(r.y), c = unpack_h2(r.z), d = unpack_h2(r.w);
    return ((v[0] * tanhf_acc(a.x + q[0]) + v[1] * tanhf_acc(a.y + q[1])) + (v[2] * tanhf_acc(b.x + q[2]) + v[3] * tanhf_acc(b.y + q[3]))) +
           ((v[4] * tanhf_acc(c.x + q[4]) + v[5] * tanhf_acc(c.y + q[5])) + (v[6] * tanhf_acc(d.x + q[6]) + v[7] * tanhf_acc(d.y + q[7])));
  }
}
// (step-latency trace of tools/ap4d_trace.py: stamps of thread 0 / CTA 0 inside the attention core, -DAP4D_TRACE only)
#ifdef AP4D_TRACE
__device__ int g_att_trace_row = -1;                 // set by the kernel at the top of a traced step, -1 otherwise
__device__ unsigned long long g_att_trace[16 * 8];
#define ATT_STAMP(k)                                                                                   \
  do {                                                                                                 \
    if (threadIdx.x == 0 && blockIdx.x == 0 && g_att_trace_row >= 0)                                   \
      g_att_trace[g_att_trace_row * 8 + (k)] = (unsigned long long)clock64();                          \
  } while (0)
#else
#define ATT_STAMP(k) do { } while (0)
#endif
// forward: scores (keys already in ra / rb) -> masked softmax -> alignments (shared memory and `arow` in HBM) ->
// context.  On return warp w4 == 0 holds the context in ctxv (tf32-rounded unless RAW_CTX).
// `hook()` is called once per batch of the two sweeps, right after the next batch's loads have been requested: the caller
// may use these (load-latency bound) moments for warp-uniform side work - the two-product kernel feeds its slow SS
// products to the tensor pipe there, two at a time (attn_persist4d.cu; one at a time from twice as many call sites
// measured slower).
struct NoHook {
  __device__ __forceinline__ void operator()() const {}
};
template <bool BAHD = false, bool RAW_CTX = false, class Hook = NoHook>
__device__ __forceinline__ void att_fwd_core(const AttRole& a, const float (&q)[8], const float (&v)[8], uint4 (&ra)[4],
                                             uint4 (&rb)[4], float* __restrict__ arow, float (&ctxv)[8], Hook hook = Hook()) {
  const int lane = a.lane, w4 = a.w4, gt = a.gt, L = a.L;
  float* sc = a.sc;
  float* red = a.red;
  // scores: rows tm = w4 + 4*i, one 9-shuffle reduction per batch of 8 rows
  const int jrow = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
  ATT_STAMP(0);
  for (int tm0 = w4; tm0 < L; tm0 += 4 * RIF) {
    float sacc[RIF];
#pragma unroll
    for (int j = 0; j < 4; ++j) sacc[j] = score8<BAHD>(ra[j], q, v);
#pragma unroll
    for (int j = 0; j < 4; ++j) ra[j] = ld_row(a.keys, tm0 + 32 + 4 * j, L, a.B, a.b_att, lane);
#pragma unroll
    for (int j = 0; j < 4; ++j) sacc[4 + j] = score8<BAHD>(rb[j], q, v);
#pragma unroll
    for (int j = 0; j < 4; ++j) rb[j] = ld_row(a.keys, tm0 + 48 + 4 * j, L, a.B, a.b_att, lane);
    hook();
    const float tot = warp_reduce8(sacc, lane);
    if ((lane & 3) == 0 && tm0 + 4 * jrow < L) sc[tm0 + 4 * jrow] = a.gs * tot;
  }
  // first batch of the values: in flight during the softmax
  ATT_STAMP(1);
  att_prefetch(a, a.values, ra, rb);
  att_bar(a.bar_id);
  ATT_STAMP(2);
  // masked softmax over the L scores (128 threads)
  float mx = -INFINITY;
  for (int tm = gt; tm < L; tm += 128) mx = fmaxf(mx, sc[tm]);
  mx = warp_max(mx);
  if (lane == 0) red[w4] = mx;
  att_bar(a.bar_id);
  ATT_STAMP(3);
  mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
  float sum = 0.0f;
  for (int tm = gt; tm < L; tm += 128) {
    const float e = __expf(sc[tm] - mx);
    sc[tm] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  if (lane == 0) red[4 + w4] = sum;
  att_bar(a.bar_id);
  ATT_STAMP(4);
  const float inv = L > 0 ? 1.0f / ((red[4] + red[5]) + (red[6] + red[7])) : 0.0f;
  for (int tm = gt; tm < a.Tm; tm += 128) {
    const float al = tm < L ? sc[tm] * inv : 0.0f;
    if (tm < L) sc[tm] = al;
    arow[tm] = al;
  }
  att_bar(a.bar_id);
  ATT_STAMP(5);
  // context: rows tm = w4 + 4*i, lane accumulates dims 8*lane .. +7
  for (int tm0 = w4; tm0 < L; tm0 += 4 * RIF) {
    float al[RIF];
#pragma unroll
    for (int j = 0; j < RIF; ++j) al[j] = tm0 + 4 * j < L ? sc[tm0 + 4 * j] : 0.0f;
#pragma unroll
    for (int j = 0; j < 4; ++j) axpy8p(al[j], ra[j], ctxv);
#pragma unroll
    for (int j = 0; j < 4; ++j) ra[j] = ld_row(a.values, tm0 + 32 + 4 * j, L, a.B, a.b_att, lane);
#pragma unroll
    for (int j = 0; j < 4; ++j) axpy8p(al[4 + j], rb[j], ctxv);
#pragma unroll
    for (int j = 0; j < 4; ++j) rb[j] = ld_row(a.values, tm0 + 48 + 4 * j, L, a.B, a.b_att, lane);
    hook();
  }
  ATT_STAMP(6);
#pragma unroll
  for (int e = 0; e < 8; ++e) a.part[w4 * DM + 8 * lane + e] = ctxv[e];
  att_bar(a.bar_id);
  if (w4 == 0) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float c = (a.part[8 * lane + e] + a.part[DM + 8 * lane + e]) + (a.part[2 * DM + 8 * lane + e] + a.part[3 * DM + 8 * lane + e]);
      ctxv[e] = RAW_CTX ? c : tf32_rn(c);
    }
  }
  ATT_STAMP(7);
}

// backward of the same: d(align) = values . dctx (values already in ra / rb), softmax backward, dq = g * ds^T keys.
// SMALL (memories of at most SMALL_TM rows): alignments (al, loaded by the caller: al[i] = a[w4 + 32 i + 4 jrow]) and
// d(score) (ds_keep, same rows) stay in registers; the caller writes ds / d(attention_g) off the critical path.
// Otherwise alignments come from a_s (shared memory, loaded by the caller) and this function writes `dsrow` (HBM) and
// accumulates d(attention_g).  On return warp w4 == 0 holds dq (already scaled by g) in dqv.
// BAHD: the second sweep forms the gradient wrt the PROCESSED query instead: dpq_u += ds[tm] v_u (1 - th^2),
// th = tanh(keys[tm]_u + q_u) recomputed (q = pq + bias, v the effective v); `values` is then whatever the first sweep
// multiplies with dctx_s (the projected values and d(attention vector) in attn_persist4d.cu).
template <bool BAHD>
__device__ __forceinline__ void key_axpy8(float w, const uint4& r, const float (&q)[8], const float (&v)[8], float (&acc)[8]) {
  if constexpr (!BAHD) {
    axpy8(w, r, acc);
  } else {
    const float2 a = unpack_h2(r.x), b = unpack_h2(r.y), c = unpack_h2(r.z), d = unpack_h2(r.w);
    const float k[8] = {a.x, a.y, b.x, b.y, c.x, c.y, d.x, d.y};
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float th = tanhf_acc(k[e] + q[e]);
      acc[e] = fmaf(w * v[e], 1.0f - th * th, acc[e]);
    }
  }
}
template <bool SMALL, bool BAHD = false>
__device__ __forceinline__ void att_bwd_core(const AttRole& a, const float* __restrict__ dctx_s, uint4 (&ra)[4],
                                             uint4 (&rb)[4], const float (&al)[SMALL ? SMALL_B : 1],
                                             float (&ds_keep)[SMALL ? SMALL_B : 1], const float* __restrict__ a_s,
                                             float* __restrict__ ds_s, float* __restrict__ dsrow, bool scaled,
                                             float* __restrict__ dg, float (&dqv)[8], const float (&q8)[8],
                                             const float (&v8)[8]) {
  constexpr int MAXB = SMALL ? SMALL_B : 1;
  const int lane = a.lane, w4 = a.w4, gt = a.gt, L = a.L;
  float* red = a.red;
  float dcx[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) dcx[e] = dctx_s[8 * lane + e];
  if constexpr (SMALL) {
    // d(align)[tm] = values[tm] . dctx: the batch totals stay in registers (da[i] = row jrow of batch i)
    float da[MAXB];
#pragma unroll
    for (int i = 0; i < MAXB; ++i) da[i] = 0.0f;
#pragma unroll
    for (int i = 0; i < MAXB; ++i) {
      const int tm0 = w4 + 32 * i;
      if (tm0 >= L) break;
      float sacc[RIF];
#pragma unroll
      for (int j = 0; j < 4; ++j) sacc[j] = dot8(ra[j], dcx);
#pragma unroll
      for (int j = 0; j < 4; ++j) ra[j] = ld_row(a.values, tm0 + 32 + 4 * j, L, a.B, a.b_att, lane);
#pragma unroll
      for (int j = 0; j < 4; ++j) sacc[4 + j] = dot8(rb[j], dcx);
#pragma unroll
      for (int j = 0; j < 4; ++j) rb[j] = ld_row(a.values, tm0 + 48 + 4 * j, L, a.B, a.b_att, lane);
      da[i] = warp_reduce8(sacc, lane);
    }
    // first batch of the keys: in flight during the softmax backward
    att_prefetch(a, a.keys, ra, rb);
    // dot = sum_tm a da over all rows of the utterance (al is zero past the memory length): one value per warp,
    // merged through shared memory - the only barrier between the two sweeps
    float dot = 0.0f;
#pragma unroll
    for (int i = 0; i < MAXB; ++i) dot = fmaf(al[i], da[i], dot);
    dot = warp_sum((lane & 3) == 0 ? dot : 0.0f);
    if (lane == 0) red[w4] = dot;
    att_bar(a.bar_id);
    dot = (red[0] + red[1]) + (red[2] + red[3]);
    // d(score) before the Luong scale: ds = a (da - dot), again in registers (da[i] is reused for it)
#pragma unroll
    for (int i = 0; i < MAXB; ++i) da[i] = al[i] * (da[i] - dot);
    // keys sweep: dq += g * ds[tm] * keys[tm]
#pragma unroll
    for (int i = 0; i < MAXB; ++i) {
      const int tm0 = w4 + 32 * i;
      if (tm0 >= L) break;
      float d[RIF];
#pragma unroll
      for (int j = 0; j < RIF; ++j) d[j] = __shfl_sync(0xffffffffu, da[i], reduce8_src(j));
#pragma unroll
      for (int j = 0; j < 4; ++j) key_axpy8<BAHD>(d[j], ra[j], q8, v8, dqv);
#pragma unroll
      for (int j = 0; j < 4; ++j) ra[j] = ld_row(a.keys, tm0 + 32 + 4 * j, L, a.B, a.b_att, lane);
#pragma unroll
      for (int j = 0; j < 4; ++j) key_axpy8<BAHD>(d[4 + j], rb[j], q8, v8, dqv);
#pragma unroll
      for (int j = 0; j < 4; ++j) rb[j] = ld_row(a.keys, tm0 + 48 + 4 * j, L, a.B, a.b_att, lane);
    }
#pragma unroll
    for (int i = 0; i < MAXB; ++i) ds_keep[i] = da[i];
  } else {
    // d(align)[tm] = values[tm] . dctx
    const int jrow = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
    for (int tm0 = w4; tm0 < L; tm0 += 4 * RIF) {
      float sacc[RIF];
#pragma unroll
      for (int j = 0; j < 4; ++j) sacc[j] = dot8(ra[j], dcx);
#pragma unroll
      for (int j = 0; j < 4; ++j) ra[j] = ld_row(a.values, tm0 + 32 + 4 * j, L, a.B, a.b_att, lane);
#pragma unroll
      for (int j = 0; j < 4; ++j) sacc[4 + j] = dot8(rb[j], dcx);
#pragma unroll
      for (int j = 0; j < 4; ++j) rb[j] = ld_row(a.values, tm0 + 48 + 4 * j, L, a.B, a.b_att, lane);
      const float tot = warp_reduce8(sacc, lane);
      if ((lane & 3) == 0 && tm0 + 4 * jrow < L) ds_s[tm0 + 4 * jrow] = tot;
    }
    // first batch of the keys: in flight during the softmax backward
    att_prefetch(a, a.keys, ra, rb);
    att_bar(a.bar_id);
    float dot = 0.0f;
    for (int tm = gt; tm < L; tm += 128) dot = fmaf(a_s[tm], ds_s[tm], dot);
    dot = warp_sum(dot);
    if (lane == 0) red[w4] = dot;
    att_bar(a.bar_id);
    dot = (red[0] + red[1]) + (red[2] + red[3]);
    // d(score) of this step, before the Luong scale (dkeys is formed after the loop).  d(attention_g) =
    // sum_tm ds[tm] (keys[tm].h) = (1/g) sum_tm ds[tm] log a[tm]: the scores are (log a + log Z) / g and
    // sum_tm ds[tm] = 0, so the normaliser drops out.
    float gacc = 0.0f;
    for (int tm = gt; tm < a.Tm; tm += 128) {
      const float av = tm < L ? a_s[tm] : 0.0f;
      const float d = tm < L ? av * (ds_s[tm] - dot) : 0.0f;
      dsrow[tm] = d;
      if (tm < L) ds_s[tm] = d;
      if (av > 0.0f) gacc = fmaf(d, __logf(av), gacc);
    }
    if (scaled && dg && a.gs != 0.0f) {
      gacc = warp_sum(gacc);
      if (lane == 0) atomicAdd(dg, gacc / a.gs);
    }
    att_bar(a.bar_id);
    // keys sweep: dq += g * ds[tm] * keys[tm]
    for (int tm0 = w4; tm0 < L; tm0 += 4 * RIF) {
      float d[RIF];
#pragma unroll
      for (int j = 0; j < RIF; ++j) d[j] = tm0 + 4 * j < L ? ds_s[tm0 + 4 * j] : 0.0f;
#pragma unroll
      for (int j = 0; j < 4; ++j) key_axpy8<BAHD>(d[j], ra[j], q8, v8, dqv);
#pragma unroll
      for (int j = 0; j < 4; ++j) ra[j] = ld_row(a.keys, tm0 + 32 + 4 * j, L, a.B, a.b_att, lane);
#pragma unroll
      for (int j = 0; j < 4; ++j) key_axpy8<BAHD>(d[4 + j], rb[j], q8, v8, dqv);
#pragma unroll
      for (int j = 0; j < 4; ++j) rb[j] = ld_row(a.keys, tm0 + 48 + 4 * j, L, a.B, a.b_att, lane);
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) a.part[w4 * DM + 8 * lane + e] = dqv[e];
  att_bar(a.bar_id);
  if (w4 == 0) {
#pragma unroll
    for (int e = 0; e < 8; ++e)
      dqv[e] = a.gs * ((a.part[8 * lane + e] + a.part[DM + 8 * lane + e]) + (a.part[2 * DM + 8 * lane + e] + a.part[3 * DM + 8 * lane + e]));
  }
}
// d(score) rows and d(attention_g) of a SMALL step, written off the critical path (see att_bwd_core)
__device__ __forceinline__ void att_bwd_small_tail(const AttRole& a, float* __restrict__ dsrow, const float (&al)[SMALL_B],
                                                   const float (&ds_keep)[SMALL_B], bool scaled, float* __restrict__ dg) {
  const int lane = a.lane;
  const int jrow = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
  float gacc = 0.0f;
  if ((lane & 3) == 0) {
#pragma unroll
    for (int i = 0; i < SMALL_B; ++i) {
      const int tm = a.w4 + 32 * i + 4 * jrow;
      if (tm < a.L) {
        dsrow[tm] = ds_keep[i];
        if (al[i] > 0.0f) gacc = fmaf(ds_keep[i], __logf(al[i]), gacc);
      }
    }
  }
  for (int tm = a.L + a.gt; tm < a.Tm; tm += 128) dsrow[tm] = 0.0f;
  if (scaled && dg && a.gs != 0.0f) {
    gacc = warp_sum(gacc);
    if (lane == 0) atomicAdd(dg, gacc / a.gs);
  }
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

}  // namespace ap4
}  // namespace avsr
