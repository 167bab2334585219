// Attention kernels: score -> masked softmax -> context (forward) and their backward.
// Scorers of reference avsr/attention.py:25-72 (tf.contrib.seq2seq Luong / Bahdanau).
//
// One CTA per utterance.  The keys/values of a batch are L2-resident (<= 40 MB per memory at the
// BASELINE sizes), so these kernels are L2-bandwidth / MUFU bound; they are written for memory-level
// parallelism: 128-bit loads, four memory rows in flight per warp, split-Tm partial sums.
//
// Backward is split in two so that nothing is accumulated on the sequential path:
//   attn_bwd_step   per query step: d(alignment) -> d(score) (saved) -> d(query)      [reads only]
//   attn_outer      after the loop: dvalues += align^T dctx ;  Luong: dkeys += ds^T q  (per utterance)
//   attn_bahdanau_post  after the loop: dkeys, dv, dbias of the tanh scorer, recomputing tanh
#include "../../include/avsr_b200.h"
#include <stdlib.h>

#include "common.cuh"

namespace avsr {

constexpr int ATT_THREADS = 256;

__device__ __forceinline__ float dot4(const float4 a, const float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ATT_THREADS)
attn_fwd_kernel(int kind, int Tm, int B, int Dm, int A, const float* __restrict__ q, int ldq,
                const float* __restrict__ keys, const float* __restrict__ values, const int* __restrict__ mem_len,
                const float* __restrict__ v, const float* __restrict__ g, const float* __restrict__ bias,
                float* __restrict__ align_t, float* __restrict__ ctx_out, int ldctx, int rnd) {
  extern __shared__ __align__(16) float sm[];
  const int Ap = (A + 3) & ~3, Tp = (Tm + 3) & ~3;
  float* q_s = sm;            // Ap
  float* v_s = q_s + Ap;      // Ap
  float* sc = v_s + Ap;       // Tp
  float* red = sc + Tp;       // 36
  float* part = red + 36;     // up to ATT_THREADS * 4 (split-Tm partial contexts)
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int L = min(mem_len[b], Tm);
  const bool luong = kind <= AVSR_ATTN_SCALED_LUONG;
  for (int u = tid; u < A; u += ATT_THREADS) {
    q_s[u] = q[(size_t)b * ldq + u] + ((!luong && bias) ? bias[u] : 0.0f);
    v_s[u] = luong ? 0.0f : v[u];
  }
  __syncthreads();
  const float gs = (kind == AVSR_ATTN_SCALED_LUONG) ? g[0] : 1.0f;
  if ((A & 3) == 0) {
    const int A4 = A >> 2;
    const float4* q4 = reinterpret_cast<const float4*>(q_s);
    const float4* v4 = reinterpret_cast<const float4*>(v_s);
    for (int tm0 = warp; tm0 < L; tm0 += 32) {  // rows tm0 + 8 j, j < 4: four key rows in flight per warp
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      for (int i = lane; i < A4; i += 32) {
        const float4 qq = q4[i], vv = v4[i];
        float4 k[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int tm = tm0 + 8 * j;
          k[j] = tm < L ? __ldg(reinterpret_cast<const float4*>(keys + ((size_t)tm * B + b) * A) + i)
                        : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (luong) acc[j] += dot4(k[j], qq);
          else
            acc[j] += vv.x * tanhf_acc(k[j].x + qq.x) + vv.y * tanhf_acc(k[j].y + qq.y) +
                      vv.z * tanhf_acc(k[j].z + qq.z) + vv.w * tanhf_acc(k[j].w + qq.w);
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j] = warp_sum(acc[j]);
      if (lane == 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (tm0 + 8 * j < L) sc[tm0 + 8 * j] = gs * acc[j];
      }
    }
  } else {
    for (int tm = warp; tm < L; tm += ATT_THREADS / 32) {
      const float* kr = keys + ((size_t)tm * B + b) * A;
      float acc = 0.0f;
      if (luong) {
        for (int u = lane; u < A; u += 32) acc = fmaf(kr[u], q_s[u], acc);
      } else {
        for (int u = lane; u < A; u += 32) acc = fmaf(v_s[u], tanhf_acc(kr[u] + q_s[u]), acc);
      }
      acc = warp_sum(acc);
      if (lane == 0) sc[tm] = gs * acc;
    }
  }
  __syncthreads();
  float mx = -INFINITY;
  for (int tm = tid; tm < L; tm += ATT_THREADS) mx = fmaxf(mx, sc[tm]);
  mx = block_max(mx, red);
  float sum = 0.0f;
  for (int tm = tid; tm < L; tm += ATT_THREADS) {
    float e = __expf(sc[tm] - mx);
    sc[tm] = e;
    sum += e;
  }
  sum = block_sum(sum, red);
  const float inv = L > 0 ? 1.0f / sum : 0.0f;
  for (int tm = tid; tm < Tm; tm += ATT_THREADS) {
    float a = tm < L ? sc[tm] * inv : 0.0f;
    sc[tm] = a;
    align_t[(size_t)b * Tm + tm] = a;
  }
  __syncthreads();
  const int D4 = Dm >> 2;
  if ((Dm & 3) == 0 && D4 <= ATT_THREADS && (ATT_THREADS % D4) == 0) {
    const int G = ATT_THREADS / D4;  // groups splitting the memory rows
    const int d4 = tid % D4, gq = tid / D4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int tm0 = gq; tm0 < L; tm0 += 4 * G) {
      float4 x[4];
      float a[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int tm = tm0 + j * G;
        a[j] = tm < L ? sc[tm] : 0.0f;
        x[j] = tm < L ? __ldg(reinterpret_cast<const float4*>(values + ((size_t)tm * B + b) * Dm) + d4)
                      : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc.x = fmaf(a[j], x[j].x, acc.x);
        acc.y = fmaf(a[j], x[j].y, acc.y);
        acc.z = fmaf(a[j], x[j].z, acc.z);
        acc.w = fmaf(a[j], x[j].w, acc.w);
      }
    }
    reinterpret_cast<float4*>(part)[gq * D4 + d4] = acc;
    __syncthreads();
    if (gq == 0) {
      for (int k = 1; k < G; ++k) {
        const float4 o = reinterpret_cast<const float4*>(part)[k * D4 + d4];
        acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
      }
      float* dst = ctx_out + (size_t)b * ldctx + 4 * d4;
      dst[0] = maybe_tf32(acc.x, rnd);
      dst[1] = maybe_tf32(acc.y, rnd);
      dst[2] = maybe_tf32(acc.z, rnd);
      dst[3] = maybe_tf32(acc.w, rnd);
    }
  } else {
    for (int d = tid; d < Dm; d += ATT_THREADS) {
      const float* vp = values + (size_t)b * Dm + d;
      const size_t stride = (size_t)B * Dm;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      int tm = 0;
      for (; tm + 3 < L; tm += 4) {
        a0 = fmaf(sc[tm], vp[(size_t)tm * stride], a0);
        a1 = fmaf(sc[tm + 1], vp[(size_t)(tm + 1) * stride], a1);
        a2 = fmaf(sc[tm + 2], vp[(size_t)(tm + 2) * stride], a2);
        a3 = fmaf(sc[tm + 3], vp[(size_t)(tm + 3) * stride], a3);
      }
      for (; tm < L; ++tm) a0 = fmaf(sc[tm], vp[(size_t)tm * stride], a0);
      ctx_out[(size_t)b * ldctx + d] = maybe_tf32((a0 + a1) + (a2 + a3), rnd);
    }
  }
}

size_t attn_fwd_smem(int Tm, int A) {
  return (size_t)(2 * ((A + 3) & ~3) + ((Tm + 3) & ~3) + 36 + ATT_THREADS * 4) * sizeof(float);
}

int attn_fwd(cudaStream_t st, int kind, int Tm, int B, int Dm, int A, const float* q, int ldq, const float* keys,
             const float* values, const int* mem_len, const float* v, const float* g, const float* bias,
             float* align_t, float* ctx_out, int ldctx, int rnd) {
  const size_t smem = attn_fwd_smem(Tm, A);
  AVSR_REQUIRE(smem <= 200 * 1024, "attention: memory too long for shared memory (Tm=%d)", Tm);
  if (smem > 48 * 1024)
    AVSR_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  AVSR_LAUNCH(attn_fwd_kernel, B, ATT_THREADS, smem, st, kind, Tm, B, Dm, A, q, ldq, keys, values, mem_len, v, g, bias,
              align_t, ctx_out, ldctx, rnd);
  return 0;
}

// ------------------------------------------------------------------------------------------------
// backward, per query step (no accumulation: reads keys / values, writes ds_t and dq_t)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ATT_THREADS)
attn_bwd_step_kernel(int kind, int t, const int* __restrict__ seq_len, int Tm, int B, int Dm, int A,
                     const float* __restrict__ q, int ldq, const float* __restrict__ keys,
                     const float* __restrict__ values, const int* __restrict__ mem_len, const float* __restrict__ v,
                     const float* __restrict__ g, const float* __restrict__ bias, const float* __restrict__ align_t,
                     const float* __restrict__ dctx, int lddctx, float* __restrict__ dq_out, int lddq,
                     float* __restrict__ ds_out, float* __restrict__ dg, int rnd) {
  extern __shared__ __align__(16) float sm[];
  const int Ap = (A + 3) & ~3, Tp = (Tm + 3) & ~3, Dp = (Dm + 3) & ~3;
  float* q_s = sm;             // Ap
  float* v_s = q_s + Ap;       // Ap
  float* dctx_s = v_s + Ap;    // Dp
  float* a_s = dctx_s + Dp;    // Tp
  float* ds_s = a_s + Tp;      // Tp
  float* raw_s = ds_s + Tp;    // Tp
  float* red = raw_s + Tp;     // 36
  float* part = red + 36;      // ATT_THREADS * 4
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (t >= seq_len[b]) {  // masked step: no gradient flows
    for (int u = tid; u < A; u += ATT_THREADS) dq_out[(size_t)b * lddq + u] = 0.0f;
    for (int tm = tid; tm < Tm; tm += ATT_THREADS) ds_out[(size_t)b * Tm + tm] = 0.0f;
    return;
  }
  const int L = min(mem_len[b], Tm);
  const bool luong = kind <= AVSR_ATTN_SCALED_LUONG;
  for (int u = tid; u < A; u += ATT_THREADS) {
    q_s[u] = q[(size_t)b * ldq + u] + ((!luong && bias) ? bias[u] : 0.0f);
    v_s[u] = luong ? 0.0f : v[u];
  }
  for (int d = tid; d < Dm; d += ATT_THREADS) dctx_s[d] = dctx[(size_t)b * lddctx + d];
  for (int tm = tid; tm < Tm; tm += ATT_THREADS) a_s[tm] = align_t[(size_t)b * Tm + tm];
  __syncthreads();
  // d(alignment)[tm] = dctx . values[tm]   (and the raw Luong score for d(attention_g))
  const bool need_raw = kind == AVSR_ATTN_SCALED_LUONG;
  if ((Dm & 3) == 0 && (A & 3) == 0) {
    const int D4 = Dm >> 2, A4 = A >> 2;
    const float4* dc4 = reinterpret_cast<const float4*>(dctx_s);
    const float4* q4 = reinterpret_cast<const float4*>(q_s);
    for (int tm0 = warp; tm0 < L; tm0 += 32) {
      float acc[4] = {0.f, 0.f, 0.f, 0.f}, raw[4] = {0.f, 0.f, 0.f, 0.f};
      for (int i = lane; i < D4; i += 32) {
        const float4 dd = dc4[i];
        float4 x[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int tm = tm0 + 8 * j;
          x[j] = tm < L ? __ldg(reinterpret_cast<const float4*>(values + ((size_t)tm * B + b) * Dm) + i)
                        : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] += dot4(x[j], dd);
      }
      if (need_raw) {
        for (int i = lane; i < A4; i += 32) {
          const float4 qq = q4[i];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int tm = tm0 + 8 * j;
            if (tm < L) raw[j] += dot4(__ldg(reinterpret_cast<const float4*>(keys + ((size_t)tm * B + b) * A) + i), qq);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc[j] = warp_sum(acc[j]);
        if (need_raw) raw[j] = warp_sum(raw[j]);
      }
      if (lane == 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (tm0 + 8 * j < L) {
            ds_s[tm0 + 8 * j] = acc[j];
            raw_s[tm0 + 8 * j] = raw[j];
          }
      }
    }
  } else {
    for (int tm = warp; tm < L; tm += ATT_THREADS / 32) {
      const size_t row = (size_t)tm * B + b;
      const float* vr = values + row * Dm;
      float acc = 0.0f, raw = 0.0f;
      for (int d = lane; d < Dm; d += 32) acc = fmaf(dctx_s[d], vr[d], acc);
      acc = warp_sum(acc);
      if (need_raw) {
        const float* kr = keys + row * A;
        for (int u = lane; u < A; u += 32) raw = fmaf(kr[u], q_s[u], raw);
        raw = warp_sum(raw);
      }
      if (lane == 0) {
        ds_s[tm] = acc;
        raw_s[tm] = raw;
      }
    }
  }
  __syncthreads();
  // softmax backward
  float dot = 0.0f;
  for (int tm = tid; tm < L; tm += ATT_THREADS) dot = fmaf(a_s[tm], ds_s[tm], dot);
  dot = block_sum(dot, red);
  float gsum = 0.0f;
  const float gs = need_raw ? g[0] : 1.0f;
  for (int tm = tid; tm < Tm; tm += ATT_THREADS) {
    float ds = 0.0f;
    if (tm < L) {
      ds = a_s[tm] * (ds_s[tm] - dot);
      gsum = fmaf(ds, raw_s[tm], gsum);
    }
    ds_out[(size_t)b * Tm + tm] = ds;  // d(score) before the Luong scale; consumed after the loop
    if (tm < L) ds_s[tm] = ds * gs;
  }
  if (need_raw) {
    gsum = block_sum(gsum, red);
    if (tid == 0) atomicAdd(dg, gsum);
  }
  __syncthreads();
  // d(query)[u] = sum_tm ds[tm] * d(score_tm)/d(q_u)
  const int A4 = A >> 2;
  if ((A & 3) == 0 && A4 <= ATT_THREADS && (ATT_THREADS % A4) == 0) {
    const int G = ATT_THREADS / A4;
    const int u4 = tid % A4, gq = tid / A4;
    const float4 qq = reinterpret_cast<const float4*>(q_s)[u4];
    const float4 vv = reinterpret_cast<const float4*>(v_s)[u4];
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int tm0 = gq; tm0 < L; tm0 += 4 * G) {
      float4 k[4];
      float d[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int tm = tm0 + j * G;
        d[j] = tm < L ? ds_s[tm] : 0.0f;
        k[j] = tm < L ? __ldg(reinterpret_cast<const float4*>(keys + ((size_t)tm * B + b) * A) + u4)
                      : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (luong) {
          acc.x = fmaf(d[j], k[j].x, acc.x);
          acc.y = fmaf(d[j], k[j].y, acc.y);
          acc.z = fmaf(d[j], k[j].z, acc.z);
          acc.w = fmaf(d[j], k[j].w, acc.w);
        } else {
          const float tx = tanhf_acc(k[j].x + qq.x), ty = tanhf_acc(k[j].y + qq.y);
          const float tz = tanhf_acc(k[j].z + qq.z), tw = tanhf_acc(k[j].w + qq.w);
          acc.x = fmaf(d[j] * vv.x, 1.0f - tx * tx, acc.x);
          acc.y = fmaf(d[j] * vv.y, 1.0f - ty * ty, acc.y);
          acc.z = fmaf(d[j] * vv.z, 1.0f - tz * tz, acc.z);
          acc.w = fmaf(d[j] * vv.w, 1.0f - tw * tw, acc.w);
        }
      }
    }
    reinterpret_cast<float4*>(part)[gq * A4 + u4] = acc;
    __syncthreads();
    if (gq == 0) {
      for (int k2 = 1; k2 < G; ++k2) {
        const float4 o = reinterpret_cast<const float4*>(part)[k2 * A4 + u4];
        acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
      }
      float* dst = dq_out + (size_t)b * lddq + 4 * u4;
      const int r2 = luong ? 0 : rnd;  // d(processed query) is a tensor-core operand, d(h) is not
      dst[0] = maybe_tf32(acc.x, r2);
      dst[1] = maybe_tf32(acc.y, r2);
      dst[2] = maybe_tf32(acc.z, r2);
      dst[3] = maybe_tf32(acc.w, r2);
    }
  } else {
    for (int u = tid; u < A; u += ATT_THREADS) {
      const size_t stride = (size_t)B * A;
      const float* kp = keys + (size_t)b * A + u;
      float dq = 0.0f;
      if (luong) {
        for (int tm = 0; tm < L; ++tm) dq = fmaf(ds_s[tm], kp[(size_t)tm * stride], dq);
      } else {
        const float qu = q_s[u], vu = v_s[u];
        for (int tm = 0; tm < L; ++tm) {
          const float th = tanhf_acc(kp[(size_t)tm * stride] + qu);
          dq = fmaf(ds_s[tm] * vu, 1.0f - th * th, dq);
        }
        dq = maybe_tf32(dq, rnd);
      }
      dq_out[(size_t)b * lddq + u] = dq;
    }
  }
}

size_t attn_bwd_smem(int Tm, int A, int Dm) {
  return (size_t)(2 * ((A + 3) & ~3) + ((Dm + 3) & ~3) + 3 * ((Tm + 3) & ~3) + 36 + ATT_THREADS * 4) * sizeof(float);
}

int attn_bwd_step(cudaStream_t st, int kind, int t, const int* seq_len, int Tm, int B, int Dm, int A, const float* q,
                  int ldq, const float* keys, const float* values, const int* mem_len, const float* v, const float* g,
                  const float* bias, const float* align_t, const float* dctx, int lddctx, float* dq_out, int lddq,
                  float* ds_out, float* dg, int rnd) {
  const size_t smem = attn_bwd_smem(Tm, A, Dm);
  AVSR_REQUIRE(smem <= 200 * 1024, "attention: memory too long for shared memory (Tm=%d)", Tm);
  if (smem > 48 * 1024)
    AVSR_CHECK_CUDA(
        cudaFuncSetAttribute(attn_bwd_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  AVSR_LAUNCH(attn_bwd_step_kernel, B, ATT_THREADS, smem, st, kind, t, seq_len, Tm, B, Dm, A, q, ldq, keys, values,
              mem_len, v, g, bias, align_t, dctx, lddctx, dq_out, lddq, ds_out, dg, rnd);
  return 0;
}

// ------------------------------------------------------------------------------------------------
// after the loop: out[tm, b, c] += scale * sum_t w[t, b, tm] * x[t, b, c]      (per utterance outer products)
//   dvalues = align^T dctx ;  Luong dkeys = g * ds^T q
// grid (B, ceil(Tm / 16)); 256 threads over columns c
// ------------------------------------------------------------------------------------------------
constexpr int OUT_TM = 16;   // memory rows per CTA
constexpr int OUT_TT = 32;   // query steps staged per chunk

constexpr int OUTER_TM = 16;  // memory rows per CTA of attn_outer (32 rows: 128 registers, fewer resident CTAs - slower)

__global__ void __launch_bounds__(256)
attn_outer_kernel(int T, int B, int Tm, int C, const int* __restrict__ seq_len, const float* __restrict__ w,
                  const float* __restrict__ x, int ldx, const float* __restrict__ scale, float* __restrict__ out) {
  __shared__ float w_s[OUT_TT][OUTER_TM];
  const int b = blockIdx.x, tm0 = blockIdx.y * OUTER_TM, tid = threadIdx.x;
  const int Tb = min(T, seq_len[b]);
  const int ntm = min(OUTER_TM, Tm - tm0);
  for (int c0 = 0; c0 < C; c0 += 256) {
    const int c = c0 + tid;
    float acc[OUTER_TM];
#pragma unroll
    for (int i = 0; i < OUTER_TM; ++i) acc[i] = 0.0f;
    for (int t0 = 0; t0 < Tb; t0 += OUT_TT) {
      __syncthreads();
      for (int e = tid; e < OUT_TT * OUTER_TM; e += 256) {
        const int tt = e / OUTER_TM, i = e % OUTER_TM;
        const int t = t0 + tt;
        w_s[tt][i] = (t < Tb && i < ntm) ? w[((size_t)t * B + b) * Tm + tm0 + i] : 0.0f;
      }
      __syncthreads();
      if (c < C) {
        const int nt = min(OUT_TT, Tb - t0);
        for (int tt0 = 0; tt0 < nt; tt0 += 8) {  // 8 independent row loads in flight per thread (L2 / HBM latency)
          float xv[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int t = t0 + tt0 + j;
            xv[j] = (tt0 + j < nt) ? __ldg(x + ((size_t)t * B + b) * ldx + c) : 0.0f;
          }
#pragma unroll
          for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int i = 0; i < OUTER_TM; ++i) acc[i] = fmaf(w_s[tt0 + j][i], xv[j], acc[i]);
        }
      }
    }
    if (c < C) {
      const float s = scale ? scale[0] : 1.0f;
#pragma unroll
      for (int i = 0; i < OUTER_TM; ++i)  // (constant indices keep acc[] in registers)
        if (i < ntm) out[((size_t)(tm0 + i) * B + b) * C + c] += s * acc[i];
    }
  }
}

// Tensor-core form of the same accumulation (tensor-core mode only): per utterance out[Tm, C] += w^T[Tm, T] x[T, C] is a
// small dense product, 256 of them per batch.  CTA = (utterance, 32 memory rows); 8 warps x 32 columns; the w chunk
// (A operand, [32 steps][32 rows]) is staged tf32-rounded in shared memory, the x rows (B operand) are read straight
// from global memory in fragment order (8 consecutive floats of 4 rows per load instruction = full 32-byte sectors) and
// rounded to tf32 on the fly; mma.sync.m16n8k8 with fp32 accumulators.  The SIMT kernel above spends 168 issue slots per
// warp and 8 steps, this one 14.
constexpr int OM_ROWS = 32;       // memory rows per CTA (two m16 tiles)
constexpr int OM_TT = 32;         // query steps per staged chunk (four k8 steps)
constexpr int OM_LD = 40;         // padded row stride of the staged chunk: conflict-free A-fragment reads

__device__ __forceinline__ void mma_tf32_16x8x8(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(256)
attn_outer_mma_kernel(int T, int B, int Tm, int C, const int* __restrict__ seq_len, const float* __restrict__ w,
                      const float* __restrict__ x, int ldx, const float* __restrict__ scale, float* __restrict__ out) {
  __shared__ float w_s[OM_TT][OM_LD];
  // the row tiles of one utterance are neighbours in launch order: they all read the same x rows, so the re-reads are
  // L2 hits (ncu showed 210 MB of DRAM reads per launch for 100 MB of operands with the utterance as the fast grid
  // index; the launch time did not move - 90 us - so the kernel is bound by its chunk-serial structure, not by DRAM)
  const int b = blockIdx.y, tm0 = blockIdx.x * OM_ROWS, tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, tg = lane & 3;
  const int Tb = min(T, seq_len[b]);
  const int ntm = min(OM_ROWS, Tm - tm0);
  for (int c0 = 0; c0 < C; c0 += 256) {
    const int cw = c0 + warp * 32;  // first column of this warp
    float acc[2][4][4];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
      for (int n = 0; n < 4; ++n)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[m][n][e] = 0.0f;
    for (int t0 = 0; t0 < Tb; t0 += OM_TT) {
      __syncthreads();
      for (int e = tid; e < OM_TT * OM_ROWS; e += 256) {
        const int tt = e / OM_ROWS, i = e % OM_ROWS;
        const int t = t0 + tt;
        w_s[tt][i] = (t < Tb && i < ntm) ? tf32_rn(w[((size_t)t * B + b) * Tm + tm0 + i]) : 0.0f;
      }
      // B fragments of the whole chunk first (independent loads in flight), then the products
      uint32_t bf[4][4][2];
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
#pragma unroll
        for (int n = 0; n < 4; ++n)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int t = t0 + ks * 8 + tg + 4 * h;
            const int c = cw + n * 8 + g;
            const float v = (t < Tb && c < C) ? __ldg(x + ((size_t)t * B + b) * ldx + c) : 0.0f;
            bf[ks][n][h] = __float_as_uint(tf32_rn(v));
          }
      __syncthreads();
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        if (t0 + ks * 8 >= Tb) break;
#pragma unroll
        for (int m = 0; m < 2; ++m) {
          uint32_t a[4];
          a[0] = __float_as_uint(w_s[ks * 8 + tg][m * 16 + g]);
          a[1] = __float_as_uint(w_s[ks * 8 + tg][m * 16 + g + 8]);
          a[2] = __float_as_uint(w_s[ks * 8 + tg + 4][m * 16 + g]);
          a[3] = __float_as_uint(w_s[ks * 8 + tg + 4][m * 16 + g + 8]);
#pragma unroll
          for (int n = 0; n < 4; ++n) mma_tf32_16x8x8(acc[m][n], a, bf[ks][n][0], bf[ks][n][1]);
        }
      }
    }
    const float s = scale ? scale[0] : 1.0f;
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
      for (int n = 0; n < 4; ++n)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int row = m * 16 + g + 8 * h, c = cw + n * 8 + 2 * tg;
          if (row < ntm && c < C) {  // (C is even for every caller: attention / memory depths)
            float* o = out + ((size_t)(tm0 + row) * B + b) * C + c;
            if (c + 1 < C) {
              float2 v = *reinterpret_cast<float2*>(o);
              v.x += s * acc[m][n][2 * h];
              v.y += s * acc[m][n][2 * h + 1];
              *reinterpret_cast<float2*>(o) = v;
            } else {
              o[0] += s * acc[m][n][2 * h];
            }
          }
        }
  }
}

int attn_outer(cudaStream_t st, int T, int B, int Tm, int C, const int* seq_len, const float* w, const float* x,
               int ldx, const float* scale, float* out) {
  if (T <= 0) return 0;
  if (tensor_cores_enabled() && (C % 2 == 0) && (((uintptr_t)out & 7) == 0) && !getenv("AVSR_OUTER_SIMT")) {
    dim3 grid(cdiv(Tm, OM_ROWS), B);
    AVSR_LAUNCH(attn_outer_mma_kernel, grid, 256, 0, st, T, B, Tm, C, seq_len, w, x, ldx, scale, out);
    return 0;
  }
  dim3 grid(B, cdiv(Tm, OUTER_TM));
  AVSR_LAUNCH(attn_outer_kernel, grid, 256, 0, st, T, B, Tm, C, seq_len, w, x, ldx, scale, out);
  return 0;
}

// ------------------------------------------------------------------------------------------------
// after the loop, Bahdanau family: with th = tanh(keys[tm,b,u] + pq[t,b,u] + bias[u])
//   dkeys[tm,b,u] += v[u] * sum_t ds[t,b,tm] (1 - th^2);  dv[u] += sum ds th;  dbias[u] += sum dE
// grid (B, ceil(Tm / 16)); threads over u
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
attn_bahdanau_post_kernel(int T, int B, int Tm, int A, const int* __restrict__ seq_len,
                          const int* __restrict__ mem_len, const float* __restrict__ ds,
                          const float* __restrict__ pq, const float* __restrict__ keys, const float* __restrict__ v,
                          const float* __restrict__ bias, float* __restrict__ dkeys, float* __restrict__ dv,
                          float* __restrict__ dbias) {
  __shared__ float w_s[OUT_TT][OUT_TM];
  const int b = blockIdx.x, tm0 = blockIdx.y * OUT_TM, tid = threadIdx.x;
  const int Tb = min(T, seq_len[b]);
  const int L = min(mem_len[b], Tm);
  const int ntm = min(OUT_TM, L - tm0);  // rows past the memory length carry no gradient
  if (ntm <= 0) return;
  for (int u0 = 0; u0 < A; u0 += 256) {
    const int u = u0 + tid;
    float kreg[OUT_TM], acc[OUT_TM];
    float dvu = 0.0f;
#pragma unroll
    for (int i = 0; i < OUT_TM; ++i) {
      acc[i] = 0.0f;
      kreg[i] = (u < A && i < ntm) ? keys[((size_t)(tm0 + i) * B + b) * A + u] + (bias ? bias[u] : 0.0f) : 0.0f;
    }
    for (int t0 = 0; t0 < Tb; t0 += OUT_TT) {
      __syncthreads();
      for (int e = tid; e < OUT_TT * OUT_TM; e += 256) {
        const int tt = e / OUT_TM, i = e % OUT_TM;
        const int t = t0 + tt;
        w_s[tt][i] = (t < Tb && i < ntm) ? ds[((size_t)t * B + b) * Tm + tm0 + i] : 0.0f;
      }
      __syncthreads();
      if (u < A) {
        const int nt = min(OUT_TT, Tb - t0);
        for (int tt = 0; tt < nt; ++tt) {
          const float pqv = __ldg(pq + ((size_t)(t0 + tt) * B + b) * A + u);
#pragma unroll
          for (int i = 0; i < OUT_TM; ++i) {
            const float d = w_s[tt][i];
            const float th = tanhf_acc(kreg[i] + pqv);
            acc[i] = fmaf(d, 1.0f - th * th, acc[i]);
            dvu = fmaf(d, th, dvu);
          }
        }
      }
    }
    if (u < A) {
      const float vu = v[u];
      float sumE = 0.0f;
      for (int i = 0; i < ntm; ++i) {
        const float e = vu * acc[i];
        dkeys[((size_t)(tm0 + i) * B + b) * A + u] += e;
        sumE += e;
      }
      atomicAdd(dv + u, dvu);
      if (dbias) atomicAdd(dbias + u, sumE);
    }
  }
}

// Contexts of all steps from the saved alignments: ctx[t,b,:] = sum_tm align[t,b,tm] values[tm,b,:] for t < len[b], else
// 0; tf32-rounded (operand of the attention-layer weight gradient).  Used after the persistent Bahdanau kernels, which
// keep the context half of the attention layer out of the recurrence (attn_persist4d.cu).  Block = (utterance, chunk
// of CTX_TC steps): the alignments of the chunk sit in shared memory, a thread owns feature columns d, d + 256.
constexpr int CTX_TC = 16;
__global__ void attn_context_all_kernel(int T, int B, int Tm, int Dm, const int* __restrict__ seq_len,
                                        const int* __restrict__ mem_len, const float* __restrict__ align,
                                        const float* __restrict__ values, float* __restrict__ ctx, int ldc, int rnd) {
  extern __shared__ float sa[];  // [CTX_TC][Tm]
  const int b = blockIdx.x, t0 = blockIdx.y * CTX_TC;
  const int n = min(CTX_TC, T - t0);
  const int L = min(mem_len[b], Tm), len = seq_len[b];
  for (int i = threadIdx.x; i < n * Tm; i += blockDim.x) {
    const int t = i / Tm, tm = i - t * Tm;
    sa[i] = (t0 + t < len && tm < L) ? align[((size_t)(t0 + t) * B + b) * Tm + tm] : 0.0f;
  }
  __syncthreads();
  for (int d0 = threadIdx.x; d0 < Dm; d0 += 2 * blockDim.x) {
    const int d1 = d0 + blockDim.x;
    float acc0[CTX_TC], acc1[CTX_TC];
#pragma unroll
    for (int t = 0; t < CTX_TC; ++t) acc0[t] = acc1[t] = 0.0f;
    for (int tm = 0; tm < L; ++tm) {
      const float* vr = values + ((size_t)tm * B + b) * Dm;
      const float v0 = vr[d0], v1 = d1 < Dm ? vr[d1] : 0.0f;
#pragma unroll
      for (int t = 0; t < CTX_TC; ++t) {
        const float a = sa[t * Tm + tm];
        acc0[t] = fmaf(a, v0, acc0[t]);
        acc1[t] = fmaf(a, v1, acc1[t]);
      }
    }
#pragma unroll
    for (int t = 0; t < CTX_TC; ++t)
      if (t < n) {
        float* o = ctx + ((size_t)(t0 + t) * B + b) * ldc;
        o[d0] = maybe_tf32(acc0[t], rnd);
        if (d1 < Dm) o[d1] = maybe_tf32(acc1[t], rnd);
      }
  }
}

int attn_context_all(cudaStream_t st, int T, int B, int Tm, int Dm, const int* seq_len, const int* mem_len,
                     const float* align, const float* values, float* ctx, int ldc) {
  if (T <= 0) return 0;
  const size_t smem = (size_t)CTX_TC * Tm * sizeof(float);
  AVSR_REQUIRE(smem <= 48 * 1024, "attn_context_all: memory of %d rows too long", Tm);
  AVSR_LAUNCH(attn_context_all_kernel, dim3(B, cdiv(T, CTX_TC)), 256, smem, st, T, B, Tm, Dm, seq_len, mem_len, align,
              values, ctx, ldc, tensor_cores_enabled());
  return 0;
}

int attn_bahdanau_post(cudaStream_t st, int T, int B, int Tm, int A, const int* seq_len, const int* mem_len,
                       const float* ds, const float* pq, const float* keys, const float* v, const float* bias,
                       float* dkeys, float* dv, float* dbias) {
  if (T <= 0) return 0;
  dim3 grid(B, cdiv(Tm, OUT_TM));
  AVSR_LAUNCH(attn_bahdanau_post_kernel, grid, 256, 0, st, T, B, Tm, A, seq_len, mem_len, ds, pq, keys, v, bias, dkeys,
              dv, dbias);
  return 0;
}

}  // namespace avsr
