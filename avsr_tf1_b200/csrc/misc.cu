// Memory-bound helpers of the hot path: input batch-norm, reverse_sequence,
// embedding, masked sequence loss, global-norm / clip / Adam, decode helpers.
#include <float.h>

#include "../../include/avsr_b200.h"
#include "common.cuh"

namespace avsr {

// ---- column statistics: out[f] += sum_r a[r,f]; out[F+f] += sum_r a[r,f]*b[r,f] (b==null -> a*a)
__global__ void colstats_kernel(const float* __restrict__ a, const float* __restrict__ b, long long rows, int F,
                                int rows_per_block, float* __restrict__ out) {
  __shared__ float s0[8][33], s1[8][33];
  const int f = blockIdx.x * 32 + threadIdx.x;
  const long long r0 = (long long)blockIdx.y * rows_per_block;
  const long long r1 = min(rows, r0 + rows_per_block);
  float p0 = 0.f, p1 = 0.f;
  if (f < F) {
    for (long long r = r0 + threadIdx.y; r < r1; r += 8) {
      float x = a[r * F + f];
      float y = b ? b[r * F + f] : x;
      p0 += x;
      p1 = fmaf(x, y, p1);
    }
  }
  s0[threadIdx.y][threadIdx.x] = p0;
  s1[threadIdx.y][threadIdx.x] = p1;
  __syncthreads();
  if (threadIdx.y == 0 && f < F) {
#pragma unroll
    for (int k = 1; k < 8; ++k) {
      p0 += s0[k][threadIdx.x];
      p1 += s1[k][threadIdx.x];
    }
    atomicAdd(out + f, p0);
    atomicAdd(out + F + f, p1);
  }
}

static int colstats(cudaStream_t st, const float* a, const float* b, long long rows, int F, float* out) {
  if (rows <= 0) return 0;
  int rpb = 256;
  while (cdiv(rows, rpb) > 65535) rpb *= 2;  // grid.y limit (the CNN's batch norms see N*H*W = tens of millions of rows)
  dim3 grid(cdiv(F, 32), cdiv(rows, rpb));
  AVSR_LAUNCH(colstats_kernel, grid, dim3(32, 8), 0, st, a, b, rows, F, rpb, out);
  return 0;
}

__global__ void colsum_kernel(const float* __restrict__ a, long long rows, int N, int ld, int rows_per_block,
                              float* __restrict__ out) {
  __shared__ float s0[8][33];
  const int f = blockIdx.x * 32 + threadIdx.x;
  const long long r0 = (long long)blockIdx.y * rows_per_block;
  const long long r1 = min(rows, r0 + rows_per_block);
  float p0 = 0.f;
  if (f < N)
    for (long long r = r0 + threadIdx.y; r < r1; r += 8) p0 += a[r * ld + f];
  s0[threadIdx.y][threadIdx.x] = p0;
  __syncthreads();
  if (threadIdx.y == 0 && f < N) {
#pragma unroll
    for (int k = 1; k < 8; ++k) p0 += s0[k][threadIdx.x];
    atomicAdd(out + f, p0);
  }
}

// d0 > 0: x is stored [d0, d1, F] and y / xhat are written [d1, d0, F] (the batch-major -> frame-major change of
// layout at the boundary, fused into the normalisation instead of a separate pass over the features)
__global__ void bn_apply_train_kernel(const float* __restrict__ x, long long n, int F, const float* __restrict__ sums,
                                      float inv_count, const float* __restrict__ gamma,
                                      const float* __restrict__ beta, float eps, float momentum,
                                      float* __restrict__ y, float* __restrict__ xhat, float* __restrict__ invstd,
                                      float* __restrict__ moving_mean, float* __restrict__ moving_var, int rnd, int d0,
                                      int d1) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int f = (int)(i % F);
  float mean = sums[f] * inv_count;
  float var = fmaxf(sums[F + f] * inv_count - mean * mean, 0.0f);
  float is = rsqrtf(var + eps);
  long long src = i;
  if (d0 > 0) {
    const long long r = i / F;  // output row = j*d0 + k  (j in d1, k in d0)
    const int k = (int)(r % d0), j = (int)(r / d0);
    src = ((long long)k * d1 + j) * F + f;
  }
  float xh = (x[src] - mean) * is;
  if (xhat) xhat[i] = xh;  // only kept when the gradient wrt the raw features will be asked for
  y[i] = maybe_tf32(fmaf(xh, gamma[f], beta[f]), rnd);
  if (i < F) {  // first row: per-feature side outputs
    invstd[f] = is;
    if (moving_mean) moving_mean[f] = moving_mean[f] * momentum + mean * (1.0f - momentum);
    if (moving_var) moving_var[f] = moving_var[f] * momentum + var * (1.0f - momentum);
  }
}

// Vectorised form for F % 4 == 0: a CTA owns a strip of BN_ROWS output rows x 1024 feature columns; every thread keeps
// the scale / shift of its 4 features in registers and streams float4s (the scalar kernel above recomputes the
// statistics per element and spends its time on 64-bit index arithmetic: 4x off the HBM roofline on the 3888-wide
// lip-crop features).
constexpr int BN_ROWS = 16;
__global__ void __launch_bounds__(256)
bn_apply_train_v4_kernel(const float* __restrict__ x, long long rows, int F, const float* __restrict__ sums,
                         float inv_count, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                         float momentum, float* __restrict__ y, float* __restrict__ xhat, float* __restrict__ invstd,
                         float* __restrict__ moving_mean, float* __restrict__ moving_var, int rnd, int d0, int d1) {
  const int f = (blockIdx.y * 256 + threadIdx.x) * 4;  // (rows on grid.x: no 65535 limit)
  if (f >= F) return;
  const float4 s1 = *reinterpret_cast<const float4*>(sums + f), s2 = *reinterpret_cast<const float4*>(sums + F + f);
  const float4 g4 = *reinterpret_cast<const float4*>(gamma + f), b4 = *reinterpret_cast<const float4*>(beta + f);
  const float sm[4] = {s1.x, s1.y, s1.z, s1.w}, sq[4] = {s2.x, s2.y, s2.z, s2.w};
  const float gm[4] = {g4.x, g4.y, g4.z, g4.w}, bt[4] = {b4.x, b4.y, b4.z, b4.w};
  float mean[4], is[4], var[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    mean[e] = sm[e] * inv_count;
    var[e] = fmaxf(sq[e] * inv_count - mean[e] * mean[e], 0.0f);
    is[e] = rsqrtf(var[e] + eps);
  }
  if (blockIdx.x == 0) {  // per-feature side outputs, once
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      invstd[f + e] = is[e];
      if (moving_mean) moving_mean[f + e] = moving_mean[f + e] * momentum + mean[e] * (1.0f - momentum);
      if (moving_var) moving_var[f + e] = moving_var[f + e] * momentum + var[e] * (1.0f - momentum);
    }
  }
  const long long r0 = (long long)blockIdx.x * BN_ROWS;
#pragma unroll 4
  for (int k = 0; k < BN_ROWS; ++k) {
    const long long r = r0 + k;  // output row
    if (r >= rows) break;
    long long src = r;
    if (d0 > 0) {  // output row = j*d0 + kk (j in d1, kk in d0)  <-  input row kk*d1 + j
      const long long j = r / d0, kk = r - j * d0;
      src = kk * d1 + j;
    }
    const float4 v = __ldcs(reinterpret_cast<const float4*>(x + src * F + f));
    const float xv[4] = {v.x, v.y, v.z, v.w};
    float xh[4], o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      xh[e] = (xv[e] - mean[e]) * is[e];
      o[e] = maybe_tf32(fmaf(xh[e], gm[e], bt[e]), rnd);
    }
    if (xhat) *reinterpret_cast<float4*>(xhat + r * F + f) = make_float4(xh[0], xh[1], xh[2], xh[3]);
    *reinterpret_cast<float4*>(y + r * F + f) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

static int bn_apply_train_launch(cudaStream_t st, const float* x, long long rows, int F, const float* sums,
                                 float inv_count, const float* gamma, const float* beta, float eps, float momentum,
                                 float* y, float* xhat, float* invstd, float* moving_mean, float* moving_var, int d0,
                                 int d1) {
  const bool aligned = (F % 4 == 0) && ((((uintptr_t)x | (uintptr_t)y | (uintptr_t)xhat | (uintptr_t)sums |
                                          (uintptr_t)gamma | (uintptr_t)beta) & 15) == 0);
  if (aligned) {
    dim3 grid(cdiv(rows, BN_ROWS), cdiv(F, 1024));
    AVSR_LAUNCH(bn_apply_train_v4_kernel, grid, 256, 0, st, x, rows, F, sums, inv_count, gamma, beta, eps, momentum, y,
                xhat, invstd, moving_mean, moving_var, tensor_cores_enabled(), d0, d1);
    return 0;
  }
  const long long n = rows * F;
  AVSR_LAUNCH(bn_apply_train_kernel, cdiv(n, 256), 256, 0, st, x, n, F, sums, inv_count, gamma, beta, eps, momentum, y,
              xhat, invstd, moving_mean, moving_var, tensor_cores_enabled(), d0, d1);
  return 0;
}

// Gradients of the input normalisation's gamma / beta WITHOUT the gradient wrt its output: with y = gamma xhat + beta
// feeding only the layer-0 gate product z = y Wx, and dWx = y^T dZ, cs = colsum(dZ) already computed,
//   dbeta_f  = sum_r dy[r,f]           = sum_n Wx[f,n] cs[n]
//   dgamma_f = sum_r dy[r,f] xhat[r,f] = sum_n Wx[f,n] (xhat^T dZ)[f,n] = sum_n Wx[f,n] (dWx[f,n] - beta_f cs[n]) / gamma_f
// so neither dy = dZ Wx^T (a [T*B, F] product) nor xhat is needed.  One block per feature f; accumulates.
__global__ void bn_input_grads_kernel(const float* __restrict__ Wx, int ldw, const float* __restrict__ dWx, int ldg,
                                      const float* __restrict__ cs, const float* __restrict__ gamma,
                                      const float* __restrict__ beta, int N, float* __restrict__ dgamma,
                                      float* __restrict__ dbeta) {
  const int f = blockIdx.x;
  const float* w = Wx + (size_t)f * ldw;
  const float* g = dWx + (size_t)f * ldg;
  const float bf = beta[f];
  float sb = 0.0f, sg = 0.0f;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const float wn = w[n], c = cs[n];
    sb = fmaf(wn, c, sb);
    sg = fmaf(wn, g[n] - bf * c, sg);
  }
  __shared__ float red[2][32];
  sb = warp_sum(sb);
  sg = warp_sum(sg);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    red[0][warp] = sb;
    red[1][warp] = sg;
  }
  __syncthreads();
  if (warp == 0) {
    const int nw = blockDim.x >> 5;
    sb = lane < nw ? red[0][lane] : 0.0f;
    sg = lane < nw ? red[1][lane] : 0.0f;
    sb = warp_sum(sb);
    sg = warp_sum(sg);
    if (lane == 0) {
      dbeta[f] += sb;
      const float gf = gamma[f];
      if (gf != 0.0f) dgamma[f] += sg / gf;  // gamma_f = 0 exactly: xhat^T dZ cannot be recovered from dWx (contributes 0)
    }
  }
}

__global__ void bn_apply_eval_kernel(const float* __restrict__ x, long long n, int F, const float* __restrict__ gamma,
                                     const float* __restrict__ beta, const float* __restrict__ mm,
                                     const float* __restrict__ mv, float eps, float* __restrict__ y, int rnd) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int f = (int)(i % F);
  y[i] = maybe_tf32(fmaf((x[i] - mm[f]) * rsqrtf(mv[f] + eps), gamma[f], beta[f]), rnd);
}

__global__ void bn_bwd_apply_kernel(const float* __restrict__ dy, const float* __restrict__ xhat, long long n, int F,
                                    const float* __restrict__ sums2, float inv_count,
                                    const float* __restrict__ gamma, const float* __restrict__ invstd,
                                    float* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int f = (int)(i % F);
  float sdy = sums2[f], sdyx = sums2[F + f];
  dx[i] = gamma[f] * invstd[f] * (dy[i] - sdy * inv_count - xhat[i] * sdyx * inv_count);
  if (i < F) {
    if (dgamma) dgamma[f] += sdyx;
    if (dbeta) dbeta[f] += sdy;
  }
}

__global__ void reverse_sequence_kernel(const float* __restrict__ x, float* __restrict__ y, int T, int B, int F,
                                        const int* __restrict__ len) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long n = (long long)T * B * F;
  if (i >= n) return;
  int f = (int)(i % F);
  long long tb = i / F;
  int b = (int)(tb % B), t = (int)(tb / B);
  int L = min(len[b], T);
  int ts = t < L ? L - 1 - t : t;
  y[i] = x[((long long)ts * B + b) * F + f];
}

__global__ void transpose01_kernel(const float* __restrict__ x, float* __restrict__ y, int d0, int d1, int F) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long n = (long long)d0 * d1 * F;
  if (i >= n) return;
  int f = (int)(i % F);
  long long r = i / F;  // output row = j*d0 + k  (j in d1, k in d0)
  int k = (int)(r % d0), j = (int)(r / d0);
  y[i] = x[((long long)k * d1 + j) * F + f];
}

// lip-crop pixels as they are stored (uint8) -> the reference's float features (v - 128) / 128 (dataset_writer.py:537):
// the crops cross PCIe as bytes and are expanded here (exact: k / 128 is a float).  16 pixels per thread.
__global__ void u8_to_f32_kernel(const uint8_t* __restrict__ x, long long n, float scale, float shift, float* __restrict__ y) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 16;
  if (i + 16 <= n) {
    const uint4 v = *reinterpret_cast<const uint4*>(x + i);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k)
      *reinterpret_cast<float4*>(y + i + 4 * k) =
          make_float4(((float)(w[k] & 255u) + shift) * scale, ((float)((w[k] >> 8) & 255u) + shift) * scale,
                      ((float)((w[k] >> 16) & 255u) + shift) * scale, ((float)(w[k] >> 24) + shift) * scale);
  } else {
    for (long long j = i; j < n; ++j) y[j] = ((float)x[j] + shift) * scale;
  }
}

__global__ void embedding_fwd_kernel(const float* __restrict__ table, int V, int E, const int* __restrict__ ids,
                                     long long n, float* __restrict__ out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * E) return;
  long long r = i / E;
  int e = (int)(i - r * E);
  int id = ids[r];
  out[i] = (id >= 0 && id < V) ? table[(size_t)id * E + e] : 0.0f;
}

// block (v, chunk): sums the rows of one chunk of ids that hit vocabulary row v, one atomic per (v, e, chunk)
constexpr int EMB_CHUNK = 512;
__global__ void embedding_bwd_kernel(const float* __restrict__ dout, const int* __restrict__ ids, long long n, int E,
                                     float* __restrict__ dtable) {
  const int v = blockIdx.x;
  const long long r0 = (long long)blockIdx.y * EMB_CHUNK, r1 = min(n, r0 + EMB_CHUNK);
  __shared__ int hit[EMB_CHUNK];
  __shared__ int nhit;
  if (threadIdx.x == 0) nhit = 0;
  __syncthreads();
  for (long long r = r0 + threadIdx.x; r < r1; r += blockDim.x)
    if (ids[r] == v) hit[atomicAdd(&nhit, 1)] = (int)(r - r0);
  __syncthreads();
  if (nhit == 0) return;
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    float acc = 0.0f;
    for (int i = 0; i < nhit; ++i) acc += dout[(r0 + hit[i]) * E + e];
    atomicAdd(dtable + (size_t)v * E + e, acc);
  }
}

// one warp per (t,b) row.  smoothing > 0: the reference's label-smoothing path (seq2seq.py:147-155 -> devel.py:54-61):
// the targets become (1 - eps) onehot + eps / V AND the loss turns into the UNMASKED mean over all T x B positions - the
// smoothed loss function returns a reduced scalar, which sequence_loss multiplies by the weights and divides by their sum
// again.  Rows past the label length hold imputed zero logits: they add log V to the sum and carry no gradient.
__global__ void seq_loss_kernel(const float* __restrict__ logits, int T, int B, int V, const int* __restrict__ labels,
                                int ldl, const int* __restrict__ labels_len, const float* __restrict__ inv_denom_dev,
                                float smoothing, float* __restrict__ loss_sum, float* __restrict__ dlogits) {
  const float inv_denom = inv_denom_dev[0];
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= T * B) return;
  const int t = warp / B, b = warp - t * B;
  const float* z = logits + (size_t)warp * V;
  float* dz = dlogits + (size_t)warp * V;
  if (t >= labels_len[b]) {
    for (int v = lane; v < V; v += 32) dz[v] = 0.0f;
    if (smoothing > 0.0f && lane == 0) atomicAdd(loss_sum, logf((float)V));
    return;
  }
  float mx = -INFINITY;
  for (int v = lane; v < V; v += 32) mx = fmaxf(mx, z[v]);
  mx = warp_max(mx);
  float s = 0.0f, zs = 0.0f;
  for (int v = lane; v < V; v += 32) {
    s += expf(z[v] - mx);
    zs += z[v];
  }
  s = warp_sum(s);
  zs = warp_sum(zs);
  const float lse = logf(s) + mx;
  const int y = labels[(size_t)b * ldl + t];
  const float on = 1.0f - smoothing, off = smoothing / (float)V;
  for (int v = lane; v < V; v += 32) {
    float p = expf(z[v] - lse);
    dz[v] = (p - (v == y ? on : 0.0f) - off) * inv_denom;
  }
  if (lane == 0) atomicAdd(loss_sum, lse - on * z[y] - off * zs);
}

// devel.py losses as `softmax_loss_function` of sequence_loss (seq2seq.py:156-163; devel.py:12-52, "not tested thoroughly"
// by its authors): per token, with p = clip(softmax(z), 1e-7, 1 - 1e-7) and y the label,
//   mc_loss    = - log p_y - sum_{k != y} log(1 - p_k)
//   focal_loss = - (1 - p_y)^g log p_y - sum_{k != y} p_k^g log(1 - p_k)          (g = 2)
// masked and normalised like the cross-entropy (weights = sequence_mask, / token count).  Gradient: dL/dp through the
// clip (zero where it clipped) and the softmax Jacobian.  One warp per (t, b) row.
__global__ void seq_loss_devel_kernel(const float* __restrict__ logits, int T, int B, int V, const int* __restrict__ labels,
                                      int ldl, const int* __restrict__ labels_len, const float* __restrict__ inv_denom_dev,
                                      int kind, float gamma, float* __restrict__ loss_sum, float* __restrict__ dlogits) {
  const float inv_denom = inv_denom_dev[0];
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= T * B) return;
  const int t = warp / B, b = warp - t * B;
  const float* z = logits + (size_t)warp * V;
  float* dz = dlogits + (size_t)warp * V;
  if (t >= labels_len[b]) {
    for (int v = lane; v < V; v += 32) dz[v] = 0.0f;
    return;
  }
  float mx = -INFINITY;
  for (int v = lane; v < V; v += 32) mx = fmaxf(mx, z[v]);
  mx = warp_max(mx);
  float se = 0.0f;
  for (int v = lane; v < V; v += 32) se += expf(z[v] - mx);
  se = warp_sum(se);
  const float lse = logf(se) + mx;
  const int y = labels[(size_t)b * ldl + t];
  float loss = 0.0f, dot = 0.0f;
  for (int v = lane; v < V; v += 32) {
    const float sv = expf(z[v] - lse);
    const float p = fminf(fmaxf(sv, 1e-7f), 1.0f - 1e-7f);
    const bool inside = sv >= 1e-7f && sv <= 1.0f - 1e-7f;
    float L, g;
    if (kind == 1) {  // mc_loss
      if (v == y) { L = -logf(p); g = -1.0f / p; }
      else { L = -logf(1.0f - p); g = 1.0f / (1.0f - p); }
    } else {          // focal_loss
      if (v == y) {
        const float q = 1.0f - p, qg = powf(q, gamma);
        L = -qg * logf(p);
        g = gamma * powf(q, gamma - 1.0f) * logf(p) - qg / p;
      } else {
        const float pg = powf(p, gamma);
        L = -pg * logf(1.0f - p);
        g = -gamma * powf(p, gamma - 1.0f) * logf(1.0f - p) + pg / (1.0f - p);
      }
    }
    loss += L;
    dot += inside ? sv * g : 0.0f;
  }
  loss = warp_sum(loss);
  dot = warp_sum(dot);
  for (int v = lane; v < V; v += 32) {  // (recomputed: V <= a few dozen classes)
    const float sv = expf(z[v] - lse);
    const float p = fminf(fmaxf(sv, 1e-7f), 1.0f - 1e-7f);
    const bool inside = sv >= 1e-7f && sv <= 1.0f - 1e-7f;
    float g;
    if (kind == 1) g = v == y ? -1.0f / p : 1.0f / (1.0f - p);
    else if (v == y) g = gamma * powf(1.0f - p, gamma - 1.0f) * logf(p) - powf(1.0f - p, gamma) / p;
    else g = -gamma * powf(p, gamma - 1.0f) * logf(1.0f - p) + powf(p, gamma) / (1.0f - p);
    dz[v] = sv * ((inside ? g : 0.0f) - dot) * inv_denom;
  }
  if (lane == 0) atomicAdd(loss_sum, loss);
}

// Action-Unit head: one thread per (t, b, k)
__global__ void au_loss_kernel(const float* __restrict__ z, int T, int B, const float* __restrict__ aus,
                               const int* __restrict__ len, const float* __restrict__ scale_dev,
                               float* __restrict__ loss_sum, float* __restrict__ dz) {
  __shared__ float red[33];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float sq = 0.0f;
  if (i < T * B * 2) {
    const int k = i & 1, row = i >> 1, t = row / B, b = row - t * B;
    float d = 0.0f;
    if (t < len[b]) {
      const float p = sigmoidf_acc(z[i]);
      const float y = fminf(fmaxf(aus[((size_t)b * T + t) * 2 + k], 0.0f), 3.0f) * (1.0f / 3.0f);
      const float e = p - y;
      sq = e * e;
      d = 2.0f * e * p * (1.0f - p) * scale_dev[0];
    }
    dz[i] = d;
  }
  sq = block_sum(sq, red);
  if (threadIdx.x == 0 && sq != 0.0f) atomicAdd(loss_sum, sq);
}

__global__ void sumsq_kernel(const float* __restrict__ x, long long n, float* __restrict__ out) {
  __shared__ float red[33];
  float acc = 0.0f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    acc = fmaf(x[i], x[i], acc);
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(out, acc);
}

__global__ void axpy_kernel(float a, const float* __restrict__ x, float* __restrict__ y, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = fmaf(a, x[i], y[i]);
}

__global__ void adam_clip_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                 float* __restrict__ v, long long n, const float* __restrict__ sumsq, float clip,
                                 const float* __restrict__ lr_t_dev, float b1, float b2, float eps,
                                 float* __restrict__ p_tf32) {
  const float lr_t = lr_t_dev[0];
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float scale = 1.0f;
  if (clip > 0.0f) {
    float norm = sqrtf(sumsq[0]);
    scale = clip / fmaxf(norm, clip);  // tf.clip_by_global_norm
  }
  float gi = g[i] * scale;
  float mi = b1 * m[i] + (1.0f - b1) * gi;
  float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
  m[i] = mi;
  v[i] = vi;
  const float pn = p[i] - lr_t * mi / (sqrtf(vi) + eps);  // TF-Adam: epsilon outside the sqrt
  p[i] = pn;
  if (p_tf32) p_tf32[i] = tf32_rn(pn);
}

// the reference's other optimisers (seq2seq.py:200-219), same single pass: Nadam (apply_adam use_nesterov=True), AdamW
// (tf.contrib.opt: var -= weight_decay * var, NOT scaled by the learning rate, then the Adam update), Momentum 0.9
__global__ void optim_clip_kernel(int kind, float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                  float* __restrict__ v, long long n, const float* __restrict__ sumsq, float clip,
                                  const float* __restrict__ lr_t_dev, float b1, float b2, float eps, float wd,
                                  float* __restrict__ p_tf32) {
  const float lr_t = lr_t_dev[0];
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float scale = 1.0f;
  if (clip > 0.0f) scale = clip / fmaxf(sqrtf(sumsq[0]), clip);
  const float gi = g[i] * scale;
  float pn = p[i];
  if (kind == AVSR_OPT_MOMENTUM) {
    const float acc = b1 * m[i] + gi;
    m[i] = acc;
    pn -= lr_t * acc;
  } else {
    if (kind == AVSR_OPT_ADAMW) pn -= wd * pn;
    const float mi = b1 * m[i] + (1.0f - b1) * gi;
    const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float num = kind == AVSR_OPT_NADAM ? mi * b1 + (1.0f - b1) * gi : mi;
    pn -= lr_t * num / (sqrtf(vi) + eps);
  }
  p[i] = pn;
  if (p_tf32) p_tf32[i] = tf32_rn(pn);
}

__global__ void round_tf32_kernel(const float* __restrict__ src, float* __restrict__ dst, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = tf32_rn(src[i]);
}

__global__ void normed_v_fwd_kernel(const float* __restrict__ v, const float* __restrict__ g, int A,
                                    float* __restrict__ veff) {
  __shared__ float red[33];
  float acc = 0.0f;
  for (int u = threadIdx.x; u < A; u += blockDim.x) acc = fmaf(v[u], v[u], acc);
  acc = block_sum(acc, red);
  float s = g[0] * rsqrtf(acc);
  for (int u = threadIdx.x; u < A; u += blockDim.x) veff[u] = v[u] * s;
}

__global__ void normed_v_bwd_kernel(const float* __restrict__ v, const float* __restrict__ g,
                                    const float* __restrict__ dveff, int A, float* __restrict__ dv,
                                    float* __restrict__ dg) {
  __shared__ float red[33];
  float n2 = 0.0f, dot = 0.0f;
  for (int u = threadIdx.x; u < A; u += blockDim.x) {
    n2 = fmaf(v[u], v[u], n2);
    dot = fmaf(dveff[u], v[u], dot);
  }
  n2 = block_sum(n2, red);
  dot = block_sum(dot, red);
  float inv = rsqrtf(n2);
  float dvhat = dot * inv;  // dveff . vhat
  if (threadIdx.x == 0) dg[0] += dvhat;
  for (int u = threadIdx.x; u < A; u += blockDim.x) dv[u] += g[0] * inv * (dveff[u] - dvhat * v[u] * inv);
}

// ---- visual front-end helpers (video.py: tf.layers.conv2d as im2col + dense product, batch_norm_relu) ----------------
// cols[row = (n, oy, ox)][col = (ky, kx, c)] = x[n, oy*stride - pad_top + ky, ox*stride - pad_left + kx, c] (0 outside);
// the column order is the kernel variable [kh, kw, Cin, Cout] flattened to [kh*kw*Cin, Cout]
__global__ void im2col_kernel(const float* __restrict__ x, int N, int H, int W, int C, int kh, int kw, int stride, int pt,
                              int pl, int Ho, int Wo, int rnd, float* __restrict__ cols) {
  const long long K = (long long)kh * kw * C, total = (long long)N * Ho * Wo * K;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / K;
    const int col = (int)(i - row * K);
    const int c = col % C, kx = (col / C) % kw, ky = col / (C * kw);
    const int ox = (int)(row % Wo), oy = (int)((row / Wo) % Ho), n = (int)(row / ((long long)Wo * Ho));
    const int iy = oy * stride - pt + ky, ix = ox * stride - pl + kx;
    float v = 0.0f;
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = x[(((long long)n * H + iy) * W + ix) * C + c];
    cols[i] = maybe_tf32(v, rnd);
  }
}

// transpose of im2col as a gather (deterministic, no atomics): dx[n, y, x, c] = sum over the (ky, kx) whose window
// position (oy, ox) covers the pixel
__global__ void col2im_kernel(const float* __restrict__ dcols, int N, int H, int W, int C, int kh, int kw, int stride,
                              int pt, int pl, int Ho, int Wo, float* __restrict__ dx) {
  const long long K = (long long)kh * kw * C, total = (long long)N * H * W * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C), xx = (int)((i / C) % W), yy = (int)((i / ((long long)C * W)) % H);
    const int n = (int)(i / ((long long)C * W * H));
    float acc = 0.0f;
    for (int ky = 0; ky < kh; ++ky) {
      const int ny = yy + pt - ky;
      if (ny < 0 || ny % stride) continue;
      const int oy = ny / stride;
      if (oy >= Ho) continue;
      for (int kx = 0; kx < kw; ++kx) {
        const int nx = xx + pl - kx;
        if (nx < 0 || nx % stride) continue;
        const int ox = nx / stride;
        if (ox >= Wo) continue;
        acc += dcols[(((long long)n * Ho + oy) * Wo + ox) * K + ((long long)ky * kw + kx) * C + c];
      }
    }
    dx[i] = acc;
  }
}

__global__ void relu_fwd_kernel(const float* __restrict__ x, long long n, float* __restrict__ y) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = fmaxf(x[i], 0.0f);
}
__global__ void relu_bwd_kernel(const float* __restrict__ y, const float* __restrict__ dy, long long n,
                                float* __restrict__ dx) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dx[i] = y[i] > 0.0f ? dy[i] : 0.0f;
}

// tf.nn.selu of the optional dense stack in front of an encoder (encoder.py:148-171): y = s * (x > 0 ? x : a * (e^x - 1));
// backward from the saved y: dy/dx = s for y > 0, else y + s a
#define AVSR_SELU_SCALE 1.0507009873554804934193349852946f
#define AVSR_SELU_ALPHA 1.6732632423543772848170429916717f
__global__ void selu_fwd_kernel(const float* __restrict__ x, long long n, float* __restrict__ y) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = x[i];
    y[i] = AVSR_SELU_SCALE * (v > 0.0f ? v : AVSR_SELU_ALPHA * expm1f(v));
  }
}
__global__ void selu_bwd_kernel(const float* __restrict__ y, const float* __restrict__ dy, long long n,
                                float* __restrict__ dx) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = y[i];
    dx[i] = dy[i] * (v > 0.0f ? AVSR_SELU_SCALE : v + AVSR_SELU_SCALE * AVSR_SELU_ALPHA);
  }
}

// tf.contrib.rnn.HighwayWrapper around an encoder layer (cells.py:89-90; coupled gates, carry bias 1):
//   carry = sigmoid(x Wc + bc) (`pre` = the product), y = x * carry + out * (1 - carry)
// backward: dx (direct part only: the carry product's share is a GEMM on dpre), dout, dpre
__global__ void highway_fwd_kernel(const float* __restrict__ x, const float* __restrict__ pre,
                                   const float* __restrict__ out, long long n, int rnd, float* __restrict__ y,
                                   float* __restrict__ y_op) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float c = 1.0f / (1.0f + expf(-pre[i]));
    const float v = x[i] * c + out[i] * (1.0f - c);
    y[i] = v;
    if (y_op) y_op[i] = maybe_tf32(v, rnd);
  }
}
__global__ void highway_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                   const float* __restrict__ pre, const float* __restrict__ out, long long n, int rnd,
                                   float* __restrict__ dx, float* __restrict__ dout, float* __restrict__ dpre) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float c = 1.0f / (1.0f + expf(-pre[i]));
    const float d = dy[i];
    dx[i] = d * c;
    dout[i] = d * (1.0f - c);
    dpre[i] = maybe_tf32(d * (x[i] - out[i]) * c * (1.0f - c), rnd);
  }
}

// ---- direct convolutions for the narrow layers of the front-end (8 / 16 output channels: almost all of its pixels) ---
// One thread per output pixel, all CO output channels in registers, the kernel [kh*kw*Ci][CO] in shared memory (broadcast
// reads); NHWC input read straight from global memory (the 3x3 neighbourhoods of adjacent threads overlap in L1).  No
// im2col buffer: at the bench batch the 8-channel layers alone would write and re-read 7 GB each.
template <int CO>
__global__ void __launch_bounds__(256)
conv_direct_kernel(const float* __restrict__ x, int N, int H, int W, int Ci, const float* __restrict__ w,
                   const float* __restrict__ bias, int kh, int kw, int stride, int pt, int pl, int Ho, int Wo,
                   float* __restrict__ y) {
  extern __shared__ __align__(16) float conv_ws[];
  const int K = kh * kw * Ci;
  for (int i = threadIdx.x; i < K * CO; i += blockDim.x) conv_ws[i] = w[i];
  __syncthreads();
  const long long total = (long long)N * Ho * Wo;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(p % Wo), oy = (int)((p / Wo) % Ho), n = (int)(p / ((long long)Wo * Ho));
    float acc[CO];
#pragma unroll
    for (int co = 0; co < CO; ++co) acc[co] = bias ? bias[co] : 0.0f;
    for (int ky = 0; ky < kh; ++ky) {
      const int iy = oy * stride - pt + ky;
      if (iy < 0 || iy >= H) continue;
      for (int kx = 0; kx < kw; ++kx) {
        const int ix = ox * stride - pl + kx;
        if (ix < 0 || ix >= W) continue;
        const float* xp = x + (((long long)n * H + iy) * W + ix) * Ci;
        const float* wp = conv_ws + (ky * kw + kx) * Ci * CO;
        for (int ci = 0; ci < Ci; ++ci) {
          const float v = __ldg(xp + ci);
#pragma unroll
          for (int co = 0; co < CO; ++co) acc[co] = fmaf(v, wp[ci * CO + co], acc[co]);
        }
      }
    }
    float4* o = reinterpret_cast<float4*>(y + p * CO);
#pragma unroll
    for (int q = 0; q < CO / 4; ++q) o[q] = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
  }
}

// dW[k = (ky, kx, ci)][co] += sum over output pixels of x-patch[k] * dy[co].  A CTA stages tiles of 64 output pixels (their
// im2col rows and dy rows) in shared memory; every thread owns up to 9 of the K*CO <= 2304 outputs and keeps them in
// registers across all the tiles the CTA visits (grid-stride), one atomicAdd per output at the end.
constexpr int WG_PIX = 64;
template <int CO>
__global__ void __launch_bounds__(256)
conv_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy, int N, int H, int W, int Ci, int kh, int kw,
                  int stride, int pt, int pl, int Ho, int Wo, float* __restrict__ dW) {
  extern __shared__ __align__(16) float wg_sm[];
  const int K = kh * kw * Ci, KP = K | 1, nout = K * CO;
  float* xs = wg_sm;                 // [WG_PIX][KP]
  float* ds = wg_sm + WG_PIX * KP;   // [WG_PIX][CO]
  const long long total = (long long)N * Ho * Wo;
  float acc[9];
#pragma unroll
  for (int j = 0; j < 9; ++j) acc[j] = 0.0f;
  for (long long tile = blockIdx.x; tile * WG_PIX < total; tile += gridDim.x) {
    const long long p0 = tile * WG_PIX;
    __syncthreads();
    for (int e = threadIdx.x; e < WG_PIX * K; e += 256) {
      const int pp = e / K, k = e - pp * K;
      const long long p = p0 + pp;
      float v = 0.0f;
      if (p < total) {
        const int ci = k % Ci, kx = (k / Ci) % kw, ky = k / (Ci * kw);
        const int ox = (int)(p % Wo), oy = (int)((p / Wo) % Ho), n = (int)(p / ((long long)Wo * Ho));
        const int iy = oy * stride - pt + ky, ix = ox * stride - pl + kx;
        if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = __ldg(x + (((long long)n * H + iy) * W + ix) * Ci + ci);
      }
      xs[pp * KP + k] = v;
    }
    for (int e = threadIdx.x; e < WG_PIX * CO; e += 256) {
      const long long p = p0 + e / CO;
      ds[e] = p < total ? __ldg(dy + p0 * CO + e) : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 9; ++j) {
      const int o = threadIdx.x + 256 * j;
      if (o < nout) {
        const int k = o / CO, co = o - k * CO;
        float a = acc[j];
#pragma unroll 8
        for (int pp = 0; pp < WG_PIX; ++pp) a = fmaf(xs[pp * KP + k], ds[pp * CO + co], a);
        acc[j] = a;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 9; ++j) {
    const int o = threadIdx.x + 256 * j;
    if (o < nout) atomicAdd(dW + o, acc[j]);
  }
}

// ---- randomness of the training graph (dropout, scheduled sampling) -----------------------------------------------
__global__ void dropout_kernel(const float* __restrict__ x, long long n, long long first,
                               const uint32_t* __restrict__ rng, uint32_t stream_id, uint32_t thr, float inv_keep,
                               int rnd, float* __restrict__ y) {
  const uint32_t seed = rng[0], step = rng[1];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float f = avsr_rand_u32(seed, step, stream_id, 0u, (uint32_t)(first + i)) < thr ? inv_keep : 0.0f;
    y[i] = maybe_tf32(x[i] * f, rnd);
  }
}

// four elements per thread (16-byte accesses): n, first and both pointers are multiples of 4 elements
__global__ void dropout_v4_kernel(const float4* __restrict__ x, long long n4, long long first,
                                  const uint32_t* __restrict__ rng, uint32_t stream_id, uint32_t thr, float inv_keep,
                                  int rnd, float4* __restrict__ y) {
  const uint32_t seed = rng[0], step = rng[1];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = x[i];
    const uint32_t lo = (uint32_t)(first + 4 * i);
    float4 o;
    o.x = maybe_tf32(v.x * (avsr_rand_u32(seed, step, stream_id, 0u, lo) < thr ? inv_keep : 0.0f), rnd);
    o.y = maybe_tf32(v.y * (avsr_rand_u32(seed, step, stream_id, 0u, lo + 1u) < thr ? inv_keep : 0.0f), rnd);
    o.z = maybe_tf32(v.z * (avsr_rand_u32(seed, step, stream_id, 0u, lo + 2u) < thr ? inv_keep : 0.0f), rnd);
    o.w = maybe_tf32(v.w * (avsr_rand_u32(seed, step, stream_id, 0u, lo + 3u) < thr ? inv_keep : 0.0f), rnd);
    y[i] = o;
  }
}

// one thread per row: V is the output alphabet (31)
__global__ void sched_sample_kernel(const float* __restrict__ logits, int B, int V, const uint32_t* __restrict__ rng,
                                    uint32_t stream_id, int t, uint32_t thr_p, const int* __restrict__ true_next,
                                    int* __restrict__ next_ids, int* __restrict__ sampled) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const uint32_t seed = rng[0], step = rng[1];
  int pick = -1;
  if (avsr_rand_u32(seed, step, stream_id, (uint32_t)t, (uint32_t)b) < thr_p) {
    const float* z = logits + (size_t)b * V;
    float m = z[0];
    for (int v = 1; v < V; ++v) m = fmaxf(m, z[v]);
    float total = 0.0f;
    for (int v = 0; v < V; ++v) total += expf(z[v] - m);
    const float u = (float)(avsr_rand_u32(seed, step, stream_id + 1u, (uint32_t)t, (uint32_t)b) >> 8) * (1.0f / 16777216.0f);
    const float target = u * total;
    float cum = 0.0f;
    pick = V - 1;
    for (int v = 0; v < V; ++v) {
      cum += expf(z[v] - m);
      if (cum > target) {
        pick = v;
        break;
      }
    }
  }
  sampled[b] = pick;
  next_ids[b] = pick >= 0 ? pick : true_next[b];
}

// ---- decoding helpers -----------------------------------------------------------
__global__ void greedy_pick_kernel(const float* __restrict__ logits, int B, int V, int eos, int* __restrict__ finished,
                                   int* __restrict__ sample_out, int* __restrict__ next_ids) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float* z = logits + (size_t)b * V;
  int best = 0;
  float bv = z[0];
  for (int v = 1; v < V; ++v)
    if (z[v] > bv) {
      bv = z[v];
      best = v;
    }
  int fin = finished[b];
  sample_out[b] = fin ? 0 : best;  // impute_finished zeroes finished rows
  next_ids[b] = best;
  finished[b] = fin | (best == eos);
}

// one CTA per utterance; W*V candidates
__global__ void beam_step_kernel(const float* __restrict__ logits, int W, int V, int eos, float lpw,
                                 float* __restrict__ log_probs, int* __restrict__ finished, int* __restrict__ lengths,
                                 int* __restrict__ word_out, int* __restrict__ parent_out,
                                 float* __restrict__ score_out) {
  extern __shared__ float sm[];
  const int N = W * V;
  float* score = sm;           // N
  float* total = score + N;    // N
  float* lp_old = total + N;   // W
  int* fin_old = (int*)(lp_old + W);
  int* len_old = fin_old + W;
  int* taken = len_old + W;    // N
  __shared__ float rv[32];
  __shared__ int ri[32];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  for (int w = tid; w < W; w += blockDim.x) {
    lp_old[w] = log_probs[b * W + w];
    fin_old[w] = finished[b * W + w];
    len_old[w] = lengths[b * W + w];
  }
  __syncthreads();
  for (int w = warp; w < W; w += nw) {
    const float* z = logits + ((size_t)b * W + w) * V;
    float mx = -INFINITY;
    for (int v = lane; v < V; v += 32) mx = fmaxf(mx, z[v]);
    mx = warp_max(mx);
    float s = 0.0f;
    for (int v = lane; v < V; v += 32) s += expf(z[v] - mx);
    s = warp_sum(s);
    const float ls = logf(s);
    for (int v = lane; v < V; v += 32) {
      float lsm = (z[v] - mx) - ls;
      if (fin_old[w]) lsm = (v == eos) ? 0.0f : -FLT_MAX;
      float tot = lp_old[w] + lsm;
      int nl = len_old[w] + ((!fin_old[w] && v != eos) ? 1 : 0);
      float pen = lpw == 0.0f ? 1.0f : powf((5.0f + (float)nl) / 6.0f, lpw);
      total[w * V + v] = tot;
      score[w * V + v] = tot / pen;
      taken[w * V + v] = 0;
    }
  }
  __syncthreads();
  for (int k = 0; k < W; ++k) {
    // argmax with lowest-index tie-break (tf.nn.top_k)
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int i = tid; i < N; i += blockDim.x) {
      if (taken[i]) continue;
      float s = score[i];
      if (bi == 0x7fffffff || s > bv) {
        bv = s;
        bi = i;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (oi != 0x7fffffff && (bi == 0x7fffffff || ov > bv || (ov == bv && oi < bi))) {
        bv = ov;
        bi = oi;
      }
    }
    if (lane == 0) {
      rv[warp] = bv;
      ri[warp] = bi;
    }
    __syncthreads();
    if (tid == 0) {
      for (int w2 = 1; w2 < nw; ++w2) {
        float ov = rv[w2];
        int oi = ri[w2];
        if (oi != 0x7fffffff && (bi == 0x7fffffff || ov > bv || (ov == bv && oi < bi))) {
          bv = ov;
          bi = oi;
        }
      }
      taken[bi] = 1;
      int word = bi % V, parent = bi / V;
      int pf = fin_old[parent];
      word_out[b * W + k] = word;
      parent_out[b * W + k] = parent;
      score_out[b * W + k] = bv;
      log_probs[b * W + k] = total[bi];
      lengths[b * W + k] = len_old[parent] + (pf ? 0 : 1);
      finished[b * W + k] = pf | (word == eos);
    }
    __syncthreads();
  }
}

__global__ void gather_rows_kernel(const float* __restrict__ src, const int* __restrict__ idx, long long n, int F,
                                   float* __restrict__ dst) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * F) return;
  long long r = i / F;
  int f = (int)(i - r * F);
  dst[i] = src[(size_t)idx[r] * F + f];
}

}  // namespace avsr

// ================================================================================
// C ABI
// ================================================================================
using namespace avsr;
#define ST(s) ((cudaStream_t)(s))

extern "C" {

int avsr_colsum(avsr_stream_t s, const float* X, int M, int N, int ldx, float* out) {
  if (M <= 0 || N <= 0) return 0;
  int rpb = 256;
  while (cdiv(M, rpb) > 65535) rpb *= 2;  // grid.y limit
  dim3 grid(cdiv(N, 32), cdiv(M, rpb));
  AVSR_LAUNCH(colsum_kernel, grid, dim3(32, 8), 0, ST(s), X, (long long)M, N, ldx, rpb, out);
  return 0;
}

int avsr_bn_stats(avsr_stream_t s, const float* x, long long rows, int F, float* sums) {
  return colstats(ST(s), x, nullptr, rows, F, sums);
}

int avsr_bn_apply_train(avsr_stream_t s, const float* x, long long rows, int F, const float* sums, double count,
                        const float* gamma, const float* beta, float eps, float momentum, float* y, float* xhat,
                        float* invstd, float* moving_mean, float* moving_var) {
  AVSR_REQUIRE(rows >= 1 && count >= 1.0, "bn: empty batch");
  return bn_apply_train_launch(ST(s), x, rows, F, sums, (float)(1.0 / count), gamma, beta, eps, momentum, y, xhat,
                               invstd, moving_mean, moving_var, 0, 0);
}

int avsr_bn_apply_train_t(avsr_stream_t s, const float* x, int d0, int d1, int F, const float* sums, double count,
                          const float* gamma, const float* beta, float eps, float momentum, float* y, float* xhat,
                          float* invstd, float* moving_mean, float* moving_var) {
  AVSR_REQUIRE(d0 >= 1 && d1 >= 1 && count >= 1.0, "bn: empty batch");
  AVSR_REQUIRE(x != y && x != xhat, "bn_apply_train_t cannot run in place");
  return bn_apply_train_launch(ST(s), x, (long long)d0 * d1, F, sums, (float)(1.0 / count), gamma, beta, eps, momentum,
                               y, xhat, invstd, moving_mean, moving_var, d0, d1);
}

int avsr_bn_apply_eval(avsr_stream_t s, const float* x, long long rows, int F, const float* gamma, const float* beta,
                       const float* moving_mean, const float* moving_var, float eps, float* y) {
  long long n = rows * F;
  if (n <= 0) return 0;
  AVSR_LAUNCH(bn_apply_eval_kernel, cdiv(n, 256), 256, 0, ST(s), x, n, F, gamma, beta, moving_mean, moving_var, eps,
              y, tensor_cores_enabled());
  return 0;
}

int avsr_bn_input_grads(avsr_stream_t s, const float* Wx, int ldw, const float* dWx, int ldg, const float* colsum_dZ,
                         const float* gamma, const float* beta, int F, int N, float* dgamma, float* dbeta) {
  if (F <= 0 || N <= 0) return 0;
  AVSR_LAUNCH(bn_input_grads_kernel, F, 128, 0, ST(s), Wx, ldw, dWx, ldg, colsum_dZ, gamma, beta, N, dgamma, dbeta);
  return 0;
}

int avsr_bn_bwd_stats(avsr_stream_t s, const float* dy, const float* xhat, long long rows, int F, float* sums2) {
  return colstats(ST(s), dy, xhat, rows, F, sums2);
}

int avsr_bn_bwd_apply(avsr_stream_t s, const float* dy, const float* xhat, long long rows, int F, const float* sums2,
                      double count, const float* gamma, const float* invstd, float* dx, float* dgamma,
                      float* dbeta) {
  long long n = rows * F;
  if (n <= 0) return 0;
  AVSR_LAUNCH(bn_bwd_apply_kernel, cdiv(n, 256), 256, 0, ST(s), dy, xhat, n, F, sums2, (float)(1.0 / count), gamma,
              invstd, dx, dgamma, dbeta);
  return 0;
}

int avsr_reverse_sequence(avsr_stream_t s, const float* x, float* y, int T, int B, int F, const int* len) {
  long long n = (long long)T * B * F;
  if (n <= 0) return 0;
  AVSR_REQUIRE(x != y, "reverse_sequence cannot run in place");
  AVSR_LAUNCH(reverse_sequence_kernel, cdiv(n, 256), 256, 0, ST(s), x, y, T, B, F, len);
  return 0;
}

int avsr_transpose01(avsr_stream_t s, const float* x, float* y, int d0, int d1, int F) {
  long long n = (long long)d0 * d1 * F;
  if (n <= 0) return 0;
  AVSR_REQUIRE(x != y, "transpose01 cannot run in place");
  AVSR_LAUNCH(transpose01_kernel, cdiv(n, 256), 256, 0, ST(s), x, y, d0, d1, F);
  return 0;
}

int avsr_u8_to_f32(avsr_stream_t s, const uint8_t* x, long long n, float scale, float shift, float* y) {
  if (n <= 0) return 0;
  AVSR_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0, "u8_to_f32: buffers must be 16-byte aligned");
  AVSR_LAUNCH(u8_to_f32_kernel, cdiv(cdiv(n, 16), 256), 256, 0, ST(s), x, n, scale, shift, y);
  return 0;
}

int avsr_embedding_fwd(avsr_stream_t s, const float* table, int V, int E, const int* ids, long long n, float* out) {
  if (n <= 0) return 0;
  AVSR_LAUNCH(embedding_fwd_kernel, cdiv(n * E, 256), 256, 0, ST(s), table, V, E, ids, n, out);
  return 0;
}

int avsr_embedding_bwd(avsr_stream_t s, const float* dout, const int* ids, long long n, int V, int E,
                       float* dtable) {
  if (n <= 0) return 0;
  AVSR_LAUNCH(embedding_bwd_kernel, dim3(V, cdiv(n, EMB_CHUNK)), 128, 0, ST(s), dout, ids, n, E, dtable);
  return 0;
}

int avsr_seq_loss(avsr_stream_t s, const float* logits, int T, int B, int V, const int* labels, int ldl,
                  const int* labels_len, const float* inv_denom, float label_smoothing, float* loss_sum, float* dlogits) {
  if (T * B <= 0) return 0;
  AVSR_LAUNCH(seq_loss_kernel, cdiv((long long)T * B * 32, 256), 256, 0, ST(s), logits, T, B, V, labels, ldl,
              labels_len, inv_denom, label_smoothing, loss_sum, dlogits);
  return 0;
}

int avsr_seq_loss_devel(avsr_stream_t s, const float* logits, int T, int B, int V, const int* labels, int ldl,
                        const int* labels_len, const float* inv_denom, int kind, float gamma, float* loss_sum,
                        float* dlogits) {
  AVSR_REQUIRE(kind == 1 || kind == 2, "seq_loss_devel: kind must be 1 (mc_loss) or 2 (focal_loss)");
  if (T * B <= 0) return 0;
  AVSR_LAUNCH(seq_loss_devel_kernel, cdiv((long long)T * B * 32, 256), 256, 0, ST(s), logits, T, B, V, labels, ldl,
              labels_len, inv_denom, kind, gamma, loss_sum, dlogits);
  return 0;
}

int avsr_au_loss(avsr_stream_t s, const float* z, int T, int B, const float* aus, const int* len,
                 const float* scale_dev, float* loss_sum, float* dz) {
  if (T * B <= 0) return 0;
  AVSR_LAUNCH(au_loss_kernel, cdiv((long long)T * B * 2, 256), 256, 0, ST(s), z, T, B, aus, len, scale_dev, loss_sum, dz);
  return 0;
}

int avsr_sumsq(avsr_stream_t s, const float* x, long long n, float* out) {
  if (n <= 0) return 0;
  int grid = (int)min((long long)148 * 8, (long long)cdiv(n, 256));
  AVSR_LAUNCH(sumsq_kernel, grid, 256, 0, ST(s), x, n, out);
  return 0;
}

int avsr_axpy(avsr_stream_t s, float a, const float* x, float* y, long long n) {
  if (n <= 0) return 0;
  AVSR_LAUNCH(axpy_kernel, cdiv(n, 256), 256, 0, ST(s), a, x, y, n);
  return 0;
}

int avsr_adam_clip_step(avsr_stream_t s, float* params, const float* grads, float* m, float* v, long long n,
                        const float* sumsq_dev, float clip_norm, const float* lr_t, float beta1, float beta2, float eps,
                        float* params_tf32) {
  if (n <= 0) return 0;
  AVSR_LAUNCH(adam_clip_kernel, cdiv(n, 256), 256, 0, ST(s), params, grads, m, v, n, sumsq_dev, clip_norm, lr_t,
              beta1, beta2, eps, params_tf32);
  return 0;
}

int avsr_optim_clip_step(avsr_stream_t s, int kind, float* params, const float* grads, float* m, float* v, long long n,
                         const float* sumsq_dev, float clip_norm, const float* lr_t, float beta1, float beta2, float eps,
                         float weight_decay, float* params_tf32) {
  AVSR_REQUIRE(kind >= AVSR_OPT_ADAM && kind <= AVSR_OPT_MOMENTUM, "optimiser kind %d unknown", kind);
  if (n <= 0) return 0;
  AVSR_LAUNCH(optim_clip_kernel, cdiv(n, 256), 256, 0, ST(s), kind, params, grads, m, v, n, sumsq_dev, clip_norm, lr_t,
              beta1, beta2, eps, weight_decay, params_tf32);
  return 0;
}

int avsr_round_tf32(avsr_stream_t s, const float* src, float* dst, long long n) {
  if (n <= 0) return 0;
  AVSR_LAUNCH(round_tf32_kernel, cdiv(n, 256), 256, 0, ST(s), src, dst, n);
  return 0;
}

int avsr_normed_v_fwd(avsr_stream_t s, const float* v, const float* g, int A, float* veff) {
  AVSR_LAUNCH(normed_v_fwd_kernel, 1, 256, 0, ST(s), v, g, A, veff);
  return 0;
}

int avsr_normed_v_bwd(avsr_stream_t s, const float* v, const float* g, const float* dveff, int A, float* dv,
                      float* dg) {
  AVSR_LAUNCH(normed_v_bwd_kernel, 1, 256, 0, ST(s), v, g, dveff, A, dv, dg);
  return 0;
}

int avsr_dropout(avsr_stream_t s, const float* x, long long n, long long first, const uint32_t* rng,
                 uint32_t stream_id, uint32_t thr, int round_out, float* y) {
  AVSR_REQUIRE(rng != nullptr && thr != 0u, "dropout: needs the rng words and a non-zero keep threshold");
  AVSR_REQUIRE(n >= 0 && first >= 0 && first + n <= (1ll << 32), "dropout: element indices must stay below 2^32");
  if (n == 0) return 0;
  if (((n | first) & 3) == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0) {
    int grid4 = cdiv(n / 4, 256);
    if (grid4 > 148 * 8) grid4 = 148 * 8;
    AVSR_LAUNCH(dropout_v4_kernel, grid4, 256, 0, ST(s), reinterpret_cast<const float4*>(x), n / 4, first, rng, stream_id,
                thr, inv_keep_of(thr), round_out && tensor_cores_enabled(), reinterpret_cast<float4*>(y));
    return 0;
  }
  int grid = cdiv(n, 256);
  if (grid > 148 * 16) grid = 148 * 16;
  AVSR_LAUNCH(dropout_kernel, grid, 256, 0, ST(s), x, n, first, rng, stream_id, thr, inv_keep_of(thr),
              round_out && tensor_cores_enabled(), y);
  return 0;
}

int avsr_sched_sample(avsr_stream_t s, const float* logits, int B, int V, const uint32_t* rng, uint32_t stream_id,
                      int t, uint32_t thr_p, const int* true_next, int* next_ids, int* sampled) {
  AVSR_REQUIRE(rng != nullptr && B > 0 && V > 0, "sched_sample: bad arguments");
  AVSR_LAUNCH(sched_sample_kernel, cdiv(B, 128), 128, 0, ST(s), logits, B, V, rng, stream_id, t, thr_p, true_next,
              next_ids, sampled);
  return 0;
}

static int grid_for(long long n) {
  long long g = (n + 255) / 256;
  return (int)(g < 1 ? 1 : (g > 148 * 32 ? 148 * 32 : g));
}

int avsr_im2col(avsr_stream_t s, const float* x, int N, int H, int W, int C, int kh, int kw, int stride, int pad_top,
                int pad_left, int Ho, int Wo, int round_out, float* cols) {
  AVSR_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && kh > 0 && kw > 0 && stride > 0 && Ho > 0 && Wo > 0, "im2col: bad geometry");
  AVSR_LAUNCH(im2col_kernel, grid_for((long long)N * Ho * Wo * kh * kw * C), 256, 0, ST(s), x, N, H, W, C, kh, kw, stride,
              pad_top, pad_left, Ho, Wo, round_out && tensor_cores_enabled(), cols);
  return 0;
}

int avsr_col2im(avsr_stream_t s, const float* dcols, int N, int H, int W, int C, int kh, int kw, int stride, int pad_top,
                int pad_left, int Ho, int Wo, float* dx) {
  AVSR_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && kh > 0 && kw > 0 && stride > 0 && Ho > 0 && Wo > 0, "col2im: bad geometry");
  AVSR_LAUNCH(col2im_kernel, grid_for((long long)N * H * W * C), 256, 0, ST(s), dcols, N, H, W, C, kh, kw, stride, pad_top,
              pad_left, Ho, Wo, dx);
  return 0;
}

int avsr_conv2d_direct(avsr_stream_t s, const float* x, int N, int H, int W, int Ci, const float* w, const float* bias,
                       int kh, int kw, int stride, int pad_top, int pad_left, int Ho, int Wo, int Co, float* y) {
  AVSR_REQUIRE(Co == 8 || Co == 16, "conv2d_direct: 8 or 16 output channels (got %d)", Co);
  AVSR_REQUIRE(N > 0 && H > 0 && W > 0 && Ci > 0 && kh > 0 && kw > 0 && stride > 0 && Ho > 0 && Wo > 0, "conv2d_direct: bad geometry");
  const size_t smem = (size_t)kh * kw * Ci * Co * sizeof(float);
  AVSR_REQUIRE(smem <= 48 * 1024, "conv2d_direct: kernel of %zu bytes does not fit shared memory", smem);
  const int grid = grid_for((long long)N * Ho * Wo);
  if (Co == 8)
    AVSR_LAUNCH(conv_direct_kernel<8>, grid, 256, smem, ST(s), x, N, H, W, Ci, w, bias, kh, kw, stride, pad_top, pad_left, Ho, Wo, y);
  else
    AVSR_LAUNCH(conv_direct_kernel<16>, grid, 256, smem, ST(s), x, N, H, W, Ci, w, bias, kh, kw, stride, pad_top, pad_left, Ho, Wo, y);
  return 0;
}

int avsr_conv2d_wgrad(avsr_stream_t s, const float* x, const float* dy, int N, int H, int W, int Ci, int kh, int kw,
                      int stride, int pad_top, int pad_left, int Ho, int Wo, int Co, float* dW) {
  AVSR_REQUIRE(Co == 8 || Co == 16, "conv2d_wgrad: 8 or 16 output channels (got %d)", Co);
  const int K = kh * kw * Ci;
  AVSR_REQUIRE(K * Co <= 9 * 256, "conv2d_wgrad: at most 2304 kernel entries (got %d)", K * Co);
  const size_t smem = (size_t)WG_PIX * ((K | 1) + Co) * sizeof(float);
  AVSR_REQUIRE(smem <= 48 * 1024, "conv2d_wgrad: tile of %zu bytes does not fit shared memory", smem);
  const long long tiles = cdiv((long long)N * Ho * Wo, WG_PIX);
  const int grid = (int)(tiles < 148 * 4 ? tiles : 148 * 4);
  if (Co == 8)
    AVSR_LAUNCH(conv_wgrad_kernel<8>, grid, 256, smem, ST(s), x, dy, N, H, W, Ci, kh, kw, stride, pad_top, pad_left, Ho, Wo, dW);
  else
    AVSR_LAUNCH(conv_wgrad_kernel<16>, grid, 256, smem, ST(s), x, dy, N, H, W, Ci, kh, kw, stride, pad_top, pad_left, Ho, Wo, dW);
  return 0;
}

int avsr_relu_fwd(avsr_stream_t s, const float* x, long long n, float* y) {
  if (n <= 0) return 0;
  AVSR_LAUNCH(relu_fwd_kernel, grid_for(n), 256, 0, ST(s), x, n, y);
  return 0;
}

int avsr_relu_bwd(avsr_stream_t s, const float* y, const float* dy, long long n, float* dx) {
  if (n <= 0) return 0;
  AVSR_LAUNCH(relu_bwd_kernel, grid_for(n), 256, 0, ST(s), y, dy, n, dx);
  return 0;
}

int avsr_selu_fwd(avsr_stream_t s, const float* x, long long n, float* y) {
  if (n <= 0) return 0;
  AVSR_LAUNCH(selu_fwd_kernel, grid_for(n), 256, 0, ST(s), x, n, y);
  return 0;
}

int avsr_selu_bwd(avsr_stream_t s, const float* y, const float* dy, long long n, float* dx) {
  if (n <= 0) return 0;
  AVSR_LAUNCH(selu_bwd_kernel, grid_for(n), 256, 0, ST(s), y, dy, n, dx);
  return 0;
}

int avsr_highway_fwd(avsr_stream_t s, const float* x, const float* pre, const float* out, long long n, float* y,
                     float* y_op) {
  if (n <= 0) return 0;
  AVSR_LAUNCH(highway_fwd_kernel, grid_for(n), 256, 0, ST(s), x, pre, out, n, tensor_cores_enabled(), y, y_op);
  return 0;
}

int avsr_highway_bwd(avsr_stream_t s, const float* dy, const float* x, const float* pre, const float* out, long long n,
                     float* dx, float* dout, float* dpre) {
  if (n <= 0) return 0;
  AVSR_LAUNCH(highway_bwd_kernel, grid_for(n), 256, 0, ST(s), dy, x, pre, out, n, tensor_cores_enabled(), dx, dout, dpre);
  return 0;
}

int avsr_greedy_pick(avsr_stream_t s, const float* logits, int B, int V, int eos, int* finished, int* sample_out,
                     int* next_ids) {
  AVSR_LAUNCH(greedy_pick_kernel, cdiv(B, 128), 128, 0, ST(s), logits, B, V, eos, finished, sample_out, next_ids);
  return 0;
}

int avsr_beam_step(avsr_stream_t s, const float* logits, int B, int W, int V, int eos, float length_penalty,
                   float* log_probs, int* finished, int* lengths, int* word_out, int* parent_out, float* score_out) {
  AVSR_REQUIRE(W >= 1 && W <= 32 && W <= V, "beam: width must be in [1, min(32, V)]");
  size_t smem = (size_t)(3 * W * V + 3 * W) * sizeof(float);
  AVSR_REQUIRE(smem <= 48 * 1024, "beam: W*V too large");
  AVSR_LAUNCH(beam_step_kernel, B, 128, smem, ST(s), logits, W, V, eos, length_penalty, log_probs, finished, lengths,
              word_out, parent_out, score_out);
  return 0;
}

int avsr_gather_rows(avsr_stream_t s, const float* src, const int* idx, long long n, int F, float* dst) {
  if (n <= 0) return 0;
  AVSR_LAUNCH(gather_rows_kernel, cdiv(n * F, 256), 256, 0, ST(s), src, idx, n, F, dst);
  return 0;
}

}  // extern "C"
