// Tensor-core convolutions of the visual front-end (avsr/video.py:17-31 conv2d_wrapper inside resnet_cnn :143-195;
// SURVEY.md 8f-3): implicit GEMM on `mma.sync.m16n8k8` TF32 with fp32 accumulation, NHWC, WHOLE FRAMES per CTA.
//
// The front-end's layers are narrow (8 .. 64 channels) over many pixels (19 200 frames x 36 x 36 per training step at
// batch 256): N = 8 output channels is exactly one n8 tile, K = 3 x 3 x Cin is a handful of k8 steps, and the operands
// are gathered, not tiled - so the register-fragment `mma.sync` path fits where a 128-row tcgen05 tile (canonical
// shared-memory layouts, i.e. an im2col pass in shared memory) does not.  What replaces the round-1 path
// (im2col + GEMM + col2im for the wide layers, SIMT direct kernels for the narrow ones, 148 ms per step):
//
//  * conv_mma_fwd_kernel: y = conv(x, w) (+ bias) (+ residual) (+ per-channel sum / sum-of-squares of y for the
//    following batch norm).  A CTA loads a group of F frames into shared memory ONCE, zero-padded (no bounds checks in
//    the inner loop) and split in planes of 4 channels (a lane's A-fragment loads are conflict-free), the weights
//    [kh*kw*Cin, Cout] next to them; a warp owns m16 tiles of 16 consecutive output pixels.  The same kernel is the
//    input gradient: of a stride-1 convolution with the kernel flipped / transposed by the host, and of a stride-2
//    convolution by ZERO-STUFFING dy while it is loaded (`ls` = 2: dy pixel (oy, ox) lands on (2 oy + pad, 2 ox + pad)).
//  * conv_mma_wgrad_kernel: dW[(ky, kx, ci), co] += sum over pixels of x[pixel shifted by (ky, kx)][ci] dy[pixel][co]:
//    M = kh*kw*Cin, N = Cout, K = pixels.  x and dy of a frame group sit in shared memory; warps split the (m, n)
//    tiles and the pixel range; partial tiles are merged with shared-memory atomics, one global atomic per entry and CTA.
//
// Operands are rounded to tf32 (round-to-nearest) while they are staged, like every other tensor-core operand of the path.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/avsr_b200.h"
#include "common.cuh"

namespace avsr {
namespace cv {

constexpr int THREADS = 256;

__device__ __forceinline__ void mma_tf32(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// exact n / d for n * d < 2^32 with m = ceil(2^32 / d) mod 2^32 (host: magic_of; d == 1 gives 0)
__device__ __forceinline__ int fdiv(int n, uint32_t m) { return m ? (int)__umulhi((uint32_t)n, m) : n; }  // m == 0: d == 1
struct Magic {
  uint32_t hw, w, hwo, wo, c;  // divisors H*W, W, Ho*Wo, Wo, channel groups of the loader (Cin / 4, or Cin)
};

struct ConvP {
  Magic mg;
  const float* x;      // [N, H, W, Cin]
  const float* w;      // [KS*KS*Cin, cout_total]
  const float* bias;   // [cout_total] or null
  const float* res;    // [N, Ho, Wo, cout_total] or null: y += res
  float* y;            // [N, Ho, Wo, cout_total]
  float* stats;        // [2*cout_total] += (sum, sum of squares) of y, or null; with mask_u: (sum d, sum d * xhat)
  const float* in_bn;  // [2*Cin] (scale, shift) or null: x is read as relu(x * scale + shift) - the batch_norm_relu between
                       // the producer of x and this convolution (video.py:4-15), applied while the frames are staged
  const float* res_bn; // [2*cout_total] or null: the residual is relu(res * scale + shift)
  const float* mask_u; // [N, Ho, Wo, cout_total] or null: backward of a batch_norm_relu whose INPUT was mask_u: the result v
  const float* mask_bn;// [4*cout_total] (scale, shift, a, b): d = v where mask_u * scale + shift > 0, else 0; y = d and
                       // stats += (sum d, sum d * (mask_u * a + b)) per channel (xhat = u * a + b)
  int N, H, W, Cin;
  int Ho, Wo, pt, pl;  // output size; padding of the gather (top / left)
  int ls;              // load stride: input pixel (iy, ix) sits at padded (ls*iy + pt, ls*ix + pl); 2 = zero-stuffed
  int Hp, Wp;          // padded frame in shared memory
  int F;               // frames per group
  int PS;              // words per channel plane (4 channels x F*Hp*Wp, padded so that PS % 32 == 16)
  int cout_total;
};

// shared-memory word of channel c of padded pixel `pix` of the group
__device__ __forceinline__ int xword(int PS, int pix, int c) { return (c >> 2) * PS + pix * 4 + (c & 3); }

// stage frames [f0, f0 + nf) of x into the channel planes (tf32-rounded).  The halo / stuffing zeros were written once.
template <int CINP>
__device__ __forceinline__ void load_frames(float* sX, const float* __restrict__ x, int f0, int nf, int H, int W, int Cin,
                                            int ls, int pt, int pl, int Hp, int Wp, int PS,
                                            const float* __restrict__ in_bn, const Magic mg) {
  const int tid = threadIdx.x;
  const int hw = H * W;
  if ((Cin & 3) == 0) {
    const int c4n = Cin >> 2;
    const int total = nf * hw * c4n;
    const float4* src = reinterpret_cast<const float4*>(x + (size_t)f0 * hw * Cin);
    constexpr int U = 8;  // loads in flight per thread (HBM latency x bandwidth needs ~32 KB per CTA in flight)
    for (int i0 = tid; i0 < total; i0 += U * THREADS) {
      float4 v[U];
#pragma unroll
      for (int j = 0; j < U; ++j) {
        const int i = i0 + j * THREADS;
        if (i < total) v[j] = __ldg(src + i);
      }
#pragma unroll
      for (int j = 0; j < U; ++j) {
        const int i = i0 + j * THREADS;
        if (i < total) {
          const int pg = fdiv(i, mg.c), c4 = i - pg * c4n;
          const int f = fdiv(pg, mg.hw), r = pg - f * hw, iy = fdiv(r, mg.w), ix = r - iy * W;
          const int pix = (f * Hp + ls * iy + pt) * Wp + ls * ix + pl;
          float4 t = v[j];
          if (in_bn) {
            const float4 sc = __ldg(reinterpret_cast<const float4*>(in_bn) + c4);
            const float4 sh = __ldg(reinterpret_cast<const float4*>(in_bn + Cin) + c4);
            t.x = fmaxf(fmaf(t.x, sc.x, sh.x), 0.0f); t.y = fmaxf(fmaf(t.y, sc.y, sh.y), 0.0f);
            t.z = fmaxf(fmaf(t.z, sc.z, sh.z), 0.0f); t.w = fmaxf(fmaf(t.w, sc.w, sh.w), 0.0f);
          }
          t.x = tf32_rn(t.x); t.y = tf32_rn(t.y); t.z = tf32_rn(t.z); t.w = tf32_rn(t.w);
          *reinterpret_cast<float4*>(sX + c4 * PS + pix * 4) = t;
        }
      }
    }
  } else {
    const int total = nf * hw * Cin;
    const float* src = x + (size_t)f0 * hw * Cin;
    for (int i = tid; i < total; i += THREADS) {
      const int pg = fdiv(i, mg.c), c = i - pg * Cin;
      const int f = fdiv(pg, mg.hw), r = pg - f * hw, iy = fdiv(r, mg.w), ix = r - iy * W;
      const int pix = (f * Hp + ls * iy + pt) * Wp + ls * ix + pl;
      float v = __ldg(src + i);
      if (in_bn) v = fmaxf(fmaf(v, __ldg(in_bn + c), __ldg(in_bn + Cin + c)), 0.0f);
      sX[xword(PS, pix, c)] = tf32_rn(v);
    }
  }
}

// CINP: channels of x padded to 4 / 8 / 16 / 32 / 64; NT: n8 tiles of output channels per CTA (blockIdx.y picks the chunk);
// KS: kernel size (square); S: stride of the gather
template <int CINP, int NT, int KS, int S>
__global__ void __launch_bounds__(THREADS) conv_mma_fwd_kernel(const ConvP p) {
  constexpr int K = KS * KS * CINP;
  constexpr int KSTEPS = (K + 7) / 8;
  constexpr int WS = NT == 1 ? 8 : NT * 8 + 8;  // weight row stride in words: B-fragment loads conflict-free
  extern __shared__ float smem[];
  float* sX = smem;
  float* sW = smem + (CINP / 4) * p.PS;
  float* sSt = sW + KSTEPS * 8 * WS;  // [2][NT*8] statistics of the CTA
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gid = lane >> 2, tig = lane & 3;
  const int n0 = blockIdx.y * NT * 8;
  const int PS = p.PS, Wp = p.Wp;

  for (int i = tid; i < (CINP / 4) * PS; i += THREADS) sX[i] = 0.0f;
  for (int i = tid; i < KSTEPS * 8 * WS; i += THREADS) {
    const int k = i / WS, n = i - k * WS;
    const int tap = k / CINP, ci = k - tap * CINP;
    float v = 0.0f;
    if (tap < KS * KS && ci < p.Cin && n < NT * 8) v = tf32_rn(p.w[(size_t)(tap * p.Cin + ci) * p.cout_total + n0 + n]);
    sW[i] = v;
  }
  if (tid < 2 * NT * 8) sSt[tid] = 0.0f;
  float st_s[NT][2], st_q[NT][2];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) st_s[nt][0] = st_s[nt][1] = st_q[nt][0] = st_q[nt][1] = 0.0f;
  float bia[NT][2];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    bia[nt][0] = p.bias ? p.bias[n0 + nt * 8 + 2 * tig] : 0.0f;
    bia[nt][1] = p.bias ? p.bias[n0 + nt * 8 + 2 * tig + 1] : 0.0f;
  }
  float rsc[NT][2], rsh[NT][2];          // residual batch_norm_relu
  float msc[NT][2], msh[NT][2], mxa[NT][2], mxb[NT][2];  // batch_norm_relu whose backward is fused into the epilogue
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int c = n0 + nt * 8 + 2 * tig + j;
      rsc[nt][j] = p.res_bn ? p.res_bn[c] : 1.0f;
      rsh[nt][j] = p.res_bn ? p.res_bn[p.cout_total + c] : 0.0f;
      msc[nt][j] = p.mask_bn ? p.mask_bn[c] : 0.0f;
      msh[nt][j] = p.mask_bn ? p.mask_bn[p.cout_total + c] : 0.0f;
      mxa[nt][j] = p.mask_bn ? p.mask_bn[2 * p.cout_total + c] : 0.0f;
      mxb[nt][j] = p.mask_bn ? p.mask_bn[3 * p.cout_total + c] : 0.0f;
    }
  // Narrow layers (one n8 tile, K <= 80: the 36 x 36 stages, most of the front-end's pixels): the B fragments of ALL k-steps
  // fit 2 KSTEPS registers and are the same for every tile, so they are read from shared memory once per CTA instead of
  // once per tile - a third of the shared-memory traffic of the product loop (4 A words + 2 B words per MMA before).
  constexpr bool BREG = (NT == 1 && KSTEPS <= 10);
  uint32_t breg[BREG ? KSTEPS : 1][2];
  if constexpr (BREG) {
    __syncthreads();  // sW is complete
#pragma unroll
    for (int ks = 0; ks < KSTEPS; ++ks) {
      breg[ks][0] = __float_as_uint(sW[(8 * ks + tig) * WS + gid]);
      breg[ks][1] = __float_as_uint(sW[(8 * ks + tig + 4) * WS + gid]);
    }
  }
  const int hwo = p.Ho * p.Wo;
  const int ngroups = (p.N + p.F - 1) / p.F;
  for (int grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
    const int f0 = grp * p.F;
    const int nf = min(p.F, p.N - f0);
    __syncthreads();  // the previous group's tiles are done with sX (and the zero fill is complete)
    load_frames<CINP>(sX, p.x, f0, nf, p.H, p.W, p.Cin, p.ls, p.pt, p.pl, p.Hp, Wp, PS, p.in_bn, p.mg);
    __syncthreads();
    const int pvalid = nf * hwo;
    const int ntiles = (pvalid + 15) >> 4;
    for (int t = warp; t < ntiles; t += THREADS / 32) {
      const int p0 = t * 16 + gid, p1 = p0 + 8;
      int base[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int pp = min(h ? p1 : p0, pvalid - 1);
        const int f = fdiv(pp, p.mg.hwo), r = pp - f * hwo, oy = fdiv(r, p.mg.wo), ox = r - oy * p.Wo;
        base[h] = ((f * p.Hp + S * oy) * Wp + S * ox) * 4;
      }
      float acc[NT][4];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.0f;
      // what the epilogue reads from HBM (residual, BN input of the mask) is requested before the products
      float2 rres[2][NT], rmu[2][NT];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int pp = h ? p1 : p0;
        const size_t row = ((size_t)f0 * hwo + min(pp, pvalid - 1)) * p.cout_total + n0 + 2 * tig;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          rres[h][nt] = p.res ? __ldg(reinterpret_cast<const float2*>(p.res + row + nt * 8)) : make_float2(0.0f, 0.0f);
          rmu[h][nt] = p.mask_u ? __ldg(reinterpret_cast<const float2*>(p.mask_u + row + nt * 8)) : make_float2(0.0f, 0.0f);
        }
      }
#pragma unroll
      for (int ks = 0; ks < KSTEPS; ++ks) {
        // K rows 8 ks + tig (a0, a1) and 8 ks + tig + 4 (a2, a3): (tap, channel) of each
        int wa, wb;
        if constexpr (CINP >= 8) {
          constexpr int CB = CINP / 8;
          const int tap = ks / CB, cb = ks - tap * CB;
          const int off = ((tap / KS) * Wp + (tap % KS)) * 4;
          wa = (2 * cb) * PS + off + tig;
          wb = (2 * cb + 1) * PS + off + tig;
        } else {  // CINP == 4: two taps per k-step
          const int ta = min(2 * ks, KS * KS - 1), tb = min(2 * ks + 1, KS * KS - 1);
          wa = ((ta / KS) * Wp + (ta % KS)) * 4 + tig;
          wb = ((tb / KS) * Wp + (tb % KS)) * 4 + tig;
        }
        const uint32_t a0 = __float_as_uint(sX[base[0] + wa]), a1 = __float_as_uint(sX[base[1] + wa]);
        const uint32_t a2 = __float_as_uint(sX[base[0] + wb]), a3 = __float_as_uint(sX[base[1] + wb]);
        if constexpr (BREG) {
          mma_tf32(acc[0], a0, a1, a2, a3, breg[ks][0], breg[ks][1]);
        } else {
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            const uint32_t b0 = __float_as_uint(sW[(8 * ks + tig) * WS + nt * 8 + gid]);
            const uint32_t b1 = __float_as_uint(sW[(8 * ks + tig + 4) * WS + nt * 8 + gid]);
            mma_tf32(acc[nt], a0, a1, a2, a3, b0, b1);
          }
        }
      }
      // epilogue: rows p0 (acc 0, 1) and p1 (acc 2, 3), columns 2 tig, 2 tig + 1 of every n tile
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int pp = h ? p1 : p0;
        if (pp < pvalid) {
          const size_t row = ((size_t)f0 * hwo + pp) * p.cout_total + n0 + 2 * tig;
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            float v0 = acc[nt][2 * h] + bia[nt][0], v1 = acc[nt][2 * h + 1] + bia[nt][1];
            if (p.res) {
              const float2 r2 = rres[h][nt];
              if (p.res_bn) {
                v0 += fmaxf(fmaf(r2.x, rsc[nt][0], rsh[nt][0]), 0.0f);
                v1 += fmaxf(fmaf(r2.y, rsc[nt][1], rsh[nt][1]), 0.0f);
              } else {
                v0 += r2.x;
                v1 += r2.y;
              }
            }
            if (p.mask_u) {
              const float2 u2 = rmu[h][nt];
              v0 = fmaf(u2.x, msc[nt][0], msh[nt][0]) > 0.0f ? v0 : 0.0f;
              v1 = fmaf(u2.y, msc[nt][1], msh[nt][1]) > 0.0f ? v1 : 0.0f;
              st_s[nt][0] += v0; st_s[nt][1] += v1;
              st_q[nt][0] = fmaf(v0, fmaf(u2.x, mxa[nt][0], mxb[nt][0]), st_q[nt][0]);
              st_q[nt][1] = fmaf(v1, fmaf(u2.y, mxa[nt][1], mxb[nt][1]), st_q[nt][1]);
            } else {
              st_s[nt][0] += v0; st_s[nt][1] += v1;
              st_q[nt][0] = fmaf(v0, v0, st_q[nt][0]); st_q[nt][1] = fmaf(v1, v1, st_q[nt][1]);
            }
            *reinterpret_cast<float2*>(p.y + row + nt * 8) = make_float2(v0, v1);
          }
        }
      }
    }
  }
  if (p.stats) {
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        float s = st_s[nt][j], q = st_q[nt][j];
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
          s += __shfl_xor_sync(0xffffffffu, s, o);
          q += __shfl_xor_sync(0xffffffffu, q, o);
        }
        if (gid == 0) {
          atomicAdd(&sSt[nt * 8 + 2 * tig + j], s);
          atomicAdd(&sSt[NT * 8 + nt * 8 + 2 * tig + j], q);
        }
      }
    __syncthreads();
    if (tid < NT * 8) {
      atomicAdd(p.stats + n0 + tid, sSt[tid]);
      atomicAdd(p.stats + p.cout_total + n0 + tid, sSt[NT * 8 + tid]);
    }
  }
}

struct WgradP {
  Magic mg;
  const float* x;   // [N, H, W, Cin]
  const float* dy;  // [N, Ho, Wo, cout_total]
  float* dW;        // [KS*KS*Cin, cout_total] +=
  float* dbias;     // [cout_total] += column sums of dy (the bias gradient), or null
  const float* in_bn;  // [2*Cin] or null: x is read as relu(x * scale + shift) (see ConvP)
  int N, H, W, Cin, Ho, Wo, pt, pl, Hp, Wp, F, PS, cout_total;
  int PK;           // pixels of a full group rounded up to 8
};

// TSPLIT: the warps split the m16 tiles TSPLIT ways and the pixel range 8 / TSPLIT ways
template <int CINP, int NT, int KS, int S, int TSPLIT>
__global__ void __launch_bounds__(THREADS) conv_mma_wgrad_kernel(const WgradP p) {
  constexpr int KW = KS * KS * CINP;
  constexpr int MT = (KW + 15) / 16;
  constexpr int MPW = (MT + TSPLIT - 1) / TSPLIT;
  constexpr int KSPLIT = (THREADS / 32) / TSPLIT;
  constexpr int DS = NT == 1 ? 8 : NT * 8 + 8;
  extern __shared__ float smem[];
  float* sX = smem;
  float* sDy = smem + (CINP / 4) * p.PS;   // [PK][DS]
  float* sAcc = sDy + p.PK * DS;           // [MT*16][NT*8] (only when the pixel range is split)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gid = lane >> 2, tig = lane & 3;
  const int ts = warp % TSPLIT, ksub = warp / TSPLIT;
  const int n0 = blockIdx.y * NT * 8;
  const int PS = p.PS, Wp = p.Wp;

  for (int i = tid; i < (CINP / 4) * PS; i += THREADS) sX[i] = 0.0f;
  if constexpr (KSPLIT > 1)
    for (int i = tid; i < MT * 16 * NT * 8; i += THREADS) sAcc[i] = 0.0f;
  float acc[MPW][NT][4];
#pragma unroll
  for (int i = 0; i < MPW; ++i)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) acc[i][nt][0] = acc[i][nt][1] = acc[i][nt][2] = acc[i][nt][3] = 0.0f;
  // rows gid and gid + 8 of m tile mt: K index = 16 mt + row -> (tap, channel); word offsets of the two rows
  int offA[MPW], offB[MPW];
#pragma unroll
  for (int i = 0; i < MPW; ++i) {
    const int mt = ts + i * TSPLIT;
    const int ka = 16 * mt + gid, kb = ka + 8;
    const int ta = min(ka / CINP, KS * KS - 1), tb = min(kb / CINP, KS * KS - 1);
    const int ca = ka % CINP, cb = kb % CINP;
    offA[i] = (ca >> 2) * PS + ((ta / KS) * Wp + (ta % KS)) * 4 + (ca & 3);
    offB[i] = (cb >> 2) * PS + ((tb / KS) * Wp + (tb % KS)) * 4 + (cb & 3);
  }
  float4 bsum = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  const int hwo = p.Ho * p.Wo;
  const int ngroups = (p.N + p.F - 1) / p.F;
  for (int grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
    const int f0 = grp * p.F;
    const int nf = min(p.F, p.N - f0);
    __syncthreads();
    load_frames<CINP>(sX, p.x, f0, nf, p.H, p.W, p.Cin, 1, p.pt, p.pl, p.Hp, Wp, PS, p.in_bn, p.mg);
    const int pvalid = nf * hwo;
    const int pk = (pvalid + 7) & ~7;
    {  // dy of the group: [pixel][this CTA's NT*8 channels], rows past the last pixel zero; a thread always meets the
       // same 4 channels (THREADS % C4 == 0), so the bias gradient is summed on the way
      constexpr int C4 = NT * 2;
      const float* src = p.dy + (size_t)f0 * hwo * p.cout_total + n0;
      for (int i = tid; i < pk * C4; i += THREADS) {
        const int pix = i / C4, c4 = i - pix * C4;
        float4 v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (pix < pvalid) v = __ldg(reinterpret_cast<const float4*>(src + (size_t)pix * p.cout_total) + c4);
        bsum.x += v.x; bsum.y += v.y; bsum.z += v.z; bsum.w += v.w;
        v.x = tf32_rn(v.x); v.y = tf32_rn(v.y); v.z = tf32_rn(v.z); v.w = tf32_rn(v.w);
        *reinterpret_cast<float4*>(sDy + pix * DS + c4 * 4) = v;
      }
    }
    __syncthreads();
    for (int kt = ksub; kt < (pk >> 3); kt += KSPLIT) {
      const int pa = kt * 8 + tig, pb = pa + 4;
      int base[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int pp = min(h ? pb : pa, pvalid - 1);
        const int f = fdiv(pp, p.mg.hwo), r = pp - f * hwo, oy = fdiv(r, p.mg.wo), ox = r - oy * p.Wo;
        base[h] = ((f * p.Hp + S * oy) * Wp + S * ox) * 4;
      }
      uint32_t b0[NT], b1[NT];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        b0[nt] = __float_as_uint(sDy[pa * DS + nt * 8 + gid]);
        b1[nt] = __float_as_uint(sDy[pb * DS + nt * 8 + gid]);
      }
#pragma unroll
      for (int i = 0; i < MPW; ++i) {
        if (ts + i * TSPLIT < MT) {
          const uint32_t a0 = __float_as_uint(sX[base[0] + offA[i]]), a1 = __float_as_uint(sX[base[0] + offB[i]]);
          const uint32_t a2 = __float_as_uint(sX[base[1] + offA[i]]), a3 = __float_as_uint(sX[base[1] + offB[i]]);
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) mma_tf32(acc[i][nt], a0, a1, a2, a3, b0[nt], b1[nt]);
        }
      }
    }
  }
  if (p.dbias) {  // threads tid, tid + C4, ... hold partial sums of the same 4 channels: lanes first, then one atomic per warp
    constexpr int C4 = NT * 2;
#pragma unroll
    for (int o = C4; o < 32; o <<= 1) {
      bsum.x += __shfl_xor_sync(0xffffffffu, bsum.x, o);
      bsum.y += __shfl_xor_sync(0xffffffffu, bsum.y, o);
      bsum.z += __shfl_xor_sync(0xffffffffu, bsum.z, o);
      bsum.w += __shfl_xor_sync(0xffffffffu, bsum.w, o);
    }
    if (lane < C4) {
      float* dst = p.dbias + n0 + lane * 4;
      atomicAdd(dst, bsum.x); atomicAdd(dst + 1, bsum.y); atomicAdd(dst + 2, bsum.z); atomicAdd(dst + 3, bsum.w);
    }
  }
  // D rows gid (acc 0, 1) and gid + 8 (acc 2, 3), columns 2 tig, 2 tig + 1
  if constexpr (KSPLIT > 1) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < MPW; ++i) {
      const int mt = ts + i * TSPLIT;
      if (mt < MT) {
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int e = 0; e < 4; ++e)
            atomicAdd(&sAcc[(16 * mt + gid + 8 * (e >> 1)) * (NT * 8) + nt * 8 + 2 * tig + (e & 1)], acc[i][nt][e]);
      }
    }
    __syncthreads();
    for (int i = tid; i < MT * 16 * NT * 8; i += THREADS) {
      const int k = i / (NT * 8), n = i - k * (NT * 8);
      const int tap = k / CINP, ci = k - tap * CINP;
      if (tap < KS * KS && ci < p.Cin) atomicAdd(p.dW + (size_t)(tap * p.Cin + ci) * p.cout_total + n0 + n, sAcc[i]);
    }
  } else {
#pragma unroll
    for (int i = 0; i < MPW; ++i) {
      const int mt = ts + i * TSPLIT;
      if (mt < MT) {
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int k = 16 * mt + gid + 8 * (e >> 1);
            const int tap = k / CINP, ci = k - tap * CINP;
            if (tap < KS * KS && ci < p.Cin)
              atomicAdd(p.dW + (size_t)(tap * p.Cin + ci) * p.cout_total + n0 + nt * 8 + 2 * tig + (e & 1), acc[i][nt][e]);
          }
      }
    }
  }
}

// batch_norm_relu (video.py:4-15) between two convolutions never materialises: the producer accumulates (sum, sum of
// squares) per channel, this kernel turns them into the per-channel coefficients the consumers apply while they load:
// coef = [scale = gamma * invstd | shift = beta - mean * scale | a = invstd | b = -mean * invstd]  (xhat = u a + b)
__global__ void bn_finalize_kernel(const float* __restrict__ sums, float inv_count, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float eps, float momentum, int C,
                                   float* __restrict__ moving_mean, float* __restrict__ moving_var, float* __restrict__ coef) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float mean = sums[c] * inv_count;
  const float var = fmaxf(sums[C + c] * inv_count - mean * mean, 0.0f);
  const float is = rsqrtf(var + eps);
  const float sc = gamma[c] * is;
  coef[c] = sc;
  coef[C + c] = beta[c] - mean * sc;
  coef[2 * C + c] = is;
  coef[3 * C + c] = -mean * is;
  if (moving_mean) moving_mean[c] = moving_mean[c] * momentum + mean * (1.0f - momentum);
  if (moving_var) moving_var[c] = moving_var[c] * momentum + var * (1.0f - momentum);
}
// same coefficients from the moving statistics (inference)
__global__ void bn_coef_eval_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                                    const float* __restrict__ moving_mean, const float* __restrict__ moving_var, float eps, int C,
                                    float* __restrict__ coef) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float is = rsqrtf(moving_var[c] + eps);
  const float sc = gamma[c] * is;
  coef[c] = sc;
  coef[C + c] = beta[c] - moving_mean[c] * sc;
  coef[2 * C + c] = is;
  coef[3 * C + c] = -moving_mean[c] * is;
}
// backward of batch_norm_relu given d = dz masked by the ReLU (formed in the epilogue of the convolution that produced dz)
// and sums2 = (sum d, sum d * xhat): du = gamma * invstd * (d - sum_d / n - xhat * sum_dxhat / n) (+ residual), xhat = u a + b
__global__ void __launch_bounds__(256) bn_relu_bwd_apply_kernel(const float4* __restrict__ d, const float4* __restrict__ u,
                                                               const float* __restrict__ coef, const float* __restrict__ sums2,
                                                               float inv_count, const float4* __restrict__ res, long long n4,
                                                               int C, float4* __restrict__ du) {
  const long long stride = (long long)gridDim.x * blockDim.x;  // a multiple of C / 4: the channels of a thread are fixed
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int c = (int)((i * 4) % C);
  float s[4], a[4], b[4], m1[4], m2[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    s[e] = coef[c + e];  // gamma * invstd
    a[e] = coef[2 * C + c + e];
    b[e] = coef[3 * C + c + e];
    m1[e] = sums2[c + e] * inv_count;
    m2[e] = sums2[C + c + e] * inv_count;
  }
  for (; i < n4; i += stride) {
    const float4 dv = __ldcs(d + i), uv = __ldcs(u + i);
    float4 o;
    o.x = s[0] * (dv.x - m1[0] - fmaf(uv.x, a[0], b[0]) * m2[0]);
    o.y = s[1] * (dv.y - m1[1] - fmaf(uv.y, a[1], b[1]) * m2[1]);
    o.z = s[2] * (dv.z - m1[2] - fmaf(uv.z, a[2], b[2]) * m2[2]);
    o.w = s[3] * (dv.w - m1[3] - fmaf(uv.w, a[3], b[3]) * m2[3]);
    if (res) {
      const float4 r = __ldcs(res + i);
      o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
    }
    du[i] = o;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------
static int g_sm_count = 0;
static int sm_count() {
  if (!g_sm_count) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (g_sm_count <= 0) g_sm_count = 148;
  }
  return g_sm_count;
}
static int cin_padded(int Ci) { return Ci <= 4 ? 4 : Ci <= 8 ? 8 : Ci <= 16 ? 16 : Ci <= 32 ? 32 : Ci <= 64 ? 64 : 0; }
static int nt_of(int Co) { return (Co % 8) ? 0 : Co == 8 ? 1 : Co == 16 ? 2 : (Co % 32 == 0) ? 4 : 0; }

struct Plan {
  int cinp, nt, Hp, Wp, F, PS, PK;
  size_t smem;
  Magic mg;
};
static uint32_t magic_of(int d) { return (uint32_t)(((1ull << 32) + (uint64_t)d - 1) / (uint64_t)d); }
// frames per group: fill about `budget` bytes of shared memory with the padded planes (+ dy rows for the weight gradient)
static bool make_plan(int N, int H, int W, int Ci, int Ho, int Wo, int Co, int KS, int S, int ls, int pt, int pl, bool wgrad,
                      Plan* pp) {
  Plan q;
  q.cinp = cin_padded(Ci);
  q.nt = nt_of(Co);
  if (!q.cinp || !q.nt || !(KS == 1 || KS == 3) || !(S == 1 || S == 2)) return false;
  q.Hp = max(ls * (H - 1) + 1 + pt, S * (Ho - 1) + KS);
  q.Wp = max(ls * (W - 1) + 1 + pl, S * (Wo - 1) + KS);
  const int ws = q.nt == 1 ? 8 : q.nt * 8 + 8;
  // two CTAs per SM when a group of >= 8 m16 tiles (128 pixels) fits 100 KB; one CTA with up to 200 KB otherwise
  // (measured: three CTAs of one frame each are slower than two CTAs of two frames for the narrow layers)
  const size_t limit = 200 * 1024;
  size_t budget = 100 * 1024;
  const int mt = (KS * KS * q.cinp + 15) / 16;
  const bool split_pixels = mt * q.nt < 64;  // (TSPLIT < 8 in launch_wgrad: partial tiles merged in shared memory)
  const size_t fixed = wgrad ? (split_pixels ? (size_t)mt * 16 * q.nt * 8 * 4 : 0) + 64
                             : (size_t)((KS * KS * q.cinp + 7) / 8) * 8 * ws * 4 + 2 * q.nt * 8 * 4;
  const size_t per_frame = (size_t)q.cinp * q.Hp * q.Wp * 4 + (wgrad ? (size_t)Ho * Wo * ws * 4 : 0);
  const int fmin = (128 + Ho * Wo - 1) / (Ho * Wo);
  if (fixed + (size_t)fmin * per_frame > budget) budget = limit - 2048;
  int F = (int)((budget > fixed ? budget - fixed : 0) / per_frame);
  F = F < 1 ? 1 : F > 32 ? 32 : F;
  if (F > N) F = N;
  q.F = F;
  q.PS = ((F * q.Hp * q.Wp * 4 + 31) & ~31) + 16;
  q.PK = (F * Ho * Wo + 7) & ~7;
  q.smem = (size_t)(q.cinp / 4) * q.PS * 4 + fixed + (wgrad ? (size_t)q.PK * ws * 4 : 0);
  if (q.smem > limit) return false;
  q.mg.hw = magic_of(H * W); q.mg.w = magic_of(W); q.mg.hwo = magic_of(Ho * Wo); q.mg.wo = magic_of(Wo);
  q.mg.c = magic_of((Ci & 3) == 0 ? Ci / 4 : Ci);
  *pp = q;
  return true;
}

template <int CINP, int NT, int KS, int S>
static int launch_fwd(cudaStream_t st, const ConvP& p, const Plan& q) {
  auto kern = conv_mma_fwd_kernel<CINP, NT, KS, S>;
  AVSR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)q.smem));
  const int ngroups = (p.N + p.F - 1) / p.F;
  const int per_sm = max(1, min(4, (int)((220 * 1024) / (q.smem + 1024))));
  const int gx = min(ngroups, sm_count() * per_sm);
  kern<<<dim3(gx, p.cout_total / (NT * 8)), THREADS, q.smem, st>>>(p);
  ++g_launch_count;
  AVSR_CHECK_CUDA(cudaGetLastError());
  return 0;
}
template <int CINP, int NT, int KS, int S>
static int launch_wgrad(cudaStream_t st, const WgradP& p, const Plan& q) {
  // split the m tiles over the warps when there are many of them, the pixels otherwise
  constexpr int MT = (KS * KS * CINP + 15) / 16;
  constexpr int TSPLIT = MT * NT >= 64 ? 8 : MT * NT >= 24 ? 4 : MT * NT >= 8 ? 2 : 1;
  auto kern = conv_mma_wgrad_kernel<CINP, NT, KS, S, TSPLIT>;
  AVSR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)q.smem));
  const int ngroups = (p.N + p.F - 1) / p.F;
  const int per_sm = max(1, min(4, (int)((220 * 1024) / (q.smem + 1024))));
  const int gx = min(ngroups, sm_count() * per_sm);
  kern<<<dim3(gx, p.cout_total / (NT * 8)), THREADS, q.smem, st>>>(p);
  ++g_launch_count;
  AVSR_CHECK_CUDA(cudaGetLastError());
  return 0;
}

#define CV_DISPATCH_KS_S(FN, CINP, NT, ...)                                  \
  do {                                                                       \
    if (KS == 3 && S == 1) return FN<CINP, NT, 3, 1>(__VA_ARGS__);           \
    if (KS == 3 && S == 2) return FN<CINP, NT, 3, 2>(__VA_ARGS__);           \
    if (KS == 1 && S == 1) return FN<CINP, NT, 1, 1>(__VA_ARGS__);           \
    if (KS == 1 && S == 2) return FN<CINP, NT, 1, 2>(__VA_ARGS__);           \
  } while (0)
#define CV_DISPATCH_NT(FN, CINP, ...)                                        \
  do {                                                                       \
    if (q.nt == 1) CV_DISPATCH_KS_S(FN, CINP, 1, __VA_ARGS__);               \
    if (q.nt == 2) CV_DISPATCH_KS_S(FN, CINP, 2, __VA_ARGS__);               \
    if (q.nt == 4) CV_DISPATCH_KS_S(FN, CINP, 4, __VA_ARGS__);               \
  } while (0)
#define CV_DISPATCH(FN, ...)                                                 \
  do {                                                                       \
    if (q.cinp == 4) CV_DISPATCH_NT(FN, 4, __VA_ARGS__);                     \
    if (q.cinp == 8) CV_DISPATCH_NT(FN, 8, __VA_ARGS__);                     \
    if (q.cinp == 16) CV_DISPATCH_NT(FN, 16, __VA_ARGS__);                   \
    if (q.cinp == 32) CV_DISPATCH_NT(FN, 32, __VA_ARGS__);                   \
    if (q.cinp == 64) CV_DISPATCH_NT(FN, 64, __VA_ARGS__);                   \
  } while (0)

static int dispatch_fwd(cudaStream_t st, const ConvP& p, const Plan& q, int KS, int S) {
  CV_DISPATCH(launch_fwd, st, p, q);
  return -1;
}
static int dispatch_wgrad(cudaStream_t st, const WgradP& p, const Plan& q, int KS, int S) {
  CV_DISPATCH(launch_wgrad, st, p, q);
  return -1;
}

}  // namespace cv
}  // namespace avsr

using namespace avsr;

extern "C" int avsr_conv2d_tc_supported(int Ci, int Co, int kh, int kw, int stride) {
  return cv::cin_padded(Ci) && cv::nt_of(Co) && kh == kw && (kh == 1 || kh == 3) && (stride == 1 || stride == 2);
}

extern "C" int avsr_conv2d_tc(avsr_stream_t stream, const float* x, int N, int H, int W, int Ci, const float* w,
                              const float* bias, int kh, int kw, int stride, int pad_top, int pad_left, int Ho, int Wo, int Co,
                              int in_dilation, const float* in_bn, const float* residual, const float* res_bn,
                              const float* mask_u, const float* mask_bn, float* stats, float* y) {
  AVSR_REQUIRE(avsr_conv2d_tc_supported(Ci, Co, kh, kw, stride), "conv2d_tc: unsupported shape Ci=%d Co=%d k=%dx%d stride=%d", Ci,
               Co, kh, kw, stride);
  AVSR_REQUIRE(in_dilation == 1 || (in_dilation == 2 && stride == 1), "conv2d_tc: zero-stuffed input needs stride 1");
  AVSR_REQUIRE((mask_u == nullptr) == (mask_bn == nullptr), "conv2d_tc: mask_u and mask_bn come together");
  AVSR_REQUIRE(!res_bn || residual, "conv2d_tc: res_bn without a residual");
  if (N <= 0) return 0;
  cv::Plan q;
  AVSR_REQUIRE(cv::make_plan(N, H, W, Ci, Ho, Wo, Co, kh, stride, in_dilation, pad_top, pad_left, false, &q),
               "conv2d_tc: frame does not fit shared memory (H=%d W=%d Ci=%d)", H, W, Ci);
  cv::ConvP p;
  p.mg = q.mg;
  p.x = x; p.w = w; p.bias = bias; p.res = residual; p.y = y; p.stats = stats;
  p.in_bn = in_bn; p.res_bn = res_bn; p.mask_u = mask_u; p.mask_bn = mask_bn;
  p.N = N; p.H = H; p.W = W; p.Cin = Ci; p.Ho = Ho; p.Wo = Wo; p.pt = pad_top; p.pl = pad_left; p.ls = in_dilation;
  p.Hp = q.Hp; p.Wp = q.Wp; p.F = q.F; p.PS = q.PS; p.cout_total = Co;
  const int rc = cv::dispatch_fwd((cudaStream_t)stream, p, q, kh, stride);
  AVSR_REQUIRE(rc >= 0, "conv2d_tc: no kernel for this shape");
  return rc;
}

extern "C" int avsr_conv2d_wgrad_tc(avsr_stream_t stream, const float* x, const float* in_bn, const float* dy, int N, int H, int W,
                                    int Ci, int kh, int kw, int stride, int pad_top, int pad_left, int Ho, int Wo, int Co,
                                    float* dW, float* dbias) {
  AVSR_REQUIRE(avsr_conv2d_tc_supported(Ci, Co, kh, kw, stride), "conv2d_wgrad_tc: unsupported shape Ci=%d Co=%d k=%dx%d stride=%d",
               Ci, Co, kh, kw, stride);
  if (N <= 0) return 0;
  cv::Plan q;
  AVSR_REQUIRE(cv::make_plan(N, H, W, Ci, Ho, Wo, Co, kh, stride, 1, pad_top, pad_left, true, &q),
               "conv2d_wgrad_tc: frame does not fit shared memory (H=%d W=%d Ci=%d)", H, W, Ci);
  cv::WgradP p;
  p.mg = q.mg;
  p.x = x; p.dy = dy; p.dW = dW; p.dbias = dbias; p.in_bn = in_bn;
  p.N = N; p.H = H; p.W = W; p.Cin = Ci; p.Ho = Ho; p.Wo = Wo; p.pt = pad_top; p.pl = pad_left;
  p.Hp = q.Hp; p.Wp = q.Wp; p.F = q.F; p.PS = q.PS; p.cout_total = Co; p.PK = q.PK;
  const int rc = cv::dispatch_wgrad((cudaStream_t)stream, p, q, kh, stride);
  AVSR_REQUIRE(rc >= 0, "conv2d_wgrad_tc: no kernel for this shape");
  return rc;
}

extern "C" int avsr_bn_finalize(avsr_stream_t stream, const float* sums, double count, const float* gamma, const float* beta,
                                float eps, float momentum, int C, float* moving_mean, float* moving_var, float* coef) {
  AVSR_REQUIRE(C > 0 && count > 0, "bn_finalize: bad C / count");
  AVSR_LAUNCH(cv::bn_finalize_kernel, cdiv(C, 128), 128, 0, (cudaStream_t)stream, sums, (float)(1.0 / count), gamma, beta, eps,
              momentum, C, moving_mean, moving_var, coef);
  return 0;
}

extern "C" int avsr_bn_coef_eval(avsr_stream_t stream, const float* gamma, const float* beta, const float* moving_mean,
                                 const float* moving_var, float eps, int C, float* coef) {
  AVSR_REQUIRE(C > 0, "bn_coef_eval: bad C");
  AVSR_LAUNCH(cv::bn_coef_eval_kernel, cdiv(C, 128), 128, 0, (cudaStream_t)stream, gamma, beta, moving_mean, moving_var, eps, C,
              coef);
  return 0;
}

extern "C" int avsr_bn_relu_bwd_apply(avsr_stream_t stream, const float* d, const float* u, const float* coef,
                                      const float* sums2, double count, const float* residual, long long rows, int C, float* du) {
  AVSR_REQUIRE(C > 0 && C % 4 == 0 && 1024 % C == 0 && count > 0, "bn_relu_bwd_apply: C must divide 1024 and be a multiple of 4");
  if (rows <= 0) return 0;
  const long long n4 = rows * C / 4;
  const int blocks = (int)(n4 < 256 ? 1 : (n4 / 256 < 148 * 16 ? n4 / 256 : 148 * 16));
  AVSR_LAUNCH(cv::bn_relu_bwd_apply_kernel, blocks, 256, 0, (cudaStream_t)stream, reinterpret_cast<const float4*>(d),
              reinterpret_cast<const float4*>(u), coef, sums2, (float)(1.0 / count), reinterpret_cast<const float4*>(residual), n4,
              C, reinterpret_cast<float4*>(du));
  return 0;
}
