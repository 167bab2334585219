// Exact-fp32 SIMT GEMM used for the small / odd-shaped products on the path
// (per-step recurrent products, N=31 logits, state projections) and as the
// exact-fp32 reference the tcgen05 TF32 kernel (gemm_tc.cu) is validated against.
// C[M,N] = beta*C + op(A) op(B) (+ bias), beta in {0,1}, optional split-K with
// fp32 atomics.
#include "common.cuh"

namespace avsr {

template <int BM, int BN, int BK, int TM, int TN, bool TA, bool TB>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
gemm_simt_kernel(int M, int N, int K, const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
                 float* __restrict__ C, int ldc, int accumulate, const float* __restrict__ bias, int splitk,
                 int round_out) {
  constexpr int NTHR = (BM / TM) * (BN / TN);
  constexpr int AE = BM * BK / NTHR, BE = BN * BK / NTHR;
  constexpr int RG = TM >= 4 ? 4 : TM, CG = TN >= 4 ? 4 : TN;
  constexpr int RSTR = BM / (TM / RG), CSTR = BN / (TN / CG);
  static_assert(BM * BK % NTHR == 0 && BN * BK % NTHR == 0, "tile/threads mismatch");
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];

  const int tid = threadIdx.x;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;  // M tiles on grid.x: no 65535 limit (im2col products: M = N*Ho*Wo)
  const int ktiles = (K + BK - 1) / BK;
  const int kt_per = (ktiles + splitk - 1) / splitk;
  const int kbeg = blockIdx.z * kt_per * BK;
  const int kend = min(K, kbeg + kt_per * BK);
  if (kbeg >= kend && blockIdx.z > 0) return;

  float ra[AE], rb[BE];
  auto gload = [&](int k0) {
#pragma unroll
    for (int i = 0; i < AE; ++i) {
      int e = tid + i * NTHR;
      int kk, mm;
      if (!TA) { kk = e % BK; mm = e / BK; } else { mm = e % BM; kk = e / BM; }
      int gm = m0 + mm, gk = k0 + kk;
      float v = 0.0f;
      if (gm < M && gk < kend) v = TA ? A[(size_t)gk * lda + gm] : A[(size_t)gm * lda + gk];
      ra[i] = v;
    }
#pragma unroll
    for (int i = 0; i < BE; ++i) {
      int e = tid + i * NTHR;
      int kk, nn;
      if (!TB) { nn = e % BN; kk = e / BN; } else { kk = e % BK; nn = e / BK; }
      int gn = n0 + nn, gk = k0 + kk;
      float v = 0.0f;
      if (gn < N && gk < kend) v = TB ? B[(size_t)gn * ldb + gk] : B[(size_t)gk * ldb + gn];
      rb[i] = v;
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int i = 0; i < AE; ++i) {
      int e = tid + i * NTHR;
      int kk, mm;
      if (!TA) { kk = e % BK; mm = e / BK; } else { mm = e % BM; kk = e / BM; }
      As[buf][kk][mm] = ra[i];
    }
#pragma unroll
    for (int i = 0; i < BE; ++i) {
      int e = tid + i * NTHR;
      int kk, nn;
      if (!TB) { nn = e % BN; kk = e / BN; } else { kk = e % BK; nn = e / BK; }
      Bs[buf][kk][nn] = rb[i];
    }
  };

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;

  int buf = 0;
  gload(kbeg);
  sstore(0);
  __syncthreads();
  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    const bool more = (k0 + BK) < kend;
    if (more) gload(k0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[buf][kk][(i / RG) * RSTR + ty * RG + (i % RG)];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs[buf][kk][(j / CG) * CSTR + tx * CG + (j % CG)];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (more) {
      sstore(buf ^ 1);
      __syncthreads();
      buf ^= 1;
    }
  }

#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int gm = m0 + (i / RG) * RSTR + ty * RG + (i % RG);
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int gn = n0 + (j / CG) * CSTR + tx * CG + (j % CG);
      if (gn >= N) continue;
      float v = acc[i][j];
      if (bias != nullptr && blockIdx.z == 0) v += bias[gn];
      float* p = C + (size_t)gm * ldc + gn;
      if (splitk > 1) atomicAdd(p, v);
      else *p = accumulate ? (*p + v) : maybe_tf32(v, round_out);
    }
  }
}

__global__ void zero_block_kernel(float* C, int M, int N, int ldc) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (size_t)M * N) C[(i / N) * ldc + (i % N)] = 0.0f;
}

template <int BM, int BN, int BK, int TM, int TN>
static int launch_cfg(cudaStream_t st, int tA, int tB, int M, int N, int K, const float* A, int lda, const float* B,
                      int ldb, float* C, int ldc, int acc, const float* bias, int splitk, int round_out) {
  dim3 grid(cdiv(M, BM), cdiv(N, BN), splitk);
  dim3 block((BM / TM) * (BN / TN));
  if (!tA && !tB) AVSR_LAUNCH((gemm_simt_kernel<BM, BN, BK, TM, TN, false, false>), grid, block, 0, st, M, N, K, A, lda, B, ldb, C, ldc, acc, bias, splitk, round_out);
  else if (!tA && tB) AVSR_LAUNCH((gemm_simt_kernel<BM, BN, BK, TM, TN, false, true>), grid, block, 0, st, M, N, K, A, lda, B, ldb, C, ldc, acc, bias, splitk, round_out);
  else if (tA && !tB) AVSR_LAUNCH((gemm_simt_kernel<BM, BN, BK, TM, TN, true, false>), grid, block, 0, st, M, N, K, A, lda, B, ldb, C, ldc, acc, bias, splitk, round_out);
  else AVSR_LAUNCH((gemm_simt_kernel<BM, BN, BK, TM, TN, true, true>), grid, block, 0, st, M, N, K, A, lda, B, ldb, C, ldc, acc, bias, splitk, round_out);
  return 0;
}

int gemm_simt(cudaStream_t st, int transA, int transB, int M, int N, int K, const float* A, int lda, const float* B,
              int ldb, float* C, int ldc, float beta, const float* bias, int round_out) {
  AVSR_REQUIRE(beta == 0.0f || beta == 1.0f, "gemm: beta must be 0 or 1 (got %f)", beta);
  AVSR_REQUIRE(!(round_out && beta != 0.0f), "gemm: round_out needs beta == 0");
  if (M <= 0 || N <= 0) return 0;
  int acc = beta == 1.0f;
  if (K <= 0) {
    AVSR_REQUIRE(bias == nullptr, "gemm: K==0 with bias unsupported");
    if (!acc) AVSR_LAUNCH(zero_block_kernel, cdiv((long long)M * N, 256), 256, 0, st, C, M, N, ldc);
    return 0;
  }
  const int target = 148 * 2;
  long long tilesL = (long long)cdiv(M, 128) * cdiv(N, 128);
  bool large = (M >= 96 && N >= 96) && (tilesL >= 64 || K >= 2048);
  int splitk = 1;
  if (large) {
    if (tilesL < target && K >= 1024) splitk = (int)min((long long)cdiv(K, 512), (long long)cdiv(target, tilesL));
  } else {
    long long tilesS = (long long)cdiv(M, 32) * cdiv(N, 64);
    if (tilesS < target && K >= 128) splitk = (int)min((long long)cdiv(K, 64), (long long)cdiv(target, tilesS));
  }
  if (round_out) splitk = 1;  // rounding needs the complete sum in one place
  if (splitk > 1 && !acc) AVSR_LAUNCH(zero_block_kernel, cdiv((long long)M * N, 256), 256, 0, st, C, M, N, ldc);
  if (large) return launch_cfg<128, 128, 8, 8, 8>(st, transA, transB, M, N, K, A, lda, B, ldb, C, ldc, acc, bias, splitk, round_out);
  return launch_cfg<32, 64, 16, 2, 4>(st, transA, transB, M, N, K, A, lda, B, ldb, C, ldc, acc, bias, splitk, round_out);
}

}  // namespace avsr
