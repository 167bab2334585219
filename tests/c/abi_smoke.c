/* Plain-C client of the C ABI (include/avsr_b200.h): no Python, no torch - only libavsr_b200.so and the CUDA runtime.
 * Runs one dense product with bias, one LSTM layer forward (avsr_rnn_seq_fwd, T = 5) and one dropout pass on cuda:0 and
 * checks them against loops written here.  Built and run by tests/test_gpu_cabi.py:
 *   gcc -std=c99 -O1 tests/c/abi_smoke.c -Iinclude -I/usr/local/cuda/include -Lavsr_tf1_b200/lib -lavsr_b200 \
 *       -L/usr/local/cuda/lib64 -lcudart -lm -o abi_smoke                                                          */
#include <cuda_runtime_api.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "avsr_b200.h"

#define CK(x)                                                                   \
  do {                                                                          \
    if ((x) != 0) {                                                             \
      fprintf(stderr, "%s failed: %s\n", #x, avsr_last_error());                \
      return 1;                                                                 \
    }                                                                           \
  } while (0)
#define CU(x)                                                                   \
  do {                                                                          \
    cudaError_t e_ = (x);                                                       \
    if (e_ != cudaSuccess) {                                                    \
      fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_));                  \
      return 1;                                                                 \
    }                                                                           \
  } while (0)

static float frand(unsigned* s) {
  *s = *s * 1664525u + 1013904223u;
  return ((*s >> 8) / 16777216.0f) * 2.0f - 1.0f;
}
static float sigm(float x) { return 1.0f / (1.0f + expf(-x)); }

int main(void) {
  enum { T = 5, B = 3, I = 7, H = 8 };
  unsigned seed = 42;
  int sizes[2];
  CK(avsr_struct_sizes(sizes));
  if (sizes[0] != (int)sizeof(AvsrAttnMech) || sizes[1] != (int)sizeof(AvsrRnnSeq)) {
    fprintf(stderr, "struct layout differs between this compiler and the library\n");
    return 1;
  }
  avsr_set_tensor_cores(0); /* exact fp32 path: bit-level comparable with the loops below */
  cudaStream_t st;
  CU(cudaStreamCreate(&st));

  /* ---- x[T*B, I] @ W[:I] + bias -> gates -------------------------------------------------- */
  float hx[T * B * I], hW[(I + H) * 4 * H], hb[4 * H], hgates[T * B * 4 * H];
  for (int i = 0; i < T * B * I; ++i) hx[i] = frand(&seed);
  for (int i = 0; i < (I + H) * 4 * H; ++i) hW[i] = 0.4f * frand(&seed);
  for (int i = 0; i < 4 * H; ++i) hb[i] = 0.1f * frand(&seed);
  float *dx, *dW, *db, *dgates;
  CU(cudaMalloc((void**)&dx, sizeof(hx)));
  CU(cudaMalloc((void**)&dW, sizeof(hW)));
  CU(cudaMalloc((void**)&db, sizeof(hb)));
  CU(cudaMalloc((void**)&dgates, sizeof(hgates)));
  CU(cudaMemcpy(dx, hx, sizeof(hx), cudaMemcpyHostToDevice));
  CU(cudaMemcpy(dW, hW, sizeof(hW), cudaMemcpyHostToDevice));
  CU(cudaMemcpy(db, hb, sizeof(hb), cudaMemcpyHostToDevice));
  CK(avsr_gemm(st, 0, 0, T * B, 4 * H, I, dx, I, dW, 4 * H, dgates, 4 * H, 0.0f, db, 0));
  CU(cudaMemcpyAsync(hgates, dgates, sizeof(hgates), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  double worst = 0.0;
  float ref_gates[T * B * 4 * H];
  for (int r = 0; r < T * B; ++r)
    for (int n = 0; n < 4 * H; ++n) {
      float acc = hb[n];
      for (int k = 0; k < I; ++k) acc += hx[r * I + k] * hW[k * 4 * H + n];
      ref_gates[r * 4 * H + n] = acc;
      double d = fabs((double)acc - hgates[r * 4 * H + n]);
      if (d > worst) worst = d;
    }
  printf("gemm+bias max abs error %.3g\n", worst);
  if (worst > 1e-5) return 1;

  /* ---- one LSTMCell layer under dynamic_rnn: lengths {5, 2, 4} ----------------------------- */
  int hlen[B] = {5, 2, 4};
  int* dlen;
  float *dS, *dcraw, *dout, *dcT, *dhT, *dwork;
  CU(cudaMalloc((void**)&dlen, sizeof(hlen)));
  CU(cudaMemcpy(dlen, hlen, sizeof(hlen), cudaMemcpyHostToDevice));
  CU(cudaMalloc((void**)&dS, (T + 1) * B * H * sizeof(float)));
  CU(cudaMemset(dS, 0, (T + 1) * B * H * sizeof(float)));
  CU(cudaMalloc((void**)&dcraw, T * B * H * sizeof(float)));
  CU(cudaMalloc((void**)&dout, T * B * H * sizeof(float)));
  CU(cudaMalloc((void**)&dcT, B * H * sizeof(float)));
  CU(cudaMalloc((void**)&dhT, B * H * sizeof(float)));
  size_t nwork = avsr_rnn_work_floats(B, H, 0, 0, 0, 0);
  CU(cudaMalloc((void**)&dwork, (nwork + 4) * sizeof(float)));
  AvsrRnnSeq r;
  memset(&r, 0, sizeof(r));
  r.T = T; r.B = B; r.H = H;
  r.len = dlen; r.gates = dgates; r.Wrec = dW + (size_t)I * 4 * H;
  r.S = dS; r.craw = dcraw; r.out = dout; r.cT = dcT; r.hT = dhT; r.work = dwork;
  CK(avsr_rnn_seq_fwd(st, &r));
  float hout[T * B * H], hhT[B * H];
  CU(cudaMemcpyAsync(hout, dout, sizeof(hout), cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(hhT, dhT, sizeof(hhT), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  worst = 0.0;
  for (int b = 0; b < B; ++b) {
    float c[H] = {0}, h[H] = {0};
    for (int t = 0; t < T; ++t) {
      float hn[H], cn[H];
      for (int u = 0; u < H; ++u) {
        float z[4];
        for (int g = 0; g < 4; ++g) {
          float acc = ref_gates[(t * B + b) * 4 * H + g * H + u];
          for (int k = 0; k < H; ++k) acc += h[k] * hW[(I + k) * 4 * H + g * H + u];
          z[g] = acc;
        }
        float cr = sigm(z[2] + 1.0f) * c[u] + sigm(z[0]) * tanhf(z[1]); /* gate order i, j, f, o; forget bias 1 */
        cn[u] = fminf(fmaxf(cr, -1.0f), 1.0f);                          /* cell_clip = 1 */
        hn[u] = sigm(z[3]) * tanhf(cn[u]);
      }
      for (int u = 0; u < H; ++u) {
        float want = t < hlen[b] ? hn[u] : 0.0f; /* zero output, carried state past the length */
        double d = fabs((double)want - hout[(t * B + b) * H + u]);
        if (d > worst) worst = d;
        if (t < hlen[b]) {
          h[u] = hn[u];
          c[u] = cn[u];
        }
      }
    }
    for (int u = 0; u < H; ++u) {
      double d = fabs((double)h[u] - hhT[b * H + u]);
      if (d > worst) worst = d;
    }
  }
  printf("lstm layer max abs error %.3g\n", worst);
  if (worst > 2e-5) return 1;

  /* ---- dropout: kept elements are scaled by 2^32 / thr, the rest are zero, about 80 % kept -- */
  uint32_t hrng[2] = {1234u, 5u}, *drng;
  CU(cudaMalloc((void**)&drng, sizeof(hrng)));
  CU(cudaMemcpy(drng, hrng, sizeof(hrng), cudaMemcpyHostToDevice));
  const uint32_t thr = 3435973836u; /* 0.8 * 2^32 */
  float* dy;
  CU(cudaMalloc((void**)&dy, sizeof(hx)));
  CK(avsr_dropout(st, dx, T * B * I, 0, drng, 9u, thr, 0, dy));
  float hy[T * B * I];
  CU(cudaMemcpyAsync(hy, dy, sizeof(hy), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  int kept = 0;
  for (int i = 0; i < T * B * I; ++i) {
    if (hy[i] != 0.0f) {
      ++kept;
      if (fabs(hy[i] - hx[i] * (float)(4294967296.0 / thr)) > 1e-6) return 1;
    }
  }
  printf("dropout kept %d of %d\n", kept, T * B * I);
  if (kept < T * B * I * 6 / 10 || kept == T * B * I) return 1;
  printf("launches %llu\nOK\n", avsr_launch_count());
  return 0;
}
