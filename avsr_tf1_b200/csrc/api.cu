// C ABI glue: error state, launch counter, GEMM dispatch, recurrent op wrappers.
#include <stdarg.h>

#include "../../include/avsr_b200.h"
#include "common.cuh"

namespace avsr {

static thread_local char g_err[1024] = "";
unsigned long long g_launch_count = 0;
static int g_use_tc = 1;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int rnn_seq_fwd(cudaStream_t st, const AvsrRnnSeq* r);
int rnn_seq_bwd(cudaStream_t st, const AvsrRnnSeq* r);
size_t rnn_work_floats(int B, int H, int At, int maxHD, int maxA, int maxTm);
// tcgen05 TF32 path (gemm_tc.cu); returns -1 when the shape is not eligible
int gemm_tc(cudaStream_t st, int transA, int transB, int M, int N, int K, const float* A, int lda, const float* B,
            int ldb, float* C, int ldc, float beta, const float* bias, int round_out);
int tensor_cores_enabled() { return g_use_tc; }

int gemm(cudaStream_t st, int transA, int transB, int M, int N, int K, const float* A, int lda, const float* B,
         int ldb, float* C, int ldc, float beta, const float* bias, int round_out) {
  round_out = (round_out && g_use_tc) ? 1 : 0;
  if (g_use_tc) {
    int r = gemm_tc(st, transA, transB, M, N, K, A, lda, B, ldb, C, ldc, beta, bias, round_out);
    if (r >= 0) return r;
  }
  return gemm_simt(st, transA, transB, M, N, K, A, lda, B, ldb, C, ldc, beta, bias, round_out);
}

}  // namespace avsr

using namespace avsr;

extern "C" {

const char* avsr_last_error(void) { return g_err; }
int avsr_version(void) { return 100; }
unsigned long long avsr_launch_count(void) { return g_launch_count; }
int avsr_set_tensor_cores(int enable) {
  int old = g_use_tc;
  g_use_tc = enable ? 1 : 0;
  return old;
}

int avsr_gemm(avsr_stream_t s, int transA, int transB, int M, int N, int K, const float* A, int lda, const float* B,
              int ldb, float* C, int ldc, float beta, const float* bias, int round_out) {
  return gemm((cudaStream_t)s, transA, transB, M, N, K, A, lda, B, ldb, C, ldc, beta, bias, round_out);
}
int avsr_get_tensor_cores(void) { return g_use_tc; }

size_t avsr_rnn_work_floats(int B, int H, int At, int maxHD, int maxA, int maxTm) {
  return rnn_work_floats(B, H, At, maxHD, maxA, maxTm);
}
int avsr_struct_sizes(int* out2) {
  out2[0] = (int)sizeof(AvsrAttnMech);
  out2[1] = (int)sizeof(AvsrRnnSeq);
  return 0;
}
int avsr_rnn_seq_fwd(avsr_stream_t s, const AvsrRnnSeq* r) { return rnn_seq_fwd((cudaStream_t)s, r); }
int avsr_rnn_seq_bwd(avsr_stream_t s, const AvsrRnnSeq* r) { return rnn_seq_bwd((cudaStream_t)s, r); }

}  // extern "C"
