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
int rnn_sampling_fused(const AvsrRnnSeq* r);
int rnn_seq_bwd(cudaStream_t st, const AvsrRnnSeq* r);
size_t rnn_work_floats(int B, int H, int At, int maxHD, int maxA, int maxTm);
// tcgen05 TF32 path (gemm_tc.cu); returns -1 when the shape is not eligible
int gemm_tc(cudaStream_t st, int transA, int transB, int M, int N, int K, const float* A, int lda, const float* B,
            int ldb, float* C, int ldc, float beta, const float* bias, int round_out);
int tensor_cores_enabled() { return g_use_tc; }

// ---- kernel timers ------------------------------------------------------------
namespace {
struct TimerSlot {
  cudaEvent_t e0, e1;
  int klass;
};
constexpr int MAX_SLOTS = 4096;
TimerSlot g_slots[MAX_SLOTS];
int g_nslots = 0;
bool g_timing = false;
}  // namespace
bool kernel_timing_enabled() { return g_timing; }
int kernel_timer_begin(cudaStream_t st, int klass) {
  if (!g_timing || g_nslots >= MAX_SLOTS) return -1;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) return -1;
  TimerSlot& s = g_slots[g_nslots];
  if (cudaEventCreate(&s.e0) != cudaSuccess || cudaEventCreate(&s.e1) != cudaSuccess) return -1;
  s.klass = klass;
  cudaEventRecord(s.e0, st);
  return g_nslots++;
}
void kernel_timer_end(cudaStream_t st, int slot) {
  if (slot >= 0) cudaEventRecord(g_slots[slot].e1, st);
}

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

int avsr_kernel_timing(int enable) {
  const int old = g_timing ? 1 : 0;
  for (int i = 0; i < g_nslots; ++i) {
    cudaEventDestroy(g_slots[i].e0);
    cudaEventDestroy(g_slots[i].e1);
  }
  g_nslots = 0;
  g_timing = enable != 0;
  return old;
}
int avsr_kernel_times(float* ms_out, int* launches_out) {
  for (int k = 0; k < AVSR_K_COUNT; ++k) {
    ms_out[k] = 0.0f;
    launches_out[k] = 0;
  }
  for (int i = 0; i < g_nslots; ++i) {
    AVSR_CHECK_CUDA(cudaEventSynchronize(g_slots[i].e1));
    float ms = 0.0f;
    AVSR_CHECK_CUDA(cudaEventElapsedTime(&ms, g_slots[i].e0, g_slots[i].e1));
    ms_out[g_slots[i].klass] += ms;
    launches_out[g_slots[i].klass] += 1;
  }
  return 0;
}

size_t avsr_rnn_work_floats(int B, int H, int At, int maxHD, int maxA, int maxTm) {
  return rnn_work_floats(B, H, At, maxHD, maxA, maxTm);
}
int avsr_struct_sizes(int* out2) {
  out2[0] = (int)sizeof(AvsrAttnMech);
  out2[1] = (int)sizeof(AvsrRnnSeq);
  return 0;
}
int avsr_rnn_sampling_fused(const AvsrRnnSeq* r) { return rnn_sampling_fused(r); }
int avsr_rnn_seq_fwd(avsr_stream_t s, const AvsrRnnSeq* r) { return rnn_seq_fwd((cudaStream_t)s, r); }
int avsr_rnn_seq_bwd(avsr_stream_t s, const AvsrRnnSeq* r) { return rnn_seq_bwd((cudaStream_t)s, r); }

}  // extern "C"
