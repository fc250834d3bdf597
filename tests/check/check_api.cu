// Entry points of the checker library (tests only): the CUDA-core GEMM / attention kernels behind the product's argument structs.
#include <stdarg.h>
#include "gemm_epilogue.cuh"

namespace davf {
static thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};
std::atomic<int64_t> g_launch_kind[kNumKinds];
std::atomic<int> g_pdl{0};
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int attn_simt_fwd(const davf_attn_fwd_args* a, cudaStream_t st);
int attn_simt_bwd(const davf_attn_bwd_args* a, cudaStream_t st);
}  // namespace davf

using namespace davf;

extern "C" {
const char* davf_check_last_error(void) { return g_err; }
int davf_check_gemm(const davf_gemm_args* a, davf_stream_t s) {
  if (!a || a->M == 0) return DAVF_OK;
  return gemm_simt_launch(*a, as_stream(s));
}
int davf_check_gemm_grouped(const davf_gemm_args* a, int count, davf_stream_t s) {
  for (int p = 0; p < count; ++p)
    if (a[p].M > 0)
      if (int rc = gemm_simt_launch(a[p], as_stream(s))) return rc;
  return DAVF_OK;
}
int davf_check_attention_fwd(const davf_attn_fwd_args* a, davf_stream_t s) { return a->B == 0 ? DAVF_OK : attn_simt_fwd(a, as_stream(s)); }
int davf_check_attention_bwd(const davf_attn_bwd_args* a, davf_stream_t s) {
  if (a->B == 0) return DAVF_OK;
  if (a->dq_dead_rows > 0) {            // the checker zero-fills the dead query slots with a memset per sample
    for (int b = 0; b < a->B; ++b)
      for (int r = 0; r < a->dq_dead_rows; ++r)
        if (cudaMemsetAsync(a->dq + (int64_t)b * a->dq_bs + (int64_t)(r - a->dq_dead_rows) * a->dq_rs, 0, (size_t)a->H * a->dqk * 2, as_stream(s)) != cudaSuccess)
          return DAVF_ECUDA;
  }
  return attn_simt_bwd(a, as_stream(s));
}
}
