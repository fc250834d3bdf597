// davf_gemm: argument validation + dispatch to the tcgen05 kernels.
#include "gemm_epilogue.cuh"

using namespace davf;

namespace davf { int gemm_set_2cta(int on); int gemm_set_sms(int n); }
extern "C" int davf_set_gemm_sms(int n) { return davf::gemm_set_sms(n); }
extern "C" int davf_set_gemm_2cta(int on) { return davf::gemm_set_2cta(on); }

static int validate_gemm(const davf_gemm_args* a) {
  DAVF_CHECK_ARG(a && a->a && a->b && a->out, "gemm: null pointer");
  DAVF_CHECK_ARG(a->M >= 0 && a->N > 0 && a->K > 0, "gemm: bad sizes M=%lld N=%lld K=%lld", (long long)a->M, (long long)a->N, (long long)a->K);
  DAVF_CHECK_ARG(a->N % 8 == 0, "gemm: N=%lld must be a multiple of 8", (long long)a->N);
  DAVF_CHECK_ARG(a->lda % 8 == 0 && a->ldb % 8 == 0, "gemm: lda=%lld ldb=%lld must be multiples of 8 (16-byte TMA strides)", (long long)a->lda, (long long)a->ldb);
  DAVF_CHECK_ARG(((uintptr_t)a->a & 15) == 0 && ((uintptr_t)a->b & 15) == 0 && ((uintptr_t)a->out & 15) == 0, "gemm: operands must be 16-byte aligned");
  DAVF_CHECK_ARG(a->lda >= (a->a_kmajor ? a->K : a->M) && a->ldb >= (a->b_kmajor ? a->K : a->N), "gemm: leading dimension smaller than the row");
  DAVF_CHECK_ARG(a->ldo % 4 == 0 && a->ldo >= a->N, "gemm: ldo=%lld", (long long)a->ldo);
  DAVF_CHECK_ARG(a->act >= 0 && a->act <= 2, "gemm: act=%d", a->act);
  DAVF_CHECK_ARG(a->act != DAVF_ACT_DGELU || a->aux_in, "gemm: DGELU needs aux_in");
  DAVF_CHECK_ARG(!(a->aux_out || a->aux_in) || (a->ldaux % 4 == 0 && a->ldaux >= a->N), "gemm: ldaux=%lld", (long long)a->ldaux);
  DAVF_CHECK_ARG(!a->res || (a->ldres % 4 == 0 && a->ldres >= a->N), "gemm: ldres=%lld", (long long)a->ldres);
  DAVF_CHECK_ARG(!(a->accumulate && a->out_bf16), "gemm: accumulate needs an f32 output");
  DAVF_CHECK_ARG(a->split_k <= 1 || a->accumulate, "gemm: split_k > 1 needs accumulate");
  DAVF_CHECK_ARG(!a->accumulate || (a->act == DAVF_ACT_NONE && !a->res && !a->aux_out), "gemm: accumulate excludes act / res / aux_out");
  DAVF_CHECK_ARG(a->g == 0 || (a->g > 0 && a->G >= a->g && a->off >= 0 && a->off + a->g <= a->G), "gemm: bad row window g=%d G=%d off=%d", a->g, a->G, a->off);
  DAVF_CHECK_ARG(!a->rowsum_out || a->accumulate, "gemm: rowsum_out is only supported on accumulate (wgrad) launches");
  return DAVF_OK;
}

extern "C" int davf_gemm(const davf_gemm_args* a, davf_stream_t s) {
  if (int rc = validate_gemm(a)) return rc;
  if (a->M == 0) return DAVF_OK;
  return gemm_tc_launch(*a, as_stream(s));
}

extern "C" int davf_gemm_grouped(const davf_gemm_args* a, int count, davf_stream_t s) {
  DAVF_CHECK_ARG(a && count >= 1 && count <= DAVF_GEMM_MAX_GROUP, "gemm_grouped: count=%d (1..%d)", count, DAVF_GEMM_MAX_GROUP);
  davf_gemm_args live[DAVF_GEMM_MAX_GROUP];
  int n = 0;
  for (int p = 0; p < count; ++p) {
    if (int rc = validate_gemm(a + p)) return rc;
    DAVF_CHECK_ARG(a[p].a_kmajor == a[0].a_kmajor && a[p].b_kmajor == a[0].b_kmajor, "gemm_grouped: problem %d has a different operand-layout class", p);
    DAVF_CHECK_ARG(!a[p].debug_clocks, "gemm_grouped: debug_clocks is a single-launch aid");
    if (a[p].M > 0) live[n++] = a[p];
  }
  if (n == 0) return DAVF_OK;
  cudaStream_t st = as_stream(s);
  if (n == 1) return gemm_tc_launch(live[0], st);
  return gemm_tc_launch_grouped(live, n, st);
}
