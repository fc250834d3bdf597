// Shared helpers for libdavf_sm100.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/davf.h"

namespace davf {

void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launches;
// launches by kernel family (davf_launch_count_kind): 1 = CTA-pair tcgen05 GEMM, 2 = tcgen05 attention, 3 = mma.sync attention
enum { kKindGemm2Cta = 1, kKindAttnTc = 2, kKindAttnMma = 3, kNumKinds = 4 };
extern std::atomic<int64_t> g_launch_kind[kNumKinds];

#define DAVF_CHECK_ARG(cond, ...)                      \
  do {                                                 \
    if (!(cond)) {                                     \
      ::davf::set_error(__VA_ARGS__);                  \
      return DAVF_EINVAL;                              \
    }                                                  \
  } while (0)

#define DAVF_CUDA(call)                                                                   \
  do {                                                                                    \
    cudaError_t e__ = (call);                                                             \
    if (e__ != cudaSuccess) {                                                             \
      ::davf::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return DAVF_ECUDA;                                                                  \
    }                                                                                     \
  } while (0)

// call after every <<<>>> launch
#define DAVF_LAUNCH_OK()                   \
  do {                                     \
    ::davf::g_launches.fetch_add(1);       \
    DAVF_CUDA(cudaGetLastError());         \
  } while (0)

static inline cudaStream_t as_stream(davf_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------------------------
// The step is ~1,200 kernel launches averaging < 20 us on a handful of streams, so the per-launch fixed cost (launch
// latency, barrier init, TMEM allocation, tensor-map prefetch) is a double-digit share of it.  Every hot kernel therefore
// (1) executes griddepcontrol.launch_dependents first, which lets the NEXT kernel of its stream be scheduled as soon as SM
// resources free up, and (2) executes griddepcontrol.wait before its first global-memory access, which blocks until the
// previous kernel has completed and its writes are visible -- so a kernel's set-up overlaps its predecessor's tail while
// every memory dependency (RAW and WAR: nothing is read or written before the wait) is ordered exactly as without PDL.
// Launches carry cudaLaunchAttributeProgrammaticStreamSerialization; stream capture turns it into programmatic graph edges.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// Measured on B200 (profiles/r2): with the step's three-to-five concurrent streams PDL LOSES 2 % (16.55 -> 16.86 ms) -- a dependent
// kernel whose CTAs become resident early holds its SMs' shared memory while it waits, and those SMs are then not available to
// the READY kernels of the other streams.  The step is bound by the SM-exclusive time of the big kernels, not by launch
// latency.  The attribute is therefore OFF by default (davf_set_pdl(1) / DAVF_PDL=1 turns it on: single-stream callers).
extern std::atomic<int> g_pdl;

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_pdl.load(std::memory_order_relaxed) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

constexpr int kNumSMs = 148;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ float bf16_to_f32(uint16_t h) { return __uint_as_float(((uint32_t)h) << 16); }
__device__ __forceinline__ uint16_t f32_to_bf16(float f) {
  return __bfloat16_as_ushort(__float2bfloat16_rn(f));
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}

__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// Gaussian CDF Phi(x) = 0.5 (1 + erf(x / sqrt 2)) and e = exp(-x^2 / 2), from the Abramowitz & Stegun
// 7.1.26 rational form of erf (|error| <= 1.5e-7) with two MUFU ops (rcp, ex2): ~15 instructions
// instead of ~30 for erff() + expf().  Every consumer rounds to bf16 (8 mantissa bits), so GELU and
// its derivative are indistinguishable from the exact-erf torch nn.GELU() / timm Mlp activation.
__device__ __forceinline__ void gauss_terms(float x, float& cdf, float& e) {
  const float ax = fabsf(x);
  const float t = rcp_approx(fmaf(0.3275911f * 0.70710678118654752f, ax, 1.0f));
  float p = fmaf(0.5f * 1.061405429f, t, 0.5f * -1.453152027f);
  p = fmaf(p, t, 0.5f * 1.421413741f);
  p = fmaf(p, t, 0.5f * -0.284496736f);
  p = fmaf(p, t, 0.5f * 0.254829592f);
  e = ex2_approx(-0.72134752044448170f * x * x);
  const float h = p * t * e;                   // 0.5 * erfc(|x| / sqrt 2)
  cdf = x >= 0.f ? 1.0f - h : h;
}
// (erf) GELU and its derivative -- torch nn.GELU() default, timm Mlp act_layer.
__device__ __forceinline__ float gelu_erf(float x) {
  float cdf, e;
  gauss_terms(x, cdf, e);
  return x * cdf;
}
__device__ __forceinline__ float dgelu_erf(float x) {
  float cdf, e;
  gauss_terms(x, cdf, e);
  return fmaf(x * 0.39894228040143268f, e, cdf);
}

// gelu(x) and gelu'(x) from one evaluation of the Gaussian terms (x may alias g)
__device__ __forceinline__ void gelu_both(float x, float& g, float& dg) {
  float cdf, e;
  gauss_terms(x, cdf, e);
  dg = fmaf(x * 0.39894228040143268f, e, cdf);
  g = x * cdf;
}

// block-wide sum for blockDim.x <= 1024 (multiple of 32); scratch >= 32 floats
__device__ __forceinline__ float block_sum(float v, float* scratch) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[w] = v;
  __syncthreads();
  float r = (threadIdx.x < nw) ? scratch[threadIdx.x] : 0.f;
  if (w == 0) r = warp_sum(r);
  if (threadIdx.x == 0) scratch[0] = r;
  __syncthreads();
  r = scratch[0];
  return r;
}

}  // namespace davf
