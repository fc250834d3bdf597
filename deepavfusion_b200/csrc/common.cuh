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

// erf to 1.5e-7 absolute error (Abramowitz & Stegun 7.1.26) with two MUFU ops (rcp, ex2): ~14
// instructions instead of ~25 for erff().  Its consumers round to bf16 (8 mantissa bits), so the
// result is indistinguishable from the exact-erf GELU of torch nn.GELU() / timm Mlp.
__device__ __forceinline__ float erf_fast(float x) {
  const float ax = fabsf(x);
  const float t = __frcp_rn(fmaf(0.3275911f, ax, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float e = exp2f(-1.4426950408889634f * ax * ax);
  return copysignf(fmaf(-p * t, e, 1.0f), x);
}
// (erf) GELU and its derivative -- torch nn.GELU() default, timm Mlp act_layer.
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erf_fast(x * 0.70710678118654752f)); }
__device__ __forceinline__ float dgelu_erf(float x) {
  const float cdf = 0.5f * (1.0f + erf_fast(x * 0.70710678118654752f));
  const float pdf = 0.39894228040143268f * exp2f(-0.72134752044448170f * x * x);
  return cdf + x * pdf;
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
