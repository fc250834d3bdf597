// K13/K14: fused multi-tensor AdamW over flat f32 buffers (replaces torch.optim.AdamW at
// train.py:93 / misc.py:126-130, the /accum_iter sweep misc.py:114-119, zero_grad, the global
// gradient norm misc.py:151-163 and the bf16 weight casts of autocast) in ONE pass.
// Bandwidth-bound: 28 B/param algorithmic (read p,g,m,v; write p,m,v) + 2 B bf16 shadow + 4 B zeroed g.
// Per-parameter-group hyper-parameters come from device tables (a group id per 64-element chunk
// and {lr, weight_decay} per group) so LR schedules never re-capture a CUDA graph and arbitrary
// groupings (weight-decay split, "pretrained" groups, layer-wise lr decay) cost nothing.
#include <cstdlib>
#include "common.cuh"

namespace davf {

constexpr int kChunk = 64;          // elements per group-table entry (== ParamStore.ALIGN)
constexpr int kFrozen = 255;

__global__ void __launch_bounds__(256) adamw_kernel(float4* __restrict__ p, float4* __restrict__ g, float4* __restrict__ m,
                                                    float4* __restrict__ v, uint2* __restrict__ pb, int64_t n4,
                                                    const uint8_t* __restrict__ chunk_group, const float* __restrict__ hp,
                                                    const float* __restrict__ scal, float beta1, float beta2, float eps,
                                                    int zero_grad, float* __restrict__ sumsq_out) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float scratch[32];
  const float bc1 = 1.0f - scal[0], bc2 = 1.0f - scal[1], gsc = scal[2];
  const float inv_sqrt_bc2 = rsqrtf(bc2);
  float ss = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const int gid = chunk_group[i / (kChunk / 4)];
    float4 G = g[i];
    float* gg = reinterpret_cast<float*>(&G);
    if (gid != kFrozen) {
      const float lr = hp[2 * gid], wd = hp[2 * gid + 1];
      const float step = lr / bc1, decay = 1.0f - lr * wd;
      float4 P = p[i], M = m[i], V = v[i];
      float* pp = reinterpret_cast<float*>(&P);
      float* mm = reinterpret_cast<float*>(&M);
      float* vv = reinterpret_cast<float*>(&V);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float gr = gg[c] * gsc;
        ss = fmaf(gr, gr, ss);
        mm[c] = beta1 * mm[c] + (1.0f - beta1) * gr;
        vv[c] = beta2 * vv[c] + (1.0f - beta2) * gr * gr;
        const float denom = sqrtf(vv[c]) * inv_sqrt_bc2 + eps;
        pp[c] = pp[c] * decay - step * (mm[c] / denom);
      }
      p[i] = P; m[i] = M; v[i] = V;
      if (pb) {
        uint2 o;
        o.x = pack_bf16x2(P.x, P.y);
        o.y = pack_bf16x2(P.z, P.w);
        pb[i] = o;
      }
    }
    if (zero_grad) g[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (sumsq_out) {
    ss = block_sum(ss, scratch);
    if (threadIdx.x == 0) atomicAdd(sumsq_out, ss);
  }
}

__global__ void __launch_bounds__(256) sumsq_kernel(const float4* __restrict__ g, int64_t n4, float* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float scratch[32];
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 x = g[i];
    acc += (x.x * x.x + x.y * x.y) + (x.z * x.z + x.w * x.w);
  }
  acc = block_sum(acc, scratch);
  if (threadIdx.x == 0) atomicAdd(out, acc);
}

}  // namespace davf

using namespace davf;

extern "C" int davf_adamw_step(float* p, float* g, float* m, float* v, davf_bf16* p_bf16, int64_t n,
                               const uint8_t* chunk_group, const float* hp, const float* scal, float beta1, float beta2,
                               float eps, int zero_grad, float* sumsq_out, davf_stream_t s) {
  DAVF_CHECK_ARG(p && g && m && v && hp && scal && chunk_group, "adamw_step: null pointer");
  DAVF_CHECK_ARG(n % kChunk == 0, "adamw_step: n=%lld must be a multiple of %d", (long long)n, kChunk);
  if (n == 0) return DAVF_OK;
  // Every CTA resident at once and half of each SM's thread slots left free: a bucket's AdamW runs beside the NEXT bucket's
  // NCCL all-reduce and the backward GEMMs.  With 16 CTAs per SM (two waves of eight) the block scheduler drained AdamW's
  // queue before it dispatched a single NCCL CTA, which put all-reduce and AdamW in series at the end of the step.
  static const int per_sm = [] { const char* e = getenv("DAVF_ADAMW_CTAS_PER_SM"); const int v = e ? atoi(e) : 4; return v < 1 ? 1 : (v > 4096 ? 4096 : v); }();
  int64_t blocks = (n / 4 + 255) / 256;
  if (blocks > (int64_t)per_sm * kNumSMs) blocks = (int64_t)per_sm * kNumSMs;
  DAVF_CUDA(launch_pdl(adamw_kernel, dim3((int)blocks), dim3(256), 0, as_stream(s), reinterpret_cast<float4*>(p), reinterpret_cast<float4*>(g), reinterpret_cast<float4*>(m),
                                                     reinterpret_cast<float4*>(v), reinterpret_cast<uint2*>(p_bf16), n / 4, chunk_group, hp,
                                                     scal, beta1, beta2, eps, zero_grad, sumsq_out));
  DAVF_LAUNCH_OK();
  return DAVF_OK;
}

extern "C" int davf_sumsq_f32(const float* g, int64_t n, float* out, davf_stream_t s) {
  DAVF_CHECK_ARG(g && out && n % 4 == 0, "sumsq: n=%lld must be a multiple of 4", (long long)n);
  if (n == 0) return DAVF_OK;
  int64_t blocks = (n / 4 + 255) / 256;
  if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
  DAVF_CUDA(launch_pdl(sumsq_kernel, dim3((int)blocks), dim3(256), 0, as_stream(s), reinterpret_cast<const float4*>(g), n / 4, out));
  DAVF_LAUNCH_OK();
  return DAVF_OK;
}
