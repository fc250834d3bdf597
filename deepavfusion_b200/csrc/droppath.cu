// Stochastic depth (timm DropPath on the residual branches: timm Block.drop_path1/2 via models/vits.py:33,
// models/fusion_blocks.py:276,283,288).  Only the fine-tuning config uses it (configs/finetune.yaml:36,
// drop_path 0.2); with drop_path = 0 none of this is launched and the residual add stays fused in the GEMM epilogue.
//   forward : out[r, :] = res[r, :] + scale[r / rows_per_sample] * y[r, :]        (scale = keep-mask / keep_prob per sample)
//   backward: branch gradient = scale[r / rows_per_sample] * dy[r, :]  as bf16 (GEMM operand) and / or f32
#include "common.cuh"

namespace davf {

__global__ void scale_rows_add_kernel(const float4* __restrict__ res, const float4* __restrict__ y, const float* __restrict__ scale,
                                      int rps, int64_t rows, int D4, float4* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t n = rows * D4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float s = scale[(i / D4) / rps];
    const float4 a = res[i], b = y[i];
    out[i] = make_float4(fmaf(s, b.x, a.x), fmaf(s, b.y, a.y), fmaf(s, b.z, a.z), fmaf(s, b.w, a.w));
  }
}

__global__ void scale_rows_kernel(const float4* __restrict__ src, const float* __restrict__ scale, int rps, int64_t rows, int D4,
                                  float4* __restrict__ dst_f32, uint2* __restrict__ dst_bf16) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t n = rows * D4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float s = scale[(i / D4) / rps];
    const float4 a = src[i];
    const float4 v = make_float4(s * a.x, s * a.y, s * a.z, s * a.w);
    if (dst_f32) dst_f32[i] = v;
    if (dst_bf16) {
      uint2 o;
      o.x = pack_bf16x2(v.x, v.y);
      o.y = pack_bf16x2(v.z, v.w);
      dst_bf16[i] = o;
    }
  }
}

static inline int grid_1d(int64_t n) {
  int64_t b = (n + 255) / 256;
  if (b > 16 * kNumSMs) b = 16 * kNumSMs;
  return (int)(b < 1 ? 1 : b);
}

}  // namespace davf

using namespace davf;

extern "C" int davf_scale_rows_add(const float* res, const float* y, const float* scale, int rows_per_sample, int64_t rows, int D,
                                   float* out, davf_stream_t s) {
  DAVF_CHECK_ARG(res && y && scale && out && rows >= 0 && rows_per_sample > 0 && D > 0 && D % 4 == 0, "scale_rows_add: bad argument");
  DAVF_CHECK_ARG(rows % rows_per_sample == 0, "scale_rows_add: rows=%lld is not a multiple of rows_per_sample=%d", (long long)rows, rows_per_sample);
  if (rows == 0) return DAVF_OK;
  DAVF_CUDA(launch_pdl(scale_rows_add_kernel, dim3(grid_1d(rows * (D / 4))), dim3(256), 0, as_stream(s), reinterpret_cast<const float4*>(res), reinterpret_cast<const float4*>(y),
                                                                          scale, rows_per_sample, rows, D / 4, reinterpret_cast<float4*>(out)));
  DAVF_LAUNCH_OK();
  return DAVF_OK;
}

extern "C" int davf_scale_rows(const float* src, const float* scale, int rows_per_sample, int64_t rows, int D, float* dst_f32,
                               davf_bf16* dst_bf16, davf_stream_t s) {
  DAVF_CHECK_ARG(src && scale && (dst_f32 || dst_bf16) && rows >= 0 && rows_per_sample > 0 && D > 0 && D % 4 == 0, "scale_rows: bad argument");
  DAVF_CHECK_ARG(rows % rows_per_sample == 0, "scale_rows: rows=%lld is not a multiple of rows_per_sample=%d", (long long)rows, rows_per_sample);
  if (rows == 0) return DAVF_OK;
  DAVF_CUDA(launch_pdl(scale_rows_kernel, dim3(grid_1d(rows * (D / 4))), dim3(256), 0, as_stream(s), reinterpret_cast<const float4*>(src), scale, rows_per_sample, rows, D / 4,
                                                                      reinterpret_cast<float4*>(dst_f32), reinterpret_cast<uint2*>(dst_bf16)));
  DAVF_LAUNCH_OK();
  return DAVF_OK;
}
