// a11: the classifier tail of AVClassifier (reference models/classifier.py:42-59): token mean-pool,
// BatchNorm1d(affine=False) on the pooled features, and the three small Linear heads, forward and backward.
// Everything here is tiny (B <= 256 rows, D = 768, C = 310 / 527 classes, any C: no multiple-of-8 rule) and
// bandwidth / latency bound: plain coalesced SIMT kernels in f32 (the reference runs this path with
// use_amp = False, configs/linprobe.yaml:35).  The encoder in front of it is the tensor-core path.
#include "common.cuh"

namespace davf {

// pooled[b, d] = mean_t x[b, t, d]          grid (ceil(D / 128), B), 128 threads, coalesced over d
__global__ void meanpool_fwd_kernel(const float* __restrict__ x, int64_t bs, int n, int D, float* __restrict__ out) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (d >= D) return;
  const float* p = x + (int64_t)b * bs + d;
  float acc = 0.f;
  for (int t = 0; t < n; ++t) acc += p[(int64_t)t * D];
  out[(int64_t)b * D + d] = acc / (float)n;
}

// dx[b, t, d] = dy[b, d] / n
__global__ void meanpool_bwd_kernel(const float* __restrict__ dy, int n, int D, float* __restrict__ dx) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (d >= D) return;
  const float g = dy[(int64_t)b * D + d] / (float)n;
  float* p = dx + (int64_t)b * n * D + d;
  for (int t = 0; t < n; ++t) p[(int64_t)t * D] = g;
}

// BatchNorm1d(affine=False) over the batch; one thread per feature (B <= a few hundred rows).
// training: batch statistics (biased variance for the output, unbiased for the running estimate), running
// statistics updated with `momentum`; eval: running statistics.  mean / rstd used are saved for backward.
__global__ void bn1d_fwd_kernel(const float* __restrict__ x, int B, int D, int training, float* run_mean, float* run_var,
                                float momentum, float eps, float* __restrict__ y, float* __restrict__ save_mean,
                                float* __restrict__ save_rstd) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= D) return;
  float mean, var;
  if (training) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += x[(int64_t)b * D + d];
    mean = s / (float)B;
    float q = 0.f;
    for (int b = 0; b < B; ++b) { const float c = x[(int64_t)b * D + d] - mean; q += c * c; }
    var = q / (float)B;
    run_mean[d] = (1.f - momentum) * run_mean[d] + momentum * mean;
    run_var[d] = (1.f - momentum) * run_var[d] + momentum * (B > 1 ? q / (float)(B - 1) : var);
  } else {
    mean = run_mean[d];
    var = run_var[d];
  }
  const float rstd = rsqrtf(var + eps);
  for (int b = 0; b < B; ++b) y[(int64_t)b * D + d] = (x[(int64_t)b * D + d] - mean) * rstd;
  save_mean[d] = mean;
  save_rstd[d] = rstd;
}

// training: dx = rstd / B * (B dy - sum dy - xh sum(dy xh));  eval: dx = dy rstd
__global__ void bn1d_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ mean,
                                const float* __restrict__ rstd, int B, int D, int training, float* __restrict__ dx) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= D) return;
  const float mu = mean[d], rs = rstd[d];
  float s1 = 0.f, s2 = 0.f;
  if (training) {
    for (int b = 0; b < B; ++b) {
      const float g = dy[(int64_t)b * D + d], xh = (x[(int64_t)b * D + d] - mu) * rs;
      s1 += g; s2 += g * xh;
    }
  }
  for (int b = 0; b < B; ++b) {
    const float g = dy[(int64_t)b * D + d], xh = (x[(int64_t)b * D + d] - mu) * rs;
    dx[(int64_t)b * D + d] = training ? rs * (g - s1 / (float)B - xh * s2 / (float)B) : g * rs;
  }
}

// y[b, c] = sum_k x[b, k] W[c, k] + bias[c]        one warp per (b, c)
__global__ void head_fwd_kernel(const float* __restrict__ x, const float* __restrict__ W, const float* __restrict__ bias,
                                int B, int C, int D, float* __restrict__ y) {
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= (int64_t)B * C) return;
  const int b = (int)(w / C), c = (int)(w - (int64_t)b * C);
  const float* xr = x + (int64_t)b * D;
  const float* wr = W + (int64_t)c * D;
  float acc = 0.f;
  for (int k = lane; k < D; k += 32) acc = fmaf(xr[k], wr[k], acc);
  acc = warp_sum(acc);
  if (lane == 0) y[w] = acc + (bias ? bias[c] : 0.f);
}

// dW[c, k] += sum_b dy[b, c] x[b, k];  db[c] += sum_b dy[b, c]        grid (ceil(D / 128), C)
__global__ void head_bwd_w_kernel(const float* __restrict__ dy, const float* __restrict__ x, int B, int C, int D,
                                  float* __restrict__ dW, float* __restrict__ db) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x, c = blockIdx.y;
  float acc = 0.f, accb = 0.f;
  for (int b = 0; b < B; ++b) {
    const float g = dy[(int64_t)b * C + c];
    accb += g;
    if (k < D) acc = fmaf(g, x[(int64_t)b * D + k], acc);
  }
  if (k < D) dW[(int64_t)c * D + k] += acc;
  if (db && k == 0) db[c] += accb;
}

// dx[b, k] = sum_c dy[b, c] W[c, k]        grid (ceil(D / 128), B)
__global__ void head_bwd_x_kernel(const float* __restrict__ dy, const float* __restrict__ W, int B, int C, int D,
                                  float* __restrict__ dx) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (k >= D) return;
  float acc = 0.f;
  for (int c = 0; c < C; ++c) acc = fmaf(dy[(int64_t)b * C + c], W[(int64_t)c * D + k], acc);
  dx[(int64_t)b * D + k] = acc;
}

}  // namespace davf

using namespace davf;

extern "C" int davf_meanpool_fwd(const float* x, int64_t batch_stride, int B, int n, int D, float* out, davf_stream_t s) {
  DAVF_CHECK_ARG(x && out && B >= 0 && n > 0 && D > 0, "meanpool_fwd: bad argument");
  if (B == 0) return DAVF_OK;
  meanpool_fwd_kernel<<<dim3((D + 127) / 128, B), 128, 0, as_stream(s)>>>(x, batch_stride, n, D, out);
  DAVF_LAUNCH_OK();
  return DAVF_OK;
}

extern "C" int davf_meanpool_bwd(const float* dy, int B, int n, int D, float* dx, davf_stream_t s) {
  DAVF_CHECK_ARG(dy && dx && B >= 0 && n > 0 && D > 0, "meanpool_bwd: bad argument");
  if (B == 0) return DAVF_OK;
  meanpool_bwd_kernel<<<dim3((D + 127) / 128, B), 128, 0, as_stream(s)>>>(dy, n, D, dx);
  DAVF_LAUNCH_OK();
  return DAVF_OK;
}

extern "C" int davf_batchnorm1d_fwd(const float* x, int B, int D, int training, float* running_mean, float* running_var,
                                    float momentum, float eps, float* y, float* save_mean, float* save_rstd, davf_stream_t s) {
  DAVF_CHECK_ARG(x && y && running_mean && running_var && save_mean && save_rstd && B > 0 && D > 0, "batchnorm1d_fwd: bad argument");
  bn1d_fwd_kernel<<<(D + 127) / 128, 128, 0, as_stream(s)>>>(x, B, D, training, running_mean, running_var, momentum, eps, y, save_mean, save_rstd);
  DAVF_LAUNCH_OK();
  return DAVF_OK;
}

extern "C" int davf_batchnorm1d_bwd(const float* dy, const float* x, const float* mean, const float* rstd, int B, int D, int training,
                                    float* dx, davf_stream_t s) {
  DAVF_CHECK_ARG(dy && x && mean && rstd && dx && B > 0 && D > 0, "batchnorm1d_bwd: bad argument");
  bn1d_bwd_kernel<<<(D + 127) / 128, 128, 0, as_stream(s)>>>(dy, x, mean, rstd, B, D, training, dx);
  DAVF_LAUNCH_OK();
  return DAVF_OK;
}

extern "C" int davf_head_fwd(const float* x, const float* W, const float* bias, int B, int C, int D, float* y, davf_stream_t s) {
  DAVF_CHECK_ARG(x && W && y && B >= 0 && C > 0 && D > 0, "head_fwd: bad argument");
  if (B == 0) return DAVF_OK;
  const int64_t warps = (int64_t)B * C;
  head_fwd_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, as_stream(s)>>>(x, W, bias, B, C, D, y);
  DAVF_LAUNCH_OK();
  return DAVF_OK;
}

extern "C" int davf_head_bwd(const float* dy, const float* x, const float* W, int B, int C, int D, float* dW, float* db, float* dx,
                             davf_stream_t s) {
  DAVF_CHECK_ARG(dy && x && W && B > 0 && C > 0 && D > 0, "head_bwd: bad argument");
  if (dW) {
    head_bwd_w_kernel<<<dim3((D + 127) / 128, C), 128, 0, as_stream(s)>>>(dy, x, B, C, D, dW, db);
    DAVF_LAUNCH_OK();
  }
  if (dx) {
    head_bwd_x_kernel<<<dim3((D + 127) / 128, B), 128, 0, as_stream(s)>>>(dy, W, B, C, D, dx);
    DAVF_LAUNCH_OK();
  }
  return DAVF_OK;
}
