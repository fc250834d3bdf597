// K12: normalised masked patch-MSE (replaces AVMAE.patchify avmae.py:201-214 + forward_loss
// :183-198).  One warp per masked patch: the patch is read straight from the NCHW input (no
// patchify copy), normalised in registers (mean / unbiased variance + 1e-6), compared with the
// prediction row; kept patches are skipped (they contribute 0 to the loss and a zero gradient).
// Algorithmic bytes per masked patch: P*4 (input) + P*4 (pred) [+ P*2 dpred in backward].
#include "common.cuh"

namespace davf {

constexpr int kMaxPerLane = 32;   // P <= 1024

struct PatchGeom { int B, C, H, W, p, gW, L, P; };

// loads the patch of (b,l) into registers in MEMORY order (c, py, px) and returns per-lane
// normalised targets together with the index e = (py*p + px)*C + c of each in the pred row.
// PS = log2(patch size) when it is a compile-time power of two (16 x 16 patches everywhere on the path: the index
// arithmetic is shifts; with run-time divisors the kernel was bound by ~150 integer divisions per lane and patch),
// 0 = generic; NT = P / 32 elements per lane when known at compile time (24: 16 x 16 x 3 frames, 8: spectrogram patches; 0 =
// generic, up to 32 with bounds checks).  ALL loads of a patch -- pixels and prediction row -- are issued unconditionally
// before the first use (ncu: with the loads inside `if (idx < P)` blocks next to their accumulation every load's latency was
// paid in turn, 92 us for 58 MB), and the backward's
// dpred row is transposed through shared memory into 16-byte stores (it was 768 two-byte stores at a 6-byte stride).
template <bool BWD, int PS, int NT>
__global__ void __launch_bounds__(256) masked_mse_kernel(const float* __restrict__ img, const float* __restrict__ pred,
                                                         const float* __restrict__ mask, float* __restrict__ loss_sum,
                                                         const float* __restrict__ gscale, float inv_count,
                                                         uint16_t* __restrict__ dpred, PatchGeom g, int pred_G, int pred_off,
                                                         int norm_pix) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[8];
  __shared__ __align__(16) uint16_t stg[BWD ? 8 : 1][BWD ? 32 * kMaxPerLane : 8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t patch = (int64_t)blockIdx.x * 8 + warp;
  const int64_t npatch = (int64_t)g.B * g.L;
  const bool vec_rows = (g.P & 7) == 0;                 // dpred rows are 16-byte multiples
  float lsum = 0.f;
  if (patch < npatch) {
    const int b = (int)(patch / g.L), l = (int)(patch - (int64_t)b * g.L);
    const bool masked = mask[patch] != 0.f;
    if (!masked) {
      if (BWD) {
        uint16_t* dr = dpred + patch * g.P;
        if (vec_rows) { for (int e = lane * 8; e < g.P; e += 256) *reinterpret_cast<uint4*>(dr + e) = make_uint4(0u, 0u, 0u, 0u); }
        else { for (int e = lane * 2; e < g.P; e += 64) *reinterpret_cast<uint32_t*>(dr + e) = 0u; }
      }
    } else {
      const int gy = l / g.gW, gx = l - gy * g.gW;
      const float* prow = pred + ((int64_t)b * pred_G + pred_off + l) * g.P;
      const int pp = g.p * g.p;
      auto split = [&](int idx, int& c, int& r, int& py, int& px) {
        if (PS > 0) { c = idx >> (2 * PS); r = idx & ((1 << (2 * PS)) - 1); py = r >> PS; px = r & ((1 << PS) - 1); }
        else { c = idx / pp; r = idx - c * pp; py = r / g.p; px = r - py * g.p; }
      };
      constexpr int TN = NT > 0 ? NT : kMaxPerLane;
      float x[TN], pr[TN];
#pragma unroll
      for (int t = 0; t < TN; ++t) {                     // every load in flight before anything is consumed
        const int idx = lane + 32 * t;
        const bool ok = NT > 0 || idx < g.P;
        int c, r, py, px;
        split(ok ? idx : 0, c, r, py, px);
        x[t] = __ldg(img + (((int64_t)b * g.C + c) * g.H + gy * g.p + py) * g.W + gx * g.p + px);
        pr[t] = __ldg(prow + r * g.C + c);               // (py, px, c) order of patchify
      }
      float s = 0.f;
#pragma unroll
      for (int t = 0; t < TN; ++t) {
        if (NT == 0 && lane + 32 * t >= g.P) { x[t] = 0.f; pr[t] = 0.f; }
        s += x[t];
      }
      float mean = 0.f, rs = 1.f;
      if (norm_pix) {
        mean = warp_sum(s) / (float)g.P;
        float sq = 0.f;
#pragma unroll
        for (int t = 0; t < TN; ++t)
          if (NT > 0 || lane + 32 * t < g.P) { const float d = x[t] - mean; sq += d * d; }
        const float var = warp_sum(sq) / (float)(g.P - 1);          // unbiased, avmae.py:191
        rs = rsqrtf(var + 1.0e-6f);
      }
      const float gs = BWD ? gscale[0] * inv_count * 2.0f / (float)g.P : 0.f;
#pragma unroll
      for (int t = 0; t < TN; ++t) {
        const int idx = lane + 32 * t;
        if (NT > 0 || idx < g.P) {
          int c, r, py, px;
          split(idx, c, r, py, px);
          const float diff = pr[t] - (x[t] - mean) * rs;
          if (BWD) stg[warp][r * g.C + c] = f32_to_bf16(gs * diff);
          else lsum += diff * diff;
        }
      }
      if (BWD) {
        __syncwarp();
        uint16_t* dr = dpred + patch * g.P;
        if (vec_rows) { for (int e = lane * 8; e < g.P; e += 256) *reinterpret_cast<uint4*>(dr + e) = *reinterpret_cast<const uint4*>(&stg[warp][e]); }
        else { for (int e = lane * 2; e < g.P; e += 64) *reinterpret_cast<uint32_t*>(dr + e) = *reinterpret_cast<const uint32_t*>(&stg[warp][e]); }
      }
      lsum = warp_sum(lsum) / (float)g.P;
    }
  }
  if (!BWD) {
    if (lane == 0) red[warp] = lsum;
    __syncthreads();
    if (threadIdx.x == 0) {
      float tot = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) tot += red[w];
      if (tot != 0.f) atomicAdd(loss_sum, tot);
    }
  }
}

static int geom(PatchGeom& g, int B, int C, int H, int W, int p) {
  if (p <= 0 || H % p || W % p) return -1;
  g.B = B; g.C = C; g.H = H; g.W = W; g.p = p; g.gW = W / p; g.L = (H / p) * (W / p); g.P = p * p * C;
  if (g.P > 32 * kMaxPerLane || g.P % 2) return -1;
  return 0;
}

}  // namespace davf

using namespace davf;

extern "C" int davf_masked_mse_fwd(const float* img, const float* pred, const float* mask, float* loss_sum, int B, int C, int H,
                                   int W, int p, int pred_G, int pred_off, int norm_pix, davf_stream_t s) {
  PatchGeom g;
  DAVF_CHECK_ARG(geom(g, B, C, H, W, p) == 0, "masked_mse_fwd: unsupported geometry C=%d H=%d W=%d p=%d", C, H, W, p);
  if (B == 0) return DAVF_OK;
  const int64_t np = (int64_t)B * g.L;
  const dim3 grid((int)((np + 7) / 8));
  if (p == 16 && g.P == 768) DAVF_CUDA(launch_pdl(masked_mse_kernel<false, 4, 24>, grid, dim3(256), 0, as_stream(s), img, pred, mask, loss_sum, nullptr, 0.f, nullptr, g, pred_G, pred_off, norm_pix));
  else if (p == 16 && g.P == 256) DAVF_CUDA(launch_pdl(masked_mse_kernel<false, 4, 8>, grid, dim3(256), 0, as_stream(s), img, pred, mask, loss_sum, nullptr, 0.f, nullptr, g, pred_G, pred_off, norm_pix));
  else DAVF_CUDA(launch_pdl(masked_mse_kernel<false, 0, 0>, grid, dim3(256), 0, as_stream(s), img, pred, mask, loss_sum, nullptr, 0.f, nullptr, g, pred_G, pred_off, norm_pix));
  DAVF_LAUNCH_OK();
  return DAVF_OK;
}

extern "C" int davf_masked_mse_bwd(const float* img, const float* pred, const float* mask, const float* gscale, float inv_count,
                                   davf_bf16* dpred, int B, int C, int H, int W, int p, int pred_G, int pred_off, int norm_pix,
                                   davf_stream_t s) {
  PatchGeom g;
  DAVF_CHECK_ARG(geom(g, B, C, H, W, p) == 0, "masked_mse_bwd: unsupported geometry C=%d H=%d W=%d p=%d", C, H, W, p);
  DAVF_CHECK_ARG(gscale && dpred, "masked_mse_bwd: null pointer");
  if (B == 0) return DAVF_OK;
  const int64_t np = (int64_t)B * g.L;
  const dim3 grid((int)((np + 7) / 8));
  if (p == 16 && g.P == 768) DAVF_CUDA(launch_pdl(masked_mse_kernel<true, 4, 24>, grid, dim3(256), 0, as_stream(s), img, pred, mask, nullptr, gscale, inv_count, dpred, g, pred_G, pred_off, norm_pix));
  else if (p == 16 && g.P == 256) DAVF_CUDA(launch_pdl(masked_mse_kernel<true, 4, 8>, grid, dim3(256), 0, as_stream(s), img, pred, mask, nullptr, gscale, inv_count, dpred, g, pred_G, pred_off, norm_pix));
  else DAVF_CUDA(launch_pdl(masked_mse_kernel<true, 0, 0>, grid, dim3(256), 0, as_stream(s), img, pred, mask, nullptr, gscale, inv_count, dpred, g, pred_G, pred_off, norm_pix));
  DAVF_LAUNCH_OK();
  return DAVF_OK;
}
