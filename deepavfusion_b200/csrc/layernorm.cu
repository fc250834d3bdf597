// K4: LayerNorm forward / backward (replaces every nn.LayerNorm on the path and the torch.cat of
// deepavfusion.py:104-105).  Bandwidth-bound: one warp per row, the row lives in registers
// (D/128 float4 per lane), two-pass statistics via warp shuffles, 16-byte loads / 8-byte bf16
// stores.  Algorithmic bytes per row: forward D*(4 in + 2 out), backward D*(4 x + 2 dy + 4 dx).
#include "common.cuh"

namespace davf {

struct RowMap {
  int n0, n1, n;          // rows per sample from x0, x1, total
  int nseg;
  int seg_start[5];
  int B;
  __device__ __forceinline__ int64_t seg_row(int b, int r) const {
    if (nseg <= 1) return (int64_t)b * n + r;
    int k = 0;
#pragma unroll
    for (int i = 1; i < 4; ++i)
      if (i < nseg && r >= seg_start[i]) k = i;
    const int st = seg_start[k], len = seg_start[k + 1] - st;
    return (int64_t)B * st + (int64_t)b * len + (r - st);
  }
};

template <int VEC>
__global__ void __launch_bounds__(256) ln_fwd_kernel(davf_ln_fwd_args a, RowMap rm, int64_t rows) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const float inv_d = 1.0f / (float)a.D;
  float4 gam[VEC], bet[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    gam[i] = *reinterpret_cast<const float4*>(a.gamma + (i * 32 + lane) * 4);
    bet[i] = *reinterpret_cast<const float4*>(a.beta + (i * 32 + lane) * 4);
  }
  for (int64_t R = (int64_t)blockIdx.x * wpb + (threadIdx.x >> 5); R < rows; R += (int64_t)gridDim.x * wpb) {
    const int b = (int)(R / rm.n), r = (int)(R - (int64_t)b * rm.n);
    const float* src = (r < rm.n0) ? a.x0 + (int64_t)b * a.bs0 + (int64_t)r * a.D
                                   : a.x1 + (int64_t)b * a.bs1 + (int64_t)(r - rm.n0) * a.D;
    float4 x[VEC];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      x[i] = *reinterpret_cast<const float4*>(src + (i * 32 + lane) * 4);
      sum += (x[i].x + x[i].y) + (x[i].z + x[i].w);
    }
    const float mean = warp_sum(sum) * inv_d;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const float d0 = x[i].x - mean, d1 = x[i].y - mean, d2 = x[i].z - mean, d3 = x[i].w - mean;
      sq += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
    }
    const float rstd = rsqrtf(warp_sum(sq) * inv_d + a.eps);
    if (lane == 0) {
      if (a.mean) a.mean[R] = mean;
      if (a.rstd) a.rstd[R] = rstd;
    }
    const int64_t orow = rm.seg_row(b, r);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      float4 y;
      y.x = (x[i].x - mean) * rstd * gam[i].x + bet[i].x;
      y.y = (x[i].y - mean) * rstd * gam[i].y + bet[i].y;
      y.z = (x[i].z - mean) * rstd * gam[i].z + bet[i].z;
      y.w = (x[i].w - mean) * rstd * gam[i].w + bet[i].w;
      const int col = (i * 32 + lane) * 4;
      if (a.y_f32) *reinterpret_cast<float4*>(a.y_f32 + R * a.D + col) = y;
      if (a.y_bf16) {
        uint2 o;
        o.x = pack_bf16x2(y.x, y.y);
        o.y = pack_bf16x2(y.z, y.w);
        *reinterpret_cast<uint2*>(a.y_bf16 + orow * a.D + col) = o;
      }
    }
  }
}

// Two CTAs (16 warps) per SM: the kernel is a pure stream (14-18 bytes per element), so what matters is bytes in
// flight per SM.  gamma is re-read through L1 per row instead of living in 4 VEC registers.
template <int VEC>
__global__ void __launch_bounds__(256, 2) ln_bwd_kernel(davf_ln_bwd_args a, RowMap rm, int64_t rows) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ __align__(16) float sm_red[];   // [warps][2*D] : per-warp dgamma | dbeta partials
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const float inv_d = 1.0f / (float)a.D;
  const float4* gam = reinterpret_cast<const float4*>(a.gamma) + lane;     // gam[i * 32] = gamma of this lane's i-th float4
  // dgamma / dbeta partials of this warp live in its private slice of shared memory (every lane only ever touches
  // its own columns: no atomics, no syncs) instead of 8 VEC registers
  float4* my_dg = reinterpret_cast<float4*>(sm_red + (size_t)(threadIdx.x >> 5) * 2 * a.D) + lane;
  float4* my_db = my_dg + a.D / 4;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    my_dg[i * 32] = make_float4(0.f, 0.f, 0.f, 0.f);
    my_db[i * 32] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int64_t R = (int64_t)blockIdx.x * wpb + (threadIdx.x >> 5); R < rows; R += (int64_t)gridDim.x * wpb) {
    const int b = (int)(R / rm.n), r = (int)(R - (int64_t)b * rm.n);
    const bool first = r < rm.n0;
    const float* src = first ? a.x0 + (int64_t)b * a.bs0 + (int64_t)r * a.D
                             : a.x1 + (int64_t)b * a.bs1 + (int64_t)(r - rm.n0) * a.D;
    const float mean = a.mean[R], rstd = a.rstd[R];
    const int64_t srow = rm.seg_row(b, r);
    float4 xh[VEC], dy[VEC];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const int col = (i * 32 + lane) * 4;
      const float4 x = *reinterpret_cast<const float4*>(src + col);
      xh[i] = make_float4((x.x - mean) * rstd, (x.y - mean) * rstd, (x.z - mean) * rstd, (x.w - mean) * rstd);
      float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
      if (a.dy_bf16) {
        const uint2 u = *reinterpret_cast<const uint2*>(a.dy_bf16 + srow * a.D + col);
        const float2 lo = unpack_bf16x2(u.x), hi = unpack_bf16x2(u.y);
        d = make_float4(lo.x, lo.y, hi.x, hi.y);
      }
      if (a.dy_f32) {
        const float4 f = *reinterpret_cast<const float4*>(a.dy_f32 + R * a.D + col);
        d.x += f.x; d.y += f.y; d.z += f.z; d.w += f.w;
      }
      dy[i] = d;
      float4 sb = my_db[i * 32], sg = my_dg[i * 32];
      sb.x += d.x; sb.y += d.y; sb.z += d.z; sb.w += d.w;
      sg.x += d.x * xh[i].x; sg.y += d.y * xh[i].y; sg.z += d.z * xh[i].z; sg.w += d.w * xh[i].w;
      my_db[i * 32] = sb; my_dg[i * 32] = sg;
      const float4 gm = __ldg(gam + i * 32);
      const float g0 = d.x * gm.x, g1 = d.y * gm.y, g2 = d.z * gm.z, g3 = d.w * gm.w;
      s1 += (g0 + g1) + (g2 + g3);
      s2 += (g0 * xh[i].x + g1 * xh[i].y) + (g2 * xh[i].z + g3 * xh[i].w);
    }
    const float c1 = warp_sum(s1) * inv_d, c2 = warp_sum(s2) * inv_d;
    float* dst = first ? (a.dx0 ? a.dx0 + (int64_t)b * a.dbs0 + (int64_t)r * a.D : nullptr)
                       : (a.dx1 ? a.dx1 + (int64_t)b * a.dbs1 + (int64_t)(r - rm.n0) * a.D : nullptr);
    const float* add = first ? (a.add0 ? a.add0 + (int64_t)b * a.dbs0 + (int64_t)r * a.D : nullptr)
                             : (a.add1 ? a.add1 + (int64_t)b * a.dbs1 + (int64_t)(r - rm.n0) * a.D : nullptr);
    uint16_t* dst_lp = first ? (a.dx0_bf16 ? a.dx0_bf16 + ((int64_t)b * rm.n0 + r) * a.D : nullptr)
                             : (a.dx1_bf16 ? a.dx1_bf16 + ((int64_t)b * rm.n1 + (r - rm.n0)) * a.D : nullptr);
    if (dst) {
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        const int col = (i * 32 + lane) * 4;
        const float4 gm = __ldg(gam + i * 32);
        float4 o;
        o.x = rstd * (dy[i].x * gm.x - c1 - xh[i].x * c2);
        o.y = rstd * (dy[i].y * gm.y - c1 - xh[i].y * c2);
        o.z = rstd * (dy[i].z * gm.z - c1 - xh[i].z * c2);
        o.w = rstd * (dy[i].w * gm.w - c1 - xh[i].w * c2);
        if (add) {
          const float4 r4 = *reinterpret_cast<const float4*>(add + col);
          o.x += r4.x; o.y += r4.y; o.z += r4.z; o.w += r4.w;
        }
        *reinterpret_cast<float4*>(dst + col) = o;
        if (dst_lp) {
          uint2 u;
          u.x = pack_bf16x2(o.x, o.y);
          u.y = pack_bf16x2(o.z, o.w);
          *reinterpret_cast<uint2*>(dst_lp + col) = u;
        }
      }
    }
  }
  // CTA-level reduction of dgamma / dbeta: every warp stores its partials (plain stores, no shared atomics),
  // 4 columns per thread are summed over the warps, then ONE vector f32 reduction per 4 columns goes to HBM.
  __syncthreads();
  const int ncol4 = 2 * a.D / 4;
  for (int i = threadIdx.x; i < ncol4; i += blockDim.x) {
    // rotate the column order per CTA: CTAs finish together, and same-address reductions serialise in L2
    const int c4 = (i + (int)blockIdx.x * 61) % ncol4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int w = 0; w < wpb; ++w) {
      const float4 v = *reinterpret_cast<const float4*>(sm_red + (size_t)w * 2 * a.D + c4 * 4);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    float* dst = c4 * 4 < a.D ? (a.dgamma ? a.dgamma + c4 * 4 : nullptr) : (a.dbeta ? a.dbeta + (c4 * 4 - a.D) : nullptr);
    if (dst) asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(acc.x), "f"(acc.y), "f"(acc.z), "f"(acc.w) : "memory");
  }
}

static int make_rowmap(RowMap& rm, int n0, int n1, int B, int nseg, const int* seg_start) {
  rm.n0 = n0; rm.n1 = n1; rm.n = n0 + n1; rm.B = B; rm.nseg = nseg;
  for (int i = 0; i < 5; ++i) rm.seg_start[i] = 0;
  if (nseg > 1) {
    if (nseg > 4) return -1;
    for (int i = 0; i <= nseg; ++i) rm.seg_start[i] = seg_start[i];
    if (rm.seg_start[0] != 0 || rm.seg_start[nseg] != rm.n) return -1;
    for (int i = 0; i < nseg; ++i)
      if (rm.seg_start[i + 1] <= rm.seg_start[i]) return -1;
  }
  return 0;
}

}  // namespace davf

using namespace davf;

extern "C" int davf_layernorm_fwd(const davf_ln_fwd_args* a, davf_stream_t s) {
  DAVF_CHECK_ARG(a && a->x0 && a->gamma && a->beta, "layernorm_fwd: null argument");
  DAVF_CHECK_ARG(a->D % 128 == 0 && a->D >= 128 && a->D <= 1024, "layernorm_fwd: D=%d must be a multiple of 128 <= 1024", a->D);
  DAVF_CHECK_ARG(a->n0 > 0 && a->n1 >= 0 && (a->n1 == 0 || a->x1), "layernorm_fwd: bad row counts");
  RowMap rm;
  DAVF_CHECK_ARG(make_rowmap(rm, a->n0, a->n1, a->B, a->nseg, a->seg_start) == 0, "layernorm_fwd: bad segments");
  const int64_t rows = (int64_t)a->B * rm.n;
  if (rows == 0) return DAVF_OK;
  const int wpb = 8;
  int64_t blocks = (rows + wpb - 1) / wpb;
  if (blocks > 4 * kNumSMs) blocks = 4 * kNumSMs;
  cudaStream_t st = as_stream(s);
  switch (a->D / 128) {
    case 4: DAVF_CUDA(launch_pdl(ln_fwd_kernel<4>, dim3((int)blocks), dim3(256), 0, st, *a, rm, rows)); break;
    case 6: DAVF_CUDA(launch_pdl(ln_fwd_kernel<6>, dim3((int)blocks), dim3(256), 0, st, *a, rm, rows)); break;
    case 8: DAVF_CUDA(launch_pdl(ln_fwd_kernel<8>, dim3((int)blocks), dim3(256), 0, st, *a, rm, rows)); break;
    case 1: DAVF_CUDA(launch_pdl(ln_fwd_kernel<1>, dim3((int)blocks), dim3(256), 0, st, *a, rm, rows)); break;
    case 2: DAVF_CUDA(launch_pdl(ln_fwd_kernel<2>, dim3((int)blocks), dim3(256), 0, st, *a, rm, rows)); break;
    default:
      set_error("layernorm_fwd: D=%d not instantiated", a->D);
      return DAVF_EUNSUPPORTED;
  }
  DAVF_LAUNCH_OK();
  return DAVF_OK;
}

extern "C" int davf_layernorm_bwd(const davf_ln_bwd_args* a, davf_stream_t s) {
  DAVF_CHECK_ARG(a && a->x0 && a->gamma && a->mean && a->rstd, "layernorm_bwd: null argument");
  DAVF_CHECK_ARG(a->dy_bf16 || a->dy_f32, "layernorm_bwd: no incoming gradient");
  DAVF_CHECK_ARG(a->D % 128 == 0 && a->D >= 128 && a->D <= 1024, "layernorm_bwd: D=%d must be a multiple of 128 <= 1024", a->D);
  DAVF_CHECK_ARG(a->n0 > 0 && a->n1 >= 0 && (a->n1 == 0 || a->x1), "layernorm_bwd: bad row counts");
  RowMap rm;
  DAVF_CHECK_ARG(make_rowmap(rm, a->n0, a->n1, a->B, a->nseg, a->seg_start) == 0, "layernorm_bwd: bad segments");
  const int64_t rows = (int64_t)a->B * rm.n;
  if (rows == 0) return DAVF_OK;
  const int wpb = 8;
  int64_t blocks = (rows + 2 * wpb - 1) / (2 * wpb);          // >= 2 rows per warp so the reduction tail amortises
  if (blocks > 2 * kNumSMs) blocks = 2 * kNumSMs;
  if (blocks < 1) blocks = 1;
  const size_t smem = (size_t)wpb * 2 * a->D * sizeof(float);      // 48 KB at D = 768
  cudaStream_t st = as_stream(s);
  static bool attr_set = false;
  if (!attr_set) {
    DAVF_CUDA(cudaFuncSetAttribute(ln_bwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    DAVF_CUDA(cudaFuncSetAttribute(ln_bwd_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    DAVF_CUDA(cudaFuncSetAttribute(ln_bwd_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    // two 48 KB CTAs per SM: ask for the large shared-memory carve-out explicitly
    DAVF_CUDA(cudaFuncSetAttribute(ln_bwd_kernel<4>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
    DAVF_CUDA(cudaFuncSetAttribute(ln_bwd_kernel<6>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
    DAVF_CUDA(cudaFuncSetAttribute(ln_bwd_kernel<8>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
    attr_set = true;
  }
  switch (a->D / 128) {
    case 4: DAVF_CUDA(launch_pdl(ln_bwd_kernel<4>, dim3((int)blocks), dim3(256), smem, st, *a, rm, rows)); break;
    case 6: DAVF_CUDA(launch_pdl(ln_bwd_kernel<6>, dim3((int)blocks), dim3(256), smem, st, *a, rm, rows)); break;
    case 8: DAVF_CUDA(launch_pdl(ln_bwd_kernel<8>, dim3((int)blocks), dim3(256), smem, st, *a, rm, rows)); break;
    case 1: DAVF_CUDA(launch_pdl(ln_bwd_kernel<1>, dim3((int)blocks), dim3(256), smem, st, *a, rm, rows)); break;
    case 2: DAVF_CUDA(launch_pdl(ln_bwd_kernel<2>, dim3((int)blocks), dim3(256), smem, st, *a, rm, rows)); break;
    default:
      set_error("layernorm_bwd: D=%d not instantiated", a->D);
      return DAVF_EUNSUPPORTED;
  }
  DAVF_LAUNCH_OK();
  return DAVF_OK;
}
