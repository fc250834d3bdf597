// K4: LayerNorm forward / backward (replaces every nn.LayerNorm on the path and the torch.cat of
// deepavfusion.py:104-105).  Bandwidth-bound: one warp per row, the row lives in registers
// (D/128 float4 per lane), two-pass statistics via warp shuffles, 16-byte loads / 8-byte bf16
// stores.  Algorithmic bytes per row: forward D*(4 in + 2 out), backward D*(4 x + 2 dy + 4 dx).
#include <stdlib.h>
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

// ---------------------------------------------------------------------------------------------
// backward, bulk-copy ring (the default): same arithmetic as ln_bwd_kernel, but every warp keeps the rows it is going to
// process NEXT in flight.  The register version above has one row per warp in flight and two dependent latency phases
// per row (x / dy, then the residual gradient), i.e. ~35 KB per SM against the ~50 KB that 6.4 TB/s times the loaded
// latency needs: ncu showed 12-24 % warps active, 16-27 % issue active, and the 133 launches of a step ran at 38 % of the
// copy bandwidth.  Here lane 0 of each warp fetches whole rows (x, dy, residual gradient) with cp.async.bulk into a
// per-warp ring of shared-memory slots that complete on an mbarrier, `slots` rows ahead: 8 warps x 2-3 slots x 7.5 KB
// = 120-180 KB in flight per SM, no registers involved.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ln_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ln_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void ln_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  }
}

template <int VEC>
__global__ void __launch_bounds__(256, 1) ln_bwd_ring_kernel(davf_ln_bwd_args a, RowMap rm, int64_t rows, int slots) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int D = VEC * 128, WPB = 8;
  extern __shared__ __align__(16) float sm_red[];   // [WPB][2*D] dgamma | dbeta partials, then the row ring, then the barriers
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float inv_d = 1.0f / (float)D;
  // slot layout: x f32 [D] | residual gradient f32 [D] | dy f32 [D] (if any) | dy bf16 [D] (if any)
  const uint32_t off_add = D * 4, off_dyf = 2 * D * 4, off_dyb = off_dyf + (a.dy_f32 ? D * 4 : 0);
  const uint32_t slot_bytes = off_dyb + (a.dy_bf16 ? D * 2 : 0);
  const uint32_t ring_base = ln_smem_u32(sm_red) + WPB * 2 * D * 4 + (uint32_t)warp * (uint32_t)slots * slot_bytes;
  const uint32_t bar_base = ln_smem_u32(sm_red) + WPB * 2 * D * 4 + WPB * (uint32_t)slots * slot_bytes + (uint32_t)warp * 4u * 8u;
  const float4* gam = reinterpret_cast<const float4*>(a.gamma) + lane;
  float4* my_dg = reinterpret_cast<float4*>(sm_red + (size_t)warp * 2 * D) + lane;
  float4* my_db = my_dg + D / 4;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    my_dg[i * 32] = make_float4(0.f, 0.f, 0.f, 0.f);
    my_db[i * 32] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (lane == 0) {
    for (int s = 0; s < slots; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_base + 8u * s) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const int64_t R0 = (int64_t)blockIdx.x * WPB + warp, stride = (int64_t)gridDim.x * WPB;

  struct RowPtr { const float* x; const float* add; float* dst; uint16_t* dst_lp; int64_t srow; };
  auto locate = [&](int64_t R) {
    RowPtr q;
    const int b = (int)(R / rm.n), r = (int)(R - (int64_t)b * rm.n);
    const bool first = r < rm.n0;
    q.x = first ? a.x0 + (int64_t)b * a.bs0 + (int64_t)r * D : a.x1 + (int64_t)b * a.bs1 + (int64_t)(r - rm.n0) * D;
    q.add = first ? (a.add0 ? a.add0 + (int64_t)b * a.dbs0 + (int64_t)r * D : nullptr)
                  : (a.add1 ? a.add1 + (int64_t)b * a.dbs1 + (int64_t)(r - rm.n0) * D : nullptr);
    q.dst = first ? (a.dx0 ? a.dx0 + (int64_t)b * a.dbs0 + (int64_t)r * D : nullptr)
                  : (a.dx1 ? a.dx1 + (int64_t)b * a.dbs1 + (int64_t)(r - rm.n0) * D : nullptr);
    q.dst_lp = first ? (a.dx0_bf16 ? a.dx0_bf16 + ((int64_t)b * rm.n0 + r) * D : nullptr)
                     : (a.dx1_bf16 ? a.dx1_bf16 + ((int64_t)b * rm.n1 + (r - rm.n0)) * D : nullptr);
    q.srow = rm.seg_row(b, r);
    return q;
  };
  auto fetch = [&](int64_t R, int slot) {                 // lane 0: request one row into a slot
    const RowPtr q = locate(R);
    const uint32_t base = ring_base + (uint32_t)slot * slot_bytes, bar = bar_base + 8u * slot;
    const bool want_add = q.add != nullptr && q.dst != nullptr;
    const uint32_t bytes = D * 4 + (want_add ? D * 4 : 0) + (a.dy_f32 ? D * 4 : 0) + (a.dy_bf16 ? D * 2 : 0);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    ln_bulk_g2s(base, q.x, D * 4, bar);
    if (want_add) ln_bulk_g2s(base + off_add, q.add, D * 4, bar);
    if (a.dy_f32) ln_bulk_g2s(base + off_dyf, a.dy_f32 + R * D, D * 4, bar);
    if (a.dy_bf16) ln_bulk_g2s(base + off_dyb, a.dy_bf16 + q.srow * D, D * 2, bar);
  };
  if (lane == 0)
    for (int s = 0; s < slots; ++s)
      if (R0 + s * stride < rows) fetch(R0 + s * stride, s);

  int k = 0;
  for (int64_t R = R0; R < rows; R += stride, ++k) {
    const int slot = k % slots;
    const uint32_t base = ring_base + (uint32_t)slot * slot_bytes;
    const RowPtr q = locate(R);
    const float mean = a.mean[R], rstd = a.rstd[R];
    ln_mbar_wait(bar_base + 8u * slot, (uint32_t)(k / slots) & 1u);
    float4 xh[VEC], dy[VEC];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const uint32_t c16 = (uint32_t)(i * 32 + lane) * 16u;
      float4 x;
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w) : "r"(base + c16));
      xh[i] = make_float4((x.x - mean) * rstd, (x.y - mean) * rstd, (x.z - mean) * rstd, (x.w - mean) * rstd);
      float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
      if (a.dy_bf16) {
        uint2 u;
        asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(u.x), "=r"(u.y) : "r"(base + off_dyb + c16 / 2));
        const float2 lo = unpack_bf16x2(u.x), hi = unpack_bf16x2(u.y);
        d = make_float4(lo.x, lo.y, hi.x, hi.y);
      }
      if (a.dy_f32) {
        float4 f;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(f.x), "=f"(f.y), "=f"(f.z), "=f"(f.w) : "r"(base + off_dyf + c16));
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
    if (q.dst) {
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        const int col = (i * 32 + lane) * 4;
        const float4 gm = __ldg(gam + i * 32);
        float4 o;
        o.x = rstd * (dy[i].x * gm.x - c1 - xh[i].x * c2);
        o.y = rstd * (dy[i].y * gm.y - c1 - xh[i].y * c2);
        o.z = rstd * (dy[i].z * gm.z - c1 - xh[i].z * c2);
        o.w = rstd * (dy[i].w * gm.w - c1 - xh[i].w * c2);
        if (q.add) {
          float4 r4;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r4.x), "=f"(r4.y), "=f"(r4.z), "=f"(r4.w) : "r"(base + off_add + (uint32_t)col * 4u));
          o.x += r4.x; o.y += r4.y; o.z += r4.z; o.w += r4.w;
        }
        *reinterpret_cast<float4*>(q.dst + col) = o;
        if (q.dst_lp) {
          uint2 u;
          u.x = pack_bf16x2(o.x, o.y);
          u.y = pack_bf16x2(o.z, o.w);
          *reinterpret_cast<uint2*>(q.dst_lp + col) = u;
        }
      }
    }
    __syncwarp();                                          // every lane is done with the slot (its values went into the stores above)
    if (lane == 0 && R + (int64_t)slots * stride < rows) fetch(R + (int64_t)slots * stride, slot);
  }
  // CTA-level reduction of dgamma / dbeta (as in ln_bwd_kernel; one CTA per SM: half the global reductions)
  __syncthreads();
  const int ncol4 = 2 * D / 4;
  for (int i = threadIdx.x; i < ncol4; i += blockDim.x) {
    const int c4 = (i + (int)blockIdx.x * 61) % ncol4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int w = 0; w < WPB; ++w) {
      const float4 v = *reinterpret_cast<const float4*>(sm_red + (size_t)w * 2 * D + c4 * 4);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    float* dst = c4 * 4 < D ? (a.dgamma ? a.dgamma + c4 * 4 : nullptr) : (a.dbeta ? a.dbeta + (c4 * 4 - D) : nullptr);
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
  cudaStream_t st = as_stream(s);
  // bulk-copy ring (default; DAVF_LN_RING=0 selects the register kernel): rows, strides and pointers must be 16-byte aligned
  static const int ring_on = [] { const char* e = getenv("DAVF_LN_RING"); return e ? atoi(e) : 1; }();
  auto al16 = [](const void* p) { return ((uintptr_t)p & 15) == 0; };
  const bool ring_ok = ring_on && (a->D == 512 || a->D == 768 || a->D == 1024) && rows >= 2 * wpb && a->bs0 % 4 == 0 && a->bs1 % 4 == 0 && a->dbs0 % 4 == 0 &&
                       a->dbs1 % 4 == 0 && al16(a->x0) && al16(a->x1) && al16(a->add0) && al16(a->add1) && al16(a->dy_f32) && al16(a->dy_bf16);
  if (ring_ok) {
    const size_t slot_bytes = (size_t)a->D * 8 + (a->dy_f32 ? (size_t)a->D * 4 : 0) + (a->dy_bf16 ? (size_t)a->D * 2 : 0);
    const size_t fixed = (size_t)wpb * 2 * a->D * 4 + wpb * 4 * 8;
    int slots = (int)((200 * 1024 - fixed) / (wpb * slot_bytes));
    if (slots > 3) slots = 3;
    int64_t blocks = (rows + 2 * wpb - 1) / (2 * wpb);
    if (blocks > kNumSMs) blocks = kNumSMs;
    const int64_t rows_per_warp = (rows + blocks * wpb - 1) / (blocks * wpb);
    if (slots > rows_per_warp) slots = (int)rows_per_warp;
    if (slots >= 1) {
      const size_t smem = fixed + (size_t)wpb * slots * slot_bytes;
      static bool ring_attr = false;
      if (!ring_attr) {
        DAVF_CUDA(cudaFuncSetAttribute(ln_bwd_ring_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        DAVF_CUDA(cudaFuncSetAttribute(ln_bwd_ring_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        DAVF_CUDA(cudaFuncSetAttribute(ln_bwd_ring_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        ring_attr = true;
      }
      switch (a->D / 128) {
        case 4: DAVF_CUDA(launch_pdl(ln_bwd_ring_kernel<4>, dim3((int)blocks), dim3(256), smem, st, *a, rm, rows, slots)); break;
        case 6: DAVF_CUDA(launch_pdl(ln_bwd_ring_kernel<6>, dim3((int)blocks), dim3(256), smem, st, *a, rm, rows, slots)); break;
        default: DAVF_CUDA(launch_pdl(ln_bwd_ring_kernel<8>, dim3((int)blocks), dim3(256), smem, st, *a, rm, rows, slots)); break;
      }
      DAVF_LAUNCH_OK();
      return DAVF_OK;
    }
  }
  int64_t blocks = (rows + 2 * wpb - 1) / (2 * wpb);          // >= 2 rows per warp so the reduction tail amortises
  if (blocks > 2 * kNumSMs) blocks = 2 * kNumSMs;
  if (blocks < 1) blocks = 1;
  const size_t smem = (size_t)wpb * 2 * a->D * sizeof(float);      // 48 KB at D = 768
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
