// Bandwidth-bound row kernels: im2col of kept patches (K1/K3), f32->bf16 row casts, bias /
// batch reductions, decoder sequence assembly (K3).  All 16-byte vectorised, grid-stride,
// grids sized in multiples of the SM count.
#include "common.cuh"

namespace davf {

static inline int grid_for(int64_t work_items, int threads, int max_waves = 8) {
  int64_t blocks = (work_items + threads - 1) / threads;
  int64_t cap = (int64_t)kNumSMs * max_waves;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

// ---------------------------------------------------------------------------------------------
// patch rows: out[(b*nK + t), c*p*p + py*p + px] = img[b, c, gy*p + py, gx*p + px]
// ---------------------------------------------------------------------------------------------
__global__ void patch_rows_kernel(const float* __restrict__ img, const int64_t* __restrict__ ids_keep,
                                  uint16_t* __restrict__ out, int B, int C, int H, int W, int p, int nK) {
  pdl_launch_dependents();
  pdl_wait();
  const int gW = W / p;
  const int qpr = p / 4;                         // float4 quads per patch row
  const int64_t quads_per_row = (int64_t)C * p * qpr;
  const int64_t total = (int64_t)B * nK * quads_per_row;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = idx / quads_per_row;
    int rem = (int)(idx - row * quads_per_row);
    const int c = rem / (p * qpr);
    rem -= c * p * qpr;
    const int py = rem / qpr, q = rem - py * qpr;
    const int b = (int)(row / nK);
    const int l = ids_keep ? (int)ids_keep[row] : (int)(row - (int64_t)b * nK);
    const int gy = l / gW, gx = l - gy * gW;
    const float4 v = *reinterpret_cast<const float4*>(img + (((int64_t)b * C + c) * H + gy * p + py) * W + gx * p + q * 4);
    uint2 o;
    o.x = pack_bf16x2(v.x, v.y);
    o.y = pack_bf16x2(v.z, v.w);
    *reinterpret_cast<uint2*>(out + row * ((int64_t)C * p * p) + (c * p + py) * p + q * 4) = o;
  }
}

// ---------------------------------------------------------------------------------------------
// cast rows f32 -> bf16 with a per-sample row window on the source
// ---------------------------------------------------------------------------------------------
__global__ void cast_rows_kernel(const float* __restrict__ src, uint16_t* __restrict__ dst, int64_t M, int D,
                                 int g, int G, int off) {
  pdl_launch_dependents();
  pdl_wait();
  const int vpr = D / 4;
  const int64_t total = M * vpr;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = idx / vpr;
    const int c = (int)(idx - m * vpr);
    const int64_t srow = (m / g) * G + off + (m % g);
    const float4 v = *reinterpret_cast<const float4*>(src + srow * D + c * 4);
    uint2 o;
    o.x = pack_bf16x2(v.x, v.y);
    o.y = pack_bf16x2(v.z, v.w);
    *reinterpret_cast<uint2*>(dst + m * D + c * 4) = o;
  }
}

// out = a + b (+ c), f32, plus a bf16 copy: the gradient fan-in of a tensor with two or three consumers
// (deepavfusion.py:104-106: x_image / x_audio feed their block and the fusion block, x_fusion feeds all three)
__global__ void sum_cast_kernel(const float4* __restrict__ a, const float4* __restrict__ b, const float4* __restrict__ c,
                                float4* __restrict__ out, uint2* __restrict__ out_lp, int64_t n4) {
  pdl_launch_dependents();
  pdl_wait();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 v = a[i];
    const float4 w = b[i];
    v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
    if (c) { const float4 u = c[i]; v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w; }
    out[i] = v;
    if (out_lp) {
      uint2 o;
      o.x = pack_bf16x2(v.x, v.y);
      o.y = pack_bf16x2(v.z, v.w);
      out_lp[i] = o;
    }
  }
}

__global__ void cast_flat_kernel(const float* __restrict__ src, uint16_t* __restrict__ dst, int64_t n4) {
  pdl_launch_dependents();
  pdl_wait();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(src)[i];
    uint2 o;
    o.x = pack_bf16x2(v.x, v.y);
    o.y = pack_bf16x2(v.z, v.w);
    reinterpret_cast<uint2*>(dst)[i] = o;
  }
}

// ---------------------------------------------------------------------------------------------
// column sums of a bf16 matrix: out[n] += sum_m x[m, n]
// block = 32 column pairs x 8 row lanes; grid = (N/64, row slabs)
// ---------------------------------------------------------------------------------------------
__global__ void colsum_bf16_kernel(const uint16_t* __restrict__ x, int64_t M, int N, int64_t ld, float* __restrict__ out,
                                   int rows_per_block) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float2 red[8][32];
  const int cp = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int n = (blockIdx.x * 32 + cp) * 2;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  int64_t r1 = r0 + rows_per_block;
  if (r1 > M) r1 = M;
  float2 acc = make_float2(0.f, 0.f);
  if (n < N) {
    for (int64_t r = r0 + rl; r < r1; r += 8) {
      const float2 v = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(x + r * ld + n));
      acc.x += v.x;
      acc.y += v.y;
    }
  }
  red[rl][cp] = acc;
  __syncthreads();
  if (rl == 0 && n < N) {
#pragma unroll
    for (int k = 1; k < 8; ++k) {
      acc.x += red[k][cp].x;
      acc.y += red[k][cp].y;
    }
    atomicAdd(out + n, acc.x);
    atomicAdd(out + n + 1, acc.y);
  }
}

// out[r*D + d] (+)= sum_b x[(b*G + off + r)*D + d]
__global__ void batchsum_f32_kernel(const float* __restrict__ x, int B, int G, int off, int g, int D,
                                    float* __restrict__ out, int accumulate) {
  pdl_launch_dependents();
  pdl_wait();
  const int vpr = D / 4;
  const int64_t total = (int64_t)g * vpr;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(idx / vpr), c = (int)(idx - (int64_t)r * vpr);
    float4 acc = accumulate ? *reinterpret_cast<const float4*>(out + (int64_t)r * D + c * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int b = 0; b < B; ++b) {
      const float4 v = *reinterpret_cast<const float4*>(x + ((int64_t)b * G + off + r) * D + c * 4);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    *reinterpret_cast<float4*>(out + (int64_t)r * D + c * 4) = acc;
  }
}

// ---------------------------------------------------------------------------------------------
// decoder sequence assembly (avmae.py:161-169)
// ---------------------------------------------------------------------------------------------
__global__ void dec_assemble_fwd_kernel(const float* __restrict__ e, const float* __restrict__ ef,
                                        const float* __restrict__ mask_token, const float* __restrict__ pos,
                                        const int64_t* __restrict__ ids_restore, float* __restrict__ seq,
                                        int B, int nK, int nF, int L, int D) {
  pdl_launch_dependents();
  pdl_wait();
  const int vpr = D / 4;
  const int S = nF + L;
  const int64_t total = (int64_t)B * S * vpr;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = idx / vpr;
    const int c = (int)(idx - row * vpr);
    const int b = (int)(row / S), r = (int)(row - (int64_t)b * S);
    float4 v;
    if (r < nF) {
      v = *reinterpret_cast<const float4*>(ef + ((int64_t)b * nF + r) * D + c * 4);
    } else {
      const int l = r - nF;
      const int64_t src = ids_restore[(int64_t)b * L + l];
      v = (src < nK) ? *reinterpret_cast<const float4*>(e + ((int64_t)b * nK + src) * D + c * 4)
                     : *reinterpret_cast<const float4*>(mask_token + c * 4);
      const float4 pe = *reinterpret_cast<const float4*>(pos + (int64_t)l * D + c * 4);
      v.x += pe.x; v.y += pe.y; v.z += pe.z; v.w += pe.w;
    }
    *reinterpret_cast<float4*>(seq + row * D + c * 4) = v;
  }
}

// rows of de (gather through ids_keep) and def (fusion rows), f32 -> bf16
__global__ void dec_assemble_bwd_rows_kernel(const float* __restrict__ dseq, const int64_t* __restrict__ ids_keep,
                                             uint16_t* __restrict__ de, uint16_t* __restrict__ def_,
                                             int B, int nK, int nF, int L, int D) {
  pdl_launch_dependents();
  pdl_wait();
  const int vpr = D / 4;
  const int S = nF + L;
  const int R = nK + nF;
  const int64_t total = (int64_t)B * R * vpr;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = idx / vpr;
    const int c = (int)(idx - row * vpr);
    const int b = (int)(row / R), r = (int)(row - (int64_t)b * R);
    int64_t srow;
    uint16_t* dst;
    if (r < nK) {
      srow = (int64_t)b * S + nF + ids_keep[(int64_t)b * nK + r];
      dst = de + ((int64_t)b * nK + r) * D;
    } else {
      srow = (int64_t)b * S + (r - nK);
      dst = def_ + ((int64_t)b * nF + (r - nK)) * D;
    }
    const float4 v = *reinterpret_cast<const float4*>(dseq + srow * D + c * 4);
    uint2 o;
    o.x = pack_bf16x2(v.x, v.y);
    o.y = pack_bf16x2(v.z, v.w);
    *reinterpret_cast<uint2*>(dst + c * 4) = o;
  }
}

// dpos[l] += sum_b dseq[b, nF+l]; dmask_token += sum over masked (b,l)
__global__ void dec_assemble_bwd_reduce_kernel(const float* __restrict__ dseq, const int64_t* __restrict__ ids_restore,
                                               float* __restrict__ dmask_token, float* __restrict__ dpos,
                                               int B, int nK, int nF, int L, int D) {
  pdl_launch_dependents();
  pdl_wait();
  const int l = blockIdx.x;
  const int S = nF + L;
  for (int c = threadIdx.x; c < D / 4; c += blockDim.x) {
    float4 ap = make_float4(0.f, 0.f, 0.f, 0.f), am = ap;
    for (int b = 0; b < B; ++b) {
      const float4 v = *reinterpret_cast<const float4*>(dseq + ((int64_t)b * S + nF + l) * D + c * 4);
      ap.x += v.x; ap.y += v.y; ap.z += v.z; ap.w += v.w;
      if (ids_restore[(int64_t)b * L + l] >= nK) { am.x += v.x; am.y += v.y; am.z += v.z; am.w += v.w; }
    }
    float4* dp = reinterpret_cast<float4*>(dpos + (int64_t)l * D + c * 4);
    float4 o = *dp;
    o.x += ap.x; o.y += ap.y; o.z += ap.z; o.w += ap.w;
    *dp = o;
    atomicAdd(dmask_token + c * 4 + 0, am.x);
    atomicAdd(dmask_token + c * 4 + 1, am.y);
    atomicAdd(dmask_token + c * 4 + 2, am.z);
    atomicAdd(dmask_token + c * 4 + 3, am.w);
  }
}

}  // namespace davf

using namespace davf;

extern "C" int davf_sum_cast(const float* a, const float* b, const float* c, float* out, davf_bf16* out_bf16, int64_t n, davf_stream_t s) {
  DAVF_CHECK_ARG(a && b && out && n % 4 == 0, "sum_cast: null pointer or n=%lld not a multiple of 4", (long long)n);
  DAVF_CHECK_ARG((((uintptr_t)a | (uintptr_t)b | (uintptr_t)c | (uintptr_t)out) & 15) == 0 && ((uintptr_t)out_bf16 & 7) == 0, "sum_cast: misaligned pointer");
  if (n == 0) return DAVF_OK;
  DAVF_CUDA(launch_pdl(sum_cast_kernel, dim3(grid_for(n / 4, 256)), dim3(256), 0, as_stream(s), reinterpret_cast<const float4*>(a),
                       reinterpret_cast<const float4*>(b), reinterpret_cast<const float4*>(c), reinterpret_cast<float4*>(out),
                       reinterpret_cast<uint2*>(out_bf16), n / 4));
  DAVF_LAUNCH_OK();
  return DAVF_OK;
}

extern "C" int davf_patch_rows(const float* img, const int64_t* ids_keep, davf_bf16* out, int B, int C, int H, int W,
                               int p, int nK, davf_stream_t s) {
  DAVF_CHECK_ARG(p > 0 && p % 4 == 0 && H % p == 0 && W % p == 0 && W % 4 == 0, "patch_rows: p=%d H=%d W=%d unsupported", p, H, W);
  DAVF_CHECK_ARG(nK > 0 && nK <= (H / p) * (W / p), "patch_rows: nK=%d", nK);
  if (B == 0) return DAVF_OK;
  const int64_t total = (int64_t)B * nK * C * p * (p / 4);
  DAVF_CUDA(launch_pdl(patch_rows_kernel, dim3(grid_for(total, 256)), dim3(256), 0, as_stream(s), img, ids_keep, out, B, C, H, W, p, nK));
  DAVF_LAUNCH_OK();
  return DAVF_OK;
}

extern "C" int davf_cast_rows_bf16(const float* src, davf_bf16* dst, int64_t M, int D, int g, int G, int off, davf_stream_t s) {
  DAVF_CHECK_ARG(D % 4 == 0 && g > 0 && G >= g && off >= 0 && off + g <= G, "cast_rows: D=%d g=%d G=%d off=%d", D, g, G, off);
  if (M == 0) return DAVF_OK;
  DAVF_CUDA(launch_pdl(cast_rows_kernel, dim3(grid_for(M * (D / 4), 256)), dim3(256), 0, as_stream(s), src, dst, M, D, g, G, off));
  DAVF_LAUNCH_OK();
  return DAVF_OK;
}

extern "C" int davf_cast_flat_bf16(const float* src, davf_bf16* dst, int64_t n, davf_stream_t s) {
  DAVF_CHECK_ARG(n % 4 == 0, "cast_flat: n=%lld must be a multiple of 4", (long long)n);
  if (n == 0) return DAVF_OK;
  DAVF_CUDA(launch_pdl(cast_flat_kernel, dim3(grid_for(n / 4, 256)), dim3(256), 0, as_stream(s), src, dst, n / 4));
  DAVF_LAUNCH_OK();
  return DAVF_OK;
}

extern "C" int davf_colsum_bf16(const davf_bf16* x, int64_t M, int N, int64_t ld, float* out, davf_stream_t s) {
  DAVF_CHECK_ARG(N % 2 == 0 && ld % 2 == 0, "colsum: N=%d ld=%lld must be even", N, (long long)ld);
  if (M == 0 || N == 0) return DAVF_OK;
  const int gx = (N + 63) / 64;
  int slabs = (2 * kNumSMs + gx - 1) / gx;                     // ~2 CTAs per SM in total
  int64_t rpb = (M + slabs - 1) / slabs;
  if (rpb < 64) rpb = 64;
  rpb = (rpb + 7) / 8 * 8;
  const int gy = (int)((M + rpb - 1) / rpb);
  DAVF_CUDA(launch_pdl(colsum_bf16_kernel, dim3(dim3(gx, gy)), dim3(256), 0, as_stream(s), x, M, N, ld, out, (int)rpb));
  DAVF_LAUNCH_OK();
  return DAVF_OK;
}

extern "C" int davf_batchsum_f32(const float* x, int B, int G, int off, int g, int D, float* out, int accumulate, davf_stream_t s) {
  DAVF_CHECK_ARG(D % 4 == 0 && g > 0 && off >= 0 && off + g <= G, "batchsum: D=%d g=%d G=%d off=%d", D, g, G, off);
  DAVF_CUDA(launch_pdl(batchsum_f32_kernel, dim3(grid_for((int64_t)g * (D / 4), 128)), dim3(128), 0, as_stream(s), x, B, G, off, g, D, out, accumulate));
  DAVF_LAUNCH_OK();
  return DAVF_OK;
}

extern "C" int davf_decoder_assemble_fwd(const float* e, const float* ef, const float* mask_token, const float* pos,
                                         const int64_t* ids_restore, float* seq, int B, int nK, int nF, int L, int D,
                                         davf_stream_t s) {
  DAVF_CHECK_ARG(D % 4 == 0 && nK <= L && nF >= 0, "decoder_assemble_fwd: D=%d nK=%d L=%d", D, nK, L);
  if (B == 0) return DAVF_OK;
  DAVF_CUDA(launch_pdl(dec_assemble_fwd_kernel, dim3(grid_for((int64_t)B * (nF + L) * (D / 4), 256)), dim3(256), 0, as_stream(s), 
      e, ef, mask_token, pos, ids_restore, seq, B, nK, nF, L, D));
  DAVF_LAUNCH_OK();
  return DAVF_OK;
}

extern "C" int davf_decoder_assemble_bwd(const float* dseq, const int64_t* ids_keep, const int64_t* ids_restore,
                                         davf_bf16* de, davf_bf16* def_, float* dmask_token, float* dpos,
                                         int B, int nK, int nF, int L, int D, davf_stream_t s) {
  DAVF_CHECK_ARG(D % 4 == 0 && nK <= L, "decoder_assemble_bwd: D=%d nK=%d L=%d", D, nK, L);
  if (B == 0) return DAVF_OK;
  DAVF_CUDA(launch_pdl(dec_assemble_bwd_rows_kernel, dim3(grid_for((int64_t)B * (nK + nF) * (D / 4), 256)), dim3(256), 0, as_stream(s), 
      dseq, ids_keep, de, def_, B, nK, nF, L, D));
  DAVF_LAUNCH_OK();
  DAVF_CUDA(launch_pdl(dec_assemble_bwd_reduce_kernel, dim3(L), dim3(128), 0, as_stream(s), dseq, ids_restore, dmask_token, dpos, B, nK, nF, L, D));
  DAVF_LAUNCH_OK();
  return DAVF_OK;
}
