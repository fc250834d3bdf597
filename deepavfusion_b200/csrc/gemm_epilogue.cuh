// Fused GEMM epilogue shared by the tcgen05 kernel and the SIMT checker kernel (see davf.h for
// the exact order of operations).  Works on NV contiguous columns of one output row.
#pragma once
#include "common.cuh"

namespace davf {

struct EpiParams {
  const float* bias;
  int act;
  uint16_t* aux_out;
  const uint16_t* aux_in;
  int64_t ldaux;
  const float* res;
  int64_t ldres;
  const int64_t* res_idx;
  void* out;
  int64_t ldo;
  int out_bf16;
  int accumulate;
  int g, G, off;
  int64_t M, N;
  float* rowsum_out;
  long long* debug_clocks;
};

static inline EpiParams make_epi(const davf_gemm_args& a) {
  EpiParams p;
  p.bias = a.bias; p.act = a.act; p.aux_out = a.aux_out; p.aux_in = a.aux_in; p.ldaux = a.ldaux;
  p.res = a.res; p.ldres = a.ldres; p.res_idx = a.res_idx; p.out = a.out; p.ldo = a.ldo;
  p.out_bf16 = a.out_bf16; p.accumulate = a.accumulate; p.g = a.g; p.G = a.G; p.off = a.off;
  p.M = a.M; p.N = a.N; p.rowsum_out = a.rowsum_out;
  p.debug_clocks = reinterpret_cast<long long*>(a.debug_clocks);
  return p;
}

// z[0..NV) are the f32 accumulators of row m, columns n0..n0+NV.  NV % 4 == 0, n0 % 4 == 0 and
// N % 4 == 0 (checked on the host) so that a 4-column group is either fully in or fully out of
// bounds and every vector access is aligned.
template <int NV>
__device__ __forceinline__ void epilogue_row(const EpiParams& p, int64_t m, int64_t n0, float (&z)[NV], bool add_bias) {
  if (m >= p.M) return;
  const int64_t orow = p.g > 0 ? (m / p.g) * (int64_t)p.G + p.off + (m % p.g) : m;
  const int64_t rrow = p.res ? (p.res_idx ? p.res_idx[m] : orow) : 0;
#pragma unroll
  for (int c = 0; c < NV; c += 4) {
    const int64_t n = n0 + c;
    if (n >= p.N) break;
    float4 v = make_float4(z[c], z[c + 1], z[c + 2], z[c + 3]);
    if (p.bias && add_bias) {
      const float4 b = *reinterpret_cast<const float4*>(p.bias + n);
      v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
    }
    float4 ax = v;                                   // what aux_out receives: z, or gelu'(z) next to a GELU
    if (p.act == DAVF_ACT_GELU) {
      gelu_both(v.x, v.x, ax.x); gelu_both(v.y, v.y, ax.y); gelu_both(v.z, v.z, ax.z); gelu_both(v.w, v.w, ax.w);
    } else if (p.act == DAVF_ACT_DGELU) {
      const uint2 u = *reinterpret_cast<const uint2*>(p.aux_in + m * p.ldaux + n);
      const float2 lo = unpack_bf16x2(u.x), hi = unpack_bf16x2(u.y);
      v.x *= lo.x; v.y *= lo.y; v.z *= hi.x; v.w *= hi.y;
    }
    if (p.aux_out) {
      uint2 o;
      o.x = pack_bf16x2(ax.x, ax.y);
      o.y = pack_bf16x2(ax.z, ax.w);
      *reinterpret_cast<uint2*>(p.aux_out + m * p.ldaux + n) = o;
    }
    if (p.res) {
      const float4 r = *reinterpret_cast<const float4*>(p.res + rrow * p.ldres + n);
      v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
    }
    if (p.accumulate) {
      float* o = reinterpret_cast<float*>(p.out) + orow * p.ldo + n;
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    } else if (p.out_bf16) {
      uint2 o;
      o.x = pack_bf16x2(v.x, v.y);
      o.y = pack_bf16x2(v.z, v.w);
      *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(p.out) + orow * p.ldo + n) = o;
    } else {
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + orow * p.ldo + n) = v;
    }
  }
}

int gemm_simt_launch(const davf_gemm_args& a, cudaStream_t st);      // tests/check/libdavf_check.so only (CUDA-core checker)
int gemm_tc_launch(const davf_gemm_args& a, cudaStream_t st);
int gemm_tc_launch_grouped(const davf_gemm_args* a, int count, cudaStream_t st);   // count <= DAVF_GEMM_MAX_GROUP, one layout class

}  // namespace davf
