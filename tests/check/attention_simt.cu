// K6/K7/K8/K11: fused small-sequence attention, forward and backward.
// Every (batch, head) problem on this path is tiny (Nk <= 228, head dims 64/32/16), so the whole
// K/V (and for backward Q/K/V/dO) of one head lives in shared memory of one CTA: no online-softmax
// loop and no HBM round trip for the score matrix.  Strided q/k/v/o addressing lets the kernel read
// packed qkv buffers, query sub-ranges (live rows only) and write packed dqkv buffers directly.
//
// This file is the CUDA-core (f32 FMA) CHECKER implementation (tests/check/libdavf_check.so, test infrastructure);
// the product kernels are csrc/attention_tc.cu (tcgen05) and csrc/attention_mma.cu (mma.sync).
#include "common.cuh"

namespace davf {

constexpr int kMaxKeys = 256;

template <int D>
__device__ __forceinline__ void load_rows_bf16(uint16_t* dst, int dst_stride, const uint16_t* src, int64_t src_rs, int rows) {
  // rows x D bf16, 8-byte (4 element) vectors; dst_stride in elements (even)
  constexpr int VPR = D / 4;
  for (int i = threadIdx.x; i < rows * VPR; i += blockDim.x) {
    const int r = i / VPR, c = i - r * VPR;
    const uint2 v = *reinterpret_cast<const uint2*>(src + (int64_t)r * src_rs + c * 4);
    uint32_t* d = reinterpret_cast<uint32_t*>(dst + r * dst_stride + c * 4);
    d[0] = v.x;
    d[1] = v.y;
  }
}

template <int D>
__device__ __forceinline__ float dot_smem(const float* __restrict__ a, const uint16_t* __restrict__ row) {
  float acc = 0.f;
#pragma unroll
  for (int d = 0; d < D; d += 2) {
    const float2 k = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(row + d));
    acc = fmaf(a[d], k.x, acc);
    acc = fmaf(a[d + 1], k.y, acc);
  }
  return acc;
}

template <int D>
__device__ __forceinline__ float dot_smem_bf(const uint16_t* __restrict__ a, const uint16_t* __restrict__ row) {
  float acc = 0.f;
#pragma unroll
  for (int d = 0; d < D; d += 2) {
    const float2 x = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(a + d));
    const float2 k = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(row + d));
    acc = fmaf(x.x, k.x, acc);
    acc = fmaf(x.y, k.y, acc);
  }
  return acc;
}

// ---------------------------------------------------------------------------------------------
// forward: grid (query tiles of QT rows, B*H), 128 threads
// ---------------------------------------------------------------------------------------------
constexpr int QT = 64;

template <int DQK, int DV>
__global__ void __launch_bounds__(128) attn_fwd_kernel(davf_attn_fwd_args a) {
  extern __shared__ __align__(16) uint8_t smem[];
  constexpr int KS = DQK + 2;                     // padded K row stride (odd number of 32-bit words)
  uint16_t* Ks = reinterpret_cast<uint16_t*>(smem);                       // [Nk][KS]
  uint16_t* Vs = Ks + a.Nk * KS + ((a.Nk * KS) & 1);                      // [Nk][DV]
  Vs = reinterpret_cast<uint16_t*>((reinterpret_cast<uintptr_t>(Vs) + 15) & ~uintptr_t(15));
  float* Ps = reinterpret_cast<float*>(Vs + a.Nk * DV);                   // [4][kMaxKeys]
  float* Qs = Ps + 4 * kMaxKeys;                                          // [4][DQK]
  const int bh = blockIdx.y, b = bh / a.H, h = bh - b * a.H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint16_t* kg = a.k + (int64_t)b * a.k_bs + h * DQK;
  const uint16_t* vg = a.v + (int64_t)b * a.v_bs + h * DV;
  load_rows_bf16<DQK>(Ks, KS, kg, a.k_rs, a.Nk);
  load_rows_bf16<DV>(Vs, DV, vg, a.v_rs, a.Nk);
  __syncthreads();
  float* ps = Ps + warp * kMaxKeys;
  float* qs = Qs + warp * DQK;
  const int q_end = min(a.Nq, (int)(blockIdx.x + 1) * QT);
  for (int i = blockIdx.x * QT + warp; i < q_end; i += 4) {
    const uint16_t* qg = a.q + (int64_t)b * a.q_bs + (int64_t)i * a.q_rs + h * DQK;
    for (int d = lane; d < DQK; d += 32) qs[d] = bf16_to_f32(qg[d]) * a.scale;
    __syncwarp();
    float s[kMaxKeys / 32];
    float mx = -INFINITY;
#pragma unroll
    for (int t = 0; t < kMaxKeys / 32; ++t) {
      const int j = lane + 32 * t;
      s[t] = -INFINITY;
      if (j < a.Nk) {
        s[t] = dot_smem<DQK>(qs, Ks + j * KS);
        mx = fmaxf(mx, s[t]);
      }
    }
    mx = warp_max(mx);
    float l = 0.f;
#pragma unroll
    for (int t = 0; t < kMaxKeys / 32; ++t) {
      const int j = lane + 32 * t;
      if (j < a.Nk) {
        const float p = __expf(s[t] - mx);
        ps[j] = p;
        l += p;
      }
    }
    l = warp_sum(l);
    __syncwarp();
    const float inv_l = 1.0f / l;
    uint16_t* og = a.o + (int64_t)b * a.o_bs + (int64_t)i * a.o_rs + h * DV;
    if (DV == 64) {
      float2 acc = make_float2(0.f, 0.f);
      for (int j = 0; j < a.Nk; ++j) {
        const float p = ps[j];
        const float2 v = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(Vs + j * DV + 2 * lane));
        acc.x = fmaf(p, v.x, acc.x);
        acc.y = fmaf(p, v.y, acc.y);
      }
      acc.x *= inv_l;
      acc.y *= inv_l;
      if (a.accumulate) {
        const float2 o = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(og + 2 * lane));
        acc.x += o.x;
        acc.y += o.y;
      }
      *reinterpret_cast<uint32_t*>(og + 2 * lane) = pack_bf16x2(acc.x, acc.y);
    } else {   // DV == 32: one column per lane
      float acc = 0.f;
      for (int j = 0; j < a.Nk; ++j) acc = fmaf(ps[j], bf16_to_f32(Vs[j * DV + lane]), acc);
      acc *= inv_l;
      if (a.accumulate) acc += bf16_to_f32(og[lane]);
      og[lane] = f32_to_bf16(acc);
    }
    if (lane == 0 && a.lse) a.lse[((int64_t)b * a.H + h) * a.Nq + i] = mx + __logf(l);
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------
// backward: one CTA per (b, h), 256 threads; Q, K, V, dO of the head in shared memory.
//   phase A (row owner = query i):  dq_i = scale * sum_j dS_ij k_j
//   phase B (row owner = key j):    dk_j = scale * sum_i dS_ij q_i ;  dv_j = sum_i P_ij dO_i
//   P_ij = exp(scale q_i.k_j - lse_i),  dS_ij = P_ij (dP_ij - D_i),  dP_ij = dO_i.v_j,
//   D_i = sum_j P_ij dP_ij  (from the recomputed probabilities; the forward output is not needed)
// ---------------------------------------------------------------------------------------------
template <int DQK, int DV>
__global__ void __launch_bounds__(256) attn_bwd_kernel(davf_attn_bwd_args a) {
  extern __shared__ __align__(16) uint8_t smem[];
  constexpr int QS = DQK + 2, VS = DV + 2;
  constexpr int NW = 8;
  const int Nq = a.Nq, Nk = a.Nk;
  uint16_t* Qs = reinterpret_cast<uint16_t*>(smem);      // [Nq][QS]
  uint16_t* Ks = Qs + Nq * QS;                            // [Nk][QS]
  uint16_t* Vs = Ks + Nk * QS;                            // [Nk][VS]
  uint16_t* dOs = Vs + Nk * VS;                           // [Nq][VS]
  float* fbase = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(dOs + Nq * VS) + 15) & ~uintptr_t(15));
  float* Ds = fbase;                                      // [Nq]
  float* Ls = Ds + kMaxKeys;                              // [Nq]
  float* W1 = Ls + kMaxKeys;                              // [NW][kMaxKeys]  dS (scaled)
  float* W2 = W1 + NW * kMaxKeys;                         // [NW][kMaxKeys]  P
  const int bh = blockIdx.x, b = bh / a.H, h = bh - b * a.H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  load_rows_bf16<DQK>(Qs, QS, a.q + (int64_t)b * a.q_bs + h * DQK, a.q_rs, Nq);
  load_rows_bf16<DQK>(Ks, QS, a.k + (int64_t)b * a.k_bs + h * DQK, a.k_rs, Nk);
  load_rows_bf16<DV>(Vs, VS, a.v + (int64_t)b * a.v_bs + h * DV, a.v_rs, Nk);
  load_rows_bf16<DV>(dOs, VS, a.d_o + (int64_t)b * a.do_bs + h * DV, a.do_rs, Nq);
  __syncthreads();
  for (int i = threadIdx.x; i < Nq; i += blockDim.x) Ls[i] = a.lse[((int64_t)b * a.H + h) * Nq + i];
  __syncthreads();
  float* w1 = W1 + warp * kMaxKeys;
  float* w2 = W2 + warp * kMaxKeys;

  // ---- phase A: D_i = sum_j P_ij dP_ij (kept in smem for phase B) and dq ----
  for (int i = warp; i < Nq; i += NW) {
    const float Li = Ls[i];
    float dsum = 0.f;
    for (int j = lane; j < Nk; j += 32) {
      const float s = dot_smem_bf<DQK>(Qs + i * QS, Ks + j * QS) * a.scale;
      const float p = __expf(s - Li);
      const float dp = dot_smem_bf<DV>(dOs + i * VS, Vs + j * VS);
      w1[j] = p;
      w2[j] = dp;
      dsum = fmaf(p, dp, dsum);
    }
    const float Di = warp_sum(dsum);
    if (lane == 0) Ds[i] = Di;
    for (int j = lane; j < Nk; j += 32) w1[j] = w1[j] * (w2[j] - Di) * a.scale;
    __syncwarp();
    uint16_t* dqg = a.dq + (int64_t)b * a.dq_bs + (int64_t)i * a.dq_rs + h * DQK;
    for (int d = lane; d < DQK; d += 32) {
      float acc = 0.f;
      for (int j = 0; j < Nk; ++j) acc = fmaf(w1[j], bf16_to_f32(Ks[j * QS + d]), acc);
      if (a.accumulate_dq) acc += bf16_to_f32(dqg[d]);
      dqg[d] = f32_to_bf16(acc);
    }
    __syncwarp();
  }
  __syncthreads();
  // ---- phase B: dk, dv ----
  for (int j = warp; j < Nk; j += NW) {
    for (int i = lane; i < Nq; i += 32) {
      const float s = dot_smem_bf<DQK>(Qs + i * QS, Ks + j * QS) * a.scale;
      const float p = __expf(s - Ls[i]);
      const float dp = dot_smem_bf<DV>(dOs + i * VS, Vs + j * VS);
      w1[i] = p * (dp - Ds[i]) * a.scale;
      w2[i] = p;
    }
    __syncwarp();
    uint16_t* dkg = a.dk + (int64_t)b * a.dk_bs + (int64_t)j * a.dk_rs + h * DQK;
    uint16_t* dvg = a.dv_ + (int64_t)b * a.dv_bs + (int64_t)j * a.dv_rs + h * DV;
    for (int d = lane; d < DQK; d += 32) {
      float acc = 0.f;
      for (int i = 0; i < Nq; ++i) acc = fmaf(w1[i], bf16_to_f32(Qs[i * QS + d]), acc);
      dkg[d] = f32_to_bf16(acc);
    }
    for (int d = lane; d < DV; d += 32) {
      float acc = 0.f;
      for (int i = 0; i < Nq; ++i) acc = fmaf(w2[i], bf16_to_f32(dOs[i * VS + d]), acc);
      dvg[d] = f32_to_bf16(acc);
    }
    __syncwarp();
  }
}

template <int DQK, int DV>
static int launch_fwd(const davf_attn_fwd_args& a, cudaStream_t st) {
  const size_t smem = (size_t)a.Nk * (DQK + 2) * 2 + 32 + (size_t)a.Nk * DV * 2 + 4 * kMaxKeys * 4 + 4 * DQK * 4;
  auto kern = attn_fwd_kernel<DQK, DV>;
  static size_t configured = 0;
  if (smem > configured) {
    DAVF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  dim3 grid((a.Nq + QT - 1) / QT, a.B * a.H);
  kern<<<grid, 128, smem, st>>>(a);
  DAVF_LAUNCH_OK();
  return DAVF_OK;
}

template <int DQK, int DV>
static int launch_bwd(const davf_attn_bwd_args& a, cudaStream_t st) {
  const size_t smem = ((size_t)(a.Nq + a.Nk) * (DQK + 2) + (size_t)(a.Nq + a.Nk) * (DV + 2)) * 2 + 16 +
                      (2 * kMaxKeys + 2 * 8 * kMaxKeys) * 4;
  auto kern = attn_bwd_kernel<DQK, DV>;
  static size_t configured = 0;
  if (smem > configured) {
    DAVF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  kern<<<a.B * a.H, 256, smem, st>>>(a);
  DAVF_LAUNCH_OK();
  return DAVF_OK;
}

static bool strides_ok(int64_t rs, int64_t bs) { return rs % 4 == 0 && bs % 4 == 0; }

}  // namespace davf


namespace davf {
int attn_simt_fwd(const davf_attn_fwd_args* a, cudaStream_t st) {
  DAVF_CHECK_ARG(a && a->q && a->k && a->v && a->o, "attention_fwd: null pointer");
  DAVF_CHECK_ARG(a->Nk > 0 && a->Nk <= kMaxKeys && a->Nq > 0 && a->H > 0 && a->B >= 0, "attention_fwd: Nq=%d Nk=%d (Nk <= %d)", a->Nq, a->Nk, kMaxKeys);
  DAVF_CHECK_ARG(strides_ok(a->q_rs, a->q_bs) && strides_ok(a->k_rs, a->k_bs) && strides_ok(a->v_rs, a->v_bs) && strides_ok(a->o_rs, a->o_bs),
                 "attention_fwd: strides must be multiples of 4 elements");
  DAVF_CHECK_ARG((((uintptr_t)a->q | (uintptr_t)a->k | (uintptr_t)a->v | (uintptr_t)a->o) & 7) == 0, "attention_fwd: pointers must be 8-byte aligned");
  if (a->B == 0) return DAVF_OK;
  if (a->dqk == 64 && a->dv == 64) return launch_fwd<64, 64>(*a, st);
  if (a->dqk == 32 && a->dv == 32) return launch_fwd<32, 32>(*a, st);
  if (a->dqk == 16 && a->dv == 64) return launch_fwd<16, 64>(*a, st);
  set_error("attention_fwd: head dims (%d,%d) unsupported", a->dqk, a->dv);
  return DAVF_EUNSUPPORTED;
}

int attn_simt_bwd(const davf_attn_bwd_args* a, cudaStream_t st) {
  DAVF_CHECK_ARG(a && a->q && a->k && a->v && a->d_o && a->lse && a->dq && a->dk && a->dv_, "attention_bwd: null pointer");
  DAVF_CHECK_ARG(a->Nk > 0 && a->Nk <= kMaxKeys && a->Nq > 0 && a->Nq <= kMaxKeys, "attention_bwd: Nq=%d Nk=%d (<= %d)", a->Nq, a->Nk, kMaxKeys);
  DAVF_CHECK_ARG(strides_ok(a->q_rs, a->q_bs) && strides_ok(a->k_rs, a->k_bs) && strides_ok(a->v_rs, a->v_bs) &&
                     strides_ok(a->do_rs, a->do_bs) && strides_ok(a->dq_rs, a->dq_bs) && strides_ok(a->dk_rs, a->dk_bs) && strides_ok(a->dv_rs, a->dv_bs),
                 "attention_bwd: strides must be multiples of 4 elements");
  if (a->B == 0) return DAVF_OK;
  if (a->dqk == 64 && a->dv == 64) return launch_bwd<64, 64>(*a, st);
  if (a->dqk == 32 && a->dv == 32) return launch_bwd<32, 32>(*a, st);
  if (a->dqk == 16 && a->dv == 64) return launch_bwd<16, 64>(*a, st);
  set_error("attention_bwd: head dims (%d,%d) unsupported", a->dqk, a->dv);
  return DAVF_EUNSUPPORTED;
}
}  // namespace davf
