// K6/K7/K8/K11: fused small-sequence attention on the tensor cores, forward and backward.
//
// Every (batch, head) problem on this path is tiny (Nq, Nk <= 228; head dims 64 / 32 / 16), so one
// CTA keeps the whole K/V (backward: Q, K, V, dO) of a head in shared memory and the score matrix
// never touches HBM.  Tiles are far too small for a 128-row tcgen05 atom to pay (an encoder problem
// is 49 x 81 x 64), so the contractions use warp-level mma.sync.m16n8k16 (bf16 in, f32 accumulate)
// with ldmatrix-fed fragments: a warp owns 16 query rows (forward / dQ) or 16 key rows (dK / dV)
// and sweeps the other sequence in chunks of 64 with an online softmax.  QK^T and PV are 3.4 % of
// the step's FLOPs; the kernel's job is removing their HBM round trips.
//
// Strided q / k / v / o addressing (element strides per batch and per row, heads packed along the
// row) reads packed qkv buffers and live-query sub-ranges in place and writes packed dqkv buffers.
#include <stdlib.h>
#include "common.cuh"

namespace davf {

constexpr int CH = 64;          // keys (or queries) per chunk = 8 n-tiles of 8

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t s_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// rows x D bf16 from global (row stride rs elements) into smem with row stride D+8; rows >= valid are zero
template <int D>
__device__ __forceinline__ void load_tile(uint16_t* dst, const uint16_t* src, int64_t rs, int valid, int rows_padded) {
  constexpr int VPR = D / 8;
  constexpr int ST = D + 8;
  for (int i = threadIdx.x; i < rows_padded * VPR; i += blockDim.x) {
    const int r = i / VPR, c = i - r * VPR;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (r < valid) v = *reinterpret_cast<const uint4*>(src + (int64_t)r * rs + c * 8);
    *reinterpret_cast<uint4*>(dst + r * ST + c * 8) = v;
  }
}

// A fragments (16 rows x D) of rows [row0, row0+16) from a [rows][D+8] smem tile
template <int D>
__device__ __forceinline__ void load_a_frags(uint32_t (&f)[D / 16][4], const uint16_t* tile, int row0, int lane) {
  constexpr int ST = D + 8;
#pragma unroll
  for (int ks = 0; ks < D / 16; ++ks)
    ldsm_x4(f[ks], s_addr(tile + (row0 + (lane & 15)) * ST + ks * 16 + (lane >> 4) * 8));
}

// acc[8][4] (16 x 64) += A(16 x D) * B^T where B rows [n0, n0+64) of an n-major [n][D+8] tile
template <int D>
__device__ __forceinline__ void mma_nmajor(float (&acc)[8][4], const uint32_t (&a)[D / 16][4], const uint16_t* tile, int n0, int lane) {
  constexpr int ST = D + 8;
#pragma unroll
  for (int ks = 0; ks < D / 16; ++ks) {
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t b[4];
      ldsm_x4(b, s_addr(tile + (n0 + np * 16 + (lane & 7) + ((lane >> 4) << 3)) * ST + ks * 16 + ((lane >> 3) & 1) * 8));
      mma_bf16(acc[2 * np], a[ks], b[0], b[1]);
      mma_bf16(acc[2 * np + 1], a[ks], b[2], b[3]);
    }
  }
}

// acc[D/8][4] (16 x D) += A(16 x 64, bf16 fragments pa[4]) * B where B rows [k0, k0+64) of a k-major [k][D+8] tile
template <int D>
__device__ __forceinline__ void mma_kmajor(float (&acc)[D / 8][4], const uint32_t (&pa)[4][4], const uint16_t* tile, int k0, int lane) {
  constexpr int ST = D + 8;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
    for (int np = 0; np < D / 16; ++np) {
      uint32_t b[4];
      ldsm_x4_t(b, s_addr(tile + (k0 + ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * ST + np * 16 + (lane >> 4) * 8));
      mma_bf16(acc[2 * np], pa[ks], b[0], b[1]);
      mma_bf16(acc[2 * np + 1], pa[ks], b[2], b[3]);
    }
  }
}

// C fragments of a 16 x 64 f32 tile -> A fragments (bf16) for a following 16 x 64 (k) contraction
__device__ __forceinline__ void c_to_a(uint32_t (&pa)[4][4], const float (&c)[8][4]) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    pa[ks][0] = pack_bf16x2(c[2 * ks][0], c[2 * ks][1]);
    pa[ks][1] = pack_bf16x2(c[2 * ks][2], c[2 * ks][3]);
    pa[ks][2] = pack_bf16x2(c[2 * ks + 1][0], c[2 * ks + 1][1]);
    pa[ks][3] = pack_bf16x2(c[2 * ks + 1][2], c[2 * ks + 1][3]);
  }
}

__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}
__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
  return v;
}

// 2^x on the MUFU pipe with no range fix-up code (exp2f without fast-math costs 2 FMUL + FSETP per call; the
// arguments here are <= 0 up to rounding, and flush-to-zero of tiny probabilities is harmless)
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}


// o (+)= (x, y) as one packed bf16 pair.  mode 0: store; 1: read-modify-write (the caller orders the launches that add into
// the same buffer); 2: red.global.add.bf16x2 -- launches that add into the same (zero-initialised) buffer may run concurrently
__device__ __forceinline__ void put_pair(uint32_t* p, float x, float y, int mode) {
  if (mode == 2) {
    atomicAdd(reinterpret_cast<__nv_bfloat162*>(p), __floats2bfloat162_rn(x, y));
    return;
  }
  if (mode == 1) { const float2 old = unpack_bf16x2(*p); x += old.x; y += old.y; }
  *p = pack_bf16x2(x, y);
}

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

// ---------------------------------------------------------------------------------------------
// forward: grid (ceil(Nq / (16 NW)), B*H), NW warps; warp = 16 query rows (NW = 1 / 2 / 4 by problem size)
// ---------------------------------------------------------------------------------------------
template <int DQK, int DV, int NW>
__global__ void __launch_bounds__(NW * 32) attn_mma_fwd_kernel(davf_attn_fwd_args a, int Nkp) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ __align__(16) uint8_t smem[];
  constexpr int QT = NW * 16;                                  // query rows per CTA
  uint16_t* Qs = reinterpret_cast<uint16_t*>(smem);            // [QT][DQK+8]
  uint16_t* Ks = Qs + QT * (DQK + 8);                          // [Nkp][DQK+8]
  uint16_t* Vs = Ks + Nkp * (DQK + 8);                         // [Nkp][DV+8]
  const int bh = blockIdx.y, b = bh / a.H, h = bh - b * a.H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int q0 = blockIdx.x * QT;
  load_tile<DQK>(Qs, a.q + (int64_t)b * a.q_bs + (int64_t)q0 * a.q_rs + h * DQK, a.q_rs, min(QT, a.Nq - q0), QT);
  load_tile<DQK>(Ks, a.k + (int64_t)b * a.k_bs + h * DQK, a.k_rs, a.Nk, Nkp);
  load_tile<DV>(Vs, a.v + (int64_t)b * a.v_bs + h * DV, a.v_rs, a.Nk, Nkp);
  __syncthreads();
  if (q0 + warp * 16 >= a.Nq) return;

  uint32_t qf[DQK / 16][4];
  load_a_frags<DQK>(qf, Qs, warp * 16, lane);
  const float sl = a.scale * kLog2e;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;      // rows g and g+8 (log2 domain max, partial sums)
  float o[DV / 8][4];
#pragma unroll
  for (int i = 0; i < DV / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;

  for (int c0 = 0; c0 < a.Nk; c0 += CH) {
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
    mma_nmajor<DQK>(s, qf, Ks, c0, lane);
    float mx0 = -INFINITY, mx1 = -INFINITY;
    if (c0 + CH <= a.Nk) {               // full chunk: no key masks
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) s[nt][e] *= sl;
        mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
        mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
      }
    } else {
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int col = c0 + nt * 8 + 2 * t;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const bool ok = (col + (e & 1)) < a.Nk;
          s[nt][e] = ok ? s[nt][e] * sl : -INFINITY;
        }
        mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
        mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
      }
    }
    const float mn0 = fmaxf(m0, quad_max(mx0)), mn1 = fmaxf(m1, quad_max(mx1));
    const float al0 = ex2(m0 - mn0), al1 = ex2(m1 - mn1);
    m0 = mn0; m1 = mn1;
    l0 *= al0; l1 *= al1;
#pragma unroll
    for (int i = 0; i < DV / 8; ++i) { o[i][0] *= al0; o[i][1] *= al0; o[i][2] *= al1; o[i][3] *= al1; }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = ex2(s[nt][0] - m0); s[nt][1] = ex2(s[nt][1] - m0);
      s[nt][2] = ex2(s[nt][2] - m1); s[nt][3] = ex2(s[nt][3] - m1);
      l0 += s[nt][0] + s[nt][1];
      l1 += s[nt][2] + s[nt][3];
    }
    uint32_t pa[4][4];
    c_to_a(pa, s);
    mma_kmajor<DV>(o, pa, Vs, c0, lane);
  }
  l0 = quad_sum(l0); l1 = quad_sum(l1);
  const float il0 = 1.0f / l0, il1 = 1.0f / l1;
  const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;
  uint16_t* ob = a.o + (int64_t)b * a.o_bs + h * DV;
#pragma unroll
  for (int nt = 0; nt < DV / 8; ++nt) {
    const int col = nt * 8 + 2 * t;
    if (r0 < a.Nq) {
      uint32_t* p = reinterpret_cast<uint32_t*>(ob + (int64_t)r0 * a.o_rs + col);
      put_pair(p, o[nt][0] * il0, o[nt][1] * il0, a.accumulate);
    }
    if (r1 < a.Nq) {
      uint32_t* p = reinterpret_cast<uint32_t*>(ob + (int64_t)r1 * a.o_rs + col);
      put_pair(p, o[nt][2] * il1, o[nt][3] * il1, a.accumulate);
    }
  }
  if (a.lse && t == 0) {
    float* lp = a.lse + ((int64_t)b * a.H + h) * a.Nq;
    if (r0 < a.Nq) lp[r0] = (m0 + log2f(l0)) * kLn2;
    if (r1 < a.Nq) lp[r1] = (m1 + log2f(l1)) * kLn2;
  }
}

// ---------------------------------------------------------------------------------------------
// backward: one CTA per (b, h), NW warps (2 / 4 / 8 by problem size); Q, K, V, dO of the head in shared memory.
//   phase A (warp = 16 query rows): pass 1  D_i = sum_j P_ij dP_ij ; pass 2  dQ = scale * dS K
//   phase B (warp = 16 key rows):   dV = P^T dO ; dK = scale * dS^T Q      (transposed tiles recomputed)
//   P = exp(scale S - lse), dP = dO V^T, dS = P (dP - D)
// ---------------------------------------------------------------------------------------------
template <int DQK, int DV, int NW>
__global__ void __launch_bounds__(NW * 32) attn_mma_bwd_kernel(davf_attn_bwd_args a, int Nqp, int Nkp) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ __align__(16) uint8_t smem[];
  uint16_t* Qs = reinterpret_cast<uint16_t*>(smem);            // [Nqp][DQK+8]
  uint16_t* Ks = Qs + Nqp * (DQK + 8);                         // [Nkp][DQK+8]
  uint16_t* Vs = Ks + Nkp * (DQK + 8);                         // [Nkp][DV+8]
  uint16_t* dOs = Vs + Nkp * (DV + 8);                         // [Nqp][DV+8]
  float* Ls = reinterpret_cast<float*>(dOs + Nqp * (DV + 8));  // [Nqp]  lse * log2e
  float* Ds = Ls + Nqp;                                        // [Nqp]
  const int bh = blockIdx.x, b = bh / a.H, h = bh - b * a.H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int Nq = a.Nq, Nk = a.Nk;
  load_tile<DQK>(Qs, a.q + (int64_t)b * a.q_bs + h * DQK, a.q_rs, Nq, Nqp);
  load_tile<DQK>(Ks, a.k + (int64_t)b * a.k_bs + h * DQK, a.k_rs, Nk, Nkp);
  load_tile<DV>(Vs, a.v + (int64_t)b * a.v_bs + h * DV, a.v_rs, Nk, Nkp);
  load_tile<DV>(dOs, a.d_o + (int64_t)b * a.do_bs + h * DV, a.do_rs, Nq, Nqp);
  for (int i = threadIdx.x; i < Nqp; i += blockDim.x) {
    Ls[i] = i < Nq ? a.lse[((int64_t)b * a.H + h) * Nq + i] * kLog2e : 0.f;
    Ds[i] = 0.f;     // padded queries are never written by phase A; 0 * (uninitialised NaN) would poison dK in phase B
  }
  __syncthreads();
  const float sl = a.scale * kLog2e;
  const bool have_o = a.o != nullptr;
  if (have_o) {      // one-pass form: D_i = dO_i . O_i  (one thread per query row, 16-byte loads, all rows in flight)
    for (int i = threadIdx.x; i < Nq; i += blockDim.x) {
      const uint16_t* orow = a.o + (int64_t)b * a.o_bs + (int64_t)i * a.o_rs + h * DV;
      float acc = 0.f;
#pragma unroll
      for (int d = 0; d < DV; d += 8) {
        const uint4 xo = *reinterpret_cast<const uint4*>(orow + d);
        const uint4 yo = *reinterpret_cast<const uint4*>(dOs + i * (DV + 8) + d);
        const uint32_t xs[4] = {xo.x, xo.y, xo.z, xo.w}, ys[4] = {yo.x, yo.y, yo.z, yo.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 x = unpack_bf16x2(xs[e]), y = unpack_bf16x2(ys[e]);
          acc = fmaf(x.x, y.x, fmaf(x.y, y.y, acc));
        }
      }
      Ds[i] = acc;
    }
    __syncthreads();
  }

  // ---------------- phase A ----------------
  for (int qt = warp; qt * 16 < Nq; qt += NW) {
    uint32_t aq[DQK / 16][4], ado[DV / 16][4];
    load_a_frags<DQK>(aq, Qs, qt * 16, lane);
    load_a_frags<DV>(ado, dOs, qt * 16, lane);
    const int r0 = qt * 16 + g, r1 = r0 + 8;
    const float L0 = Ls[r0], L1 = Ls[r1];
    float D0 = have_o ? Ds[r0] : 0.f, D1 = have_o ? Ds[r1] : 0.f;
    for (int pass = have_o ? 1 : 0; pass < 2; ++pass) {
      float dq[DQK / 8][4];
#pragma unroll
      for (int i = 0; i < DQK / 8; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;
      for (int c0 = 0; c0 < Nk; c0 += CH) {
        float s[8][4], dp[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f; }
        mma_nmajor<DQK>(s, aq, Ks, c0, lane);
        mma_nmajor<DV>(dp, ado, Vs, c0, lane);
        // Padded key columns (rows >= Nk of the K / V tiles are zero): P is masked to 0 there, in the last chunk only.
        const bool partial = c0 + CH > Nk;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          float p0 = ex2(fmaf(s[nt][0], sl, -L0)), p1 = ex2(fmaf(s[nt][1], sl, -L0));
          float p2 = ex2(fmaf(s[nt][2], sl, -L1)), p3 = ex2(fmaf(s[nt][3], sl, -L1));
          if (partial) {
            const int col = c0 + nt * 8 + 2 * t;
            const bool ok0 = col < Nk, ok1 = col + 1 < Nk;
            p0 = ok0 ? p0 : 0.f; p1 = ok1 ? p1 : 0.f; p2 = ok0 ? p2 : 0.f; p3 = ok1 ? p3 : 0.f;
          }
          if (pass == 0) {
            D0 += p0 * dp[nt][0] + p1 * dp[nt][1];
            D1 += p2 * dp[nt][2] + p3 * dp[nt][3];
          } else {       // dS without the softmax scale: it is applied once to the dQ accumulators
            s[nt][0] = p0 * (dp[nt][0] - D0); s[nt][1] = p1 * (dp[nt][1] - D0);
            s[nt][2] = p2 * (dp[nt][2] - D1); s[nt][3] = p3 * (dp[nt][3] - D1);
          }
        }
        if (pass == 1) {
          uint32_t dsa[4][4];
          c_to_a(dsa, s);
          mma_kmajor<DQK>(dq, dsa, Ks, c0, lane);
        }
      }
      if (pass == 0) {
        D0 = quad_sum(D0); D1 = quad_sum(D1);
        if (t == 0) { Ds[r0] = D0; Ds[r1] = D1; }
      } else {
        uint16_t* qb = a.dq + (int64_t)b * a.dq_bs + h * DQK;
#pragma unroll
        for (int nt = 0; nt < DQK / 8; ++nt) {
          const int col = nt * 8 + 2 * t;
          if (r0 < Nq) {
            uint32_t* p = reinterpret_cast<uint32_t*>(qb + (int64_t)r0 * a.dq_rs + col);
            put_pair(p, dq[nt][0] * a.scale, dq[nt][1] * a.scale, a.accumulate_dq);
          }
          if (r1 < Nq) {
            uint32_t* p = reinterpret_cast<uint32_t*>(qb + (int64_t)r1 * a.dq_rs + col);
            put_pair(p, dq[nt][2] * a.scale, dq[nt][3] * a.scale, a.accumulate_dq);
          }
        }
      }
    }
  }
  __syncthreads();

  // ---------------- phase B ----------------
  for (int kt = warp; kt * 16 < Nk; kt += NW) {
    uint32_t ak[DQK / 16][4], av[DV / 16][4];
    load_a_frags<DQK>(ak, Ks, kt * 16, lane);
    load_a_frags<DV>(av, Vs, kt * 16, lane);
    float dk[DQK / 8][4], dv[DV / 8][4];
#pragma unroll
    for (int i = 0; i < DQK / 8; ++i) dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f;
#pragma unroll
    for (int i = 0; i < DV / 8; ++i) dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f;
    for (int c0 = 0; c0 < Nq; c0 += CH) {
      float st[8][4], dpt[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i) { st[i][0] = st[i][1] = st[i][2] = st[i][3] = 0.f; dpt[i][0] = dpt[i][1] = dpt[i][2] = dpt[i][3] = 0.f; }
      mma_nmajor<DQK>(st, ak, Qs, c0, lane);        // S^T tile: rows = keys, cols = queries
      mma_nmajor<DV>(dpt, av, dOs, c0, lane);       // dP^T tile
      // Padded query columns need no masks: their dO rows are zero (dV += P^T dO, dP = 0) and D = 0, so dS = 0;
      // their lse is 0 and their Q row is zero, so P = 1 stays finite.
      float pt[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int qc = c0 + nt * 8 + 2 * t;
        const float2 L = *reinterpret_cast<const float2*>(Ls + qc), Dd = *reinterpret_cast<const float2*>(Ds + qc);
        pt[nt][0] = ex2(fmaf(st[nt][0], sl, -L.x)); pt[nt][1] = ex2(fmaf(st[nt][1], sl, -L.y));
        pt[nt][2] = ex2(fmaf(st[nt][2], sl, -L.x)); pt[nt][3] = ex2(fmaf(st[nt][3], sl, -L.y));
        st[nt][0] = pt[nt][0] * (dpt[nt][0] - Dd.x); st[nt][1] = pt[nt][1] * (dpt[nt][1] - Dd.y);
        st[nt][2] = pt[nt][2] * (dpt[nt][2] - Dd.x); st[nt][3] = pt[nt][3] * (dpt[nt][3] - Dd.y);
      }
      uint32_t fa[4][4];
      c_to_a(fa, pt);
      mma_kmajor<DV>(dv, fa, dOs, c0, lane);        // dV += P^T dO
      c_to_a(fa, st);
      mma_kmajor<DQK>(dk, fa, Qs, c0, lane);        // dK += dS^T Q
    }
    const int r0 = kt * 16 + g, r1 = r0 + 8;
    uint16_t* kb = a.dk + (int64_t)b * a.dk_bs + h * DQK;
    uint16_t* vb = a.dv_ + (int64_t)b * a.dv_bs + h * DV;
#pragma unroll
    for (int nt = 0; nt < DQK / 8; ++nt) {
      const int col = nt * 8 + 2 * t;
      if (r0 < Nk) *reinterpret_cast<uint32_t*>(kb + (int64_t)r0 * a.dk_rs + col) = pack_bf16x2(dk[nt][0] * a.scale, dk[nt][1] * a.scale);
      if (r1 < Nk) *reinterpret_cast<uint32_t*>(kb + (int64_t)r1 * a.dk_rs + col) = pack_bf16x2(dk[nt][2] * a.scale, dk[nt][3] * a.scale);
    }
#pragma unroll
    for (int nt = 0; nt < DV / 8; ++nt) {
      const int col = nt * 8 + 2 * t;
      if (r0 < Nk) *reinterpret_cast<uint32_t*>(vb + (int64_t)r0 * a.dv_rs + col) = pack_bf16x2(dv[nt][0], dv[nt][1]);
      if (r1 < Nk) *reinterpret_cast<uint32_t*>(vb + (int64_t)r1 * a.dv_rs + col) = pack_bf16x2(dv[nt][2], dv[nt][3]);
    }
  }
}

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

template <int DQK, int DV, int NW>
static int launch_fwd_nw(const davf_attn_fwd_args& a, cudaStream_t st) {
  const int Nkp = round_up(a.Nk, CH);
  constexpr int QT = NW * 16;
  const size_t smem = (size_t)(QT + Nkp) * (DQK + 8) * 2 + (size_t)Nkp * (DV + 8) * 2;
  auto kern = attn_mma_fwd_kernel<DQK, DV, NW>;
  static size_t configured = 0;
  if (smem > configured) {
    DAVF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    DAVF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
    configured = smem;
  }
  dim3 grid((a.Nq + QT - 1) / QT, a.B * a.H);
  DAVF_CUDA(launch_pdl(kern, grid, dim3(NW * 32), smem, st, a, Nkp));
  g_launch_kind[kKindAttnMma].fetch_add(1);
  DAVF_LAUNCH_OK();
  return DAVF_OK;
}
template <int DQK, int DV>
static int launch_fwd(const davf_attn_fwd_args& a, cudaStream_t st) {
  // fusion-token attentions (8 / 16 queries): one or two warps compute; with >= 32 keys FOUR warps stage the head's K / V into
  // shared memory (the extra warps leave after the load: with one warp the 12 KB of a 49-key head were 25 dependent 16-byte
  // loads per lane; measured 14.0 -> 10.9 us at 49 keys, while at 19 keys the wider CTA loses: 7.6 -> 8.7 us)
  if (a.Nq <= 32 && a.Nk >= 32) return launch_fwd_nw<DQK, DV, 4>(a, st);
  if (a.Nq <= 16) return launch_fwd_nw<DQK, DV, 1>(a, st);
  if (a.Nq <= 32) return launch_fwd_nw<DQK, DV, 2>(a, st);
  // long sequences (decoders: 228 / 128 queries): more query rows per CTA, so the head's K / V tile is staged
  // into shared memory by fewer CTAs (DAVF_ATTN_FWD_NW = 4 / 8 / 16 overrides, for experiments)
  static const int force = [] { const char* e = getenv("DAVF_ATTN_FWD_NW"); return e ? atoi(e) : 0; }();
  const int nw = force ? force : (a.Nq > 64 ? 8 : 4);        // measured: decoder image 104 -> 64 us, decoder audio 30.5 -> 25.1 us
  if (nw == 16) return launch_fwd_nw<DQK, DV, 16>(a, st);
  if (nw == 8) return launch_fwd_nw<DQK, DV, 8>(a, st);
  return launch_fwd_nw<DQK, DV, 4>(a, st);
}

template <int DQK, int DV, int NW>
static int launch_bwd_nw(const davf_attn_bwd_args& a, cudaStream_t st) {
  const int Nqp = round_up(a.Nq, CH), Nkp = round_up(a.Nk, CH);
  const size_t smem = (size_t)(Nqp + Nkp) * (DQK + 8) * 2 + (size_t)(Nqp + Nkp) * (DV + 8) * 2 + (size_t)2 * Nqp * 4;
  auto kern = attn_mma_bwd_kernel<DQK, DV, NW>;
  static size_t configured = 0;
  if (smem > configured) {
    DAVF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // without this the driver picks a carve-out that fits ONE 84 KB decoder CTA per SM (ncu: occupancy limit 1)
    DAVF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
    configured = smem;
  }
  DAVF_CUDA(launch_pdl(kern, dim3(a.B * a.H), dim3(NW * 32), smem, st, a, Nqp, Nkp));
  g_launch_kind[kKindAttnMma].fetch_add(1);
  DAVF_LAUNCH_OK();
  return DAVF_OK;
}
template <int DQK, int DV>
static int launch_bwd(const davf_attn_bwd_args& a, cudaStream_t st) {
  const int n = a.Nq > a.Nk ? a.Nq : a.Nk;
  if (n <= 32) return launch_bwd_nw<DQK, DV, 2>(a, st);         // one or two 16-row tiles per phase
  if (n <= 128) return launch_bwd_nw<DQK, DV, 4>(a, st);
  return launch_bwd_nw<DQK, DV, 8>(a, st);
}

// tcgen05 / TMEM / TMA kernels (attention_tc.cu): every problem with at least 8 query rows and head dim 64 / 32
bool attn_tc_fwd_ok(const davf_attn_fwd_args& a);
bool attn_tc_bwd_ok(const davf_attn_bwd_args& a);
int attn_tc_fwd(const davf_attn_fwd_args& a, cudaStream_t st);
int attn_tc_bwd(const davf_attn_bwd_args& a, cudaStream_t st);
static std::atomic<int> g_attn_impl{0};

static bool s8(int64_t x) { return x % 8 == 0; }

// dq_dead_rows for the kernels that do not zero-fill themselves: [B][dead][cols] bf16 pairs in front of dq
__global__ void attn_zero_dead_kernel(uint16_t* dq, int64_t bs, int64_t rs, int B, int dead, int cols) {
  pdl_launch_dependents();
  pdl_wait();
  const int per_row = cols / 2;
  const int64_t total = (int64_t)B * dead * per_row;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % per_row);
    const int64_t r = i / per_row;
    const int row = (int)(r % dead);
    const int64_t b = r / dead;
    *reinterpret_cast<uint32_t*>(dq + b * bs + (int64_t)(row - dead) * rs + 2 * c) = 0u;
  }
}
static int zero_dead_rows(const davf_attn_bwd_args& a, cudaStream_t st) {
  const int64_t total = (int64_t)a.B * a.dq_dead_rows * (a.H * a.dqk / 2);
  const int grid = (int)((total + 255) / 256 < 2048 ? (total + 255) / 256 : 2048);
  DAVF_CUDA(launch_pdl(attn_zero_dead_kernel, dim3(grid), dim3(256), 0, st, a.dq, a.dq_bs, a.dq_rs, a.B, a.dq_dead_rows, a.H * a.dqk));
  DAVF_LAUNCH_OK();
  return DAVF_OK;
}

}  // namespace davf

using namespace davf;

extern "C" int davf_set_attn_impl(int impl) {
  DAVF_CHECK_ARG(impl == 0 || impl == 2, "set_attn_impl: %d (0 = default dispatch, 2 = mma.sync everywhere)", impl);
  g_attn_impl.store(impl);
  return DAVF_OK;
}

extern "C" int davf_attention_fwd(const davf_attn_fwd_args* a, davf_stream_t s) {
  DAVF_CHECK_ARG(a && a->q && a->k && a->v && a->o, "attention_fwd: null pointer");
  DAVF_CHECK_ARG(a->Nk > 0 && a->Nk <= 256 && a->Nq > 0 && a->H > 0 && a->B >= 0, "attention_fwd: Nq=%d Nk=%d (Nk <= 256)", a->Nq, a->Nk);
  DAVF_CHECK_ARG(s8(a->q_rs) && s8(a->q_bs) && s8(a->k_rs) && s8(a->k_bs) && s8(a->v_rs) && s8(a->v_bs) && a->o_rs % 2 == 0 && a->o_bs % 2 == 0,
                 "attention_fwd: q/k/v strides must be multiples of 8 elements (16-byte rows)");
  DAVF_CHECK_ARG((((uintptr_t)a->q | (uintptr_t)a->k | (uintptr_t)a->v) & 15) == 0 && ((uintptr_t)a->o & 3) == 0, "attention_fwd: q/k/v must be 16-byte aligned");
  if (a->B == 0) return DAVF_OK;
  cudaStream_t st = as_stream(s);
  if (g_attn_impl.load() == 0 && attn_tc_fwd_ok(*a)) return attn_tc_fwd(*a, st);
  if (a->dqk == 64 && a->dv == 64) return launch_fwd<64, 64>(*a, st);
  if (a->dqk == 32 && a->dv == 32) return launch_fwd<32, 32>(*a, st);
  if (a->dqk == 16 && a->dv == 64) return launch_fwd<16, 64>(*a, st);
  set_error("attention_fwd: head dims (%d,%d) unsupported", a->dqk, a->dv);
  return DAVF_EUNSUPPORTED;
}

extern "C" int davf_attention_bwd(const davf_attn_bwd_args* a, davf_stream_t s) {
  DAVF_CHECK_ARG(a && a->q && a->k && a->v && a->d_o && a->lse && a->dq && a->dk && a->dv_, "attention_bwd: null pointer");
  DAVF_CHECK_ARG(a->Nk > 0 && a->Nk <= 256 && a->Nq > 0 && a->Nq <= 256, "attention_bwd: Nq=%d Nk=%d (<= 256)", a->Nq, a->Nk);
  DAVF_CHECK_ARG(s8(a->q_rs) && s8(a->q_bs) && s8(a->k_rs) && s8(a->k_bs) && s8(a->v_rs) && s8(a->v_bs) && s8(a->do_rs) && s8(a->do_bs),
                 "attention_bwd: q/k/v/dO strides must be multiples of 8 elements (16-byte rows)");
  DAVF_CHECK_ARG((((uintptr_t)a->q | (uintptr_t)a->k | (uintptr_t)a->v | (uintptr_t)a->d_o) & 15) == 0, "attention_bwd: q/k/v/dO must be 16-byte aligned");
  DAVF_CHECK_ARG(a->dq_rs % 2 == 0 && a->dk_rs % 2 == 0 && a->dv_rs % 2 == 0 && a->dq_bs % 2 == 0 && a->dk_bs % 2 == 0 && a->dv_bs % 2 == 0,
                 "attention_bwd: gradient strides must be even");
  DAVF_CHECK_ARG(a->dq_dead_rows >= 0 && a->dq_dead_rows <= 4096, "attention_bwd: dq_dead_rows=%d", a->dq_dead_rows);
  if (a->B == 0) return DAVF_OK;
  cudaStream_t st = as_stream(s);
  if (g_attn_impl.load() == 0 && attn_tc_bwd_ok(*a)) return attn_tc_bwd(*a, st);     // (zero-fills the dead rows itself)
  if (a->dq_dead_rows > 0)
    if (int rc = zero_dead_rows(*a, st)) return rc;
  if (a->dqk == 64 && a->dv == 64) return launch_bwd<64, 64>(*a, st);
  if (a->dqk == 32 && a->dv == 32) return launch_bwd<32, 32>(*a, st);
  if (a->dqk == 16 && a->dv == 64) return launch_bwd<16, 64>(*a, st);
  set_error("attention_bwd: head dims (%d,%d) unsupported", a->dqk, a->dv);
  return DAVF_EUNSUPPORTED;
}
