// K6/K8: fused self-attention on the 5th-generation tensor cores, forward and backward.
//
// Replaces timm Attention's SDPA call (vits.py:32-34, avmae.py:53-55,83-85) for every (batch, head) problem with more
// than 16 query rows: the modality encoder blocks (49 / 19 live queries against 81 / 51 keys, or 196 / 96 against
// 228 / 128 unmasked; head dim 64) and the MAE decoders (228 / 128 tokens, head dim 32).
//
// Data flow (no score matrix, no transposed copy ever touches HBM):
//   TMA (cp.async.bulk.tensor.3d) reads the head's Q / K / V (/ dO) rows straight out of the packed [B, S, 3, H, d]
//   qkv buffer into 128B-swizzled shared memory -- [rows][64 bf16] tiles; out-of-range rows are zero-filled by the
//   TMA unit, so ragged sequence lengths need no masking of the operands.  tcgen05.mma accumulates S = Q K^T (forward)
//   or S^T = K Q^T and dP^T = V dO^T (backward) in TENSOR MEMORY; the softmax warps read them with tcgen05.ld (one row
//   per lane: row max / row sum are plain register reductions, no shuffles), write P (forward) or P^T and dS^T
//   (backward) as bf16 into swizzled shared memory, and a second round of tcgen05.mma produces O = P V or
//   dV = P^T dO, dK = dS^T Q, dQ = dS K, again in tensor memory.  Results leave through a swizzled staging tile and a
//   TMA store (which clips the ragged tail).  The same [key][query] shared-memory tile serves as K-major A operand
//   (dK) and as MN-major A operand (dQ): one write, two descriptors.
//
// Head dim 32: two heads share a 128-byte row of the packed buffer.  K-major operands select the head with a +64 B
// start-address offset inside the swizzle atom (as a K-loop step does); MN-major B operands are used at N = 64 (both
// heads' columns) and the epilogue keeps the 32 accumulator columns of its head.
//
// Warp roles (192 threads, one CTA per SM, persistent over work items):
//   warp 0  TMA producer      warp 1  TMEM allocator + MMA issuer      warps 2..5  softmax / epilogue (TMEM lane quadrant = warp % 4)
#include <cuda.h>
#include <mutex>
#include <unordered_map>
#include <stdlib.h>

#include "tc_ptx.cuh"

namespace davf {

constexpr float kLog2eTc = 1.4426950408889634f;
constexpr float kLn2Tc = 0.6931471805599453f;
constexpr uint32_t kTile = 16384;            // one [128 rows][128 B] swizzled tile

// ---------------------------------------------------------------------------------------------------------------
// rank-3 tensor maps: dims {columns, rows, batch}, element strides {1, rs, bs}
// ---------------------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiledA)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                     const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiledA attn_get_encode() {
  static PFN_encodeTiledA fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiledA>(p);
  });
  return fn;
}
struct Map3Key {
  const void* ptr; int64_t cols, rows, batch, rs, bs; int box_cols, box_rows;
  bool operator==(const Map3Key& o) const {
    return ptr == o.ptr && cols == o.cols && rows == o.rows && batch == o.batch && rs == o.rs && bs == o.bs && box_cols == o.box_cols && box_rows == o.box_rows;
  }
};
struct Map3Hash {
  size_t operator()(const Map3Key& k) const {
    size_t h = (size_t)k.ptr;
    for (int64_t v : {k.cols, k.rows, k.batch, k.rs, k.bs, (int64_t)k.box_cols, (int64_t)k.box_rows}) h = h * 1000003u ^ (size_t)v;
    return h;
  }
};
// box = {box_cols, box_rows, 1}; 128B swizzle when a box row is 128 bytes, none otherwise (64-byte rows of a d = 32 store)
static int get_map3(const void* ptr, int64_t cols, int64_t rows, int64_t batch, int64_t rs, int64_t bs, int box_cols, int box_rows, CUtensorMap* out) {
  static std::unordered_map<Map3Key, CUtensorMap, Map3Hash> cache;
  static std::mutex mu;
  if (batch <= 1) bs = rows * rs;                 // a single sample: any legal stride
  Map3Key key{ptr, cols, rows, batch, rs, bs, box_cols, box_rows};
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return DAVF_OK; }
  }
  PFN_encodeTiledA enc = attn_get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return DAVF_ECUDA; }
  cuuint64_t gdim[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)(batch < 1 ? 1 : batch)};
  cuuint64_t gstride[2] = {(cuuint64_t)rs * 2, (cuuint64_t)bs * 2};
  cuuint32_t box[3] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, 1u};
  cuuint32_t estr[3] = {1u, 1u, 1u};
  CUtensorMap m;
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   box_cols * 2 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("attention: cuTensorMapEncodeTiled failed (%d) ptr=%p cols=%lld rows=%lld batch=%lld rs=%lld bs=%lld box=%dx%d", (int)r, ptr,
              (long long)cols, (long long)rows, (long long)batch, (long long)rs, (long long)bs, box_cols, box_rows);
    return DAVF_ECUDA;
  }
  {
    std::lock_guard<std::mutex> lk(mu);
    if (cache.size() > 65536) cache.clear();
    cache[key] = m;
  }
  *out = m;
  return DAVF_OK;
}

__device__ __forceinline__ float ex2_tc(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ---------------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------------
struct AttnTcFwdParams {
  CUtensorMap tq, tk, tv, to;
  float* lse;
  int B, H, Nq, Nk;
  float scale;
};

// D = head dim (64 or 32), NKP = key capacity of the shared-memory tiles (128 or 256)
template <int D, int NKP>
__global__ void __launch_bounds__(192, 1) attn_tc_fwd_kernel(const __grid_constant__ AttnTcFwdParams p) {
  constexpr uint32_t KV_BYTES = NKP * 128;
  constexpr uint32_t STAGE_BYTES = kTile + 2 * KV_BYTES;          // Q tile | K | V
  constexpr uint32_t P_BYTES = (NKP / 64) * kTile;
  constexpr uint32_t TMEM_COLS = NKP == 256 ? 512 : 256;
  constexpr uint32_t O_COL = NKP;                                   // S: columns [0, NKP), O: [NKP, NKP + 64)
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t p_base = smem_base + 2 * STAGE_BYTES;
  const uint32_t bar_base = p_base + P_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 16u + 8u * s; };
  const uint32_t s_ready = bar_base + 32u, p_ready = bar_base + 40u, o_ready = bar_base + 48u, tmem_ptr_addr = bar_base + 56u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.tq)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.tk)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.tv)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.to)) : "memory");
    for (int s = 0; s < 2; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(s_ready, 1);
    mbar_init(p_ready, 4);
    mbar_init(o_ready, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr) : "memory");

  const int nqt = (p.Nq + 127) >> 7;
  const int items = p.B * p.H * nqt;
  const int nk16 = (p.Nk + 15) & ~15;
  auto decode = [&](int w, int& b, int& h, int& qt) { qt = w % nqt; const int bh = w / nqt; h = bh % p.H; b = bh / p.H; };
  auto col_of = [&](int h) { return D == 64 ? h * 64 : (h >> 1) * 64; };        // first column of the 128-byte row segment holding head h
  auto koff_of = [&](int h) { return D == 64 ? 0u : (uint32_t)(h & 1) * 64u; };   // byte offset of head h inside it

  if (warp == 0) {
    // ===================== TMA producer =====================
    const bool leader = elect_one();
    int it = 0;
    for (int w = blockIdx.x; w < items; w += gridDim.x, ++it) {
      int b, h, qt;
      decode(w, b, h, qt);
      const int s = it & 1;
      mbar_wait(empty_bar(s), ((it >> 1) & 1) ^ 1u);
      if (leader) {
        const uint32_t qs = smem_base + s * STAGE_BYTES, ks = qs + kTile, vs = ks + KV_BYTES;
        mbar_expect_tx(full_bar(s), STAGE_BYTES);
        tma_load_3d(qs, &p.tq, full_bar(s), col_of(h), qt * 128, b);
        tma_load_3d(ks, &p.tk, full_bar(s), col_of(h), 0, b);
        tma_load_3d(vs, &p.tv, full_bar(s), col_of(h), 0, b);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const bool leader = elect_one();
    const uint32_t idesc_s = make_idesc(128, nk16, 0, 0);        // S = Q K^T : both K-major
    const uint32_t idesc_o = make_idesc(128, 64, 0, 1);          // O = P V   : A K-major (keys), B = V MN-major
    int it = 0;
    for (int w = blockIdx.x; w < items; w += gridDim.x, ++it) {
      int b, h, qt;
      decode(w, b, h, qt);
      const int s = it & 1;
      const uint32_t qs = smem_base + s * STAGE_BYTES, ks = qs + kTile, vs = ks + KV_BYTES;
      const uint32_t koff = koff_of(h);
      mbar_wait(full_bar(s), (it >> 1) & 1);
      tc_fence_after();
      // the S columns are free: the softmax warps signalled p_ready of the previous item after their last read
      if (leader) {
#pragma unroll
        for (int k = 0; k < D / 16; ++k)
          umma_f16(tmem_base, make_smem_desc(qs + koff + k * 32, 16, 1024), make_smem_desc(ks + koff + k * 32, 16, 1024), idesc_s, k > 0 ? 1u : 0u);
        umma_commit(s_ready);
      }
      __syncwarp();
      mbar_wait(p_ready, it & 1);
      tc_fence_after();
      if (leader) {
        for (int kk = 0; kk < nk16 / 16; ++kk)
          umma_f16(tmem_base + O_COL, make_smem_desc(p_base + (kk >> 2) * kTile + (kk & 3) * 32, 16, 1024),
                   make_smem_desc(vs + kk * 2048, 8192, 1024), idesc_o, kk > 0 ? 1u : 0u);
        umma_commit(o_ready);
      }
      __syncwarp();
    }
  } else {
    // ===================== softmax + epilogue (lane = query row) =====================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16);
    const float sl = p.scale * kLog2eTc;
    const uint32_t sw = (uint32_t)(row & 7);
    int it = 0;
    for (int w = blockIdx.x; w < items; w += gridDim.x, ++it) {
      int b, h, qt;
      decode(w, b, h, qt);
      const int s = it & 1;
      const bool live = qt * 128 + quad * 32 < p.Nq;               // warp-uniform: this warp owns at least one real query
      mbar_wait(s_ready, it & 1);
      tc_fence_after();
      float mx = -INFINITY, l = 0.f;
      if (live) {
        for (int c = 0; c < nk16; c += 32) {                       // pass 1: row max
          uint32_t v[32];
          tmem_ld_32x32b_x32_issue(trow + c, v);
          tmem_ld_wait(v);
          if (c + 32 <= p.Nk) {
#pragma unroll
            for (int e = 0; e < 32; ++e) mx = fmaxf(mx, __uint_as_float(v[e]));
          } else {
#pragma unroll
            for (int e = 0; e < 32; ++e) mx = (c + e < p.Nk) ? fmaxf(mx, __uint_as_float(v[e])) : mx;
          }
        }
        const float m2 = mx * sl;
        for (int c = 0; c < nk16; c += 32) {                       // pass 2: p = 2^(s sl - m2), row sum, bf16 P tile
          uint32_t v[32];
          tmem_ld_32x32b_x32_issue(trow + c, v);
          tmem_ld_wait(v);
          float pr[32];
          if (c + 32 <= p.Nk) {
#pragma unroll
            for (int e = 0; e < 32; ++e) pr[e] = ex2_tc(fmaf(__uint_as_float(v[e]), sl, -m2));
          } else {
#pragma unroll
            for (int e = 0; e < 32; ++e) pr[e] = (c + e < p.Nk) ? ex2_tc(fmaf(__uint_as_float(v[e]), sl, -m2)) : 0.f;
          }
#pragma unroll
          for (int e = 0; e < 32; ++e) l += pr[e];
          const uint32_t chunk = p_base + (uint32_t)(c >> 6) * kTile + (uint32_t)row * 128u;
          const uint32_t j0 = (uint32_t)((c & 63) >> 3);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            sts_128(chunk + (((j0 + j) ^ sw) << 4), pack_bf16x2(pr[8 * j], pr[8 * j + 1]), pack_bf16x2(pr[8 * j + 2], pr[8 * j + 3]),
                    pack_bf16x2(pr[8 * j + 4], pr[8 * j + 5]), pack_bf16x2(pr[8 * j + 6], pr[8 * j + 7]));
        }
      }
      fence_async_smem();                                          // generic-proxy P writes -> visible to tcgen05.mma
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready);
      const int q = qt * 128 + row;
      if (live && q < p.Nq && p.lse) p.lse[((int64_t)b * p.H + h) * p.Nq + q] = (mx * sl + log2f(l)) * kLn2Tc;
      const float inv_l = live ? 1.0f / l : 0.f;

      mbar_wait(o_ready, it & 1);
      tc_fence_after();
      // every MMA of this item has retired: the stage's Q tile is dead and becomes the output staging tile
      const uint32_t stg = smem_base + s * STAGE_BYTES;
      if (live) {
        if (D == 64) {
          uint32_t o0[32], o1[32];
          tmem_ld_32x32b_x32_issue(trow + O_COL, o0);
          tmem_ld_32x32b_x32_issue(trow + O_COL + 32, o1);
          tmem_ld_wait(o0);
          tmem_ld_wait(o1);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t* z = j < 4 ? o0 + 8 * j : o1 + 8 * (j - 4);
            sts_128(stg + (uint32_t)row * 128u + (((uint32_t)j ^ sw) << 4),
                    pack_bf16x2(__uint_as_float(z[0]) * inv_l, __uint_as_float(z[1]) * inv_l), pack_bf16x2(__uint_as_float(z[2]) * inv_l, __uint_as_float(z[3]) * inv_l),
                    pack_bf16x2(__uint_as_float(z[4]) * inv_l, __uint_as_float(z[5]) * inv_l), pack_bf16x2(__uint_as_float(z[6]) * inv_l, __uint_as_float(z[7]) * inv_l));
          }
        } else {
          uint32_t o0[32];
          tmem_ld_32x32b_x32_issue(trow + O_COL + (uint32_t)(h & 1) * 32u, o0);     // this head's half of the N = 64 accumulator
          tmem_ld_wait(o0);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t* z = o0 + 8 * j;
            sts_128(stg + (uint32_t)row * 64u + ((uint32_t)j << 4),
                    pack_bf16x2(__uint_as_float(z[0]) * inv_l, __uint_as_float(z[1]) * inv_l), pack_bf16x2(__uint_as_float(z[2]) * inv_l, __uint_as_float(z[3]) * inv_l),
                    pack_bf16x2(__uint_as_float(z[4]) * inv_l, __uint_as_float(z[5]) * inv_l), pack_bf16x2(__uint_as_float(z[6]) * inv_l, __uint_as_float(z[7]) * inv_l));
          }
        }
      }
      fence_async_smem();
      tc_fence_before();
      named_bar_sync(1, 128);
      if (warp == 2 && lane == 0) {
        tma_store_3d(&p.to, stg, h * D, qt * 128, b);              // rows >= Nq are clipped by the TMA unit
        bulk_commit();
        bulk_wait_read<0>();
        mbar_arrive(empty_bar(s));                                  // Q / K / V of this stage may be overwritten
      }
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------------------------
struct AttnTcBwdParams {
  CUtensorMap tq, tk, tv, tdo;          // loads: box {64, NP, 1}
  CUtensorMap tdq, tdk, tdv;            // stores: box {D, 128, 1}
  const float* lse;
  const uint16_t* o; int64_t o_bs, o_rs;
  int B, H, Nq, Nk;
  float scale;
};

// Work item = one (batch, head).  Key tiles of 128 rows (MMA M), query tiles of up to 128 columns (MMA N).
//   per (kt, qt):  S^T = K Q^T, dP^T = V dO^T  ->  TMEM           (warp 1)
//                  P^T = 2^(S^T sl - L), dS^T = P^T (dP^T - D)  -> smem bf16   (warps 2..5, lane = key row)
//                  dV[kt] += P^T dO, dK[kt] += dS^T Q, dQ[qt] += dS K         (warp 1; accumulators stay in TMEM)
//   TMEM columns: S^T [0,128) | dP^T [128,256) | dV [256,320) | dK [320,384) | dQ tile 0 [384,448) | dQ tile 1 [448,512)
template <int D, int NP>
__global__ void __launch_bounds__(192, 1) attn_tc_bwd_kernel(const __grid_constant__ AttnTcBwdParams p) {
  constexpr uint32_t OP_BYTES = NP * 128;                          // one operand tile set: Q, K, V or dO
  constexpr uint32_t TMEM_COLS = 512;
  constexpr uint32_t C_ST = 0, C_DP = 128, C_DV = 256, C_DK = 320, C_DQ = 384;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t qs = smem_base, ks = qs + OP_BYTES, vs = ks + OP_BYTES, dos = vs + OP_BYTES;
  const uint32_t pt_base = dos + OP_BYTES;                         // P^T  [128 keys][128 queries] bf16: 2 chunks of 64 queries
  const uint32_t dst_base = pt_base + 2 * kTile;                   // dS^T, same layout
  const uint32_t ls_base = dst_base + 2 * kTile;                   // L[q] = lse * log2(e)   (f32, NP)
  const uint32_t ds_base = ls_base + NP * 4;                       // D[q] = dO_q . O_q       (f32, NP)
  const uint32_t bar_base = ds_base + NP * 4;
  const uint32_t full_bar = bar_base, empty_bar = bar_base + 8u, s_ready = bar_base + 16u, p_ready = bar_base + 24u,
                 acc_ready = bar_base + 32u, tmem_ptr_addr = bar_base + 40u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    const CUtensorMap* maps[7] = {&p.tq, &p.tk, &p.tv, &p.tdo, &p.tdq, &p.tdk, &p.tdv};
    for (int i = 0; i < 7; ++i) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(maps[i])) : "memory");
    mbar_init(full_bar, 1);
    mbar_init(empty_bar, 1);
    mbar_init(s_ready, 1);
    mbar_init(p_ready, 4);
    mbar_init(acc_ready, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr) : "memory");

  const int items = p.B * p.H;
  const int nqt = (p.Nq + 127) >> 7, nkt = (p.Nk + 127) >> 7;
  const int npairs = nqt * nkt;                                    // (kt, qt) pairs per item, qt fastest
  auto col_of = [&](int h) { return D == 64 ? h * 64 : (h >> 1) * 64; };
  auto koff_of = [&](int h) { return D == 64 ? 0u : (uint32_t)(h & 1) * 64u; };
  auto n16 = [](int n) { return (n + 15) & ~15; };

  if (warp == 0) {
    // ===================== TMA producer (single stage: the next item's tiles follow the last MMA of this one) =====================
    const bool leader = elect_one();
    int it = 0;
    for (int w = blockIdx.x; w < items; w += gridDim.x, ++it) {
      const int h = w % p.H, b = w / p.H;
      mbar_wait(empty_bar, (it & 1) ^ 1u);
      if (leader) {
        mbar_expect_tx(full_bar, 4 * OP_BYTES);
        tma_load_3d(qs, &p.tq, full_bar, col_of(h), 0, b);
        tma_load_3d(ks, &p.tk, full_bar, col_of(h), 0, b);
        tma_load_3d(vs, &p.tv, full_bar, col_of(h), 0, b);
        tma_load_3d(dos, &p.tdo, full_bar, col_of(h), 0, b);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const bool leader = elect_one();
    const uint32_t idesc_acc = make_idesc(128, 64, 0, 1);          // dV / dK: A = P^T / dS^T K-major (queries), B = dO / Q MN-major
    const uint32_t idesc_dq = make_idesc(128, 64, 1, 1);           // dQ: A = dS^T read MN-major (M = queries), B = K MN-major
    int it = 0;
    uint32_t pair_phase = 0;                                       // s_ready / p_ready complete once per (kt, qt) pair
    for (int w = blockIdx.x; w < items; w += gridDim.x, ++it) {
      const int h = w % p.H;
      const uint32_t koff = koff_of(h);
      mbar_wait(full_bar, it & 1);
      tc_fence_after();
      auto issue_scores = [&](int kt, int qt) {                    // S^T and dP^T of one pair
        const uint32_t idesc_s = make_idesc(128, n16(min(128, p.Nq - qt * 128)), 0, 0);
#pragma unroll
        for (int k = 0; k < D / 16; ++k)
          umma_f16(tmem_base + C_ST, make_smem_desc(ks + kt * kTile + koff + k * 32, 16, 1024),
                   make_smem_desc(qs + qt * kTile + koff + k * 32, 16, 1024), idesc_s, k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < D / 16; ++k)
          umma_f16(tmem_base + C_DP, make_smem_desc(vs + kt * kTile + koff + k * 32, 16, 1024),
                   make_smem_desc(dos + qt * kTile + koff + k * 32, 16, 1024), idesc_s, k > 0 ? 1u : 0u);
      };
      if (leader) { issue_scores(0, 0); umma_commit(s_ready); }
      __syncwarp();
      for (int pr = 0; pr < npairs; ++pr) {
        const int kt = pr / nqt, qt = pr % nqt;
        mbar_wait(p_ready, pair_phase);
        pair_phase ^= 1u;
        tc_fence_after();
        if (leader) {
          const int nq_steps = n16(min(128, p.Nq - qt * 128)) / 16;      // K extent of the dV / dK products (queries of this tile)
          const int nk_steps = n16(min(128, p.Nk - kt * 128)) / 16;      // K extent of the dQ product (keys of this tile)
          for (int kk = 0; kk < nq_steps; ++kk) {
            const uint32_t aoff = (uint32_t)(kk >> 2) * kTile + (uint32_t)(kk & 3) * 32u;
            const uint32_t boff = (uint32_t)(qt * 128 + kk * 16) * 128u;
            const uint32_t acc = (qt > 0 || kk > 0) ? 1u : 0u;
            umma_f16(tmem_base + C_DV, make_smem_desc(pt_base + aoff, 16, 1024), make_smem_desc(dos + boff, 8192, 1024), idesc_acc, acc);
            umma_f16(tmem_base + C_DK, make_smem_desc(dst_base + aoff, 16, 1024), make_smem_desc(qs + boff, 8192, 1024), idesc_acc, acc);
          }
          for (int kk = 0; kk < nk_steps; ++kk)
            umma_f16(tmem_base + C_DQ + qt * 64, make_smem_desc(dst_base + kk * 2048, kTile, 1024),
                     make_smem_desc(ks + (uint32_t)(kt * 128 + kk * 16) * 128u, 8192, 1024), idesc_dq, (kt > 0 || kk > 0) ? 1u : 0u);
          if (qt == nqt - 1) umma_commit(acc_ready);               // dV / dK of this key tile complete (last pair: dQ too)
          if (pr + 1 < npairs) { issue_scores((pr + 1) / nqt, (pr + 1) % nqt); umma_commit(s_ready); }
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== compute + epilogue warps (lane = key row of the current key tile) =====================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const int tid = row;                                            // 0..127 over the four warps
    const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16);
    const float sl = p.scale * kLog2eTc;
    const uint32_t sw = (uint32_t)(row & 7);
    int it = 0;
    uint32_t pair_phase = 0, acc_phase = 0;
    // one output tile: 64 accumulator columns of this lane's row -> (x mul) -> bf16 -> staging -> TMA store (clipped at `rows`)
    auto store_tile = [&](uint32_t tcol, uint32_t stg, const CUtensorMap* map, float mul, int h, int row0, int b) {
      if (D == 64) {
        uint32_t z0[32], z1[32];
        tmem_ld_32x32b_x32_issue(trow + tcol, z0);
        tmem_ld_32x32b_x32_issue(trow + tcol + 32, z1);
        tmem_ld_wait(z0);
        tmem_ld_wait(z1);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t* z = j < 4 ? z0 + 8 * j : z1 + 8 * (j - 4);
          sts_128(stg + (uint32_t)row * 128u + (((uint32_t)j ^ sw) << 4),
                  pack_bf16x2(__uint_as_float(z[0]) * mul, __uint_as_float(z[1]) * mul), pack_bf16x2(__uint_as_float(z[2]) * mul, __uint_as_float(z[3]) * mul),
                  pack_bf16x2(__uint_as_float(z[4]) * mul, __uint_as_float(z[5]) * mul), pack_bf16x2(__uint_as_float(z[6]) * mul, __uint_as_float(z[7]) * mul));
        }
      } else {
        uint32_t z0[32];
        tmem_ld_32x32b_x32_issue(trow + tcol + (uint32_t)(h & 1) * 32u, z0);
        tmem_ld_wait(z0);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t* z = z0 + 8 * j;
          sts_128(stg + (uint32_t)row * 64u + ((uint32_t)j << 4),
                  pack_bf16x2(__uint_as_float(z[0]) * mul, __uint_as_float(z[1]) * mul), pack_bf16x2(__uint_as_float(z[2]) * mul, __uint_as_float(z[3]) * mul),
                  pack_bf16x2(__uint_as_float(z[4]) * mul, __uint_as_float(z[5]) * mul), pack_bf16x2(__uint_as_float(z[6]) * mul, __uint_as_float(z[7]) * mul));
        }
      }
      fence_async_smem();
      tc_fence_before();
      named_bar_sync(1, 128);
      if (warp == 2 && lane == 0) {
        tma_store_3d(map, stg, h * D, row0, b);
        bulk_commit();
      }
    };
    for (int w = blockIdx.x; w < items; w += gridDim.x, ++it) {
      const int h = w % p.H, b = w / p.H;
      const uint32_t koff = koff_of(h);
      mbar_wait(full_bar, it & 1);
      // ---- L[q] and D[q] = dO_q . O_q for every query of the head (thread per query row) ----
      for (int q = tid; q < NP; q += 128) {
        float Lq = 1e30f, Dq = 0.f;                                  // padded queries: P = 2^(-inf) = 0, dS = 0
        if (q < p.Nq) {
          Lq = p.lse[((int64_t)b * p.H + h) * p.Nq + q] * kLog2eTc;
          const uint16_t* orow = p.o + (int64_t)b * p.o_bs + (int64_t)q * p.o_rs + (int64_t)h * D;
          const uint32_t drow = dos + (uint32_t)q * 128u;
          const uint32_t qsw = (uint32_t)(q & 7);
#pragma unroll
          for (int j = 0; j < D / 8; ++j) {
            const uint4 xo = *reinterpret_cast<const uint4*>(orow + 8 * j);
            const uint4 yo = lds_128(drow + ((((koff >> 4) + (uint32_t)j) ^ qsw) << 4));
            const uint32_t xs[4] = {xo.x, xo.y, xo.z, xo.w}, ys[4] = {yo.x, yo.y, yo.z, yo.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 x = unpack_bf16x2(xs[e]), y = unpack_bf16x2(ys[e]);
              Dq = fmaf(x.x, y.x, fmaf(x.y, y.y, Dq));
            }
          }
        }
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(ls_base + 4u * q), "f"(Lq) : "memory");
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(ds_base + 4u * q), "f"(Dq) : "memory");
      }
      named_bar_sync(1, 128);
      for (int pr = 0; pr < npairs; ++pr) {
        const int kt = pr / nqt, qt = pr % nqt;
        const int nq_t = n16(min(128, p.Nq - qt * 128));
        const bool key_ok = kt * 128 + row < p.Nk;
        mbar_wait(s_ready, pair_phase);
        pair_phase ^= 1u;
        tc_fence_after();
        // s_ready also tells that the accumulate MMAs of the previous pair have retired: P^T / dS^T may be overwritten
        for (int c = 0; c < nq_t; c += 32) {
          uint32_t sv[32], dv[32];
          tmem_ld_32x32b_x32_issue(trow + C_ST + c, sv);
          tmem_ld_32x32b_x32_issue(trow + C_DP + c, dv);
          tmem_ld_wait(sv);
          tmem_ld_wait(dv);
          const uint32_t lq = ls_base + 4u * (uint32_t)(qt * 128 + c), dq_ = ds_base + 4u * (uint32_t)(qt * 128 + c);
          const uint32_t chunk = (uint32_t)(c >> 6) * kTile + (uint32_t)row * 128u;
          const uint32_t j0 = (uint32_t)((c & 63) >> 3);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float pp[8], dd[8];
            const uint4 L0 = lds_128(lq + 32u * j), L1 = lds_128(lq + 32u * j + 16u);
            const uint4 D0 = lds_128(dq_ + 32u * j), D1 = lds_128(dq_ + 32u * j + 16u);
            const float Lv[8] = {__uint_as_float(L0.x), __uint_as_float(L0.y), __uint_as_float(L0.z), __uint_as_float(L0.w),
                                 __uint_as_float(L1.x), __uint_as_float(L1.y), __uint_as_float(L1.z), __uint_as_float(L1.w)};
            const float Dv[8] = {__uint_as_float(D0.x), __uint_as_float(D0.y), __uint_as_float(D0.z), __uint_as_float(D0.w),
                                 __uint_as_float(D1.x), __uint_as_float(D1.y), __uint_as_float(D1.z), __uint_as_float(D1.w)};
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float pe = key_ok ? ex2_tc(fmaf(__uint_as_float(sv[8 * j + e]), sl, -Lv[e])) : 0.f;
              pp[e] = pe;
              dd[e] = pe * (__uint_as_float(dv[8 * j + e]) - Dv[e]);          // dS without the softmax scale (applied to dQ / dK once)
            }
            const uint32_t pos = chunk + (((j0 + j) ^ sw) << 4);
            sts_128(pt_base + pos, pack_bf16x2(pp[0], pp[1]), pack_bf16x2(pp[2], pp[3]), pack_bf16x2(pp[4], pp[5]), pack_bf16x2(pp[6], pp[7]));
            sts_128(dst_base + pos, pack_bf16x2(dd[0], dd[1]), pack_bf16x2(dd[2], dd[3]), pack_bf16x2(dd[4], dd[5]), pack_bf16x2(dd[6], dd[7]));
          }
        }
        fence_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_ready);
        if (qt == nqt - 1) {
          // ---- dV / dK of this key tile (and, after the last one, dQ): every MMA issued so far has retired ----
          mbar_wait(acc_ready, acc_phase);
          acc_phase ^= 1u;
          tc_fence_after();
          // staging tiles alias P^T / dS^T: free (their readers have retired) until the next pair's s_ready
          store_tile(C_DV, pt_base, &p.tdv, 1.0f, h, kt * 128, b);
          store_tile(C_DK, dst_base, &p.tdk, p.scale, h, kt * 128, b);
          if (kt == nkt - 1) {
            for (int t = 0; t < nqt; ++t) store_tile(C_DQ + t * 64, (t == 0 ? pt_base : dst_base) + kTile, &p.tdq, p.scale, h, t * 128, b);
          }
          if (warp == 2 && lane == 0) bulk_wait_read<0>();          // the staging tiles are about to be overwritten
          tc_fence_before();
          named_bar_sync(1, 128);
          if (kt == nkt - 1 && warp == 2 && lane == 0) mbar_arrive(empty_bar);     // Q / K / V / dO may be overwritten
        }
      }
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
static bool al16(const void* p) { return ((uintptr_t)p & 15) == 0; }
static bool m8(int64_t x) { return x % 8 == 0; }

static int attn_tc_min_q() {
  static const int v = [] { const char* e = getenv("DAVF_ATTN_TC_MINQ"); return e ? atoi(e) : 17; }();
  return v;
}

bool attn_tc_fwd_ok(const davf_attn_fwd_args& a) {
  if (!((a.dqk == 64 && a.dv == 64) || (a.dqk == 32 && a.dv == 32))) return false;
  if (a.accumulate || a.Nq < attn_tc_min_q() || a.Nk > 256 || a.Nk < 1) return false;
  return al16(a.q) && al16(a.k) && al16(a.v) && al16(a.o) && m8(a.q_rs) && m8(a.q_bs) && m8(a.k_rs) && m8(a.k_bs) && m8(a.v_rs) && m8(a.v_bs) &&
         m8(a.o_rs) && m8(a.o_bs);
}

bool attn_tc_bwd_ok(const davf_attn_bwd_args& a) {
  if (!((a.dqk == 64 && a.dv == 64) || (a.dqk == 32 && a.dv == 32))) return false;
  if (a.accumulate_dq || !a.o || a.Nq < attn_tc_min_q() || a.Nq > 256 || a.Nk > 256 || a.Nk < 1) return false;
  return al16(a.q) && al16(a.k) && al16(a.v) && al16(a.d_o) && al16(a.o) && al16(a.dq) && al16(a.dk) && al16(a.dv_) &&
         m8(a.q_rs) && m8(a.q_bs) && m8(a.k_rs) && m8(a.k_bs) && m8(a.v_rs) && m8(a.v_bs) && m8(a.do_rs) && m8(a.do_bs) && m8(a.o_rs) && m8(a.o_bs) &&
         m8(a.dq_rs) && m8(a.dq_bs) && m8(a.dk_rs) && m8(a.dk_bs) && m8(a.dv_rs) && m8(a.dv_bs);
}

template <int D, int NKP>
static int launch_tc_fwd(const davf_attn_fwd_args& a, cudaStream_t st) {
  AttnTcFwdParams p;
  const int64_t cols = (int64_t)a.H * D;
  int rc;
  if ((rc = get_map3(a.q, cols, a.Nq, a.B, a.q_rs, a.q_bs, 64, 128, &p.tq))) return rc;
  if ((rc = get_map3(a.k, cols, a.Nk, a.B, a.k_rs, a.k_bs, 64, NKP, &p.tk))) return rc;
  if ((rc = get_map3(a.v, cols, a.Nk, a.B, a.v_rs, a.v_bs, 64, NKP, &p.tv))) return rc;
  if ((rc = get_map3(a.o, cols, a.Nq, a.B, a.o_rs, a.o_bs, D, 128, &p.to))) return rc;
  p.lse = a.lse; p.B = a.B; p.H = a.H; p.Nq = a.Nq; p.Nk = a.Nk; p.scale = a.scale;
  constexpr size_t smem = 2 * (size_t)(kTile + 2 * NKP * 128) + (size_t)(NKP / 64) * kTile + 64 + 1024;
  static_assert(smem <= 232448, "exceeds the 227 KB shared memory of an SM");
  auto kern = attn_tc_fwd_kernel<D, NKP>;
  static bool attr_set = false;
  if (!attr_set) {
    DAVF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  const int items = a.B * a.H * ((a.Nq + 127) / 128);
  const int grid = items < kNumSMs ? items : kNumSMs;
  kern<<<grid, 192, smem, st>>>(p);
  g_launch_kind[kKindAttnTc].fetch_add(1);
  DAVF_LAUNCH_OK();
  return DAVF_OK;
}

template <int D, int NP>
static int launch_tc_bwd(const davf_attn_bwd_args& a, cudaStream_t st) {
  AttnTcBwdParams p;
  const int64_t cols = (int64_t)a.H * D;
  int rc;
  if ((rc = get_map3(a.q, cols, a.Nq, a.B, a.q_rs, a.q_bs, 64, NP, &p.tq))) return rc;
  if ((rc = get_map3(a.k, cols, a.Nk, a.B, a.k_rs, a.k_bs, 64, NP, &p.tk))) return rc;
  if ((rc = get_map3(a.v, cols, a.Nk, a.B, a.v_rs, a.v_bs, 64, NP, &p.tv))) return rc;
  if ((rc = get_map3(a.d_o, cols, a.Nq, a.B, a.do_rs, a.do_bs, 64, NP, &p.tdo))) return rc;
  if ((rc = get_map3(a.dq, cols, a.Nq, a.B, a.dq_rs, a.dq_bs, D, 128, &p.tdq))) return rc;
  if ((rc = get_map3(a.dk, cols, a.Nk, a.B, a.dk_rs, a.dk_bs, D, 128, &p.tdk))) return rc;
  if ((rc = get_map3(a.dv_, cols, a.Nk, a.B, a.dv_rs, a.dv_bs, D, 128, &p.tdv))) return rc;
  p.lse = a.lse; p.o = a.o; p.o_bs = a.o_bs; p.o_rs = a.o_rs;
  p.B = a.B; p.H = a.H; p.Nq = a.Nq; p.Nk = a.Nk; p.scale = a.scale;
  constexpr size_t smem = 4 * (size_t)NP * 128 + 4 * (size_t)kTile + 2 * (size_t)NP * 4 + 64 + 1024;
  static_assert(smem <= 232448, "exceeds the 227 KB shared memory of an SM");
  auto kern = attn_tc_bwd_kernel<D, NP>;
  static bool attr_set = false;
  if (!attr_set) {
    DAVF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  const int items = a.B * a.H;
  const int grid = items < kNumSMs ? items : kNumSMs;
  kern<<<grid, 192, smem, st>>>(p);
  g_launch_kind[kKindAttnTc].fetch_add(1);
  DAVF_LAUNCH_OK();
  return DAVF_OK;
}

int attn_tc_fwd(const davf_attn_fwd_args& a, cudaStream_t st) {
  if (a.dqk == 64) return a.Nk <= 128 ? launch_tc_fwd<64, 128>(a, st) : launch_tc_fwd<64, 256>(a, st);
  return a.Nk <= 128 ? launch_tc_fwd<32, 128>(a, st) : launch_tc_fwd<32, 256>(a, st);
}

int attn_tc_bwd(const davf_attn_bwd_args& a, cudaStream_t st) {
  const int n = a.Nq > a.Nk ? a.Nq : a.Nk;
  if (a.dqk == 64) return n <= 128 ? launch_tc_bwd<64, 128>(a, st) : launch_tc_bwd<64, 256>(a, st);
  return n <= 128 ? launch_tc_bwd<32, 128>(a, st) : launch_tc_bwd<32, 256>(a, st);
}

}  // namespace davf
