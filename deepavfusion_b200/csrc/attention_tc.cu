// K6/K8: fused self-attention on the 5th-generation tensor cores, forward and backward.
//
// Replaces timm Attention's SDPA call (vits.py:32-34, avmae.py:53-55,83-85) and the fusion block's CrossAttention
// (fusion_blocks.py:46-59) for every (batch, head) problem with at least 8 query rows and head dim 64 / 32: the modality encoder blocks (49 / 19 live queries against 81 / 51 keys, or 196 / 96 against
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

// Work item = (batch, head, 128-query tile).  Every phase of an item hangs on the previous one (load -> S -> softmax ->
// P V -> store), so ONE item per SM leaves the tensor pipe, the MUFU and the TMA unit idle in turn.  The kernel is
// therefore sized for TWO resident CTAs per SM that interleave their phases:
//   TMEM   NKP columns per CTA: S = Q K^T in [0, NKP); O = P V overwrites columns [0, 64) once the softmax has consumed S
//   smem   Q | K | X | V.  P (bf16, NKP / 64 tiles of [128 q][64 keys]) is written over Q and K, which are dead once S
//          has been computed (NKP = 256: P also covers X); X is the output staging tile.  96 KB (NKP = 256) / 64 KB (128).
// D = head dim (64 or 32), NKP = key capacity of the shared-memory tiles (128 or 256)
template <int D, int NKP>
__global__ void __launch_bounds__(192, 2) attn_tc_fwd_kernel(const __grid_constant__ AttnTcFwdParams p) {
  constexpr uint32_t KV_BYTES = NKP * 128;
  constexpr uint32_t LOAD_BYTES = kTile + 2 * KV_BYTES;
  constexpr uint32_t TMEM_COLS = NKP;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t qs = smem_base, ks = qs + kTile, xs = ks + KV_BYTES, vs = xs + kTile;
  const uint32_t p_base = smem_base;                                // P tiles alias Q | K (| X)
  const uint32_t bar_base = vs + KV_BYTES;
  const uint32_t full_bar = bar_base, s_ready = bar_base + 8u, p_ready = bar_base + 16u, o_ready = bar_base + 24u,
                 o_cons = bar_base + 32u, stg_free = bar_base + 40u, tmem_ptr_addr = bar_base + 48u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.tq)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.tk)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.tv)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.to)) : "memory");
    mbar_init(full_bar, 1);
    mbar_init(s_ready, 1);
    mbar_init(p_ready, 4);
    mbar_init(o_ready, 1);
    mbar_init(o_cons, 4);
    mbar_init(stg_free, 1);
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
  pdl_launch_dependents();     // set-up done, nothing global touched yet (PDL, common.cuh)
  pdl_wait();

  const int nqt = (p.Nq + 127) >> 7;
  const int items = p.B * p.H * nqt;
  const int nk16 = (p.Nk + 15) & ~15;
  auto decode = [&](int w, int& b, int& h, int& qt) { qt = w % nqt; const int bh = w / nqt; h = bh % p.H; b = bh / p.H; };
  auto col_of = [&](int h) { return D == 64 ? h * 64 : (h >> 1) * 64; };        // first column of the 128-byte row segment holding head h
  auto koff_of = [&](int h) { return D == 64 ? 0u : (uint32_t)(h & 1) * 64u; };   // byte offset of head h inside it

  if (warp == 0) {
    // ===================== TMA producer: loads of the next item, store of the finished one =====================
    const bool leader = elect_one();
    auto load = [&](int w) {
      int b, h, qt;
      decode(w, b, h, qt);
      if (leader) {
        mbar_expect_tx(full_bar, LOAD_BYTES);
        tma_load_3d(qs, &p.tq, full_bar, col_of(h), qt * 128, b);
        tma_load_3d(ks, &p.tk, full_bar, col_of(h), 0, b);
        tma_load_3d(vs, &p.tv, full_bar, col_of(h), 0, b);
      }
    };
    if ((int)blockIdx.x < items) load(blockIdx.x);
    __syncwarp();
    int it = 0;
    for (int w = blockIdx.x; w < items; w += gridDim.x, ++it) {
      int b, h, qt;
      decode(w, b, h, qt);
      mbar_wait(o_ready, it & 1);                                  // every MMA of this item has retired: Q / K / V (and P) are dead
      if (w + (int)gridDim.x < items) load(w + gridDim.x);
      __syncwarp();
      mbar_wait(o_cons, it & 1);                                   // the staging tile is written
      if (leader) {
        tma_store_3d(&p.to, xs, h * D, qt * 128, b);               // rows >= Nq are clipped by the TMA unit
        bulk_commit();
        bulk_wait_read<0>();
        mbar_arrive(stg_free);
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
      const uint32_t koff = koff_of(h);
      mbar_wait(full_bar, it & 1);
      if (it > 0) mbar_wait(o_cons, (it - 1) & 1);               // O of the previous item has been read out of columns [0, 64)
      tc_fence_after();
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
          umma_f16(tmem_base, make_smem_desc(p_base + (kk >> 2) * kTile + (kk & 3) * 32, 16, 1024),
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
      const bool live = qt * 128 + quad * 32 < p.Nq;               // warp-uniform: this warp owns at least one real query
      mbar_wait(s_ready, it & 1);
      tc_fence_after();
      float mx = -INFINITY, l = 0.f, m2 = 0.f;
      const int nch = (nk16 + 31) >> 5;                            // 32-column chunks of the score row
      uint32_t va[32], vb[32];
      // Softmax is shift-invariant, and bf16 / f32 share one exponent range, so the shift only has to keep 2^(s sl - m2)
      // finite: ANY m2 within +-64 of the true row maximum gives the same P to the last bit of its mantissa.  TMEM reads
      // run at 64 B/clk/SM, i.e. a second pass over the 128 x Nk score tile costs as much as all its ex2 -- so the
      // shift is the maximum of the FIRST 32 columns (a lower bound of the row maximum, hence p >= 1 somewhere: no
      // underflow), the exact maximum is tracked on the fly, and only a row whose maximum exceeds the estimate by more
      // than 64 (p > 2^64; never seen on real activations) repeats the pass with the exact value.  LSE = m2 + log2(l)
      // holds for any shift.
      if (live) {
        tmem_ld_32x32b_x32_issue(trow, va);
        tmem_ld_wait(va);
#pragma unroll
        for (int e = 0; e < 32; ++e) mx = (e < p.Nk) ? fmaxf(mx, __uint_as_float(va[e])) : mx;
        m2 = mx * sl;
      }
      if (it > 0) mbar_wait(stg_free, (it - 1) & 1);               // the TMA store of the previous item has read X (P's last tile when NKP = 256)
      if (live) {
        // p = 2^(s sl - m2), row sum, bf16 P tile (Q and K are dead: S is complete)
        auto emit = [&](const uint32_t (&v)[32], int c) {
          float pr[32];
          if (c * 32 + 32 <= p.Nk) {
#pragma unroll
            for (int e = 0; e < 32; ++e) { mx = fmaxf(mx, __uint_as_float(v[e])); pr[e] = ex2_tc(fmaf(__uint_as_float(v[e]), sl, -m2)); }
          } else {
#pragma unroll
            for (int e = 0; e < 32; ++e) {
              const bool ok = c * 32 + e < p.Nk;
              mx = ok ? fmaxf(mx, __uint_as_float(v[e])) : mx;
              pr[e] = ok ? ex2_tc(fmaf(__uint_as_float(v[e]), sl, -m2)) : 0.f;
            }
          }
          float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
          for (int e = 0; e < 32; e += 4) { s0 += pr[e]; s1 += pr[e + 1]; s2 += pr[e + 2]; s3 += pr[e + 3]; }
          l += (s0 + s1) + (s2 + s3);
          const uint32_t chunk = p_base + (uint32_t)(c >> 1) * kTile + (uint32_t)row * 128u;
          const uint32_t j0 = (uint32_t)(c & 1) * 4u;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            sts_128(chunk + (((j0 + j) ^ sw) << 4), pack_bf16x2(pr[8 * j], pr[8 * j + 1]), pack_bf16x2(pr[8 * j + 2], pr[8 * j + 3]),
                    pack_bf16x2(pr[8 * j + 4], pr[8 * j + 5]), pack_bf16x2(pr[8 * j + 6], pr[8 * j + 7]));
        };
        auto sweep = [&](bool first_loaded) {                       // the TMEM load of chunk c + 1 is in flight while chunk c is processed
          if (!first_loaded) { tmem_ld_32x32b_x32_issue(trow, va); tmem_ld_wait(va); }
          for (int c = 0; c < nch; c += 2) {
            if (c + 1 < nch) tmem_ld_32x32b_x32_issue(trow + (c + 1) * 32, vb);
            emit(va, c);
            if (c + 1 < nch) {
              tmem_ld_wait(vb);
              if (c + 2 < nch) tmem_ld_32x32b_x32_issue(trow + (c + 2) * 32, va);
              emit(vb, c + 1);
              if (c + 2 < nch) tmem_ld_wait(va);
            }
          }
        };
        sweep(true);
        if (__any_sync(0xffffffffu, mx * sl - m2 > 64.0f)) {        // (warp-uniform) exact repeat: the row maximum is known now
          m2 = mx * sl;
          l = 0.f;
          sweep(false);
        }
      }
      fence_async_smem();                                          // generic-proxy P writes -> visible to tcgen05.mma
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready);
      const int q = qt * 128 + row;
      if (live && q < p.Nq && p.lse) p.lse[((int64_t)b * p.H + h) * p.Nq + q] = (m2 + log2f(l)) * kLn2Tc;
      const float inv_l = live ? 1.0f / l : 0.f;

      mbar_wait(o_ready, it & 1);
      tc_fence_after();
      if (live) {
        if (D == 64) {
          tmem_ld_32x32b_x32_issue(trow, va);
          tmem_ld_32x32b_x32_issue(trow + 32, vb);
          tmem_ld_wait(va);
          tmem_ld_wait(vb);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t* z = j < 4 ? va + 8 * j : vb + 8 * (j - 4);
            sts_128(xs + (uint32_t)row * 128u + (((uint32_t)j ^ sw) << 4),
                    pack_bf16x2(__uint_as_float(z[0]) * inv_l, __uint_as_float(z[1]) * inv_l), pack_bf16x2(__uint_as_float(z[2]) * inv_l, __uint_as_float(z[3]) * inv_l),
                    pack_bf16x2(__uint_as_float(z[4]) * inv_l, __uint_as_float(z[5]) * inv_l), pack_bf16x2(__uint_as_float(z[6]) * inv_l, __uint_as_float(z[7]) * inv_l));
          }
        } else {
          tmem_ld_32x32b_x32_issue(trow + (uint32_t)(h & 1) * 32u, va);        // this head's half of the N = 64 accumulator
          tmem_ld_wait(va);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t* z = va + 8 * j;
            sts_128(xs + (uint32_t)row * 64u + ((uint32_t)j << 4),
                    pack_bf16x2(__uint_as_float(z[0]) * inv_l, __uint_as_float(z[1]) * inv_l), pack_bf16x2(__uint_as_float(z[2]) * inv_l, __uint_as_float(z[3]) * inv_l),
                    pack_bf16x2(__uint_as_float(z[4]) * inv_l, __uint_as_float(z[5]) * inv_l), pack_bf16x2(__uint_as_float(z[6]) * inv_l, __uint_as_float(z[7]) * inv_l));
          }
        }
      }
      fence_async_smem();                                          // staging writes -> visible to the TMA store
      tc_fence_before();                                           // O reads complete -> the next S MMA may overwrite the columns
      __syncwarp();
      if (lane == 0) mbar_arrive(o_cons);
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
  CUtensorMap tdq0;                     // the dead query rows in front of dq (zero-filled): box {D, 32, 1}
  int dead;
  const float* lse;
  const uint16_t* o; int64_t o_bs, o_rs;
  int B, H, Nq, Nk;
  float scale;
};

// Work item = one (batch, head).  Key tiles of 128 rows (MMA M), query tiles of up to 128 columns (MMA N).
//   per (kt, qt):  S^T = K Q^T, dP^T = V dO^T  ->  TMEM           (warp 1)
//                  P^T = 2^(S^T sl - L), dS^T = P^T (dP^T - D)  -> smem bf16   (warps 2..9: lane = key row, the two warps of a
//                                                                 TMEM lane quadrant split the query columns)
//                  dV[kt] += P^T dO, dK[kt] += dS^T Q, dQ[qt] += dS K         (warp 1; accumulators stay in TMEM)
//   TMEM columns: S^T [0,128) | dP^T [128,256) | dV [256,320) | dK [320,384) | dQ tile 0 [384,448) | dQ tile 1 [448,512)
//   warp 0 loads the operand tiles (NP = 128: two stages, the next item's tiles land while this one computes) and issues
//   the TMA stores of the finished output tiles, so the compute warps never wait for a store.
template <int D, int NP>
__global__ void __launch_bounds__(320, 1) attn_tc_bwd_kernel(const __grid_constant__ AttnTcBwdParams p) {
  constexpr int STAGES = NP == 128 ? 2 : 1;
  constexpr uint32_t OP_BYTES = NP * 128;                          // one operand tile set: Q, K, V or dO
  constexpr uint32_t STAGE_BYTES = 4 * OP_BYTES;
  constexpr uint32_t TMEM_COLS = 512;
  constexpr uint32_t C_ST = 0, C_DP = 128, C_DV = 256, C_DK = 320, C_DQ = 384;
  constexpr int NCW = 8;                                            // compute warps
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t pt_base = smem_base + STAGES * STAGE_BYTES;       // P^T  [128 keys][128 queries] bf16: 2 tiles of 64 queries
  const uint32_t dst_base = pt_base + 2 * kTile;                   // dS^T, same layout
  const uint32_t ls_base = dst_base + 2 * kTile;                   // L[q] = lse * log2(e)   (f32, NP)
  const uint32_t ds_base = ls_base + NP * 4;                       // D[q] = dO_q . O_q       (f32, NP)
  const uint32_t zero_base = ds_base + NP * 4;                     // [32 rows][128 B] of zeros: source of the dead-row stores
  const uint32_t bar_base = zero_base + 4096;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 16u + 8u * s; };
  const uint32_t s_ready = bar_base + 32u, p_ready = bar_base + 40u, acc_ready = bar_base + 48u, out_ready = bar_base + 56u,
                 stg_free = bar_base + 64u, tmem_ptr_addr = bar_base + 72u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    const CUtensorMap* maps[7] = {&p.tq, &p.tk, &p.tv, &p.tdo, &p.tdq, &p.tdk, &p.tdv};
    for (int i = 0; i < 7; ++i) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(maps[i])) : "memory");
    for (int s = 0; s < 2; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(s_ready, 1);
    mbar_init(p_ready, NCW);
    mbar_init(acc_ready, 1);
    mbar_init(out_ready, NCW);
    mbar_init(stg_free, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp >= 2) {
    for (uint32_t i = threadIdx.x - 64; i < 4096 / 16; i += NCW * 32) sts_128(zero_base + 16u * i, 0u, 0u, 0u, 0u);
    fence_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr) : "memory");
  pdl_launch_dependents();     // set-up done, nothing global touched yet (PDL, common.cuh)
  pdl_wait();

  const int items = p.B * p.H;
  const int nqt = (p.Nq + 127) >> 7, nkt = (p.Nk + 127) >> 7;
  const int npairs = nqt * nkt;                                    // (kt, qt) pairs per item, qt fastest
  auto col_of = [&](int h) { return D == 64 ? h * 64 : (h >> 1) * 64; };
  auto koff_of = [&](int h) { return D == 64 ? 0u : (uint32_t)(h & 1) * 64u; };
  auto n16 = [](int n) { return (n + 15) & ~15; };
  auto stage_of = [&](int it) { return smem_base + (uint32_t)(it % STAGES) * STAGE_BYTES; };
  // staging tiles of the outputs alias P^T / dS^T (free between the last accumulate MMA of a key tile and the next pair)
  const uint32_t stg_dv = pt_base, stg_dk = dst_base, stg_dq0 = pt_base + kTile, stg_dq1 = dst_base + kTile;

  if (warp == 0) {
    // ===================== TMA producer + output stores =====================
    const bool leader = elect_one();
    auto load = [&](int w, int it) {
      const int h = w % p.H, b = w / p.H;
      const int s = it % STAGES;
      mbar_wait(empty_bar(s), ((it / STAGES) & 1) ^ 1u);
      if (leader) {
        const uint32_t qs = stage_of(it), ks = qs + OP_BYTES, vs = ks + OP_BYTES, dos = vs + OP_BYTES;
        mbar_expect_tx(full_bar(s), STAGE_BYTES);
        tma_load_3d(qs, &p.tq, full_bar(s), col_of(h), 0, b);
        tma_load_3d(ks, &p.tk, full_bar(s), col_of(h), 0, b);
        tma_load_3d(vs, &p.tv, full_bar(s), col_of(h), 0, b);
        tma_load_3d(dos, &p.tdo, full_bar(s), col_of(h), 0, b);
      }
      __syncwarp();
    };
    int it = 0;
    for (int k = 0; k < STAGES; ++k)
      if ((int)blockIdx.x + k * (int)gridDim.x < items) load(blockIdx.x + k * gridDim.x, k);
    uint32_t out_phase = 0;
    for (int w = blockIdx.x; w < items; w += gridDim.x, ++it) {
      const int h = w % p.H, b = w / p.H;
      for (int kt = 0; kt < nkt; ++kt) {
        mbar_wait(out_ready, out_phase);                           // dV / dK (/ dQ) staging tiles of this key tile are written
        out_phase ^= 1u;
        if (leader) {
          if (kt == 0)
            for (int r0 = 0; r0 < p.dead; r0 += 32) tma_store_3d(&p.tdq0, zero_base, h * D, r0, b);      // dead query slots := 0
          tma_store_3d(&p.tdv, stg_dv, h * D, kt * 128, b);
          tma_store_3d(&p.tdk, stg_dk, h * D, kt * 128, b);
          if (kt == nkt - 1) {
            tma_store_3d(&p.tdq, stg_dq0, h * D, 0, b);
            if (nqt > 1) tma_store_3d(&p.tdq, stg_dq1, h * D, 128, b);
          }
          bulk_commit();
          bulk_wait_read<0>();
          mbar_arrive(stg_free);
        }
        __syncwarp();
      }
      if (w + STAGES * (int)gridDim.x < items) load(w + STAGES * gridDim.x, it + STAGES);
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
      const int s = it % STAGES;
      const uint32_t qs = stage_of(it), ks = qs + OP_BYTES, vs = ks + OP_BYTES, dos = vs + OP_BYTES;
      mbar_wait(full_bar(s), (it / STAGES) & 1);
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
          else umma_commit(empty_bar(s));                          // every read of this stage's Q / K / V / dO has retired
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== compute + epilogue warps =====================
    const int quad = warp & 3;                                      // TMEM lane quadrant
    const int half = (warp - 2) >> 2;                               // which half of the query columns / output columns
    const int row = quad * 32 + lane;                               // key row inside the key tile
    const int tid = half * 128 + row;                               // 0..255 over the eight warps
    const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16);
    const float sl = p.scale * kLog2eTc;
    const uint32_t sw = (uint32_t)(row & 7);
    int it = 0;
    uint32_t pair_phase = 0, acc_phase = 0, stg_phase = 0;
    bool stg_pending = false;                                       // a TMA store of the staging tiles (= P^T / dS^T) may still be reading them
    // this warp's half (32 columns; d = 32: 16) of one 64-column accumulator row -> (x mul) -> bf16 -> staging tile
    auto stage_tile = [&](uint32_t tcol, uint32_t stg, float mul, int h) {
      if (D == 64) {
        uint32_t z[32];
        tmem_ld_32x32b_x32_issue(trow + tcol + (uint32_t)half * 32u, z);
        tmem_ld_wait(z);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          sts_128(stg + (uint32_t)row * 128u + ((((uint32_t)half * 4u + j) ^ sw) << 4),
                  pack_bf16x2(__uint_as_float(z[8 * j]) * mul, __uint_as_float(z[8 * j + 1]) * mul), pack_bf16x2(__uint_as_float(z[8 * j + 2]) * mul, __uint_as_float(z[8 * j + 3]) * mul),
                  pack_bf16x2(__uint_as_float(z[8 * j + 4]) * mul, __uint_as_float(z[8 * j + 5]) * mul), pack_bf16x2(__uint_as_float(z[8 * j + 6]) * mul, __uint_as_float(z[8 * j + 7]) * mul));
      } else {
        float z[16];
        tmem_ld_32x32b_x16(trow + tcol + (uint32_t)(h & 1) * 32u + (uint32_t)half * 16u, z);
#pragma unroll
        for (int j = 0; j < 2; ++j)
          sts_128(stg + (uint32_t)row * 64u + (((uint32_t)half * 2u + j) << 4),
                  pack_bf16x2(z[8 * j] * mul, z[8 * j + 1] * mul), pack_bf16x2(z[8 * j + 2] * mul, z[8 * j + 3] * mul),
                  pack_bf16x2(z[8 * j + 4] * mul, z[8 * j + 5] * mul), pack_bf16x2(z[8 * j + 6] * mul, z[8 * j + 7] * mul));
      }
    };
    for (int w = blockIdx.x; w < items; w += gridDim.x, ++it) {
      const int h = w % p.H, b = w / p.H;
      const uint32_t koff = koff_of(h);
      const uint32_t dos = stage_of(it) + 3 * OP_BYTES;
      // ---- L[q] and D[q] = dO_q . O_q for every query of the head (thread per query row); the forward output row is
      //      fetched from global memory before the wait on the operand tiles ----
      uint4 orow[D / 8];
      const bool have_q = tid < p.Nq;
      float Lq = 1e30f;                                              // padded queries: P = 2^(-inf) = 0, dS = 0
      if (have_q) {
        const uint16_t* op = p.o + (int64_t)b * p.o_bs + (int64_t)tid * p.o_rs + (int64_t)h * D;
#pragma unroll
        for (int j = 0; j < D / 8; ++j) orow[j] = *reinterpret_cast<const uint4*>(op + 8 * j);
        Lq = p.lse[((int64_t)b * p.H + h) * p.Nq + tid] * kLog2eTc;
      }
      mbar_wait(full_bar(it % STAGES), (it / STAGES) & 1);
      if (tid < NP) {
        float Dq = 0.f;
        if (have_q) {
          const uint32_t drow = dos + (uint32_t)tid * 128u;
          const uint32_t qsw = (uint32_t)(tid & 7);
#pragma unroll
          for (int j = 0; j < D / 8; ++j) {
            const uint4 yo = lds_128(drow + ((((koff >> 4) + (uint32_t)j) ^ qsw) << 4));
            const uint32_t xs[4] = {orow[j].x, orow[j].y, orow[j].z, orow[j].w}, ys[4] = {yo.x, yo.y, yo.z, yo.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 x = unpack_bf16x2(xs[e]), y = unpack_bf16x2(ys[e]);
              Dq = fmaf(x.x, y.x, fmaf(x.y, y.y, Dq));
            }
          }
        }
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(ls_base + 4u * tid), "f"(Lq) : "memory");
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(ds_base + 4u * tid), "f"(Dq) : "memory");
      }
      named_bar_sync(1, NCW * 32);
      for (int pr = 0; pr < npairs; ++pr) {
        const int kt = pr / nqt, qt = pr % nqt;
        const int nq_t = n16(min(128, p.Nq - qt * 128));
        const bool key_ok = kt * 128 + row < p.Nk;
        mbar_wait(s_ready, pair_phase);
        pair_phase ^= 1u;
        tc_fence_after();
        // s_ready also tells that the accumulate MMAs of the previous pair have retired: P^T / dS^T may be overwritten --
        // once the TMA stores that used them as staging tiles have read them
        if (stg_pending) { mbar_wait(stg_free, stg_phase); stg_phase ^= 1u; stg_pending = false; }
        for (int c = half * 32; c < nq_t; c += 64) {               // the two warps of a lane quadrant alternate 32-column chunks
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
          // ---- dV / dK of this key tile (and, after the last one, dQ): every MMA issued so far has retired, so the
          //      accumulators are final and P^T / dS^T are free to serve as staging tiles ----
          mbar_wait(acc_ready, acc_phase);
          acc_phase ^= 1u;
          tc_fence_after();
          stage_tile(C_DV, stg_dv, 1.0f, h);
          stage_tile(C_DK, stg_dk, p.scale, h);
          if (kt == nkt - 1) {
            stage_tile(C_DQ, stg_dq0, p.scale, h);
            if (nqt > 1) stage_tile(C_DQ + 64, stg_dq1, p.scale, h);
          }
          fence_async_smem();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(out_ready);                    // warp 0 issues the TMA stores
          stg_pending = true;
        }
      }
    }
    if (stg_pending) mbar_wait(stg_free, stg_phase);                // (keeps the CTA alive until its last stores have read shared memory)
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
  // 8: the fusion block's 8-query cross attentions too (backward 29.8 -> 20.2 us at 49 keys, step 16.29 -> 16.08 ms); below that
  // (and for the dqk = 16 pair attention) the mma.sync kernels serve
  static const int v = [] { const char* e = getenv("DAVF_ATTN_TC_MINQ"); return e ? atoi(e) : 8; }();
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
  constexpr size_t smem = 2 * (size_t)kTile + 2 * (size_t)NKP * 128 + 64 + 1024;      // Q | K | X | V (+ barriers, alignment slack)
  static_assert(2 * (smem + 1024) <= 233472, "two CTAs per SM must fit the 228 KB shared memory");
  auto kern = attn_tc_fwd_kernel<D, NKP>;
  static bool attr_set = false;
  if (!attr_set) {
    DAVF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  const int items = a.B * a.H * ((a.Nq + 127) / 128);
  const int grid = items < 2 * kNumSMs ? items : 2 * kNumSMs;        // two resident CTAs per SM
  DAVF_CUDA(launch_pdl(kern, dim3(grid), dim3(192), smem, st, p));
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
  p.dead = a.dq_dead_rows;
  p.tdq0 = p.tdq;
  if (a.dq_dead_rows > 0 &&
      (rc = get_map3(a.dq - (int64_t)a.dq_dead_rows * a.dq_rs, cols, a.dq_dead_rows, a.B, a.dq_rs, a.dq_bs, D, 32, &p.tdq0))) return rc;
  p.lse = a.lse; p.o = a.o; p.o_bs = a.o_bs; p.o_rs = a.o_rs;
  p.B = a.B; p.H = a.H; p.Nq = a.Nq; p.Nk = a.Nk; p.scale = a.scale;
  constexpr size_t smem = (NP == 128 ? 2 : 1) * 4 * (size_t)NP * 128 + 4 * (size_t)kTile + 2 * (size_t)NP * 4 + 4096 + 128 + 1024;
  static_assert(smem <= 232448, "exceeds the 227 KB shared memory of an SM");
  auto kern = attn_tc_bwd_kernel<D, NP>;
  static bool attr_set = false;
  if (!attr_set) {
    DAVF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  const int items = a.B * a.H;
  const int grid = items < kNumSMs ? items : kNumSMs;
  DAVF_CUDA(launch_pdl(kern, dim3(grid), dim3(320), smem, st, p));
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
