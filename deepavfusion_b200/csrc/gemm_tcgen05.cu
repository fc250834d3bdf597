// K1/K5/K9/K10: bf16 GEMM on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM),
// operands staged by TMA into 128B-swizzled shared memory, persistent over output tiles, warp
// specialised:
//     warp 0      TMA producer            (one elected lane)
//     warp 1      TMEM allocator + MMA issuer (one elected lane issues tcgen05.mma / commit)
//     warps 2..9  epilogue (two warps per TMEM lane quadrant, each owning half of the tile's columns):
//                 tcgen05.ld TMEM -> registers -> per-warp smem staging (transposes "one row per lane"
//                 into "4 lanes per row") -> fused epilogue with fully coalesced global loads / stores
// Pipelines: smem ring (full/empty mbarriers, TMA <-> MMA) and a double-buffered TMEM accumulator
// (tmem_full/tmem_empty, MMA <-> epilogue) so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// Operand layouts.  Both operands may be K-major (reduction dim contiguous in global memory) or
// MN-major (row dim contiguous), which covers forward (K,K), dgrad (K,MN) and wgrad (MN,MN)
// without any transposed copies in HBM:
//   K-major  tile: [BLOCK rows][64 k] bf16, one TMA box {64,BLOCK}, canonical SW128 K-major layout
//                  (SBO = 1024 B between 8-row groups, +32 B per UMMA_K step)
//   MN-major tile: BLOCK/64 boxes {64 rows, 64 k}: [64 k][64 rows] each 8 KB, canonical SW128
//                  MN-major layout (LBO = 8192 B between 64-row groups, SBO = 1024 B between
//                  8-k groups, +2048 B per UMMA_K step)
// Descriptor bit layouts follow the PTX ISA "tcgen05 shared memory descriptor" / "instruction
// descriptor" tables (cross-checked against cute/arch/mma_sm100_desc.hpp).
#include <cuda.h>
#include <mutex>
#include <unordered_map>
#include <string.h>
#include <stdlib.h>

#include "gemm_epilogue.cuh"
#include "tc_ptx.cuh"

namespace davf {

constexpr int BM = 128;
constexpr int BK = 64;            // 64 bf16 = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int kStageCols = 16;                         // columns per epilogue chunk
constexpr int kStagePitch = 20;                        // floats per staged row (80 B: conflict-free float4 writes)
__host__ __device__ constexpr uint32_t staging_bytes(int epi_warps) { return epi_warps * 32 * kStagePitch * 4; }
__host__ __device__ constexpr bool direct_epi(int epi) {   // compile-time epilogues: registers -> swizzled smem box -> TMA store / reduce
  return epi >= 0;
}
constexpr uint32_t kTmaBox = 4096;                     // one TMA box of the direct epilogue: 32 rows x 128 B (64 bf16), SW128
// epilogues with a matrix operand (f32 residual: EPI_RES = 16, saved gelu': EPI_DGELU = 4) run a ring of three boxes per warp
__host__ __device__ constexpr bool ring_epi(int epi) { return epi >= 0 && (epi & (16 | 4)) != 0; }
// (two for the single-CTA 256-wide tile, whose 48 KB pipeline stages leave no room for a third)
__host__ __device__ constexpr uint32_t ring_boxes(int bn, int cg) { return (bn == 256 && cg == 1) ? 2u : 3u; }
__host__ __device__ constexpr uint32_t staging_bytes(int epi, int epi_warps, int bn, int cg) {
  return direct_epi(epi) ? epi_warps * (ring_epi(epi) ? ring_boxes(bn, cg) : 2u) * kTmaBox : staging_bytes(epi_warps);
}
__host__ __device__ constexpr uint32_t ones_bytes(int epi) { return (epi >= 0 && (epi & 128) == 0) ? 0u : 2048u; }   // all-ones B tile of the row-sum MMA
// pipeline depth: what the caller asks for, capped by what fits next to the epilogue staging in 227 KB
__host__ __device__ constexpr int fit_stages(int want, int epi, int ew, int bn, int cg) {
  const int stage = 128 * 64 * 2 + (bn / cg) * 64 * 2;
  const int avail = 232448 - 1024 - 16 - 8 * (2 * want + 5 + 3 * ew) - (int)ones_bytes(epi) - (int)staging_bytes(epi, ew, bn, cg);
  return avail / stage < want ? avail / stage : want;
}

constexpr int kMaxGroup = 6;                          // problems per grouped launch (kernel parameter block ~2.6 KB)
// Work decomposition.  Classic: every (m, n) tile is cut into `splits` equal K ranges.  Stream-K (sk_units > 0, wgrad
// only: partial products are f32-ADDED to the output, so any partition of the K loop is valid): the linearised
// (tile, k-block) space is cut into sk_units equal ranges of sk_len k-blocks, one per CTA (pair); a range touches at
// most two tiles (sk_len <= kb_total), which appear as "tile" t = unit and t = unit + sk_units (possibly empty:
// kb1 <= kb0).  Every CTA pair then runs the same number of k-blocks, whatever the tile count.
struct TileSched {
  int m_tiles, n_tiles, splits, kb_total, kb_per_split;
  int sk_units, sk_len;
  __host__ __device__ __forceinline__ int num_tiles() const { return sk_units > 0 ? 2 * sk_units : m_tiles * n_tiles * splits; }
  __device__ __forceinline__ void decode(int t, int& m_blk, int& n_blk, int& sp, int& kb0, int& kb1) const {
    if (sk_units > 0) {
      const int unit = t % sk_units, second = t / sk_units;
      const int total = m_tiles * n_tiles * kb_total;
      const int lo = unit * sk_len, hi = min(total, lo + sk_len);
      int mn = lo / kb_total;
      const int bnd = (mn + 1) * kb_total;
      int a = lo, b = min(hi, bnd);
      if (second) { a = bnd; b = hi; ++mn; }
      kb0 = a - mn * kb_total;
      kb1 = b - mn * kb_total;           // <= kb0 when the range does not reach a second tile
      m_blk = mn % m_tiles;
      n_blk = mn / m_tiles;
      sp = kb0 == 0 ? 0 : 1;
      return;
    }
    sp = t % splits;
    const int mn = t / splits;
    m_blk = mn % m_tiles;
    n_blk = mn / m_tiles;
    kb0 = sp * kb_per_split;
    kb1 = min(kb_total, kb0 + kb_per_split);
  }
};

// Kernel parameters: NG independent GEMM problems served by ONE launch (NG = 1: the ordinary case).  The
// tile index space is the concatenation of the problems' tiles (tile_end = exclusive prefix sums; unused
// slots repeat the last value), so the many tiny GEMMs of a fusion block (M = 512 rows, 24 tiles each,
// latency-bound at ~10 us per launch) share one launch, one prologue and one wave of CTAs.
template <int NG>
struct GemmGroup {
  CUtensorMap ta[NG], tb[NG];
  CUtensorMap tc[NG], taux[NG];      // TMA epilogue: out / (aux_out | aux_in | res) as [M][N] tensors, box = 32 rows x 128 bytes
  EpiParams ep[NG];
  TileSched ts[NG];
  int tile_end[NG];
};

// Compile-time epilogue specialisation.  EPI < 0: every option is a run-time flag (rare shapes: row
// windows, gathered residuals, ...).  EPI >= 0: bit mask of the options below, fixed at compile time
// for the six epilogues that carry > 90 % of the GEMM time, so that no flag tests, dead operand
// prefetches or dead address arithmetic remain in the (instruction-bound) epilogue loop.
enum : int { EPI_BIAS = 1, EPI_GELU = 2, EPI_DGELU = 4, EPI_AUX = 8, EPI_RES = 16, EPI_BF16 = 32, EPI_RED = 64, EPI_ROWSUM = 128 };
template <int EPI>
struct EpiSel {
  static constexpr bool kStatic = EPI >= 0;
  __device__ __forceinline__ static bool bias(const EpiParams& p) { return kStatic ? (EPI & EPI_BIAS) != 0 : p.bias != nullptr; }
  __device__ __forceinline__ static bool gelu(const EpiParams& p) { return kStatic ? (EPI & EPI_GELU) != 0 : p.act == DAVF_ACT_GELU; }
  __device__ __forceinline__ static bool dgelu(const EpiParams& p) { return kStatic ? (EPI & EPI_DGELU) != 0 : p.act == DAVF_ACT_DGELU; }
  __device__ __forceinline__ static bool aux(const EpiParams& p) { return kStatic ? (EPI & EPI_AUX) != 0 : p.aux_out != nullptr; }
  __device__ __forceinline__ static bool res(const EpiParams& p) { return kStatic ? (EPI & EPI_RES) != 0 : p.res != nullptr; }
  __device__ __forceinline__ static bool bf16(const EpiParams& p) { return kStatic ? (EPI & EPI_BF16) != 0 : p.out_bf16 != 0; }
  __device__ __forceinline__ static bool red(const EpiParams& p) { return kStatic ? (EPI & EPI_RED) != 0 : p.accumulate != 0; }
  __device__ __forceinline__ static bool remap(const EpiParams& p) { return kStatic ? false : (p.g > 0 || p.res_idx != nullptr); }
  __device__ __forceinline__ static bool rowsum(const EpiParams& p) { return kStatic ? (EPI & EPI_ROWSUM) != 0 : p.rowsum_out != nullptr; }
};

// CG = 1: one CTA computes a 128 x BN tile.  CG = 2: a CTA pair (cluster of 2 on one TPC) computes a 256 x BN
// tile with tcgen05.mma.cta_group::2 -- each CTA stages its own 128 rows of A and HALF of B (BN/2 rows), the
// MMA reads both halves, so the shared-memory fill traffic per FLOP halves (the 1-CTA kernel is bound by the
// per-SM TMA fill rate, not by the tensor pipe).  TileSched m-blocks are 128*CG rows.
// EW = number of epilogue warps (8, or 16 for the transcendental-heavy GELU / dGELU epilogues, which are
// instruction-bound: EW/4 warps share a TMEM lane quadrant and split the tile's columns).
template <int BN, int STAGES, bool A_KMAJOR, bool B_KMAJOR, int EPI, int CG, int EW, int NG>
__global__ void __launch_bounds__(64 + 32 * EW, 1)
gemm_tc_kernel(const __grid_constant__ GemmGroup<NG> gp) {
  static_assert(NG == 1 || EPI < 0 || (EPI & (EPI_GELU | EPI_DGELU | EPI_AUX | EPI_RES)) == 0, "grouped compile-time epilogues: bias / bf16 / f32 reduce only");
  // tile t of the launch -> problem p, tile tl of that problem
  auto locate = [&](int t, int& p, int& tl) {
    p = 0; tl = t;
    if (NG > 1) {
      while (p < NG - 1 && t >= gp.tile_end[p]) ++p;
      if (p) tl = t - gp.tile_end[p - 1];
    }
  };
  constexpr int BNL = BN / CG;                       // B rows staged by this CTA
  constexpr uint32_t A_BYTES = BM * BK * 2;
  constexpr uint32_t B_BYTES = BNL * BK * 2;
  constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  const uint32_t cta_rank = CG == 2 ? cluster_ctarank() : 0u;
  const bool lead_cta = cta_rank == 0;
  const int cluster_id = blockIdx.x / CG, num_clusters = gridDim.x / CG;
  constexpr uint32_t TMEM_COLS = 512;
  // accumulator stages: two (epilogue of tile i overlaps the MMAs of tile i+1) unless the tile is 256 columns wide AND
  // carries the 16 row-sum columns (wgrad + bias gradient): TMEM has 512 columns, so that case is single-buffered
  constexpr bool kStaticRowsum = EPI >= 0 && (EPI & EPI_ROWSUM) != 0;
  constexpr int NACC = (BN == 256 && kStaticRowsum) ? 1 : 2;
  constexpr uint32_t IDESC = make_idesc(BM * CG, BN, A_KMAJOR ? 0 : 1, B_KMAJOR ? 0 : 1);
  constexpr uint32_t IDESC_ONES = make_idesc(BM * CG, 16, A_KMAJOR ? 0 : 1, 0);
  constexpr uint32_t ROWSUM_COL = NACC * BN;        // row-sum columns follow the accumulator stages (BN = 256 only when NACC = 1)
  static_assert(NACC * BN + NACC * 16 <= 512 || !kStaticRowsum, "TMEM overflow");

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;   // SW128 atoms need 1024 B alignment
  const uint32_t stage_base = smem_base + STAGES * STAGE_BYTES;          // epilogue staging, kStagingBytes
  const uint32_t ones_base = stage_base + staging_bytes(EPI, EW, BN, CG);              // 1024-byte aligned, kOnesBytes
  const uint32_t bar_base = ones_base + ones_bytes(EPI);
  bool want_rowsum = false;
#pragma unroll
  for (int p = 0; p < NG; ++p) want_rowsum |= EpiSel<EPI>::rowsum(gp.ep[p]);
  // barrier addresses: full[s] | empty[s] | tmem_full[2] | tmem_empty[2] | tmem_ptr
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
  const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * STAGES + 4);
  auto aux_bar = [&](int w, int b) { return bar_base + 8u * (2 * STAGES + 5 + 3 * w + b); };     // TMA loads of the epilogue's matrix operand: 3 boxes per warp

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
#pragma unroll
    for (int p = 0; p < NG; ++p) {
      if (p == 0 || gp.tile_end[p] > gp.tile_end[p - 1]) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&gp.ta[p])) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&gp.tb[p])) : "memory");
      }
    }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), EW * CG);     // the leader's barrier collects the epilogue warps of both CTAs
    }
    for (int w = 0; w < EW; ++w)
      for (int b = 0; b < 3; ++b) mbar_init(aux_bar(w, b), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if constexpr (ones_bytes(EPI) > 0) if (want_rowsum) {      // bf16 1.0 everywhere: layout-agnostic B operand
    for (uint32_t i = threadIdx.x; i < ones_bytes(EPI) / 4; i += blockDim.x)
      asm volatile("st.shared.u32 [%0], %1;" ::"r"(ones_base + 4 * i), "r"(0x3F803F80u) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to tcgen05.mma
  }
  if (warp == 1) {
    if (CG == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr), "r"(TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr), "r"(TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();   // barriers of both CTAs initialised before any remote arrive / TMA
  tc_fence_after();
  // set-up done (barriers, TMEM, descriptor prefetch: nothing global was touched): let the next kernel of the stream start
  // its own set-up, then wait until the previous kernel's results are visible (see common.cuh, PDL)
  pdl_launch_dependents();
  pdl_wait();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr) : "memory");

  const int num_tiles = gp.tile_end[NG - 1];
  long long* dbg = (NG == 1 && CG == 1 && gp.ep[0].debug_clocks && blockIdx.x == 0) ? gp.ep[0].debug_clocks : nullptr;   // optional timeline probe of CTA 0
  if (dbg && threadIdx.x == 0) dbg[0] = clock64();

  if (warp == 0) {
    // ===================== TMA producer =====================
    // The whole warp runs the loop (so ptxas keeps the loop state and addresses in uniform registers);
    // one elected lane issues the expect_tx + TMA instructions.
    {
      const bool leader = elect_one();
      int stage = 0;
      uint32_t phase = 0;
      auto tma = [&](uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
        if (CG == 2) tma_load_2d_2cta(dst, map, bar, c0, c1); else tma_load_2d(dst, map, bar, c0, c1);
      };
      for (int t = cluster_id; t < num_tiles; t += num_clusters) {
        int p, tl, m_blk, n_blk, sp, kb0, kb1;
        locate(t, p, tl);
        gp.ts[p].decode(tl, m_blk, n_blk, sp, kb0, kb1);
        if (kb1 <= kb0) continue;                   // empty stream-K segment
        const CUtensorMap* tmap_a = &gp.ta[p];
        const CUtensorMap* tmap_b = &gp.tb[p];
        const int m0 = m_blk * (BM * CG) + (int)cta_rank * BM;      // this CTA's rows of A
        const int n0 = n_blk * BN + (int)cta_rank * BNL;             // this CTA's rows of B
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          if (dbg && t == 0 && kb == kb0 && leader) dbg[1] = clock64();
          const uint32_t sa = smem_base + stage * STAGE_BYTES;
          const uint32_t sb = sa + A_BYTES;
          if (leader) {
            if (lead_cta) mbar_expect_tx(full_bar(stage), STAGE_BYTES * CG);   // bytes of both CTAs land on the leader's barrier
            if (A_KMAJOR) {
              tma(sa, tmap_a, full_bar(stage), kb * BK, m0);
            } else {
#pragma unroll
              for (int j = 0; j < BM / 64; ++j) tma(sa + j * 8192, tmap_a, full_bar(stage), m0 + j * 64, kb * BK);
            }
            if (B_KMAJOR) {
              tma(sb, tmap_b, full_bar(stage), kb * BK, n0);
            } else {
#pragma unroll
              for (int j = 0; j < BNL / 64; ++j) tma(sb + j * 8192, tmap_b, full_bar(stage), n0 + j * 64, kb * BK);
            }
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // Warp-uniform loop (barrier waits, descriptor arithmetic in uniform registers); one elected lane
    // issues tcgen05.mma / tcgen05.commit.  With the loop inside `if (lane == 0)` every operand had to be
    // moved vector->uniform register per instruction and the issue path, not the tensor pipe, set the pace.
    if (lead_cta) {                                  // the pair's MMAs are issued by the leader CTA only
      const bool leader = elect_one();
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      for (int t = cluster_id; t < num_tiles; t += num_clusters) {
        int p, tl, m_blk, n_blk, sp, kb0, kb1;
        locate(t, p, tl);
        gp.ts[p].decode(tl, m_blk, n_blk, sp, kb0, kb1);
        if (kb1 <= kb0) continue;                   // empty stream-K segment (no accumulator stage is consumed)
        const int acc = local % NACC;
        const uint32_t acc_phase = (local / NACC) & 1u;
        ++local;
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        const bool rowsum_tile = EpiSel<EPI>::rowsum(gp.ep[p]) && n_blk == 0;
        const uint64_t ones_desc = make_smem_desc(ones_base, 16, 1024);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          if (dbg && t == 0 && kb == kb0 && leader) dbg[2] = clock64();
          const uint32_t sa = smem_base + stage * STAGE_BYTES;
          const uint32_t sb = sa + A_BYTES;
          const uint64_t adesc0 = A_KMAJOR ? make_smem_desc(sa, 16, 1024) : make_smem_desc(sa, 8192, 1024);
          const uint64_t bdesc0 = B_KMAJOR ? make_smem_desc(sb, 16, 1024) : make_smem_desc(sb, 8192, 1024);
          constexpr uint32_t A_STEP = (A_KMAJOR ? UMMA_K * 2 : UMMA_K * 128) >> 4;   // descriptor address units of 16 B
          constexpr uint32_t B_STEP = (B_KMAJOR ? UMMA_K * 2 : UMMA_K * 128) >> 4;
          if (leader) {
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              const uint64_t adesc = adesc0 + (uint64_t)(k * A_STEP), bdesc = bdesc0 + (uint64_t)(k * B_STEP);
              if (CG == 2) umma_f16_2cta(tmem_d, adesc, bdesc, IDESC, (kb > kb0 || k > 0) ? 1u : 0u);
              else umma_f16(tmem_d, adesc, bdesc, IDESC, (kb > kb0 || k > 0) ? 1u : 0u);
              if (rowsum_tile) {     // D[:, 0:16] += A * ones^T  ->  every column holds sum_k A(m, k)
                if (CG == 2) umma_f16_2cta(tmem_base + ROWSUM_COL + acc * 16, adesc, ones_desc, IDESC_ONES, (kb > kb0 || k > 0) ? 1u : 0u);
                else umma_f16(tmem_base + ROWSUM_COL + acc * 16, adesc, ones_desc, IDESC_ONES, (kb > kb0 || k > 0) ? 1u : 0u);
              }
            }
            if (CG == 2) umma_commit_2cta(empty_bar(stage)); else umma_commit(empty_bar(stage));   // frees the smem slot(s) once these MMAs retire
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        if (leader) { if (CG == 2) umma_commit_2cta(tfull_bar(acc)); else umma_commit(tfull_bar(acc)); }   // accumulator complete -> epilogue(s)
        __syncwarp();
        if (dbg && t == 0 && leader) dbg[3] = clock64();
      }
    }
  } else if constexpr (ring_epi(EPI)) {
    // ===================== TMA epilogue with a matrix operand: f32 residual (proj / fc2) or saved gelu' (fc2 dgrad) =====================
    // As below, but the operand box is loaded by TMA TWO groups ahead into a ring of three boxes per warp and the result
    // is computed IN PLACE in the box the operand arrived in, which then leaves by a TMA store.  With one operand box
    // the L2 / HBM latency of every box load (~1 k clk) was exposed once per 32-row x 128-byte group, i.e. the
    // epilogue of a K = 512 tile took longer than its MMAs.
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    const int ew = warp - 2;
    constexpr int SLICE = BN / (EW / 4);
    constexpr bool kBias = (EPI & EPI_BIAS) != 0, kMul = (EPI & EPI_DGELU) != 0, kRes = (EPI & EPI_RES) != 0, kF32 = (EPI & EPI_BF16) == 0;
    static_assert((kRes && kF32 && !kMul) || (kMul && !kF32 && !kRes), "ring epilogue: f32 + residual, or bf16 x gelu'");
    static_assert((EPI & (EPI_GELU | EPI_AUX | EPI_RED | EPI_ROWSUM)) == 0, "ring epilogue: no second output / reduction");
    constexpr int GW = kF32 ? 32 : 64;                  // columns per box (128 bytes per row)
    constexpr int NGRP = SLICE / GW;
    static_assert(SLICE % GW == 0, "TMA epilogue works on whole boxes");
    constexpr int NBOX = (int)ring_boxes(BN, CG);       // operand boxes in flight: NBOX - 1 groups ahead
    const uint32_t boxes = stage_base + (uint32_t)ew * (uint32_t)NBOX * kTmaBox;
    const uint32_t my_row = (uint32_t)lane * 128u, sw = (uint32_t)(lane & 7);
    // load cursor: the next ACTIVE group (tile lt, group lg) whose operand box has not been requested yet
    int lt = cluster_id - num_clusters, lg = NGRP - 1, l_p = 0, l_m0 = 0, l_nbase = 0;
    bool l_ok = false, l_valid = true;
    int issued = 0, done = 0;
    auto lc_advance = [&]() {
      while (true) {
        if (++lg == NGRP) {
          lg = 0;
          lt += num_clusters;
          if (lt >= num_tiles) { l_valid = false; return; }
          int tl, m_blk, n_blk, sp, kb0, kb1;
          locate(lt, l_p, tl);
          gp.ts[l_p].decode(tl, m_blk, n_blk, sp, kb0, kb1);
          l_m0 = m_blk * (BM * CG) + (int)cta_rank * BM + quad * 32;
          l_nbase = n_blk * BN + half * SLICE;
          l_ok = kb1 > kb0 && l_m0 < (int)gp.ep[l_p].M;
        }
        if (l_ok && l_nbase + lg * GW < (int)gp.ep[l_p].N) return;
      }
    };
    auto issue_load = [&]() {                            // (warp-uniform bookkeeping; lane 0 talks to the TMA unit)
      if (!l_valid) return;
      const int b = issued % NBOX;
      if (lane == 0) {
        mbar_expect_tx(aux_bar(ew, b), kTmaBox);
        tma_load_2d(boxes + (uint32_t)b * kTmaBox, &gp.taux[l_p], aux_bar(ew, b), l_nbase + lg * GW, l_m0);
      }
      ++issued;
      lc_advance();
    };
    lc_advance();
    for (int i = 0; i < NBOX - 1; ++i) issue_load();      // boxes in flight before the first accumulator is ready
    int local = 0;
    for (int t = cluster_id; t < num_tiles; t += num_clusters) {
      int p, tl, m_blk, n_blk, sp, kb0, kb1;
      locate(t, p, tl);
      gp.ts[p].decode(tl, m_blk, n_blk, sp, kb0, kb1);
      if (kb1 <= kb0) continue;
      const EpiParams& ep = gp.ep[p];
      const CUtensorMap* tmap_c = &gp.tc[p];
      const int acc = local % NACC;
      const uint32_t acc_phase = (local / NACC) & 1u;
      ++local;
      const int m0 = m_blk * (BM * CG) + (int)cta_rank * BM + quad * 32;     // first row of this warp's boxes
      const int n_base = n_blk * BN + half * SLICE;
      const bool live = m0 < (int)ep.M;
      const bool add_bias = kBias && sp == 0 && ep.bias != nullptr;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN + half * SLICE);
#pragma unroll 1
      for (int g = 0; g < NGRP; ++g) {
        uint32_t z[GW / 32][32];
#pragma unroll
        for (int h = 0; h < GW / 32; ++h) tmem_ld_32x32b_x32_issue(trow + g * GW + h * 32, z[h]);
        const int n0 = n_base + g * GW;
        const bool active = live && n0 < (int)ep.N;   // warp-uniform; the same predicate drives the load cursor
        const uint32_t box = boxes + (uint32_t)(done % NBOX) * kTmaBox;
        if (active) mbar_wait(aux_bar(ew, done % NBOX), (uint32_t)(done / NBOX) & 1u);       // the operand of this group has landed
#pragma unroll
        for (int h = 0; h < GW / 32; ++h) tmem_ld_wait(z[h]);
        if (g == NGRP - 1) {                          // accumulator fully read: hand the TMEM stage back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (CG == 2 && !lead_cta) mbar_arrive_remote(tempty_bar(acc), 0);
            else mbar_arrive(tempty_bar(acc));
          }
        }
        if (active) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {               // one 16-byte piece of the row segment: 8 bf16 or 4 f32 columns, in place
            const uint32_t pos = box + my_row + (((uint32_t)j ^ sw) << 4);
            const uint4 u = lds_128(pos);
            if constexpr (kF32) {
              float v[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) v[e] = __uint_as_float(z[0][j * 4 + e]);
              if (add_bias && n0 + j * 4 < (int)ep.N) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(ep.bias + n0 + j * 4));
                v[0] += b.x; v[1] += b.y; v[2] += b.z; v[3] += b.w;
              }
              sts_128(pos, __float_as_uint(v[0] + __uint_as_float(u.x)), __float_as_uint(v[1] + __uint_as_float(u.y)),
                      __float_as_uint(v[2] + __uint_as_float(u.z)), __float_as_uint(v[3] + __uint_as_float(u.w)));
            } else {
              float v[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(z[j >> 2][(j & 3) * 8 + e]);
              const float2 a0 = unpack_bf16x2(u.x), a1 = unpack_bf16x2(u.y), a2 = unpack_bf16x2(u.z), a3 = unpack_bf16x2(u.w);
              sts_128(pos, pack_bf16x2(v[0] * a0.x, v[1] * a0.y), pack_bf16x2(v[2] * a1.x, v[3] * a1.y),
                      pack_bf16x2(v[4] * a2.x, v[5] * a2.y), pack_bf16x2(v[6] * a3.x, v[7] * a3.y));
            }
          }
          fence_async_smem();                          // generic-proxy writes -> visible to the TMA engine
        }
        __syncwarp();
        if (active) {
          if (lane == 0) {
            tma_store_2d(tmap_c, box, n0, m0);
            bulk_commit();
            bulk_wait_read<1>();                       // every store but this one has read its box: the box of the previous group is free
          }
          ++done;
          issue_load();                                // operand of the group after next, into the box just freed
        }
      }
    }
    if (lane == 0) bulk_wait_read<0>();                // the boxes must outlive the stores that read them
    __syncwarp();
  } else if constexpr (direct_epi(EPI)) {
    // ===================== TMA epilogue (every compile-time epilogue) =====================
    // tcgen05.ld hands every lane consecutive columns of ITS row.  The lane applies the epilogue in registers and
    // writes its 128-byte row segment (64 bf16 or 32 f32 columns) into a per-warp 32 x 128 B box in the
    // 128B-swizzled layout (conflict-free 16-byte stores); one lane hands the box to the TMA engine, which writes
    // (or, for wgrad, f32-ADDS: cp.reduce.async.bulk) full lines to global memory and clips the M / N tails.
    // Operands of the epilogue that are matrices (the saved gelu' of a backward launch, the f32 residual) arrive
    // the same way, by a TMA load into the warp's second box.  No per-element address arithmetic and no LSU-side
    // scatter: 32 lanes storing 32 different rows with ordinary vector stores were measured LSU-bound at ~2 TB/s.
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    const int ew = warp - 2;
    constexpr int SLICE = BN / (EW / 4);
    constexpr bool kBias = (EPI & EPI_BIAS) != 0, kGelu = (EPI & EPI_GELU) != 0, kMul = (EPI & EPI_DGELU) != 0, kAux = (EPI & EPI_AUX) != 0;
    constexpr bool kRes = (EPI & EPI_RES) != 0, kRed = (EPI & EPI_RED) != 0, kRowsum = (EPI & EPI_ROWSUM) != 0, kF32 = (EPI & EPI_BF16) == 0;
    static_assert(!(kF32 && (kGelu || kMul || kAux)) && !(kRes && !kF32) && !(kRed && !kF32), "unsupported compile-time epilogue");
    constexpr int GW = kF32 ? 32 : 64;                  // columns per box (128 bytes per row)
    constexpr int NGRP = SLICE / GW;
    static_assert(SLICE % GW == 0, "TMA epilogue works on whole boxes");
    constexpr bool kLoad = kMul || kRes;                // box1 receives a TMA load
    constexpr bool kPingPong = !kAux && !kLoad;         // both boxes serve the single output stream alternately
    const uint32_t box0 = stage_base + (uint32_t)ew * 2u * kTmaBox, box1 = box0 + kTmaBox;
    const uint32_t my_row = (uint32_t)lane * 128u, sw = (uint32_t)(lane & 7);
    uint32_t aux_phase = 0;
    int local = 0, grp_count = 0;
    for (int t = cluster_id; t < num_tiles; t += num_clusters) {
      int p, tl, m_blk, n_blk, sp, kb0, kb1;
      locate(t, p, tl);
      gp.ts[p].decode(tl, m_blk, n_blk, sp, kb0, kb1);
      if (kb1 <= kb0) continue;
      const EpiParams& ep = gp.ep[p];
      const CUtensorMap* tmap_c = &gp.tc[p];
      const CUtensorMap* tmap_x = &gp.taux[p];
      const int acc = local % NACC;
      const uint32_t acc_phase = (local / NACC) & 1u;
      ++local;
      const int m0 = m_blk * (BM * CG) + (int)cta_rank * BM + quad * 32;     // first row of this warp's boxes
      const int n_base = n_blk * BN + half * SLICE;
      const bool live = m0 < (int)ep.M;                // a box entirely below the matrix is neither loaded nor stored
      const bool add_bias = kBias && sp == 0 && ep.bias != nullptr;
      if (kLoad && live && lane == 0 && n_base < (int)ep.N) {     // operand box of group 0: in flight while the MMAs of this tile still run
        mbar_expect_tx(aux_bar(ew, 0), kTmaBox);
        tma_load_2d(box1, tmap_x, aux_bar(ew, 0), n_base, m0);
      }
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      if (kRowsum && ep.rowsum_out != nullptr && n_blk == 0 && half == 0) {
        const float rs = tmem_ld_32x32b_x1(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(ROWSUM_COL + acc * 16));
        if (m0 + lane < (int)ep.M) atomicAdd(ep.rowsum_out + m0 + lane, rs);
      }
      const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN + half * SLICE);
#pragma unroll 1
      for (int g = 0; g < NGRP; ++g, ++grp_count) {
        uint32_t z[GW / 32][32];
#pragma unroll
        for (int h = 0; h < GW / 32; ++h) tmem_ld_32x32b_x32_issue(trow + g * GW + h * 32, z[h]);
        const int n0 = n_base + g * GW;
        const uint32_t obox = kPingPong ? ((grp_count & 1) ? box1 : box0) : box0;
        // the TMA store that last read this box must have finished reading it
        if (lane == 0) { if (kPingPong) bulk_wait_read<1>(); else bulk_wait_read<0>(); }
        __syncwarp();
#pragma unroll
        for (int h = 0; h < GW / 32; ++h) tmem_ld_wait(z[h]);
        if (g == NGRP - 1) {                          // accumulator fully read: hand the TMEM stage back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (CG == 2 && !lead_cta) mbar_arrive_remote(tempty_bar(acc), 0);
            else mbar_arrive(tempty_bar(acc));
          }
        }
        const bool active = live && n0 < (int)ep.N;   // warp-uniform
        if (active) {
          if (kLoad) { mbar_wait(aux_bar(ew, 0), aux_phase); aux_phase ^= 1u; }
#pragma unroll
          for (int j = 0; j < 8; ++j) {               // one 16-byte piece of the row segment: 8 bf16 or 4 f32 columns
            const uint32_t pos = my_row + (((uint32_t)j ^ sw) << 4);
            if constexpr (kF32) {
              float v[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) v[e] = __uint_as_float(z[0][j * 4 + e]);
              if (add_bias && n0 + j * 4 < (int)ep.N) {   // same address in every lane: one broadcast
                const float4 b = __ldg(reinterpret_cast<const float4*>(ep.bias + n0 + j * 4));
                v[0] += b.x; v[1] += b.y; v[2] += b.z; v[3] += b.w;
              }
              if (kRes) {
                const uint4 u = lds_128(box1 + pos);
                v[0] += __uint_as_float(u.x); v[1] += __uint_as_float(u.y); v[2] += __uint_as_float(u.z); v[3] += __uint_as_float(u.w);
              }
              sts_128(obox + pos, __float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3]));
            } else {
              float v[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(z[j >> 2][(j & 3) * 8 + e]);
              if (add_bias && n0 + j * 8 < (int)ep.N) {
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(ep.bias + n0 + j * 8));
                const float4 b1 = __ldg(reinterpret_cast<const float4*>(ep.bias + n0 + j * 8 + 4));
                v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w; v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
              }
              float d[8];
              if (kGelu) {
#pragma unroll
                for (int e = 0; e < 8; ++e) gelu_both(v[e], v[e], d[e]);
              } else if (kMul) {
                const uint4 u = lds_128(box1 + pos);
                const float2 a0 = unpack_bf16x2(u.x), a1 = unpack_bf16x2(u.y), a2 = unpack_bf16x2(u.z), a3 = unpack_bf16x2(u.w);
                v[0] *= a0.x; v[1] *= a0.y; v[2] *= a1.x; v[3] *= a1.y; v[4] *= a2.x; v[5] *= a2.y; v[6] *= a3.x; v[7] *= a3.y;
              }
              sts_128(obox + pos, pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
              if (kAux && kGelu) sts_128(box1 + pos, pack_bf16x2(d[0], d[1]), pack_bf16x2(d[2], d[3]), pack_bf16x2(d[4], d[5]), pack_bf16x2(d[6], d[7]));
            }
          }
          fence_async_smem();                          // generic-proxy writes -> visible to the TMA engine
        }
        __syncwarp();                                  // (also: every lane has consumed the operand box)
        if (lane == 0) {
          if (active) {
            if (kRed) tma_reduce_add_2d(tmap_c, obox, n0, m0); else tma_store_2d(tmap_c, obox, n0, m0);
            if (kAux) tma_store_2d(tmap_x, box1, n0, m0);
          }
          bulk_commit();
          if (kLoad && live && g + 1 < NGRP && n0 + GW < (int)ep.N) {     // next group's operand box
            mbar_expect_tx(aux_bar(ew, 0), kTmaBox);
            tma_load_2d(box1, tmap_x, aux_bar(ew, 0), n0 + GW, m0);
          }
        }
      }
    }
    if (lane == 0) bulk_wait_read<0>();                // the boxes must outlive the stores that read them
    __syncwarp();
  } else {
    // ===================== epilogue warps (TMEM lane quadrant = warp % 4) =====================
    // Per 16-column chunk: TMEM -> registers (one row per lane) -> per-warp smem staging -> "4 lanes per
    // row" so that every global access is a fully used 32/64-byte row segment.  The global operands of
    // the fused epilogue (bias, residual, pre-activation) for chunk c+1 are loaded into registers while
    // chunk c is processed (and for the first chunk before the accumulator is even ready), so their
    // L2/HBM latency is off the critical path of this 10-warp, low-occupancy CTA.
    const int quad = warp & 3;                 // TMEM lanes [32*quad, 32*quad+32) are the only ones this warp may read
    const int half = (warp - 2) >> 2;          // which slice of the tile's columns (EW/4 slices)
    constexpr int SLICE = BN / (EW / 4);       // columns per epilogue warp
    float* stg = reinterpret_cast<float*>(smem_raw + (stage_base - smem_u32(smem_raw))) + (warp - 2) * 32 * kStagePitch;
    const int rr = lane >> 2, cc = (lane & 3) * 4;   // coalesced phase: 4 lanes per row, 8 rows per iteration
    constexpr int NCHUNK = SLICE / kStageCols;
    using F = EpiSel<EPI>;
    struct Pre { float4 bias; float4 res[4]; uint2 aux[4]; };
    int local = 0;
    for (int t = cluster_id; t < num_tiles; t += num_clusters) {
      int p, tl, m_blk, n_blk, sp, kb0, kb1;
      locate(t, p, tl);
      gp.ts[p].decode(tl, m_blk, n_blk, sp, kb0, kb1);
      if (kb1 <= kb0) continue;
      const EpiParams& ep = gp.ep[p];
      const bool has_bias = F::bias(ep), has_res = F::res(ep), has_auxin = F::dgelu(ep);
      const int acc = local % NACC;
      const uint32_t acc_phase = (local / NACC) & 1u;
      ++local;
      const int64_t m_base = (int64_t)m_blk * (BM * CG) + cta_rank * BM + quad * 32;
      // per-tile row bookkeeping for the 4 rows this lane touches in the coalesced phase
      int64_t out_off[4], res_off[4], aux_off[4];
      bool rvalid[4];
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int64_t m = m_base + it * 8 + rr;
        rvalid[it] = m < ep.M;
        const int64_t mm = rvalid[it] ? m : 0;
        int64_t orow = mm, rrow = mm;
        if (F::remap(ep)) {
          if (ep.g > 0) orow = (mm / ep.g) * (int64_t)ep.G + ep.off + (mm % ep.g);
          rrow = (has_res && ep.res_idx) ? ep.res_idx[mm] : orow;
        }
        out_off[it] = orow * ep.ldo;
        res_off[it] = has_res ? rrow * ep.ldres : 0;
        aux_off[it] = (F::aux(ep) || has_auxin) ? mm * ep.ldaux : 0;
      }
      const bool add_bias = has_bias && sp == 0;
      auto prefetch = [&](int c, Pre& pr) {
        const int64_t n = (int64_t)n_blk * BN + half * SLICE + c * kStageCols + cc;
        const bool nv = n < ep.N;
        pr.bias = (add_bias && nv) ? *reinterpret_cast<const float4*>(ep.bias + n) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const bool ok = nv && rvalid[it];
          pr.res[it] = (has_res && ok) ? *reinterpret_cast<const float4*>(ep.res + res_off[it] + n) : make_float4(0.f, 0.f, 0.f, 0.f);
          pr.aux[it] = (has_auxin && ok) ? *reinterpret_cast<const uint2*>(ep.aux_in + aux_off[it] + n) : make_uint2(0u, 0u);
        }
      };
      Pre cur, nxt;
      prefetch(0, cur);                       // in flight while the MMAs of this tile still run
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      if (dbg && t == 0 && warp == 2 && lane == 0) dbg[4] = clock64();
      if (F::rowsum(ep) && n_blk == 0 && half == 0) {
        const float rs = tmem_ld_32x32b_x1(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(ROWSUM_COL + acc * 16));
        if (m_base + lane < ep.M) atomicAdd(ep.rowsum_out + m_base + lane, rs);
      }
      // One chunk: `use` holds this chunk's prefetched operands, `fill` receives the next chunk's.  The loop
      // below is unrolled by two with the buffers swapped, so no register copies (which would force a wait on
      // the in-flight loads) sit between chunks.
      auto process = [&](int c, const Pre& use, Pre& fill) {
        const int col0 = half * SLICE + c * kStageCols;
        float z[kStageCols];
        tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN + col0), z);
        if (c == NCHUNK - 1) {                 // accumulator fully read: hand the TMEM stage back to the MMA warp early
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (CG == 2 && !lead_cta) mbar_arrive_remote(tempty_bar(acc), 0);   // the leader CTA's MMA warp owns the hand-shake
            else mbar_arrive(tempty_bar(acc));
          }
        }
#pragma unroll
        for (int j = 0; j < kStageCols; j += 4)
          *reinterpret_cast<float4*>(stg + lane * kStagePitch + j) = make_float4(z[j], z[j + 1], z[j + 2], z[j + 3]);
        __syncwarp();
        if (c + 1 < NCHUNK) prefetch(c + 1, fill);
        const int64_t n = (int64_t)n_blk * BN + col0 + cc;
        if (n < ep.N) {
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            if (!rvalid[it]) continue;
            float4 v = *reinterpret_cast<const float4*>(stg + (it * 8 + rr) * kStagePitch + cc);
            if (has_bias) { v.x += use.bias.x; v.y += use.bias.y; v.z += use.bias.z; v.w += use.bias.w; }
            float4 ax = v;                       // aux_out receives z, or gelu'(z) next to a GELU
            if (F::gelu(ep)) {
              gelu_both(v.x, v.x, ax.x); gelu_both(v.y, v.y, ax.y); gelu_both(v.z, v.z, ax.z); gelu_both(v.w, v.w, ax.w);
            } else if (has_auxin) {
              const float2 lo = unpack_bf16x2(use.aux[it].x), hi = unpack_bf16x2(use.aux[it].y);
              v.x *= lo.x; v.y *= lo.y; v.z *= hi.x; v.w *= hi.y;
            }
            if (F::aux(ep)) {
              uint2 o;
              o.x = pack_bf16x2(ax.x, ax.y);
              o.y = pack_bf16x2(ax.z, ax.w);
              *reinterpret_cast<uint2*>(ep.aux_out + aux_off[it] + n) = o;
            }
            if (has_res) { v.x += use.res[it].x; v.y += use.res[it].y; v.z += use.res[it].z; v.w += use.res[it].w; }
            if (F::red(ep)) {
              float* o = reinterpret_cast<float*>(ep.out) + out_off[it] + n;
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
            } else if (F::bf16(ep)) {
              uint2 o;
              o.x = pack_bf16x2(v.x, v.y);
              o.y = pack_bf16x2(v.z, v.w);
              *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(ep.out) + out_off[it] + n) = o;
            } else {
              *reinterpret_cast<float4*>(reinterpret_cast<float*>(ep.out) + out_off[it] + n) = v;
            }
          }
        }
        __syncwarp();
      };
      static_assert(NCHUNK % 2 == 0, "chunk loop is unrolled by two");
#pragma unroll 1
      for (int c = 0; c < NCHUNK; c += 2) {
        process(c, cur, nxt);
        process(c + 1, nxt, cur);
      }
      if (dbg && t == 0 && warp == 2 && lane == 0) dbg[5] = clock64();
    }
  }

  __syncwarp();          // lanes 1..31 of the producer / MMA warps skipped the role loops: reconverge, so that the
                         // aligned barrier below is executed once per warp (a divergent bar.sync counts a warp twice)
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();   // the peer's smem / barriers / TMEM stay valid until both CTAs are done
  if (dbg && threadIdx.x == 0) dbg[6] = clock64();
  if (warp == 1) {
    tc_fence_after();
    if (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------
// host side: TMA descriptor cache + launch
// ------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  });
  return fn;
}

struct MapKey {
  const void* ptr; int64_t inner, outer, ld; int box_outer;     // box_outer < 0: f32 tensor, box {32, -box_outer}
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && inner == o.inner && outer == o.outer && ld == o.ld && box_outer == o.box_outer;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = (size_t)k.ptr;
    h = h * 1000003u ^ (size_t)k.inner;
    h = h * 1000003u ^ (size_t)k.outer;
    h = h * 1000003u ^ (size_t)k.ld;
    h = h * 1000003u ^ (size_t)k.box_outer;
    return h;
  }
};

// 2-D bf16 tensor [outer][inner] with row pitch ld elements; box = {64, box_outer}, 128B swizzle,
// out-of-bounds elements read as zero (M/N/K tails need no special casing in the kernel).
static int get_tensor_map(const void* ptr, int64_t inner, int64_t outer, int64_t ld, int box_outer, CUtensorMap* out) {
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  static std::mutex mu;
  MapKey key{ptr, inner, outer, ld, box_outer};
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return DAVF_OK; }
  }
  PFN_encodeTiled enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return DAVF_ECUDA; }
  const bool f32 = box_outer < 0;                  // f32 epilogue tensors: 32 columns = 128 bytes per box row
  cuuint64_t gdim[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * (f32 ? 4 : 2)};
  cuuint32_t box[2] = {f32 ? 32u : 64u, (cuuint32_t)(f32 ? -box_outer : box_outer)};
  cuuint32_t estr[2] = {1u, 1u};
  CUtensorMap m;
  CUresult r = enc(&m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) ptr=%p inner=%lld outer=%lld ld=%lld box=%d", (int)r, ptr, (long long)inner,
              (long long)outer, (long long)ld, box_outer);
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

// SMs the persistent GEMM grids are sized for.  Data-parallel runs leave a few SMs to the NCCL kernels that run under
// backward: a 148-CTA persistent launch that finds some SMs taken runs its last CTAs as a second wave (davf_set_gemm_sms).
static std::atomic<int> g_gemm_sms{[] { const char* e = getenv("DAVF_GEMM_SMS"); const int n = e ? atoi(e) : kNumSMs; return n < 2 ? 2 : (n > kNumSMs ? kNumSMs : (n & ~1)); }()};
static inline int num_sms() { return g_gemm_sms.load(); }

template <int BN, int STAGES_, bool AK, bool BKM, int EPI, int CG, int NG, int EW>
static int launch_group_ew(const GemmGroup<NG>& gp, cudaStream_t st) {
  // 16 epilogue warps trade one pipeline stage for their staging (staged epilogues only)
  constexpr int STAGES = fit_stages(STAGES_, EPI, EW, BN, CG);
  static_assert(STAGES >= 3, "pipeline too shallow");
  constexpr int kNumThreads = 64 + 32 * EW;
  constexpr size_t smem = (size_t)STAGES * (BM * BK * 2 + (BN / CG) * BK * 2) + staging_bytes(EPI, EW, BN, CG) + ones_bytes(EPI) + 8 * (2 * STAGES + 5 + 3 * EW) + 16 + 1024;
  static_assert(smem <= 232448, "exceeds the 227 KB shared memory of an SM");
  static bool attr_set = false;
  auto kern = gemm_tc_kernel<BN, STAGES, AK, BKM, EPI, CG, EW, NG>;
  if (!attr_set) {
    DAVF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  const int tiles = gp.tile_end[NG - 1];
  int clusters = tiles < num_sms() / CG ? tiles : num_sms() / CG;
  if (NG == 1 && gp.ts[0].sk_units > 0) clusters = gp.ts[0].sk_units;      // stream-K: unit u and its second segment belong to cluster u
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(clusters * CG);
  cfg.blockDim = dim3(kNumThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (CG > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = CG;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (g_pdl.load(std::memory_order_relaxed)) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  DAVF_CUDA(cudaLaunchKernelEx(&cfg, kern, gp));
  g_launches.fetch_add(1);
  if (CG == 2) g_launch_kind[kKindGemm2Cta].fetch_add(1);
  return DAVF_OK;
}

template <int BN, int STAGES_, bool AK, bool BKM, int EPI, int CG, int NG = 1>
static int launch_group(const GemmGroup<NG>& gp, cudaStream_t st) {
  // GELU / x gelu' epilogues of the CTA-pair kernel: 16 epilogue warps (DAVF_HEAVY_EW=8 selects 8, for experiments)
  constexpr bool kHeavy = EPI >= 0 && (EPI & (EPI_GELU | EPI_DGELU)) != 0 && CG == 2 && !direct_epi(EPI);
  if constexpr (kHeavy) {
    static const int ew = [] { const char* e = getenv("DAVF_HEAVY_EW"); return e ? atoi(e) : 16; }();
    if (ew == 16) return launch_group_ew<BN, STAGES_, AK, BKM, EPI, CG, NG, 16>(gp, st);
  }
  return launch_group_ew<BN, STAGES_, AK, BKM, EPI, CG, NG, 8>(gp, st);
}

template <int BN, int STAGES_, bool AK, bool BKM, int EPI, int CG>
static int launch_cfg(const CUtensorMap& ta, const CUtensorMap& tb, const TileSched& ts, const EpiParams& ep, cudaStream_t st) {
  GemmGroup<1> gp;
  gp.ta[0] = ta; gp.tb[0] = tb; gp.ep[0] = ep; gp.ts[0] = ts;
  gp.tc[0] = ta; gp.taux[0] = ta;
  if constexpr (direct_epi(EPI)) {           // epilogue tensors as [M][N] TMA tensors, box = 32 rows x 128 bytes
    constexpr bool kF32 = (EPI & EPI_BF16) == 0;
    int rc = get_tensor_map(ep.out, ep.N, ep.M, ep.ldo, kF32 ? -32 : 32, &gp.tc[0]);
    if (rc) return rc;
    if constexpr ((EPI & EPI_RES) != 0) {
      rc = get_tensor_map(ep.res, ep.N, ep.M, ep.ldres, -32, &gp.taux[0]);
    } else if constexpr ((EPI & (EPI_AUX | EPI_DGELU)) != 0) {
      const void* aux = (EPI & EPI_AUX) ? (const void*)ep.aux_out : (const void*)ep.aux_in;
      rc = get_tensor_map(aux, ep.N, ep.M, ep.ldaux, 32, &gp.taux[0]);
    }
    if (rc) return rc;
  }
  gp.tile_end[0] = ts.num_tiles();
  return launch_group<BN, STAGES_, AK, BKM, EPI, CG, 1>(gp, st);
}

// run-time epilogue mask of a launch, or -1 if it needs the generic kernel
static bool direct_ok(const davf_gemm_args& a);
static int epi_mask(const davf_gemm_args& a) {
  if (a.g > 0 || a.res_idx || a.debug_clocks) return -1;
  if (!direct_ok(a)) return -1;
  int m = 0;
  if (a.rowsum_out) m |= EPI_ROWSUM;
  if (a.bias) m |= EPI_BIAS;
  if (a.act == DAVF_ACT_GELU) m |= EPI_GELU;
  if (a.act == DAVF_ACT_DGELU) m |= EPI_DGELU;
  if (a.aux_out) m |= EPI_AUX;
  if (a.res) m |= EPI_RES;
  if (a.out_bf16) m |= EPI_BF16;
  if (a.accumulate) m |= EPI_RED;
  return m;
}

// the TMA-store epilogue of the bf16-output kernels needs TMA-legal output tensors (16-byte aligned base and row pitch)
static bool direct_ok(const davf_gemm_args& a) {
  auto al = [](const void* p, int64_t ld, int esz) { return ((uintptr_t)p & 15) == 0 && (ld * esz) % 16 == 0; };
  return al(a.out, a.ldo, a.out_bf16 ? 2 : 4) && (!a.aux_out || al(a.aux_out, a.ldaux, 2)) && (!a.aux_in || al(a.aux_in, a.ldaux, 2)) &&
         (!a.res || al(a.res, a.ldres, 4));
}

static bool has_static_epi(const davf_gemm_args& a) {
  const int em = epi_mask(a);
  if (em < 0) return false;
  if (a.a_kmajor && a.b_kmajor) return em == (EPI_BIAS | EPI_BF16) || em == (EPI_BIAS | EPI_GELU | EPI_AUX | EPI_BF16) || em == (EPI_BIAS | EPI_RES);
  if (a.a_kmajor && !a.b_kmajor) return em == EPI_BF16 || em == (EPI_DGELU | EPI_BF16);
  if (!a.a_kmajor && !a.b_kmajor) return em == EPI_RED || em == (EPI_RED | EPI_ROWSUM);
  return false;
}

template <int BN, int STAGES, int CG>
static int launch_major(const davf_gemm_args& a, const CUtensorMap& ta, const CUtensorMap& tb, const TileSched& ts, cudaStream_t st) {
  const EpiParams ep = make_epi(a);
  const int em = epi_mask(a);
  if (a.a_kmajor && a.b_kmajor) {            // forward
    if (em == (EPI_BIAS | EPI_BF16)) return launch_cfg<BN, STAGES, true, true, EPI_BIAS | EPI_BF16, CG>(ta, tb, ts, ep, st);
    if (em == (EPI_BIAS | EPI_GELU | EPI_AUX | EPI_BF16)) return launch_cfg<BN, STAGES, true, true, EPI_BIAS | EPI_GELU | EPI_AUX | EPI_BF16, CG>(ta, tb, ts, ep, st);
    if (em == (EPI_BIAS | EPI_RES)) return launch_cfg<BN, STAGES, true, true, EPI_BIAS | EPI_RES, CG>(ta, tb, ts, ep, st);
    if constexpr (CG == 1) return launch_cfg<BN, STAGES, true, true, -1, 1>(ta, tb, ts, ep, st);
  } else if (a.a_kmajor && !a.b_kmajor) {    // dgrad
    if (em == EPI_BF16) return launch_cfg<BN, STAGES, true, false, EPI_BF16, CG>(ta, tb, ts, ep, st);
    if (em == (EPI_DGELU | EPI_BF16)) return launch_cfg<BN, STAGES, true, false, EPI_DGELU | EPI_BF16, CG>(ta, tb, ts, ep, st);
    if constexpr (CG == 1) return launch_cfg<BN, STAGES, true, false, -1, 1>(ta, tb, ts, ep, st);
  } else if (!a.a_kmajor && !a.b_kmajor) {   // wgrad, with or without the bias-gradient row sums
    if (em == EPI_RED) return launch_cfg<BN, STAGES, false, false, EPI_RED, CG>(ta, tb, ts, ep, st);
    if constexpr (CG == 2 || BN == 128) {
      if (em == (EPI_RED | EPI_ROWSUM)) return launch_cfg<BN, STAGES, false, false, EPI_RED | EPI_ROWSUM, CG>(ta, tb, ts, ep, st);
    }
    if constexpr (CG == 1) return launch_cfg<BN, STAGES, false, false, -1, 1>(ta, tb, ts, ep, st);
  } else {
    if constexpr (CG == 1) return launch_cfg<BN, STAGES, false, true, -1, 1>(ta, tb, ts, ep, st);
  }
  set_error("gemm: no kernel for this configuration");
  return DAVF_EUNSUPPORTED;
}

static std::atomic<int> g_allow_2cta{1};

int gemm_tc_launch(const davf_gemm_args& a, cudaStream_t st) {
  const int kb_total = (int)((a.K + BK - 1) / BK);
  auto pick_splits = [&](int64_t mn_tiles, int units) {
    int splits = a.split_k;
    if (splits <= 0) {   // auto: fill the machine when the caller allows atomic accumulation
      splits = 1;
      if (a.accumulate) {      // split K so that the tiles fill (at most) one wave of CTAs / CTA pairs
        if (mn_tiles < units) splits = (int)(units / mn_tiles);
        if (splits > kb_total / 4) splits = kb_total / 4 > 0 ? kb_total / 4 : 1;
      }
    }
    if (splits > kb_total) splits = kb_total;
    if (splits < 1) splits = 1;
    const int per = (kb_total + splits - 1) / splits;
    return std::make_pair((kb_total + per - 1) / per, per);      // no empty split
  };
  // Stream-K for accumulating (wgrad) launches whose tiles do not fill the machine: equal K-block ranges per CTA (pair)
  // instead of whole splits (DAVF_STREAMK=0 turns it off).  Needs a range no longer than one tile's K loop.
  static const int streamk_on = [] { const char* e = getenv("DAVF_STREAMK"); return e ? atoi(e) : 1; }();
  auto make_sched = [&](int64_t m_t, int64_t n_t, int units) {
    const auto sp = pick_splits(m_t * n_t, units);
    TileSched ts{(int)m_t, (int)n_t, sp.first, kb_total, sp.second, 0, 0};
    const int64_t mn = m_t * n_t, total = mn * kb_total;
    if (streamk_on && a.accumulate && a.split_k <= 0 && mn < units && total >= (int64_t)units * 4) {
      int len = (int)((total + units - 1) / units);
      if (len < 4) len = 4;
      // cost in k-block units: a CTA (pair) runs its K range plus up to two f32 reduce epilogues against
      // rounds x (split length + one epilogue) of the classic decomposition.  Measured on B200 (gemm_shapes_bench with
      // DAVF_STREAMK=2): the reduce epilogue of a 256 x 256 tile costs about as much as 12 k-blocks, so for this model's
      // wgrad shapes the balanced ranges rarely pay for the second epilogue.
      constexpr int kEpi = 12;
      const int64_t rounds = (mn * sp.first + units - 1) / units;
      if (len <= kb_total && (streamk_on == 2 || len + 2 * kEpi < rounds * (sp.second + kEpi))) {     // DAVF_STREAMK=2: whenever legal (tests)
        ts.sk_len = len;
        ts.sk_units = (int)((total + len - 1) / len);
      }
    }
    return ts;
  };
  CUtensorMap ta, tb;
  int rc;
  // ---- CTA-pair path: 256 x 256 tiles, tcgen05.mma.cta_group::2 -----------------------------------
  static const int class_mask = [] { const char* e = getenv("DAVF_2CTA_CLASSES"); return e ? atoi(e) : 7; }();   // debug: 1 fwd, 2 dgrad, 4 wgrad
  const int cls = (a.a_kmajor && a.b_kmajor) ? 1 : (a.a_kmajor ? 2 : 4);
  // wgrad + bias-gradient row sums: a 256-wide tile leaves no TMEM for a second accumulator stage next to the 16 row-sum
  // columns (2 x 256 + 16 > 512), so its epilogue is not overlapped; 256 x 128 pair tiles are double-buffered
  // (DAVF_WGRAD_ROWSUM_BN = 128 | 256, measured in tools/gemm_shapes_bench.py)
  static const int rowsum_bn = [] { const char* e = getenv("DAVF_WGRAD_ROWSUM_BN"); return e ? atoi(e) : 256; }();
  if (g_allow_2cta.load() && (class_mask & cls) && cls == 4 && rowsum_bn == 128 && a.rowsum_out && has_static_epi(a) && a.M >= 256 && a.N >= 128) {
    const int64_t m2 = (a.M + 2 * BM - 1) / (2 * BM), n2 = (a.N + 127) / 128;
    const TileSched ts = make_sched(m2, n2, num_sms() / 2);
    if (ts.sk_units > 0 || m2 * n2 * ts.splits >= 36) {
      rc = get_tensor_map(a.a, a.M, a.K, a.lda, BK, &ta);
      if (rc) return rc;
      rc = get_tensor_map(a.b, a.N, a.K, a.ldb, BK, &tb);
      if (rc) return rc;
      return launch_cfg<128, 6, false, false, EPI_RED | EPI_ROWSUM, 2>(ta, tb, ts, make_epi(a), st);
    }
  }
  if (g_allow_2cta.load() && (class_mask & cls) && has_static_epi(a) && a.M >= 256 && a.N >= 256) {
    const int64_t m2 = (a.M + 2 * BM - 1) / (2 * BM), n2 = (a.N + 255) / 256;
    const TileSched ts = make_sched(m2, n2, num_sms() / 2);
    // worth it when the pair tiles fill at least ~half of the 74 SM pairs; otherwise 128-wide 1-CTA tiles
    // give more parallelism
    if (ts.sk_units >= 36 || m2 * n2 * ts.splits >= 36) {
      if (a.a_kmajor) rc = get_tensor_map(a.a, a.K, a.M, a.lda, BM, &ta);
      else rc = get_tensor_map(a.a, a.M, a.K, a.lda, BK, &ta);
      if (rc) return rc;
      if (a.b_kmajor) rc = get_tensor_map(a.b, a.K, a.N, a.ldb, 128, &tb);
      else rc = get_tensor_map(a.b, a.N, a.K, a.ldb, BK, &tb);
      if (rc) return rc;
      return launch_major<256, 6, 2>(a, ta, tb, ts, st);
    }
  }
  // ---- single-CTA path --------------------------------------------------------------------------------
  // tile-N choice: 256-wide tiles halve the shared-memory operand traffic per FLOP; use them when
  // the problem still yields at least ~one full wave of CTAs, otherwise 128 for parallelism.
  const int64_t m_tiles = (a.M + BM - 1) / BM;
  int bn = 128;
  if (a.N % 256 == 0 && m_tiles * (a.N / 256) >= num_sms() && !a.rowsum_out) bn = 256;   // row-sum columns need BN = 128
  const int64_t n_tiles = (a.N + bn - 1) / bn;
  const TileSched ts = make_sched(m_tiles, n_tiles, num_sms());
  if (a.a_kmajor) rc = get_tensor_map(a.a, a.K, a.M, a.lda, BM, &ta);
  else rc = get_tensor_map(a.a, a.M, a.K, a.lda, BK, &ta);
  if (rc) return rc;
  if (a.b_kmajor) rc = get_tensor_map(a.b, a.K, a.N, a.ldb, bn, &tb);
  else rc = get_tensor_map(a.b, a.N, a.K, a.ldb, BK, &tb);
  if (rc) return rc;
  if (bn == 256) return launch_major<256, 4, 1>(a, ta, tb, ts, st);
  return launch_major<128, 6, 1>(a, ta, tb, ts, st);
}

// ---- grouped launch: up to kMaxGroup problems of ONE operand-layout class in one launch (1-CTA 128 x 128 tiles,
// run-time-flag epilogue).  Accumulating problems (wgrad) are split along K so that the group fills about one wave.
int gemm_tc_launch_grouped(const davf_gemm_args* a, int count, cudaStream_t st) {
  GemmGroup<kMaxGroup> gp;
  memset(&gp, 0, sizeof(gp));
  const bool ak = a[0].a_kmajor != 0, bk = a[0].b_kmajor != 0;
  // A group whose members all have one of the simple epilogues runs the compile-time (TMA) epilogue kernel:
  //   forward  bias (or none) -> bf16      dgrad  -> bf16      wgrad  f32 reduce (+ row sums where requested)
  int gmask = -2;
  for (int p = 0; p < count; ++p) {
    int em = epi_mask(a[p]);
    if (em >= 0) {
      if (em == EPI_BF16 && ak && bk) em = EPI_BIAS | EPI_BF16;            // bias pointer is checked at run time
      if (em == EPI_RED) em = EPI_RED | EPI_ROWSUM;                          // rowsum_out pointer is checked at run time
      const bool ok = (ak && bk && em == (EPI_BIAS | EPI_BF16)) || (ak && !bk && em == EPI_BF16) || (!ak && !bk && em == (EPI_RED | EPI_ROWSUM));
      if (!ok) em = -1;
    }
    gmask = (gmask == -2 || gmask == em) ? em : -1;
  }
  // wgrad groups of large problems (a ViT block's Linears) use CTA-pair 256 x 256 tiles, everything else 128 x 128
  int64_t pair_tiles = 0, mn_total = 0;
  for (int p = 0; p < count; ++p) {
    pair_tiles += ((a[p].M + 255) / 256) * ((a[p].N + 255) / 256);
    mn_total += ((a[p].M + BM - 1) / BM) * ((a[p].N + 127) / 128);
  }
  bool pair = gmask == (EPI_RED | EPI_ROWSUM) && g_allow_2cta.load() && pair_tiles >= 8;
  for (int p = 0; p < count && pair; ++p) pair = a[p].M >= 256 && a[p].N >= 256;
  const int tm = pair ? 256 : 128, tn = pair ? 256 : 128;
  const int64_t units = pair ? num_sms() / 2 : num_sms(), tiles_total = pair ? pair_tiles : mn_total;
  int end = 0, rc;
  for (int p = 0; p < kMaxGroup; ++p) {
    if (p >= count) {                 // unused slot: no tiles; keep valid (never dereferenced) descriptors
      gp.ta[p] = gp.ta[0]; gp.tb[p] = gp.tb[0]; gp.tc[p] = gp.tc[0]; gp.taux[p] = gp.taux[0]; gp.ep[p] = gp.ep[0]; gp.ts[p] = gp.ts[0];
      gp.tile_end[p] = end;
      continue;
    }
    const davf_gemm_args& g = a[p];
    const int kb_total = (int)((g.K + BK - 1) / BK);
    const int m_tiles = (int)((g.M + tm - 1) / tm), n_tiles = (int)((g.N + tn - 1) / tn);
    int splits = g.split_k > 0 ? g.split_k : 1;
    if (g.split_k <= 0 && g.accumulate && tiles_total < units) {
      splits = (int)(units / tiles_total);
      if (splits > kb_total / 4) splits = kb_total / 4 > 0 ? kb_total / 4 : 1;
    }
    if (splits > kb_total) splits = kb_total;
    if (splits < 1) splits = 1;
    const int per = (kb_total + splits - 1) / splits;
    splits = (kb_total + per - 1) / per;
    gp.ts[p] = TileSched{m_tiles, n_tiles, splits, kb_total, per, 0, 0};
    if (g.a_kmajor) rc = get_tensor_map(g.a, g.K, g.M, g.lda, BM, &gp.ta[p]);
    else rc = get_tensor_map(g.a, g.M, g.K, g.lda, BK, &gp.ta[p]);
    if (rc) return rc;
    if (g.b_kmajor) rc = get_tensor_map(g.b, g.K, g.N, g.ldb, 128, &gp.tb[p]);
    else rc = get_tensor_map(g.b, g.N, g.K, g.ldb, BK, &gp.tb[p]);
    if (rc) return rc;
    gp.ep[p] = make_epi(g);
    gp.ep[p].debug_clocks = nullptr;
    gp.tc[p] = gp.ta[p]; gp.taux[p] = gp.ta[p];
    if (gmask >= 0) {
      rc = get_tensor_map(g.out, g.N, g.M, g.ldo, g.out_bf16 ? 32 : -32, &gp.tc[p]);
      if (rc) return rc;
    }
    end += m_tiles * n_tiles * splits;
    gp.tile_end[p] = end;
  }
  if (gmask == (EPI_BIAS | EPI_BF16)) return launch_group<128, 6, true, true, EPI_BIAS | EPI_BF16, 1, kMaxGroup>(gp, st);
  if (gmask == EPI_BF16) return launch_group<128, 6, true, false, EPI_BF16, 1, kMaxGroup>(gp, st);
  if (gmask == (EPI_RED | EPI_ROWSUM)) {
    if (pair) return launch_group<256, 6, false, false, EPI_RED | EPI_ROWSUM, 2, kMaxGroup>(gp, st);
    return launch_group<128, 6, false, false, EPI_RED | EPI_ROWSUM, 1, kMaxGroup>(gp, st);
  }
  if (ak && bk) return launch_group<128, 6, true, true, -1, 1, kMaxGroup>(gp, st);
  if (ak && !bk) return launch_group<128, 6, true, false, -1, 1, kMaxGroup>(gp, st);
  if (!ak && !bk) return launch_group<128, 6, false, false, -1, 1, kMaxGroup>(gp, st);
  return launch_group<128, 6, false, true, -1, 1, kMaxGroup>(gp, st);
}

int gemm_set_sms(int n) { g_gemm_sms.store(n < 2 ? 2 : (n > kNumSMs ? kNumSMs : (n & ~1))); return g_gemm_sms.load(); }
int gemm_set_2cta(int on) { g_allow_2cta.store(on ? 1 : 0); return 0; }

}  // namespace davf
