// K2: MAE random masking by rank counting (replaces avmae.py:127-140).
//   ids_restore[i] = #{ j : n_j < n_i  or (n_j == n_i and j < i) }      (stable rank)
//   ids_keep[rank] = i for rank < len_keep ;  mask[i] = rank >= len_keep
// One CTA per sample row; the row lives in shared memory; O(L^2) compares (L <= 196 here).
#include "common.cuh"

namespace davf {
__global__ void mask_rank_kernel(const float* __restrict__ noise, int L, int len_keep,
                                 int64_t* __restrict__ ids_restore, int64_t* __restrict__ ids_keep,
                                 float* __restrict__ mask) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float row[];
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < L; i += blockDim.x) row[i] = noise[(int64_t)b * L + i];
  __syncthreads();
  for (int i = threadIdx.x; i < L; i += blockDim.x) {
    const float ni = row[i];
    int rank = 0;
    for (int j = 0; j < L; ++j) {
      const float nj = row[j];
      rank += (nj < ni) || (nj == ni && j < i);
    }
    ids_restore[(int64_t)b * L + i] = rank;
    if (rank < len_keep) ids_keep[(int64_t)b * len_keep + rank] = i;
    mask[(int64_t)b * L + i] = rank >= len_keep ? 1.0f : 0.0f;
  }
}
}  // namespace davf

extern "C" int davf_mask_rank(const float* noise, int B, int L, int len_keep, int64_t* ids_restore,
                              int64_t* ids_keep, float* mask, davf_stream_t s) {
  DAVF_CHECK_ARG(B >= 0 && L > 0 && len_keep >= 0 && len_keep <= L, "mask_rank: bad sizes B=%d L=%d keep=%d", B, L, len_keep);
  DAVF_CHECK_ARG(L <= 8192, "mask_rank: L=%d too large for one CTA", L);
  if (B == 0) return DAVF_OK;
  int threads = L < 256 ? ((L + 31) / 32) * 32 : 256;
  DAVF_CUDA(davf::launch_pdl(davf::mask_rank_kernel, dim3(B), dim3(threads), L * sizeof(float), davf::as_stream(s), noise, L, len_keep, ids_restore, ids_keep, mask));
  DAVF_LAUNCH_OK();
  return DAVF_OK;
}
