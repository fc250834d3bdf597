// SIMT (CUDA-core) GEMM with the same argument struct and epilogue as the tcgen05 kernel.
// It is the on-device CHECKER for the tensor-core kernel (tests compare the two on the GPU) and a
// (tests/check/libdavf_check.so; test infrastructure -- the product library does not contain it).
#include "gemm_epilogue.cuh"

namespace davf {

constexpr int ST = 64;   // tile M = N
constexpr int SK = 16;   // tile K

__global__ void __launch_bounds__(256) gemm_simt_kernel(const uint16_t* __restrict__ A, int64_t lda, int a_kmajor,
                                                        const uint16_t* __restrict__ B, int64_t ldb, int b_kmajor,
                                                        int64_t M, int64_t N, int64_t K, int64_t k_chunk, EpiParams ep) {
  __shared__ float As[SK][ST + 1];
  __shared__ float Bs[SK][ST + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int64_t m0 = (int64_t)blockIdx.y * ST, n0 = (int64_t)blockIdx.x * ST;
  const int64_t kb = (int64_t)blockIdx.z * k_chunk;
  int64_t ke = kb + k_chunk;
  if (ke > K) ke = K;
  float acc[4][4] = {};
  float rs[4] = {};
  for (int64_t k0 = kb; k0 < ke; k0 += SK) {
    for (int i = threadIdx.x; i < ST * SK; i += 256) {
      int mm, kk;
      if (a_kmajor) { mm = i / SK; kk = i % SK; } else { kk = i / ST; mm = i % ST; }
      const int64_t m = m0 + mm, k = k0 + kk;
      float v = 0.f;
      if (m < M && k < ke) v = bf16_to_f32(a_kmajor ? A[m * lda + k] : A[k * lda + m]);
      As[kk][mm] = v;
      int nn;
      if (b_kmajor) { nn = i / SK; kk = i % SK; } else { kk = i / ST; nn = i % ST; }
      const int64_t n = n0 + nn, k2 = k0 + kk;
      v = 0.f;
      if (n < N && k2 < ke) v = bf16_to_f32(b_kmajor ? B[n * ldb + k2] : B[k2 * ldb + n]);
      Bs[kk][nn] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[kk][ty * 4 + i]; b[i] = Bs[kk][tx * 4 + i]; rs[i] += a[i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) epilogue_row<4>(ep, m0 + ty * 4 + i, n0 + tx * 4, acc[i], blockIdx.z == 0);
  if (ep.rowsum_out && blockIdx.x == 0 && tx == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (m0 + ty * 4 + i < M) atomicAdd(ep.rowsum_out + m0 + ty * 4 + i, rs[i]);
  }
}

int gemm_simt_launch(const davf_gemm_args& a, cudaStream_t st) {
  const int splits = a.split_k > 1 ? a.split_k : 1;
  int64_t k_chunk = (a.K + splits - 1) / splits;
  k_chunk = (k_chunk + SK - 1) / SK * SK;
  dim3 grid((unsigned)((a.N + ST - 1) / ST), (unsigned)((a.M + ST - 1) / ST), (unsigned)splits);
  gemm_simt_kernel<<<grid, 256, 0, st>>>(a.a, a.lda, a.a_kmajor, a.b, a.ldb, a.b_kmajor, a.M, a.N, a.K, k_chunk, make_epi(a));
  DAVF_LAUNCH_OK();
  return DAVF_OK;
}

}  // namespace davf
