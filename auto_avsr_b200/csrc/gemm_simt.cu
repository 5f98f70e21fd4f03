// fp32 CUDA-core GEMM with the fused epilogue family -- the on-device EXACT reference path
// (AVSR_PREC_FP32).  acc[m][n] = sum_k A[m][k] * Bw[n][k], both operands K-major (torch.nn.Linear).
// Classic 128x128x16 shared-memory tiling, 8x8 register micro-tile per thread, float4 everywhere.
// It exists so that parity failures of the tensor-core path can be split into "algorithm" and "TF32
// rounding" on the GPU itself; it is not the product path.
#include "common.cuh"

namespace avsr {

constexpr int SB_M = 128, SB_N = 128, SB_K = 16;

template <int MODE>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const float* __restrict__ A, const float* __restrict__ Bw,
                                                        int M, int N, int K, EpiParams ep) {
  __shared__ float As[SB_K][SB_M + 4];
  __shared__ float Bs[SB_K][SB_N + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * SB_M, n0 = blockIdx.x * SB_N;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  // each thread stages 2 float4 of A and 2 of B per k-block: row = tid/4 (+64), k-chunk = tid%4
  const int lr = tid >> 2, lk = (tid & 3) * 4;
  for (int k0 = 0; k0 < K; k0 += SB_K) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = lr + h * 64;
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
      if (m0 + r < M) a = *reinterpret_cast<const float4*>(A + (long)(m0 + r) * K + k0 + lk);
      if (n0 + r < N) b = *reinterpret_cast<const float4*>(Bw + (long)(n0 + r) * K + k0 + lk);
      As[lk + 0][r] = a.x; As[lk + 1][r] = a.y; As[lk + 2][r] = a.z; As[lk + 3][r] = a.w;
      Bs[lk + 0][r] = b.x; Bs[lk + 1][r] = b.y; Bs[lk + 2][r] = b.z; Bs[lk + 3][r] = b.w;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SB_K; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if constexpr (MODE == EPI_GLU) {
      // tile columns [0,64) = value channels, [64,128) = their gates (interleaved weight layout)
#pragma unroll
      for (int j = 0; j < 4; ++j) epi_store_glu(ep, m, n0 + tx * 4 + j, acc[i][j], acc[i][j + 4]);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
        epi_store<MODE>(ep, m, n, acc[i][j]);
      }
    }
  }
}

int gemm_simt(int mode, const float* A, const float* Bw, int M, int N, int K, const EpiParams& ep, cudaStream_t st) {
  AVSR_REQUIRE(K > 0 && K % SB_K == 0, "gemm_simt: K=%d must be a multiple of %d", K, SB_K);
  AVSR_REQUIRE(mode != EPI_GLU || N % 128 == 0, "gemm_simt: GLU needs N %% 128 == 0 (N=%d)", N);
  if (M <= 0 || N <= 0) return AVSR_OK;
  dim3 grid(cdiv(N, SB_N), cdiv(M, SB_M));
  switch (mode) {
    case EPI_LINEAR: gemm_simt_kernel<EPI_LINEAR><<<grid, 256, 0, st>>>(A, Bw, M, N, K, ep); break;
    case EPI_QK: gemm_simt_kernel<EPI_QK><<<grid, 256, 0, st>>>(A, Bw, M, N, K, ep); break;
    case EPI_VT: gemm_simt_kernel<EPI_VT><<<grid, 256, 0, st>>>(A, Bw, M, N, K, ep); break;
    case EPI_GLU: gemm_simt_kernel<EPI_GLU><<<grid, 256, 0, st>>>(A, Bw, M, N, K, ep); break;
    case EPI_POS: gemm_simt_kernel<EPI_POS><<<grid, 256, 0, st>>>(A, Bw, M, N, K, ep); break;
    default: AVSR_REQUIRE(false, "gemm_simt: bad epilogue mode %d", mode);
  }
  AVSR_CHECK_LAUNCH();
  return AVSR_OK;
}

}  // namespace avsr
