// CUDA-core GEMM  out = act(A W^T + bias + addmat)  with fp32 accumulation.
// fp32 validation mode runs every nn.Linear of the reference through this kernel (1e-5 parity gate);
// bf16 mode uses it only for the odd tiny-K layers (K = 12, 17, 20, 50, 68) that do not map on UMMA tiles.
#include "context.cuh"

namespace {

constexpr int TM = 64, TN = 64, TK = 16;

template <typename TA, bool kFast>
__global__ void __launch_bounds__(256) gemm_simt_kernel(GemmDesc g) {
  pdl_trigger();
  pdl_wait();
  __shared__ float As[TK][TM + 4];
  __shared__ float Ws[TK][TN + 4];
  const TA* __restrict__ A = reinterpret_cast<const TA*>(g.A);
  const TA* __restrict__ W = reinterpret_cast<const TA*>(g.W);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.x * TM, n0 = blockIdx.y * TN;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < g.K; k0 += TK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = threadIdx.x + i * 256;
      const int r = idx >> 4, c = idx & 15;
      const int k = k0 + c;
      float a = 0.f, w = 0.f;
      if (k < g.K) {
        if (m0 + r < g.M) a = to_f32(A[static_cast<size_t>(m0 + r) * g.lda + k]);
        if (n0 + r < g.N) w = to_f32(W[static_cast<size_t>(n0 + r) * g.ldw + k]);
      }
      As[c][r] = a;
      Ws[c][r] = w;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      float a[4], w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) w[j] = Ws[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = m0 + ty * 4 + i;
    if (row >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = n0 + tx * 4 + j;
      if (col >= g.N) continue;
      float v = acc[i][j];
      if (g.bias) v += g.bias[col];
      if (g.addmat) v += g.addmat[static_cast<size_t>(row) * g.ldadd + col];
      v = apply_act<kFast>(v, g.act);
      if (g.out_dtype == DT_BF16)
        reinterpret_cast<bf16*>(g.out)[static_cast<size_t>(row) * g.ldo + col] = __float2bfloat16_rn(v);
      else
        reinterpret_cast<float*>(g.out)[static_cast<size_t>(row) * g.ldo + col] = v;
    }
  }
}

}  // namespace

int gemm_simt_launch(const GemmDesc& g, bool fast_math, cudaStream_t s) {
  DS_CHECK(g.M > 0 && g.N > 0 && g.K > 0, DS_ERR_INVALID, "gemm_simt: empty problem M=%d N=%d K=%d", g.M, g.N, g.K);
  dim3 grid((g.M + TM - 1) / TM, (g.N + TN - 1) / TN);
  if (g.a_dtype == DT_BF16) {
    if (fast_math)
      ds_launch(gemm_simt_kernel<bf16, true>, dim3(grid), dim3(256), 0, s, g);
    else
      ds_launch(gemm_simt_kernel<bf16, false>, dim3(grid), dim3(256), 0, s, g);
  } else {
    if (fast_math)
      ds_launch(gemm_simt_kernel<float, true>, dim3(grid), dim3(256), 0, s, g);
    else
      ds_launch(gemm_simt_kernel<float, false>, dim3(grid), dim3(256), 0, s, g);
  }
  DS_CUDA_CHECK(cudaGetLastError());
  return DS_OK;
}
