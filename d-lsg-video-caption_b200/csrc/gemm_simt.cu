// fp32 FFMA GEMM with arbitrary element strides: the strict-parity ("fp32 mode") path and the path for
// tiny / oddly-shaped products (26x26 attention, P<=8 node pooling) that do not belong on tensor cores.
// D[b](M,N) = epi(alpha * sum_k A[m*sam+k*sak] * B[n*sbn+k*sbk] + bias), 64x64x16 tiles, 4x4 per thread.
#include "common.cuh"

namespace dlsg {

constexpr int SB = 64, SK = 16;

template <typename TA>
__device__ __forceinline__ float ldf(const TA* p, int64_t i);
template <> __device__ __forceinline__ float ldf<float>(const float* p, int64_t i) { return p[i]; }
template <> __device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p, int64_t i) { return __bfloat162float(p[i]); }

template <typename TA, typename TB>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const TA* __restrict__ A, const TB* __restrict__ B, void* __restrict__ D, const float* __restrict__ bias,
                 int M, int N, int K, int64_t sam, int64_t sak, int64_t sbn, int64_t sbk, int64_t ldd,
                 int64_t stride_a, int64_t stride_b, int64_t stride_d, int d_dtype, int flags, float alpha) {
  pdl_prologue();
  __shared__ float As[SK][SB + 4];
  __shared__ float Bs[SK][SB + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * SB, n0 = blockIdx.x * SB;
  A += (int64_t)blockIdx.z * stride_a;
  B += (int64_t)blockIdx.z * stride_b;
  const int tx = tid & 15, ty = tid >> 4;      // 16 x 16 threads, each a 4x4 micro-tile
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  // loader mapping: k fastest when the operand is K-contiguous, row fastest otherwise
  const bool a_kfast = (sak == 1), b_kfast = (sbk == 1);
  for (int k0 = 0; k0 < K; k0 += SK) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int e = tid + r * 256;
      int m, k;
      if (a_kfast) { k = e & 15; m = e >> 4; } else { m = e & 63; k = e >> 6; }
      float v = 0.f;
      if (m0 + m < M && k0 + k < K) v = ldf<TA>(A, (int64_t)(m0 + m) * sam + (int64_t)(k0 + k) * sak);
      As[k][m] = v;
      int n, kk;
      if (b_kfast) { kk = e & 15; n = e >> 4; } else { n = e & 63; kk = e >> 6; }
      float w = 0.f;
      if (n0 + n < N && k0 + kk < K) w = ldf<TB>(B, (int64_t)(n0 + n) * sbn + (int64_t)(k0 + kk) * sbk);
      Bs[kk][n] = w;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SK; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  const bool store_t = flags & DLSG_EPI_STORE_T;
  const int64_t dbase = (int64_t)blockIdx.z * stride_d;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float x = acc[i][j] * alpha;
      if (bias) {
        if (flags & DLSG_EPI_BIAS_N) x += bias[n];
        if (flags & DLSG_EPI_BIAS_M) x += bias[m];
      }
      if (flags & DLSG_EPI_TANH) x = tanhf(x);
      const int64_t idx = dbase + (store_t ? ((int64_t)n * ldd + m) : ((int64_t)m * ldd + n));
      if (flags & DLSG_EPI_ACCUM) x += ld_as_float(D, d_dtype, idx);
      st_from_float(D, d_dtype, idx, x);
    }
  }
}

int gemm_simt_dispatch(const dlsg_gemm_t* g, cudaStream_t st) {
  DLSG_REQUIRE(g->M > 0 && g->N > 0 && g->K >= 0, "gemm_simt: empty problem");
  DLSG_REQUIRE(g->splitk <= 1, "gemm_simt: split-K unsupported");
  const int batch = g->batch < 1 ? 1 : g->batch;
  dim3 grid((g->N + SB - 1) / SB, (g->M + SB - 1) / SB, batch);
  DLSG_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "gemm_simt: grid too large");
#define DLSG_SIMT_LAUNCH(TA, TB)                                                                                     \
  DLSG_LAUNCH((gemm_simt_kernel<TA, TB>), grid, 256, 0, st, (const TA*)g->A, (const TB*)g->B, g->D, g->bias, g->M, g->N, g->K,   \
                                                 g->sam, g->sak, g->sbn, g->sbk, g->ldd, g->stride_a, g->stride_b,     \
                                                 g->stride_d, g->d_dtype, g->flags, g->alpha)
  if (g->a_dtype == DLSG_F32 && g->b_dtype == DLSG_F32) DLSG_SIMT_LAUNCH(float, float);
  else if (g->a_dtype == DLSG_F32) DLSG_SIMT_LAUNCH(float, __nv_bfloat16);
  else if (g->b_dtype == DLSG_F32) DLSG_SIMT_LAUNCH(__nv_bfloat16, float);
  else DLSG_SIMT_LAUNCH(__nv_bfloat16, __nv_bfloat16);
#undef DLSG_SIMT_LAUNCH
  return check_launch("gemm_simt_kernel");
}

}  // namespace dlsg
