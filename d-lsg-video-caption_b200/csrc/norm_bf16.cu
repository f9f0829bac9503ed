// LayerNorm forward / backward for the big bf16 activations of the graph encoder (obj_norm on the 59904 x 1024 region
// projections, layer.py:184-186): x, y (dy, dx) all bf16, plain LayerNorm (backward optionally through the tanh that
// the GEMM epilogue already applied to x).  HBM-bound streaming kernels:
//   * one warp per row, 16-byte loads (8 bf16 per lane per load), the NEXT row's loads are issued before the current
//     row is processed (software pipeline), gamma / beta live in shared memory;
//   * backward keeps the per-column dgamma / dbeta partial sums in REGISTERS across the row loop (each lane always
//     owns the same columns), reduces them across the CTA's warps once at the end, then one atomicAdd per column per CTA.
// Algorithmic bytes per row of D elements: forward 2D read + 2D written (+8 stats), backward 4D read + 2D written.
#include "common.cuh"

namespace dlsg {

__device__ __forceinline__ void unpack8(const uint4 u, float (&f)[8]) {
  f[0] = __uint_as_float(u.x << 16); f[1] = __uint_as_float(u.x & 0xffff0000u);
  f[2] = __uint_as_float(u.y << 16); f[3] = __uint_as_float(u.y & 0xffff0000u);
  f[4] = __uint_as_float(u.z << 16); f[5] = __uint_as_float(u.z & 0xffff0000u);
  f[6] = __uint_as_float(u.w << 16); f[7] = __uint_as_float(u.w & 0xffff0000u);
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&v);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  return make_uint4(pack2(f[0], f[1]), pack2(f[2], f[3]), pack2(f[4], f[5]), pack2(f[6], f[7]));
}

template <int NC>     // 16-byte chunks per lane: D <= 256 * NC
__global__ void __launch_bounds__(128)
norm_fwd_bf16_kernel(const dlsg_norm_fwd_t p) {
  pdl_prologue();
  extern __shared__ float sgb[];                 // gamma[D] | beta[D]
  const int D = p.D;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int c = threadIdx.x * 4; c < D; c += blockDim.x * 4) {
    *reinterpret_cast<float4*>(sgb + c) = *reinterpret_cast<const float4*>(p.gamma + c);
    *reinterpret_cast<float4*>(sgb + D + c) = *reinterpret_cast<const float4*>(p.beta + c);
  }
  __syncthreads();
  const __nv_bfloat16* X = reinterpret_cast<const __nv_bfloat16*>(p.x);
  __nv_bfloat16* Y = reinterpret_cast<__nv_bfloat16*>(p.y);
  const float invD = 1.f / (float)D;
  const int64_t stride = (int64_t)gridDim.x * nw;
  int64_t row = (int64_t)blockIdx.x * nw + w;
  uint4 cur[NC], nxt[NC];
#pragma unroll
  for (int j = 0; j < NC; ++j) {
    const int c = 256 * j + 8 * lane;
    cur[j] = make_uint4(0u, 0u, 0u, 0u);
    if (c < D && row < p.rows) cur[j] = *reinterpret_cast<const uint4*>(X + row * p.ldx + c);
  }
  for (; row < p.rows; row += stride) {
    const int64_t rn = row + stride;
#pragma unroll
    for (int j = 0; j < NC; ++j) {               // next row's loads in flight while this row is processed
      const int c = 256 * j + 8 * lane;
      nxt[j] = make_uint4(0u, 0u, 0u, 0u);
      if (c < D && rn < p.rows) nxt[j] = *reinterpret_cast<const uint4*>(X + rn * p.ldx + c);
    }
    float v[NC][8];
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < NC; ++j) {
      unpack8(cur[j], v[j]);                      // (columns >= D unpack to zeros)
#pragma unroll
      for (int u = 0; u < 8; ++u) sum += v[j][u];
    }
    const float mean = warp_sum(sum) * invD;
    float sq = 0.f;
#pragma unroll
    for (int j = 0; j < NC; ++j) {
      if (256 * j + 8 * lane < D) {
#pragma unroll
        for (int u = 0; u < 8; ++u) { const float a = v[j][u] - mean; sq = fmaf(a, a, sq); }
      }
    }
    const float rstd = rsqrtf(warp_sum(sq) * invD + 1e-5f);
    if (p.stats && lane == 0) { p.stats[row * 2] = mean; p.stats[row * 2 + 1] = rstd; }
#pragma unroll
    for (int j = 0; j < NC; ++j) {
      const int c = 256 * j + 8 * lane;
      if (c < D) {
        const float4 g0 = *reinterpret_cast<const float4*>(sgb + c), g1 = *reinterpret_cast<const float4*>(sgb + c + 4);
        const float4 b0 = *reinterpret_cast<const float4*>(sgb + D + c), b1 = *reinterpret_cast<const float4*>(sgb + D + c + 4);
        const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w}, b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        float y[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) y[u] = fmaf((v[j][u] - mean) * rstd, g[u], b[u]);
        *reinterpret_cast<uint4*>(Y + row * p.ldy + c) = pack8(y);
      }
    }
#pragma unroll
    for (int j = 0; j < NC; ++j) cur[j] = nxt[j];
  }
}

template <int NC>     // D <= 256 * NC, NC <= 4 (the dgamma / dbeta partial sums take 16 * NC registers)
__global__ void __launch_bounds__(128, 2)
norm_bwd_bf16_kernel(const dlsg_norm_bwd_t p) {
  pdl_prologue();
  extern __shared__ float sm[];                  // gamma[D], then (reused at the end) [nw][2][D] partial sums
  const int D = p.D;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float* sg = sm;
  for (int c = threadIdx.x * 4; c < D; c += blockDim.x * 4) *reinterpret_cast<float4*>(sg + c) = *reinterpret_cast<const float4*>(p.gamma + c);
  __syncthreads();
  const __nv_bfloat16* X = reinterpret_cast<const __nv_bfloat16*>(p.x);
  const __nv_bfloat16* DY = reinterpret_cast<const __nv_bfloat16*>(p.dy);
  __nv_bfloat16* DX = reinterpret_cast<__nv_bfloat16*>(p.dx);
  const bool dtanh = (p.flags & DLSG_NORM_IN_IS_TANH) != 0;
  const float invD = 1.f / (float)D;
  const int64_t stride = (int64_t)gridDim.x * nw;
  int64_t row = (int64_t)blockIdx.x * nw + w;
  float ag[NC][8], ab[NC][8], ax[NC][8];          // per-lane column partial sums: dgamma, dbeta, sum(dx)
#pragma unroll
  for (int j = 0; j < NC; ++j)
#pragma unroll
    for (int u = 0; u < 8; ++u) { ag[j][u] = 0.f; ab[j][u] = 0.f; ax[j][u] = 0.f; }
  const bool want_dxsum = p.dxsum != nullptr;
  uint4 cx[NC], cd[NC], nx[NC], nd[NC];
  float2 st = make_float2(0.f, 0.f), stn = st;
#pragma unroll
  for (int j = 0; j < NC; ++j) {
    const int c = 256 * j + 8 * lane;
    cx[j] = make_uint4(0u, 0u, 0u, 0u); cd[j] = cx[j];
    if (c < D && row < p.rows) {
      cx[j] = *reinterpret_cast<const uint4*>(X + row * p.ldx + c);
      cd[j] = *reinterpret_cast<const uint4*>(DY + row * p.lddy + c);
    }
  }
  if (row < p.rows) st = *reinterpret_cast<const float2*>(p.stats + row * 2);
  for (; row < p.rows; row += stride) {
    const int64_t rn = row + stride;
#pragma unroll
    for (int j = 0; j < NC; ++j) {
      const int c = 256 * j + 8 * lane;
      nx[j] = make_uint4(0u, 0u, 0u, 0u); nd[j] = nx[j];
      if (c < D && rn < p.rows) {
        nx[j] = *reinterpret_cast<const uint4*>(X + rn * p.ldx + c);
        nd[j] = *reinterpret_cast<const uint4*>(DY + rn * p.lddy + c);
      }
    }
    if (rn < p.rows) stn = *reinterpret_cast<const float2*>(p.stats + rn * 2);
    const float mean = st.x, rstd = st.y;
    float xh[NC][8], d[NC][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < NC; ++j) {
      const int c = 256 * j + 8 * lane;
      float xv[8], dy[8];
      unpack8(cx[j], xv); unpack8(cd[j], dy);
      float g[8];
      if (c < D) {
        const float4 g0 = *reinterpret_cast<const float4*>(sg + c), g1 = *reinterpret_cast<const float4*>(sg + c + 4);
        g[0] = g0.x; g[1] = g0.y; g[2] = g0.z; g[3] = g0.w; g[4] = g1.x; g[5] = g1.y; g[6] = g1.z; g[7] = g1.w;
      } else {
#pragma unroll
        for (int u = 0; u < 8; ++u) g[u] = 0.f;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        xh[j][u] = (c < D) ? (xv[u] - mean) * rstd : 0.f;
        ag[j][u] = fmaf(dy[u], xh[j][u], ag[j][u]);
        ab[j][u] += dy[u];
        d[j][u] = dy[u] * g[u];
        s1 += d[j][u];
        s2 = fmaf(d[j][u], xh[j][u], s2);
      }
    }
    s1 = warp_sum(s1) * invD; s2 = warp_sum(s2) * invD;
    const float istd = 1.f / rstd;
#pragma unroll
    for (int j = 0; j < NC; ++j) {
      const int c = 256 * j + 8 * lane;
      if (c < D) {
        float o[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          o[u] = rstd * (d[j][u] - s1 - xh[j][u] * s2);
          if (dtanh) { const float t = fmaf(xh[j][u], istd, mean); o[u] *= (1.f - t * t); }   // t = x (the tanh output)
          if (want_dxsum) ax[j][u] += o[u];
        }
        *reinterpret_cast<uint4*>(DX + row * p.lddx + c) = pack8(o);
      }
    }
#pragma unroll
    for (int j = 0; j < NC; ++j) { cx[j] = nx[j]; cd[j] = nd[j]; }
    st = stn;
  }
  if (want_dxsum) {                              // uniform; reduce sum(dx) across the CTA's warps, one atomic per column
    __syncthreads();                             // gamma in smem no longer needed
    float* mx = sm + (size_t)w * D;
#pragma unroll
    for (int j = 0; j < NC; ++j) {
      const int c = 256 * j + 8 * lane;
      if (c < D) {
        *reinterpret_cast<float4*>(mx + c) = make_float4(ax[j][0], ax[j][1], ax[j][2], ax[j][3]);
        *reinterpret_cast<float4*>(mx + c + 4) = make_float4(ax[j][4], ax[j][5], ax[j][6], ax[j][7]);
      }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
      float tx = 0.f;
      for (int ww = 0; ww < nw; ++ww) tx += sm[(size_t)ww * D + c];
      atomicAdd(&p.dxsum[c], tx);
    }
  }
  if (p.dgamma == nullptr) return;               // uniform
  __syncthreads();                               // gamma / dxsum staging in smem no longer needed
  float* mg = sm + (size_t)w * 2 * D;
#pragma unroll
  for (int j = 0; j < NC; ++j) {
    const int c = 256 * j + 8 * lane;
    if (c < D) {
      *reinterpret_cast<float4*>(mg + c) = make_float4(ag[j][0], ag[j][1], ag[j][2], ag[j][3]);
      *reinterpret_cast<float4*>(mg + c + 4) = make_float4(ag[j][4], ag[j][5], ag[j][6], ag[j][7]);
      *reinterpret_cast<float4*>(mg + D + c) = make_float4(ab[j][0], ab[j][1], ab[j][2], ab[j][3]);
      *reinterpret_cast<float4*>(mg + D + c + 4) = make_float4(ab[j][4], ab[j][5], ab[j][6], ab[j][7]);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float tg = 0.f, tb = 0.f;
    for (int ww = 0; ww < nw; ++ww) { tg += sm[(size_t)ww * 2 * D + c]; tb += sm[(size_t)ww * 2 * D + D + c]; }
    atomicAdd(&p.dgamma[c], tg);
    atomicAdd(&p.dbeta[c], tb);
  }
}

static inline bool al16p(const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; }

// Eligibility: the streaming bf16 form (everything else goes to the generic kernels in rowops.cu)
bool norm_fwd_bf16_ok(const dlsg_norm_fwd_t* p) {
  return p->x_dtype == DLSG_BF16 && p->y && p->y_dtype == DLSG_BF16 && !p->y2 && !p->res && p->flags == 0 && p->drop_p <= 0.f &&
         p->D % 8 == 0 && p->D <= 2048 && p->ldx % 8 == 0 && p->ldy % 8 == 0 && al16p(p->x) && al16p(p->y) && al16p(p->gamma) &&
         al16p(p->beta) && p->rows >= 2048;
}
bool norm_bwd_bf16_ok(const dlsg_norm_bwd_t* p) {
  return p->x_dtype == DLSG_BF16 && p->dy_dtype == DLSG_BF16 && p->dx && p->dx_dtype == DLSG_BF16 && !p->res && !p->dx_accum &&
         (p->flags & ~DLSG_NORM_IN_IS_TANH) == 0 && p->drop_p <= 0.f && p->D % 8 == 0 && p->D <= 1024 && p->ldx % 8 == 0 &&
         p->lddy % 8 == 0 && p->lddx % 8 == 0 && al16p(p->x) && al16p(p->dy) && al16p(p->dx) && al16p(p->gamma) &&
         (p->dgamma == nullptr) == (p->dbeta == nullptr) && p->rows >= 2048;
}

template <int NC>
static int fwd_launch(const dlsg_norm_fwd_t* p, cudaStream_t st) {
  const int nw = 4;
  int64_t blocks = (p->rows + nw - 1) / nw;
  if (blocks > kNumSM * 8) blocks = kNumSM * 8;
  DLSG_LAUNCH(norm_fwd_bf16_kernel<NC>, (unsigned)blocks, nw * 32, (size_t)2 * p->D * sizeof(float), st, *p);
  return check_launch("norm_fwd_bf16_kernel");
}
int norm_fwd_bf16_launch(const dlsg_norm_fwd_t* p, cudaStream_t st) {
  const int nc = (p->D + 255) / 256;
  if (nc <= 1) return fwd_launch<1>(p, st);
  if (nc <= 2) return fwd_launch<2>(p, st);
  if (nc <= 4) return fwd_launch<4>(p, st);
  return fwd_launch<8>(p, st);
}

template <int NC>
static int bwd_launch(const dlsg_norm_bwd_t* p, cudaStream_t st) {
  const int nw = 4;
  const size_t smem = (size_t)nw * 2 * p->D * sizeof(float);
  int64_t blocks = (p->rows + nw - 1) / nw;
  if (blocks > kNumSM * 2) blocks = kNumSM * 2;
  DLSG_LAUNCH(norm_bwd_bf16_kernel<NC>, (unsigned)blocks, nw * 32, smem, st, *p);
  return check_launch("norm_bwd_bf16_kernel");
}
int norm_bwd_bf16_launch(const dlsg_norm_bwd_t* p, cudaStream_t st) {
  const int nc = (p->D + 255) / 256;
  if (nc <= 1) return bwd_launch<1>(p, st);
  if (nc <= 2) return bwd_launch<2>(p, st);
  return bwd_launch<4>(p, st);
}

}  // namespace dlsg
