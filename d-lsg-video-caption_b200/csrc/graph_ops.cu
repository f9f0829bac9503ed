// Latent-semantic-graph pooling (LatentPSL, models/sublayer.py:189-198) as one fused kernel per direction:
//   G = X theta^T (T x P), Gs = softmax over the T frames, N = Gs^T X (P x H)          [forward]
//   dX = Gs dN + dG theta, dtheta += dG^T X, dG = softmax-bwd(X dN^T)                    [backward]
// One CTA per clip; theta / dN (P x H) staged in shared memory, X rows stream through L1 (read twice).
#include "common.cuh"

namespace dlsg {

constexpr int LP_MAXP = 8, LP_MAXT = 32;

// dots[t][p] = X[t,:] . W[p,:]  for one clip; W already in smem (P x H). warp per t.
__device__ __forceinline__ void lp_row_dots(const float* __restrict__ X, const float* __restrict__ Wsm, int T, int P, int H,
                                            float (*out)[LP_MAXP]) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int t = w; t < T; t += nw) {
    float acc[LP_MAXP];
#pragma unroll
    for (int j = 0; j < LP_MAXP; ++j) acc[j] = 0.f;
    for (int c = lane * 4; c < H; c += 128) {
      const float4 x = *reinterpret_cast<const float4*>(X + (int64_t)t * H + c);
#pragma unroll
      for (int j = 0; j < LP_MAXP; ++j) {
        if (j < P) {
          const float4 wv = *reinterpret_cast<const float4*>(Wsm + j * H + c);
          acc[j] += (x.x * wv.x + x.y * wv.y) + (x.z * wv.z + x.w * wv.w);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < LP_MAXP; ++j) {
      const float s = warp_sum(acc[j]);
      if (lane == 0 && j < P) out[t][j] = s;
    }
  }
}

// up to two independent poolings (the object and the motion encoder) per launch: blockIdx.y selects the set of pointers
struct LpFwdArgs { const float* X[2]; const float* theta[2]; float* Gs[2]; float* N[2]; };
struct LpBwdArgs { const float* X[2]; const float* theta[2]; const float* Gs[2]; const float* dN[2]; float* dX[2]; float* dtheta[2]; };

__global__ void __launch_bounds__(256)
latent_psl_fwd_kernel(const LpFwdArgs a, int T, int P, int H) {
  pdl_prologue();
  extern __shared__ float sm[];                // theta (P x H)
  __shared__ float g[LP_MAXT][LP_MAXP];
  const int b = blockIdx.x, e = blockIdx.y;
  const float* __restrict__ X = a.X[e];
  const float* __restrict__ theta = a.theta[e];
  float* __restrict__ Gs = a.Gs[e];
  float* __restrict__ N = a.N[e];
  X += (int64_t)b * T * H;
  for (int i = threadIdx.x * 4; i < P * H; i += blockDim.x * 4)
    *reinterpret_cast<float4*>(sm + i) = *reinterpret_cast<const float4*>(theta + i);
  __syncthreads();
  lp_row_dots(X, sm, T, P, H, g);
  __syncthreads();
  if (threadIdx.x < P) {                        // softmax over the T frames for node p
    const int p = threadIdx.x;
    float mx = -INFINITY;
    for (int t = 0; t < T; ++t) mx = fmaxf(mx, g[t][p]);
    float s = 0.f;
    for (int t = 0; t < T; ++t) { const float e = expf(g[t][p] - mx); g[t][p] = e; s += e; }
    const float inv = 1.f / s;
    for (int t = 0; t < T; ++t) { g[t][p] *= inv; Gs[((int64_t)b * T + t) * P + p] = g[t][p]; }
  }
  __syncthreads();
  for (int c = threadIdx.x * 4; c < H; c += blockDim.x * 4) {
    float4 acc[LP_MAXP];
#pragma unroll
    for (int j = 0; j < LP_MAXP; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int t = 0; t < T; ++t) {
      const float4 x = *reinterpret_cast<const float4*>(X + (int64_t)t * H + c);
#pragma unroll
      for (int j = 0; j < LP_MAXP; ++j) {
        if (j < P) { const float a = g[t][j]; acc[j].x = fmaf(a, x.x, acc[j].x); acc[j].y = fmaf(a, x.y, acc[j].y); acc[j].z = fmaf(a, x.z, acc[j].z); acc[j].w = fmaf(a, x.w, acc[j].w); }
      }
    }
#pragma unroll
    for (int j = 0; j < LP_MAXP; ++j)
      if (j < P) *reinterpret_cast<float4*>(N + ((int64_t)b * P + j) * H + c) = acc[j];
  }
}

__global__ void __launch_bounds__(256)
latent_psl_bwd_kernel(const LpBwdArgs a, int T, int P, int H) {
  pdl_prologue();
  extern __shared__ float sm[];                // dN (P x H) then theta (P x H)
  __shared__ float gs[LP_MAXT][LP_MAXP];
  __shared__ float dg[LP_MAXT][LP_MAXP];
  const int b = blockIdx.x, e = blockIdx.y;
  const float* __restrict__ X = a.X[e];
  const float* __restrict__ theta = a.theta[e];
  const float* __restrict__ Gs = a.Gs[e];
  const float* __restrict__ dN = a.dN[e];
  float* __restrict__ dX = a.dX[e];
  float* __restrict__ dtheta = a.dtheta[e];
  X += (int64_t)b * T * H;
  float* sdn = sm;
  float* sth = sm + P * H;
  for (int i = threadIdx.x * 4; i < P * H; i += blockDim.x * 4) {
    *reinterpret_cast<float4*>(sdn + i) = *reinterpret_cast<const float4*>(dN + (int64_t)b * P * H + i);
    *reinterpret_cast<float4*>(sth + i) = *reinterpret_cast<const float4*>(theta + i);
  }
  for (int i = threadIdx.x; i < T * P; i += blockDim.x) gs[i / P][i % P] = Gs[(int64_t)b * T * P + i];
  __syncthreads();
  lp_row_dots(X, sdn, T, P, H, dg);             // dGs[t][p] = X[t,:] . dN[p,:]
  __syncthreads();
  if (threadIdx.x < P) {                        // softmax (over t) backward
    const int p = threadIdx.x;
    float dot = 0.f;
    for (int t = 0; t < T; ++t) dot = fmaf(gs[t][p], dg[t][p], dot);
    for (int t = 0; t < T; ++t) dg[t][p] = gs[t][p] * (dg[t][p] - dot);
  }
  __syncthreads();
  for (int c = threadIdx.x * 4; c < H; c += blockDim.x * 4) {
    float4 dth[LP_MAXP];
#pragma unroll
    for (int j = 0; j < LP_MAXP; ++j) dth[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int t = 0; t < T; ++t) {
      const float4 x = *reinterpret_cast<const float4*>(X + (int64_t)t * H + c);
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int j = 0; j < LP_MAXP; ++j) {
        if (j < P) {
          const float a = gs[t][j], d = dg[t][j];
          const float4 n = *reinterpret_cast<const float4*>(sdn + j * H + c), th = *reinterpret_cast<const float4*>(sth + j * H + c);
          o.x += a * n.x + d * th.x; o.y += a * n.y + d * th.y; o.z += a * n.z + d * th.z; o.w += a * n.w + d * th.w;
          dth[j].x = fmaf(d, x.x, dth[j].x); dth[j].y = fmaf(d, x.y, dth[j].y); dth[j].z = fmaf(d, x.z, dth[j].z); dth[j].w = fmaf(d, x.w, dth[j].w);
        }
      }
      *reinterpret_cast<float4*>(dX + ((int64_t)b * T + t) * H + c) = o;
    }
#pragma unroll
    for (int j = 0; j < LP_MAXP; ++j) {
      if (j < P) {
        float* d = dtheta + (int64_t)j * H + c;
        atomicAdd(d, dth[j].x); atomicAdd(d + 1, dth[j].y); atomicAdd(d + 2, dth[j].z); atomicAdd(d + 3, dth[j].w);
      }
    }
  }
}

}  // namespace dlsg

using namespace dlsg;

extern "C" {

static int lp_check(const char* what, int32_t E, int32_t T, int32_t P, int32_t H, size_t smem) {
  DLSG_REQUIRE(E >= 1 && E <= 2, "%s: 1 or 2 poolings per launch", what);
  DLSG_REQUIRE(P >= 1 && P <= LP_MAXP && T >= 1 && T <= LP_MAXT && H % 4 == 0, "%s: unsupported shape T=%d P=%d H=%d", what, T, P, H);
  DLSG_REQUIRE(smem <= 200 * 1024, "%s: P*H too large", what);
  return 0;
}

int dlsg_latent_psl_fwd_multi(const float* const* X, const float* const* theta, float* const* Gs, float* const* N, int32_t E,
                              int32_t B, int32_t T, int32_t P, int32_t H, void* stream) {
  const size_t smem = (size_t)P * H * sizeof(float);
  if (int rc = lp_check("latent_psl_fwd", E, T, P, H, smem)) return rc;
  if (B <= 0) return 0;
  LpFwdArgs a = {};
  for (int e = 0; e < E; ++e) {
    DLSG_REQUIRE(X[e] && theta[e] && Gs[e] && N[e], "latent_psl_fwd: null operand (set %d)", e);
    a.X[e] = X[e]; a.theta[e] = theta[e]; a.Gs[e] = Gs[e]; a.N[e] = N[e];
  }
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(latent_psl_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); attr = true; }
  DLSG_LAUNCH(latent_psl_fwd_kernel, dim3(B, E), 256, smem, (cudaStream_t)stream, a, T, P, H);
  return check_launch("latent_psl_fwd_kernel");
}

int dlsg_latent_psl_bwd_multi(const float* const* X, const float* const* theta, const float* const* Gs, const float* const* dN,
                              float* const* dX, float* const* dtheta, int32_t E, int32_t B, int32_t T, int32_t P, int32_t H, void* stream) {
  const size_t smem = (size_t)2 * P * H * sizeof(float);
  if (int rc = lp_check("latent_psl_bwd", E, T, P, H, smem)) return rc;
  if (B <= 0) return 0;
  LpBwdArgs a = {};
  for (int e = 0; e < E; ++e) {
    DLSG_REQUIRE(X[e] && theta[e] && Gs[e] && dN[e] && dX[e] && dtheta[e], "latent_psl_bwd: null operand (set %d)", e);
    a.X[e] = X[e]; a.theta[e] = theta[e]; a.Gs[e] = Gs[e]; a.dN[e] = dN[e]; a.dX[e] = dX[e]; a.dtheta[e] = dtheta[e];
  }
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(latent_psl_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); attr = true; }
  DLSG_LAUNCH(latent_psl_bwd_kernel, dim3(B, E), 256, smem, (cudaStream_t)stream, a, T, P, H);
  return check_launch("latent_psl_bwd_kernel");
}

int dlsg_latent_psl_fwd(const float* X, const float* theta, float* Gs, float* N, int32_t B, int32_t T, int32_t P, int32_t H, void* stream) {
  return dlsg_latent_psl_fwd_multi(&X, &theta, &Gs, &N, 1, B, T, P, H, stream);
}

int dlsg_latent_psl_bwd(const float* X, const float* theta, const float* Gs, const float* dN, float* dX, float* dtheta,
                        int32_t B, int32_t T, int32_t P, int32_t H, void* stream) {
  return dlsg_latent_psl_bwd_multi(&X, &theta, &Gs, &dN, &dX, &dtheta, 1, B, T, P, H, stream);
}

}  // extern "C"
