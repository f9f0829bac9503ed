// HBM-bound row / pointwise kernels: dtype+layout conversion, the Tanh->LayerNorm family, the LSTM cell,
// axis softmax, embedding gather/scatter and small elementwise helpers.  One warp owns one row (warp-shuffle
// statistics, 128-bit / 64-bit vector accesses when the row is aligned), rows are staged once in shared memory.
#include "common.cuh"

namespace dlsg {

// ------------------------------------------------------------------------------------------- convert2d
// 32x32 tiles through smem: dst (cast) and optional dstT (transposed cast) in one pass over src.
__global__ void __launch_bounds__(256)
convert2d_kernel(const void* __restrict__ src_, int sdt, int64_t lds, void* __restrict__ dst_, int ddt, int64_t ldd,
                 void* __restrict__ dstT_, int64_t ldt, int64_t rows, int64_t cols, int64_t bs_src, int64_t bs_dst, int64_t bs_dstT) {
  pdl_prologue();
  __shared__ float tile[32][33];
  const int es = sdt == DLSG_F32 ? 4 : 2, ed = ddt == DLSG_F32 ? 4 : 2;
  const void* src = reinterpret_cast<const uint8_t*>(src_) + (int64_t)blockIdx.z * bs_src * es;
  void* dst = dst_ ? reinterpret_cast<uint8_t*>(dst_) + (int64_t)blockIdx.z * bs_dst * ed : nullptr;
  void* dstT = dstT_ ? reinterpret_cast<uint8_t*>(dstT_) + (int64_t)blockIdx.z * bs_dstT * ed : nullptr;
  const int64_t c0 = (int64_t)blockIdx.x * 32, r0 = (int64_t)blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t r = r0 + ty + i * 8, c = c0 + tx;
    float v = 0.f;
    if (r < rows && c < cols) {
      v = ld_as_float(src, sdt, r * lds + c);
      if (dst) st_from_float(dst, ddt, r * ldd + c, v);
    }
    tile[ty + i * 8][tx] = v;
  }
  if (!dstT) return;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t c = c0 + ty + i * 8, r = r0 + tx;
    if (r < rows && c < cols) st_from_float(dstT, ddt, c * ldt + r, tile[tx][ty + i * 8]);
  }
}

// flat vectorised cast fp32 -> bf16 (contiguous case): 8 elements / thread, 2x128-bit loads, 1x128-bit store
__global__ void __launch_bounds__(256)
cast_f32_bf16_flat(const float4* __restrict__ src, uint4* __restrict__ dst, int64_t n8) {
  pdl_prologue();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 a = __ldg(src + 2 * i), b = __ldg(src + 2 * i + 1);
    __nv_bfloat162 p0 = __floats2bfloat162_rn(a.x, a.y), p1 = __floats2bfloat162_rn(a.z, a.w);
    __nv_bfloat162 p2 = __floats2bfloat162_rn(b.x, b.y), p3 = __floats2bfloat162_rn(b.z, b.w);
    uint4 o;
    o.x = *reinterpret_cast<uint32_t*>(&p0); o.y = *reinterpret_cast<uint32_t*>(&p1);
    o.z = *reinterpret_cast<uint32_t*>(&p2); o.w = *reinterpret_cast<uint32_t*>(&p3);
    dst[i] = o;
  }
}

// ------------------------------------------------------------------------------------------- row access
struct RowIO {
  // load 4 consecutive elements starting at element c (c % 4 == 0) of a row
  __device__ static __forceinline__ float4 ld4(const void* base, int dt, int64_t off) {
    if (dt == DLSG_F32) return *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + off);
    const uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(base) + off);
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&u.x), b = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
    return make_float4(__low2float(a), __high2float(a), __low2float(b), __high2float(b));
  }
  __device__ static __forceinline__ void st4(void* base, int dt, int64_t off, float4 v) {
    if (dt == DLSG_F32) { *reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + off) = v; return; }
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 u; u.x = *reinterpret_cast<uint32_t*>(&a); u.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(base) + off) = u;
  }
};
static inline bool vec_ok(const void* p, int dt, int64_t ld, int D) {
  if (!p) return true;
  const int es = dt == DLSG_F32 ? 4 : 2;
  return (D % 4 == 0) && (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(p) % (4 * es)) == 0);
}

// ------------------------------------------------------------------------------------------- norm fwd
template <bool VEC>
__global__ void __launch_bounds__(256)
norm_fwd_kernel(const dlsg_norm_fwd_t p) {
  pdl_prologue();
  extern __shared__ float srow[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float* t = srow + (size_t)w * p.D;
  const int D = p.D;
  for (int64_t row = (int64_t)blockIdx.x * nw + w; row < p.rows; row += (int64_t)gridDim.x * nw) {
    float sum = 0.f;
    if (VEC) {
      for (int c = lane * 4; c < D; c += 128) {
        float4 v = RowIO::ld4(p.x, p.x_dtype, row * p.ldx + c);
        if (p.res) { const float4 r = RowIO::ld4(p.res, p.res_dtype, row * p.ldres + c); v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w; }
        if (p.flags & DLSG_NORM_PRE_TANH) { v.x = tanhf(v.x); v.y = tanhf(v.y); v.z = tanhf(v.z); v.w = tanhf(v.w); }
        *reinterpret_cast<float4*>(t + c) = v;
        sum += (v.x + v.y) + (v.z + v.w);
      }
    } else {
      for (int c = lane; c < D; c += 32) {
        float v = ld_as_float(p.x, p.x_dtype, row * p.ldx + c);
        if (p.res) v += ld_as_float(p.res, p.res_dtype, row * p.ldres + c);
        if (p.flags & DLSG_NORM_PRE_TANH) v = tanhf(v);
        t[c] = v;
        sum += v;
      }
    }
    const float mean = warp_sum(sum) / (float)D;
    float sq = 0.f;
    __syncwarp();
    for (int c = lane; c < D; c += 32) { const float d = t[c] - mean; sq = fmaf(d, d, sq); }
    const float rstd = rsqrtf(warp_sum(sq) / (float)D + 1e-5f);
    if (p.stats && lane == 0) { p.stats[row * 2] = mean; p.stats[row * 2 + 1] = rstd; }
    if (VEC) {
      for (int c = lane * 4; c < D; c += 128) {
        const float4 v = *reinterpret_cast<const float4*>(t + c);
        const float4 g = *reinterpret_cast<const float4*>(p.gamma + c), b = *reinterpret_cast<const float4*>(p.beta + c);
        float4 y;
        y.x = (v.x - mean) * rstd * g.x + b.x; y.y = (v.y - mean) * rstd * g.y + b.y;
        y.z = (v.z - mean) * rstd * g.z + b.z; y.w = (v.w - mean) * rstd * g.w + b.w;
        if (p.flags & DLSG_NORM_POST_TANH) { y.x = tanhf(y.x); y.y = tanhf(y.y); y.z = tanhf(y.z); y.w = tanhf(y.w); }
        if (p.drop_p > 0.f) {
          const uint64_t i0 = p.offset + (uint64_t)row * D + c;
          y.x *= drop_scale(p.drop_p, p.seed, i0); y.y *= drop_scale(p.drop_p, p.seed, i0 + 1);
          y.z *= drop_scale(p.drop_p, p.seed, i0 + 2); y.w *= drop_scale(p.drop_p, p.seed, i0 + 3);
        }
        if (p.y) RowIO::st4(p.y, p.y_dtype, row * p.ldy + c, y);
        if (p.y2) RowIO::st4(p.y2, p.y2_dtype, row * p.ldy2 + c, y);
      }
    } else {
      for (int c = lane; c < D; c += 32) {
        float y = (t[c] - mean) * rstd * p.gamma[c] + p.beta[c];
        if (p.flags & DLSG_NORM_POST_TANH) y = tanhf(y);
        if (p.drop_p > 0.f) y *= drop_scale(p.drop_p, p.seed, p.offset + (uint64_t)row * D + c);
        if (p.y) st_from_float(p.y, p.y_dtype, row * p.ldy + c, y);
        if (p.y2) st_from_float(p.y2, p.y2_dtype, row * p.ldy2 + c, y);
      }
    }
    __syncwarp();
  }
}

// Register-resident fast path: one warp per row, lane owns float4 chunks {128*j + 4*lane}, j < NV (D <= 128*NV).
// All loads of a row are issued back to back (one memory latency), statistics by warp shuffles, no shared memory.
template <int NV>
__global__ void __launch_bounds__(128)
norm_fwd_vec_kernel(const dlsg_norm_fwd_t p) {
  pdl_prologue();
  const int D = p.D;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const float invD = 1.f / (float)D;
  for (int64_t row = (int64_t)blockIdx.x * nw + w; row < p.rows; row += (int64_t)gridDim.x * nw) {
    float4 t[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int c = 128 * j + 4 * lane;
      t[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < D) t[j] = RowIO::ld4(p.x, p.x_dtype, row * p.ldx + c);
    }
    if (p.res) {
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int c = 128 * j + 4 * lane;
        if (c < D) { const float4 r = RowIO::ld4(p.res, p.res_dtype, row * p.ldres + c); t[j].x += r.x; t[j].y += r.y; t[j].z += r.z; t[j].w += r.w; }
      }
    }
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int c = 128 * j + 4 * lane;
      if (c < D) {
        if (p.flags & DLSG_NORM_PRE_TANH) { t[j].x = tanhf(t[j].x); t[j].y = tanhf(t[j].y); t[j].z = tanhf(t[j].z); t[j].w = tanhf(t[j].w); }
        sum += (t[j].x + t[j].y) + (t[j].z + t[j].w);
      }
    }
    const float mean = warp_sum(sum) * invD;
    float sq = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int c = 128 * j + 4 * lane;
      if (c < D) {
        const float a = t[j].x - mean, b = t[j].y - mean, cc = t[j].z - mean, d = t[j].w - mean;
        sq += (a * a + b * b) + (cc * cc + d * d);
      }
    }
    const float rstd = rsqrtf(warp_sum(sq) * invD + 1e-5f);
    if (p.stats && lane == 0) { p.stats[row * 2] = mean; p.stats[row * 2 + 1] = rstd; }
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int c = 128 * j + 4 * lane;
      if (c < D) {
        const float4 g = *reinterpret_cast<const float4*>(p.gamma + c), b = *reinterpret_cast<const float4*>(p.beta + c);
        float4 y;
        y.x = (t[j].x - mean) * rstd * g.x + b.x; y.y = (t[j].y - mean) * rstd * g.y + b.y;
        y.z = (t[j].z - mean) * rstd * g.z + b.z; y.w = (t[j].w - mean) * rstd * g.w + b.w;
        if (p.flags & DLSG_NORM_POST_TANH) { y.x = tanhf(y.x); y.y = tanhf(y.y); y.z = tanhf(y.z); y.w = tanhf(y.w); }
        if (p.drop_p > 0.f) {
          const float4 m = drop_mask4(p.drop_p, 1.f / (1.f - p.drop_p), p.seed, p.offset + (uint64_t)row * D + c);
          y.x *= m.x; y.y *= m.y; y.z *= m.z; y.w *= m.w;
        }
        if (p.y) RowIO::st4(p.y, p.y_dtype, row * p.ldy + c, y);
        if (p.y2) RowIO::st4(p.y2, p.y2_dtype, row * p.ldy2 + c, y);
      }
    }
  }
}

template <int NV>
static int norm_fwd_vec_launch(const dlsg_norm_fwd_t* p, cudaStream_t st) {
  const int nw = 4;
  int64_t blocks = (p->rows + nw - 1) / nw;
  if (blocks > kNumSM * 8) blocks = kNumSM * 8;
  DLSG_LAUNCH(norm_fwd_vec_kernel<NV>, (unsigned)blocks, nw * 32, 0, st, *p);
  return check_launch("norm_fwd_vec_kernel");
}

bool norm_fwd_bf16_ok(const dlsg_norm_fwd_t* p);
int norm_fwd_bf16_launch(const dlsg_norm_fwd_t* p, cudaStream_t st);
bool norm_bwd_bf16_ok(const dlsg_norm_bwd_t* p);
int norm_bwd_bf16_launch(const dlsg_norm_bwd_t* p, cudaStream_t st);

int norm_fwd_launch(const dlsg_norm_fwd_t* p, cudaStream_t st) {
  DLSG_REQUIRE(p->rows >= 0 && p->D > 0 && p->D <= 8192, "norm_fwd: bad shape rows=%lld D=%d", (long long)p->rows, p->D);
  if (p->rows == 0) return 0;
  if (norm_fwd_bf16_ok(p)) return norm_fwd_bf16_launch(p, st);      // big bf16 activations: streaming kernel (norm_bf16.cu)
  if (p->D <= 2048 && vec_ok(p->x, p->x_dtype, p->ldx, p->D) && vec_ok(p->res, p->res_dtype, p->ldres, p->D) &&
      vec_ok(p->y, p->y_dtype, p->ldy, p->D) && vec_ok(p->y2, p->y2_dtype, p->ldy2, p->D) &&
      vec_ok(p->gamma, DLSG_F32, 4, p->D) && vec_ok(p->beta, DLSG_F32, 4, p->D)) {
    const int nv = (p->D + 127) / 128;
    if (nv <= 1) return norm_fwd_vec_launch<1>(p, st);
    if (nv <= 2) return norm_fwd_vec_launch<2>(p, st);
    if (nv <= 4) return norm_fwd_vec_launch<4>(p, st);
    if (nv <= 8) return norm_fwd_vec_launch<8>(p, st);
    if (nv <= 12) return norm_fwd_vec_launch<12>(p, st);
    return norm_fwd_vec_launch<16>(p, st);
  }
  const int nw = p->D <= 2048 ? 8 : (p->D <= 4096 ? 4 : 2);
  const size_t smem = (size_t)nw * p->D * sizeof(float);
  const bool vec = vec_ok(p->x, p->x_dtype, p->ldx, p->D) && vec_ok(p->res, p->res_dtype, p->ldres, p->D) &&
                   vec_ok(p->y, p->y_dtype, p->ldy, p->D) && vec_ok(p->y2, p->y2_dtype, p->ldy2, p->D) &&
                   vec_ok(p->gamma, DLSG_F32, 4, p->D) && vec_ok(p->beta, DLSG_F32, 4, p->D);
  int64_t blocks = (p->rows + nw - 1) / nw;
  if (blocks > kNumSM * 16) blocks = kNumSM * 16;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(norm_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaFuncSetAttribute(norm_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    attr = true;
  }
  if (vec) DLSG_LAUNCH(norm_fwd_kernel<true>, (unsigned)blocks, nw * 32, smem, st, *p);
  else DLSG_LAUNCH(norm_fwd_kernel<false>, (unsigned)blocks, nw * 32, smem, st, *p);
  return check_launch("norm_fwd_kernel");
}

// ------------------------------------------------------------------------------------------- norm bwd
// Each warp owns rows; t=pre(x+res) and dxhat staged in smem; dgamma/dbeta reduced per CTA in smem, then
// one global atomicAdd per column per CTA (grid is capped, rows are looped).
template <bool VEC>
__global__ void __launch_bounds__(128)
norm_bwd_kernel(const dlsg_norm_bwd_t p) {
  pdl_prologue();
  extern __shared__ float sm[];
  const int D = p.D;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float* sdg = sm;                       // D
  float* sdb = sm + D;                   // D
  float* t = sm + 2 * (size_t)D + (size_t)w * 2 * D;
  float* dxh = t + D;
  const bool want_param = (p.dgamma != nullptr);
  if (want_param) {
    for (int c = threadIdx.x; c < D; c += blockDim.x) { sdg[c] = 0.f; sdb[c] = 0.f; }
  }
  __syncthreads();
  for (int64_t row = (int64_t)blockIdx.x * nw + w; row < p.rows; row += (int64_t)gridDim.x * nw) {
    const float mean = p.stats[row * 2], rstd = p.stats[row * 2 + 1];
    float s1 = 0.f, s2 = 0.f;
    for (int c = lane; c < D; c += 32) {
      float x = ld_as_float(p.x, p.x_dtype, row * p.ldx + c);
      if (p.res) x += ld_as_float(p.res, p.res_dtype, row * p.ldres + c);
      const float tv = (p.flags & DLSG_NORM_PRE_TANH) ? tanhf(x) : x;
      const float xh = (tv - mean) * rstd;
      float dy = ld_as_float(p.dy, p.dy_dtype, row * p.lddy + c);
      if (p.drop_p > 0.f) dy *= drop_scale(p.drop_p, p.seed, p.offset + (uint64_t)row * D + c);
      const float g = p.gamma[c];
      if (p.flags & DLSG_NORM_POST_TANH) { const float yt = tanhf(xh * g + p.beta[c]); dy *= (1.f - yt * yt); }
      if (want_param) { atomicAdd(&sdg[c], dy * xh); atomicAdd(&sdb[c], dy); }
      const float d = dy * g;
      t[c] = tv; dxh[c] = d;
      s1 += d; s2 = fmaf(d, xh, s2);
    }
    s1 = warp_sum(s1) / (float)D; s2 = warp_sum(s2) / (float)D;
    __syncwarp();
    if (p.dx) {
      for (int c = lane; c < D; c += 32) {
        const float tv = t[c], xh = (tv - mean) * rstd;
        float dx = rstd * (dxh[c] - s1 - xh * s2);
        if (p.flags & (DLSG_NORM_PRE_TANH | DLSG_NORM_IN_IS_TANH)) dx *= (1.f - tv * tv);
        const int64_t idx = row * p.lddx + c;
        if (p.dx_accum) dx += ld_as_float(p.dx, p.dx_dtype, idx);
        st_from_float(p.dx, p.dx_dtype, idx, dx);
      }
    }
    __syncwarp();
  }
  if (want_param) {
    __syncthreads();
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
      atomicAdd(&p.dgamma[c], sdg[c]);
      atomicAdd(&p.dbeta[c], sdb[c]);
    }
  }
}

// Fast path (D % 4 == 0, D <= 128*NV, aligned rows): one warp per row, each lane owns the float4 chunks
// {128*j + 4*lane}, j < NV.  The row (t = pre(x+res) and d = dl*gamma) stays in registers between the statistics
// and the dx computation (single pass over global memory); dgamma/dbeta are accumulated in a warp-private shared-
// memory slab (plain read-modify-write, no atomics), reduced across the CTA's warps once at the end, then one
// global atomicAdd per column per CTA.
template <int NV>
__global__ void __launch_bounds__(128)
norm_bwd_vec_kernel(const dlsg_norm_bwd_t p) {
  pdl_prologue();
  extern __shared__ float sm[];          // [nw][2][D]
  const int D = p.D;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const bool want_param = (p.dgamma != nullptr);
  float* mg = sm + (size_t)w * 2 * D;
  float* mb = mg + D;
  if (want_param) {
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int c = 128 * j + 4 * lane;
      if (c < D) {
        *reinterpret_cast<float4*>(mg + c) = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(mb + c) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }
  const float invD = 1.f / (float)D;
  const bool pre = p.flags & DLSG_NORM_PRE_TANH, post = p.flags & DLSG_NORM_POST_TANH;
  const bool dtanh = p.flags & (DLSG_NORM_PRE_TANH | DLSG_NORM_IN_IS_TANH);
  const bool drop = p.drop_p > 0.f;
  const float keep = drop ? 1.f / (1.f - p.drop_p) : 1.f;
  for (int64_t row = (int64_t)blockIdx.x * nw + w; row < p.rows; row += (int64_t)gridDim.x * nw) {
    const float mean = p.stats[row * 2], rstd = p.stats[row * 2 + 1];
    float4 tv[NV], dv[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int c = 128 * j + 4 * lane;
      tv[j] = make_float4(0.f, 0.f, 0.f, 0.f); dv[j] = tv[j];
      if (c < D) {
        tv[j] = RowIO::ld4(p.x, p.x_dtype, row * p.ldx + c);
        dv[j] = RowIO::ld4(p.dy, p.dy_dtype, row * p.lddy + c);
      }
    }
    if (p.res) {
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int c = 128 * j + 4 * lane;
        if (c < D) { const float4 r = RowIO::ld4(p.res, p.res_dtype, row * p.ldres + c); tv[j].x += r.x; tv[j].y += r.y; tv[j].z += r.z; tv[j].w += r.w; }
      }
    }
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int c = 128 * j + 4 * lane;
      if (c < D) {
        if (pre) { tv[j].x = tanhf(tv[j].x); tv[j].y = tanhf(tv[j].y); tv[j].z = tanhf(tv[j].z); tv[j].w = tanhf(tv[j].w); }
        if (drop) {
          const float4 m = drop_mask4(p.drop_p, keep, p.seed, p.offset + (uint64_t)row * D + c);
          dv[j].x *= m.x; dv[j].y *= m.y; dv[j].z *= m.z; dv[j].w *= m.w;
        }
        const float4 g = *reinterpret_cast<const float4*>(p.gamma + c);
        float4 xh;
        xh.x = (tv[j].x - mean) * rstd; xh.y = (tv[j].y - mean) * rstd; xh.z = (tv[j].z - mean) * rstd; xh.w = (tv[j].w - mean) * rstd;
        if (post) {
          const float4 b = *reinterpret_cast<const float4*>(p.beta + c);
          float yt;
          yt = tanhf(xh.x * g.x + b.x); dv[j].x *= (1.f - yt * yt);
          yt = tanhf(xh.y * g.y + b.y); dv[j].y *= (1.f - yt * yt);
          yt = tanhf(xh.z * g.z + b.z); dv[j].z *= (1.f - yt * yt);
          yt = tanhf(xh.w * g.w + b.w); dv[j].w *= (1.f - yt * yt);
        }
        if (want_param) {
          float4 a = *reinterpret_cast<float4*>(mg + c), bsum = *reinterpret_cast<float4*>(mb + c);
          a.x = fmaf(dv[j].x, xh.x, a.x); a.y = fmaf(dv[j].y, xh.y, a.y); a.z = fmaf(dv[j].z, xh.z, a.z); a.w = fmaf(dv[j].w, xh.w, a.w);
          bsum.x += dv[j].x; bsum.y += dv[j].y; bsum.z += dv[j].z; bsum.w += dv[j].w;
          *reinterpret_cast<float4*>(mg + c) = a; *reinterpret_cast<float4*>(mb + c) = bsum;
        }
        dv[j].x *= g.x; dv[j].y *= g.y; dv[j].z *= g.z; dv[j].w *= g.w;
        s1 += (dv[j].x + dv[j].y) + (dv[j].z + dv[j].w);
        s2 += (dv[j].x * xh.x + dv[j].y * xh.y) + (dv[j].z * xh.z + dv[j].w * xh.w);
      }
    }
    s1 = warp_sum(s1) * invD; s2 = warp_sum(s2) * invD;
    if (p.dx) {
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int c = 128 * j + 4 * lane;
        if (c < D) {
          float4 o;
          o.x = rstd * (dv[j].x - s1 - (tv[j].x - mean) * rstd * s2);
          o.y = rstd * (dv[j].y - s1 - (tv[j].y - mean) * rstd * s2);
          o.z = rstd * (dv[j].z - s1 - (tv[j].z - mean) * rstd * s2);
          o.w = rstd * (dv[j].w - s1 - (tv[j].w - mean) * rstd * s2);
          if (dtanh) {
            o.x *= (1.f - tv[j].x * tv[j].x); o.y *= (1.f - tv[j].y * tv[j].y);
            o.z *= (1.f - tv[j].z * tv[j].z); o.w *= (1.f - tv[j].w * tv[j].w);
          }
          const int64_t idx = row * p.lddx + c;
          if (p.dx_accum) { const float4 old = RowIO::ld4(p.dx, p.dx_dtype, idx); o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w; }
          RowIO::st4(p.dx, p.dx_dtype, idx, o);
        }
      }
    }
  }
  if (want_param) {
    __syncthreads();
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
      float tg = 0.f, tb = 0.f;
      for (int ww = 0; ww < nw; ++ww) { tg += sm[(size_t)ww * 2 * D + c]; tb += sm[(size_t)ww * 2 * D + D + c]; }
      atomicAdd(&p.dgamma[c], tg);
      atomicAdd(&p.dbeta[c], tb);
    }
  }
}

template <int NV>
static int norm_bwd_vec_launch(const dlsg_norm_bwd_t* p, cudaStream_t st) {
  const int nw = 4;
  const size_t smem = (size_t)nw * 2 * p->D * sizeof(float);
  int64_t blocks = (p->rows + nw - 1) / nw;
  if (blocks > kNumSM * 4) blocks = kNumSM * 4;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(norm_bwd_vec_kernel<NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, nw * 2 * 128 * NV * 4);
    attr = true;
  }
  DLSG_LAUNCH(norm_bwd_vec_kernel<NV>, (unsigned)blocks, nw * 32, smem, st, *p);
  return check_launch("norm_bwd_vec_kernel");
}

int norm_bwd_launch(const dlsg_norm_bwd_t* p, cudaStream_t st) {
  DLSG_REQUIRE(p->rows >= 0 && p->D > 0 && p->D <= 4096, "norm_bwd: bad shape rows=%lld D=%d", (long long)p->rows, p->D);
  if (p->rows == 0) return 0;
  if (norm_bwd_bf16_ok(p)) return norm_bwd_bf16_launch(p, st);
  DLSG_REQUIRE(p->dxsum == nullptr, "norm_bwd: dxsum is only produced by the streaming bf16 form (dlsg_norm_bwd_streaming)");
  const bool vec = p->D <= 2048 && vec_ok(p->x, p->x_dtype, p->ldx, p->D) && vec_ok(p->res, p->res_dtype, p->ldres, p->D) &&
                   vec_ok(p->dy, p->dy_dtype, p->lddy, p->D) && vec_ok(p->dx, p->dx_dtype, p->lddx, p->D) &&
                   vec_ok(p->gamma, DLSG_F32, 4, p->D) && vec_ok(p->beta, DLSG_F32, 4, p->D);
  if (vec) {
    const int nv = (p->D + 127) / 128;
    if (nv <= 1) return norm_bwd_vec_launch<1>(p, st);
    if (nv <= 2) return norm_bwd_vec_launch<2>(p, st);
    if (nv <= 4) return norm_bwd_vec_launch<4>(p, st);
    if (nv <= 8) return norm_bwd_vec_launch<8>(p, st);
    if (nv <= 12) return norm_bwd_vec_launch<12>(p, st);
    return norm_bwd_vec_launch<16>(p, st);
  }
  const int nw = 4;
  const size_t smem = (size_t)(2 + 2 * nw) * p->D * sizeof(float);
  int64_t blocks = (p->rows + nw - 1) / nw;
  if (blocks > kNumSM * 4) blocks = kNumSM * 4;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(norm_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 10 * 4096 * 4);
    attr = true;
  }
  DLSG_LAUNCH(norm_bwd_kernel<false>, (unsigned)blocks, nw * 32, smem, st, *p);
  return check_launch("norm_bwd_kernel");
}

// ------------------------------------------------------------------------------------------- LSTM cell
__global__ void __launch_bounds__(256)
lstm_cell_fwd_kernel(const dlsg_lstm_cell_fwd_t p) {
  pdl_prologue();
  const int64_t n = (int64_t)p.B * p.H;
  const int H = p.H;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(e / H), h = (int)(e % H);
    float g[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int64_t gi = (int64_t)b * 4 * H + (int64_t)k * H + h;
      float v = 0.f;
#pragma unroll
      for (int s = 0; s < 8; ++s)            // unrolled + predicated: all partial loads are in flight together
        if (s < p.nsplit) v += p.gates[gi + (int64_t)s * p.stride_split];
      for (int s = 8; s < p.nsplit; ++s) v += p.gates[gi + (int64_t)s * p.stride_split];
      if (p.row_bias) v += p.row_bias[(int64_t)b * p.ld_row_bias + (int64_t)k * H + h];
      if (p.bias) v += p.bias[k * H + h];
      g[k] = v;
    }
    const float ig = sigmoidf_(g[0]), fg = sigmoidf_(g[1]), gg = tanhf(g[2]), og = sigmoidf_(g[3]);
    const float c = fg * (p.c_prev ? p.c_prev[e] : 0.f) + ig * gg;
    float hval = og * tanhf(c);
    const int64_t g0 = (int64_t)b * 4 * H + h;
    p.gates[g0] = ig; p.gates[g0 + H] = fg; p.gates[g0 + 2 * (int64_t)H] = gg; p.gates[g0 + 3 * (int64_t)H] = og;
    p.c_out[e] = c;
    if (p.drop_p > 0.f) hval *= drop_scale(p.drop_p, p.seed, p.offset + (uint64_t)e);
    if (p.h_out) p.h_out[e] = hval;
    if (p.h2) st_from_float(p.h2, p.h2_dtype, (int64_t)b * p.ldh2 + h, hval);
    if (p.h3) st_from_float(p.h3, p.h3_dtype, (int64_t)b * p.ldh3 + h, hval);
  }
}

// ---- vectorised variants for the recurrent loops (BiLSTM of EncoderVisual, the critic's LSTM): 4 hidden units per thread,
// split counts as template parameters, every 128-bit load issued before its first use (the scalar kernels sum a run-time
// number of partials through one or two registers, i.e. one L2 round trip after the other).
__device__ __forceinline__ float4 ldf4(const float* q) { return *reinterpret_cast<const float4*>(q); }
__device__ __forceinline__ float4 addf4(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ void st4dt(void* base, int dt, int64_t off, float4 v) {      // off % 4 == 0, base aligned (host-checked)
  if (dt == DLSG_F32) { *reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + off) = v; return; }
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  uint2 u; u.x = *reinterpret_cast<uint32_t*>(&a); u.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(base) + off) = u;
}

template <int S>
__global__ void __launch_bounds__(256)
lstm_cell_fwd_fast(const dlsg_lstm_cell_fwd_t p) {
  pdl_prologue();
  const int H = p.H, H4 = H >> 2;
  const int64_t e4 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e4 >= (int64_t)p.B * H4) return;
  const int b = (int)(e4 / H4), h = (int)(e4 % H4) * 4;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 part[4][S], rb[4], bs[4], cp = z4;
  const float* g0p = p.gates + (int64_t)b * 4 * H + h;
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int s_ = 0; s_ < S; ++s_) part[k][s_] = ldf4(g0p + (int64_t)k * H + (int64_t)s_ * p.stride_split);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    rb[k] = p.row_bias ? ldf4(p.row_bias + (int64_t)b * p.ld_row_bias + (int64_t)k * H + h) : z4;
    bs[k] = p.bias ? ldf4(p.bias + k * H + h) : z4;
  }
  const int64_t ei = (int64_t)b * H + h;
  if (p.c_prev) cp = ldf4(p.c_prev + ei);
  float g[4][4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float4 v = addf4(rb[k], bs[k]);
#pragma unroll
    for (int s_ = 0; s_ < S; ++s_) v = addf4(v, part[k][s_]);
    g[k][0] = v.x; g[k][1] = v.y; g[k][2] = v.z; g[k][3] = v.w;
  }
  const float cp_[4] = {cp.x, cp.y, cp.z, cp.w};
  float ai[4], af[4], ag[4], ao[4], cc[4], hh[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    ai[u] = sigmoidf_(g[0][u]); af[u] = sigmoidf_(g[1][u]); ag[u] = tanhf(g[2][u]); ao[u] = sigmoidf_(g[3][u]);
    cc[u] = af[u] * cp_[u] + ai[u] * ag[u];
    hh[u] = ao[u] * tanhf(cc[u]);
    if (p.drop_p > 0.f) hh[u] *= drop_scale(p.drop_p, p.seed, p.offset + (uint64_t)(ei + u));
  }
  float* g0 = p.gates + (int64_t)b * 4 * H + h;
  *reinterpret_cast<float4*>(g0) = make_float4(ai[0], ai[1], ai[2], ai[3]);
  *reinterpret_cast<float4*>(g0 + H) = make_float4(af[0], af[1], af[2], af[3]);
  *reinterpret_cast<float4*>(g0 + 2 * (int64_t)H) = make_float4(ag[0], ag[1], ag[2], ag[3]);
  *reinterpret_cast<float4*>(g0 + 3 * (int64_t)H) = make_float4(ao[0], ao[1], ao[2], ao[3]);
  *reinterpret_cast<float4*>(p.c_out + ei) = make_float4(cc[0], cc[1], cc[2], cc[3]);
  const float4 h4 = make_float4(hh[0], hh[1], hh[2], hh[3]);
  if (p.h_out) *reinterpret_cast<float4*>(p.h_out + ei) = h4;
  if (p.h2) st4dt(p.h2, p.h2_dtype, (int64_t)b * p.ldh2 + h, h4);
  if (p.h3) st4dt(p.h3, p.h3_dtype, (int64_t)b * p.ldh3 + h, h4);
}

template <int S2>
__global__ void __launch_bounds__(256)
lstm_cell_bwd_fast(const dlsg_lstm_cell_bwd_t p) {
  pdl_prologue();
  const int H = p.H, H4 = H >> 2;
  const int64_t e4 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e4 >= (int64_t)p.B * H4) return;
  const int b = (int)(e4 / H4), h = (int)(e4 % H4) * 4;
  const int64_t ei = (int64_t)b * H + h, g0 = (int64_t)b * 4 * H + h;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 ai = ldf4(p.acts + g0), af = ldf4(p.acts + g0 + H), ag = ldf4(p.acts + g0 + 2 * (int64_t)H), ao = ldf4(p.acts + g0 + 3 * (int64_t)H);
  const float4 cn = ldf4(p.c_new + ei);
  const float4 r1 = ldf4(p.dh + (int64_t)b * p.lddh + h);
  float4 r2[S2], cp = z4, dcn = z4;
#pragma unroll
  for (int s_ = 0; s_ < S2; ++s_)      // (S2 > 4: the wide instantiation, partial count at run time - predicated, all loads still issued together)
    r2[s_] = (p.dh2 && (S2 <= 4 || s_ < p.dh2_nsplit)) ? ldf4(p.dh2 + (int64_t)b * p.lddh2 + h + (int64_t)s_ * p.dh2_stride_split) : z4;
  if (p.c_prev) cp = ldf4(p.c_prev + ei);
  if (p.dc_next) dcn = ldf4(p.dc_next + ei);
  float4 ga[4] = {z4, z4, z4, z4};
  if (p.dc_next2) dcn = addf4(dcn, ldf4(p.dc_next2 + ei));
  if (p.dgates_add) {
#pragma unroll
    for (int k = 0; k < 4; ++k) ga[k] = ldf4(p.dgates_add + g0 + (int64_t)k * H);
  }
  float4 r = r1;
#pragma unroll
  for (int s_ = 0; s_ < S2; ++s_) r = addf4(r, r2[s_]);
  if (p.dh_total) *reinterpret_cast<float4*>(p.dh_total + ei) = r;
  const float dh_[4] = {r.x, r.y, r.z, r.w};
  const float i_[4] = {ai.x, ai.y, ai.z, ai.w}, f_[4] = {af.x, af.y, af.z, af.w}, g_[4] = {ag.x, ag.y, ag.z, ag.w}, o_[4] = {ao.x, ao.y, ao.z, ao.w};
  const float cn_[4] = {cn.x, cn.y, cn.z, cn.w}, cp_[4] = {cp.x, cp.y, cp.z, cp.w}, dcn_[4] = {dcn.x, dcn.y, dcn.z, dcn.w};
  float d[4][4], dcp[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    float dh = dh_[u];
    if (p.drop_p > 0.f) dh *= drop_scale(p.drop_p, p.seed, p.offset + (uint64_t)(ei + u));
    const float tc = tanhf(cn_[u]);
    const float dc = dh * o_[u] * (1.f - tc * tc) + dcn_[u];
    d[0][u] = dc * g_[u] * i_[u] * (1.f - i_[u]);
    d[1][u] = dc * cp_[u] * f_[u] * (1.f - f_[u]);
    d[2][u] = dc * i_[u] * (1.f - g_[u] * g_[u]);
    d[3][u] = dh * tc * o_[u] * (1.f - o_[u]);
    dcp[u] = dc * f_[u];
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) { d[k][0] += ga[k].x; d[k][1] += ga[k].y; d[k][2] += ga[k].z; d[k][3] += ga[k].w; }
  if (p.dc_prev) *reinterpret_cast<float4*>(p.dc_prev + ei) = make_float4(dcp[0], dcp[1], dcp[2], dcp[3]);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int64_t col = (int64_t)k * H + h;
    const float4 d4 = make_float4(d[k][0], d[k][1], d[k][2], d[k][3]);
    if (p.dgates) *reinterpret_cast<float4*>(p.dgates + (int64_t)b * (p.ld_dgates ? p.ld_dgates : 4 * (int64_t)H) + col) = d4;
    if (p.dgates2) st4dt(p.dgates2, p.dgates2_dtype, (int64_t)b * p.ld_dgates2 + col, d4);
    if (p.dgatesT) {
#pragma unroll
      for (int u = 0; u < 4; ++u) st_from_float(p.dgatesT, p.dgatesT_dtype, (col + u) * p.ld_dgatesT + b, d[k][u]);
    }
  }
}

__global__ void __launch_bounds__(256)
lstm_cell_bwd_kernel(const dlsg_lstm_cell_bwd_t p) {
  pdl_prologue();
  const int64_t n = (int64_t)p.B * p.H;
  const int H = p.H;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(e / H), h = (int)(e % H);
    const int64_t g0 = (int64_t)b * 4 * H + h;
    const float ig = p.acts[g0], fg = p.acts[g0 + H], gg = p.acts[g0 + 2 * (int64_t)H], og = p.acts[g0 + 3 * (int64_t)H];
    float dh = p.dh[(int64_t)b * p.lddh + h];
    if (p.dh2) {
      const float* d2 = p.dh2 + (int64_t)b * p.lddh2 + h;
      float v = d2[0];
#pragma unroll
      for (int s = 1; s < 16; ++s)              // unrolled + predicated: all partial loads in flight together
        if (s < p.dh2_nsplit) v += d2[(int64_t)s * p.dh2_stride_split];
      dh += v;
    }
    if (p.dh_total) p.dh_total[e] = dh;
    if (p.drop_p > 0.f) dh *= drop_scale(p.drop_p, p.seed, p.offset + (uint64_t)e);
    const float tc = tanhf(p.c_new[e]);
    float dc = dh * og * (1.f - tc * tc);
    if (p.dc_next) dc += p.dc_next[e];
    if (p.dc_next2) dc += p.dc_next2[e];
    const float cp = p.c_prev ? p.c_prev[e] : 0.f;
    float d[4];
    d[0] = dc * gg * ig * (1.f - ig);
    d[1] = dc * cp * fg * (1.f - fg);
    d[2] = dc * ig * (1.f - gg * gg);
    d[3] = dh * tc * og * (1.f - og);
    if (p.dc_prev) p.dc_prev[e] = dc * fg;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int64_t col = (int64_t)k * H + h;
      if (p.dgates_add) d[k] += p.dgates_add[(int64_t)b * 4 * H + col];
      if (p.dgates) p.dgates[(int64_t)b * (p.ld_dgates ? p.ld_dgates : 4 * (int64_t)H) + col] = d[k];
      if (p.dgates2) st_from_float(p.dgates2, p.dgates2_dtype, (int64_t)b * p.ld_dgates2 + col, d[k]);
      if (p.dgatesT) st_from_float(p.dgatesT, p.dgatesT_dtype, col * p.ld_dgatesT + b, d[k]);
    }
  }
}

// ------------------------------------------------------------------------------------------- LayerNorm backward, differentiated
// Backward of dx = rstd * P(dy*gamma), P(v) = v - mean(v) - xh * mean(v * xh), xh = (x - mean) * rstd, with respect to
// (dy, x, gamma) for a cotangent u of dx (the WGAN-GP penalty differentiates the critic's input gradient, run_gun.py:362-375):
//   g_dy = rstd * P(u) * gamma            g_gamma += sum_rows rstd * P(u) * dy
//   g_x  = rstd * P(q) - (K rstd^2 / D) xh,   q = -rstd (u c_g + c_u g),  K = <u,g> - D mean(u) mean(g) - D c_u c_g,
//   c_v = mean(v * xh), g = dy * gamma.   One warp per row, lanes stride over columns (D <= 32 * KC).
template <int KC>
__global__ void __launch_bounds__(256)
norm_bwd2_kernel(const dlsg_norm_bwd2_t p) {
  pdl_prologue();
  extern __shared__ float sm2[];                 // [nw][D] partial g_gamma
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int D = p.D;
  const float invD = 1.f / (float)D;
  float gam[KC], acc[KC];
#pragma unroll
  for (int k = 0; k < KC; ++k) {
    const int c = lane + 32 * k;
    gam[k] = c < D ? p.gamma[c] : 0.f;
    acc[k] = 0.f;
  }
  for (int64_t row = (int64_t)blockIdx.x * nw + w; row < p.rows; row += (int64_t)gridDim.x * nw) {
    const float mean = p.stats[row * 2], r = p.stats[row * 2 + 1];
    float a[KC], uu[KC], dv[KC];
    float s_u = 0.f, s_ua = 0.f, s_g = 0.f, s_ga = 0.f, s_ug = 0.f;
#pragma unroll
    for (int k = 0; k < KC; ++k) {
      const int c = lane + 32 * k;
      a[k] = 0.f; uu[k] = 0.f; dv[k] = 0.f;
      if (c < D) {
        a[k] = (p.x[row * D + c] - mean) * r;
        uu[k] = p.u[row * D + c];
        dv[k] = p.dy[row * D + c];
      }
      const float g = dv[k] * gam[k];
      s_u += uu[k]; s_ua = fmaf(uu[k], a[k], s_ua); s_g += g; s_ga = fmaf(g, a[k], s_ga); s_ug = fmaf(uu[k], g, s_ug);
    }
    s_u = warp_sum(s_u); s_ua = warp_sum(s_ua); s_g = warp_sum(s_g); s_ga = warp_sum(s_ga); s_ug = warp_sum(s_ug);
    const float mu_u = s_u * invD, c_u = s_ua * invD, mu_g = s_g * invD, c_g = s_ga * invD;
    const float K = s_ug - (float)D * (mu_u * mu_g + c_u * c_g);
    const float mq = -r * (mu_u * c_g + c_u * mu_g), mqa = -2.f * r * c_u * c_g, kx = K * r * r * invD;
#pragma unroll
    for (int k = 0; k < KC; ++k) {
      const int c = lane + 32 * k;
      if (c < D) {
        const float pu = uu[k] - mu_u - a[k] * c_u;
        const float g = dv[k] * gam[k];
        if (p.g_dy) p.g_dy[row * D + c] = r * pu * gam[k];
        acc[k] = fmaf(r * pu, dv[k], acc[k]);
        const float q = -r * (uu[k] * c_g + c_u * g);
        if (p.g_x) p.g_x[row * D + c] = r * (q - mq - a[k] * mqa) - kx * a[k];
      }
    }
  }
  if (p.g_gamma == nullptr) return;              // uniform
#pragma unroll
  for (int k = 0; k < KC; ++k) {
    const int c = lane + 32 * k;
    if (c < D) sm2[w * D + c] = acc[k];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float t = 0.f;
    for (int ww = 0; ww < nw; ++ww) t += sm2[ww * D + c];
    atomicAdd(&p.g_gamma[c], t);
  }
}

// ------------------------------------------------------------------------------------------- multi-segment convert
// One launch refreshes every GEMM-operand copy of the parameters after an optimizer step: a device table of 2-D
// segments dst[r,c] = cast(src[r,c] (+ src2[r,c])) (fp32 sources; bf16 or fp32 destinations with their own pitch: row
// slices of stacked weights, column segments of the packed LSTM gate matrices, summed bias pairs) and a table of
// (segment, first row, rows) chunks of ~16K elements, one CTA each.
__device__ __forceinline__ void convert_rows(const dlsg_seg_t& sg, const int64_t row0, const int nrows) {
  if (sg.src_dtype == DLSG_BF16) {               // bf16 source (a reduced gradient bucket read back as fp32): no src2
    const __nv_bfloat16* s16 = static_cast<const __nv_bfloat16*>(sg.src) + row0 * sg.ld_src;
    const bool v4 = (sg.cols % 4 == 0) && (sg.ld_src % 4 == 0) && (sg.ld_dst % 4 == 0) && sg.dst_dtype == DLSG_F32 &&
                    ((reinterpret_cast<uintptr_t>(sg.src) & 7) == 0) && ((reinterpret_cast<uintptr_t>(sg.dst) & 15) == 0);
    if (v4) {
      const int c4 = (int)(sg.cols >> 2);
      const int64_t n = (int64_t)nrows * c4;
      for (int64_t e = threadIdx.x; e < n; e += blockDim.x) {
        const int64_t r = e / c4;
        const int c = (int)(e % c4) * 4;
        const uint2 raw = *reinterpret_cast<const uint2*>(s16 + r * sg.ld_src + c);
        const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.x));
        const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.y));
        *reinterpret_cast<float4*>(static_cast<float*>(sg.dst) + (row0 + r) * sg.ld_dst + c) = make_float4(a.x, a.y, b.x, b.y);
      }
    } else {
      const int64_t n = (int64_t)nrows * sg.cols;
      for (int64_t e = threadIdx.x; e < n; e += blockDim.x) {
        const int64_t r = e / sg.cols, c = e % sg.cols;
        st_from_float(sg.dst, sg.dst_dtype, (row0 + r) * sg.ld_dst + c, __bfloat162float(s16[r * sg.ld_src + c]));
      }
    }
    return;
  }
  const float* src = static_cast<const float*>(sg.src) + row0 * sg.ld_src;
  const float* src2 = sg.src2 ? static_cast<const float*>(sg.src2) + row0 * sg.ld_src : nullptr;
  const bool bf = sg.dst_dtype == DLSG_BF16;
  const bool vec = (sg.cols % 4 == 0) && (sg.ld_src % 4 == 0) && (sg.ld_dst % 4 == 0) &&
                   ((reinterpret_cast<uintptr_t>(sg.src) & 15) == 0) && (!sg.src2 || (reinterpret_cast<uintptr_t>(sg.src2) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(sg.dst) & (bf ? 7 : 15)) == 0);
  if (vec) {
    const int c4 = (int)(sg.cols >> 2);
    const int64_t n = (int64_t)nrows * c4;
    for (int64_t e = threadIdx.x; e < n; e += blockDim.x) {
      const int64_t r = e / c4;
      const int c = (int)(e % c4) * 4;
      float4 v = *reinterpret_cast<const float4*>(src + r * sg.ld_src + c);
      if (src2) {
        const float4 w = *reinterpret_cast<const float4*>(src2 + r * sg.ld_src + c);
        v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
      }
      if (bf) {
        __nv_bfloat16* d = static_cast<__nv_bfloat16*>(sg.dst) + (row0 + r) * sg.ld_dst + c;
        __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
        uint2 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&lo); pk.y = *reinterpret_cast<uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(d) = pk;
      } else {
        *reinterpret_cast<float4*>(static_cast<float*>(sg.dst) + (row0 + r) * sg.ld_dst + c) = v;
      }
    }
  } else {
    const int64_t n = (int64_t)nrows * sg.cols;
    for (int64_t e = threadIdx.x; e < n; e += blockDim.x) {
      const int64_t r = e / sg.cols, c = e % sg.cols;
      float v = src[r * sg.ld_src + c];
      if (src2) v += src2[r * sg.ld_src + c];
      st_from_float(sg.dst, sg.dst_dtype, (row0 + r) * sg.ld_dst + c, v);
    }
  }
}

__global__ void __launch_bounds__(256)
multi_convert_kernel(const dlsg_seg_t* __restrict__ segs, const int32_t* __restrict__ chunks) {
  pdl_prologue();
  const dlsg_seg_t sg = segs[chunks[3 * blockIdx.x]];
  convert_rows(sg, chunks[3 * blockIdx.x + 1], chunks[3 * blockIdx.x + 2]);
}

// Same work with the segment table passed BY VALUE in the kernel parameters (like adam_multi_kernel): nothing is uploaded,
// so it can be issued with fresh pointers in the middle of a CUDA-graph capture (the per-block gradient packs of a
// data-parallel step, dlsg.functional.GradSync).
constexpr int CONV_MAX_SEGS = 384;
struct ConvArgs {
  dlsg_seg_t seg[CONV_MAX_SEGS];
  int32_t chunk_start[CONV_MAX_SEGS + 1];
  int32_t rows_per_chunk[CONV_MAX_SEGS];
  int32_t nsegs;
};

__global__ void __launch_bounds__(256)
multi_convert_args_kernel(const __grid_constant__ ConvArgs a) {
  pdl_prologue();
  int lo = 0, hi = a.nsegs - 1;
  while (lo < hi) {
    const int md = (lo + hi + 1) >> 1;
    if (a.chunk_start[md] <= (int)blockIdx.x) lo = md; else hi = md - 1;
  }
  const dlsg_seg_t& sg = a.seg[lo];
  const int64_t row0 = (int64_t)((int)blockIdx.x - a.chunk_start[lo]) * a.rows_per_chunk[lo];
  convert_rows(sg, row0, (int)min((int64_t)a.rows_per_chunk[lo], sg.rows - row0));
}

// ------------------------------------------------------------------------------------------- multi-tensor Adam
// torch.optim.Adam's update (run_gun.py:91: lr 1.6e-4, betas (0.5, 0.9), no weight decay / amsgrad) over a device table of
// 2-D segments, one launch per parameter block, writing the bf16 GEMM-operand copy of each weight in the same pass:
//   m = m + (g - m)(1 - b1);  v = b2 v + (1 - b2) g^2;  p -= (lr / (1 - b1^t)) m / (sqrt(v) / sqrt(1 - b2^t) + eps)
constexpr int ADAM_MAX_SEGS = 256;
struct AdamArgs {                       // passed BY VALUE as a kernel parameter (a CUDA-graph kernel node keeps it: no table upload)
  dlsg_adam_seg_t seg[ADAM_MAX_SEGS];
  int32_t chunk_start[ADAM_MAX_SEGS + 1];   // first chunk (CTA) of each segment
  int32_t rows_per_chunk[ADAM_MAX_SEGS];
  int32_t nsegs;
};

__global__ void __launch_bounds__(256)
adam_multi_kernel(const __grid_constant__ AdamArgs a, const float* __restrict__ step, const float* __restrict__ lr_dev, float lr_host,
                  float beta1, float beta2, float eps) {
  pdl_prologue();
  __shared__ float sh[2];
  int lo = 0, hi = a.nsegs - 1;                     // segment of this CTA: last one whose first chunk <= blockIdx.x
  while (lo < hi) {
    const int md = (lo + hi + 1) >> 1;
    if (a.chunk_start[md] <= (int)blockIdx.x) lo = md; else hi = md - 1;
  }
  const dlsg_adam_seg_t& sg = a.seg[lo];
  if (threadIdx.x == 0) {
    // bias correction from THIS segment's step counter when it has one (torch keeps one `step` per parameter: they
    // diverge for a parameter that was frozen for a while or restored from another checkpoint), else the launch-wide one
    const double t = (double)(sg.step ? *sg.step : *step);
    const float lr = lr_dev ? *lr_dev : lr_host;
    const double bc1 = 1.0 - pow((double)beta1, t), bc2 = 1.0 - pow((double)beta2, t);
    sh[0] = (float)((double)lr / bc1);
    sh[1] = (float)sqrt(bc2);
  }
  __syncthreads();
  const float step_size = sh[0], bc2_sqrt = sh[1];
  const int64_t row0 = (int64_t)((int)blockIdx.x - a.chunk_start[lo]) * a.rows_per_chunk[lo];
  const int nrows = (int)min((int64_t)a.rows_per_chunk[lo], sg.rows - row0);
  float* P = sg.p + row0 * sg.ld;
  const bool g16 = sg.g_dtype == DLSG_BF16;
  const float* G = static_cast<const float*>(sg.g) + (g16 ? 0 : row0 * sg.ld_g);
  const __nv_bfloat16* G16 = static_cast<const __nv_bfloat16*>(sg.g) + (g16 ? row0 * sg.ld_g : 0);
  float* M = sg.m + row0 * sg.ld;
  float* V = sg.v + row0 * sg.ld;
  __nv_bfloat16* S16 = sg.dst16 ? static_cast<__nv_bfloat16*>(sg.dst16) + row0 * sg.ld_dst : nullptr;
  const float omb1 = 1.f - beta1, omb2 = 1.f - beta2;
  const bool vec = (sg.cols % 4 == 0) && (sg.ld % 4 == 0) && (sg.ld_g % 4 == 0) &&
                   (((reinterpret_cast<uintptr_t>(sg.p) | reinterpret_cast<uintptr_t>(sg.m) | reinterpret_cast<uintptr_t>(sg.v)) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(sg.g) & (g16 ? 7 : 15)) == 0) &&
                   (!sg.dst16 || ((sg.ld_dst % 4 == 0) && (reinterpret_cast<uintptr_t>(sg.dst16) & 7) == 0));
  if (vec) {
    const int c4 = (int)(sg.cols >> 2);
    const int64_t n = (int64_t)nrows * c4;
    for (int64_t e = threadIdx.x; e < n; e += blockDim.x) {
      const int64_t r = e / c4;
      const int c = (int)(e % c4) * 4;
      const int64_t o = r * sg.ld + c;
      float4 p4 = *reinterpret_cast<const float4*>(P + o);
      float4 g4;
      if (g16) {
        const uint2 raw = *reinterpret_cast<const uint2*>(G16 + r * sg.ld_g + c);
        const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.x));
        const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.y));
        g4 = make_float4(a.x, a.y, b.x, b.y);
      } else {
        g4 = *reinterpret_cast<const float4*>(G + r * sg.ld_g + c);
      }
      float4 m4 = *reinterpret_cast<const float4*>(M + o), v4 = *reinterpret_cast<const float4*>(V + o);
      float pp[4] = {p4.x, p4.y, p4.z, p4.w}, gg[4] = {g4.x, g4.y, g4.z, g4.w};
      float mm[4] = {m4.x, m4.y, m4.z, m4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        mm[u] = mm[u] + (gg[u] - mm[u]) * omb1;
        vv[u] = beta2 * vv[u] + omb2 * gg[u] * gg[u];
        pp[u] -= step_size * mm[u] / (sqrtf(vv[u]) / bc2_sqrt + eps);
      }
      *reinterpret_cast<float4*>(P + o) = make_float4(pp[0], pp[1], pp[2], pp[3]);
      *reinterpret_cast<float4*>(M + o) = make_float4(mm[0], mm[1], mm[2], mm[3]);
      *reinterpret_cast<float4*>(V + o) = make_float4(vv[0], vv[1], vv[2], vv[3]);
      if (S16) {
        __nv_bfloat162 lo = __floats2bfloat162_rn(pp[0], pp[1]), hi = __floats2bfloat162_rn(pp[2], pp[3]);
        uint2 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&lo); pk.y = *reinterpret_cast<uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(S16 + r * sg.ld_dst + c) = pk;
      }
    }
  } else {
    const int64_t n = (int64_t)nrows * sg.cols;
    for (int64_t e = threadIdx.x; e < n; e += blockDim.x) {
      const int64_t r = e / sg.cols, c = e % sg.cols;
      const int64_t o = r * sg.ld + c;
      const float g = g16 ? __bfloat162float(G16[r * sg.ld_g + c]) : G[r * sg.ld_g + c];
      const float m = M[o] + (g - M[o]) * omb1;
      const float v = beta2 * V[o] + omb2 * g * g;
      const float pnew = P[o] - step_size * m / (sqrtf(v) / bc2_sqrt + eps);
      P[o] = pnew; M[o] = m; V[o] = v;
      if (S16) S16[r * sg.ld_dst + c] = __float2bfloat16_rn(pnew);
    }
  }
}

// Backward OF the cell backward (WGAN-GP double backward through the discriminator LSTM, run_gun.py:362-375).
// The cell backward maps (dh, dc; a=pre-activations, c0) -> (dpre[4], dc0) with D = dh*o*(1-tc^2) + dc:
//   dpre_i = D*g*i(1-i)  dpre_f = D*c0*f(1-f)  dpre_g = D*i*(1-g^2)  dpre_o = dh*tc*o(1-o)  dc0 = D*f .
// Given cotangents u[4] (for dpre) and w (for dc0) this kernel returns the cotangents of dh, dc, a[4] and c0
// (c = f*c0 + i*g is treated as a function of a and c0).
__global__ void __launch_bounds__(256)
lstm_cell_bwd2_kernel(const dlsg_lstm_cell_bwd2_t p) {
  pdl_prologue();
  const int64_t n = (int64_t)p.B * p.H;
  const int H = p.H;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(e / H), h = (int)(e % H);
    const int64_t g0 = (int64_t)b * 4 * H + h;
    const float ig = p.acts[g0], fg = p.acts[g0 + H], gg = p.acts[g0 + 2 * (int64_t)H], og = p.acts[g0 + 3 * (int64_t)H];
    const float dh = p.dh[e];
    const float dcn = p.dc_next ? p.dc_next[e] : 0.f;
    const float c0 = p.c_prev ? p.c_prev[e] : 0.f;
    const float tc = tanhf(p.c_new[e]);
    const float omt = 1.f - tc * tc;
    float ui = 0.f, uf = 0.f, ug = 0.f, uo = 0.f;
    if (p.u) {
      const float* q = p.u + (p.ld_u ? (int64_t)b * p.ld_u + h : g0);
      ui = q[0]; uf = q[H]; ug = q[2 * (int64_t)H]; uo = q[3 * (int64_t)H];
    }
    if (p.u2) {
      const int ns = p.u2_nsplit > 1 ? p.u2_nsplit : 1;
      for (int s_ = 0; s_ < ns; ++s_) {
        const float* q = p.u2 + (int64_t)s_ * p.u2_stride_split + g0;
        ui += q[0]; uf += q[H]; ug += q[2 * (int64_t)H]; uo += q[3 * (int64_t)H];
      }
    }
    const float w = p.w ? p.w[e] : 0.f;
    const float si = ig * (1.f - ig), sf = fg * (1.f - fg), sg = 1.f - gg * gg, so = og * (1.f - og);
    const float D = dh * og * omt + dcn;
    const float S = ui * gg * si + uf * c0 * sf + ug * ig * sg + w * fg;       // d(L2)/dD
    const float Q = S * dh * og * (-2.f * tc * omt) + uo * dh * so * omt;       // d(L2)/dc through tc
    const float gdh = S * og * omt + uo * tc * so;
    if (p.g_dh) p.g_dh[p.ld_g_dh ? (int64_t)b * p.ld_g_dh + h : e] = gdh;
    if (p.g_dh2) st_from_float(p.g_dh2, p.g_dh2_dtype, (int64_t)b * p.ld_g_dh2 + h, gdh);
    if (p.g_dc) p.g_dc[e] = S;
    if (p.g_cprev) p.g_cprev[e] = D * uf * sf + Q * fg;
    if (p.g_pre) {
      p.g_pre[g0] = D * (ui * gg * si * (1.f - 2.f * ig) + ug * sg * si) + Q * gg * si;
      p.g_pre[g0 + H] = D * (uf * c0 * sf * (1.f - 2.f * fg) + w * sf) + Q * c0 * sf;
      p.g_pre[g0 + 2 * (int64_t)H] = D * (ui * si * sg + ug * ig * (-2.f * gg) * sg) + Q * ig * sg;
      p.g_pre[g0 + 3 * (int64_t)H] = S * dh * omt * so + uo * dh * tc * so * (1.f - 2.f * og);
    }
  }
}

// ------------------------------------------------------------------------------------------- axis softmax
// one warp per (outer, inner) pair, lanes stride over the softmax axis
__global__ void __launch_bounds__(256)
softmax_fwd_kernel(const dlsg_softmax_t p) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int64_t pairs = p.outer * p.inner;
  for (int64_t pr = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; pr < pairs; pr += ((int64_t)gridDim.x * blockDim.x) >> 5) {
    const int64_t base = (pr / p.inner) * p.so + (pr % p.inner) * p.si;
    float mx = -INFINITY;
    for (int64_t j = lane; j < p.n; j += 32) {
      float v = p.x[base + j * p.sn] * p.scale;
      if (p.mask_mode == 1 && !(p.mask[base + j * p.sn] > 0.f)) v = -9e15f;
      mx = fmaxf(mx, v);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int64_t j = lane; j < p.n; j += 32) {
      float v = p.x[base + j * p.sn] * p.scale;
      if (p.mask_mode == 1 && !(p.mask[base + j * p.sn] > 0.f)) v = -9e15f;
      sum += expf(v - mx);
    }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    for (int64_t j = lane; j < p.n; j += 32) {
      float v = p.x[base + j * p.sn] * p.scale;
      if (p.mask_mode == 1 && !(p.mask[base + j * p.sn] > 0.f)) v = -9e15f;
      float y = expf(v - mx) * inv;
      if (p.mask_mode == 2 && !(p.mask[base + j * p.sn] > 0.f)) y = 0.f;
      p.y[base + j * p.sn] = y;
    }
  }
}

__global__ void __launch_bounds__(256)
softmax_bwd_kernel(const dlsg_softmax_t p, const float* __restrict__ dy, float* __restrict__ dx) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int64_t pairs = p.outer * p.inner;
  for (int64_t pr = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; pr < pairs; pr += ((int64_t)gridDim.x * blockDim.x) >> 5) {
    const int64_t base = (pr / p.inner) * p.so + (pr % p.inner) * p.si;
    float mx = -INFINITY;
    for (int64_t j = lane; j < p.n; j += 32) {
      float v = p.x[base + j * p.sn] * p.scale;
      if (p.mask_mode == 1 && !(p.mask[base + j * p.sn] > 0.f)) v = -9e15f;
      mx = fmaxf(mx, v);
    }
    mx = warp_max(mx);
    float sum = 0.f, dot = 0.f;
    for (int64_t j = lane; j < p.n; j += 32) {
      const int64_t i = base + j * p.sn;
      float v = p.x[i] * p.scale;
      if (p.mask_mode == 1 && !(p.mask[i] > 0.f)) v = -9e15f;
      const float e = expf(v - mx);
      float g = dy[i];
      if (p.mask_mode == 2 && !(p.mask[i] > 0.f)) g = 0.f;
      sum += e; dot = fmaf(e, g, dot);
    }
    sum = warp_sum(sum); dot = warp_sum(dot);
    const float inv = 1.f / sum;
    dot *= inv;
    for (int64_t j = lane; j < p.n; j += 32) {
      const int64_t i = base + j * p.sn;
      float v = p.x[i] * p.scale;
      const bool masked_pre = (p.mask_mode == 1 && !(p.mask[i] > 0.f));
      if (masked_pre) v = -9e15f;
      const float s = expf(v - mx) * inv;
      float g = dy[i];
      if (p.mask_mode == 2 && !(p.mask[i] > 0.f)) g = 0.f;
      dx[i] = masked_pre ? 0.f : p.scale * s * (g - dot);
    }
  }
}

// Backward of softmax_bwd_kernel with respect to (dy, x): see dlsg_softmax_bwd2 in dlsg.h.  One warp per (outer, inner) pair.
__global__ void __launch_bounds__(256)
softmax_bwd2_kernel(const dlsg_softmax_t p, const float* __restrict__ dy, const float* __restrict__ u, float* __restrict__ g_dy,
                    float* __restrict__ g_x) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int64_t pairs = p.outer * p.inner;
  for (int64_t pr = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; pr < pairs; pr += ((int64_t)gridDim.x * blockDim.x) >> 5) {
    const int64_t base = (pr / p.inner) * p.so + (pr % p.inner) * p.si;
    float mx = -INFINITY;
    for (int64_t j = lane; j < p.n; j += 32) {
      float v = p.x[base + j * p.sn] * p.scale;
      if (p.mask_mode == 1 && !(p.mask[base + j * p.sn] > 0.f)) v = -9e15f;
      mx = fmaxf(mx, v);
    }
    mx = warp_max(mx);
    float sum = 0.f, A = 0.f, Bq = 0.f;
    for (int64_t j = lane; j < p.n; j += 32) {
      const int64_t i = base + j * p.sn;
      const bool keep = p.mask_mode == 0 || p.mask[i] > 0.f;
      float v = p.x[i] * p.scale;
      if (p.mask_mode == 1 && !keep) v = -9e15f;
      const float e = expf(v - mx);
      const float g = (p.mask_mode == 2 && !keep) ? 0.f : dy[i];
      const float uu = (p.mask_mode == 1 && !keep) ? 0.f : u[i];
      sum += e; A = fmaf(e, g, A); Bq = fmaf(e, uu, Bq);
    }
    sum = warp_sum(sum); A = warp_sum(A); Bq = warp_sum(Bq);
    const float inv = 1.f / sum;
    A *= inv; Bq *= inv;
    float Cq = 0.f;
    for (int64_t j = lane; j < p.n; j += 32) {
      const int64_t i = base + j * p.sn;
      const bool keep = p.mask_mode == 0 || p.mask[i] > 0.f;
      float v = p.x[i] * p.scale;
      if (p.mask_mode == 1 && !keep) v = -9e15f;
      const float s = expf(v - mx) * inv;
      const float g = (p.mask_mode == 2 && !keep) ? 0.f : dy[i];
      const float uu = (p.mask_mode == 1 && !keep) ? 0.f : u[i];
      Cq = fmaf(s, p.scale * (uu * (g - A) - g * Bq), Cq);
    }
    Cq = warp_sum(Cq);
    for (int64_t j = lane; j < p.n; j += 32) {
      const int64_t i = base + j * p.sn;
      const bool keep = p.mask_mode == 0 || p.mask[i] > 0.f;
      float v = p.x[i] * p.scale;
      if (p.mask_mode == 1 && !keep) v = -9e15f;
      const float s = expf(v - mx) * inv;
      const float g = (p.mask_mode == 2 && !keep) ? 0.f : dy[i];
      const float uu = (p.mask_mode == 1 && !keep) ? 0.f : u[i];
      const float q = p.scale * (uu * (g - A) - g * Bq);
      if (g_dy) g_dy[i] = (p.mask_mode == 2 && !keep) ? 0.f : p.scale * s * (uu - Bq);
      if (g_x) g_x[i] = (p.mask_mode == 1 && !keep) ? 0.f : p.scale * s * (q - Cq);
    }
  }
}

// Small fused element-wise forms (dlsg_ew in dlsg.h): one float4 (or one element) per thread and trip.
template <int W>
__device__ __forceinline__ void ew_apply(const dlsg_ew_t& p, int64_t i) {
  float a[5][W], o[3][W];
  const int nin = p.op == DLSG_EW_MUL_BWD2 ? 5 : (p.op == DLSG_EW_TANH_BWD || p.op == DLSG_EW_LERP_ROWS_BWD ? 2 : 3);
  const bool rows_e = p.op == DLSG_EW_LERP_ROWS || p.op == DLSG_EW_LERP_ROWS_BWD;
  const int e_slot = p.op == DLSG_EW_LERP_ROWS ? 2 : 1;
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    if (k >= nin) continue;
    if (rows_e && k == e_slot) {
      const float e = p.in[k][i / p.cols];                 // W consecutive elements share a row (cols % W == 0)
#pragma unroll
      for (int w = 0; w < W; ++w) a[k][w] = e;
    } else if constexpr (W == 4) {
      const float4 v = *reinterpret_cast<const float4*>(p.in[k] + i);
      a[k][0] = v.x; a[k][1] = v.y; a[k][2] = v.z; a[k][3] = v.w;
    } else {
      a[k][0] = p.in[k][i];
    }
  }
  int nout = 2;
#pragma unroll
  for (int w = 0; w < W; ++w) {
    switch (p.op) {
      case DLSG_EW_TANH_BWD: o[0][w] = a[0][w] * (1.f - a[1][w] * a[1][w]); nout = 1; break;
      case DLSG_EW_TANH_BWD2:
        o[0][w] = a[2][w] * (1.f - a[1][w] * a[1][w]);
        o[1][w] = -2.f * a[1][w] * a[0][w] * a[2][w];
        break;
      case DLSG_EW_MUL_BWD: o[0][w] = a[0][w] * a[2][w]; o[1][w] = a[0][w] * a[1][w]; break;
      case DLSG_EW_MUL_BWD2:
        o[0][w] = a[3][w] * a[2][w] + a[4][w] * a[1][w];
        o[1][w] = a[4][w] * a[0][w];
        o[2][w] = a[3][w] * a[0][w];
        nout = 3;
        break;
      case DLSG_EW_LERP_ROWS: o[0][w] = a[0][w] * a[2][w] + a[1][w] * (1.f - a[2][w]); nout = 1; break;
      default: o[0][w] = a[0][w] * a[1][w]; o[1][w] = a[0][w] * (1.f - a[1][w]); break;      // DLSG_EW_LERP_ROWS_BWD
    }
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    if (k >= nout || !p.out[k]) continue;
    if constexpr (W == 4) *reinterpret_cast<float4*>(p.out[k] + i) = make_float4(o[k][0], o[k][1], o[k][2], o[k][3]);
    else p.out[k][i] = o[k][0];
  }
}

template <int W>
__global__ void __launch_bounds__(256)
ew_kernel(const dlsg_ew_t p) {
  pdl_prologue();
  const int64_t nv = p.n / W;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nv; e += (int64_t)gridDim.x * blockDim.x) ew_apply<W>(p, e * W);
}

// ------------------------------------------------------------------------------------------- embedding etc.
__global__ void __launch_bounds__(128)
embedding_gather_kernel(const float* __restrict__ table, const int64_t* __restrict__ ids, int64_t ld_ids, int rows, int W,
                        void* out, int odt, int64_t ldo, void* out2, int odt2, int64_t ldo2, float drop_p, uint64_t seed,
                        uint64_t offset) {
  pdl_prologue();
  const int r = blockIdx.x;
  if (r >= rows) return;
  const int64_t id = ids[(int64_t)r * ld_ids];
  for (int c = threadIdx.x; c < W; c += blockDim.x) {
    float v = table[id * W + c];
    if (drop_p > 0.f) v *= drop_scale(drop_p, seed, offset + (uint64_t)r * W + c);
    if (out) st_from_float(out, odt, (int64_t)r * ldo + c, v);
    if (out2) st_from_float(out2, odt2, (int64_t)r * ldo2 + c, v);
  }
}
__global__ void __launch_bounds__(128)
embedding_scatter_kernel(float* __restrict__ dtable, const int64_t* __restrict__ ids, int64_t ld_ids, int rows, int W,
                         const float* __restrict__ dout, int64_t lddo, float drop_p, uint64_t seed, uint64_t offset) {
  pdl_prologue();
  const int r = blockIdx.x;
  if (r >= rows) return;
  const int64_t id = ids[(int64_t)r * ld_ids];
  for (int c = threadIdx.x; c < W; c += blockDim.x) {
    float v = dout[(int64_t)r * lddo + c];
    if (drop_p > 0.f) v *= drop_scale(drop_p, seed, offset + (uint64_t)r * W + c);
    atomicAdd(&dtable[id * W + c], v);
  }
}
__global__ void mean_nodes_fwd_kernel(const float* __restrict__ x, int B, int P, int H, float* __restrict__ y, int64_t ldy) {
  pdl_prologue();
  const int64_t n = (int64_t)B * H;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(e / H), h = (int)(e % H);
    float s = 0.f;
    for (int q = 0; q < P; ++q) s += x[((int64_t)b * P + q) * H + h];
    y[(int64_t)b * ldy + h] = s / (float)P;
  }
}
__global__ void mean_nodes_bwd_kernel(const float* __restrict__ dy, int64_t lddy, int B, int P, int H, float* __restrict__ dx) {
  pdl_prologue();
  const int64_t n = (int64_t)B * P * H;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const int h = (int)(e % H);
    const int b = (int)(e / ((int64_t)P * H));
    dx[e] += dy[(int64_t)b * lddy + h] / (float)P;
  }
}
__global__ void axpby_kernel(const float* __restrict__ x, float a, float* __restrict__ y, float b, int64_t n) {
  pdl_prologue();
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x)
    y[e] = (b == 0.f) ? a * x[e] : fmaf(a, x[e], b * y[e]);
}
__global__ void add_rowbcast_kernel(const float* __restrict__ x, const float* __restrict__ pe, float* __restrict__ y, int64_t batch, int64_t inner,
                                    float drop_p, uint64_t seed, uint64_t offset) {
  pdl_prologue();
  const int64_t n = batch * inner;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    float v = x[e] + pe[e % inner];
    if (drop_p > 0.f) v *= drop_scale(drop_p, seed, offset + (uint64_t)e);
    y[e] = v;
  }
}
// y = x * dropmask (inverted dropout); also its own backward (same mask)
__global__ void dropout_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n, float drop_p, uint64_t seed, uint64_t offset) {
  pdl_prologue();
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x)
    y[e] = x[e] * drop_scale(drop_p, seed, offset + (uint64_t)e);
}
__global__ void relu_kernel(float* __restrict__ x, int64_t n) {
  pdl_prologue();
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) x[e] = fmaxf(x[e], 0.f);
}
__global__ void relu_bwd_kernel(const float* __restrict__ r, const float* __restrict__ dr, float* __restrict__ dx, int64_t n) {
  pdl_prologue();
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) dx[e] = r[e] > 0.f ? dr[e] : 0.f;
}
__global__ void mul_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ y, int64_t n) {
  pdl_prologue();
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) y[e] = a[e] * b[e];
}


// out[c] += sum_r x[r, c]  (bias gradients).  block = 32 columns x 8 row-lanes, rows chunked over grid.y
__global__ void __launch_bounds__(256)
colsum_kernel(const void* __restrict__ x, int dt, int64_t ld, int64_t rows, int64_t cols, float* __restrict__ out) {
  pdl_prologue();
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t c = (int64_t)blockIdx.x * 32 + tx;
  const int64_t rchunk = (rows + gridDim.y - 1) / gridDim.y;
  const int64_t r0 = (int64_t)blockIdx.y * rchunk, r1 = min(rows, r0 + rchunk);
  float s = 0.f;
  if (c < cols) for (int64_t r = r0 + ty; r < r1; r += 8) s += ld_as_float(x, dt, r * ld + c);
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][tx];
    atomicAdd(&out[c], t);
  }
}

static inline unsigned ew_blocks(int64_t n, int threads = 256) {
  int64_t b = (n + threads - 1) / threads;
  if (b > kNumSM * 8) b = kNumSM * 8;
  if (b < 1) b = 1;
  return (unsigned)b;
}

}  // namespace dlsg

using namespace dlsg;

extern "C" {

int dlsg_convert2d_batched(const void* src, int sdt, int64_t lds, void* dst, int ddt, int64_t ldd, void* dstT, int64_t ldt,
                           int64_t rows, int64_t cols, int64_t batch, int64_t bs_src, int64_t bs_dst, int64_t bs_dstT, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (rows <= 0 || cols <= 0 || batch <= 0) return 0;
  if (batch == 1 && !dstT && sdt == DLSG_F32 && ddt == DLSG_BF16 && lds == cols && ldd == cols && ((rows * cols) % 8 == 0) &&
      (reinterpret_cast<uintptr_t>(src) % 16 == 0) && (reinterpret_cast<uintptr_t>(dst) % 16 == 0)) {
    const int64_t n8 = rows * cols / 8;
    DLSG_LAUNCH(cast_f32_bf16_flat, ew_blocks(n8), 256, 0, st, (const float4*)src, (uint4*)dst, n8);
    return check_launch("cast_f32_bf16_flat");
  }
  // rows ride grid.x (2^31 limit), columns grid.y
  dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32), (unsigned)batch);
  if (grid.y > 65535 || grid.z > 65535) {
    // split the row range / batches into several launches
    DLSG_REQUIRE(grid.z <= 65535, "convert2d: batch too large");
    const int64_t step = 65535LL * 32;
    for (int64_t r0 = 0; r0 < rows; r0 += step) {
      const int64_t nr = rows - r0 < step ? rows - r0 : step;
      const int es = sdt == DLSG_F32 ? 4 : 2, ed = ddt == DLSG_F32 ? 4 : 2;
      dim3 g2((unsigned)((cols + 31) / 32), (unsigned)((nr + 31) / 32), (unsigned)batch);
      DLSG_LAUNCH(convert2d_kernel, g2, 256, 0, st, reinterpret_cast<const uint8_t*>(src) + r0 * lds * es, sdt, lds,
                                           dst ? reinterpret_cast<uint8_t*>(dst) + r0 * ldd * ed : nullptr, ddt, ldd,
                                           dstT ? reinterpret_cast<uint8_t*>(dstT) + r0 * ed : nullptr, ldt, nr, cols, bs_src, bs_dst, bs_dstT);
    }
    return check_launch("convert2d_kernel");
  }
  DLSG_LAUNCH(convert2d_kernel, grid, 256, 0, st, src, sdt, lds, dst, ddt, ldd, dstT, ldt, rows, cols, bs_src, bs_dst, bs_dstT);
  return check_launch("convert2d_kernel");
}
int dlsg_convert2d(const void* src, int sdt, int64_t lds, void* dst, int ddt, int64_t ldd, void* dstT, int64_t ldt,
                   int64_t rows, int64_t cols, void* stream) {
  return dlsg_convert2d_batched(src, sdt, lds, dst, ddt, ldd, dstT, ldt, rows, cols, 1, 0, 0, 0, stream);
}

int dlsg_norm_fwd(const dlsg_norm_fwd_t* p, void* stream) { return norm_fwd_launch(p, (cudaStream_t)stream); }
int dlsg_norm_bwd(const dlsg_norm_bwd_t* p, void* stream) { return norm_bwd_launch(p, (cudaStream_t)stream); }
int dlsg_norm_bwd_streaming(const dlsg_norm_bwd_t* p) { return norm_bwd_bf16_ok(p) ? 1 : 0; }
int dlsg_norm_bwd2(const dlsg_norm_bwd2_t* p, void* stream) {
  DLSG_REQUIRE(p->rows >= 0 && p->D > 0 && p->D <= 1024, "norm_bwd2: D=%d unsupported (<= 1024)", p->D);
  DLSG_REQUIRE(p->x && p->dy && p->u && p->gamma && p->stats, "norm_bwd2: null input");
  if (p->rows == 0) return 0;
  const int nw = 8;
  int64_t blocks = (p->rows + nw - 1) / nw;
  if (blocks > kNumSM * 4) blocks = kNumSM * 4;
  const size_t smem = (size_t)nw * p->D * sizeof(float);
  if (p->D <= 512) DLSG_LAUNCH(norm_bwd2_kernel<16>, (unsigned)blocks, nw * 32, smem, (cudaStream_t)stream, *p);
  else DLSG_LAUNCH(norm_bwd2_kernel<32>, (unsigned)blocks, nw * 32, smem, (cudaStream_t)stream, *p);
  return check_launch("norm_bwd2_kernel");
}

int dlsg_lstm_cell_fwd(const dlsg_lstm_cell_fwd_t* p, void* stream) {
  DLSG_REQUIRE(p->B > 0 && p->H > 0 && p->nsplit >= 1, "lstm_cell_fwd: bad shape");
  auto a16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  auto okdt = [&](const void* q, int dt, int64_t ld) { return !q || ((ld % 4 == 0) && (reinterpret_cast<uintptr_t>(q) % (dt == DLSG_F32 ? 16 : 8)) == 0); };
  if (p->H % 4 == 0 && p->nsplit <= 4 && a16(p->gates) && p->stride_split % 4 == 0 && (!p->row_bias || (a16(p->row_bias) && p->ld_row_bias % 4 == 0)) &&
      (!p->bias || a16(p->bias)) && (!p->c_prev || a16(p->c_prev)) && a16(p->c_out) && (!p->h_out || a16(p->h_out)) &&
      okdt(p->h2, p->h2_dtype, p->ldh2) && okdt(p->h3, p->h3_dtype, p->ldh3)) {
    const unsigned nb = (unsigned)(((int64_t)p->B * (p->H / 4) + 255) / 256);
    cudaStream_t st = (cudaStream_t)stream;
    switch (p->nsplit) {
      case 1: DLSG_LAUNCH(lstm_cell_fwd_fast<1>, nb, 256, 0, st, *p); break;
      case 2: DLSG_LAUNCH(lstm_cell_fwd_fast<2>, nb, 256, 0, st, *p); break;
      case 3: DLSG_LAUNCH(lstm_cell_fwd_fast<3>, nb, 256, 0, st, *p); break;
      default: DLSG_LAUNCH(lstm_cell_fwd_fast<4>, nb, 256, 0, st, *p); break;
    }
    return check_launch("lstm_cell_fwd_fast");
  }
  DLSG_LAUNCH(lstm_cell_fwd_kernel, ew_blocks((int64_t)p->B * p->H), 256, 0, (cudaStream_t)stream, *p);
  return check_launch("lstm_cell_fwd_kernel");
}
int dlsg_lstm_cell_bwd(const dlsg_lstm_cell_bwd_t* p, void* stream) {
  DLSG_REQUIRE(p->B > 0 && p->H > 0, "lstm_cell_bwd: bad shape");
  auto a16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  const int ns = p->dh2 ? (p->dh2_nsplit > 1 ? p->dh2_nsplit : 1) : 1;
  if (p->H % 4 == 0 && ns <= 16 && a16(p->acts) && a16(p->c_new) && p->dh && a16(p->dh) && p->lddh % 4 == 0 &&
      (!p->dh2 || (a16(p->dh2) && p->lddh2 % 4 == 0 && p->dh2_stride_split % 4 == 0)) && (!p->c_prev || a16(p->c_prev)) &&
      (!p->dc_next || a16(p->dc_next)) && (!p->dc_prev || a16(p->dc_prev)) && (!p->dgates || a16(p->dgates)) &&
      (!p->dc_next2 || a16(p->dc_next2)) && (!p->dgates_add || a16(p->dgates_add)) && (!p->dh_total || a16(p->dh_total)) &&
      p->ld_dgates % 4 == 0 &&
      (!p->dgates2 || (p->ld_dgates2 % 4 == 0 && (reinterpret_cast<uintptr_t>(p->dgates2) % (p->dgates2_dtype == DLSG_F32 ? 16 : 8)) == 0))) {
    const unsigned nb = (unsigned)(((int64_t)p->B * (p->H / 4) + 255) / 256);
    cudaStream_t st = (cudaStream_t)stream;
    switch (ns) {
      case 1: DLSG_LAUNCH(lstm_cell_bwd_fast<1>, nb, 256, 0, st, *p); break;
      case 2: DLSG_LAUNCH(lstm_cell_bwd_fast<2>, nb, 256, 0, st, *p); break;
      case 3: DLSG_LAUNCH(lstm_cell_bwd_fast<3>, nb, 256, 0, st, *p); break;
      case 4: DLSG_LAUNCH(lstm_cell_bwd_fast<4>, nb, 256, 0, st, *p); break;
      default: DLSG_LAUNCH(lstm_cell_bwd_fast<16>, nb, 256, 0, st, *p); break;      // 5..16 partials (the BiLSTM's two half-GPU chains: 9)
    }
    return check_launch("lstm_cell_bwd_fast");
  }
  DLSG_LAUNCH(lstm_cell_bwd_kernel, ew_blocks((int64_t)p->B * p->H), 256, 0, (cudaStream_t)stream, *p);
  return check_launch("lstm_cell_bwd_kernel");
}

int dlsg_multi_convert(const dlsg_seg_t* segs_dev, const int32_t* chunks_dev, int32_t nchunks, void* stream) {
  if (nchunks <= 0) return 0;
  DLSG_REQUIRE(segs_dev && chunks_dev, "multi_convert: null tables");
  DLSG_LAUNCH(multi_convert_kernel, (unsigned)nchunks, 256, 0, (cudaStream_t)stream, segs_dev, chunks_dev);
  return check_launch("multi_convert_kernel");
}

int dlsg_multi_convert_host(const dlsg_seg_t* segs_host, int32_t nsegs, int32_t chunk_elems, void* stream) {
  if (nsegs <= 0) return 0;
  DLSG_REQUIRE(segs_host && chunk_elems > 0, "multi_convert_host: bad arguments");
  static thread_local ConvArgs args;
  for (int s0 = 0; s0 < nsegs; s0 += CONV_MAX_SEGS) {
    const int n = nsegs - s0 < CONV_MAX_SEGS ? nsegs - s0 : CONV_MAX_SEGS;
    int64_t chunks = 0;
    for (int i = 0; i < n; ++i) {
      const dlsg_seg_t& sg = segs_host[s0 + i];
      DLSG_REQUIRE(sg.rows > 0 && sg.cols > 0 && sg.src && sg.dst, "multi_convert_host: empty segment %d", s0 + i);
      DLSG_REQUIRE(sg.src_dtype == DLSG_F32 || (sg.src_dtype == DLSG_BF16 && !sg.src2), "multi_convert_host: source dtype of segment %d", s0 + i);
      int64_t per = chunk_elems / sg.cols;
      if (per < 1) per = 1;
      args.seg[i] = sg;
      args.chunk_start[i] = (int32_t)chunks;
      args.rows_per_chunk[i] = (int32_t)per;
      chunks += (sg.rows + per - 1) / per;
    }
    args.chunk_start[n] = (int32_t)chunks;
    args.nsegs = n;
    DLSG_REQUIRE(chunks < (1ll << 31), "multi_convert_host: too many chunks");
    DLSG_LAUNCH(multi_convert_args_kernel, (unsigned)chunks, 256, 0, (cudaStream_t)stream, args);
    if (int rc = check_launch("multi_convert_args_kernel")) return rc;
  }
  return 0;
}

int dlsg_adam_multi(const dlsg_adam_seg_t* segs_host, int32_t nsegs, int32_t chunk_elems, const float* step_dev,
                    const float* lr_dev, float lr, float beta1, float beta2, float eps, void* stream) {
  if (nsegs <= 0) return 0;
  DLSG_REQUIRE(segs_host && step_dev && chunk_elems > 0, "adam_multi: bad arguments");
  static thread_local AdamArgs args;                                   // 24.6 KB: filled per launch, copied into the launch by value
  for (int s0 = 0; s0 < nsegs; s0 += ADAM_MAX_SEGS) {
    const int n = nsegs - s0 < ADAM_MAX_SEGS ? nsegs - s0 : ADAM_MAX_SEGS;
    int64_t chunks = 0;
    for (int i = 0; i < n; ++i) {
      const dlsg_adam_seg_t& sg = segs_host[s0 + i];
      DLSG_REQUIRE(sg.rows > 0 && sg.cols > 0 && sg.p && sg.g && sg.m && sg.v, "adam_multi: empty segment %d", s0 + i);
      DLSG_REQUIRE(sg.g_dtype == DLSG_F32 || sg.g_dtype == DLSG_BF16, "adam_multi: gradient dtype of segment %d", s0 + i);
      int64_t per = chunk_elems / sg.cols;
      if (per < 1) per = 1;
      args.seg[i] = sg;
      args.chunk_start[i] = (int32_t)chunks;
      args.rows_per_chunk[i] = (int32_t)per;
      chunks += (sg.rows + per - 1) / per;
    }
    args.chunk_start[n] = (int32_t)chunks;
    args.nsegs = n;
    DLSG_REQUIRE(chunks < (1ll << 31), "adam_multi: too many chunks");
    DLSG_LAUNCH(adam_multi_kernel, (unsigned)chunks, 256, 0, (cudaStream_t)stream, args, step_dev, lr_dev, lr, beta1, beta2, eps);
    if (int rc = check_launch("adam_multi_kernel")) return rc;
  }
  return 0;
}

int dlsg_lstm_cell_bwd2(const dlsg_lstm_cell_bwd2_t* p, void* stream) {
  DLSG_REQUIRE(p->B > 0 && p->H > 0 && p->acts && p->c_new && p->dh, "lstm_cell_bwd2: bad arguments");
  DLSG_LAUNCH(lstm_cell_bwd2_kernel, ew_blocks((int64_t)p->B * p->H), 256, 0, (cudaStream_t)stream, *p);
  return check_launch("lstm_cell_bwd2_kernel");
}

int dlsg_softmax_fwd(const dlsg_softmax_t* p, void* stream) {
  const int64_t pairs = p->outer * p->inner;
  if (pairs <= 0 || p->n <= 0) return 0;
  DLSG_REQUIRE(p->mask_mode == 0 || p->mask != nullptr, "softmax: mask_mode set without mask");
  DLSG_LAUNCH(softmax_fwd_kernel, ew_blocks(pairs * 32), 256, 0, (cudaStream_t)stream, *p);
  return check_launch("softmax_fwd_kernel");
}
int dlsg_softmax_bwd(const dlsg_softmax_t* p, const float* dy, float* dx, void* stream) {
  const int64_t pairs = p->outer * p->inner;
  if (pairs <= 0 || p->n <= 0) return 0;
  DLSG_REQUIRE(p->mask_mode == 0 || p->mask != nullptr, "softmax: mask_mode set without mask");
  DLSG_LAUNCH(softmax_bwd_kernel, ew_blocks(pairs * 32), 256, 0, (cudaStream_t)stream, *p, dy, dx);
  return check_launch("softmax_bwd_kernel");
}

int dlsg_softmax_bwd2(const dlsg_softmax_t* p, const float* dy, const float* u, float* g_dy, float* g_x, void* stream) {
  const int64_t pairs = p->outer * p->inner;
  if (pairs <= 0 || p->n <= 0) return 0;
  DLSG_REQUIRE(p->mask_mode == 0 || p->mask != nullptr, "softmax_bwd2: mask_mode set without mask");
  DLSG_REQUIRE(dy && u, "softmax_bwd2: dy and u are required");
  DLSG_LAUNCH(softmax_bwd2_kernel, ew_blocks(pairs * 32), 256, 0, (cudaStream_t)stream, *p, dy, u, g_dy, g_x);
  return check_launch("softmax_bwd2_kernel");
}
int dlsg_ew(const dlsg_ew_t* p, void* stream) {
  if (p->n <= 0) return 0;
  DLSG_REQUIRE(p->op >= DLSG_EW_TANH_BWD && p->op <= DLSG_EW_LERP_ROWS_BWD, "ew: unknown op");
  const int nin = p->op == DLSG_EW_MUL_BWD2 ? 5 : (p->op == DLSG_EW_TANH_BWD || p->op == DLSG_EW_LERP_ROWS_BWD ? 2 : 3);
  const bool rows_e = p->op == DLSG_EW_LERP_ROWS || p->op == DLSG_EW_LERP_ROWS_BWD;
  const int e_slot = p->op == DLSG_EW_LERP_ROWS ? 2 : 1;
  DLSG_REQUIRE(!rows_e || p->cols > 0, "ew: cols required for a per-row operand");
  bool vec = p->n % 4 == 0 && (!rows_e || p->cols % 4 == 0);
  for (int k = 0; k < nin; ++k) {
    DLSG_REQUIRE(p->in[k] != nullptr, "ew: missing input");
    if (!(rows_e && k == e_slot)) vec = vec && (reinterpret_cast<uintptr_t>(p->in[k]) & 15) == 0;
  }
  for (int k = 0; k < 3; ++k) vec = vec && (reinterpret_cast<uintptr_t>(p->out[k]) & 15) == 0;
  if (vec) {
    DLSG_LAUNCH(ew_kernel<4>, ew_blocks(p->n / 4), 256, 0, (cudaStream_t)stream, *p);
  } else {
    DLSG_LAUNCH(ew_kernel<1>, ew_blocks(p->n), 256, 0, (cudaStream_t)stream, *p);
  }
  return check_launch("ew_kernel");
}

int dlsg_embedding_gather(const float* table, const int64_t* ids, int64_t ld_ids, int32_t rows, int32_t W, void* out,
                          int odt, int64_t ldo, void* out2, int odt2, int64_t ldo2, float drop_p, uint64_t seed,
                          uint64_t offset, void* stream) {
  if (rows <= 0) return 0;
  DLSG_LAUNCH(embedding_gather_kernel, rows, 128, 0, (cudaStream_t)stream, table, ids, ld_ids, rows, W, out, odt, ldo, out2, odt2,
                                                                 ldo2, drop_p, seed, offset);
  return check_launch("embedding_gather_kernel");
}
int dlsg_embedding_scatter_add(float* dtable, const int64_t* ids, int64_t ld_ids, int32_t rows, int32_t W,
                               const float* dout, int64_t lddo, float drop_p, uint64_t seed, uint64_t offset, void* stream) {
  if (rows <= 0) return 0;
  DLSG_LAUNCH(embedding_scatter_kernel, rows, 128, 0, (cudaStream_t)stream, dtable, ids, ld_ids, rows, W, dout, lddo, drop_p, seed, offset);
  return check_launch("embedding_scatter_kernel");
}
int dlsg_mean_nodes_fwd(const float* x, int32_t B, int32_t P, int32_t H, float* y, int64_t ldy, void* stream) {
  DLSG_LAUNCH(mean_nodes_fwd_kernel, ew_blocks((int64_t)B * H), 256, 0, (cudaStream_t)stream, x, B, P, H, y, ldy);
  return check_launch("mean_nodes_fwd_kernel");
}
int dlsg_mean_nodes_bwd(const float* dy, int64_t lddy, int32_t B, int32_t P, int32_t H, float* dx, void* stream) {
  DLSG_LAUNCH(mean_nodes_bwd_kernel, ew_blocks((int64_t)B * P * H), 256, 0, (cudaStream_t)stream, dy, lddy, B, P, H, dx);
  return check_launch("mean_nodes_bwd_kernel");
}
int dlsg_axpby(const float* x, float a, float* y, float b, int64_t n, void* stream) {
  if (n <= 0) return 0;
  DLSG_LAUNCH(axpby_kernel, ew_blocks(n), 256, 0, (cudaStream_t)stream, x, a, y, b, n);
  return check_launch("axpby_kernel");
}
int dlsg_add_rowbcast(const float* x, const float* pe, float* y, int64_t batch, int64_t inner, float drop_p, uint64_t seed,
                      uint64_t offset, void* stream) {
  if (batch * inner <= 0) return 0;
  DLSG_LAUNCH(add_rowbcast_kernel, ew_blocks(batch * inner), 256, 0, (cudaStream_t)stream, x, pe, y, batch, inner, drop_p, seed, offset);
  return check_launch("add_rowbcast_kernel");
}
int dlsg_dropout(const float* x, float* y, int64_t n, float drop_p, uint64_t seed, uint64_t offset, void* stream) {
  if (n <= 0) return 0;
  DLSG_LAUNCH(dropout_kernel, ew_blocks(n), 256, 0, (cudaStream_t)stream, x, y, n, drop_p, seed, offset);
  return check_launch("dropout_kernel");
}
int dlsg_relu(float* x, int64_t n, void* stream) {
  if (n <= 0) return 0;
  DLSG_LAUNCH(relu_kernel, ew_blocks(n), 256, 0, (cudaStream_t)stream, x, n);
  return check_launch("relu_kernel");
}
int dlsg_relu_bwd(const float* r, const float* dr, float* dx, int64_t n, void* stream) {
  if (n <= 0) return 0;
  DLSG_LAUNCH(relu_bwd_kernel, ew_blocks(n), 256, 0, (cudaStream_t)stream, r, dr, dx, n);
  return check_launch("relu_bwd_kernel");
}

int dlsg_colsum(const void* x, int dtype, int64_t ld, int64_t rows, int64_t cols, float* out, void* stream) {
  if (rows <= 0 || cols <= 0) return 0;
  int64_t chunks = (rows + 255) / 256;
  if (chunks > 64) chunks = 64;
  DLSG_LAUNCH(colsum_kernel, dim3((unsigned)((cols + 31) / 32), (unsigned)chunks), 256, 0, (cudaStream_t)stream, x, dtype, ld, rows, cols, out);
  return check_launch("colsum_kernel");
}
int dlsg_mul(const float* a, const float* b, float* y, int64_t n, void* stream) {
  if (n <= 0) return 0;
  DLSG_LAUNCH(mul_kernel, ew_blocks(n), 256, 0, (cudaStream_t)stream, a, b, y, n);
  return check_launch("mul_kernel");
}

}  // extern "C"
