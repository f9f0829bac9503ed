// Per-step fusions of the decoder's LSTM stages (one CTA per batch row, H/4 threads (<= 512), 4 consecutive hidden units per
// thread so every access is a 128-bit load and all loads of a thread are issued back to back):
//   lstm_cell_norm_fwd : split-K partial sums + hoisted row bias -> gates -> c,h (dropout) -> LayerNorm(h) [-> tanh]
//                        [-> dropout]; writes h into the next step's operand rows and LN(h) into its consumer's row
//                        (layer.py:571-574 query LSTM + LN, layer.py:593-599 lang LSTM + LN + tanh)
//   norm_lstm_cell_bwd : LayerNorm backward + LSTM cell backward (+ running sum of the gate gradients); the per-row
//                        dgamma / dbeta contributions are written out (reduced over time by one colsum after BPTT)
// They replace 3 launches each (reduce / cell / norm, resp. norm_bwd / cell_bwd / axpby) in the 26-step loop.
#include "common.cuh"

namespace dlsg {

constexpr int FS_NG = 1;          // float4 groups per thread (threads = H/4 <= 512 cover H <= 2048)
constexpr int FS_MAXH = 2048;
constexpr int FS_MAXS = 8;        // split-K partials summed in registers

__device__ __forceinline__ float4 ld4f(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4any(void* base, int dt, int64_t off, float4 v, bool vec) {
  if (dt == DLSG_F32) {
    if (vec) { *reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + off) = v; return; }
    float* d = reinterpret_cast<float*>(base) + off;
    d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
    return;
  }
  __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(base) + off;
  if (vec) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 u; u.x = *reinterpret_cast<uint32_t*>(&a); u.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(d) = u;
    return;
  }
  d[0] = __float2bfloat16_rn(v.x); d[1] = __float2bfloat16_rn(v.y); d[2] = __float2bfloat16_rn(v.z); d[3] = __float2bfloat16_rn(v.w);
}
__device__ __forceinline__ bool vecok(const void* p, int dt, int64_t ld) {
  const int es = dt == DLSG_F32 ? 4 : 2;
  return (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(p) % (4 * es)) == 0);
}
__device__ __forceinline__ float4 f4add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

// Fast path (the shapes of the decoder loop): the number of split-K partials is a template parameter and EVERY global load
// of a thread (4 gates x S partials, the bias, c_prev, the LayerNorm parameters: 20-30 independent 128-bit loads) is issued
// before the first use, into registers of its own.  The generic kernel below sums the partials in a loop over a run-time
// count; ptxas reuses two or three destination registers there, which serialises the loads into a chain of L2 round
// trips: 10 us per launch inside the 26-step loop for 5 MB of data (ncu source view, profiles/r02_step_kernels_ncu.json).
template <int S>
__device__ __forceinline__ float4 cell_norm_fwd_body(const dlsg_lstm_cell_norm_fwd_t& q, float* red) {
  const dlsg_lstm_cell_fwd_t& p = q.cell;
  const int b = blockIdx.x, H = p.H, tid = threadIdx.x;
  const int h = tid * 4;
  const bool act = h < H;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 part[4][S], eb[4], cp = z4, ga = z4, bt = z4;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    eb[k] = z4;
#pragma unroll
    for (int s_ = 0; s_ < S; ++s_) part[k][s_] = z4;
  }
  if (act) {
    const float* g0p = p.gates + (int64_t)b * 4 * H + h;
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int s_ = 0; s_ < S; ++s_) part[k][s_] = ld4f(g0p + (int64_t)k * H + (int64_t)s_ * p.stride_split);
    if (p.row_bias) {
#pragma unroll
      for (int k = 0; k < 4; ++k) eb[k] = ld4f(p.row_bias + (int64_t)b * p.ld_row_bias + (int64_t)k * H + h);
    } else if (p.bias) {
#pragma unroll
      for (int k = 0; k < 4; ++k) eb[k] = ld4f(p.bias + k * H + h);
    }
    if (p.c_prev) cp = ld4f(p.c_prev + (int64_t)b * H + h);
    ga = ld4f(q.gamma + h);
    bt = ld4f(q.beta + h);
  }
  float4 g[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float4 v = part[k][0];
#pragma unroll
    for (int s_ = 1; s_ < S; ++s_) v = f4add(v, part[k][s_]);
    g[k] = f4add(v, eb[k]);
  }
  if (p.row_bias && p.bias && act) {            // both given (not the decoder's case): one more round trip
#pragma unroll
    for (int k = 0; k < 4; ++k) g[k] = f4add(g[k], ld4f(p.bias + k * H + h));
  }
  const bool v2 = p.h2 && vecok(p.h2, p.h2_dtype, p.ldh2), v3 = p.h3 && vecok(p.h3, p.h3_dtype, p.ldh3);
  const bool vy = q.y && vecok(q.y, q.y_dtype, q.ldy), vy2 = q.y2 && vecok(q.y2, q.y2_dtype, q.ldy2);
  float4 h4 = z4;
  float sum = 0.f;
  if (act) {
    const int64_t ei = (int64_t)b * H + h;
    const float gi_[4] = {g[0].x, g[0].y, g[0].z, g[0].w}, gf_[4] = {g[1].x, g[1].y, g[1].z, g[1].w};
    const float gg_[4] = {g[2].x, g[2].y, g[2].z, g[2].w}, go_[4] = {g[3].x, g[3].y, g[3].z, g[3].w};
    const float cp_[4] = {cp.x, cp.y, cp.z, cp.w};
    float ai[4], af[4], ag[4], ao[4], cc[4], hh[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      ai[u] = sigmoidf_(gi_[u]); af[u] = sigmoidf_(gf_[u]); ag[u] = tanhf(gg_[u]); ao[u] = sigmoidf_(go_[u]);
      cc[u] = af[u] * cp_[u] + ai[u] * ag[u];
      hh[u] = ao[u] * tanhf(cc[u]);
    }
    if (p.drop_p > 0.f) {
      const float4 m = drop_mask4(p.drop_p, 1.f / (1.f - p.drop_p), p.seed, p.offset + (uint64_t)ei);
      hh[0] *= m.x; hh[1] *= m.y; hh[2] *= m.z; hh[3] *= m.w;
    }
    float* g0 = p.gates + (int64_t)b * 4 * H + h;
    *reinterpret_cast<float4*>(g0) = make_float4(ai[0], ai[1], ai[2], ai[3]);
    *reinterpret_cast<float4*>(g0 + H) = make_float4(af[0], af[1], af[2], af[3]);
    *reinterpret_cast<float4*>(g0 + 2 * (int64_t)H) = make_float4(ag[0], ag[1], ag[2], ag[3]);
    *reinterpret_cast<float4*>(g0 + 3 * (int64_t)H) = make_float4(ao[0], ao[1], ao[2], ao[3]);
    *reinterpret_cast<float4*>(p.c_out + ei) = make_float4(cc[0], cc[1], cc[2], cc[3]);
    h4 = make_float4(hh[0], hh[1], hh[2], hh[3]);
    if (p.h_out) *reinterpret_cast<float4*>(p.h_out + ei) = h4;
    if (p.h2) st4any(p.h2, p.h2_dtype, (int64_t)b * p.ldh2 + h, h4, v2);
    if (p.h3) st4any(p.h3, p.h3_dtype, (int64_t)b * p.ldh3 + h, h4, v3);
    sum = (h4.x + h4.y) + (h4.z + h4.w);
  }
  const float mean = block_sum(sum, red) / (float)H;
  float sq = 0.f;
  if (act) {
    const float a = h4.x - mean, b2 = h4.y - mean, c = h4.z - mean, d = h4.w - mean;
    sq = (a * a + b2 * b2) + (c * c + d * d);
  }
  const float rstd = rsqrtf(block_sum(sq, red) / (float)H + 1e-5f);
  if (q.stats && tid == 0) { q.stats[b * 2] = mean; q.stats[b * 2 + 1] = rstd; }
  if (act) {
    float y[4] = {(h4.x - mean) * rstd * ga.x + bt.x, (h4.y - mean) * rstd * ga.y + bt.y,
                  (h4.z - mean) * rstd * ga.z + bt.z, (h4.w - mean) * rstd * ga.w + bt.w};
    if (q.post_tanh) {
#pragma unroll
      for (int u = 0; u < 4; ++u) y[u] = tanhf(y[u]);
    }
    if (q.ydrop_p > 0.f) {
      const float4 m = drop_mask4(q.ydrop_p, 1.f / (1.f - q.ydrop_p), q.yseed, q.yoffset + (uint64_t)b * H + h);
      y[0] *= m.x; y[1] *= m.y; y[2] *= m.z; y[3] *= m.w;
    }
    const float4 y4 = make_float4(y[0], y[1], y[2], y[3]);
    if (q.y) st4any(q.y, q.y_dtype, (int64_t)b * q.ldy + h, y4, vy);
    if (q.y2) st4any(q.y2, q.y2_dtype, (int64_t)b * q.ldy2 + h, y4, vy2);
    return y4;
  }
  return z4;
}

template <int S>
__global__ void __launch_bounds__(384)
lstm_cell_norm_fwd_fast(const dlsg_lstm_cell_norm_fwd_t q) {
  pdl_prologue();
  __shared__ float red[32];
  cell_norm_fwd_body<S>(q, red);
}

// ---- query-LSTM cell + LayerNorm AND the hoisted attention step + context output layer in ONE launch (layer.py:571-591):
// one CTA per batch row, 256 threads per attention head.  Threads [0, H/4) run the cell + LayerNorm exactly as
// lstm_cell_norm_fwd_fast does and leave q = dropout(LN(query_h)) in shared memory; then every head's 256 threads run
// attn2_fwd_kernel's step (scores over the P latent nodes, softmax, weighted sum, tanh -> LayerNorm -> dropout) on it.
// Same arithmetic, same Philox sites, same outputs as the two kernels launched back to back - one launch gap and one
// kernel latency less per decode step.
constexpr int FA_PM = 8;          // max latent nodes (= APM of decode_ops.cu)
__device__ __forceinline__ float fa_dot4(const float4 a, const float4 b) { return (a.x * b.x + a.y * b.y) + (a.z * b.z + a.w * b.w); }
// sum over the 256 threads of one head; every thread of the block calls it
__device__ __forceinline__ float fa_head_sum(float v, float (*sh)[8], int hd, int w, int lane) {
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) sh[hd][w] = v;
  __syncthreads();
  float r = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) r += sh[hd][i];
  return r;
}

template <int S>
__global__ void __launch_bounds__(512)
cell_norm_attn2_fwd_kernel(const dlsg_cell_norm_attn2_fwd_t f) {
  pdl_prologue();
  __shared__ float red[32];
  __shared__ __align__(16) float qs[1024];
  __shared__ float ared[2][8][FA_PM];
  __shared__ float al[2][FA_PM];
  __shared__ float hred[2][8];
  const dlsg_attn2_fwd_t& p = f.at;
  const int tid = threadIdx.x;
  {
    const float4 y4 = cell_norm_fwd_body<S>(f.cn, red);
    if (tid * 4 < f.cn.cell.H) *reinterpret_cast<float4*>(qs + tid * 4) = y4;
  }
  __syncthreads();
  const int r = blockIdx.x, hd = tid >> 8, t = tid & 255, lane = tid & 31, w = t >> 5;
  const int node = r / p.rows_per_node;
  const int c = t * 4;
  const bool ak = c < p.Hk, av = c < p.Hv;
  const float* KW = p.KW + (((int64_t)hd * p.nodes + node) * p.P) * p.Hk;
  const float* VW = p.VW + (((int64_t)hd * p.nodes + node) * p.P) * p.Hv;
  float4 k4[FA_PM], v4[FA_PM], q4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (ak) q4 = *reinterpret_cast<const float4*>(qs + c);
#pragma unroll
  for (int j = 0; j < FA_PM; ++j) {
    k4[j] = make_float4(0.f, 0.f, 0.f, 0.f); v4[j] = k4[j];
    if (j < p.P) {
      if (ak) k4[j] = *reinterpret_cast<const float4*>(KW + (int64_t)j * p.Hk + c);
      if (av) v4[j] = *reinterpret_cast<const float4*>(VW + (int64_t)j * p.Hv + c);
    }
  }
#pragma unroll
  for (int j = 0; j < FA_PM; ++j) {
    const float s_ = warp_sum(fa_dot4(k4[j], q4));
    if (lane == 0) ared[hd][w][j] = s_;
  }
  __syncthreads();
  if (t == 0) {
    float lg[FA_PM], mx = -INFINITY;
    for (int j = 0; j < p.P; ++j) {
      float s_ = 0.f;
      for (int ww = 0; ww < 8; ++ww) s_ += ared[hd][ww][j];
      lg[j] = s_ * p.scale; mx = fmaxf(mx, lg[j]);
    }
    float sum = 0.f;
    for (int j = 0; j < p.P; ++j) { lg[j] = expf(lg[j] - mx); sum += lg[j]; }
    const float inv = 1.f / sum;
    for (int j = 0; j < p.P; ++j) {
      const float a = lg[j] * inv;
      al[hd][j] = a;
      if (p.alpha) p.alpha[(int64_t)r * p.ldalpha + hd * p.P + j] = a;
    }
  }
  __syncthreads();
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (av) {
#pragma unroll
    for (int j = 0; j < FA_PM; ++j) {
      if (j < p.P) { const float a = al[hd][j]; acc.x = fmaf(a, v4[j].x, acc.x); acc.y = fmaf(a, v4[j].y, acc.y); acc.z = fmaf(a, v4[j].z, acc.z); acc.w = fmaf(a, v4[j].w, acc.w); }
    }
    *reinterpret_cast<float4*>(p.co + (int64_t)r * p.ldco + hd * p.Hv + c) = acc;
  }
  if (p.y == nullptr) return;                       // uniform
  float4 t4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (av) t4 = make_float4(tanhf(acc.x), tanhf(acc.y), tanhf(acc.z), tanhf(acc.w));
  const float* gam = hd ? p.gamma[1] : p.gamma[0];
  const float* bet = hd ? p.beta[1] : p.beta[0];
  float4 g4 = t4, b4 = t4;
  if (av) { g4 = *reinterpret_cast<const float4*>(gam + c); b4 = *reinterpret_cast<const float4*>(bet + c); }
  const float invH = 1.f / (float)p.Hv;
  const float mean = fa_head_sum(av ? (t4.x + t4.y) + (t4.z + t4.w) : 0.f, hred, hd, w, lane) * invH;
  float sq = 0.f;
  if (av) { const float a = t4.x - mean, b = t4.y - mean, cc = t4.z - mean, d = t4.w - mean; sq = (a * a + b * b) + (cc * cc + d * d); }
  const float rstd = rsqrtf(fa_head_sum(sq, hred, hd, w, lane) * invH + 1e-5f);
  if (p.stats && t == 0) {
    float* st = p.stats + (int64_t)hd * p.stats_head_stride + 2 * (int64_t)r;
    st[0] = mean; st[1] = rstd;
  }
  if (av) {
    float4 y;
    y.x = (t4.x - mean) * rstd * g4.x + b4.x; y.y = (t4.y - mean) * rstd * g4.y + b4.y;
    y.z = (t4.z - mean) * rstd * g4.z + b4.z; y.w = (t4.w - mean) * rstd * g4.w + b4.w;
    if (p.drop_p > 0.f) {
      const float4 m = drop_mask4(p.drop_p, 1.f / (1.f - p.drop_p), p.seed,
                                  p.offset + (uint64_t)hd * p.offset_head_stride + (uint64_t)r * p.Hv + c);
      y.x *= m.x; y.y *= m.y; y.z *= m.z; y.w *= m.w;
    }
    st4any(p.y, p.y_dtype, (int64_t)r * p.ldy + hd * p.Hv + c, y, true);
  }
}

__global__ void __launch_bounds__(512)
lstm_cell_norm_fwd_kernel(const dlsg_lstm_cell_norm_fwd_t q) {
  pdl_prologue();
  __shared__ float red[32];
  const dlsg_lstm_cell_fwd_t& p = q.cell;
  const int b = blockIdx.x, H = p.H, tid = threadIdx.x;
  const bool v2 = p.h2 && vecok(p.h2, p.h2_dtype, p.ldh2), v3 = p.h3 && vecok(p.h3, p.h3_dtype, p.ldh3);
  const bool vy = q.y && vecok(q.y, q.y_dtype, q.ldy), vy2 = q.y2 && vecok(q.y2, q.y2_dtype, q.ldy2);
  float4 hv[FS_NG];
  float sum = 0.f;
#pragma unroll
  for (int e = 0; e < FS_NG; ++e) {
    const int h = (tid + e * (int)blockDim.x) * 4;
    hv[e] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (h < H) {
      float4 g[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int64_t gi = (int64_t)b * 4 * H + (int64_t)k * H + h;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int s = 0; s < FS_MAXS; ++s)
          if (s < p.nsplit) v = f4add(v, ld4f(p.gates + gi + (int64_t)s * p.stride_split));
        for (int s = FS_MAXS; s < p.nsplit; ++s) v = f4add(v, ld4f(p.gates + gi + (int64_t)s * p.stride_split));
        if (p.row_bias) v = f4add(v, ld4f(p.row_bias + (int64_t)b * p.ld_row_bias + (int64_t)k * H + h));
        if (p.bias) v = f4add(v, ld4f(p.bias + k * H + h));
        g[k] = v;
      }
      const int64_t ei = (int64_t)b * H + h;
      float4 cp = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p.c_prev) cp = ld4f(p.c_prev + ei);
      const float gi_[4] = {g[0].x, g[0].y, g[0].z, g[0].w}, gf_[4] = {g[1].x, g[1].y, g[1].z, g[1].w};
      const float gg_[4] = {g[2].x, g[2].y, g[2].z, g[2].w}, go_[4] = {g[3].x, g[3].y, g[3].z, g[3].w};
      const float cp_[4] = {cp.x, cp.y, cp.z, cp.w};
      float ai[4], af[4], ag[4], ao[4], cc[4], hh[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        ai[u] = sigmoidf_(gi_[u]); af[u] = sigmoidf_(gf_[u]); ag[u] = tanhf(gg_[u]); ao[u] = sigmoidf_(go_[u]);
        cc[u] = af[u] * cp_[u] + ai[u] * ag[u];
        hh[u] = ao[u] * tanhf(cc[u]);
      }
      if (p.drop_p > 0.f) {       // offsets are multiples of 4 (checked on the host): one Philox evaluation per float4
        const float4 m = drop_mask4(p.drop_p, 1.f / (1.f - p.drop_p), p.seed, p.offset + (uint64_t)ei);
        hh[0] *= m.x; hh[1] *= m.y; hh[2] *= m.z; hh[3] *= m.w;
      }
      const int64_t g0 = (int64_t)b * 4 * H + h;
      *reinterpret_cast<float4*>(p.gates + g0) = make_float4(ai[0], ai[1], ai[2], ai[3]);
      *reinterpret_cast<float4*>(p.gates + g0 + H) = make_float4(af[0], af[1], af[2], af[3]);
      *reinterpret_cast<float4*>(p.gates + g0 + 2 * (int64_t)H) = make_float4(ag[0], ag[1], ag[2], ag[3]);
      *reinterpret_cast<float4*>(p.gates + g0 + 3 * (int64_t)H) = make_float4(ao[0], ao[1], ao[2], ao[3]);
      *reinterpret_cast<float4*>(p.c_out + ei) = make_float4(cc[0], cc[1], cc[2], cc[3]);
      const float4 h4 = make_float4(hh[0], hh[1], hh[2], hh[3]);
      if (p.h_out) *reinterpret_cast<float4*>(p.h_out + ei) = h4;
      if (p.h2) st4any(p.h2, p.h2_dtype, (int64_t)b * p.ldh2 + h, h4, v2);
      if (p.h3) st4any(p.h3, p.h3_dtype, (int64_t)b * p.ldh3 + h, h4, v3);
      hv[e] = h4;
      sum += (h4.x + h4.y) + (h4.z + h4.w);
    }
  }
  float4 ga[FS_NG], bt[FS_NG];                      // LayerNorm parameters: loads in flight during the two reductions
#pragma unroll
  for (int e = 0; e < FS_NG; ++e) {
    const int h = (tid + e * (int)blockDim.x) * 4;
    ga[e] = make_float4(0.f, 0.f, 0.f, 0.f); bt[e] = ga[e];
    if (h < H) { ga[e] = ld4f(q.gamma + h); bt[e] = ld4f(q.beta + h); }
  }
  const float mean = block_sum(sum, red) / (float)H;
  float sq = 0.f;
#pragma unroll
  for (int e = 0; e < FS_NG; ++e) {
    const int h = (tid + e * (int)blockDim.x) * 4;
    if (h < H) {
      const float a = hv[e].x - mean, b2 = hv[e].y - mean, c = hv[e].z - mean, d = hv[e].w - mean;
      sq += (a * a + b2 * b2) + (c * c + d * d);
    }
  }
  const float rstd = rsqrtf(block_sum(sq, red) / (float)H + 1e-5f);
  if (q.stats && tid == 0) { q.stats[b * 2] = mean; q.stats[b * 2 + 1] = rstd; }
#pragma unroll
  for (int e = 0; e < FS_NG; ++e) {
    const int h = (tid + e * (int)blockDim.x) * 4;
    if (h < H) {
      float y[4] = {(hv[e].x - mean) * rstd * ga[e].x + bt[e].x, (hv[e].y - mean) * rstd * ga[e].y + bt[e].y,
                    (hv[e].z - mean) * rstd * ga[e].z + bt[e].z, (hv[e].w - mean) * rstd * ga[e].w + bt[e].w};
      if (q.post_tanh) {
#pragma unroll
        for (int u = 0; u < 4; ++u) y[u] = tanhf(y[u]);
      }
      if (q.ydrop_p > 0.f) {
        const float4 m = drop_mask4(q.ydrop_p, 1.f / (1.f - q.ydrop_p), q.yseed, q.yoffset + (uint64_t)b * H + h);
        y[0] *= m.x; y[1] *= m.y; y[2] *= m.z; y[3] *= m.w;
      }
      const float4 y4 = make_float4(y[0], y[1], y[2], y[3]);
      if (q.y) st4any(q.y, q.y_dtype, (int64_t)b * q.ldy + h, y4, vy);
      if (q.y2) st4any(q.y2, q.y2_dtype, (int64_t)b * q.ldy2 + h, y4, vy2);
    }
  }
}

// Backward counterpart of lstm_cell_norm_fwd_fast: every load of the thread (LayerNorm operands, the four saved gate
// activations, cell states, the recurrent gradients with their S2 split-K partials, the running gate-gradient sum) is issued
// before the first reduction, so the two block-wide sums overlap the memory round trip instead of following it.
template <int S2>
__global__ void __launch_bounds__(384)
norm_lstm_cell_bwd_fast(const dlsg_norm_lstm_cell_bwd_t q) {
  pdl_prologue();
  __shared__ float red[32];
  const dlsg_lstm_cell_bwd_t& p = q.cell;
  const int b = blockIdx.x, H = p.H, tid = threadIdx.x;
  const int h = tid * 4;
  const bool act = h < H;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 x = z4, dy = z4, g = z4, be = z4, ai = z4, af = z4, ag = z4, ao = z4, cn = z4, cp = z4, dcn = z4, r1 = z4, r2[S2], gs[4];
#pragma unroll
  for (int s_ = 0; s_ < S2; ++s_) r2[s_] = z4;
#pragma unroll
  for (int k = 0; k < 4; ++k) gs[k] = z4;
  float mean = 0.f, rstd = 0.f;
  const int64_t ei = (int64_t)b * H + h, g0 = (int64_t)b * 4 * H + h;
  if (act) {
    x = ld4f(q.x + (int64_t)b * q.ldx + h);
    dy = ld4f(q.dy + (int64_t)b * q.lddy + h);
    g = ld4f(q.gamma + h);
    if (q.post_tanh) be = ld4f(q.beta + h);
    ai = ld4f(p.acts + g0); af = ld4f(p.acts + g0 + H); ag = ld4f(p.acts + g0 + 2 * (int64_t)H); ao = ld4f(p.acts + g0 + 3 * (int64_t)H);
    cn = ld4f(p.c_new + ei);
    if (p.c_prev) cp = ld4f(p.c_prev + ei);
    if (p.dc_next) dcn = ld4f(p.dc_next + ei);
    if (p.dh) r1 = ld4f(p.dh + (int64_t)b * p.lddh + h);
    if (p.dh2) {
#pragma unroll
      for (int s_ = 0; s_ < S2; ++s_) r2[s_] = ld4f(p.dh2 + (int64_t)b * p.lddh2 + h + (int64_t)s_ * p.dh2_stride_split);
    }
    if (q.dgates_sum) {
#pragma unroll
      for (int k = 0; k < 4; ++k) gs[k] = ld4f(q.dgates_sum + g0 + (int64_t)k * H);
    }
  }
  mean = q.stats[b * 2]; rstd = q.stats[b * 2 + 1];
  const bool vd2 = p.dgates2 && vecok(p.dgates2, p.dgates2_dtype, p.ld_dgates2);
  float4 xn = z4, d = z4;
  float s1 = 0.f, s2 = 0.f;
  if (act) {
    if (q.ydrop_p > 0.f) {
      const float4 m = drop_mask4(q.ydrop_p, 1.f / (1.f - q.ydrop_p), q.yseed, q.yoffset + (uint64_t)b * H + h);
      dy.x *= m.x; dy.y *= m.y; dy.z *= m.z; dy.w *= m.w;
    }
    xn = make_float4((x.x - mean) * rstd, (x.y - mean) * rstd, (x.z - mean) * rstd, (x.w - mean) * rstd);
    if (q.post_tanh) {
      float yt;
      yt = tanhf(xn.x * g.x + be.x); dy.x *= (1.f - yt * yt);
      yt = tanhf(xn.y * g.y + be.y); dy.y *= (1.f - yt * yt);
      yt = tanhf(xn.z * g.z + be.z); dy.z *= (1.f - yt * yt);
      yt = tanhf(xn.w * g.w + be.w); dy.w *= (1.f - yt * yt);
    }
    *reinterpret_cast<float4*>(q.dgamma + (int64_t)b * q.ld_dparam + h) = make_float4(dy.x * xn.x, dy.y * xn.y, dy.z * xn.z, dy.w * xn.w);
    *reinterpret_cast<float4*>(q.dbeta + (int64_t)b * q.ld_dparam + h) = dy;
    d = make_float4(dy.x * g.x, dy.y * g.y, dy.z * g.z, dy.w * g.w);
    s1 = (d.x + d.y) + (d.z + d.w);
    s2 = (d.x * xn.x + d.y * xn.y) + (d.z * xn.z + d.w * xn.w);
  }
  s1 = block_sum(s1, red) / (float)H;
  s2 = block_sum(s2, red) / (float)H;
  if (!act) return;
  float4 r2s = r2[0];
#pragma unroll
  for (int s_ = 1; s_ < S2; ++s_) r2s = f4add(r2s, r2[s_]);
  const float dxl[4] = {rstd * (d.x - s1 - xn.x * s2), rstd * (d.y - s1 - xn.y * s2), rstd * (d.z - s1 - xn.z * s2), rstd * (d.w - s1 - xn.w * s2)};
  const float r1_[4] = {r1.x, r1.y, r1.z, r1.w}, r2_[4] = {r2s.x, r2s.y, r2s.z, r2s.w};
  const float i_[4] = {ai.x, ai.y, ai.z, ai.w}, f_[4] = {af.x, af.y, af.z, af.w}, g_[4] = {ag.x, ag.y, ag.z, ag.w}, o_[4] = {ao.x, ao.y, ao.z, ao.w};
  const float cn_[4] = {cn.x, cn.y, cn.z, cn.w}, cp_[4] = {cp.x, cp.y, cp.z, cp.w}, dcn_[4] = {dcn.x, dcn.y, dcn.z, dcn.w};
  float dd[4][4], dcp[4];
  float4 hm = make_float4(1.f, 1.f, 1.f, 1.f);
  if (p.drop_p > 0.f) hm = drop_mask4(p.drop_p, 1.f / (1.f - p.drop_p), p.seed, p.offset + (uint64_t)ei);
  const float hm_[4] = {hm.x, hm.y, hm.z, hm.w};
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const float dh = (dxl[u] + r1_[u] + r2_[u]) * hm_[u];   // gradient wrt the (dropped) h: LayerNorm path + recurrent paths
    const float tc = tanhf(cn_[u]);
    const float dc = dh * o_[u] * (1.f - tc * tc) + dcn_[u];
    dd[0][u] = dc * g_[u] * i_[u] * (1.f - i_[u]);
    dd[1][u] = dc * cp_[u] * f_[u] * (1.f - f_[u]);
    dd[2][u] = dc * i_[u] * (1.f - g_[u] * g_[u]);
    dd[3][u] = dh * tc * o_[u] * (1.f - o_[u]);
    dcp[u] = dc * f_[u];
  }
  if (p.dc_prev) *reinterpret_cast<float4*>(p.dc_prev + ei) = make_float4(dcp[0], dcp[1], dcp[2], dcp[3]);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int64_t col = (int64_t)k * H + h;
    const float4 d4 = make_float4(dd[k][0], dd[k][1], dd[k][2], dd[k][3]);
    if (p.dgates) *reinterpret_cast<float4*>(p.dgates + (int64_t)b * 4 * H + col) = d4;
    if (q.dgates_sum) *reinterpret_cast<float4*>(q.dgates_sum + (int64_t)b * 4 * H + col) = f4add(gs[k], d4);
    if (p.dgates2) st4any(p.dgates2, p.dgates2_dtype, (int64_t)b * p.ld_dgates2 + col, d4, vd2);
    if (p.dgatesT) {
#pragma unroll
      for (int u = 0; u < 4; ++u) st_from_float(p.dgatesT, p.dgatesT_dtype, (col + u) * p.ld_dgatesT + b, dd[k][u]);
    }
  }
}

__global__ void __launch_bounds__(512)
norm_lstm_cell_bwd_kernel(const dlsg_norm_lstm_cell_bwd_t q) {
  pdl_prologue();
  __shared__ float red[32];
  const dlsg_lstm_cell_bwd_t& p = q.cell;
  const int b = blockIdx.x, H = p.H, tid = threadIdx.x;
  const float mean = q.stats[b * 2], rstd = q.stats[b * 2 + 1];
  const bool vd2 = p.dgates2 && vecok(p.dgates2, p.dgates2_dtype, p.ld_dgates2);
  float4 xh[FS_NG], dv[FS_NG];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int e = 0; e < FS_NG; ++e) {
    const int h = (tid + e * (int)blockDim.x) * 4;
    xh[e] = make_float4(0.f, 0.f, 0.f, 0.f); dv[e] = xh[e];
    if (h < H) {
      const float4 x = ld4f(q.x + (int64_t)b * q.ldx + h);
      float4 dy = ld4f(q.dy + (int64_t)b * q.lddy + h);
      const float4 g = ld4f(q.gamma + h);
      if (q.ydrop_p > 0.f) {
        const float4 m = drop_mask4(q.ydrop_p, 1.f / (1.f - q.ydrop_p), q.yseed, q.yoffset + (uint64_t)b * H + h);
        dy.x *= m.x; dy.y *= m.y; dy.z *= m.z; dy.w *= m.w;
      }
      float4 xn = make_float4((x.x - mean) * rstd, (x.y - mean) * rstd, (x.z - mean) * rstd, (x.w - mean) * rstd);
      if (q.post_tanh) {
        const float4 be = ld4f(q.beta + h);
        float yt;
        yt = tanhf(xn.x * g.x + be.x); dy.x *= (1.f - yt * yt);
        yt = tanhf(xn.y * g.y + be.y); dy.y *= (1.f - yt * yt);
        yt = tanhf(xn.z * g.z + be.z); dy.z *= (1.f - yt * yt);
        yt = tanhf(xn.w * g.w + be.w); dy.w *= (1.f - yt * yt);
      }
      // per-row LayerNorm parameter-gradient contributions (summed over rows / time by the caller)
      *reinterpret_cast<float4*>(q.dgamma + (int64_t)b * q.ld_dparam + h) = make_float4(dy.x * xn.x, dy.y * xn.y, dy.z * xn.z, dy.w * xn.w);
      *reinterpret_cast<float4*>(q.dbeta + (int64_t)b * q.ld_dparam + h) = dy;
      const float4 d = make_float4(dy.x * g.x, dy.y * g.y, dy.z * g.z, dy.w * g.w);
      xh[e] = xn; dv[e] = d;
      s1 += (d.x + d.y) + (d.z + d.w);
      s2 += (d.x * xn.x + d.y * xn.y) + (d.z * xn.z + d.w * xn.w);
    }
  }
  s1 = block_sum(s1, red) / (float)H;
  s2 = block_sum(s2, red) / (float)H;
#pragma unroll
  for (int e = 0; e < FS_NG; ++e) {
    const int h = (tid + e * (int)blockDim.x) * 4;
    if (h < H) {
      const int64_t ei = (int64_t)b * H + h;
      const int64_t g0 = (int64_t)b * 4 * H + h;
      // all loads first
      const float4 ai = ld4f(p.acts + g0), af = ld4f(p.acts + g0 + H), ag = ld4f(p.acts + g0 + 2 * (int64_t)H), ao = ld4f(p.acts + g0 + 3 * (int64_t)H);
      const float4 cn = ld4f(p.c_new + ei);
      float4 cp = make_float4(0.f, 0.f, 0.f, 0.f), dcn = cp, r1 = cp, r2 = cp, gs[4];
      if (p.c_prev) cp = ld4f(p.c_prev + ei);
      if (p.dc_next) dcn = ld4f(p.dc_next + ei);
      if (p.dh) r1 = ld4f(p.dh + (int64_t)b * p.lddh + h);
      if (p.dh2) {
        r2 = ld4f(p.dh2 + (int64_t)b * p.lddh2 + h);
#pragma unroll
        for (int s = 1; s < 16; ++s)
          if (s < p.dh2_nsplit) r2 = f4add(r2, ld4f(p.dh2 + (int64_t)b * p.lddh2 + h + (int64_t)s * p.dh2_stride_split));
      }
      if (q.dgates_sum) {
#pragma unroll
        for (int k = 0; k < 4; ++k) gs[k] = ld4f(q.dgates_sum + (int64_t)b * 4 * H + (int64_t)k * H + h);
      }
      const float dxl[4] = {rstd * (dv[e].x - s1 - xh[e].x * s2), rstd * (dv[e].y - s1 - xh[e].y * s2),
                            rstd * (dv[e].z - s1 - xh[e].z * s2), rstd * (dv[e].w - s1 - xh[e].w * s2)};
      const float r1_[4] = {r1.x, r1.y, r1.z, r1.w}, r2_[4] = {r2.x, r2.y, r2.z, r2.w};
      const float i_[4] = {ai.x, ai.y, ai.z, ai.w}, f_[4] = {af.x, af.y, af.z, af.w}, g_[4] = {ag.x, ag.y, ag.z, ag.w}, o_[4] = {ao.x, ao.y, ao.z, ao.w};
      const float cn_[4] = {cn.x, cn.y, cn.z, cn.w}, cp_[4] = {cp.x, cp.y, cp.z, cp.w}, dcn_[4] = {dcn.x, dcn.y, dcn.z, dcn.w};
      float d[4][4], dcp[4];
      float4 hm = make_float4(1.f, 1.f, 1.f, 1.f);
      if (p.drop_p > 0.f) hm = drop_mask4(p.drop_p, 1.f / (1.f - p.drop_p), p.seed, p.offset + (uint64_t)ei);
      const float hm_[4] = {hm.x, hm.y, hm.z, hm.w};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float dh = (dxl[u] + r1_[u] + r2_[u]) * hm_[u];   // gradient wrt the (dropped) h: LayerNorm path + recurrent paths
        const float tc = tanhf(cn_[u]);
        const float dc = dh * o_[u] * (1.f - tc * tc) + dcn_[u];
        d[0][u] = dc * g_[u] * i_[u] * (1.f - i_[u]);
        d[1][u] = dc * cp_[u] * f_[u] * (1.f - f_[u]);
        d[2][u] = dc * i_[u] * (1.f - g_[u] * g_[u]);
        d[3][u] = dh * tc * o_[u] * (1.f - o_[u]);
        dcp[u] = dc * f_[u];
      }
      if (p.dc_prev) *reinterpret_cast<float4*>(p.dc_prev + ei) = make_float4(dcp[0], dcp[1], dcp[2], dcp[3]);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int64_t col = (int64_t)k * H + h;
        const float4 d4 = make_float4(d[k][0], d[k][1], d[k][2], d[k][3]);
        if (p.dgates) *reinterpret_cast<float4*>(p.dgates + (int64_t)b * 4 * H + col) = d4;
        if (q.dgates_sum) *reinterpret_cast<float4*>(q.dgates_sum + (int64_t)b * 4 * H + col) = f4add(gs[k], d4);
        if (p.dgates2) st4any(p.dgates2, p.dgates2_dtype, (int64_t)b * p.ld_dgates2 + col, d4, vd2);
        if (p.dgatesT) {
#pragma unroll
          for (int u = 0; u < 4; ++u) st_from_float(p.dgatesT, p.dgatesT_dtype, (col + u) * p.ld_dgatesT + b, d[k][u]);
        }
      }
    }
  }
}

}  // namespace dlsg

using namespace dlsg;

static inline bool a16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
// one float4 group per thread up to H = 2048 (512 threads); whole warps
static inline int fs_threads(int H) { int t = ((H / 4 + 31) / 32) * 32; return t < 64 ? 64 : (t > 512 ? 512 : t); }

extern "C" {

int dlsg_fused_step_supported(int32_t H) { return (H % 4 == 0 && H <= FS_MAXH) ? 1 : 0; }

int dlsg_lstm_cell_norm_fwd(const dlsg_lstm_cell_norm_fwd_t* q, void* stream) {
  const dlsg_lstm_cell_fwd_t& p = q->cell;
  DLSG_REQUIRE(p.B > 0 && p.H > 0 && dlsg_fused_step_supported(p.H) && p.nsplit >= 1, "lstm_cell_norm_fwd: H=%d unsupported (H %% 4 == 0, H <= %d)", p.H, FS_MAXH);
  DLSG_REQUIRE(a16(p.gates) && p.stride_split % 4 == 0 && (!p.row_bias || (a16(p.row_bias) && p.ld_row_bias % 4 == 0)) && (!p.bias || a16(p.bias)) &&
               (!p.c_prev || a16(p.c_prev)) && a16(p.c_out) && (!p.h_out || a16(p.h_out)) && a16(q->gamma) && a16(q->beta),
               "lstm_cell_norm_fwd: fp32 operands must be 16-byte aligned");
  DLSG_REQUIRE(p.offset % 4 == 0 && q->yoffset % 4 == 0, "lstm_cell_norm_fwd: dropout offsets must be multiples of 4");
  const int nt = fs_threads(p.H);
  if (p.H <= 1536 && p.nsplit <= 4) {        // <= 384 threads: the register budget of the all-loads-first kernel
    cudaStream_t st = (cudaStream_t)stream;
    switch (p.nsplit) {
      case 1: DLSG_LAUNCH(lstm_cell_norm_fwd_fast<1>, p.B, nt, 0, st, *q); break;
      case 2: DLSG_LAUNCH(lstm_cell_norm_fwd_fast<2>, p.B, nt, 0, st, *q); break;
      case 3: DLSG_LAUNCH(lstm_cell_norm_fwd_fast<3>, p.B, nt, 0, st, *q); break;
      default: DLSG_LAUNCH(lstm_cell_norm_fwd_fast<4>, p.B, nt, 0, st, *q); break;
    }
    return check_launch("lstm_cell_norm_fwd_fast");
  }
  DLSG_LAUNCH(lstm_cell_norm_fwd_kernel, p.B, nt, 0, (cudaStream_t)stream, *q);
  return check_launch("lstm_cell_norm_fwd_kernel");
}

int dlsg_cell_norm_attn2_supported(const dlsg_cell_norm_attn2_fwd_t* f) {
  const dlsg_lstm_cell_fwd_t& p = f->cn.cell;
  const dlsg_attn2_fwd_t& a = f->at;
  return (p.H % 4 == 0 && p.H <= 1024 && p.H == a.Hk && a.Hv % 4 == 0 && a.Hv <= 1024 && a.nh >= 1 && a.nh <= 2 && a.P >= 1 && a.P <= FA_PM &&
          p.nsplit >= 1 && p.nsplit <= 4 && p.B == a.rows && a.rows_per_node >= 1) ? 1 : 0;
}
int dlsg_cell_norm_attn2_fwd(const dlsg_cell_norm_attn2_fwd_t* f, void* stream) {
  const dlsg_lstm_cell_fwd_t& p = f->cn.cell;
  const dlsg_attn2_fwd_t& a = f->at;
  DLSG_REQUIRE(dlsg_cell_norm_attn2_supported(f), "cell_norm_attn2_fwd: unsupported shape (H=%d Hk=%d Hv=%d nh=%d P=%d nsplit=%d)", p.H, a.Hk, a.Hv, a.nh, a.P, p.nsplit);
  DLSG_REQUIRE(a16(p.gates) && p.stride_split % 4 == 0 && (!p.row_bias || (a16(p.row_bias) && p.ld_row_bias % 4 == 0)) && (!p.bias || a16(p.bias)) &&
               (!p.c_prev || a16(p.c_prev)) && a16(p.c_out) && (!p.h_out || a16(p.h_out)) && a16(f->cn.gamma) && a16(f->cn.beta),
               "cell_norm_attn2_fwd: fp32 cell operands must be 16-byte aligned");
  DLSG_REQUIRE(p.offset % 4 == 0 && f->cn.yoffset % 4 == 0 && a.offset % 4 == 0, "cell_norm_attn2_fwd: dropout offsets must be multiples of 4");
  DLSG_REQUIRE(a.KW && a.VW && a.co && a16(a.KW) && a16(a.VW) && a16(a.co) && a.ldco % 4 == 0, "cell_norm_attn2_fwd: attention operands must be 16-byte aligned");
  DLSG_REQUIRE(!a.y || (a.gamma[0] && a.beta[0] && (a.nh < 2 || (a.gamma[1] && a.beta[1])) && a.ldy % 4 == 0 &&
                        (reinterpret_cast<uintptr_t>(a.y) % (a.y_dtype == DLSG_F32 ? 16 : 8)) == 0),
               "cell_norm_attn2_fwd: output-layer operands missing or unaligned");
  const int nt = 256 * a.nh;
  cudaStream_t st = (cudaStream_t)stream;
  switch (p.nsplit) {
    case 1: DLSG_LAUNCH(cell_norm_attn2_fwd_kernel<1>, p.B, nt, 0, st, *f); break;
    case 2: DLSG_LAUNCH(cell_norm_attn2_fwd_kernel<2>, p.B, nt, 0, st, *f); break;
    case 3: DLSG_LAUNCH(cell_norm_attn2_fwd_kernel<3>, p.B, nt, 0, st, *f); break;
    default: DLSG_LAUNCH(cell_norm_attn2_fwd_kernel<4>, p.B, nt, 0, st, *f); break;
  }
  return check_launch("cell_norm_attn2_fwd_kernel");
}

int dlsg_norm_lstm_cell_bwd(const dlsg_norm_lstm_cell_bwd_t* q, void* stream) {
  const dlsg_lstm_cell_bwd_t& p = q->cell;
  DLSG_REQUIRE(p.B > 0 && p.H > 0 && dlsg_fused_step_supported(p.H), "norm_lstm_cell_bwd: H=%d unsupported", p.H);
  DLSG_REQUIRE(q->dgamma && q->dbeta && q->stats && q->ld_dparam % 4 == 0 && a16(q->dgamma) && a16(q->dbeta), "norm_lstm_cell_bwd: dgamma/dbeta rows (B,H) and stats required");
  DLSG_REQUIRE(a16(q->x) && q->ldx % 4 == 0 && a16(q->dy) && q->lddy % 4 == 0 && a16(q->gamma) && a16(q->beta) && a16(p.acts) && a16(p.c_new) &&
               (!p.c_prev || a16(p.c_prev)) && (!p.dc_next || a16(p.dc_next)) && (!p.dc_prev || a16(p.dc_prev)) &&
               (!p.dh || (a16(p.dh) && p.lddh % 4 == 0)) && (!p.dh2 || (a16(p.dh2) && p.lddh2 % 4 == 0)) &&
               (!p.dgates || a16(p.dgates)) && (!q->dgates_sum || a16(q->dgates_sum)),
               "norm_lstm_cell_bwd: fp32 operands must be 16-byte aligned with row pitches multiple of 4");
  DLSG_REQUIRE(p.offset % 4 == 0 && q->yoffset % 4 == 0, "norm_lstm_cell_bwd: dropout offsets must be multiples of 4");
  const int nt = fs_threads(p.H);
  const int ns = (p.dh2 && p.dh2_nsplit > 1) ? p.dh2_nsplit : 1;
  if (p.H <= 1536 && ns <= 4) {
    cudaStream_t st = (cudaStream_t)stream;
    switch (ns) {
      case 1: DLSG_LAUNCH(norm_lstm_cell_bwd_fast<1>, p.B, nt, 0, st, *q); break;
      case 2: DLSG_LAUNCH(norm_lstm_cell_bwd_fast<2>, p.B, nt, 0, st, *q); break;
      case 3: DLSG_LAUNCH(norm_lstm_cell_bwd_fast<3>, p.B, nt, 0, st, *q); break;
      default: DLSG_LAUNCH(norm_lstm_cell_bwd_fast<4>, p.B, nt, 0, st, *q); break;
    }
    return check_launch("norm_lstm_cell_bwd_fast");
  }
  DLSG_LAUNCH(norm_lstm_cell_bwd_kernel, p.B, nt, 0, (cudaStream_t)stream, *q);
  return check_launch("norm_lstm_cell_bwd_kernel");
}

}  // extern "C"
