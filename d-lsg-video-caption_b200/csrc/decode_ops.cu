// Per-decode-step fused kernels: the AttentionShare core (score / softmax / weighted sum), per-row
// vocabulary kernels (argmax, log-softmax, masked cross-entropy fwd+bwd) and the beam-search kernels
// (log-softmax + after-<end> mask + warp-level top-k, candidate merge, state gather, back-track).
#include "common.cuh"

namespace dlsg {

constexpr int MAXP = 128;   // max nodes per attention (P=5/8 latent nodes, 26/52 frames in the baseline)
constexpr int MAXK = 16;    // max beam / per-node beam

// ------------------------------------------------------------------------------------------- node attention
// grid (rows, nh); 128 threads.  logits_p = K_p . q / sqrt(H); alpha = softmax_p; ctx = sum_p alpha_p V_p
__global__ void __launch_bounds__(128)
node_attn_fwd_kernel(const dlsg_node_attn_fwd_t p) {
  pdl_prologue();
  __shared__ float lg[MAXP];
  const int r = blockIdx.x, hd = blockIdx.y;
  const int node = r / p.rows_per_node;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const float* q = p.qp + ((int64_t)r * p.nh + hd) * p.H;
  const float* K = p.Kp + (((int64_t)hd * p.nodes + node) * p.P) * p.H;
  const float* V = p.Vp + (((int64_t)hd * p.nodes + node) * p.P) * p.H;
  const float scale = 1.0f / sqrtf((float)p.H);
  for (int j = w; j < p.P; j += nw) {
    float s = 0.f;
    for (int c = lane; c < p.H; c += 32) s = fmaf(K[(int64_t)j * p.H + c], q[c], s);
    s = warp_sum(s);
    if (lane == 0) lg[j] = s * scale;
  }
  __syncthreads();
  float mx = -INFINITY;
  for (int j = 0; j < p.P; ++j) mx = fmaxf(mx, lg[j]);
  float sum = 0.f;
  for (int j = 0; j < p.P; ++j) sum += expf(lg[j] - mx);
  const float inv = 1.f / sum;
  __syncthreads();
  for (int j = threadIdx.x; j < p.P; j += blockDim.x) {
    const float a = expf(lg[j] - mx) * inv;
    lg[j] = a;
    if (p.alpha) p.alpha[(int64_t)r * p.ldalpha + hd * p.P + j] = a;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < p.H; c += blockDim.x) {
    float acc = 0.f;
    for (int j = 0; j < p.P; ++j) acc = fmaf(lg[j], V[(int64_t)j * p.H + c], acc);
    st_from_float(p.ctx, p.ctx_dtype, (int64_t)r * p.ldctx + hd * p.H + c, acc);
  }
}

// backward of one step (training: one row per node set).  dKp / dVp accumulate across the 26 steps.
__global__ void __launch_bounds__(128)
node_attn_bwd_kernel(const dlsg_node_attn_bwd_t p) {
  pdl_prologue();
  __shared__ float da[MAXP];
  __shared__ float dl[MAXP];
  const int r = blockIdx.x, hd = blockIdx.y;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const float* q = p.qp + ((int64_t)r * p.nh + hd) * p.H;
  const int64_t nb = (((int64_t)hd * p.rows + r) * p.P) * p.H;   // layout (nh, rows, P, H)
  const float* K = p.Kp + nb;
  const float* V = p.Vp + nb;
  const float* dctx = p.dctx + (int64_t)r * p.lddctx + hd * p.H;
  const float* al = p.alpha + (int64_t)r * p.ldalpha + hd * p.P;
  const float scale = 1.0f / sqrtf((float)p.H);
  for (int j = w; j < p.P; j += nw) {
    float s = 0.f;
    for (int c = lane; c < p.H; c += 32) s = fmaf(V[(int64_t)j * p.H + c], dctx[c], s);
    s = warp_sum(s);
    if (lane == 0) da[j] = s + (p.dalpha_ext ? p.dalpha_ext[(int64_t)r * p.ldalpha + hd * p.P + j] : 0.f);
  }
  __syncthreads();
  float dot = 0.f;
  for (int j = 0; j < p.P; ++j) dot = fmaf(al[j], da[j], dot);
  __syncthreads();
  for (int j = threadIdx.x; j < p.P; j += blockDim.x) dl[j] = al[j] * (da[j] - dot) * scale;
  __syncthreads();
  const int64_t dq0 = ((int64_t)r * p.nh + hd) * p.H;
  for (int c = threadIdx.x; c < p.H; c += blockDim.x) {
    float acc = 0.f;
    const float qc = q[c], dc = dctx[c];
    for (int j = 0; j < p.P; ++j) {
      acc = fmaf(dl[j], K[(int64_t)j * p.H + c], acc);
      p.dKp[nb + (int64_t)j * p.H + c] += dl[j] * qc;
      p.dVp[nb + (int64_t)j * p.H + c] += al[j] * dc;
    }
    st_from_float(p.dqp, p.dqp_dtype, dq0 + c, acc);
  }
}


// ---- fast paths: P <= 8 nodes, H <= 1024, H % 4 == 0.  256 threads, thread t owns columns [4t, 4t+4); all K / V / q
// (and dctx) loads of the step are issued back to back (one memory latency), P dot products reduced by shuffles.
constexpr int APM = 8;

__device__ __forceinline__ float dot4(const float4 a, const float4 b) { return (a.x * b.x + a.y * b.y) + (a.z * b.z + a.w * b.w); }

__global__ void __launch_bounds__(256)
node_attn_fwd_fast(const dlsg_node_attn_fwd_t p) {
  pdl_prologue();
  __shared__ float red[8][APM];
  __shared__ float al[APM];
  const int r = blockIdx.x, hd = blockIdx.y, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int node = r / p.rows_per_node;
  const int c = tid * 4;
  const bool act = c < p.H;
  const float* q = p.qp + ((int64_t)r * p.nh + hd) * p.H;
  const float* K = p.Kp + (((int64_t)hd * p.nodes + node) * p.P) * p.H;
  const float* V = p.Vp + (((int64_t)hd * p.nodes + node) * p.P) * p.H;
  float4 k4[APM], v4[APM], q4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (act) q4 = *reinterpret_cast<const float4*>(q + c);
#pragma unroll
  for (int j = 0; j < APM; ++j) {
    k4[j] = make_float4(0.f, 0.f, 0.f, 0.f); v4[j] = k4[j];
    if (act && j < p.P) {
      k4[j] = *reinterpret_cast<const float4*>(K + (int64_t)j * p.H + c);
      v4[j] = *reinterpret_cast<const float4*>(V + (int64_t)j * p.H + c);
    }
  }
#pragma unroll
  for (int j = 0; j < APM; ++j) {
    const float s = warp_sum(dot4(k4[j], q4));
    if (lane == 0) red[w][j] = s;
  }
  __syncthreads();
  if (tid == 0) {
    const float scale = 1.0f / sqrtf((float)p.H);
    float lg[APM], mx = -INFINITY;
    for (int j = 0; j < p.P; ++j) {
      float s = 0.f;
      for (int ww = 0; ww < 8; ++ww) s += red[ww][j];
      lg[j] = s * scale; mx = fmaxf(mx, lg[j]);
    }
    float sum = 0.f;
    for (int j = 0; j < p.P; ++j) { lg[j] = expf(lg[j] - mx); sum += lg[j]; }
    const float inv = 1.f / sum;
    for (int j = 0; j < p.P; ++j) {
      const float a = lg[j] * inv;
      al[j] = a;
      if (p.alpha) p.alpha[(int64_t)r * p.ldalpha + hd * p.P + j] = a;
    }
  }
  __syncthreads();
  if (act) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < APM; ++j) {
      if (j < p.P) { const float a = al[j]; acc.x = fmaf(a, v4[j].x, acc.x); acc.y = fmaf(a, v4[j].y, acc.y); acc.z = fmaf(a, v4[j].z, acc.z); acc.w = fmaf(a, v4[j].w, acc.w); }
    }
    const int64_t o = (int64_t)r * p.ldctx + hd * p.H + c;
    if (p.ctx_dtype == DLSG_F32) *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.ctx) + o) = acc;
    else {
      __nv_bfloat162 a = __floats2bfloat162_rn(acc.x, acc.y), b = __floats2bfloat162_rn(acc.z, acc.w);
      uint2 u; u.x = *reinterpret_cast<uint32_t*>(&a); u.y = *reinterpret_cast<uint32_t*>(&b);
      *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.ctx) + o) = u;
    }
  }
}

__global__ void __launch_bounds__(256)
node_attn_bwd_fast(const dlsg_node_attn_bwd_t p) {
  pdl_prologue();
  __shared__ float red[8][APM];
  __shared__ float dl[APM];
  const int r = blockIdx.x, hd = blockIdx.y, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int c = tid * 4;
  const bool act = c < p.H;
  const float* q = p.qp + ((int64_t)r * p.nh + hd) * p.H;
  const int64_t nb = (((int64_t)hd * p.rows + r) * p.P) * p.H;
  const float* al = p.alpha + (int64_t)r * p.ldalpha + hd * p.P;
  float4 k4[APM], v4[APM], q4 = make_float4(0.f, 0.f, 0.f, 0.f), d4 = q4;
  if (act) {
    q4 = *reinterpret_cast<const float4*>(q + c);
    d4 = *reinterpret_cast<const float4*>(p.dctx + (int64_t)r * p.lddctx + hd * p.H + c);
  }
#pragma unroll
  for (int j = 0; j < APM; ++j) {
    k4[j] = make_float4(0.f, 0.f, 0.f, 0.f); v4[j] = k4[j];
    if (act && j < p.P) {
      k4[j] = *reinterpret_cast<const float4*>(p.Kp + nb + (int64_t)j * p.H + c);
      v4[j] = *reinterpret_cast<const float4*>(p.Vp + nb + (int64_t)j * p.H + c);
    }
  }
#pragma unroll
  for (int j = 0; j < APM; ++j) {
    const float s = warp_sum(dot4(v4[j], d4));
    if (lane == 0) red[w][j] = s;
  }
  __syncthreads();
  if (tid == 0) {
    const float scale = 1.0f / sqrtf((float)p.H);
    float da[APM], dot = 0.f;
    for (int j = 0; j < p.P; ++j) {
      float s = 0.f;
      for (int ww = 0; ww < 8; ++ww) s += red[ww][j];
      if (p.dalpha_ext) s += p.dalpha_ext[(int64_t)r * p.ldalpha + hd * p.P + j];
      da[j] = s; dot = fmaf(al[j], s, dot);
    }
    for (int j = 0; j < p.P; ++j) dl[j] = al[j] * (da[j] - dot) * scale;
  }
  __syncthreads();
  if (act) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < APM; ++j) {
      if (j < p.P) {
        const float g = dl[j], a = al[j];
        acc.x = fmaf(g, k4[j].x, acc.x); acc.y = fmaf(g, k4[j].y, acc.y); acc.z = fmaf(g, k4[j].z, acc.z); acc.w = fmaf(g, k4[j].w, acc.w);
        float4* dk = reinterpret_cast<float4*>(p.dKp + nb + (int64_t)j * p.H + c);
        float4* dv = reinterpret_cast<float4*>(p.dVp + nb + (int64_t)j * p.H + c);
        float4 ok = *dk, ov = *dv;
        ok.x = fmaf(g, q4.x, ok.x); ok.y = fmaf(g, q4.y, ok.y); ok.z = fmaf(g, q4.z, ok.z); ok.w = fmaf(g, q4.w, ok.w);
        ov.x = fmaf(a, d4.x, ov.x); ov.y = fmaf(a, d4.y, ov.y); ov.z = fmaf(a, d4.z, ov.z); ov.w = fmaf(a, d4.w, ov.w);
        *dk = ok; *dv = ov;
      }
    }
    const int64_t o = ((int64_t)r * p.nh + hd) * p.H + c;
    if (p.dqp_dtype == DLSG_F32) *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.dqp) + o) = acc;
    else {
      __nv_bfloat162 a = __floats2bfloat162_rn(acc.x, acc.y), b = __floats2bfloat162_rn(acc.z, acc.w);
      uint2 u; u.x = *reinterpret_cast<uint32_t*>(&a); u.y = *reinterpret_cast<uint32_t*>(&b);
      *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.dqp) + o) = u;
    }
  }
}

// ---- hoisted AttentionShare (P <= 8, Hk,Hv <= 1024): the query / output projections are folded into the node
// tensors once per sequence (KW = K Wq, VW = V Wo^T), so a decode step needs NO per-step attention GEMM:
//   logits_p = KW_p . q * scale ; alpha = softmax_p ; co = sum_p alpha_p VW_p   (pre-LayerNorm context, sublayer.py:29-41)
__global__ void __launch_bounds__(256)
attn2_fwd_kernel(const dlsg_attn2_fwd_t p) {
  pdl_prologue();
  __shared__ float red[8][APM];
  __shared__ float al[APM];
  const int r = blockIdx.x, hd = blockIdx.y, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int node = r / p.rows_per_node;
  const int c = tid * 4;
  const bool ak = c < p.Hk, av = c < p.Hv;
  const float* KW = p.KW + (((int64_t)hd * p.nodes + node) * p.P) * p.Hk;
  const float* VW = p.VW + (((int64_t)hd * p.nodes + node) * p.P) * p.Hv;
  float4 k4[APM], v4[APM], q4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (ak) q4 = *reinterpret_cast<const float4*>(p.q + (int64_t)r * p.ldq + c);
#pragma unroll
  for (int j = 0; j < APM; ++j) {
    k4[j] = make_float4(0.f, 0.f, 0.f, 0.f); v4[j] = k4[j];
    if (j < p.P) {
      if (ak) k4[j] = *reinterpret_cast<const float4*>(KW + (int64_t)j * p.Hk + c);
      if (av) v4[j] = *reinterpret_cast<const float4*>(VW + (int64_t)j * p.Hv + c);
    }
  }
#pragma unroll
  for (int j = 0; j < APM; ++j) {
    const float s = warp_sum(dot4(k4[j], q4));
    if (lane == 0) red[w][j] = s;
  }
  __syncthreads();
  if (tid == 0) {
    float lg[APM], mx = -INFINITY;
    for (int j = 0; j < p.P; ++j) {
      float s = 0.f;
      for (int ww = 0; ww < 8; ++ww) s += red[ww][j];
      lg[j] = s * p.scale; mx = fmaxf(mx, lg[j]);
    }
    float sum = 0.f;
    for (int j = 0; j < p.P; ++j) { lg[j] = expf(lg[j] - mx); sum += lg[j]; }
    const float inv = 1.f / sum;
    for (int j = 0; j < p.P; ++j) {
      const float a = lg[j] * inv;
      al[j] = a;
      if (p.alpha) p.alpha[(int64_t)r * p.ldalpha + hd * p.P + j] = a;
    }
  }
  __syncthreads();
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (av) {
#pragma unroll
    for (int j = 0; j < APM; ++j) {
      if (j < p.P) { const float a = al[j]; acc.x = fmaf(a, v4[j].x, acc.x); acc.y = fmaf(a, v4[j].y, acc.y); acc.z = fmaf(a, v4[j].z, acc.z); acc.w = fmaf(a, v4[j].w, acc.w); }
    }
    *reinterpret_cast<float4*>(p.co + (int64_t)r * p.ldco + hd * p.Hv + c) = acc;
  }
  if (p.y == nullptr) return;                       // uniform
  // fused context output layer: tanh -> LayerNorm (two-pass statistics over the head's Hv values) -> dropout
  __shared__ float red2[32];
  float4 t4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (av) t4 = make_float4(tanhf(acc.x), tanhf(acc.y), tanhf(acc.z), tanhf(acc.w));
  const float* gam = hd ? p.gamma[1] : p.gamma[0];
  const float* bet = hd ? p.beta[1] : p.beta[0];
  float4 g4 = t4, b4 = t4;
  if (av) { g4 = *reinterpret_cast<const float4*>(gam + c); b4 = *reinterpret_cast<const float4*>(bet + c); }   // in flight during the reductions
  const float invH = 1.f / (float)p.Hv;
  const float mean = block_sum(av ? (t4.x + t4.y) + (t4.z + t4.w) : 0.f, red2) * invH;
  float sq = 0.f;
  if (av) { const float a = t4.x - mean, b = t4.y - mean, cc = t4.z - mean, d = t4.w - mean; sq = (a * a + b * b) + (cc * cc + d * d); }
  const float rstd = rsqrtf(block_sum(sq, red2) * invH + 1e-5f);
  if (p.stats && tid == 0) {
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
    const int64_t o = (int64_t)r * p.ldy + hd * p.Hv + c;
    if (p.y_dtype == DLSG_F32) *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.y) + o) = y;
    else {
      __nv_bfloat162 a = __floats2bfloat162_rn(y.x, y.y), b = __floats2bfloat162_rn(y.z, y.w);
      uint2 u; u.x = *reinterpret_cast<uint32_t*>(&a); u.y = *reinterpret_cast<uint32_t*>(&b);
      *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.y) + o) = u;
    }
  }
}

// grid (rows, nh), block 256: one CTA per (row, head); the heads' dq contributions are ADDED to dq with fp32 atomics when
// nh > 1 (half the work per CTA and twice the CTAs of the former (rows) x 512-thread form: 10.5 -> ~6 us per decode step)
__global__ void __launch_bounds__(256)
attn2_bwd_kernel(const dlsg_attn2_bwd_t p) {
  pdl_prologue();
  __shared__ float red[2][8][APM];
  __shared__ float dl[2][APM];
  // one CTA per (row, head): the heads only meet in dq, which they accumulate with fp32 atomics (nh > 1)
  const int r = blockIdx.x, hd = blockIdx.y, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int c = tid * 4;
  const bool ak = c < p.Hk, av = c < p.Hv;
  const int64_t nk = (((int64_t)hd * p.rows + r) * p.P) * p.Hk, nv = (((int64_t)hd * p.rows + r) * p.P) * p.Hv;
  const float* al = p.alpha + (int64_t)r * p.ldalpha + hd * p.P;
  float4 k4[APM], v4[APM], q4 = make_float4(0.f, 0.f, 0.f, 0.f), d4 = q4;
  if (ak) q4 = *reinterpret_cast<const float4*>(p.q + (int64_t)r * p.ldq + c);
  if (av && !p.dy) d4 = *reinterpret_cast<const float4*>(p.dco + (int64_t)r * p.lddco + hd * p.Hv + c);
#pragma unroll
  for (int j = 0; j < APM; ++j) {
    k4[j] = make_float4(0.f, 0.f, 0.f, 0.f); v4[j] = k4[j];
    if (j < p.P) {
      if (ak) k4[j] = *reinterpret_cast<const float4*>(p.KW + nk + (int64_t)j * p.Hk + c);
      if (av) v4[j] = *reinterpret_cast<const float4*>(p.VW + nv + (int64_t)j * p.Hv + c);
    }
  }
  if (p.dy) {                                       // uniform
    // fused head: backward of dropout(LN(tanh(co))) for this (row, head): d4 = gradient wrt co
    __shared__ float gr[2][2][8];
    const float* st = p.stats + (int64_t)hd * p.stats_head_stride + 2 * (int64_t)r;
    const float mean = st[0], rstd = st[1];
    float4 t4 = make_float4(0.f, 0.f, 0.f, 0.f), xh = t4, d = t4;
    float s1 = 0.f, s2 = 0.f;
    if (av) {
      const int64_t o = (int64_t)hd * p.Hv + c;
      const float4 x4 = *reinterpret_cast<const float4*>(p.co + (int64_t)r * p.ldco + o);
      float4 dy4 = *reinterpret_cast<const float4*>(p.dy + (int64_t)r * p.lddy + o);
      const float4 g4 = *reinterpret_cast<const float4*>((hd ? p.gamma[1] : p.gamma[0]) + c);
      if (p.drop_p > 0.f) {
        const float4 m = drop_mask4(p.drop_p, 1.f / (1.f - p.drop_p), p.seed,
                                    p.offset + (uint64_t)hd * p.offset_head_stride + (uint64_t)r * p.Hv + c);
        dy4.x *= m.x; dy4.y *= m.y; dy4.z *= m.z; dy4.w *= m.w;
      }
      t4 = make_float4(tanhf(x4.x), tanhf(x4.y), tanhf(x4.z), tanhf(x4.w));
      xh = make_float4((t4.x - mean) * rstd, (t4.y - mean) * rstd, (t4.z - mean) * rstd, (t4.w - mean) * rstd);
      *reinterpret_cast<float4*>(p.dgamma_rows + (int64_t)r * p.ld_dparam + o) = make_float4(dy4.x * xh.x, dy4.y * xh.y, dy4.z * xh.z, dy4.w * xh.w);
      *reinterpret_cast<float4*>(p.dbeta_rows + (int64_t)r * p.ld_dparam + o) = dy4;
      d = make_float4(dy4.x * g4.x, dy4.y * g4.y, dy4.z * g4.z, dy4.w * g4.w);
      s1 = (d.x + d.y) + (d.z + d.w);
      s2 = (d.x * xh.x + d.y * xh.y) + (d.z * xh.z + d.w * xh.w);
    }
    s1 = warp_sum(s1); s2 = warp_sum(s2);
    if (lane == 0) { gr[0][hd][w] = s1; gr[1][hd][w] = s2; }
    __syncthreads();
    s1 = 0.f; s2 = 0.f;
#pragma unroll
    for (int ww = 0; ww < 8; ++ww) { s1 += gr[0][hd][ww]; s2 += gr[1][hd][ww]; }
    const float invH = 1.f / (float)p.Hv;
    s1 *= invH; s2 *= invH;
    d4.x = rstd * (d.x - s1 - xh.x * s2) * (1.f - t4.x * t4.x);
    d4.y = rstd * (d.y - s1 - xh.y * s2) * (1.f - t4.y * t4.y);
    d4.z = rstd * (d.z - s1 - xh.z * s2) * (1.f - t4.z * t4.z);
    d4.w = rstd * (d.w - s1 - xh.w * s2) * (1.f - t4.w * t4.w);
    if (!av) d4 = make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int j = 0; j < APM; ++j) {
    const float s = warp_sum(dot4(v4[j], d4));
    if (lane == 0) red[hd][w][j] = s;
  }
  __syncthreads();
  if (tid == 0) {
    float da[APM], dot = 0.f;
    for (int j = 0; j < p.P; ++j) {
      float s = 0.f;
      for (int ww = 0; ww < 8; ++ww) s += red[hd][ww][j];
      if (p.dalpha_ext) s += p.dalpha_ext[(int64_t)r * p.ldalpha + hd * p.P + j];
      da[j] = s; dot = fmaf(al[j], s, dot);
    }
    for (int j = 0; j < p.P; ++j) {
      const float v = al[j] * (da[j] - dot) * p.scale;
      dl[hd][j] = v;
      if (p.dl_save) p.dl_save[(int64_t)r * p.ld_dl_save + hd * p.P + j] = v;
    }
  }
  __syncthreads();
  // deferred node gradients (dlsg_attn2_bwd_nodes after the time loop): this step only records dl and d(co); the per-step
  // read-modify-write of dKW / dVW (2 x 80 KB per row and step at nh = 2, P = 5, H = 1024) disappears
  const bool defer = p.dl_save != nullptr;
  if (defer && av) *reinterpret_cast<float4*>(p.dco_save + (int64_t)r * p.ld_dco_save + hd * p.Hv + c) = d4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int j = 0; j < APM; ++j) {
    if (j < p.P) {
      const float g = dl[hd][j], a = al[j];
      if (ak) {
        acc.x = fmaf(g, k4[j].x, acc.x); acc.y = fmaf(g, k4[j].y, acc.y); acc.z = fmaf(g, k4[j].z, acc.z); acc.w = fmaf(g, k4[j].w, acc.w);
      }
      if (defer) continue;
      if (ak) {
        float4* dk = reinterpret_cast<float4*>(p.dKW + nk + (int64_t)j * p.Hk + c);
        float4 ok = *dk;
        ok.x = fmaf(g, q4.x, ok.x); ok.y = fmaf(g, q4.y, ok.y); ok.z = fmaf(g, q4.z, ok.z); ok.w = fmaf(g, q4.w, ok.w);
        *dk = ok;
      }
      if (av) {
        float4* dv = reinterpret_cast<float4*>(p.dVW + nv + (int64_t)j * p.Hv + c);
        float4 ov = *dv;
        ov.x = fmaf(a, d4.x, ov.x); ov.y = fmaf(a, d4.y, ov.y); ov.z = fmaf(a, d4.z, ov.z); ov.w = fmaf(a, d4.w, ov.w);
        *dv = ov;
      }
    }
  }
  if (ak) {
    float* dq = p.dq + (int64_t)r * p.lddq + c;
    if (p.nh > 1) {
      atomicAdd(dq, acc.x); atomicAdd(dq + 1, acc.y); atomicAdd(dq + 2, acc.z); atomicAdd(dq + 3, acc.w);
    } else {
      float4 old = *reinterpret_cast<float4*>(dq);
      old.x += acc.x; old.y += acc.y; old.z += acc.z; old.w += acc.w;
      *reinterpret_cast<float4*>(dq) = old;
    }
  }
}

static inline bool al16(const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; }

// ------------------------------------------------------------------------------------------- vocab rows
struct ArgMax { float v; int i; };
__device__ __forceinline__ ArgMax better(ArgMax a, ArgMax b) {   // max value, lowest index on ties
  if (b.v > a.v || (b.v == a.v && b.i < a.i)) return b;
  return a;
}
__device__ __forceinline__ ArgMax warp_argmax(ArgMax a) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ArgMax b; b.v = __shfl_xor_sync(0xffffffffu, a.v, o); b.i = __shfl_xor_sync(0xffffffffu, a.i, o);
    a = better(a, b);
  }
  return a;
}

// ---- vocabulary rows: one CTA per row.  A row (V ~ 1e4 fp32, arbitrary pitch, so its base is only 4-byte aligned) is walked
// as [scalar head up to the first 16-byte boundary | float4 body | scalar tail]; every thread issues VR_UNROLL independent
// 128-bit loads before it touches any of them (the old one-scalar-load-per-iteration loops were a chain of L2 round trips:
// 31 us for a 27 MB top-k).  `F(v, col)` is applied to every element.
constexpr int VR_UNROLL = 8;

template <typename F1, typename F4>
__device__ __forceinline__ void row_foreach(const float* __restrict__ x, int V, F1&& f1, F4&& f4) {
  const int tid = threadIdx.x, nt = blockDim.x;
  int head = (int)(((16u - (uint32_t)(reinterpret_cast<uintptr_t>(x) & 15u)) & 15u) >> 2);
  if (head > V) head = V;
  if (tid < head) f1(x[tid], tid);
  const float4* __restrict__ x4 = reinterpret_cast<const float4*>(x + head);
  const int n4 = (V - head) >> 2;
  int j = tid;
  for (; j + (VR_UNROLL - 1) * nt < n4; j += VR_UNROLL * nt) {
    float4 v[VR_UNROLL];
#pragma unroll
    for (int u = 0; u < VR_UNROLL; ++u) v[u] = x4[j + u * nt];
#pragma unroll
    for (int u = 0; u < VR_UNROLL; ++u) f4(v[u], head + 4 * (j + u * nt));
  }
  for (; j < n4; j += nt) f4(x4[j], head + 4 * j);
  const int t0 = head + 4 * n4;
  if (t0 + tid < V) f1(x[t0 + tid], t0 + tid);
}
template <typename F>
__device__ __forceinline__ void row_foreach(const float* __restrict__ x, int V, F&& f) {
  row_foreach(x, V, f, [&](const float4 v, int c) { f(v.x, c); f(v.y, c + 1); f(v.z, c + 2); f(v.w, c + 3); });
}

// running (max, sum of exp(x - max)) of the values a thread has seen; add4 rescales at most once per 128-bit load, so the
// common case is 4 x (subtract, ex2, add) with no per-element branch
struct OnlineLse {
  float m, s;
  __device__ __forceinline__ void init() { m = -INFINITY; s = 0.f; }
  __device__ __forceinline__ void add(float v) {
    if (v > m) { s = s * __expf(m - v) + 1.f; m = v; }       // (m = -inf: s = 0 * exp(-inf) + 1 = 1)
    else if (m > -INFINITY) s += __expf(v - m);              // (only -inf seen so far: nothing to add, exp(-inf + inf) is NaN)
  }
  __device__ __forceinline__ void add4(const float4 v) {
    const float c = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
    if (c > m) { s *= __expf(m - c); m = c; }
    if (m > -INFINITY) s += (__expf(v.x - m) + __expf(v.y - m)) + (__expf(v.z - m) + __expf(v.w - m));
  }
  __device__ __forceinline__ void merge(float m2, float s2) {
    const float mm = fmaxf(m, m2);
    if (mm == -INFINITY) return;
    s = s * __expf(m - mm) + s2 * __expf(m2 - mm);
    m = mm;
  }
};
// block-wide log-sum-exp from per-thread OnlineLse (blockDim <= 1024); `sh` needs 64 floats.  All threads get the result.
__device__ __forceinline__ float block_lse(OnlineLse o, float* sh) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) o.merge(__shfl_xor_sync(0xffffffffu, o.m, d), __shfl_xor_sync(0xffffffffu, o.s, d));
  __syncthreads();
  if (lane == 0) { sh[w] = o.m; sh[32 + w] = o.s; }
  __syncthreads();
  OnlineLse r; r.m = lane < nw ? sh[lane] : -INFINITY; r.s = lane < nw ? sh[32 + lane] : 0.f;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) r.merge(__shfl_xor_sync(0xffffffffu, r.m, d), __shfl_xor_sync(0xffffffffu, r.s, d));
  return r.m + logf(r.s);
}

__global__ void __launch_bounds__(128)
row_argmax_kernel(const float* __restrict__ logits, int64_t ld, int V, int64_t* __restrict__ ids, int64_t ld_ids) {
  pdl_prologue();
  __shared__ float sv[4]; __shared__ int si[4];
  const float* x = logits + (int64_t)blockIdx.x * ld;
  ArgMax a; a.v = -INFINITY; a.i = 0x7fffffff;
  row_foreach(x, V, [&](float v, int c) { if (v > a.v || (v == a.v && c < a.i)) { a.v = v; a.i = c; } });
  a = warp_argmax(a);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (lane == 0) { sv[w] = a.v; si[w] = a.i; }
  __syncthreads();
  if (w == 0) {
    ArgMax b; b.v = lane < nw ? sv[lane] : -INFINITY; b.i = lane < nw ? si[lane] : 0x7fffffff;
    b = warp_argmax(b);
    if (lane == 0) ids[(int64_t)blockIdx.x * ld_ids] = b.i;
  }
}

__global__ void __launch_bounds__(256)
log_softmax_kernel(const float* __restrict__ logits, int64_t ld, int V, float* __restrict__ out, int64_t ldo) {
  pdl_prologue();
  __shared__ float red[64];
  const float* x = logits + (int64_t)blockIdx.x * ld;
  float* y = out + (int64_t)blockIdx.x * ldo;
  OnlineLse o; o.init();
  row_foreach(x, V, [&](float v, int) { o.add(v); }, [&](const float4 v, int) { o.add4(v); });
  const float lse = block_lse(o, red);
  if (((reinterpret_cast<uintptr_t>(x) ^ reinterpret_cast<uintptr_t>(y)) & 15u) == 0) {
    row_foreach(x, V, [&](float v, int c) { y[c] = v - lse; },
                [&](const float4 v, int c) { *reinterpret_cast<float4*>(y + c) = make_float4(v.x - lse, v.y - lse, v.z - lse, v.w - lse); });
  } else {
    row_foreach(x, V, [&](float v, int c) { y[c] = v - lse; });
  }
}

// masked CE: one CTA per (b,t) row.  loss_sum += -log p[target] for t < len[b];
// dlogits = (softmax - onehot) * inv_count for counted rows, 0 for padded rows.
__global__ void __launch_bounds__(256)
ce_masked_kernel(const float* __restrict__ logits, const int64_t* __restrict__ targets, const int32_t* __restrict__ lens,
                 int L, int V, float* __restrict__ loss_sum, float* __restrict__ dlogits, float inv_count_host,
                 const float* __restrict__ inv_count_dev, float* __restrict__ row_loss) {
  pdl_prologue();
  __shared__ float red[64];
  const float inv_count = inv_count_dev ? *inv_count_dev : inv_count_host;
  const int row = blockIdx.x, b = row / L, t = row % L;
  const float* x = logits + (int64_t)row * V;
  if (t >= lens[b]) {
    if (dlogits) for (int c = threadIdx.x; c < V; c += blockDim.x) dlogits[(int64_t)row * V + c] = 0.f;
    if (row_loss && threadIdx.x == 0) row_loss[row] = 0.f;
    return;
  }
  OnlineLse o; o.init();
  row_foreach(x, V, [&](float v, int) { o.add(v); }, [&](const float4 v, int) { o.add4(v); });
  const float lse = block_lse(o, red);
  const int tgt = (int)targets[row];
  if (threadIdx.x == 0) {
    const float l = (lse - x[tgt]) * inv_count;
    if (row_loss) row_loss[row] = l;               // summed in row order by ce_sum_rows_kernel: deterministic
    else atomicAdd(loss_sum, l);
  }
  if (dlogits) {
    float* __restrict__ d = dlogits + (int64_t)row * V;          // same pitch as x: same alignment pattern
    const bool same = ((reinterpret_cast<uintptr_t>(d) ^ reinterpret_cast<uintptr_t>(x)) & 15u) == 0;
    if (same) {
      const int tid = threadIdx.x, nt = blockDim.x;
      int head = (int)(((16u - (uint32_t)(reinterpret_cast<uintptr_t>(x) & 15u)) & 15u) >> 2);
      if (head > V) head = V;
      if (tid < head) d[tid] = (expf(x[tid] - lse) - (tid == tgt ? 1.f : 0.f)) * inv_count;
      const float4* __restrict__ x4 = reinterpret_cast<const float4*>(x + head);
      float4* __restrict__ d4 = reinterpret_cast<float4*>(d + head);
      const int n4 = (V - head) >> 2;
      for (int j = tid; j < n4; j += nt) {
        const float4 v = x4[j];
        const int c = head + 4 * j;
        float4 g = make_float4(expf(v.x - lse), expf(v.y - lse), expf(v.z - lse), expf(v.w - lse));
        if ((unsigned)(tgt - c) < 4u) { if (tgt == c) g.x -= 1.f; else if (tgt == c + 1) g.y -= 1.f; else if (tgt == c + 2) g.z -= 1.f; else g.w -= 1.f; }
        d4[j] = make_float4(g.x * inv_count, g.y * inv_count, g.z * inv_count, g.w * inv_count);
      }
      const int t0 = head + 4 * n4;
      if (t0 + tid < V) d[t0 + tid] = (expf(x[t0 + tid] - lse) - (t0 + tid == tgt ? 1.f : 0.f)) * inv_count;
    } else {
      for (int c = threadIdx.x; c < V; c += blockDim.x) d[c] = (expf(x[c] - lse) - (c == tgt ? 1.f : 0.f)) * inv_count;
    }
  }
}

// loss_sum += sum(row_loss[0..n)) with a fixed summation tree (one block): run-to-run deterministic
__global__ void __launch_bounds__(1024)
ce_sum_rows_kernel(const float* __restrict__ row_loss, int n, float* __restrict__ loss_sum) {
  pdl_prologue();
  __shared__ float red[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += row_loss[i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) *loss_sum += s;
}

// ------------------------------------------------------------------------------------------- beam search
// Per row: lse over V, then top-k of (x - lse) by value desc / index asc.  Rows whose last token is <end>
// emit {<end>: 0, then the k-1 lowest other indices at -inf} exactly as topk over the forced one-hot row
// would on ties... (torch.topk's order among equal -inf entries is unspecified; only the first slot is used
// downstream with a finite score).
__device__ __forceinline__ bool kv_before(float v1, int i1, float v2, int i2) {
  return v1 > v2 || (v1 == v2 && i1 < i2);
}

template <int KT>
__device__ __forceinline__ void insert_topk(float (&tv)[KT], int (&ti)[KT], float v, int i) {
  if (!kv_before(v, i, tv[KT - 1], ti[KT - 1])) return;
  // fully unrolled insertion (static register indexing): walk up from the tail, shifting worse entries down
#pragma unroll
  for (int j = KT - 1; j >= 0; --j) {
    if (kv_before(v, i, tv[j], ti[j])) {
      if (j + 1 < KT) { tv[j + 1] = tv[j]; ti[j + 1] = ti[j]; }
      tv[j] = v; ti[j] = i;
    }
  }
}

// Two walks over the row (the second one hits L2): (1) per-thread maxima (+ the online log-sum-exp); the k-th largest of the
// 128 thread maxima is a LOWER bound T of the row's k-th largest value (they are k distinct elements); (2) every element >= T
// (a handful) is appended to a shared candidate list, from which one warp picks the top k by (value desc, index asc).
// Per-thread sorted insertion - the old scheme - made every warp run the ~60-instruction insertion path on almost every
// element (some lane always had a new top-5 entry among its first ~80 values): 33 us for 27 MB.  A degenerate row (more than
// TOPK_CAP candidates, e.g. thousands of equal logits) falls back to that scheme.
constexpr int TOPK_CAP = 512;

template <int KT>
__global__ void __launch_bounds__(128)
beam_topk_kernel(const float* __restrict__ logits, int64_t ld, int V, const int64_t* __restrict__ last, int end_index, int k,
                 float* __restrict__ top_lp, int64_t* __restrict__ top_id, int normalize) {
  pdl_prologue();
  __shared__ float red[64];
  __shared__ float tmax[128];
  __shared__ float cand_v[TOPK_CAP];
  __shared__ int cand_i[TOPK_CAP];
  __shared__ float cv[4 * MAXK];
  __shared__ int ci[4 * MAXK];
  __shared__ int ncand;
  __shared__ float thr;
  const int row = blockIdx.x;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (last && last[row] == end_index) {
    // forced <end>: log-prob 0 at end_index, -inf elsewhere
    if (threadIdx.x < k) {
      int idx;
      if (threadIdx.x == 0) idx = end_index;
      else { idx = threadIdx.x - 1; if (idx >= end_index) idx += 1; }
      top_lp[(int64_t)row * k + threadIdx.x] = threadIdx.x == 0 ? 0.f : -INFINITY;
      top_id[(int64_t)row * k + threadIdx.x] = idx;
    }
    return;
  }
  const float* x = logits + (int64_t)row * ld;
  // ---- walk 1: thread maxima, log-sum-exp
  OnlineLse o; o.init();
  float mx = -INFINITY;
  if (normalize) row_foreach(x, V, [&](float v, int) { o.add(v); }, [&](const float4 v, int) { o.add4(v); });
  else row_foreach(x, V, [&](float v, int) { mx = fmaxf(mx, v); },
                   [&](const float4 v, int) { mx = fmaxf(mx, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w))); });
  if (normalize) mx = o.m;
  tmax[threadIdx.x] = mx;
  if (threadIdx.x == 0) ncand = 0;
  const float lse = normalize ? block_lse(o, red) : 0.f;     // (contains the __syncthreads that publish tmax / ncand)
  if (!normalize) __syncthreads();
  if (w == 0) {
    // k-th largest of the 128 thread maxima (duplicates count separately): k rounds of warp arg-max with removal
    float a[4] = {tmax[lane], tmax[lane + 32], tmax[lane + 64], tmax[lane + 96]};
    float t = -INFINITY;
    const int rounds = k < 128 ? k : 128;
    for (int j = 0; j < rounds; ++j) {
      int bu = 0;
      float bv = a[0];
#pragma unroll
      for (int u = 1; u < 4; ++u) if (a[u] > bv) { bv = a[u]; bu = u; }
      ArgMax c; c.v = bv; c.i = lane * 4 + bu;
      const ArgMax best = warp_argmax(c);
      t = best.v;
      if ((best.i >> 2) == lane) {
#pragma unroll
        for (int u = 0; u < 4; ++u) if (u == (best.i & 3)) a[u] = -INFINITY;
      }
    }
    if (lane == 0) thr = (V >= 128 * 1 && k <= 128) ? t : -INFINITY;
  }
  __syncthreads();
  const float T = thr;
  // ---- walk 2: candidates >= T
  auto push = [&](float v, int c) {
    if (v >= T) {
      const int slot = atomicAdd(&ncand, 1);
      if (slot < TOPK_CAP) { cand_v[slot] = v; cand_i[slot] = c; }
    }
  };
  row_foreach(x, V, push);
  __syncthreads();
  const int n = ncand;
  if (n <= TOPK_CAP) {
    if (w == 0) {
      for (int j = 0; j < k; ++j) {
        ArgMax a; a.v = -INFINITY; a.i = 0x7fffffff;
        int slot = -1;
        for (int e = lane; e < n; e += 32) {
          const float v = cand_v[e]; const int i = cand_i[e];
          if (i != 0x7fffffff && kv_before(v, i, a.v, a.i)) { a.v = v; a.i = i; slot = e; }
        }
        const ArgMax best = warp_argmax(a);
        if (slot >= 0 && a.i == best.i) cand_i[slot] = 0x7fffffff;     // (column indices are unique: exactly one lane removes)
        __syncwarp();
        if (lane == 0) {
          top_lp[(int64_t)row * k + j] = best.v - lse;
          top_id[(int64_t)row * k + j] = best.i;
        }
      }
    }
    return;
  }
  // ---- degenerate row: per-thread sorted lists + two-level merge
  float tv[KT]; int ti[KT];
#pragma unroll
  for (int j = 0; j < KT; ++j) { tv[j] = -INFINITY; ti[j] = 0x7fffffff; }
  row_foreach(x, V, [&](float v, int c) { insert_topk<KT>(tv, ti, v, c); });
  int head = 0;
  for (int j = 0; j < k; ++j) {
    ArgMax a; a.v = -INFINITY; a.i = 0x7fffffff;
#pragma unroll
    for (int u = 0; u < KT; ++u) if (u == head && u < k) { a.v = tv[u]; a.i = ti[u]; }
    const ArgMax best = warp_argmax(a);
    if (a.i == best.i && a.v == best.v && best.i != 0x7fffffff) ++head;
    if (lane == 0) { cv[w * MAXK + j] = best.v; ci[w * MAXK + j] = best.i; }
  }
  __syncthreads();
  if (w == 0) {
    const int nwarps = blockDim.x >> 5;
    int hp = 0;                       // lane l (< nwarps) walks warp l's sorted candidate list
    for (int j = 0; j < k; ++j) {
      ArgMax a; a.v = -INFINITY; a.i = 0x7fffffff;
      if (lane < nwarps && hp < k) { a.v = cv[lane * MAXK + hp]; a.i = ci[lane * MAXK + hp]; }
      const ArgMax best = warp_argmax(a);
      if (lane < nwarps && hp < k && a.i == best.i && a.v == best.v && best.i != 0x7fffffff) ++hp;
      if (lane == 0) {
        top_lp[(int64_t)row * k + j] = best.v - lse;
        top_id[(int64_t)row * k + j] = best.i;
      }
    }
  }
}

// one warp per batch element: top `beam` of beam*k summed candidates (value desc, index asc)
__global__ void __launch_bounds__(32)
beam_merge_kernel(const float* __restrict__ top_lp, const int64_t* __restrict__ top_id, const float* __restrict__ last_lp,
                  int beam, int k, float* __restrict__ new_lp, int64_t* __restrict__ new_cls, int64_t* __restrict__ backptr,
                  int32_t* __restrict__ all_end, int end_index) {
  pdl_prologue();
  const int b = blockIdx.x, lane = threadIdx.x;
  const int n = beam * k;
  // each lane holds up to 8 candidates (n <= 256)
  float v[8]; int idx[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const int c = lane + u * 32;
    if (c < n) { v[u] = top_lp[(int64_t)b * n + c] + last_lp[(int64_t)b * beam + c / k]; idx[u] = c; }
    else { v[u] = -INFINITY; idx[u] = 0x7fffffff; }
  }
  bool not_end = false;
  for (int j = 0; j < beam; ++j) {
    ArgMax a; a.v = -INFINITY; a.i = 0x7fffffff;
#pragma unroll
    for (int u = 0; u < 8; ++u) { ArgMax c; c.v = v[u]; c.i = idx[u]; if (idx[u] != 0x7fffffff) a = better(a, c); }
    // -inf candidates must still be selectable (index order) once finite ones are exhausted
    const ArgMax best = warp_argmax(a);
#pragma unroll
    for (int u = 0; u < 8; ++u) if (idx[u] == best.i) { idx[u] = 0x7fffffff; v[u] = -INFINITY; }
    if (lane == 0) {
      const int64_t cls = top_id[(int64_t)b * n + best.i];
      new_lp[(int64_t)b * beam + j] = best.v;
      new_cls[(int64_t)b * beam + j] = cls;
      backptr[(int64_t)b * beam + j] = best.i / k;
      if (cls != end_index) not_end = true;
    }
  }
  if (lane == 0 && not_end && all_end) atomicAnd(all_end, 0);
}

__global__ void beam_gather_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, const int64_t* __restrict__ backptr,
                                   int beam, int W) {
  pdl_prologue();
  const int row = blockIdx.x;                 // b*beam + j ; W = row length in 32-bit words
  const int b = row / beam;
  const int64_t srow = (int64_t)b * beam + backptr[row];
  for (int c = threadIdx.x; c < W; c += blockDim.x) dst[(int64_t)row * W + c] = src[srow * W + c];
}

// the same re-indexing for up to four state buffers in ONE launch (h / c rows of both LSTMs after a beam step): grid (rows, n),
// 16-byte copies when every row is 16-byte sized and aligned
struct BeamGatherMulti { const void* src[4]; void* dst[4]; int32_t row_bytes[4]; int32_t n; };
__global__ void beam_gather_multi_kernel(const BeamGatherMulti a, const int64_t* __restrict__ backptr, int beam) {
  pdl_prologue();
  const int row = blockIdx.x, k = blockIdx.y;
  const int b = row / beam;
  const int64_t srow = (int64_t)b * beam + backptr[row];
  const int rb = a.row_bytes[k];
  const uint8_t* s = static_cast<const uint8_t*>(a.src[k]) + srow * rb;
  uint8_t* d = static_cast<uint8_t*>(a.dst[k]) + (int64_t)row * rb;
  if ((rb & 15) == 0 && ((reinterpret_cast<uintptr_t>(a.src[k]) | reinterpret_cast<uintptr_t>(a.dst[k])) & 15) == 0) {
    for (int c = threadIdx.x; c < (rb >> 4); c += blockDim.x) reinterpret_cast<uint4*>(d)[c] = reinterpret_cast<const uint4*>(s)[c];
  } else {
    for (int c = threadIdx.x; c < (rb >> 2); c += blockDim.x) reinterpret_cast<uint32_t*>(d)[c] = reinterpret_cast<const uint32_t*>(s)[c];
  }
}

__global__ void beam_backtrack_kernel(const int64_t* __restrict__ preds, const int64_t* __restrict__ backs, int S, int B, int beam,
                                      int64_t* __restrict__ out) {
  pdl_prologue();
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= B * beam) return;
  const int b = e / beam;
  int cur = e % beam;
  for (int t = S - 1; t >= 0; --t) {
    out[(int64_t)e * S + t] = preds[((int64_t)t * B + b) * beam + cur];
    if (t > 0) cur = (int)backs[((int64_t)(t - 1) * B + b) * beam + cur];
  }
}


// dKW[hd][r][j][:] (+)= sum_t dl[t][r][hd*P+j] q[t][r][:] ;  dVW[hd][r][j][:] (+)= sum_t alpha[t][r][hd*P+j] dco[t][r][hd*Hv + :]
// (the node gradients of the hoisted attention, accumulated over the T decode steps in ONE launch after the time loop;
// grid (rows, nh), 256 threads x 4 columns, all loads of a step independent)
__global__ void __launch_bounds__(256)
attn2_bwd_nodes_kernel(const float* __restrict__ q_all, int64_t ldq, int64_t q_ts, const float* __restrict__ dl_all, int64_t lddl, int64_t dl_ts,
                       const float* __restrict__ al_all, int64_t ldal, int64_t al_ts, const float* __restrict__ dco_all, int64_t lddco,
                       int64_t dco_ts, float* __restrict__ dKW, float* __restrict__ dVW, int T, int rows, int P, int Hk, int Hv, int accum) {
  pdl_prologue();
  const int r = blockIdx.x, hd = blockIdx.y, c = threadIdx.x * 4;
  const bool ak = c < Hk, av = c < Hv;
  float4 ka[APM], va[APM];
#pragma unroll
  for (int j = 0; j < APM; ++j) { ka[j] = make_float4(0.f, 0.f, 0.f, 0.f); va[j] = ka[j]; }
  for (int t = 0; t < T; ++t) {
    float4 q4 = make_float4(0.f, 0.f, 0.f, 0.f), d4 = q4;
    if (ak) q4 = *reinterpret_cast<const float4*>(q_all + (int64_t)t * q_ts + (int64_t)r * ldq + c);
    if (av) d4 = *reinterpret_cast<const float4*>(dco_all + (int64_t)t * dco_ts + (int64_t)r * lddco + hd * Hv + c);
    const float* dlp = dl_all + (int64_t)t * dl_ts + (int64_t)r * lddl + hd * P;
    const float* alp = al_all + (int64_t)t * al_ts + (int64_t)r * ldal + hd * P;
#pragma unroll
    for (int j = 0; j < APM; ++j) {
      if (j < P) {
        const float g = dlp[j], a = alp[j];
        ka[j].x = fmaf(g, q4.x, ka[j].x); ka[j].y = fmaf(g, q4.y, ka[j].y); ka[j].z = fmaf(g, q4.z, ka[j].z); ka[j].w = fmaf(g, q4.w, ka[j].w);
        va[j].x = fmaf(a, d4.x, va[j].x); va[j].y = fmaf(a, d4.y, va[j].y); va[j].z = fmaf(a, d4.z, va[j].z); va[j].w = fmaf(a, d4.w, va[j].w);
      }
    }
  }
  const int64_t nk = (((int64_t)hd * rows + r) * P) * Hk, nv = (((int64_t)hd * rows + r) * P) * Hv;
#pragma unroll
  for (int j = 0; j < APM; ++j) {
    if (j < P) {
      if (ak) {
        float4* o = reinterpret_cast<float4*>(dKW + nk + (int64_t)j * Hk + c);
        float4 v = ka[j];
        if (accum) { const float4 u = *o; v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w; }
        *o = v;
      }
      if (av) {
        float4* o = reinterpret_cast<float4*>(dVW + nv + (int64_t)j * Hv + c);
        float4 v = va[j];
        if (accum) { const float4 u = *o; v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w; }
        *o = v;
      }
    }
  }
}

}  // namespace dlsg

using namespace dlsg;

extern "C" {

int dlsg_node_attn_fwd(const dlsg_node_attn_fwd_t* p, void* stream) {
  DLSG_REQUIRE(p->P >= 1 && p->P <= MAXP, "node_attn: P=%d out of range (1..%d)", p->P, MAXP);
  DLSG_REQUIRE(p->rows > 0 && p->nh > 0 && p->rows_per_node >= 1, "node_attn: bad shape");
  const int es = p->ctx_dtype == DLSG_F32 ? 4 : 2;
  if (p->P <= APM && p->H <= 1024 && p->H % 4 == 0 && al16(p->Kp) && al16(p->Vp) && al16(p->qp) &&
      (reinterpret_cast<uintptr_t>(p->ctx) % (4 * es)) == 0 && p->ldctx % 4 == 0) {
    DLSG_LAUNCH(node_attn_fwd_fast, dim3(p->rows, p->nh), 256, 0, (cudaStream_t)stream, *p);
    return check_launch("node_attn_fwd_fast");
  }
  DLSG_LAUNCH(node_attn_fwd_kernel, dim3(p->rows, p->nh), 128, 0, (cudaStream_t)stream, *p);
  return check_launch("node_attn_fwd_kernel");
}
int dlsg_node_attn_bwd(const dlsg_node_attn_bwd_t* p, void* stream) {
  DLSG_REQUIRE(p->P >= 1 && p->P <= MAXP, "node_attn_bwd: P=%d out of range (1..%d)", p->P, MAXP);
  const int es = p->dqp_dtype == DLSG_F32 ? 4 : 2;
  if (p->P <= APM && p->H <= 1024 && p->H % 4 == 0 && al16(p->Kp) && al16(p->Vp) && al16(p->qp) && al16(p->dKp) && al16(p->dVp) &&
      al16(p->dctx) && p->lddctx % 4 == 0 && (reinterpret_cast<uintptr_t>(p->dqp) % (4 * es)) == 0) {
    DLSG_LAUNCH(node_attn_bwd_fast, dim3(p->rows, p->nh), 256, 0, (cudaStream_t)stream, *p);
    return check_launch("node_attn_bwd_fast");
  }
  DLSG_LAUNCH(node_attn_bwd_kernel, dim3(p->rows, p->nh), 128, 0, (cudaStream_t)stream, *p);
  return check_launch("node_attn_bwd_kernel");
}

int dlsg_attn2_supported(int32_t nh, int32_t P, int32_t Hk, int32_t Hv) {
  return (nh >= 1 && nh <= 2 && P >= 1 && P <= APM && Hk <= 1024 && Hv <= 1024 && Hk % 4 == 0 && Hv % 4 == 0) ? 1 : 0;
}
int dlsg_attn2_fwd(const dlsg_attn2_fwd_t* p, void* stream) {
  DLSG_REQUIRE(dlsg_attn2_supported(p->nh, p->P, p->Hk, p->Hv), "attn2_fwd: unsupported shape nh=%d P=%d Hk=%d Hv=%d", p->nh, p->P, p->Hk, p->Hv);
  DLSG_REQUIRE(al16(p->KW) && al16(p->VW) && al16(p->q) && al16(p->co) && p->ldq % 4 == 0 && p->ldco % 4 == 0, "attn2_fwd: unaligned operands");
  if (p->y) {
    const int es = p->y_dtype == DLSG_F32 ? 4 : 2;
    DLSG_REQUIRE((reinterpret_cast<uintptr_t>(p->y) % (4 * es)) == 0 && p->ldy % 4 == 0 && al16(p->gamma[0]) && al16(p->beta[0]) &&
                 (p->nh < 2 || (al16(p->gamma[1]) && al16(p->beta[1]))) && (p->offset % 4 == 0) && (p->offset_head_stride % 4 == 0),
                 "attn2_fwd: unaligned fused-LayerNorm operands");
  }
  if (p->rows <= 0) return 0;
  DLSG_LAUNCH(attn2_fwd_kernel, dim3(p->rows, p->nh), 256, 0, (cudaStream_t)stream, *p);
  return check_launch("attn2_fwd_kernel");
}
int dlsg_attn2_bwd(const dlsg_attn2_bwd_t* p, void* stream) {
  DLSG_REQUIRE(dlsg_attn2_supported(p->nh, p->P, p->Hk, p->Hv), "attn2_bwd: unsupported shape nh=%d P=%d Hk=%d Hv=%d", p->nh, p->P, p->Hk, p->Hv);
  DLSG_REQUIRE(al16(p->KW) && al16(p->VW) && al16(p->q) && al16(p->dq) && al16(p->dKW) && al16(p->dVW) &&
               p->ldq % 4 == 0 && p->lddq % 4 == 0, "attn2_bwd: unaligned operands");
  if (p->dy) {
    DLSG_REQUIRE(al16(p->dy) && al16(p->co) && al16(p->gamma[0]) && (p->nh < 2 || al16(p->gamma[1])) && al16(p->dgamma_rows) && al16(p->dbeta_rows) &&
                 p->lddy % 4 == 0 && p->ldco % 4 == 0 && p->ld_dparam % 4 == 0 && p->stats && (p->offset % 4 == 0) &&
                 (p->offset_head_stride % 4 == 0), "attn2_bwd: unaligned fused-LayerNorm operands");
  } else {
    DLSG_REQUIRE(al16(p->dco) && p->lddco % 4 == 0, "attn2_bwd: unaligned dco");
  }
  if (p->rows <= 0) return 0;
  DLSG_LAUNCH(attn2_bwd_kernel, dim3(p->rows, p->nh), 256, 0, (cudaStream_t)stream, *p);
  return check_launch("attn2_bwd_kernel");
}
int dlsg_row_argmax(const float* logits, int64_t ld, int32_t rows, int32_t V, int64_t* ids, int64_t ld_ids, void* stream) {
  if (rows <= 0) return 0;
  DLSG_LAUNCH(row_argmax_kernel, rows, 128, 0, (cudaStream_t)stream, logits, ld, V, ids, ld_ids);
  return check_launch("row_argmax_kernel");
}
int dlsg_log_softmax(const float* logits, int64_t ld, int32_t rows, int32_t V, float* out, int64_t ldo, void* stream) {
  if (rows <= 0) return 0;
  DLSG_LAUNCH(log_softmax_kernel, rows, 256, 0, (cudaStream_t)stream, logits, ld, V, out, ldo);
  return check_launch("log_softmax_kernel");
}
int dlsg_ce_masked(const float* logits, const int64_t* targets, const int32_t* lens, int32_t B, int32_t L, int32_t V,
                   float* loss_sum, float* dlogits, float inv_count, const float* inv_count_dev, float* row_loss, void* stream) {
  if (B * L <= 0) return 0;
  DLSG_LAUNCH(ce_masked_kernel, B * L, 256, 0, (cudaStream_t)stream, logits, targets, lens, L, V, loss_sum, dlogits, inv_count, inv_count_dev,
              row_loss);
  if (int rc = check_launch("ce_masked_kernel")) return rc;
  if (!row_loss) return 0;
  DLSG_LAUNCH(ce_sum_rows_kernel, 1, 1024, 0, (cudaStream_t)stream, (const float*)row_loss, B * L, loss_sum);
  return check_launch("ce_sum_rows_kernel");
}
int dlsg_beam_topk(const float* logits, int64_t ld, int32_t rows, int32_t V, const int64_t* last, int32_t end_index,
                   int32_t k, float* top_lp, int64_t* top_id, int32_t normalize, void* stream) {
  DLSG_REQUIRE(k >= 1 && k <= MAXK, "beam_topk: k=%d out of range (1..%d)", k, MAXK);
  DLSG_REQUIRE(k <= V, "beam_topk: Target vocab size (%d) too small relative to per_node_beam_size (%d)", V, k);
  if (rows <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (k <= 1) DLSG_LAUNCH(beam_topk_kernel<1>, rows, 128, 0, st, logits, ld, V, last, end_index, k, top_lp, top_id, normalize);
  else if (k <= 3) DLSG_LAUNCH(beam_topk_kernel<3>, rows, 128, 0, st, logits, ld, V, last, end_index, k, top_lp, top_id, normalize);
  else if (k <= 5) DLSG_LAUNCH(beam_topk_kernel<5>, rows, 128, 0, st, logits, ld, V, last, end_index, k, top_lp, top_id, normalize);
  else if (k <= 8) DLSG_LAUNCH(beam_topk_kernel<8>, rows, 128, 0, st, logits, ld, V, last, end_index, k, top_lp, top_id, normalize);
  else DLSG_LAUNCH(beam_topk_kernel<16>, rows, 128, 0, st, logits, ld, V, last, end_index, k, top_lp, top_id, normalize);
  return check_launch("beam_topk_kernel");
}
int dlsg_beam_merge(const float* top_lp, const int64_t* top_id, const float* last_lp, int32_t B, int32_t beam, int32_t k,
                    float* new_lp, int64_t* new_cls, int64_t* backptr, int32_t* all_end, int32_t end_index, void* stream) {
  DLSG_REQUIRE(beam * k <= 256 && beam >= 1 && k >= 1, "beam_merge: beam*k=%d > 256", beam * k);
  if (B <= 0) return 0;
  DLSG_LAUNCH(beam_merge_kernel, B, 32, 0, (cudaStream_t)stream, top_lp, top_id, last_lp, beam, k, new_lp, new_cls, backptr, all_end, end_index);
  return check_launch("beam_merge_kernel");
}
int dlsg_beam_gather(const void* src, void* dst, const int64_t* backptr, int32_t B, int32_t beam, int32_t row_bytes, void* stream) {
  if (B * beam <= 0) return 0;
  DLSG_REQUIRE(row_bytes % 4 == 0, "beam_gather: row_bytes must be a multiple of 4");
  DLSG_LAUNCH(beam_gather_kernel, B * beam, 256, 0, (cudaStream_t)stream, (const uint32_t*)src, (uint32_t*)dst, backptr, beam, row_bytes / 4);
  return check_launch("beam_gather_kernel");
}
int dlsg_beam_gather_multi(const void* const* src, void* const* dst, const int32_t* row_bytes, int32_t n, const int64_t* backptr,
                           int32_t B, int32_t beam, void* stream) {
  DLSG_REQUIRE(n >= 1 && n <= 4, "beam_gather_multi: 1..4 buffers per launch");
  if (B * beam <= 0) return 0;
  BeamGatherMulti a = {};
  for (int k = 0; k < n; ++k) {
    DLSG_REQUIRE(src[k] && dst[k] && row_bytes[k] > 0 && row_bytes[k] % 4 == 0, "beam_gather_multi: buffer %d: null or row_bytes not a multiple of 4", k);
    a.src[k] = src[k]; a.dst[k] = dst[k]; a.row_bytes[k] = row_bytes[k];
  }
  a.n = n;
  DLSG_LAUNCH(beam_gather_multi_kernel, dim3(B * beam, n), 256, 0, (cudaStream_t)stream, a, backptr, beam);
  return check_launch("beam_gather_multi_kernel");
}
int dlsg_beam_backtrack(const int64_t* preds, const int64_t* backs, int32_t S, int32_t B, int32_t beam, int64_t* out, void* stream) {
  if (B * beam <= 0 || S <= 0) return 0;
  DLSG_LAUNCH(beam_backtrack_kernel, (B * beam + 127) / 128, 128, 0, (cudaStream_t)stream, preds, backs, S, B, beam, out);
  return check_launch("beam_backtrack_kernel");
}


int dlsg_attn2_bwd_nodes(const float* q_all, int64_t ldq, int64_t q_step_stride, const float* dl_all, int64_t lddl, int64_t dl_step_stride,
                         const float* alpha_all, int64_t ldalpha, int64_t alpha_step_stride, const float* dco_all, int64_t lddco,
                         int64_t dco_step_stride, float* dKW, float* dVW, int32_t T, int32_t rows, int32_t nh, int32_t P, int32_t Hk,
                         int32_t Hv, int32_t accumulate, void* stream) {
  DLSG_REQUIRE(dlsg_attn2_supported(nh, P, Hk, Hv), "attn2_bwd_nodes: unsupported shape nh=%d P=%d Hk=%d Hv=%d", nh, P, Hk, Hv);
  DLSG_REQUIRE(q_all && dl_all && alpha_all && dco_all && dKW && dVW, "attn2_bwd_nodes: null operand");
  DLSG_REQUIRE(ldq % 4 == 0 && q_step_stride % 4 == 0 && lddco % 4 == 0 && dco_step_stride % 4 == 0 &&
               ((reinterpret_cast<uintptr_t>(q_all) | reinterpret_cast<uintptr_t>(dco_all) | reinterpret_cast<uintptr_t>(dKW) |
                 reinterpret_cast<uintptr_t>(dVW)) & 15) == 0, "attn2_bwd_nodes: rows must be 16-byte aligned");
  if (rows <= 0 || T <= 0) return 0;
  DLSG_LAUNCH(attn2_bwd_nodes_kernel, dim3(rows, nh), 256, 0, (cudaStream_t)stream, q_all, ldq, q_step_stride, dl_all, lddl, dl_step_stride,
              alpha_all, ldalpha, alpha_step_stride, dco_all, lddco, dco_step_stride, dKW, dVW, T, rows, P, Hk, Hv, accumulate);
  return check_launch("attn2_bwd_nodes_kernel");
}

}  // extern "C"
