// Per-step fusions of the decoder's LSTM stages (one CTA per batch row, 256 threads):
//   lstm_cell_norm_fwd : split-K partial sums + hoisted row bias -> gates -> c,h (dropout) -> LayerNorm(h) [-> tanh]
//                        [-> dropout]; writes h into the next step's operand rows and LN(h) into its consumer's row
//                        (layer.py:571-574 query LSTM + LN, layer.py:593-599 lang LSTM + LN + tanh)
//   norm_lstm_cell_bwd : LayerNorm backward + LSTM cell backward (+ running sum of the gate gradients)
// They replace 3 launches each (reduce / cell / norm, resp. norm_bwd / cell_bwd / axpby) in the 26-step loop.
#include "common.cuh"

namespace dlsg {

constexpr int FS_MAXE = 8;        // elements per thread: H <= 2048

__global__ void __launch_bounds__(256)
lstm_cell_norm_fwd_kernel(const dlsg_lstm_cell_norm_fwd_t q) {
  pdl_prologue();
  __shared__ float red[32];
  const dlsg_lstm_cell_fwd_t& p = q.cell;
  const int b = blockIdx.x, H = p.H, tid = threadIdx.x;
  float hv[FS_MAXE];
  float sum = 0.f;
  const float hkeep = p.drop_p > 0.f ? 1.f / (1.f - p.drop_p) : 1.f;
#pragma unroll
  for (int e = 0; e < FS_MAXE; ++e) {
    const int h = tid + e * 256;
    hv[e] = 0.f;
    if (h < H) {
      float g[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int64_t gi = (int64_t)b * 4 * H + (int64_t)k * H + h;
        float v = 0.f;
        for (int s = 0; s < p.nsplit; ++s) v += p.gates[gi + (int64_t)s * p.stride_split];
        if (p.row_bias) v += p.row_bias[(int64_t)b * p.ld_row_bias + (int64_t)k * H + h];
        if (p.bias) v += p.bias[k * H + h];
        g[k] = v;
      }
      const float ig = sigmoidf_(g[0]), fg = sigmoidf_(g[1]), gg = tanhf(g[2]), og = sigmoidf_(g[3]);
      const int64_t ei = (int64_t)b * H + h;
      const float c = fg * (p.c_prev ? p.c_prev[ei] : 0.f) + ig * gg;
      float hval = og * tanhf(c);
      const int64_t g0 = (int64_t)b * 4 * H + h;
      p.gates[g0] = ig; p.gates[g0 + H] = fg; p.gates[g0 + 2 * (int64_t)H] = gg; p.gates[g0 + 3 * (int64_t)H] = og;
      p.c_out[ei] = c;
      if (p.drop_p > 0.f) hval *= drop_scale(p.drop_p, p.seed, p.offset + (uint64_t)ei);
      if (p.h_out) p.h_out[ei] = hval;
      if (p.h2) st_from_float(p.h2, p.h2_dtype, (int64_t)b * p.ldh2 + h, hval);
      if (p.h3) st_from_float(p.h3, p.h3_dtype, (int64_t)b * p.ldh3 + h, hval);
      hv[e] = hval;
      sum += hval;
    }
  }
  (void)hkeep;
  const float mean = block_sum(sum, red) / (float)H;
  float sq = 0.f;
#pragma unroll
  for (int e = 0; e < FS_MAXE; ++e) {
    const int h = tid + e * 256;
    if (h < H) { const float d = hv[e] - mean; sq = fmaf(d, d, sq); }
  }
  const float rstd = rsqrtf(block_sum(sq, red) / (float)H + 1e-5f);
  if (q.stats && tid == 0) { q.stats[b * 2] = mean; q.stats[b * 2 + 1] = rstd; }
#pragma unroll
  for (int e = 0; e < FS_MAXE; ++e) {
    const int h = tid + e * 256;
    if (h < H) {
      float y = (hv[e] - mean) * rstd * q.gamma[h] + q.beta[h];
      if (q.post_tanh) y = tanhf(y);
      if (q.ydrop_p > 0.f) y *= drop_scale(q.ydrop_p, q.yseed, q.yoffset + (uint64_t)b * H + h);
      if (q.y) st_from_float(q.y, q.y_dtype, (int64_t)b * q.ldy + h, y);
      if (q.y2) st_from_float(q.y2, q.y2_dtype, (int64_t)b * q.ldy2 + h, y);
    }
  }
}

__global__ void __launch_bounds__(256)
norm_lstm_cell_bwd_kernel(const dlsg_norm_lstm_cell_bwd_t q) {
  pdl_prologue();
  __shared__ float red[32];
  const dlsg_lstm_cell_bwd_t& p = q.cell;
  const int b = blockIdx.x, H = p.H, tid = threadIdx.x;
  const float mean = q.stats[b * 2], rstd = q.stats[b * 2 + 1];
  float xh[FS_MAXE], dv[FS_MAXE];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int e = 0; e < FS_MAXE; ++e) {
    const int h = tid + e * 256;
    xh[e] = 0.f; dv[e] = 0.f;
    if (h < H) {
      const float x = q.x[(int64_t)b * q.ldx + h];
      float dy = q.dy[(int64_t)b * q.lddy + h];
      if (q.ydrop_p > 0.f) dy *= drop_scale(q.ydrop_p, q.yseed, q.yoffset + (uint64_t)b * H + h);
      const float g = q.gamma[h];
      const float xn = (x - mean) * rstd;
      if (q.post_tanh) { const float yt = tanhf(xn * g + q.beta[h]); dy *= (1.f - yt * yt); }
      atomicAdd(&q.dgamma[h], dy * xn);
      atomicAdd(&q.dbeta[h], dy);
      const float d = dy * g;
      xh[e] = xn; dv[e] = d;
      s1 += d; s2 = fmaf(d, xn, s2);
    }
  }
  s1 = block_sum(s1, red) / (float)H;
  s2 = block_sum(s2, red) / (float)H;
#pragma unroll
  for (int e = 0; e < FS_MAXE; ++e) {
    const int h = tid + e * 256;
    if (h < H) {
      // gradient wrt the (dropped) h: LayerNorm path + recurrent paths
      float dh = rstd * (dv[e] - s1 - xh[e] * s2);
      if (p.dh) dh += p.dh[(int64_t)b * p.lddh + h];
      if (p.dh2) dh += p.dh2[(int64_t)b * p.lddh2 + h];
      const int64_t ei = (int64_t)b * H + h;
      if (p.drop_p > 0.f) dh *= drop_scale(p.drop_p, p.seed, p.offset + (uint64_t)ei);
      const int64_t g0 = (int64_t)b * 4 * H + h;
      const float ig = p.acts[g0], fg = p.acts[g0 + H], gg = p.acts[g0 + 2 * (int64_t)H], og = p.acts[g0 + 3 * (int64_t)H];
      const float tc = tanhf(p.c_new[ei]);
      float dc = dh * og * (1.f - tc * tc);
      if (p.dc_next) dc += p.dc_next[ei];
      const float cp = p.c_prev ? p.c_prev[ei] : 0.f;
      float d[4];
      d[0] = dc * gg * ig * (1.f - ig);
      d[1] = dc * cp * fg * (1.f - fg);
      d[2] = dc * ig * (1.f - gg * gg);
      d[3] = dh * tc * og * (1.f - og);
      if (p.dc_prev) p.dc_prev[ei] = dc * fg;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int64_t col = (int64_t)k * H + h;
        if (p.dgates) p.dgates[(int64_t)b * 4 * H + col] = d[k];
        if (q.dgates_sum) q.dgates_sum[(int64_t)b * 4 * H + col] += d[k];
        if (p.dgates2) st_from_float(p.dgates2, p.dgates2_dtype, (int64_t)b * p.ld_dgates2 + col, d[k]);
        if (p.dgatesT) st_from_float(p.dgatesT, p.dgatesT_dtype, col * p.ld_dgatesT + b, d[k]);
      }
    }
  }
}

}  // namespace dlsg

using namespace dlsg;

extern "C" {

int dlsg_lstm_cell_norm_fwd(const dlsg_lstm_cell_norm_fwd_t* q, void* stream) {
  DLSG_REQUIRE(q->cell.B > 0 && q->cell.H > 0 && q->cell.H <= 256 * FS_MAXE && q->cell.nsplit >= 1, "lstm_cell_norm_fwd: bad shape (H <= %d)", 256 * FS_MAXE);
  DLSG_LAUNCH(lstm_cell_norm_fwd_kernel, q->cell.B, 256, 0, (cudaStream_t)stream, *q);
  return check_launch("lstm_cell_norm_fwd_kernel");
}

int dlsg_norm_lstm_cell_bwd(const dlsg_norm_lstm_cell_bwd_t* q, void* stream) {
  DLSG_REQUIRE(q->cell.B > 0 && q->cell.H > 0 && q->cell.H <= 256 * FS_MAXE, "norm_lstm_cell_bwd: bad shape (H <= %d)", 256 * FS_MAXE);
  DLSG_REQUIRE(q->dgamma && q->dbeta && q->stats, "norm_lstm_cell_bwd: dgamma/dbeta/stats required");
  DLSG_LAUNCH(norm_lstm_cell_bwd_kernel, q->cell.B, 256, 0, (cudaStream_t)stream, *q);
  return check_launch("norm_lstm_cell_bwd_kernel");
}

}  // extern "C"
