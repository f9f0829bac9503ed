// Region -> frame aggregation of the latent-semantic-graph encoder (EncoderVisualGraphTUN, models/layer.py:184-192) as
// streaming kernels ("graph aggregation", BASELINE north_star class (b); SURVEY 2.2 "flash-style over the 936 axis"):
//
//   O_r   = LayerNorm_obj(Y_r)                    Y = tanh(region projection), bf16, (T*R rows, H) per clip and encoder
//   S_tr  = F_t . O_r                             F = LayerNorm(tanh(frame projection)), (T, H)
//   A_tr  = softmax_r(S_tr / sqrt(Dr))            over ALL T*R regions of the clip (layer.py:188, dim=1)
//   agg_t = sum_r A_tr O_r
//
// The unfused path reads the 123 MB of Y / O per encoder four times forward (LayerNorm read + write, scores GEMM,
// aggregation GEMM) and six times backward.  Here:
//  * forward: one CTA per (clip, encoder) streams its T*R x H tile of Y ONCE through shared memory (cp.async.bulk row
//    copies, two stages, mbarrier completion) and keeps everything else on chip.  LayerNorm is folded algebraically, so both
//    products run on the RAW bf16 rows that sit in shared memory:
//       S_tr  = rstd_r (F'_t . Y_r - mu_r sum_h F'_th) + F_t . beta,                F' = bf16(F o gamma)
//       U_t   = sum_r (A_tr rstd_r) Y_r - sum_r A_tr rstd_r mu_r ;  agg_t = gamma o U_t + beta
//    The two products are mma.sync m16n8k16 bf16 (fp32 accumulate): 32 x 32 x 1024 scores per 32-row tile (K split over the
//    8 warps, partials reduced through shared memory) and a 32 x 1024 x 32 update of the aggregate (each warp owns 128
//    columns); online softmax over the tiles (running max / sum per frame, accumulator rescale); raw scores are written as
//    they are produced, the normalised weights by a short pass at the end (both (T, T*R) fp32: 0.1 MB per clip).
//  * backward pass 1 = the same kernel with scores_only (F := dA): dSm_tr = dA_t . O_r, one more read of Y.
//  * backward pass 2 (region_aggregate_bwd_kernel): one CTA per (256-column slice, clip, encoder), 3-stage ring, reads its
//    slice of Y once and writes d(pre-activation) in place in shared memory, then bulk-stores it.  Nothing in it reduces over
//    H: with W = [A ; dS] (dS = scale A o (dSm - dA.agg), the softmax backward),
//       dxhat_r = sum_t A_tr dA'_t + dS_tr F'_t              (32 x 64 x 256 mma per tile, dA' = dA o gamma, F' = F o gamma)
//       mean_h dxhat_r      = (sum_t A_tr sum_h dA'_t + dS_tr sum_h F'_t) / H
//       mean_h dxhat_r xhat = (sum_t A_tr (dSm_tr - dA_t.beta) + dS_tr (S_tr - F_t.beta)) / H
//    so the LayerNorm backward's two row statistics come from the small (T, T*R) matrices, and
//       V_t = sum_r dS_tr xhat_r  (second mma, accumulated over the tiles) gives dF = dA + gamma o V,
//       dgamma = sum_t dA_t o U_t + F_t o V_t,  dbeta = sum_t dA_t  (sum_r A_tr = 1, sum_r dS_tr = 0).
// HBM-bound by design: algorithmic bytes per clip and encoder = T*R*H*2 (forward), 3 x that (backward: two reads, one write).
#include "common.cuh"

namespace dlsg {

constexpr int RA_H = 1024;                 // node width these kernels are built for (region_projected_size)
constexpr int RA_ROWS = 32;                // region rows per tile
constexpr int RA_TPAD = 32;                // frames padded to two m16 tiles
constexpr int RA_TMAX = 26;                // shared memory is sized for T <= 26 frames (checked on the host)
constexpr int RA_PITCH = RA_H + 8;         // bf16 elements: row pitch 2064 B = 16 (mod 128) -> conflict-free ldmatrix
constexpr int RA_THREADS = 256;
constexpr int RA_STAGES = 2;
constexpr int RA_PP = RA_ROWS + 8;         // pitch of the bf16 weight tile P' (80 B rows: conflict-free ldmatrix)
constexpr int RA_SP = RA_ROWS + 1;         // pitch of the fp32 score partials

struct RaSmem {
  static constexpr int TILE_BYTES = RA_ROWS * RA_PITCH * 2;                  // 66048
  static constexpr int OFF_TILE = 0;
  static constexpr int OFF_F = OFF_TILE + RA_STAGES * TILE_BYTES;            // F' : T rows, same pitch
  static constexpr int F_BYTES_MAX = (RA_TMAX + 1) * RA_PITCH * 2;           // 55728: T rows of F' + one row of ones (row sums)
  static constexpr int OFF_SRED = OFF_F + F_BYTES_MAX;                       // [8 warps][32][33] fp32 score partials
  static constexpr int SRED_BYTES = 8 * RA_TPAD * RA_SP * 4;                 // 33792
  static constexpr int OFF_P = OFF_SRED + SRED_BYTES;                        // [32][40] bf16
  static constexpr int P_BYTES = RA_TPAD * RA_PP * 2;                        // 2560
  static constexpr int OFF_SMALL = OFF_P + P_BYTES;                          // floats: m[32] l[32] corr[32] cmu[32] sF[32] cF[32] gsq[8 warps][32]
  static constexpr int SMALL_BYTES = (6 + 8) * 32 * 4;
  static constexpr int OFF_BAR = OFF_SMALL + SMALL_BYTES;                    // 2 mbarriers
  static constexpr int TOTAL = OFF_BAR + 64;
};
static_assert(RaSmem::TOTAL <= 227 * 1024, "region_aggregate: shared memory budget");

__device__ __forceinline__ uint32_t ra_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ra_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void ra_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ra_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  while (true) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) break;
    if (++spins > (1u << 26)) __trap();          // fail loudly, never hang the box
  }
}
__device__ __forceinline__ void ra_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void ra_bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ra_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void ra_bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void ra_bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void ra_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void ra_ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ra_ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ra_mma(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float ra_bf16_round(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

// =========================================================================================== forward / scores pass
__global__ void __launch_bounds__(RA_THREADS, 1)
region_aggregate_fwd_kernel(const dlsg_region_agg_fwd_t p) {
  extern __shared__ __align__(128) uint8_t ra_smem[];
  __nv_bfloat16* tiles = reinterpret_cast<__nv_bfloat16*>(ra_smem + RaSmem::OFF_TILE);
  __nv_bfloat16* Fs = reinterpret_cast<__nv_bfloat16*>(ra_smem + RaSmem::OFF_F);
  float* sred = reinterpret_cast<float*>(ra_smem + RaSmem::OFF_SRED);
  __nv_bfloat16* Ps = reinterpret_cast<__nv_bfloat16*>(ra_smem + RaSmem::OFF_P);
  float* small = reinterpret_cast<float*>(ra_smem + RaSmem::OFF_SMALL);
  float *s_m = small, *s_l = small + 32, *s_corr = small + 64, *s_cmu = small + 96, *s_sF = small + 128, *s_cF = small + 160,
        *s_gsq = small + 192;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ra_smem + RaSmem::OFF_BAR);

  const int b = blockIdx.x, e = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int T = p.T, TR = p.TR, H = RA_H;
  const int ntiles = (TR + RA_ROWS - 1) / RA_ROWS;
  const bool scores_only = p.scores_only != 0;
  const __nv_bfloat16* Y = reinterpret_cast<const __nv_bfloat16*>(p.Y[e]) + (int64_t)b * TR * p.ldy;
  const float* F = p.F[e] + (int64_t)b * T * p.ldf;
  const float* gamma = p.gamma[e];
  const float* beta = p.beta[e];
  float* St = p.St[e] ? p.St[e] + (int64_t)b * T * TR : nullptr;
  float* stats = p.stats[e] ? p.stats[e] + (int64_t)b * TR * 2 : nullptr;

  if (tid == 0) {
    ra_mbar_init(ra_smem_u32(&bars[0]), 1);
    ra_mbar_init(ra_smem_u32(&bars[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  pdl_prologue();

  // producer: every warp issues the bulk copies of 4 region rows of a tile (UBLKCP is a uniform-datapath instruction: the
  // compiler serialises per-lane addresses, so 32 rows from one warp cost ~2000 cycles of that warp per tile - and the other
  // warps wait for it at the next barrier); rows past TR are zero-filled by the consumer
  auto issue = [&](int tile, int stage) {
    const int r0 = tile * RA_ROWS;
    const int nrows = min(RA_ROWS, TR - r0);
    const uint32_t bar = ra_smem_u32(&bars[stage]);
    if (tid == 0) ra_mbar_expect_tx(bar, (uint32_t)nrows * H * 2);
    const int row = warp * 4 + lane;
    if (lane < 4 && row < nrows)
      ra_bulk_g2s(ra_smem_u32(tiles + (size_t)stage * RA_ROWS * RA_PITCH + (size_t)row * RA_PITCH), Y + (int64_t)(r0 + row) * p.ldy,
                  (uint32_t)H * 2, bar);
  };
  issue(0, 0);
  if (ntiles > 1) issue(1, 1);

  // ---- F' = bf16(F o gamma) into shared memory; sF_t = sum_h F'_th (of the ROUNDED values), cF_t = F_t . beta
  for (int t = warp; t < RA_TPAD; t += 8) {
    float sf = 0.f, cf = 0.f;
    if (t < T) {
      for (int c = lane * 4; c < H; c += 128) {
        const float4 f = *reinterpret_cast<const float4*>(F + (int64_t)t * p.ldf + c);
        const float4 g = *reinterpret_cast<const float4*>(gamma + c);
        const float4 bt = *reinterpret_cast<const float4*>(beta + c);
        __nv_bfloat162 lo = __floats2bfloat162_rn(f.x * g.x, f.y * g.y), hi = __floats2bfloat162_rn(f.z * g.z, f.w * g.w);
        const float2 l2 = __bfloat1622float2(lo), h2 = __bfloat1622float2(hi);
        sf += (l2.x + l2.y) + (h2.x + h2.y);
        cf += (f.x * bt.x + f.y * bt.y) + (f.z * bt.z + f.w * bt.w);
        uint2 u; u.x = *reinterpret_cast<uint32_t*>(&lo); u.y = *reinterpret_cast<uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(Fs + (size_t)t * RA_PITCH + c) = u;
      }
    }
    if (t == T) {                              // row T = ones: its "score" is the row sum of Y (LayerNorm mean for free)
      for (int c = lane * 8; c < H; c += 256)
        *reinterpret_cast<uint4*>(Fs + (size_t)t * RA_PITCH + c) = make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);
    }
    sf = warp_sum(sf); cf = warp_sum(cf);
    if (lane == 0) {
      s_sF[t] = sf; s_cF[t] = cf; s_m[t] = -INFINITY; s_l[t] = 0.f; s_cmu[t] = 0.f; s_corr[t] = 1.f;
      if (t < T && p.tconst[e]) *reinterpret_cast<float4*>(p.tconst[e] + ((int64_t)b * T + t) * 4) = make_float4(sf, cf, 0.f, 0.f);
    }
  }
  __syncthreads();

  float acc[2][16][4];                       // aggregate accumulators: frames (2 m-tiles) x this warp's 128 columns (16 n-tiles)
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 16; ++j)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[i][j][c] = 0.f;

  const int g = lane >> 2, tq = lane & 3;
  const uint32_t Fs_u = ra_smem_u32(Fs), Ps_u = ra_smem_u32(Ps);
  // ldmatrix row addresses of the A operand F' (rows >= T read the row of ones)
  const int a_mat = lane >> 3, a_row = lane & 7;
  int fa_row[2];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) { const int r = mt * 16 + (a_mat & 1) * 8 + a_row; fa_row[mt] = r < T ? r : T; }
  const int a_koff = (a_mat >> 1) * 8;

  for (int tile = 0; tile < ntiles; ++tile) {
    const int stage = tile & 1;
    const int r0 = tile * RA_ROWS;
    const int nrows = min(RA_ROWS, TR - r0);
    __nv_bfloat16* tl = tiles + (size_t)stage * RA_ROWS * RA_PITCH;
    const uint32_t tl_u = ra_smem_u32(tl);
    ra_mbar_wait(ra_smem_u32(&bars[stage]), (uint32_t)((tile >> 1) & 1));
    if (nrows < RA_ROWS) {                     // last, ragged tile: zero the missing rows (stale bits could be NaN patterns)
      for (int i = tid; i < (RA_ROWS - nrows) * (H / 8); i += RA_THREADS) {
        const int rr = nrows + i / (H / 8), cc = (i % (H / 8)) * 8;
        *reinterpret_cast<uint4*>(tl + (size_t)rr * RA_PITCH + cc) = make_uint4(0u, 0u, 0u, 0u);
      }
      __syncthreads();
    }
    if (p.scores_only == 2) {                    // measurement only: stream the tiles, no arithmetic (tools/bench_region_agg.py)
      __syncthreads();
      if (tile + 2 < ntiles) { ra_fence_async(); issue(tile + 2, stage); }
      continue;
    }
    // ---- scores: S'[t][r] = sum_h F'[t][h] Y[r][h]; this warp covers h in [warp*128, warp*128+128).  Row T of F' is ones, so
    // S'[T][r] = sum_h Y[r][h]; the diagonal blocks of Y Y^T (4 more mma per k-step) give sum_h Y[r][h]^2: LayerNorm statistics
    // on the tensor cores, no separate pass over the tile.
    float sp[2][4][4], gd[2][2][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int c = 0; c < 4; ++c) sp[i][j][c] = 0.f;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int c = 0; c < 4; ++c) gd[i][j][c] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      const int k0 = warp * 128 + ks * 16;
      uint32_t a[2][4], ay[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        ra_ldsm_x4(Fs_u + (uint32_t)(fa_row[mt] * RA_PITCH + k0 + a_koff) * 2, a[mt][0], a[mt][1], a[mt][2], a[mt][3]);
        ra_ldsm_x4(tl_u + (uint32_t)((mt * 16 + (a_mat & 1) * 8 + a_row) * RA_PITCH + k0 + a_koff) * 2, ay[mt][0], ay[mt][1], ay[mt][2], ay[mt][3]);
      }
      uint32_t bq[4][2];
#pragma unroll
      for (int np = 0; np < 2; ++np) {           // two n-tiles (16 region rows) per ldmatrix.x4
        const int n = np * 16 + (lane >> 4) * 8 + (lane & 7);
        const int kh = ((lane >> 3) & 1) * 8;
        ra_ldsm_x4(tl_u + (uint32_t)(n * RA_PITCH + k0 + kh) * 2, bq[2 * np][0], bq[2 * np][1], bq[2 * np + 1][0], bq[2 * np + 1][1]);
      }
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) ra_mma(sp[mt][nt], a[mt][0], a[mt][1], a[mt][2], a[mt][3], bq[nt][0], bq[nt][1]);
#pragma unroll
        for (int sub = 0; sub < 2; ++sub)
          ra_mma(gd[mt][sub], ay[mt][0], ay[mt][1], ay[mt][2], ay[mt][3], bq[2 * mt + sub][0], bq[2 * mt + sub][1]);
      }
    }
    if ((g >> 1) == tq) {                        // this thread holds the diagonal entries of rows mt*16 + sub*8 + g
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int sub = 0; sub < 2; ++sub)
          s_gsq[warp * 32 + mt * 16 + sub * 8 + g] = (g & 1) ? gd[mt][sub][sub * 2 + 1] : gd[mt][sub][sub * 2];
    }
    {
      float* my = sred + (size_t)warp * RA_TPAD * RA_SP;
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const int t0 = mt * 16 + g, c0 = nt * 8 + 2 * tq;
          my[t0 * RA_SP + c0] = sp[mt][nt][0]; my[t0 * RA_SP + c0 + 1] = sp[mt][nt][1];
          my[(t0 + 8) * RA_SP + c0] = sp[mt][nt][2]; my[(t0 + 8) * RA_SP + c0 + 1] = sp[mt][nt][3];
        }
    }
    __syncthreads();
    // ---- online softmax: warp w handles frames w, w+8, w+16, w+24 (interleaved for ILP); lane = region row of the tile
    {
      float rsum = 0.f, rsq = 0.f;
#pragma unroll
      for (int w8 = 0; w8 < 8; ++w8) { rsum += sred[(size_t)w8 * RA_TPAD * RA_SP + T * RA_SP + lane]; rsq += s_gsq[w8 * 32 + lane]; }
      const float mu = rsum * (1.f / H);
      const float rstd = rsqrtf(fmaxf(rsq * (1.f / H) - mu * mu, 0.f) + 1e-5f);
      const bool rvalid = lane < nrows;
      if (warp == 0 && rvalid && stats) *reinterpret_cast<float2*>(stats + (int64_t)(r0 + lane) * 2) = make_float2(mu, rstd);
      float s[4], mx[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int t = warp + 8 * j;
        float raw = 0.f;
#pragma unroll
        for (int w8 = 0; w8 < 8; ++w8) raw += sred[(size_t)w8 * RA_TPAD * RA_SP + t * RA_SP + lane];
        const float sraw = rstd * (raw - mu * s_sF[t]) + s_cF[t];           // F_t . O_r  (unscaled, what the backward differentiates)
        const bool valid = rvalid && (t < T);
        if (valid && St) St[(int64_t)t * TR + r0 + lane] = sraw;
        s[j] = valid ? sraw * p.scale : -INFINITY;
        mx[j] = s[j];
      }
      if (!scores_only) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
          for (int j = 0; j < 4; ++j) mx[j] = fmaxf(mx[j], __shfl_xor_sync(0xffffffffu, mx[j], o));
        float pr[4], cm[4], corr[4], mnew[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int t = warp + 8 * j;
          const float m_old = s_m[t];
          mnew[j] = fmaxf(m_old, mx[j]);
          pr[j] = (s[j] > -INFINITY) ? __expf(s[j] - mnew[j]) : 0.f;
          corr[j] = (m_old > -INFINITY) ? __expf(m_old - mnew[j]) : 0.f;
          const __nv_bfloat16 wq = __float2bfloat16_rn(pr[j] * rstd);           // A'_tr rstd_r (unnormalised), as the mma sees it
          Ps[t * RA_PP + lane] = wq;
          cm[j] = __bfloat162float(wq) * mu;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
          for (int j = 0; j < 4; ++j) { pr[j] += __shfl_xor_sync(0xffffffffu, pr[j], o); cm[j] += __shfl_xor_sync(0xffffffffu, cm[j], o); }
        if (lane == 0) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int t = warp + 8 * j;
            s_m[t] = mnew[j]; s_corr[t] = corr[j];
            s_l[t] = s_l[t] * corr[j] + pr[j];
            s_cmu[t] = s_cmu[t] * corr[j] + cm[j];
          }
        }
      }
    }
    __syncthreads();
    // ---- aggregate: acc[t][h] = acc[t][h] * corr_t + sum_r P'[t][r] Y[r][h]   (this warp: h in [warp*128, +128))
    if (!scores_only) {
      float cr[2][2];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) { cr[mt][0] = s_corr[mt * 16 + g]; cr[mt][1] = s_corr[mt * 16 + g + 8]; }
      const bool resc = (cr[0][0] != 1.f) || (cr[0][1] != 1.f) || (cr[1][0] != 1.f) || (cr[1][1] != 1.f);
      if (__any_sync(0xffffffffu, resc)) {       // the running maxima settle after a few tiles: mostly skipped
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int nt = 0; nt < 16; ++nt) {
            acc[mt][nt][0] *= cr[mt][0]; acc[mt][nt][1] *= cr[mt][0];
            acc[mt][nt][2] *= cr[mt][1]; acc[mt][nt][3] *= cr[mt][1];
          }
      }
      uint32_t pa[2][2][4];                     // [k-step][m-tile]
#pragma unroll
      for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          const int row = mt * 16 + (a_mat & 1) * 8 + a_row;
          ra_ldsm_x4(Ps_u + (uint32_t)(row * RA_PP + ks * 16 + a_koff) * 2, pa[ks][mt][0], pa[ks][mt][1], pa[ks][mt][2], pa[ks][mt][3]);
        }
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
        for (int np = 0; np < 8; ++np) {         // two n-tiles (16 columns) per transposed ldmatrix.x4
          const int rr = ks * 16 + ((lane >> 3) & 1) * 8 + (lane & 7);
          const int hh = warp * 128 + np * 16 + (lane >> 4) * 8;
          uint32_t b0, b1, b2, b3;
          ra_ldsm_x4_t(tl_u + (uint32_t)(rr * RA_PITCH + hh) * 2, b0, b1, b2, b3);
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            ra_mma(acc[mt][2 * np], pa[ks][mt][0], pa[ks][mt][1], pa[ks][mt][2], pa[ks][mt][3], b0, b1);
            ra_mma(acc[mt][2 * np + 1], pa[ks][mt][0], pa[ks][mt][1], pa[ks][mt][2], pa[ks][mt][3], b2, b3);
          }
        }
      }
      __syncthreads();                                       // every warp is done with this stage (and with sred / Ps)
    }
    if (tile + 2 < ntiles) {
      ra_fence_async();                                      // generic-proxy reads above, async-proxy writes below
      issue(tile + 2, stage);
    }
  }
  if (scores_only) return;
  if (p.tconst[e] && tid < T)                    // softmax normalisers for the backward: A_tr = exp(scale S_tr - m_t) / l_t
    *reinterpret_cast<float2*>(p.tconst[e] + ((int64_t)b * T + tid) * 4 + 2) = make_float2(s_m[tid], 1.f / s_l[tid]);

  // ---- epilogue: U[t][h] = (acc[t][h] - cmu_t) / l_t ; agg = gamma o U + beta
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int t = mt * 16 + g + half * 8;
      if (t < T) {
        const float inv = 1.f / s_l[t], cm = s_cmu[t];
        float* out = p.agg[e] + ((int64_t)b * T + t) * p.ldagg;
        float* uout = p.U[e] ? p.U[e] + ((int64_t)b * T + t) * p.ldu : nullptr;
#pragma unroll
        for (int nt = 0; nt < 16; ++nt) {
          const int c = warp * 128 + nt * 8 + 2 * tq;
          const float2 gg = *reinterpret_cast<const float2*>(gamma + c);
          const float2 bb = *reinterpret_cast<const float2*>(beta + c);
          float2 u, o;
          u.x = (acc[mt][nt][half * 2] - cm) * inv;
          u.y = (acc[mt][nt][half * 2 + 1] - cm) * inv;
          o.x = gg.x * u.x + bb.x;
          o.y = gg.y * u.y + bb.y;
          *reinterpret_cast<float2*>(out + c) = o;
          if (uout) *reinterpret_cast<float2*>(uout + c) = u;
        }
      }
    }
  }
}

// =========================================================================================== backward pass 2
// (2a) region_aggregate_prep_kernel: everything that lives on the small (T, T*R) matrices, once per (clip, encoder):
//      A = exp(scale S - m) / l, c_t = sum_r A_tr dSm_tr, dS = scale A o (dSm - c)  (softmax backward), and per 32-row tile
//      the two mma A operands of (2b) as bf16 blocks [A ; dS]^T (32 x 72) and dS o rstd (32 x 40), the row scalars
//      (mu, rstd, mean_h dxhat, mean_h dxhat xhat) and cmuV_t = sum_r bf16(dS_tr rstd_r) mu_r - laid out so that (2b) fetches
//      each block with ONE bulk copy.
// (2b) region_aggregate_bwd_kernel: one CTA per (256-column slice, clip, encoder), 3-stage ring; per tile nothing but bulk
//      loads, two mma products, the fused LayerNorm / tanh backward in place, one bulk store.
constexpr int RB_HC = 256;                 // columns per CTA
constexpr int RB_PITCH = RB_HC + 8;        // 528 B rows = 16 (mod 128)
constexpr int RB_STAGES = 3;
constexpr int RB_WP = 64 + 8;              // pitch of W = [A ; dS]^T tile: (32 region rows) x (64 k), bf16
constexpr int RB_THREADS = 256;
constexpr int RB_KROWS = 2 * RA_TMAX + 1;  // rows of [dA' ; F' ; 0]
constexpr int RB_W_BYTES = RA_ROWS * RB_WP * 2;      // 4608
constexpr int RB_PT_BYTES = RA_TPAD * RA_PP * 2;     // 2560
constexpr int RB_RS_BYTES = RA_ROWS * 4 * 4;         // 512
constexpr int RB_BLK_BYTES = RB_W_BYTES + RB_PT_BYTES + RB_RS_BYTES;    // 7680 per tile: one bulk copy
constexpr int RB_PREP_SPLIT = 4;           // CTAs per (clip, encoder) in the prep kernel (each recomputes c_t: L2 reads)

struct RbSmem {
  static constexpr int TILE_BYTES = RA_ROWS * RB_PITCH * 2;                  // 16896
  static constexpr int STAGE_BYTES = TILE_BYTES + RB_BLK_BYTES;              // 24576
  static constexpr int OFF_STAGE = 0;
  static constexpr int OFF_B = OFF_STAGE + RB_STAGES * STAGE_BYTES;          // 73728
  static constexpr int B_BYTES = RB_KROWS * RB_PITCH * 2;                    // 27984
  static constexpr int OFF_BAR = OFF_B + B_BYTES;
  static constexpr int TOTAL = OFF_BAR + 64;
};
static_assert(2 * (RbSmem::TOTAL + 1024) <= 228 * 1024, "region_aggregate_bwd: two CTAs per SM");

__host__ __device__ inline int64_t rb_work_bytes_per_clip(int T, int TR) {
  const int ntiles = (TR + RA_ROWS - 1) / RA_ROWS;
  return (int64_t)ntiles * RB_BLK_BYTES + RB_PREP_SPLIT * 128;       // + per-split partial cmuV[32] floats
}

__global__ void __launch_bounds__(256)
region_aggregate_prep_kernel(const dlsg_region_agg_bwd_t p) {
  __shared__ __align__(16) uint8_t blk[RB_BLK_BYTES];
  __shared__ float red[8 * 32 * 2];
  __shared__ float s_c[32], s_m[32], s_il[32], s_cbA[32], s_cbF[32], s_sdA[32], s_sF[32], s_cmuV[32];
  __nv_bfloat16* Ws = reinterpret_cast<__nv_bfloat16*>(blk);
  __nv_bfloat16* PT = reinterpret_cast<__nv_bfloat16*>(blk + RB_W_BYTES);
  float* Rs = reinterpret_cast<float*>(blk + RB_W_BYTES + RB_PT_BYTES);
  const int split = blockIdx.x, b = blockIdx.y, e = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int T = p.T, TR = p.TR;
  const int ntiles = (TR + RA_ROWS - 1) / RA_ROWS;
  const int tiles_per = (ntiles + RB_PREP_SPLIT - 1) / RB_PREP_SPLIT;
  const int tile_lo = split * tiles_per, tile_hi = min(ntiles, tile_lo + tiles_per);
  const float* Stp = p.St[e] + (int64_t)b * T * TR;
  const float* dSp = p.dSm[e] + (int64_t)b * T * TR;
  const float* stats = p.stats[e] + (int64_t)b * TR * 2;
  uint8_t* work = reinterpret_cast<uint8_t*>(p.work[e]) + (int64_t)b * rb_work_bytes_per_clip(T, TR);
  const float invH = 1.f / (float)p.H;
  pdl_prologue();
  if (tid < 32) {
    const int t = tid;
    float m = 0.f, il = 0.f, cbA = 0.f, cbF = 0.f, sdA = 0.f, sF = 0.f;
    if (t < T) {
      const float4 a = *reinterpret_cast<const float4*>(p.tcA[e] + ((int64_t)b * T + t) * 4);
      const float4 f = *reinterpret_cast<const float4*>(p.tcF[e] + ((int64_t)b * T + t) * 4);
      sdA = a.x; cbA = a.y; sF = f.x; cbF = f.y; m = f.z; il = f.w;
    }
    s_m[t] = m; s_il[t] = il; s_cbA[t] = cbA; s_cbF[t] = cbF; s_sdA[t] = sdA; s_sF[t] = sF; s_cmuV[t] = 0.f; s_c[t] = 0.f;
  }
  for (int i = tid; i < RB_BLK_BYTES / 16; i += 256) reinterpret_cast<uint4*>(blk)[i] = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
  // ---- c_t = sum_r A_tr dSm_tr (one warp per frame)
  for (int t = warp; t < T; t += 8) {
    const float m = s_m[t], il = s_il[t];
    float acc = 0.f;
    const float* srow = Stp + (int64_t)t * TR;
    const float* drow = dSp + (int64_t)t * TR;
    int r = lane;
    for (; r + 96 < TR; r += 128) {              // eight loads in flight
      const float s0 = srow[r], s1 = srow[r + 32], s2 = srow[r + 64], s3 = srow[r + 96];
      const float d0 = drow[r], d1 = drow[r + 32], d2 = drow[r + 64], d3 = drow[r + 96];
      acc += __expf(s0 * p.scale - m) * d0 + __expf(s1 * p.scale - m) * d1 + __expf(s2 * p.scale - m) * d2 + __expf(s3 * p.scale - m) * d3;
    }
    for (; r < TR; r += 32) acc += __expf(srow[r] * p.scale - m) * drow[r];
    acc = warp_sum(acc) * il;
    if (lane == 0) s_c[t] = acc;
  }
  __syncthreads();
  for (int tile = tile_lo; tile < tile_hi; ++tile) {
    const int r0 = tile * RA_ROWS;
    const int nrows = min(RA_ROWS, TR - r0);
    const bool rvalid = lane < nrows;
    float mu = 0.f, rstd = 0.f;
    if (rvalid) { const float2 st = *reinterpret_cast<const float2*>(stats + (int64_t)(r0 + lane) * 2); mu = st.x; rstd = st.y; }
    float pa = 0.f, pb = 0.f;
    float sr[4], ds[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int t = warp + 8 * j;
      const bool v = rvalid && t < T;
      const int64_t idx = (int64_t)t * TR + r0 + lane;
      sr[j] = v ? Stp[idx] : 0.f; ds[j] = v ? dSp[idx] : 0.f;
    }
    float cm[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int t = warp + 8 * j;
      cm[j] = 0.f;
      if (t < T) {
        const float sm = rvalid ? __expf(sr[j] * p.scale - s_m[t]) * s_il[t] : 0.f;
        const float dst = p.scale * sm * (ds[j] - s_c[t]);
        const __nv_bfloat16 wa = __float2bfloat16_rn(sm), wd = __float2bfloat16_rn(dst), wp = __float2bfloat16_rn(dst * rstd);
        Ws[lane * RB_WP + t] = wa;
        Ws[lane * RB_WP + T + t] = wd;
        PT[t * RA_PP + lane] = wp;
        const float fa = __bfloat162float(wa), fd = __bfloat162float(wd);
        pa += fa * s_sdA[t] + fd * s_sF[t];
        pb += fa * (ds[j] - s_cbA[t]) + fd * (sr[j] - s_cbF[t]);
        cm[j] = __bfloat162float(wp) * mu;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int j = 0; j < 4; ++j) cm[j] += __shfl_xor_sync(0xffffffffu, cm[j], o);
    if (lane == 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j) if (warp + 8 * j < T) s_cmuV[warp + 8 * j] += cm[j];
    }
    red[(warp * 32 + lane) * 2] = pa; red[(warp * 32 + lane) * 2 + 1] = pb;
    __syncthreads();
    if (warp == 0) {
      float ar = 0.f, br = 0.f;
#pragma unroll
      for (int w8 = 0; w8 < 8; ++w8) { ar += red[(w8 * 32 + lane) * 2]; br += red[(w8 * 32 + lane) * 2 + 1]; }
      // row scalars in the form the epilogue consumes: x = y*rstd + nmr ; dpre = (dx*rstd - ars - x*brs) (1 - y^2)
      *reinterpret_cast<float4*>(Rs + lane * 4) = make_float4(rstd, -mu * rstd, ar * invH * rstd, br * invH * rstd);
    }
    __syncthreads();
    uint4* dst = reinterpret_cast<uint4*>(work + (int64_t)tile * RB_BLK_BYTES);
    for (int i = tid; i < RB_BLK_BYTES / 16; i += 256) dst[i] = reinterpret_cast<const uint4*>(blk)[i];
    __syncthreads();
  }
  if (tid < 32) reinterpret_cast<float*>(work + (int64_t)ntiles * RB_BLK_BYTES)[split * 32 + tid] = s_cmuV[tid];
}

__global__ void __launch_bounds__(RB_THREADS, 2)
region_aggregate_bwd_kernel(const dlsg_region_agg_bwd_t p) {
  extern __shared__ __align__(128) uint8_t ra_smem[];
  __nv_bfloat16* Bs = reinterpret_cast<__nv_bfloat16*>(ra_smem + RbSmem::OFF_B);
  uint64_t* bars = reinterpret_cast<uint64_t*>(ra_smem + RbSmem::OFF_BAR);

  const int cs = blockIdx.x, b = blockIdx.y, e = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int T = p.T, TR = p.TR;
  const int c0 = cs * RB_HC;
  const int ntiles = (TR + RA_ROWS - 1) / RA_ROWS;
  const __nv_bfloat16* Y = reinterpret_cast<const __nv_bfloat16*>(p.Y[e]) + (int64_t)b * TR * p.ldy + c0;
  __nv_bfloat16* dpre = reinterpret_cast<__nv_bfloat16*>(p.dpre[e]) + (int64_t)b * TR * p.ldd + c0;
  const float* Fp = p.F[e] + (int64_t)b * T * p.ldf + c0;
  const float* dAp = p.dA[e] + (int64_t)b * T * p.ldda + c0;
  const float* Up = p.U[e] + (int64_t)b * T * p.ldu + c0;
  const float* gamma = p.gamma[e] + c0;
  const uint8_t* work = reinterpret_cast<const uint8_t*>(p.work[e]) + (int64_t)b * rb_work_bytes_per_clip(T, TR);

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < RB_STAGES; ++s) ra_mbar_init(ra_smem_u32(&bars[s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  pdl_prologue();

  auto issue = [&](int tile, int stage) {     // lanes 0-3 of every warp: rows 4*warp + lane; thread 0 also fetches the prepared block
    const int r0 = tile * RA_ROWS;
    const int nrows = min(RA_ROWS, TR - r0);
    const uint32_t bar = ra_smem_u32(&bars[stage]);
    const uint32_t st_u = ra_smem_u32(ra_smem + RbSmem::OFF_STAGE + (size_t)stage * RbSmem::STAGE_BYTES);
    if (tid == 0) {
      ra_mbar_expect_tx(bar, (uint32_t)nrows * RB_HC * 2 + RB_BLK_BYTES);
      ra_bulk_g2s(st_u + RbSmem::TILE_BYTES, work + (int64_t)tile * RB_BLK_BYTES, RB_BLK_BYTES, bar);
    }
    const int row = warp * 4 + lane;
    if (lane < 4 && row < nrows) ra_bulk_g2s(st_u + (uint32_t)row * RB_PITCH * 2, Y + (int64_t)(r0 + row) * p.ldy, (uint32_t)RB_HC * 2, bar);
  };
  issue(0, 0);
  if (ntiles > 1) issue(1, 1);

  // ---- [dA o gamma ; F o gamma ; 0] as bf16, this CTA's 256 columns (lane = 8 columns)
  for (int row = warp; row < 2 * T + 1; row += 8) {
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    if (row < 2 * T) {
      const float* src = row < T ? dAp + (int64_t)row * p.ldda : Fp + (int64_t)(row - T) * p.ldf;
      const float4 x0 = *reinterpret_cast<const float4*>(src + lane * 8), x1 = *reinterpret_cast<const float4*>(src + lane * 8 + 4);
      const float4 g0 = *reinterpret_cast<const float4*>(gamma + lane * 8), g1 = *reinterpret_cast<const float4*>(gamma + lane * 8 + 4);
      __nv_bfloat162 q0 = __floats2bfloat162_rn(x0.x * g0.x, x0.y * g0.y), q1 = __floats2bfloat162_rn(x0.z * g0.z, x0.w * g0.w);
      __nv_bfloat162 q2 = __floats2bfloat162_rn(x1.x * g1.x, x1.y * g1.y), q3 = __floats2bfloat162_rn(x1.z * g1.z, x1.w * g1.w);
      o.x = *reinterpret_cast<uint32_t*>(&q0); o.y = *reinterpret_cast<uint32_t*>(&q1);
      o.z = *reinterpret_cast<uint32_t*>(&q2); o.w = *reinterpret_cast<uint32_t*>(&q3);
    }
    *reinterpret_cast<uint4*>(Bs + (size_t)row * RB_PITCH + lane * 8) = o;
  }
  __syncthreads();

  float accV[2][4][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int c = 0; c < 4; ++c) accV[i][j][c] = 0.f;
  float csum[4][2];
#pragma unroll
  for (int j = 0; j < 4; ++j) { csum[j][0] = 0.f; csum[j][1] = 0.f; }

  const int g = lane >> 2, tq = lane & 3;
  const int a_mat = lane >> 3, a_row = lane & 7;
  const int a_koff = (a_mat >> 1) * 8;
  const uint32_t Bs_u = ra_smem_u32(Bs);

  for (int tile = 0; tile < ntiles; ++tile) {
    const int stage = tile % RB_STAGES;
    const int r0 = tile * RA_ROWS;
    const int nrows = min(RA_ROWS, TR - r0);
    uint8_t* st = ra_smem + RbSmem::OFF_STAGE + (size_t)stage * RbSmem::STAGE_BYTES;
    __nv_bfloat16* tl = reinterpret_cast<__nv_bfloat16*>(st);
    const float* Rs = reinterpret_cast<const float*>(st + RbSmem::TILE_BYTES + RB_W_BYTES + RB_PT_BYTES);
    const uint32_t tl_u = ra_smem_u32(st), Ws_u = tl_u + RbSmem::TILE_BYTES, PT_u = Ws_u + RB_W_BYTES;
    ra_mbar_wait(ra_smem_u32(&bars[stage]), (uint32_t)((tile / RB_STAGES) & 1));
    if (nrows < RA_ROWS) {                     // ragged last tile: zero the missing rows
      for (int i = tid; i < (RA_ROWS - nrows) * (RB_HC / 8); i += RB_THREADS) {
        const int rr = nrows + i / (RB_HC / 8), cc = (i % (RB_HC / 8)) * 8;
        *reinterpret_cast<uint4*>(tl + (size_t)rr * RB_PITCH + cc) = make_uint4(0u, 0u, 0u, 0u);
      }
      __syncthreads();
    }
    // ---- V[t][h] += sum_r (dS_tr rstd_r) Y[r][h]     (this warp: 32 columns)
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      uint32_t a[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const int row = mt * 16 + (a_mat & 1) * 8 + a_row;
        ra_ldsm_x4(PT_u + (uint32_t)(row * RA_PP + ks * 16 + a_koff) * 2, a[mt][0], a[mt][1], a[mt][2], a[mt][3]);
      }
#pragma unroll
      for (int np = 0; np < 2; ++np) {
        const int rr = ks * 16 + ((lane >> 3) & 1) * 8 + (lane & 7);
        const int hh = warp * 32 + np * 16 + (lane >> 4) * 8;
        uint32_t b0, b1, b2, b3;
        ra_ldsm_x4_t(tl_u + (uint32_t)(rr * RB_PITCH + hh) * 2, b0, b1, b2, b3);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          ra_mma(accV[mt][2 * np], a[mt][0], a[mt][1], a[mt][2], a[mt][3], b0, b1);
          ra_mma(accV[mt][2 * np + 1], a[mt][0], a[mt][1], a[mt][2], a[mt][3], b2, b3);
        }
      }
    }
    // ---- dxhat[r][h] = sum_k W[r][k] B[k][h],  k = (A_t | dS_t)
    float dx[2][4][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int c = 0; c < 4; ++c) dx[i][j][c] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t a[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const int row = mt * 16 + (a_mat & 1) * 8 + a_row;
        ra_ldsm_x4(Ws_u + (uint32_t)(row * RB_WP + ks * 16 + a_koff) * 2, a[mt][0], a[mt][1], a[mt][2], a[mt][3]);
      }
#pragma unroll
      for (int np = 0; np < 2; ++np) {
        int k = ks * 16 + ((lane >> 3) & 1) * 8 + (lane & 7);
        k = k < 2 * T ? k : 2 * T;             // rows past 2T read the zero row
        const int hh = warp * 32 + np * 16 + (lane >> 4) * 8;
        uint32_t b0, b1, b2, b3;
        ra_ldsm_x4_t(Bs_u + (uint32_t)(k * RB_PITCH + hh) * 2, b0, b1, b2, b3);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          ra_mma(dx[mt][2 * np], a[mt][0], a[mt][1], a[mt][2], a[mt][3], b0, b1);
          ra_mma(dx[mt][2 * np + 1], a[mt][0], a[mt][1], a[mt][2], a[mt][3], b2, b3);
        }
      }
    }
    __syncwarp();                              // every lane's ldmatrix reads of this warp's columns precede the in-place writes
    // ---- LayerNorm backward + tanh derivative, written over the Y tile in place (this warp owns its 32 columns)
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int r = mt * 16 + g + 8 * half;
        const float4 rs = *reinterpret_cast<const float4*>(Rs + r * 4);      // rstd, -mu rstd, mean(dxhat) rstd, mean(dxhat xhat) rstd
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          __nv_bfloat162* yp = reinterpret_cast<__nv_bfloat162*>(tl + (size_t)r * RB_PITCH + warp * 32 + nt * 8 + 2 * tq);
          const float2 y = __bfloat1622float2(*yp);
          const float x0 = fmaf(y.x, rs.x, rs.y), x1 = fmaf(y.y, rs.x, rs.y);
          const float d0 = fmaf(-x0, rs.w, fmaf(dx[mt][nt][half * 2], rs.x, -rs.z)) * fmaf(-y.x, y.x, 1.f);
          const float d1 = fmaf(-x1, rs.w, fmaf(dx[mt][nt][half * 2 + 1], rs.x, -rs.z)) * fmaf(-y.y, y.y, 1.f);
          csum[nt][0] += d0; csum[nt][1] += d1;
          *yp = __floats2bfloat162_rn(d0, d1);
        }
      }
    ra_fence_async();                          // generic-proxy writes above, async-proxy reads (bulk store) below
    __syncthreads();
    if (lane < 4) {                            // each thread stores, and later refills, its own row: bulk groups are per thread
      const int row = warp * 4 + lane;
      if (row < nrows) ra_bulk_s2g(dpre + (int64_t)(r0 + row) * p.ldd, tl_u + (uint32_t)row * RB_PITCH * 2, (uint32_t)RB_HC * 2);
      ra_bulk_commit();
      if (tile + 2 < ntiles) ra_bulk_wait_read<1>();     // the store of tile-1 (same stage as tile+2) has read its shared memory
    }
    if (tile + 2 < ntiles) issue(tile + 2, (tile + 2) % RB_STAGES);
  }
  if (lane < 4) ra_bulk_wait_all();

  // ---- dF = dA + gamma o V ; dgamma += sum_t dA o U + F o V ; dbeta += sum_t dA ; dbias += column sums of dpre
  const float* cmuV = reinterpret_cast<const float*>(work + (int64_t)ntiles * RB_BLK_BYTES);
  float dg[4][2], db[4][2];
#pragma unroll
  for (int j = 0; j < 4; ++j) { dg[j][0] = dg[j][1] = 0.f; db[j][0] = db[j][1] = 0.f; }
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int t = mt * 16 + g + 8 * half;
      if (t < T) {
        float cm = 0.f;
#pragma unroll
        for (int sp = 0; sp < RB_PREP_SPLIT; ++sp) cm += cmuV[sp * 32 + t];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const int c = warp * 32 + nt * 8 + 2 * tq;
          const float2 gg = *reinterpret_cast<const float2*>(gamma + c);
          const float2 da = *reinterpret_cast<const float2*>(dAp + (int64_t)t * p.ldda + c);
          const float2 uu = *reinterpret_cast<const float2*>(Up + (int64_t)t * p.ldu + c);
          const float2 ff = *reinterpret_cast<const float2*>(Fp + (int64_t)t * p.ldf + c);
          const float v0 = accV[mt][nt][half * 2] - cm, v1 = accV[mt][nt][half * 2 + 1] - cm;
          *reinterpret_cast<float2*>(p.dF[e] + ((int64_t)b * T + t) * p.lddf + c0 + c) = make_float2(da.x + gg.x * v0, da.y + gg.y * v1);
          dg[nt][0] += da.x * uu.x + ff.x * v0; dg[nt][1] += da.y * uu.y + ff.y * v1;
          db[nt][0] += da.x; db[nt][1] += da.y;
        }
      }
    }
#pragma unroll
  for (int o = 4; o < 32; o <<= 1)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        dg[j][k] += __shfl_xor_sync(0xffffffffu, dg[j][k], o);
        db[j][k] += __shfl_xor_sync(0xffffffffu, db[j][k], o);
        csum[j][k] += __shfl_xor_sync(0xffffffffu, csum[j][k], o);
      }
  if (g == 0) {
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int c = c0 + warp * 32 + nt * 8 + 2 * tq + k;
        atomicAdd(p.dgamma[e] + c, dg[nt][k]);
        atomicAdd(p.dbeta[e] + c, db[nt][k]);
        if (p.dbias[e]) atomicAdd(p.dbias[e] + c, csum[nt][k]);
      }
  }
}

}  // namespace dlsg

using namespace dlsg;

extern "C" {

int dlsg_region_aggregate_supported(int32_t T, int32_t TR, int32_t H) {
  return (H == RA_H && T >= 1 && T <= RA_TMAX && TR >= 1) ? 1 : 0;
}

static inline bool ra_a16(const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; }

int dlsg_region_aggregate_fwd(const dlsg_region_agg_fwd_t* p, void* stream) {
  DLSG_REQUIRE(p, "region_aggregate_fwd: null params");
  DLSG_REQUIRE(dlsg_region_aggregate_supported(p->T, p->TR, p->H), "region_aggregate_fwd: unsupported shape T=%d TR=%d H=%d (H must be %d, T <= %d)",
               p->T, p->TR, p->H, RA_H, RA_TMAX);
  DLSG_REQUIRE(p->E >= 0 && p->E <= 2, "region_aggregate_fwd: E must be <= 2");
  if (p->B <= 0 || p->E <= 0) return 0;
  DLSG_REQUIRE(p->ldy % 8 == 0 && p->ldf % 4 == 0 && (p->scores_only || p->ldagg % 2 == 0) && p->ldu % 2 == 0,
               "region_aggregate_fwd: row pitches must keep 16-byte (bf16 / fp32 input) or 8-byte (output) alignment");
  for (int e = 0; e < p->E; ++e) {
    DLSG_REQUIRE(p->Y[e] && p->F[e] && p->gamma[e] && p->beta[e] && (p->scores_only || p->agg[e]), "region_aggregate_fwd: null operand (encoder %d)", e);
    DLSG_REQUIRE(ra_a16(p->Y[e]) && ra_a16(p->F[e]) && ra_a16(p->gamma[e]) && ra_a16(p->beta[e]) && ra_a16(p->agg[e]) && ra_a16(p->U[e]) &&
                 ra_a16(p->tconst[e]) && ra_a16(p->stats[e]),
                 "region_aggregate_fwd: operands must be 16-byte aligned (encoder %d)", e);
    DLSG_REQUIRE(!p->scores_only || p->St[e], "region_aggregate_fwd: scores_only needs St");
  }
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(region_aggregate_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RaSmem::TOTAL);
    attr = true;
  }
  DLSG_LAUNCH(region_aggregate_fwd_kernel, dim3(p->B, p->E), RA_THREADS, RaSmem::TOTAL, (cudaStream_t)stream, *p);
  return check_launch("region_aggregate_fwd_kernel");
}

int64_t dlsg_region_aggregate_bwd_workspace(int32_t B, int32_t T, int32_t TR) {
  return (int64_t)(B > 0 ? B : 0) * rb_work_bytes_per_clip(T, TR);
}

int dlsg_region_aggregate_bwd(const dlsg_region_agg_bwd_t* p, void* stream) {
  DLSG_REQUIRE(p, "region_aggregate_bwd: null params");
  DLSG_REQUIRE(dlsg_region_aggregate_supported(p->T, p->TR, p->H), "region_aggregate_bwd: unsupported shape T=%d TR=%d H=%d (H must be %d, T <= %d)",
               p->T, p->TR, p->H, RA_H, RA_TMAX);
  DLSG_REQUIRE(p->E >= 0 && p->E <= 2, "region_aggregate_bwd: E must be <= 2");
  if (p->B <= 0 || p->E <= 0) return 0;
  DLSG_REQUIRE(p->B <= 65535, "region_aggregate_bwd: B too large for one launch");
  DLSG_REQUIRE(p->ldy % 8 == 0 && p->ldd % 8 == 0 && p->ldf % 4 == 0 && p->ldda % 4 == 0 && p->ldu % 2 == 0 && p->lddf % 2 == 0,
               "region_aggregate_bwd: row pitches must keep 16-byte (inputs, dpre) / 8-byte alignment");
  for (int e = 0; e < p->E; ++e) {
    DLSG_REQUIRE(p->Y[e] && p->stats[e] && p->work[e] && p->St[e] && p->dSm[e] && p->F[e] && p->dA[e] && p->U[e] && p->tcF[e] && p->tcA[e] &&
                 p->gamma[e] && p->dpre[e] && p->dF[e] && p->dgamma[e] && p->dbeta[e], "region_aggregate_bwd: null operand (encoder %d)", e);
    DLSG_REQUIRE(ra_a16(p->Y[e]) && ra_a16(p->dpre[e]) && ra_a16(p->F[e]) && ra_a16(p->dA[e]) && ra_a16(p->U[e]) && ra_a16(p->gamma[e]) &&
                 ra_a16(p->tcF[e]) && ra_a16(p->tcA[e]) && ra_a16(p->dF[e]) && ra_a16(p->stats[e]) && ra_a16(p->work[e]),
                 "region_aggregate_bwd: operands must be 16-byte aligned (encoder %d)", e);
  }
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(region_aggregate_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RbSmem::TOTAL);
    attr = true;
  }
  DLSG_LAUNCH(region_aggregate_prep_kernel, dim3(RB_PREP_SPLIT, p->B, p->E), 256, 0, (cudaStream_t)stream, *p);
  DLSG_LAUNCH(region_aggregate_bwd_kernel, dim3(p->H / RB_HC, p->B, p->E), RB_THREADS, RbSmem::TOTAL, (cudaStream_t)stream, *p);
  return check_launch("region_aggregate_bwd_kernel");
}

}  // extern "C"
