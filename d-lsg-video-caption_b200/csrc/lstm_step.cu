// One LSTM time step in ONE launch: recurrent product h_{t-1} W_hh^T + cell, for up to two independent directions
// (the BiLSTM of EncoderVisual, models/layer.py:52 nn.LSTM(bidirectional=True); torch gate order i, f, g, o).
//
// The per-step chain of the BiLSTM used to be a tcgen05 split-K GEMM (8 MB of weights per direction, ~10 us: mostly the
// fixed cost of that kernel - barrier / TMEM set-up, tensor-map fetch, split-K partials) followed by the cell kernel
// (~5 us).  Here each CTA owns 16 hidden units of one direction, i.e. the 64 rows [i | f | g | o] x 16 of W_hh that
// produce them, so the whole cell is local to the CTA and nothing is exchanged between CTAs inside a step:
//   * the 64 x H weight slab and the (B <= 64) x H previous hidden state stream through a 4-stage cp.async ring in K-chunks
//     of 128 (bf16, rows padded to 272 B: conflict-free ldmatrix),
//   * mma.sync m16n8k16 (bf16, fp32 accumulate): warp = (gate, half of the batch), 64 k-steps,
//   * the four gate tiles meet in shared memory, 4 (unit, batch) cells per thread: + input projection, sigmoid / tanh,
//     c and h, written as fp32 (h for the layer output), bf16 (h as the next step's operand) and the activated gates (BPTT).
// grid (H / 16, directions), 256 threads, ~153 KB of shared memory.  Bound: L2 -> SM streaming of (64 + B) x H x 2 B per CTA.
#include "common.cuh"

namespace dlsg {

constexpr int LS_UNITS = 16;              // hidden units per CTA
constexpr int LS_ROWS = 4 * LS_UNITS;     // weight rows per CTA (gate-major: [i | f | g | o] x 16)
constexpr int LS_BMAX = 64;               // batch rows per launch
constexpr int LS_KC = 128;                // K chunk
constexpr int LS_PITCH = LS_KC + 8;       // bf16 elements: 272 B rows = 16 (mod 128)
constexpr int LS_STAGES = 4;
constexpr int LS_THREADS = 256;
constexpr int LS_GP = LS_BMAX + 1;        // pitch of the gate exchange buffer

struct LsSmem {
  static constexpr int CHUNK_BYTES = (LS_ROWS + LS_BMAX) * LS_PITCH * 2;     // 34816: W rows then h rows
  static constexpr int OFF_G = LS_STAGES * CHUNK_BYTES;                       // [4 gates][16 units][65] fp32
  static constexpr int G_BYTES = 4 * LS_UNITS * LS_GP * 4;
  static constexpr int TOTAL = OFF_G + G_BYTES;
};

__device__ __forceinline__ void ls_cp16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void ls_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void ls_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void ls_ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ls_mma(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(LS_THREADS, 1)
lstm_step_fwd_kernel(const dlsg_lstm_step_t p) {
  extern __shared__ __align__(128) uint8_t ls_smem[];
  float* G = reinterpret_cast<float*>(ls_smem + LsSmem::OFF_G);
  const int d = blockIdx.y, j0 = blockIdx.x * LS_UNITS;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int H = p.H, B = p.B;
  const __nv_bfloat16* W = reinterpret_cast<const __nv_bfloat16*>(p.W[d]);
  const __nv_bfloat16* hin = reinterpret_cast<const __nv_bfloat16*>(p.h_in[d]);
  const int nk = H / LS_KC;
  pdl_prologue();

  const int mt = warp & 3, nh = warp >> 2;           // gate, half of the batch
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[i][c] = 0.f;

  if (hin != nullptr) {
    // chunk loader: 2048 16-byte pieces per chunk (64 weight rows + 64 state rows, 16 pieces each), 8 per thread
    auto load_chunk = [&](int kc, int stage) {
      const uint32_t base = (uint32_t)__cvta_generic_to_shared(ls_smem + (size_t)stage * LsSmem::CHUNK_BYTES);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int piece = tid + LS_THREADS * i;
        const int row = piece >> 4, c16 = piece & 15;
        const uint32_t dst = base + (uint32_t)(row * LS_PITCH * 2 + c16 * 16);
        if (row < LS_ROWS) {
          const int grow = (row >> 4) * H + j0 + (row & 15);               // gate-major rows of W_hh
          ls_cp16(dst, W + (int64_t)grow * H + kc * LS_KC + c16 * 8);
        } else {
          const int b = row - LS_ROWS;
          if (b < B) ls_cp16(dst, hin + (int64_t)b * p.ldh_in + kc * LS_KC + c16 * 8);
          else *reinterpret_cast<uint4*>(ls_smem + (size_t)stage * LsSmem::CHUNK_BYTES + (size_t)row * LS_PITCH * 2 + c16 * 16) = make_uint4(0u, 0u, 0u, 0u);
        }
      }
    };
#pragma unroll
    for (int s = 0; s < LS_STAGES - 1; ++s) {
      if (s < nk) load_chunk(s, s);
      ls_commit();
    }
    const int a_mat = lane >> 3, a_row = lane & 7;
    for (int kc = 0; kc < nk; ++kc) {
      ls_wait<LS_STAGES - 2>();                       // chunk kc has landed (this thread's pieces) ...
      __syncthreads();                                // ... and everybody's; the stage refilled below was consumed at kc-1
      if (kc + LS_STAGES - 1 < nk) load_chunk(kc + LS_STAGES - 1, (kc + LS_STAGES - 1) % LS_STAGES);
      ls_commit();
      const uint32_t wb = (uint32_t)__cvta_generic_to_shared(ls_smem + (size_t)(kc % LS_STAGES) * LsSmem::CHUNK_BYTES);
      const uint32_t hb = wb + LS_ROWS * LS_PITCH * 2;
#pragma unroll
      for (int ks = 0; ks < LS_KC / 16; ++ks) {
        uint32_t a0, a1, a2, a3;
        ls_ldsm_x4(wb + (uint32_t)((mt * 16 + (a_mat & 1) * 8 + a_row) * LS_PITCH + ks * 16 + (a_mat >> 1) * 8) * 2, a0, a1, a2, a3);
#pragma unroll
        for (int np = 0; np < 2; ++np) {
          const int n = nh * 32 + np * 16 + (lane >> 4) * 8 + (lane & 7);
          uint32_t b0, b1, b2, b3;
          ls_ldsm_x4(hb + (uint32_t)(n * LS_PITCH + ks * 16 + ((lane >> 3) & 1) * 8) * 2, b0, b1, b2, b3);
          ls_mma(acc[2 * np], a0, a1, a2, a3, b0, b1);
          ls_mma(acc[2 * np + 1], a0, a1, a2, a3, b2, b3);
        }
      }
    }
    ls_wait<0>();
  }
  // ---- the four gate tiles of every (unit, batch) meet in shared memory
  {
    const int g = lane >> 2, tq = lane & 3;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int bcol = nh * 32 + nt * 8 + 2 * tq;
      float* g0 = G + (size_t)(mt * LS_UNITS + g) * LS_GP + bcol;
      float* g1 = G + (size_t)(mt * LS_UNITS + g + 8) * LS_GP + bcol;
      g0[0] = acc[nt][0]; g0[1] = acc[nt][1];
      g1[0] = acc[nt][2]; g1[1] = acc[nt][3];
    }
  }
  __syncthreads();
  // ---- cell: thread -> batch row b, 4 consecutive units
  const int b = tid >> 2, u0 = (tid & 3) * 4;
  if (b < B) {
    const float* gin = p.gin[d] + (int64_t)b * p.ldgin + j0 + u0;
    float pre[4][4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float4 x = *reinterpret_cast<const float4*>(gin + (int64_t)k * H);
      pre[k][0] = x.x + G[(size_t)(k * LS_UNITS + u0 + 0) * LS_GP + b];
      pre[k][1] = x.y + G[(size_t)(k * LS_UNITS + u0 + 1) * LS_GP + b];
      pre[k][2] = x.z + G[(size_t)(k * LS_UNITS + u0 + 2) * LS_GP + b];
      pre[k][3] = x.w + G[(size_t)(k * LS_UNITS + u0 + 3) * LS_GP + b];
    }
    const int64_t ei = (int64_t)b * H + j0 + u0;
    float4 cp = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.c_in[d]) cp = *reinterpret_cast<const float4*>(p.c_in[d] + ei);
    const float cp_[4] = {cp.x, cp.y, cp.z, cp.w};
    float ai[4], af[4], ag[4], ao[4], cc[4], hh[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      ai[u] = sigmoidf_(pre[0][u]); af[u] = sigmoidf_(pre[1][u]); ag[u] = tanhf(pre[2][u]); ao[u] = sigmoidf_(pre[3][u]);
      cc[u] = af[u] * cp_[u] + ai[u] * ag[u];
      hh[u] = ao[u] * tanhf(cc[u]);
    }
    float* a0 = p.acts[d] + (int64_t)b * 4 * H + j0 + u0;
    *reinterpret_cast<float4*>(a0) = make_float4(ai[0], ai[1], ai[2], ai[3]);
    *reinterpret_cast<float4*>(a0 + H) = make_float4(af[0], af[1], af[2], af[3]);
    *reinterpret_cast<float4*>(a0 + 2 * (int64_t)H) = make_float4(ag[0], ag[1], ag[2], ag[3]);
    *reinterpret_cast<float4*>(a0 + 3 * (int64_t)H) = make_float4(ao[0], ao[1], ao[2], ao[3]);
    *reinterpret_cast<float4*>(p.c_out[d] + ei) = make_float4(cc[0], cc[1], cc[2], cc[3]);
    if (p.h_out[d]) *reinterpret_cast<float4*>(p.h_out[d] + (int64_t)b * p.ldh_out + j0 + u0) = make_float4(hh[0], hh[1], hh[2], hh[3]);
    if (p.h_op[d]) {
      __nv_bfloat162 lo = __floats2bfloat162_rn(hh[0], hh[1]), hi = __floats2bfloat162_rn(hh[2], hh[3]);
      uint2 u; u.x = *reinterpret_cast<uint32_t*>(&lo); u.y = *reinterpret_cast<uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.h_op[d]) + (int64_t)b * p.ldh_op + j0 + u0) = u;
    }
  }
}

}  // namespace dlsg

using namespace dlsg;

extern "C" {

int dlsg_lstm_step_supported(int32_t B, int32_t H) {
  return (B >= 1 && B <= LS_BMAX && H >= LS_KC && H % LS_KC == 0) ? 1 : 0;
}

int dlsg_lstm_step_fwd(const dlsg_lstm_step_t* p, void* stream) {
  DLSG_REQUIRE(p, "lstm_step_fwd: null params");
  DLSG_REQUIRE(dlsg_lstm_step_supported(p->B, p->H), "lstm_step_fwd: unsupported shape B=%d H=%d (B <= %d, H a multiple of %d)", p->B, p->H,
               LS_BMAX, LS_KC);
  DLSG_REQUIRE(p->ndir >= 1 && p->ndir <= DLSG_LSTM_STEP_MAXG, "lstm_step_fwd: ndir must be 1..%d", DLSG_LSTM_STEP_MAXG);
  auto a16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  DLSG_REQUIRE(p->ldgin % 4 == 0 && p->ldh_in % 8 == 0 && p->ldh_out % 4 == 0 && p->ldh_op % 4 == 0, "lstm_step_fwd: unaligned row pitches");
  for (int d = 0; d < p->ndir; ++d) {
    DLSG_REQUIRE(p->W[d] && p->gin[d] && p->c_out[d] && p->acts[d], "lstm_step_fwd: null operand (direction %d)", d);
    DLSG_REQUIRE((p->h_in[d] != nullptr) == (p->h_in[0] != nullptr), "lstm_step_fwd: both directions must be first steps or neither");
    DLSG_REQUIRE(a16(p->W[d]) && a16(p->h_in[d]) && a16(p->gin[d]) && a16(p->c_in[d]) && a16(p->c_out[d]) && a16(p->acts[d]) &&
                 a16(p->h_out[d]) && (reinterpret_cast<uintptr_t>(p->h_op[d]) & 7) == 0, "lstm_step_fwd: unaligned operand (direction %d)", d);
  }
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(lstm_step_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LsSmem::TOTAL);
    attr = true;
  }
  DLSG_LAUNCH(lstm_step_fwd_kernel, dim3(p->H / LS_UNITS, p->ndir), LS_THREADS, LsSmem::TOTAL, (cudaStream_t)stream, *p);
  return check_launch("lstm_step_fwd_kernel");
}

}  // extern "C"
