// tcgen05 / TMA / TMEM GEMM for sm_100a:  C(P,Q) = A(P,K) . B(Q,K)^T, bf16 operands, fp32 accumulate.
//
// Persistent CTAs (one per SM) walk 128 x BN output tiles (BN in {32,64,128,256}); TMEM holds two accumulator
// stages so the epilogue of one tile overlaps the main loop of the next.  6 warps:
//   warp 0      : TMA producer  (cp.async.bulk.tensor.3d, 128B swizzle, mbarrier complete_tx)
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer (UMMA 128 x BN x 16, kind::f16)
//   warps 2..5  : epilogue - tcgen05.ld 32x32b.x32 from their TMEM lane quarter (warp_id % 4),
//                 bias / tanh / scale, smem-transposed coalesced stores (or direct transposed stores)
// smem ring of STAGES x {A 128x64 bf16, B BNx64 bf16}; full/empty mbarriers; tcgen05.commit frees slots.
// tile index -> (batch*splitk, P tile, Q tile) with the Q tile fastest.  Every spin-wait is bounded and traps
// instead of hanging.
#include <cuda.h>
#include <stdlib.h>
#include <cstdlib>
#include "common.cuh"

namespace dlsg {

constexpr int BM = 128;
constexpr int BK = 64;           // 64 bf16 = 128 bytes = one swizzle-128B row
constexpr int UMMA_K = 16;
constexpr int TC_THREADS = 192;
constexpr int EPI_PAD = 33;      // scalar staging pitch (conflict-free 4-byte accesses)
constexpr int EPI_PADV = 36;     // vector staging pitch (16-byte aligned rows, conflict-free 16-byte accesses)
constexpr int MN_BLOCK_BYTES = 64 * BK * 2;   // one 64(MN) x 64(K) bf16 box of an MN-major operand

struct TcParams {
  void* D; const float* bias;
  int64_t ldd, stride_d, stride_split;
  int P, Q;            // valid extents of the tile axes (rows of A / rows of B)
  int bias_mode;       // 0 none, 1 per-p, 2 per-q
  int store_t;         // 0: D[p*ldd+q]   1: D[q*ldd+p]
  int d_dtype, do_tanh, accum;
  int atomic;          // D += result with red.global.add.f32 (split-K CTAs all land in the same D)
  float alpha;
  int splitk, kb_total, kb_per_split;
  int ntm, ntn, tiles_total;
  int a_mn, b_mn;      // operand is MN-major in global memory (unit stride along P resp. Q, not along K)
  int pre_a, pre_b;    // operand is STATIC (DLSG_GEMM_*_STATIC: not written by the preceding kernels of the stream, e.g. a weight
                       // inside a recurrent loop): its first ring-full of tiles is fetched BEFORE griddepcontrol.wait
  unsigned long long* trace;   // debug: per-CTA phase timestamps (dlsg_debug_gemm_trace), nullptr in production
};

// phase timestamps: slot i of CTA b -> trace[b*16 + 2i] = %globaltimer (ns), trace[b*16 + 2i + 1] = clock64
__device__ __forceinline__ void tc_trace(const TcParams& prm, int i) {
  if (prm.trace) {
    unsigned long long g, c;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(c));
    prm.trace[blockIdx.x * 16 + 2 * i] = g;
    prm.trace[blockIdx.x * 16 + 2 * i + 1] = c;
  }
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  uint32_t spins = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) break;
    if (++spins > (1u << 26)) __trap();   // deadlock guard: fail loudly, never hang the box
  }
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// ---- CTA-pair (cta_group::2) helpers: PTX forms as in CUTLASS cute/arch/{copy_sm100_tma,mma_sm100_umma}.hpp, cutlass/arch/barrier.h
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;      // shared::cluster address of the same offset in the EVEN CTA of the pair
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
}
// both CTAs of a pair load their own tile; the bytes are counted on the LEADER's barrier (peer bit cleared)
__device__ __forceinline__ void tma_load_3d_2sm(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
// MMA completion -> the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}
// arrive on the barrier at this offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t rank) {
  asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\t"
               "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(bar), "r"(rank) : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, sm_100 version=1)
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);        // start address  [0,14)
  d |= (uint64_t)1 << 16;                          // LBO (unused for swizzled K-major) [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;                // SBO = 8 rows * 128 B           [32,46)
  d |= (uint64_t)1 << 46;                          // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                          // SWIZZLE_128B
  return d;
}
// MN-major, SWIZZLE_128B: the tile is stored as 64-element (128 B) MN blocks, each block = K rows of 128 B
// (8-row swizzle atoms of 1024 B).  Canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units
// (cute make_umma_desc<Major::MN>): LBO = byte distance between MN blocks, SBO = between 8-row K groups.
__device__ __forceinline__ uint64_t make_sw128_mn_desc(uint32_t saddr, uint32_t mn_block_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(mn_block_bytes >> 4) << 16;      // LBO
  d |= (uint64_t)(1024 >> 4) << 32;                // SBO = 8 K-rows * 128 B
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr) : "memory");
}

// NCTA = 2: a CTA pair (thread-block cluster of two, one TPC) computes a 256 x BN tile with tcgen05.mma.cta_group::2 - each CTA
// stages its own 128 rows of A and HALF of the B tile, the leader issues the MMAs for both, each CTA's TMEM receives its
// 128 accumulator rows.  Per k-block a CTA fills 16 KB + BN/2 rows instead of 16 KB + BN rows: the L2 -> SM operand traffic
// per flop drops by 1/3 at BN = 256 (the 1-CTA 128 x 256 tile is exactly L2-bandwidth bound at 85 flop/B).
template <int BN, int NCTA = 1> struct TcCfg {
#ifndef DLSG_SKINNY_STAGES
#define DLSG_SKINNY_STAGES 6      /* 8 measured 0.08 ms/step slower (216 KB CTAs co-reside less with the preceding kernel) */
#endif
  static constexpr int STAGES = (NCTA == 2) ? 6 : ((BN >= 256) ? 4 : (BN <= 64 ? DLSG_SKINNY_STAGES : 6));
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_ROWS = BN / NCTA;                                    // rows of B this CTA stages
  static constexpr int B_BYTES = B_ROWS * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int EPI_BYTES = 4 * 32 * EPI_PADV * 4;
  static constexpr int SMEM = STAGES * STAGE_BYTES + EPI_BYTES + 256 + 1024;  // +1024 alignment slack
  static constexpr int ACC_COLS = BN < 32 ? 32 : BN;                          // one accumulator stage
  static constexpr int TMEM_COLS = 2 * ACC_COLS;                              // double-buffered (<= 512)
};

__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Persistent: grid = min(#tiles, #SMs); CTA walks tiles t = blockIdx.x + i*gridDim.x (n-tile fastest so that the CTAs
// running concurrently share A row-panels and the whole weight matrix in L2).  Two TMEM accumulator stages let the
// epilogue of tile i overlap the TMA/MMA main loop of tile i+1.
template <int BN, int NCTA>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcParams prm) {
  using Cfg = TcCfg<BN, NCTA>;
  // CTA pair: rank 0 (leader) issues the MMAs; unit = the pair.  prm.ntm then counts 256-row tile PAIRS.
  const uint32_t crank = (NCTA == 2) ? cluster_ctarank() : 0u;
  const int unit = (NCTA == 2) ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int nunits = (NCTA == 2) ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* tiles = smem;
  float* epi = reinterpret_cast<float*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES + Cfg::EPI_BYTES);
  // bars[0..S) full, [S..2S) empty, [2S..2S+2) tmem_full, [2S+2..2S+4) tmem_empty ; then the TMEM base-address slot
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + Cfg::STAGES;
  uint64_t* tfull_bar = bars + 2 * Cfg::STAGES;
  uint64_t* tempty_bar = bars + 2 * Cfg::STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * Cfg::STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntn = prm.ntn, ntm = prm.ntm;
  const int tiles_total = prm.tiles_total;

  if (threadIdx.x == 0) {
    tc_trace(prm, 0);                                   // CTA entry
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");     // descriptor fetch overlaps the set-up below
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&tfull_bar[s]), 1);
      mbar_init(smem_u32(&tempty_bar[s]), 4 * NCTA);   // one arrival per epilogue warp (of both CTAs of a pair, on the leader)
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if constexpr (NCTA == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                   ::"r"(smem_u32(tmem_slot)), "n"(Cfg::TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                   ::"r"(smem_u32(tmem_slot)), "n"(Cfg::TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if constexpr (NCTA == 2) cluster_sync_all(); else __syncthreads();   // the pair's barriers exist before any remote arrive
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) tc_trace(prm, 1);               // barriers + TMEM ready
  // barrier init / TMEM allocation above overlap the previous kernel's tail (programmatic dependent launch).  The producer
  // thread goes further for a STATIC operand: the first ring-full of its tiles (a weight slab of a per-step recurrent GEMM:
  // up to STAGES x 16 KB per CTA) is requested before the dependency wait, so the weight stream overlaps the preceding
  // latency-bound kernels of the chain; only the dependent operand (64 activation rows) waits.
  const bool is_producer = (warp == 0 && lane == 0);
  int pre = 0;
  if (NCTA == 1 && is_producer && (prm.pre_a | prm.pre_b) && (int)blockIdx.x < tiles_total) {
    const int tile = blockIdx.x;
    const int tn = tile % ntn, tm = (tile / ntn) % ntm, z = tile / (ntn * ntm);
    const int zb = z / prm.splitk, zs = z % prm.splitk;
    const int kb_begin = zs * prm.kb_per_split;
    const int nkb = min(prm.kb_total, kb_begin + prm.kb_per_split) - kb_begin;
    pre = min(nkb, Cfg::STAGES);
    for (int kb = 0; kb < pre; ++kb) {
      const uint32_t full = smem_u32(&full_bar[kb]);
      mbar_expect_tx(full, Cfg::STAGE_BYTES);
      const uint32_t a_dst = smem_u32(tiles + kb * Cfg::STAGE_BYTES);
      const int kc = (kb_begin + kb) * BK;
      if (prm.pre_a) {
        if (!prm.a_mn) {
          tma_load_3d(a_dst, &tmA, full, kc, tm * BM, zb);
        } else {
          tma_load_3d(a_dst, &tmA, full, tm * BM, kc, zb);
          tma_load_3d(a_dst + MN_BLOCK_BYTES, &tmA, full, tm * BM + 64, kc, zb);
        }
      }
      if (prm.pre_b) {
        if (!prm.b_mn) {
          tma_load_3d(a_dst + Cfg::A_BYTES, &tmB, full, kc, tn * BN, zb);
        } else {
#pragma unroll
          for (int jb = 0; jb < BN / 64; ++jb)
            tma_load_3d(a_dst + Cfg::A_BYTES + jb * MN_BLOCK_BYTES, &tmB, full, tn * BN + 64 * jb, kc, zb);
        }
      }
    }
  }
  pdl_prologue();      // operands that depend on the preceding kernel are read below

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (int tile = unit; tile < tiles_total; tile += nunits) {
        const int tn = tile % ntn, tmu = (tile / ntn) % ntm, z = tile / (ntn * ntm);
        const int tm = (NCTA == 2) ? tmu * 2 + (int)crank : tmu;       // this CTA's 128-row tile
        const int zb = z / prm.splitk, zs = z % prm.splitk;
        const int kb_begin = zs * prm.kb_per_split;
        const int nkb = min(prm.kb_total, kb_begin + prm.kb_per_split) - kb_begin;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1);
          const uint32_t full = smem_u32(&full_bar[s]);
          if constexpr (NCTA == 2) {
            // both CTAs stage their own A rows and their half of the B tile; all bytes are counted on the leader's barrier
            if (crank == 0) mbar_expect_tx(full, 2 * Cfg::STAGE_BYTES);
            const uint32_t a_dst = smem_u32(tiles + s * Cfg::STAGE_BYTES);
            const int kc = (kb_begin + kb) * BK;
            const int bq0 = tn * BN + (int)crank * Cfg::B_ROWS;
            if (!prm.a_mn) {
              tma_load_3d_2sm(a_dst, &tmA, full, kc, tm * BM, zb);
            } else {
              tma_load_3d_2sm(a_dst, &tmA, full, tm * BM, kc, zb);
              tma_load_3d_2sm(a_dst + MN_BLOCK_BYTES, &tmA, full, tm * BM + 64, kc, zb);
            }
            if (!prm.b_mn) {
              tma_load_3d_2sm(a_dst + Cfg::A_BYTES, &tmB, full, kc, bq0, zb);
            } else {
#pragma unroll
              for (int jb = 0; jb < Cfg::B_ROWS / 64; ++jb)
                tma_load_3d_2sm(a_dst + Cfg::A_BYTES + jb * MN_BLOCK_BYTES, &tmB, full, bq0 + 64 * jb, kc, zb);
            }
            if (++s == Cfg::STAGES) { s = 0; ph ^= 1; }
            continue;
          }
          // the first `pre` k-blocks of this CTA's first tile: transaction count armed and static operand(s) already in flight
          const bool early = (tile == (int)blockIdx.x) && (kb < pre);
          if (!early) mbar_expect_tx(full, Cfg::STAGE_BYTES);
          const uint32_t a_dst = smem_u32(tiles + s * Cfg::STAGE_BYTES);
          const int kc = (kb_begin + kb) * BK;
          if (!(early && prm.pre_a)) {
            if (!prm.a_mn) {
              tma_load_3d(a_dst, &tmA, full, kc, tm * BM, zb);
            } else {                                     // two 64(P) x 64(K) boxes
              tma_load_3d(a_dst, &tmA, full, tm * BM, kc, zb);
              tma_load_3d(a_dst + MN_BLOCK_BYTES, &tmA, full, tm * BM + 64, kc, zb);
            }
          }
          if (!(early && prm.pre_b)) {
            if (!prm.b_mn) {
              tma_load_3d(a_dst + Cfg::A_BYTES, &tmB, full, kc, tn * BN, zb);
            } else {
#pragma unroll
              for (int jb = 0; jb < BN / 64; ++jb)
                tma_load_3d(a_dst + Cfg::A_BYTES + jb * MN_BLOCK_BYTES, &tmB, full, tn * BN + 64 * jb, kc, zb);
            }
          }
          if (tile == (int)blockIdx.x && kb == 0) tc_trace(prm, 2);     // first TMA issued
          if (++s == Cfg::STAGES) { s = 0; ph ^= 1; }
        }
      }
      tc_trace(prm, 3);                                 // all loads issued
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one elected thread; in a CTA pair only the leader's) =====
    if (lane == 0 && crank == 0) {
      // instruction descriptor: c=F32 (1<<4), a=b=BF16 (1<<7, 1<<10), K-major both, N>>3 @17, M>>4 @24
      //                         a_major @15, b_major @16 (1 = MN-major)
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((BM * NCTA) >> 4) << 24) |
                             (prm.a_mn ? (1u << 15) : 0u) | (prm.b_mn ? (1u << 16) : 0u);
      // descriptor advance per UMMA_K=16 step: K-major +32 B inside the swizzle row, MN-major +16 K-rows (2048 B)
      const uint32_t a_step = prm.a_mn ? (16 * 128) >> 4 : 2, b_step = prm.b_mn ? (16 * 128) >> 4 : 2;
      int s = 0; uint32_t ph = 0;
      int acc = 0; uint32_t acc_ph = 0;
      for (int tile = unit; tile < tiles_total; tile += nunits) {
        const int z = tile / (ntn * ntm);
        const int zs = z % prm.splitk;
        const int kb_begin = zs * prm.kb_per_split;
        const int nkb = min(prm.kb_total, kb_begin + prm.kb_per_split) - kb_begin;
        mbar_wait(smem_u32(&tempty_bar[acc]), acc_ph ^ 1);           // epilogue has drained this accumulator stage
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * Cfg::ACC_COLS);
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(smem_u32(&full_bar[s]), ph);
          if (tile == (int)blockIdx.x && kb == 0) tc_trace(prm, 4);     // first operands landed
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a_addr = smem_u32(tiles + s * Cfg::STAGE_BYTES);
          const uint64_t adesc = prm.a_mn ? make_sw128_mn_desc(a_addr, MN_BLOCK_BYTES) : make_sw128_desc(a_addr);
          const uint64_t bdesc = prm.b_mn ? make_sw128_mn_desc(a_addr + Cfg::A_BYTES, MN_BLOCK_BYTES) : make_sw128_desc(a_addr + Cfg::A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            if constexpr (NCTA == 2)
              umma_bf16_2sm(tmem_d, adesc + (uint64_t)(a_step * k), bdesc + (uint64_t)(b_step * k), idesc, (kb | k) != 0 ? 1u : 0u);
            else
              umma_bf16(tmem_d, adesc + (uint64_t)(a_step * k), bdesc + (uint64_t)(b_step * k), idesc, (kb | k) != 0 ? 1u : 0u);
          }
          if constexpr (NCTA == 2) umma_commit_2sm(smem_u32(&empty_bar[s]));   // frees the slot in BOTH CTAs
          else umma_commit(smem_u32(&empty_bar[s]));                 // frees the smem slot when the MMAs retire
          if (++s == Cfg::STAGES) { s = 0; ph ^= 1; }
        }
        if constexpr (NCTA == 2) umma_commit_2sm(smem_u32(&tfull_bar[acc]));   // both CTAs' epilogues
        else umma_commit(smem_u32(&tfull_bar[acc]));                 // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_ph ^= 1; }
      }
      tc_trace(prm, 5);                                 // all MMAs issued
    }
  } else {
    // ===== epilogue warps =====
    const int g = warp & 3;                                // TMEM lane quarter this warp may access
    float* st = epi + (warp - 2) * 32 * EPI_PADV;
    uint8_t* Dbase = reinterpret_cast<uint8_t*>(prm.D);
    const bool bf16_out = prm.d_dtype != DLSG_F32;
    // packed bf16x2 stores need 4-byte aligned pairs
    const bool pair_ok = bf16_out && !prm.store_t && !prm.accum && (prm.ldd % 2 == 0) && ((prm.stride_d | prm.stride_split) % 2 == 0) &&
                         ((reinterpret_cast<uintptr_t>(prm.D) & 3) == 0);
    const bool plain_f32 = !bf16_out && !prm.accum && !prm.do_tanh && !prm.atomic;   // fp32 store, nothing else: the per-step GEMMs
    const bool atomic_f32 = prm.atomic != 0;                            // (host guarantees fp32 D, no tanh)
    const bool accum_f32 = !bf16_out && prm.accum && !prm.do_tanh;      // D += ...: all 32 loads in flight before the first store
    // 16-byte row stores (lanes cover 64 B (bf16) / 128 B (fp32) contiguous per row, 8 / 4 rows per instruction)
    const bool al16 = (reinterpret_cast<uintptr_t>(prm.D) & 15) == 0;
    const bool bias16 = prm.bias == nullptr || (reinterpret_cast<uintptr_t>(prm.bias) & 15) == 0;
    const bool vec_bf16 = !prm.store_t && bf16_out && !prm.accum && !prm.atomic && al16 && bias16 && (prm.ldd % 8 == 0) &&
                          ((prm.stride_d | prm.stride_split) % 8 == 0);
    const bool vec_f32 = !prm.store_t && !bf16_out && !prm.do_tanh && !prm.atomic && al16 && bias16 && (prm.ldd % 4 == 0) &&
                         ((prm.stride_d | prm.stride_split) % 4 == 0);
    int acc = 0; uint32_t acc_ph = 0;
    for (int tile = unit; tile < tiles_total; tile += nunits) {
      const int tn = tile % ntn, tmu = (tile / ntn) % ntm, z = tile / (ntn * ntm);
      const int tm = (NCTA == 2) ? tmu * 2 + (int)crank : tmu;
      const int zb = z / prm.splitk, zs = z % prm.splitk;
      const int p0 = tm * BM, q0 = tn * BN;
      mbar_wait(smem_u32(&tfull_bar[acc]), acc_ph);
      if (threadIdx.x == 64 && tile + nunits >= tiles_total) tc_trace(prm, 6);   // last accumulator complete
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * Cfg::ACC_COLS) + ((uint32_t)(g * 32) << 16);
      const int p_row = p0 + g * 32 + lane;                  // this thread's accumulator row
      const int64_t doff = (int64_t)zb * prm.stride_d + (int64_t)zs * prm.stride_split;
      const bool add_bias = (prm.bias_mode != 0) && (zs == 0);
      float bias_p = 0.f;
      if (add_bias && prm.bias_mode == 1 && p_row < prm.P) bias_p = prm.bias[p_row];
      int nchunks = 0;
      for (int c0 = 0; c0 < BN; c0 += 32) if (q0 + c0 < prm.Q) ++nchunks;
#pragma unroll 1
      for (int ci = 0; ci < nchunks; ++ci) {
        const int c0 = ci * 32;
        uint32_t v[32];
        tmem_ld32(tmem_d + (uint32_t)c0, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (ci == nchunks - 1) {
          // all TMEM reads of this accumulator stage are complete: hand it back to the MMA warp
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) {
            if constexpr (NCTA == 2) mbar_arrive_cluster(smem_u32(&tempty_bar[acc]), 0);     // the leader's MMA thread waits for both CTAs
            else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tempty_bar[acc])) : "memory");
          }
        }
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) * prm.alpha + bias_p;
        const int qb = q0 + c0;
        const int qn = min(32, prm.Q - qb);                  // valid columns of this chunk (>= 1)
        const int pb = p0 + g * 32;
        const int pn = min(32, prm.P - pb);                  // valid rows of this warp's slab (may be <= 0)
        // Mode decisions are warp-uniform and hoisted: each store loop below is branch-free (predicated stores only).
        if (prm.store_t) {
          // element (p,q) -> D[q*ldd + p]: lanes are consecutive p -> coalesced
          if (plain_f32) {
            if (add_bias && prm.bias_mode == 2) {
              float bq[32];                                 // all loads in flight before the first add
#pragma unroll
              for (int j = 0; j < 32; ++j) bq[j] = (j < qn) ? __ldg(prm.bias + qb + j) : 0.f;
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] += bq[j];
            }
            if (lane < pn) {
              float* d = reinterpret_cast<float*>(Dbase) + doff + (int64_t)qb * prm.ldd + p_row;
              if (qn == 32) {
#pragma unroll
                for (int j = 0; j < 32; ++j) { *d = f[j]; d += prm.ldd; }
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) { if (j < qn) *d = f[j]; d += prm.ldd; }
              }
            }
          } else if (atomic_f32) {
            if (add_bias && prm.bias_mode == 2) {
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] += (j < qn) ? __ldg(prm.bias + qb + j) : 0.f;
            }
            if (lane < pn) {
              float* d = reinterpret_cast<float*>(Dbase) + doff + (int64_t)qb * prm.ldd + p_row;
#pragma unroll
              for (int j = 0; j < 32; ++j) { if (j < qn) atomicAdd(d, f[j]); d += prm.ldd; }
            }
          } else if (accum_f32) {
            if (add_bias && prm.bias_mode == 2) {
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] += (j < qn) ? __ldg(prm.bias + qb + j) : 0.f;
            }
            if (lane < pn) {
              float* d = reinterpret_cast<float*>(Dbase) + doff + (int64_t)qb * prm.ldd + p_row;
              float o[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) o[j] = (j < qn) ? d[(int64_t)j * prm.ldd] : 0.f;
#pragma unroll
              for (int j = 0; j < 32; ++j) if (j < qn) d[(int64_t)j * prm.ldd] = o[j] + f[j];
            }
          } else if (lane < pn) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {            // static register indexing (a rolled loop would spill f[] to local memory)
              if (j >= qn) break;
              float x = f[j];
              if (add_bias && prm.bias_mode == 2) x += prm.bias[qb + j];
              if (prm.do_tanh) x = bf16_out ? tanh_fast(x) : tanhf(x);
              const int64_t idx = doff + (int64_t)(qb + j) * prm.ldd + p_row;
              if (!bf16_out) {
                float* d = reinterpret_cast<float*>(Dbase) + idx;
                *d = prm.accum ? (*d + x) : x;
              } else {
                __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(Dbase) + idx;
                *d = __float2bfloat16_rn(prm.accum ? (__bfloat162float(*d) + x) : x);
              }
            }
          }
        } else if ((vec_bf16 || vec_f32) && qn == 32) {
          // element (p,q) -> D[p*ldd + q], full 32-column chunk, 16-byte aligned rows: stage through smem with
          // 16-byte accesses and store 16 bytes per lane
          {
            float4* s4 = reinterpret_cast<float4*>(st + lane * EPI_PADV);
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) s4[jj] = make_float4(f[4 * jj], f[4 * jj + 1], f[4 * jj + 2], f[4 * jj + 3]);
          }
          __syncwarp();
          if (vec_bf16) {
            const int cg = (lane & 3) * 8, rsel = lane >> 2;
            float bq[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) bq[u] = 0.f;
            if (add_bias && prm.bias_mode == 2) {
              const float4 b0 = __ldg(reinterpret_cast<const float4*>(prm.bias + qb + cg));
              const float4 b1 = __ldg(reinterpret_cast<const float4*>(prm.bias + qb + cg + 4));
              bq[0] = b0.x; bq[1] = b0.y; bq[2] = b0.z; bq[3] = b0.w; bq[4] = b1.x; bq[5] = b1.y; bq[6] = b1.z; bq[7] = b1.w;
            }
            __nv_bfloat16* dbase = reinterpret_cast<__nv_bfloat16*>(Dbase) + doff + (int64_t)pb * prm.ldd + qb + cg;
#pragma unroll
            for (int it = 0; it < 4; ++it) {
              const int r = it * 8 + rsel;
              const float4 a0 = *reinterpret_cast<const float4*>(st + r * EPI_PADV + cg);
              const float4 a1 = *reinterpret_cast<const float4*>(st + r * EPI_PADV + cg + 4);
              float x[8] = {a0.x + bq[0], a0.y + bq[1], a0.z + bq[2], a0.w + bq[3], a1.x + bq[4], a1.y + bq[5], a1.z + bq[6], a1.w + bq[7]};
              if (prm.do_tanh) {
#pragma unroll
                for (int u = 0; u < 8; ++u) x[u] = tanh_fast(x[u]);
              }
              __nv_bfloat162 h0 = __floats2bfloat162_rn(x[0], x[1]), h1 = __floats2bfloat162_rn(x[2], x[3]);
              __nv_bfloat162 h2 = __floats2bfloat162_rn(x[4], x[5]), h3 = __floats2bfloat162_rn(x[6], x[7]);
              uint4 pk;
              pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
              pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
              if (r < pn) *reinterpret_cast<uint4*>(dbase + (int64_t)r * prm.ldd) = pk;
            }
          } else {
            const int cg = (lane & 7) * 4, rsel = lane >> 3;
            float4 bq = make_float4(0.f, 0.f, 0.f, 0.f);
            if (add_bias && prm.bias_mode == 2) bq = __ldg(reinterpret_cast<const float4*>(prm.bias + qb + cg));
            float* dbase = reinterpret_cast<float*>(Dbase) + doff + (int64_t)pb * prm.ldd + qb + cg;
            float4 o[8];
            if (prm.accum) {
#pragma unroll
              for (int it = 0; it < 8; ++it) {
                const int r = it * 4 + rsel;
                o[it] = (r < pn) ? *reinterpret_cast<const float4*>(dbase + (int64_t)r * prm.ldd) : make_float4(0.f, 0.f, 0.f, 0.f);
              }
            }
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              const int r = it * 4 + rsel;
              float4 a = *reinterpret_cast<const float4*>(st + r * EPI_PADV + cg);
              a.x += bq.x; a.y += bq.y; a.z += bq.z; a.w += bq.w;
              if (prm.accum) { a.x += o[it].x; a.y += o[it].y; a.z += o[it].z; a.w += o[it].w; }
              if (r < pn) *reinterpret_cast<float4*>(dbase + (int64_t)r * prm.ldd) = a;
            }
          }
          __syncwarp();
        } else {
          // element (p,q) -> D[p*ldd + q]: transpose through smem so lanes walk q
#pragma unroll
          for (int j = 0; j < 32; ++j) st[lane * EPI_PAD + j] = f[j];
          __syncwarp();
          if (pair_ok && qn == 32) {
            // bf16 output: lane -> (row 2*it + lane/16, column pair 2*(lane%16)): 64 B contiguous per half-warp
            const int cp = 2 * (lane & 15), rsel = lane >> 4;
            float b0 = 0.f, b1 = 0.f;
            if (add_bias && prm.bias_mode == 2) { b0 = prm.bias[qb + cp]; b1 = prm.bias[qb + cp + 1]; }
            __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(Dbase) + doff + (int64_t)(pb + rsel) * prm.ldd + qb + cp;
            const float* sp = st + rsel * EPI_PAD + cp;
            const int64_t ld2 = 2 * prm.ldd;
            const int pv = pn >= 32 ? 32 : pn;              // full slab: no predicates at all
            if (prm.do_tanh) {
#pragma unroll
              for (int it = 0; it < 16; ++it) {
                if (pv == 32 || 2 * it + rsel < pv)
                  *reinterpret_cast<__nv_bfloat162*>(d) =
                      __floats2bfloat162_rn(tanh_fast(sp[2 * it * EPI_PAD] + b0), tanh_fast(sp[2 * it * EPI_PAD + 1] + b1));
                d += ld2;
              }
            } else {
#pragma unroll
              for (int it = 0; it < 16; ++it) {
                if (pv == 32 || 2 * it + rsel < pv)
                  *reinterpret_cast<__nv_bfloat162*>(d) = __floats2bfloat162_rn(sp[2 * it * EPI_PAD] + b0, sp[2 * it * EPI_PAD + 1] + b1);
                d += ld2;
              }
            }
          } else if (lane < qn) {
            const int q = qb + lane;
            float bias_q = 0.f;
            if (add_bias && prm.bias_mode == 2) bias_q = prm.bias[q];
            if (plain_f32) {
              float* d = reinterpret_cast<float*>(Dbase) + doff + (int64_t)pb * prm.ldd + q;
              if (pn >= 32) {
#pragma unroll
                for (int i = 0; i < 32; ++i) { *d = st[i * EPI_PAD + lane] + bias_q; d += prm.ldd; }
              } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) { if (i < pn) *d = st[i * EPI_PAD + lane] + bias_q; d += prm.ldd; }
              }
            } else if (atomic_f32) {
              float* d = reinterpret_cast<float*>(Dbase) + doff + (int64_t)pb * prm.ldd + q;
#pragma unroll
              for (int i = 0; i < 32; ++i) { if (i < pn) atomicAdd(d, st[i * EPI_PAD + lane] + bias_q); d += prm.ldd; }
            } else if (accum_f32) {
              float* d = reinterpret_cast<float*>(Dbase) + doff + (int64_t)pb * prm.ldd + q;
              float o[32];
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] = (i < pn) ? d[(int64_t)i * prm.ldd] : 0.f;
#pragma unroll
              for (int i = 0; i < 32; ++i) if (i < pn) d[(int64_t)i * prm.ldd] = o[i] + st[i * EPI_PAD + lane] + bias_q;
            } else {
#pragma unroll 1
              for (int i = 0; i < pn; ++i) {
                float x = st[i * EPI_PAD + lane] + bias_q;
                if (prm.do_tanh) x = bf16_out ? tanh_fast(x) : tanhf(x);
                const int64_t idx = doff + (int64_t)(pb + i) * prm.ldd + q;
                if (!bf16_out) {
                  float* d = reinterpret_cast<float*>(Dbase) + idx;
                  *d = prm.accum ? (*d + x) : x;
                } else {
                  __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(Dbase) + idx;
                  *d = __float2bfloat16_rn(prm.accum ? (__bfloat162float(*d) + x) : x);
                }
              }
            }
          }
          __syncwarp();
        }
      }
      if (++acc == 2) { acc = 0; acc_ph ^= 1; }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if constexpr (NCTA == 2) cluster_sync_all(); else __syncthreads();   // neither CTA of a pair leaves while the other may still touch it
  if (threadIdx.x == 0) tc_trace(prm, 7);               // epilogue stores issued, CTA about to exit
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if constexpr (NCTA == 2)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(Cfg::TMEM_COLS) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(Cfg::TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------- host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
  }
  return fn;
}

// K-major operand (mn = false): dims {K, rows, batch}, ld = row pitch, box {64 K, box_rows}.
// MN-major operand (mn = true):  dims {rows, K, batch}, ld = pitch between consecutive k, box {64 rows, 64 K}.
static int make_map(CUtensorMap* tm, const void* ptr, int64_t rows, int64_t K, int64_t ld, int64_t batch,
                    int64_t stride_batch, int box_rows, bool mn) {
  EncodeTiledFn enc = get_encode();
  DLSG_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not found");
  if (batch <= 1 || stride_batch == 0) stride_batch = (mn ? K : rows) * ld;
  cuuint64_t gdim[3] = {(cuuint64_t)(mn ? rows : K), (cuuint64_t)(mn ? K : rows), (cuuint64_t)(batch < 1 ? 1 : batch)};
  cuuint64_t gstr[2] = {(cuuint64_t)ld * 2, (cuuint64_t)stride_batch * 2};
  cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)(mn ? BK : box_rows), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  DLSG_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "gemm_tc: operand base not 16-byte aligned");
  DLSG_REQUIRE((gstr[0] % 16) == 0 && (gstr[1] % 16) == 0, "gemm_tc: ld/stride must be multiples of 8 elements (ld=%lld stride=%lld)",
               (long long)ld, (long long)stride_batch);
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DLSG_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) rows=%lld K=%lld ld=%lld", (int)r, (long long)rows,
               (long long)K, (long long)ld);
  return 0;
}

// sum split-K partials (S, batch, M, N) fp32 from the workspace and apply the epilogue into the user's D
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const float* __restrict__ ws, int S, int batch, int M, int N, void* __restrict__ D, int d_dtype, int64_t ldd,
                     int64_t stride_d, const float* __restrict__ bias, int flags, float alpha) {
  pdl_prologue();
  const int64_t per = (int64_t)M * N, total = per * batch;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(e / per);
    const int64_t r = e % per;
    const int m = (int)(r / N), n = (int)(r % N);
    float x = 0.f;
#pragma unroll
    for (int s = 0; s < 16; ++s)             // unrolled + predicated: all partial loads are in flight together
      if (s < S) x += ws[(int64_t)s * total + e];
    for (int s = 16; s < S; ++s) x += ws[(int64_t)s * total + e];
    x *= alpha;
    if (bias) {
      if (flags & DLSG_EPI_BIAS_N) x += bias[n];
      if (flags & DLSG_EPI_BIAS_M) x += bias[m];
    }
    if (flags & DLSG_EPI_TANH) x = tanhf(x);
    const int64_t idx = (int64_t)b * stride_d + ((flags & DLSG_EPI_STORE_T) ? ((int64_t)n * ldd + m) : ((int64_t)m * ldd + n));
    if (flags & DLSG_EPI_ACCUM) x += ld_as_float(D, d_dtype, idx);
    st_from_float(D, d_dtype, idx, x);
  }
}

template <int BN>
static int launch_tc(const CUtensorMap& ta, const CUtensorMap& tb, const TcParams& prm, dim3 grid, cudaStream_t st) {
  using Cfg = TcCfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BN, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
    DLSG_REQUIRE(e == cudaSuccess, "gemm_tc: cudaFuncSetAttribute(%d) failed: %s", Cfg::SMEM, cudaGetErrorString(e));
    attr_set = true;
  }
  DLSG_LAUNCH((gemm_tc_kernel<BN, 1>), grid, TC_THREADS, Cfg::SMEM, st, ta, tb, prm);
  return check_launch("gemm_tc_kernel");
}

// CTA-pair variant: thread-block cluster (2,1,1) + programmatic dependent launch
template <int BN>
static int launch_tc_pair(const CUtensorMap& ta, const CUtensorMap& tb, const TcParams& prm, dim3 grid, cudaStream_t st) {
  using Cfg = TcCfg<BN, 2>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BN, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
    DLSG_REQUIRE(e == cudaSuccess, "gemm_tc: cudaFuncSetAttribute(%d) failed: %s", Cfg::SMEM, cudaGetErrorString(e));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = dim3(TC_THREADS); cfg.dynamicSmemBytes = Cfg::SMEM; cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 2 : 1;
  cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, 2>, ta, tb, prm);
  return check_launch("gemm_tc_kernel<pair>");
}

static int gemm_tc_direct(const dlsg_gemm_t* g, cudaStream_t st);
static unsigned long long* g_tc_trace = nullptr;
void gemm_tc_set_trace(void* p) { g_tc_trace = static_cast<unsigned long long*>(p); }

int gemm_tc_dispatch(const dlsg_gemm_t* g, cudaStream_t st) {
  if (g->flags & DLSG_EPI_ATOMIC) {
    // D += A.B^T with atomic adds: split K over the idle SMs, every split's epilogue lands in D itself
    DLSG_REQUIRE(g->d_dtype == DLSG_F32 && !(g->flags & (DLSG_EPI_TANH | DLSG_EPI_ACCUM)) && g->splitk <= 1,
                 "gemm_tc: DLSG_EPI_ATOMIC needs an fp32 D, no tanh / accumulate flag and no explicit splitk");
    dlsg_gemm_t part = *g;
    part.workspace = nullptr;
    const int batch = g->batch < 1 ? 1 : g->batch;
    const bool swap = (g->M <= 64 && g->N > g->M);
    const int P = swap ? g->N : g->M, Q = swap ? g->M : g->N;
    const int bn = Q <= 32 ? 32 : (Q <= 64 ? 64 : 128);
    const int64_t tiles = (int64_t)((P + BM - 1) / BM) * ((Q + bn - 1) / bn) * batch;
    const int kb = (g->K + BK - 1) / BK;
    int S = 1;
    if (tiles * 2 <= kNumSM && kb >= 8) {
      S = (int)(kNumSM / tiles);
      if (S > kb / 4) S = kb / 4;
      if (S > 16) S = 16;
      const int per = (kb + S - 1) / S;
      S = (kb + per - 1) / per;
    }
    part.splitk = S < 1 ? 1 : S;
    part.stride_split = 0;
    return gemm_tc_direct(&part, st);
  }
  // Automatic split-K for skinny problems: too few 128-row tiles to fill 148 SMs and a long K loop
  // (the per-step recurrent GEMMs: M = batch = 64, weights streamed once).
  if (g->splitk <= 1 && g->workspace && g->M > 0 && g->N > 0) {
    const int batch = g->batch < 1 ? 1 : g->batch;
    const bool swap = (g->M <= 64 && g->N > g->M);
    const int P = swap ? g->N : g->M, Q = swap ? g->M : g->N;
    const int bn = Q <= 32 ? 32 : (Q <= 64 ? 64 : 128);
    const int64_t tiles = (int64_t)((P + BM - 1) / BM) * ((Q + bn - 1) / bn) * batch;
    const int kb = (g->K + BK - 1) / BK;
    // the reduce launch costs ~2-3 us in a dependent chain: only worth it when every split still streams >= 16 k-blocks
    // (8 -> 16 measured: step 6.28 -> 6.24 ms, greedy B=256 35.4 -> 37.8 k captions/s, beam-5 18.6 -> 19.6 k; profiles/r04e_*)
    // (DLSG_AUTOSPLIT_MIN_KB: measurement switch for that threshold, k-blocks per split)
    static const int min_kb = [] { const char* e = getenv("DLSG_AUTOSPLIT_MIN_KB"); const int v = e ? atoi(e) : 16; return v < 2 ? 2 : v; }();
    if (tiles * 2 <= kNumSM && kb >= 2 * min_kb) {
      int S = (int)(kNumSM / tiles);      // floor: all tiles*S CTAs run in ONE wave of the persistent grid
      if (S > kb / min_kb) S = kb / min_kb;
      if (S > 16) S = 16;
      const int per = (kb + S - 1) / S;
      S = (kb + per - 1) / per;
      const int64_t need = (int64_t)S * batch * g->M * g->N * 4;
      if (S >= 2 && need <= g->workspace_bytes) {
        dlsg_gemm_t part = *g;
        part.D = g->workspace; part.d_dtype = DLSG_F32; part.bias = nullptr; part.flags = g->flags & (DLSG_GEMM_A_STATIC | DLSG_GEMM_B_STATIC); part.alpha = 1.f;
        part.ldd = g->N; part.stride_d = (int64_t)g->M * g->N; part.splitk = S;
        part.stride_split = (int64_t)batch * g->M * g->N; part.workspace = nullptr;
        if (int rc = gemm_tc_direct(&part, st)) return rc;
        const int64_t total = (int64_t)batch * g->M * g->N;
        int64_t blocks = (total + 255) / 256;
        if (blocks > kNumSM * 8) blocks = kNumSM * 8;
        DLSG_LAUNCH(splitk_reduce_kernel, (unsigned)blocks, 256, 0, st, (const float*)g->workspace, S, batch, g->M, g->N, g->D, g->d_dtype,
                                                               g->ldd, g->stride_d, g->bias, g->flags, g->alpha);
        return check_launch("splitk_reduce_kernel");
      }
    }
  }
  return gemm_tc_direct(g, st);
}

static int gemm_tc_direct(const dlsg_gemm_t* g, cudaStream_t st) {
  DLSG_REQUIRE(g->a_dtype == DLSG_BF16 && g->b_dtype == DLSG_BF16, "gemm_tc: operands must be bf16");
  // each operand is K-major (unit stride along K) or MN-major (unit stride along its row axis: a transposed view)
  const bool a_mn_user = !(g->sak == 1 || g->K == 1), b_mn_user = !(g->sbk == 1 || g->K == 1);
  DLSG_REQUIRE((!a_mn_user || g->sam == 1 || g->M == 1) && (!b_mn_user || g->sbn == 1 || g->N == 1),
               "gemm_tc: each operand needs a unit stride along K or along its row axis");
  DLSG_REQUIRE(g->M > 0 && g->N > 0 && g->K > 0, "gemm_tc: empty problem");
  const int batch = g->batch < 1 ? 1 : g->batch;
  // Put the wide operand on the 128-row UMMA-M side; skinny activations (M<=64) ride the N side ("swap-AB").
  const bool swap = (g->M <= 64 && g->N > g->M);
  const void* Ap = swap ? g->B : g->A;
  const void* Bq = swap ? g->A : g->B;
  const int P = swap ? g->N : g->M, Q = swap ? g->M : g->N;
  const bool p_mn = swap ? b_mn_user : a_mn_user, q_mn = swap ? a_mn_user : b_mn_user;
  const int64_t ldp = p_mn ? (swap ? g->sbk : g->sak) : (swap ? g->sbn : g->sam);
  const int64_t ldq = q_mn ? (swap ? g->sak : g->sbk) : (swap ? g->sam : g->sbn);
  const int64_t strp = swap ? g->stride_b : g->stride_a, strq = swap ? g->stride_a : g->stride_b;
  TcParams prm;
  prm.trace = g_tc_trace;
  prm.D = g->D; prm.bias = g->bias; prm.ldd = g->ldd; prm.stride_d = g->stride_d; prm.stride_split = g->stride_split;
  prm.P = P; prm.Q = Q;
  const bool user_t = (g->flags & DLSG_EPI_STORE_T) != 0;
  prm.store_t = (swap != user_t) ? 1 : 0;
  prm.bias_mode = 0;
  if (g->bias && (g->flags & DLSG_EPI_BIAS_N)) prm.bias_mode = swap ? 1 : 2;
  if (g->bias && (g->flags & DLSG_EPI_BIAS_M)) prm.bias_mode = swap ? 2 : 1;
  prm.d_dtype = g->d_dtype; prm.do_tanh = (g->flags & DLSG_EPI_TANH) ? 1 : 0; prm.accum = (g->flags & DLSG_EPI_ACCUM) ? 1 : 0;
  prm.atomic = (g->flags & DLSG_EPI_ATOMIC) ? 1 : 0;
  prm.alpha = g->alpha;
  prm.kb_total = (g->K + BK - 1) / BK;
  int splitk = g->splitk < 1 ? 1 : g->splitk;
  if (splitk > prm.kb_total) splitk = prm.kb_total;
  prm.kb_per_split = (prm.kb_total + splitk - 1) / splitk;
  splitk = (prm.kb_total + prm.kb_per_split - 1) / prm.kb_per_split;   // no empty splits
  DLSG_REQUIRE(splitk == (g->splitk < 1 ? 1 : g->splitk), "gemm_tc: splitk=%d does not divide %d k-blocks without empty splits (use %d)",
               g->splitk, prm.kb_total, splitk);
  prm.splitk = splitk;
  DLSG_REQUIRE(!(splitk > 1 && (prm.do_tanh)), "gemm_tc: tanh epilogue is incompatible with split-K");

  const int ptiles = (P + BM - 1) / BM;
  auto tiles_for = [&](int bn) { return (int64_t)ptiles * ((Q + bn - 1) / bn) * batch * splitk; };
  // Tile width: fewest (waves of the persistent grid) x (time of one k-block of a 128 x bn tile).  A k-block costs the
  // larger of its operand fill ((128 + bn) rows x 128 B at ~106 B/ns per SM, measured) and its MMAs (11.1 TFLOP/s per SM).
  auto cost = [&](int bn) {
    const int64_t waves = (tiles_for(bn) + kNumSM - 1) / kNumSM;
    const double fill = (128.0 + bn) * 128.0 / 106.0, mma = 128.0 * bn * 128.0 / 11100.0;
    return (double)waves * (fill > mma ? fill : mma);
  };
  int bn;
  if (Q <= 32 && !q_mn) bn = 32;             // (an MN-major Q operand is loaded in 64-element blocks)
  else if (Q <= 64) bn = 64;
  else {
    bn = 128;
    if (Q > 128 && cost(256) <= cost(bn)) bn = 256;
    if (cost(64) < cost(bn)) bn = 64;        // small problems: more, narrower tiles cover more SMs
  }
  // CTA pairs (256 x bn tiles, tcgen05.mma.cta_group::2) for the big tensor-bound products: many tiles, no split-K.  The
  // 1-CTA 128 x 256 tile moves (128 + 256) rows per k-block for 128 x 256 outputs = 85 flop/B and is L2 -> SM bandwidth bound
  // (measured 15.3 TB/s aggregate = 1300 TFLOP/s); a pair moves (256 + 256) rows for 256 x 256 outputs = 128 flop/B.
  static const bool pair_on = [] { const char* e = getenv("DLSG_GEMM_2CTA"); return !(e && e[0] == '0'); }();
  const bool pair = pair_on && splitk == 1 && bn >= 128 && ptiles >= 2 && tiles_for(bn) >= 40;   // a pair occupies two SMs: never fewer SMs busy
  CUtensorMap ta, tb;
  if (make_map(&ta, Ap, P, g->K, ldp, batch, strp, BM, p_mn)) return -1;
  if (make_map(&tb, Bq, Q, g->K, ldq, batch, strq, pair ? bn / 2 : bn, q_mn)) return -1;
  prm.a_mn = p_mn ? 1 : 0; prm.b_mn = q_mn ? 1 : 0;
  if (pair) {
    prm.pre_a = prm.pre_b = 0;
    prm.ntm = (ptiles + 1) / 2;                  // 256-row tile pairs (an odd last 128-row tile is zero-filled / masked)
    prm.ntn = (Q + bn - 1) / bn;
    const int64_t units = (int64_t)prm.ntm * prm.ntn * batch;
    DLSG_REQUIRE(units < (1ll << 31), "gemm_tc: too many tiles");
    prm.tiles_total = (int)units;
    const int64_t npairs = units < kNumSM / 2 ? units : kNumSM / 2;
    dim3 grid2((unsigned)(2 * npairs));
    return bn == 128 ? launch_tc_pair<128>(ta, tb, prm, grid2, st) : launch_tc_pair<256>(ta, tb, prm, grid2, st);
  }
  {
    const bool a_static = (g->flags & DLSG_GEMM_A_STATIC) != 0, b_static = (g->flags & DLSG_GEMM_B_STATIC) != 0;
    prm.pre_a = (swap ? b_static : a_static) ? 1 : 0;
    prm.pre_b = (swap ? a_static : b_static) ? 1 : 0;
  }
  prm.ntm = ptiles;
  prm.ntn = (Q + bn - 1) / bn;
  const int64_t tiles_total = (int64_t)prm.ntm * prm.ntn * batch * splitk;
  DLSG_REQUIRE(tiles_total < (1ll << 31), "gemm_tc: too many tiles");
  prm.tiles_total = (int)tiles_total;
  dim3 grid((unsigned)(tiles_total < kNumSM ? tiles_total : kNumSM));
  switch (bn) {
    case 32: return launch_tc<32>(ta, tb, prm, grid, st);
    case 64: return launch_tc<64>(ta, tb, prm, grid, st);
    case 128: return launch_tc<128>(ta, tb, prm, grid, st);
    default: return launch_tc<256>(ta, tb, prm, grid, st);
  }
}

}  // namespace dlsg
