// Shared device/host helpers for libdlsg (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/dlsg.h"

namespace dlsg {

void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define DLSG_REQUIRE(cond, ...)                      \
  do {                                               \
    if (!(cond)) {                                   \
      ::dlsg::set_error(__VA_ARGS__);                \
      return -1;                                     \
    }                                                \
  } while (0)

constexpr int kNumSM = 148;

// ---- launch helper: every kernel is launched with programmatic stream serialization (PDL) so that, inside a stream or
// a captured CUDA graph, the next kernel's CTAs are scheduled and run their prologue while the previous kernel drains.
// Kernels call pdl_prologue() before touching global memory (griddepcontrol.wait = all prerequisite grids complete and
// their writes visible).  With ~5-10 us kernels in the recurrent loops it hides the launch gap + prologue of the next kernel:
// 9.43 -> 9.07 ms/step on B200 inside the captured step.  On by default; DLSG_PDL=0 switches it off.
bool pdl_enabled();
template <typename K, typename... Args>
inline void launch(K kernel, dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, args...);
}
#define DLSG_LAUNCH(kernel, grid, block, smem, st, ...) ::dlsg::launch(kernel, dim3(grid), dim3(block), smem, st, __VA_ARGS__)

__device__ __forceinline__ void pdl_prologue() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}


__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// block-wide sum for blockDim.x <= 1024; `sh` needs 32 floats. All threads get the result.
__device__ __forceinline__ float block_sum(float v, float* sh) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  float r = (lane < nw) ? sh[lane] : 0.f;
  r = warp_sum(r);
  return r;
}
__device__ __forceinline__ float block_max(float v, float* sh) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  float r = (lane < nw) ? sh[lane] : -INFINITY;
  r = warp_max(r);
  return r;
}

__device__ __forceinline__ float ld_as_float(const void* p, int dtype, int64_t i) {
  return dtype == DLSG_F32 ? reinterpret_cast<const float*>(p)[i]
                           : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
}
__device__ __forceinline__ void st_from_float(void* p, int dtype, int64_t i, float v) {
  if (dtype == DLSG_F32) reinterpret_cast<float*>(p)[i] = v;
  else reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// Philox4x32-10 counter RNG -> one uniform in [0,1) per (seed, index).  Used for dropout so that
// forward and backward regenerate the same mask without storing it.
__device__ __forceinline__ uint32_t mulhi32(uint32_t a, uint32_t b) { return __umulhi(a, b); }
// Philox4x32-10: counter = element index / 4, the four 32-bit outputs serve four consecutive elements.
__device__ __forceinline__ uint4 philox4(uint64_t seed, uint64_t ctr) {
  uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = 0x243F6A88u, c3 = 0x85A308D3u;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = mulhi32(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = mulhi32(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}
__device__ __forceinline__ float u01(uint32_t r) { return (float)(r >> 8) * (1.0f / 16777216.0f); }
// multiplier of inverted dropout for element idx: 0 or 1/(1-p)
__device__ __forceinline__ float drop_scale(float p, uint64_t seed, uint64_t idx) {
  if (p <= 0.f) return 1.f;
  const uint4 r = philox4(seed, idx >> 2);
  const uint32_t k = (uint32_t)idx & 3u;
  const uint32_t v = k == 0 ? r.x : (k == 1 ? r.y : (k == 2 ? r.z : r.w));
  return u01(v) >= p ? 1.f / (1.f - p) : 0.f;
}
// the four multipliers of elements idx..idx+3 (idx % 4 == 0): one Philox evaluation
__device__ __forceinline__ float4 drop_mask4(float p, float keep, uint64_t seed, uint64_t idx) {
  const uint4 r = philox4(seed, idx >> 2);
  return make_float4(u01(r.x) >= p ? keep : 0.f, u01(r.y) >= p ? keep : 0.f, u01(r.z) >= p ? keep : 0.f, u01(r.w) >= p ? keep : 0.f);
}

}  // namespace dlsg
