// Shared device/host helpers for libax2d (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/ax2d.h"

namespace ax2d {

void set_error(const char* fmt, ...);
void count_launches(int n);   // host-side tally of kernels enqueued by this library (ax2d_launch_count)

#define AX2D_CHECK_ARG(cond, ...)                    \
  do {                                               \
    if (!(cond)) {                                   \
      ::ax2d::set_error(__VA_ARGS__);                \
      return AX2D_ERR_ARG;                           \
    }                                                \
  } while (0)

#define AX2D_CHECK_ALIGN(ptr)                                              \
  do {                                                                     \
    if ((reinterpret_cast<uintptr_t>(ptr) & 15u) != 0) {                   \
      ::ax2d::set_error("%s: pointer %s is not 16-byte aligned", __func__, #ptr); \
      return AX2D_ERR_ALIGN;                                               \
    }                                                                      \
  } while (0)

inline int launch_status(const char* what, int n_kernels = 1) {
  count_launches(n_kernels);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: kernel launch failed: %s", what, cudaGetErrorString(e));
    return AX2D_ERR_LAUNCH;
  }
  return AX2D_OK;
}

constexpr int kNumSMs = 148;   // B200

// ---------------------------------------------------------------------------------- programmatic dependent launch
// Every kernel of the training / inference step is enqueued through launch_k and starts with pdl_wait() (nothing of a
// previous kernel is read or overwritten before it) followed by pdl_trigger(): inside the captured step graph the
// next kernel's CTAs are scheduled, run their prologue (barrier init, TMEM allocation, descriptor fetch) and park in
// griddepcontrol.wait while the tail of the current kernel drains, instead of paying a full launch after it.
// The launch attribute is opt-in (ax2d_set_pdl(1) / AX2D_PDL=1; without it the two instructions are no-ops): measured on
// the captured C2 step it is neutral (3.057 vs 3.067 ms), on 48 back-to-back aggregation launches 19.6 -> 18.9 us.
bool pdl_enabled();
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() {
  pdl_wait();
  pdl_trigger();
}
template <typename... KArgs, typename... Args>
inline void launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1u : 0u;
  (void)cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);   // the error is picked up by launch_status
}

// ---------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  uint32_t addr = smem_u32(bar);
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}
// 1-D bulk async copy global -> shared (TMA engine, no tensor map).  dst/src 16-B aligned, bytes % 16 == 0.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ float4 ld_nc_f4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_na_f4(float4* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
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

// ---------------------------------------------------------------------------------- activations
// utils/activation.py:23-29 -- torch defaults: LeakyReLU(0.01), ELU(alpha=1), GELU(erf).
__device__ __forceinline__ float act_fwd(int act, float v) {
  switch (act) {
    case AX2D_ACT_RELU: return v > 0.f ? v : 0.f;
    case AX2D_ACT_LEAKYRELU: return v > 0.f ? v : 0.01f * v;
    case AX2D_ACT_ELU: return v > 0.f ? v : expm1f(v);
    case AX2D_ACT_GELU: return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));
    case AX2D_ACT_SILU: return v / (1.f + expf(-v));
    default: return v;
  }
}
__device__ __forceinline__ float act_bwd(int act, float v) {   // d act / d v at pre-activation v
  switch (act) {
    case AX2D_ACT_RELU: return v > 0.f ? 1.f : 0.f;
    case AX2D_ACT_LEAKYRELU: return v > 0.f ? 1.f : 0.01f;
    case AX2D_ACT_ELU: return v > 0.f ? 1.f : expf(v);
    case AX2D_ACT_GELU: {
      const float cdf = 0.5f * (1.f + erff(v * 0.70710678118654752440f));
      const float pdf = 0.39894228040143267794f * expf(-0.5f * v * v);
      return cdf + v * pdf;
    }
    case AX2D_ACT_SILU: {
      const float s = 1.f / (1.f + expf(-v));
      return s * (1.f + v * (1.f - s));
    }
    default: return 1.f;
  }
}

// Counter-based dropout: keep-scale in {0, 1/(1-p)} from a 64-bit mix of (seed, element index).
// The same (seed, m, n) reproduces the same decision in the backward pass, so no mask is stored.
__device__ __forceinline__ float drop_scale(uint64_t seed, uint64_t idx, float p, float inv_keep) {
  uint64_t z = seed + idx * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  const float u = static_cast<float>(static_cast<uint32_t>(z >> 40)) * (1.0f / 16777216.0f);
  return u >= p ? inv_keep : 0.f;
}

}  // namespace ax2d
