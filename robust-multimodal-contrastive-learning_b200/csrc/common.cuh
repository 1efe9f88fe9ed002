// Shared helpers for the rmcl_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/rmcl_b200.h"

namespace rmcl {

// ---- per-thread error string -------------------------------------------------------------
void set_error(const char* fmt, ...);

#define RMCL_CHECK_ARG(cond, ...)          \
  do {                                     \
    if (!(cond)) {                         \
      rmcl::set_error(__VA_ARGS__);        \
      return RMCL_E_BADARG;                \
    }                                      \
  } while (0)

#define RMCL_CUDA_OK(expr)                                                                  \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      rmcl::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return RMCL_E_CUDA;                                                                   \
    }                                                                                       \
  } while (0)

#define RMCL_LAUNCH_OK(name)                                                     \
  do {                                                                           \
    cudaError_t _e = cudaGetLastError();                                         \
    if (_e != cudaSuccess) {                                                     \
      rmcl::set_error("launch of %s failed: %s", name, cudaGetErrorString(_e));  \
      return RMCL_E_CUDA;                                                        \
    }                                                                            \
  } while (0)

int sm_count();  // cached per device; <=0 on failure

inline size_t dtype_size(int dt) { return dt == RMCL_BF16 ? 2 : 4; }
inline bool dtype_ok(int dt) { return dt == RMCL_F32 || dt == RMCL_BF16; }

// ---- element access ------------------------------------------------------------------------
__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) {
  return __float2bfloat16_rn(v);
}

// 128-bit streaming accesses (read-once / write-once data: bypass L1 allocation).
__device__ __forceinline__ uint4 ld_stream_u4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
// plain 128-bit load for data that is also written by this kernel (no .nc)
__device__ __forceinline__ uint4 ld_u4(const void* p) {
  uint4 r;
  asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream_u4(void* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w)
               : "memory");
}

// Programmatic dependent launch: a kernel launched with launch_pdl() may start while its
// predecessor in the stream is still running; it must call pdl_wait() before touching anything the
// predecessor writes.  pdl_trigger() in the predecessor lets the dependent's CTAs become resident early.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// Every kernel of the step is launched with the programmatic-dependent-launch attribute and starts with pdl_wait(): its
// launch is processed while the kernel in front of it still runs, and it touches nothing before that kernel has completed
// (cfg2 step 256.7 -> 250.5 us).  RMCL_PDL_EARLY_TRIGGER bits say which kernels ALSO let their successor's CTAs become
// resident before they end: 1 the EMA, from each CTA's last chunk on — the InfoNCE prep and flash kernels then set up (TMEM,
// barriers, the first queue tiles) in the slots the EMA's tail frees: 250.5 -> 239.5 us per step, 3480 -> 3590 steps/s end
// to end; 2 the enqueue and 4 the finalize kernel at entry — neutral alone, harmful together (263 us: the next step's EMA
// grid then becomes resident and idles beside the finalize kernel).  profiles/r2_pdl_experiments.txt.
#ifndef RMCL_PDL_EARLY_TRIGGER
#define RMCL_PDL_EARLY_TRIGGER 1
#endif

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                              Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// L2 eviction-priority policies for data with a known reuse distance
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint4 ld_u4_hint(const void* p, uint64_t pol) {
  uint4 r;
  asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ void st_u4_hint(void* p, const uint4& v, uint64_t pol) {
  asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.u32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w), "l"(pol)
               : "memory");
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace rmcl
