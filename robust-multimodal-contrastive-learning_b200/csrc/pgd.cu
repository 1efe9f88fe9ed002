// K4 — fused PGD perturbation update on delta[B,N] given grad[B,N].
//
// Replaces attack/pgd_attack_vilt.py:162-173:
//   g = grad.clone().detach().float()
//   d = clamp(norm(g.view(B,-1), dim=1, p=inf), min=1e-8)
//   delta = (delta + (lr*g/d).to(delta)).detach(); if eps > 0: delta = clamp(delta, -eps, eps)
// (7 kernels / 5 full passes in eager torch).  REF_LINF reproduces that arithmetic operation by
// operation — fadd(delta, fdiv(fmul(lr,g), d)) — so fp32 results are bit-identical.
// SIGN_LINF and L2 are the north-star's extra modes (specified in DESIGN.md).
//
// Two implementations, both HBM-bound, selected by size:
//  * cluster path (pgd_cluster_kernel): one thread-block cluster per sample.  Each CTA stages its
//    slice of the gradient in shared memory with one pass over HBM, the per-sample norm is
//    reduced warp-shuffle -> CTA -> cluster (distributed shared memory), and the update runs out
//    of shared memory: 12 B/element (read g, read delta, write delta) — the algorithmic minimum.
//  * streaming path (norm kernel + update kernel [+ projection kernel for L2]): any N; re-reads
//    the gradient (16 B/element; from L2 when B*N*4 fits the 126 MB L2).
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace rmcl {

constexpr int kPgdThreads = 256;

__device__ __forceinline__ float block_reduce(float v, bool is_max, float* red /*[32]*/) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = is_max ? warp_max(v) : warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float r = (lane < nw) ? red[lane] : 0.f;
  r = is_max ? warp_max(r) : warp_sum(r);
  return r;  // valid in every thread
}

// One element of the update, given the per-sample denominator.
template <typename TD>
__device__ __forceinline__ float pgd_apply(float delta, float g, float lr, float denom, int mode);
template <>
__device__ __forceinline__ float pgd_apply<float>(float delta, float g, float lr, float denom, int mode) {
  float step;
  if (mode == RMCL_PGD_SIGN_LINF) {
    const float sg = (g > 0.f) ? 1.f : ((g < 0.f) ? -1.f : (g == g ? 0.f : g));
    step = __fmul_rn(lr, sg);
  } else {
    step = __fdiv_rn(__fmul_rn(lr, g), denom);
  }
  return __fadd_rn(delta, step);
}
template <>
__device__ __forceinline__ float pgd_apply<__nv_bfloat16>(float delta, float g, float lr, float denom, int mode) {
  // (lr*g/d).to(delta) rounds the step to bf16 first, the add rounds again
  const float step = __bfloat162float(__float2bfloat16_rn(pgd_apply<float>(0.f, g, lr, denom, mode)));
  return __bfloat162float(__float2bfloat16_rn(__fadd_rn(delta, step)));
}
template <typename TD> __device__ __forceinline__ float clamp_eps(float v, float eps);
template <> __device__ __forceinline__ float clamp_eps<float>(float v, float eps) {
  return fminf(fmaxf(v, -eps), eps);
}
template <> __device__ __forceinline__ float clamp_eps<__nv_bfloat16>(float v, float eps) {
  const float e = __bfloat162float(__float2bfloat16_rn(eps));  // ATen casts the bound to the tensor dtype
  return fminf(fmaxf(v, -e), e);
}

// ------------------------------------------------------------------ streaming path
// norms[b] (max|g| as uint bits, or sum g^2) accumulated with one atomic per CTA.
template <typename TG>
__global__ void __launch_bounds__(kPgdThreads) pgd_norm_kernel(const TG* __restrict__ grad, long long N, int mode,
                                                               float* __restrict__ norms) {
  __shared__ float red[32];
  const int b = blockIdx.y;
  const TG* g = grad + (long long)b * N;
  const bool is_max = (mode == RMCL_PGD_REF_LINF);
  float acc = 0.f;
  constexpr int VE = 16 / sizeof(TG);
  const bool vec = (N % VE == 0) && ((reinterpret_cast<uintptr_t>(grad) & 15u) == 0);
  if (vec) {
    const long long nv = N / VE;
    const uint4* gv = reinterpret_cast<const uint4*>(g);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (long long)gridDim.x * blockDim.x) {
      uint4 u = ld_stream_u4(gv + i);
      const TG* e = reinterpret_cast<const TG*>(&u);
#pragma unroll
      for (int j = 0; j < VE; ++j) {
        const float x = to_f32(e[j]);
        acc = is_max ? fmaxf(acc, fabsf(x)) : fmaf(x, x, acc);
      }
    }
  } else {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
      const float x = to_f32(g[i]);
      acc = is_max ? fmaxf(acc, fabsf(x)) : fmaf(x, x, acc);
    }
  }
  acc = block_reduce(acc, is_max, red);
  if (threadIdx.x == 0) {
    if (is_max) atomicMax(reinterpret_cast<unsigned int*>(norms + b), __float_as_uint(acc));
    else atomicAdd(norms + b, acc);
  }
}

template <typename TD, typename TG>
__global__ void __launch_bounds__(kPgdThreads) pgd_update_kernel(TD* __restrict__ delta, const TG* __restrict__ grad,
                                                                 long long N, float lr, float eps, int mode,
                                                                 const float* __restrict__ norms,
                                                                 float* __restrict__ dnorm2) {
  __shared__ float red[32];
  const int b = blockIdx.y;
  TD* d = delta + (long long)b * N;
  const TG* g = grad + (long long)b * N;
  float denom = 1.f;
  if (mode == RMCL_PGD_REF_LINF) denom = fmaxf(norms[b], 1e-8f);
  else if (mode == RMCL_PGD_L2) denom = fmaxf(sqrtf(norms[b]), 1e-8f);
  const bool do_clamp = (eps > 0.f) && (mode != RMCL_PGD_L2);
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
    float v = pgd_apply<TD>(to_f32(d[i]), to_f32(g[i]), lr, denom, mode);
    if (do_clamp) v = clamp_eps<TD>(v, eps);
    d[i] = from_f32<TD>(v);
    acc = fmaf(v, v, acc);
  }
  if (mode == RMCL_PGD_L2 && eps > 0.f) {
    acc = block_reduce(acc, false, red);
    if (threadIdx.x == 0) atomicAdd(dnorm2 + b, acc);
  }
}

template <typename TD>
__global__ void __launch_bounds__(kPgdThreads) pgd_project_l2_kernel(TD* __restrict__ delta, long long N, float eps,
                                                                     const float* __restrict__ dnorm2) {
  const int b = blockIdx.y;
  const float dn = fmaxf(sqrtf(dnorm2[b]), 1e-12f);
  const float s = fminf(__fdiv_rn(eps, dn), 1.f);
  if (s >= 1.f) return;
  TD* d = delta + (long long)b * N;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x)
    d[i] = from_f32<TD>(__fmul_rn(to_f32(d[i]), s));
}

// ------------------------------------------------------------------ cluster path
// grid = (cluster_size, B), cluster dims (cluster_size,1,1): cluster y == sample.
// dynamic smem: slice_elems floats (the CTA's slice of g as fp32; reused for delta' in L2 mode).
template <typename T>
__device__ __forceinline__ void unpack_to_smem(const uint4& u, float* dst, float& acc, bool is_max) {
  constexpr int V = 16 / sizeof(T);
  const T* e = reinterpret_cast<const T*>(&u);
  float x[V];
#pragma unroll
  for (int j = 0; j < V; ++j) {
    x[j] = to_f32(e[j]);
    acc = is_max ? fmaxf(acc, fabsf(x[j])) : fmaf(x[j], x[j], acc);
  }
#pragma unroll
  for (int j = 0; j < V; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(x[j], x[j + 1], x[j + 2], x[j + 3]);
}

template <typename TD, typename TG>
__global__ void __launch_bounds__(kPgdThreads) pgd_cluster_kernel(TD* __restrict__ delta, const TG* __restrict__ grad,
                                                                  long long N, long long slice_elems, float lr,
                                                                  float eps, int mode) {
  extern __shared__ __align__(16) float sg[];
  __shared__ float red[32];
  __shared__ float partial[2];  // [0]: norm of g, [1]: norm of delta'  (read by peer CTAs over DSMEM)
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned rank = cluster.block_rank(), csz = cluster.num_blocks();
  const int b = blockIdx.y;
  const long long lo = (long long)rank * slice_elems;
  long long n = N - lo;
  n = n < 0 ? 0 : (n > slice_elems ? slice_elems : n);
  const TG* g = grad + (long long)b * N + lo;
  TD* d = delta + (long long)b * N + lo;
  const bool is_max = (mode == RMCL_PGD_REF_LINF);
  constexpr int VG = 16 / sizeof(TG), VD = 16 / sizeof(TD);
  // slice_elems % 64 == 0, so N % 8 == 0 makes every sample/slice start 16B-aligned and n % 8 == 0
  const bool vec = (N % 8 == 0) &&
                   (((reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(delta)) & 15u) == 0);

  // pass 1: HBM -> smem (as fp32), 4 independent 16-byte loads in flight per thread; norm on the fly
  float acc = 0.f;
  if (vec) {
    const uint4* gv = reinterpret_cast<const uint4*>(g);
    const long long nv = n / VG;
    long long i = threadIdx.x;
    for (; i + 3 * kPgdThreads < nv; i += 4 * kPgdThreads) {
      uint4 u[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) u[k] = ld_stream_u4(gv + i + k * kPgdThreads);
#pragma unroll
      for (int k = 0; k < 4; ++k) unpack_to_smem<TG>(u[k], sg + (i + k * kPgdThreads) * VG, acc, is_max);
    }
    for (; i < nv; i += kPgdThreads) unpack_to_smem<TG>(ld_stream_u4(gv + i), sg + i * VG, acc, is_max);
  } else {
    for (long long i = threadIdx.x; i < n; i += kPgdThreads) {
      const float x = to_f32(g[i]);
      sg[i] = x;
      acc = is_max ? fmaxf(acc, fabsf(x)) : fmaf(x, x, acc);
    }
  }
  float denom = 1.f;
  if (mode != RMCL_PGD_SIGN_LINF) {
    acc = block_reduce(acc, is_max, red);
    if (threadIdx.x == 0) partial[0] = acc;
    cluster.sync();
    float tot = 0.f;
    for (unsigned r = 0; r < csz; ++r) {  // same order in every CTA -> identical denominators
      const float p = *cluster.map_shared_rank(&partial[0], r);
      tot = is_max ? fmaxf(tot, p) : tot + p;
    }
    denom = fmaxf(is_max ? tot : sqrtf(tot), 1e-8f);
  } else {
    __syncthreads();
  }

  // pass 2: update out of smem; delta is read and written exactly once
  const bool l2proj = (mode == RMCL_PGD_L2) && (eps > 0.f);
  const bool do_clamp = (eps > 0.f) && (mode != RMCL_PGD_L2);
  float acc2 = 0.f;
  if (vec) {
    uint4* dv = reinterpret_cast<uint4*>(d);
    const long long nv = n / VD;
    for (long long i = threadIdx.x; i < nv; i += kPgdThreads) {
      uint4 u = ld_u4(dv + i);
      TD* e = reinterpret_cast<TD*>(&u);
      float gs[VD];
#pragma unroll
      for (int j = 0; j < VD; j += 4) {
        const float4 t = *reinterpret_cast<const float4*>(sg + i * VD + j);
        gs[j] = t.x; gs[j + 1] = t.y; gs[j + 2] = t.z; gs[j + 3] = t.w;
      }
#pragma unroll
      for (int j = 0; j < VD; ++j) {
        float v = pgd_apply<TD>(to_f32(e[j]), gs[j], lr, denom, mode);
        if (do_clamp) v = clamp_eps<TD>(v, eps);
        gs[j] = v;
        acc2 = fmaf(v, v, acc2);
        e[j] = from_f32<TD>(v);
      }
      if (l2proj) {
#pragma unroll
        for (int j = 0; j < VD; j += 4)
          *reinterpret_cast<float4*>(sg + i * VD + j) = make_float4(gs[j], gs[j + 1], gs[j + 2], gs[j + 3]);
      } else {
        st_stream_u4(dv + i, u);
      }
    }
  } else {
    for (long long i = threadIdx.x; i < n; i += kPgdThreads) {
      float v = pgd_apply<TD>(to_f32(d[i]), sg[i], lr, denom, mode);
      if (do_clamp) v = clamp_eps<TD>(v, eps);
      acc2 = fmaf(v, v, acc2);
      if (l2proj) sg[i] = v;
      else d[i] = from_f32<TD>(v);
    }
  }
  if (l2proj) {
    acc2 = block_reduce(acc2, false, red);
    if (threadIdx.x == 0) partial[1] = acc2;
    cluster.sync();
    float tot = 0.f;
    for (unsigned r = 0; r < csz; ++r) tot += *cluster.map_shared_rank(&partial[1], r);
    const float s = fminf(__fdiv_rn(eps, fmaxf(sqrtf(tot), 1e-12f)), 1.f);
    for (long long i = threadIdx.x; i < n; i += kPgdThreads) {
      const float v = sg[i];
      d[i] = from_f32<TD>(s < 1.f ? __fmul_rn(v, s) : v);
    }
  }
  cluster.sync();  // keep partial[] alive until every peer has read it
}

template <typename TD, typename TG>
static int launch_pgd(void* delta, const void* grad, int B, long long N, float lr, float eps, int mode, float* ws,
                      cudaStream_t s) {
  TD* d = (TD*)delta;
  const TG* g = (const TG*)grad;
  const int sms = sm_count();
  if (sms <= 0) return RMCL_E_CUDA;

  // ---- cluster path: smallest cluster whose per-CTA slice fits ~100 KB (2 CTAs/SM), else up to 200 KB
  int csz = 0;
  long long slice = 0;
  for (int c = 1; c <= 16; c *= 2) {
    long long sl = ((N + c - 1) / c + 63) / 64 * 64;
    if (sl * 4 <= 100 * 1024) { csz = c; slice = sl; break; }
  }
  if (!csz) {
    long long sl = ((N + 15) / 16 + 63) / 64 * 64;
    if (sl * 4 <= 200 * 1024) { csz = 16; slice = sl; }
  }
  if (csz) {
    auto kern = pgd_cluster_kernel<TD, TG>;
    const size_t smem = (size_t)slice * 4;
    RMCL_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (csz > 8) RMCL_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(csz, B, 1);
    cfg.blockDim = dim3(kPgdThreads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = csz;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    RMCL_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, d, g, N, slice, lr, eps, mode));
    return RMCL_OK;
  }

  // ---- streaming path
  RMCL_CHECK_ARG(ws != nullptr, "rmcl_pgd_step: norms_ws is required for N=%lld", N);
  RMCL_CUDA_OK(cudaMemsetAsync(ws, 0, sizeof(float) * 2 * (size_t)B, s));
  long long bps = (N + (long long)kPgdThreads * 16 - 1) / ((long long)kPgdThreads * 16);
  long long cap = ((long long)sms * 8 + B - 1) / B;
  if (bps > cap) bps = cap;
  if (bps < 1) bps = 1;
  dim3 grid((unsigned)bps, B);
  if (mode != RMCL_PGD_SIGN_LINF) {
    pgd_norm_kernel<TG><<<grid, kPgdThreads, 0, s>>>(g, N, mode, ws);
    RMCL_LAUNCH_OK("pgd_norm_kernel");
  }
  pgd_update_kernel<TD, TG><<<grid, kPgdThreads, 0, s>>>(d, g, N, lr, eps, mode, ws, ws + B);
  RMCL_LAUNCH_OK("pgd_update_kernel");
  if (mode == RMCL_PGD_L2 && eps > 0.f) {
    pgd_project_l2_kernel<TD><<<grid, kPgdThreads, 0, s>>>(d, N, eps, ws + B);
    RMCL_LAUNCH_OK("pgd_project_l2_kernel");
  }
  return RMCL_OK;
}

}  // namespace rmcl

extern "C" int rmcl_pgd_step(void* delta, rmcl_dtype delta_dtype, const void* grad, rmcl_dtype grad_dtype, int B,
                             int64_t N, float lr, float eps, int mode, float* norms_ws, void* stream) {
  RMCL_CHECK_ARG(delta && grad, "rmcl_pgd_step: null pointer");
  RMCL_CHECK_ARG(B > 0 && N > 0 && B <= 65535, "rmcl_pgd_step: bad sizes B=%d N=%lld", B, (long long)N);
  RMCL_CHECK_ARG(mode >= RMCL_PGD_REF_LINF && mode <= RMCL_PGD_L2, "rmcl_pgd_step: bad mode %d", mode);
  RMCL_CHECK_ARG(rmcl::dtype_ok(delta_dtype) && rmcl::dtype_ok(grad_dtype), "rmcl_pgd_step: bad dtype");
  cudaStream_t s = (cudaStream_t)stream;
  using bf16 = __nv_bfloat16;
  if (delta_dtype == RMCL_F32 && grad_dtype == RMCL_F32)
    return rmcl::launch_pgd<float, float>(delta, grad, B, N, lr, eps, mode, norms_ws, s);
  if (delta_dtype == RMCL_F32 && grad_dtype == RMCL_BF16)
    return rmcl::launch_pgd<float, bf16>(delta, grad, B, N, lr, eps, mode, norms_ws, s);
  if (delta_dtype == RMCL_BF16 && grad_dtype == RMCL_F32)
    return rmcl::launch_pgd<bf16, float>(delta, grad, B, N, lr, eps, mode, norms_ws, s);
  return rmcl::launch_pgd<bf16, bf16>(delta, grad, B, N, lr, eps, mode, norms_ws, s);
}
