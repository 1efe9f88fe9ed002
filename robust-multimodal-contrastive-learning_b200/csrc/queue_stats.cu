// Queue statistics for the per-view diagnostics of compute_moco_contrastive
// (vilt/modules/objectives.py:337-349, 300-312, 375-387).
//
// The reference loops over the B samples of a view in Python and, per sample, reduces the whole [K,C]
// queue three times (L2 distance, cosine, dot): 3*B passes over the queue per view.  Two of the three
// are linear in the queue:
//     mean_j  q^.queue_j                 = q^ . (sum_j queue_j) / K
//     mean_j  cos(q^, queue_j)           = q^/|q^| . (sum_j queue_j / max(|queue_j|, eps)) / K
// so they need two [C] vectors, computed here once per step; the third needs |queue_j|^2 per column
// (also computed here) plus the q^.queue_j the fused InfoNCE kernels hold anyway.
//
//   colnorm   CTA = 128 columns x all C rows: lane = 4 consecutive columns (8/16-byte loads), the 8 warps take rows
//             w, w+8, ... with four loads in flight each and are combined through shared memory in warp order
//             (K/128 CTAs: 512 at K = 65536, several per SM)
//   rowsum    one CTA per row c: sum_j queue[c][j] and sum_j queue[c][j]/max(|queue_j|, eps), 16-byte loads, four in flight;
//             the queue (33-67 MB) was just read by colnorm and is served by L2; fixed-shape reduction tree
// One reduction of the queue per training step: 33.5 MB at C128 K65536 fp32, i.e. ~15 us for both kernels (the round-1
// version — 64 CTAs marching down the rows one dependent load at a time, scalar row sums — took 120 us).
#include "common.cuh"

namespace rmcl {

template <typename TQ>
__device__ __forceinline__ float4 ld4(const TQ* p);
template <>
__device__ __forceinline__ float4 ld4<float>(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}
template <>
__device__ __forceinline__ float4 ld4<__nv_bfloat16>(const __nv_bfloat16* p) {
  const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
  const __nv_bfloat162 lo = *reinterpret_cast<const __nv_bfloat162*>(&u.x);
  const __nv_bfloat162 hi = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
  return make_float4(__low2float(lo), __high2float(lo), __low2float(hi), __high2float(hi));
}

template <typename TQ>
__global__ void __launch_bounds__(256) queue_colnorm_kernel(const TQ* __restrict__ queue, int C, long long K, long long ldq,
                                                            float* __restrict__ colnorm2) {
  __shared__ float4 part[8][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long long j0 = (long long)blockIdx.x * 128 + lane * 4;
  const bool vec = (ldq % 4 == 0) && ((reinterpret_cast<uintptr_t>(queue) & 15u) == 0);
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  if (vec && j0 + 3 < K) {
    int c = w;
    for (; c + 24 < C; c += 32) {          // four independent loads in flight
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = ld4<TQ>(queue + (size_t)(c + 8 * u) * ldq + j0);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        a.x = fmaf(v[u].x, v[u].x, a.x); a.y = fmaf(v[u].y, v[u].y, a.y);
        a.z = fmaf(v[u].z, v[u].z, a.z); a.w = fmaf(v[u].w, v[u].w, a.w);
      }
    }
    for (; c < C; c += 8) {
      const float4 v = ld4<TQ>(queue + (size_t)c * ldq + j0);
      a.x = fmaf(v.x, v.x, a.x); a.y = fmaf(v.y, v.y, a.y); a.z = fmaf(v.z, v.z, a.z); a.w = fmaf(v.w, v.w, a.w);
    }
  } else {
    float* af = reinterpret_cast<float*>(&a);
    for (int u = 0; u < 4; ++u)
      if (j0 + u < K)
        for (int c = w; c < C; c += 8) {
          const float v = to_f32(queue[(size_t)c * ldq + j0 + u]);
          af[u] = fmaf(v, v, af[u]);
        }
  }
  part[w][lane] = a;
  __syncthreads();
  if (w == 0) {
    float4 t = part[0][lane];
#pragma unroll
    for (int g = 1; g < 8; ++g) {          // warp order: deterministic
      const float4 o = part[g][lane];
      t.x += o.x; t.y += o.y; t.z += o.z; t.w += o.w;
    }
    if (j0 + 3 < K && (reinterpret_cast<uintptr_t>(colnorm2 + j0) & 15u) == 0) {
      *reinterpret_cast<float4*>(colnorm2 + j0) = t;
    } else {
      const float* tf = reinterpret_cast<const float*>(&t);
      for (int u = 0; u < 4; ++u)
        if (j0 + u < K) colnorm2[j0 + u] = tf[u];
    }
  }
}

constexpr int kRowsumThreads = 1024;   // one CTA per row and at most one row per SM: the bytes in flight have to come from the CTA

template <typename TQ>
__global__ void __launch_bounds__(kRowsumThreads) queue_rowsum_kernel(const TQ* __restrict__ queue, long long K, long long ldq,
                                                           const float* __restrict__ colnorm2, float eps,
                                                           float* __restrict__ sum_vec, float* __restrict__ sum_unit) {
  __shared__ float red[2][kRowsumThreads / 32];
  const int c = blockIdx.x;
  const TQ* row = queue + (size_t)c * ldq;
  float s0 = 0.f, s1 = 0.f;
  const bool vec = (ldq % 4 == 0) && ((reinterpret_cast<uintptr_t>(queue) & 15u) == 0) &&
                   ((reinterpret_cast<uintptr_t>(colnorm2) & 15u) == 0);
  const long long K4 = vec ? (K / 4) * 4 : 0;
  auto acc4 = [&](const float4& v, const float4& n) {
    s0 += (v.x + v.y) + (v.z + v.w);
    s1 = fmaf(v.x, __fdiv_rn(1.f, fmaxf(sqrtf(n.x), eps)), s1);
    s1 = fmaf(v.y, __fdiv_rn(1.f, fmaxf(sqrtf(n.y), eps)), s1);
    s1 = fmaf(v.z, __fdiv_rn(1.f, fmaxf(sqrtf(n.z), eps)), s1);
    s1 = fmaf(v.w, __fdiv_rn(1.f, fmaxf(sqrtf(n.w), eps)), s1);
  };
  long long j = (long long)threadIdx.x * 4;
  constexpr long long kStep = 4ll * kRowsumThreads;
  for (; j + 3 * kStep < K4; j += 4 * kStep) {     // four independent 16-byte loads of the row (+ four of the norms) in flight
    float4 v[4], n[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      v[u] = ld4<TQ>(row + j + u * kStep);
      n[u] = __ldg(reinterpret_cast<const float4*>(colnorm2 + j + u * kStep));
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) acc4(v[u], n[u]);
  }
  for (; j < K4; j += kStep) acc4(ld4<TQ>(row + j), __ldg(reinterpret_cast<const float4*>(colnorm2 + j)));
  for (long long t = K4 + threadIdx.x; t < K; t += kRowsumThreads) {
    const float v = to_f32(row[t]);
    s0 += v;
    s1 = fmaf(v, __fdiv_rn(1.f, fmaxf(sqrtf(colnorm2[t]), eps)), s1);
  }
  s0 = warp_sum(s0);
  s1 = warp_sum(s1);
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = s0;
    red[1][threadIdx.x >> 5] = s1;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float t0 = 0.f, t1 = 0.f;
    for (int w = 0; w < kRowsumThreads / 32; ++w) {
      t0 += red[0][w];
      t1 += red[1][w];
    }
    sum_vec[c] = t0;
    sum_unit[c] = t1;
  }
}

}  // namespace rmcl

extern "C" int rmcl_queue_stats(const void* queue, rmcl_dtype queue_dtype, int C, int64_t K, int64_t ldq, float cos_eps,
                                float* colnorm2, float* sum_vec, float* sum_unit, void* stream) {
  RMCL_CHECK_ARG(queue && colnorm2 && sum_vec && sum_unit, "rmcl_queue_stats: null pointer");
  RMCL_CHECK_ARG(C > 0 && K > 0 && ldq >= K, "rmcl_queue_stats: bad sizes C=%d K=%lld ldq=%lld", C, (long long)K,
                 (long long)ldq);
  RMCL_CHECK_ARG(rmcl::dtype_ok(queue_dtype), "rmcl_queue_stats: bad dtype");
  RMCL_CHECK_ARG((reinterpret_cast<uintptr_t>(colnorm2) & 15u) == 0, "rmcl_queue_stats: colnorm2 must be 16B aligned");
  cudaStream_t s = (cudaStream_t)stream;
  const unsigned grid = (unsigned)((K + 127) / 128);
  if (queue_dtype == RMCL_F32) {
    rmcl::queue_colnorm_kernel<float><<<grid, 256, 0, s>>>((const float*)queue, C, K, ldq, colnorm2);
    RMCL_LAUNCH_OK("queue_colnorm_kernel");
    rmcl::queue_rowsum_kernel<float><<<C, rmcl::kRowsumThreads, 0, s>>>((const float*)queue, K, ldq, colnorm2, cos_eps, sum_vec, sum_unit);
  } else {
    using bf16 = __nv_bfloat16;
    rmcl::queue_colnorm_kernel<bf16><<<grid, 256, 0, s>>>((const bf16*)queue, C, K, ldq, colnorm2);
    RMCL_LAUNCH_OK("queue_colnorm_kernel");
    rmcl::queue_rowsum_kernel<bf16><<<C, rmcl::kRowsumThreads, 0, s>>>((const bf16*)queue, K, ldq, colnorm2, cos_eps, sum_vec, sum_unit);
  }
  RMCL_LAUNCH_OK("queue_rowsum_kernel");
  return RMCL_OK;
}
