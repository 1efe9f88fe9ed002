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
//   colnorm   one thread per 4 consecutive columns, marching down the C rows: coalesced 8/16-byte loads
//   rowsum    one CTA per row c: sum_j queue[c][j] and sum_j queue[c][j]/max(|queue_j|, eps); the queue
//             (33-67 MB) was just read by colnorm and is served by L2; fixed-shape reduction tree
#include "common.cuh"

namespace rmcl {

template <typename TQ>
__global__ void __launch_bounds__(256) queue_colnorm_kernel(const TQ* __restrict__ queue, int C, long long K, long long ldq,
                                                            float* __restrict__ colnorm2) {
  const long long j0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (j0 >= K) return;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  const bool vec = (j0 + 3 < K) && (ldq % 4 == 0) && ((reinterpret_cast<uintptr_t>(queue) & 15u) == 0);
  if (vec) {
    for (int c = 0; c < C; ++c) {
      const TQ* p = queue + (size_t)c * ldq + j0;
      float v0, v1, v2, v3;
      if (sizeof(TQ) == 4) {
        const float4 u = __ldg(reinterpret_cast<const float4*>(p));
        v0 = u.x; v1 = u.y; v2 = u.z; v3 = u.w;
      } else {
        const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
        const __nv_bfloat162 lo = *reinterpret_cast<const __nv_bfloat162*>(&u.x);
        const __nv_bfloat162 hi = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
        v0 = __low2float(lo); v1 = __high2float(lo); v2 = __low2float(hi); v3 = __high2float(hi);
      }
      a0 = fmaf(v0, v0, a0); a1 = fmaf(v1, v1, a1); a2 = fmaf(v2, v2, a2); a3 = fmaf(v3, v3, a3);
    }
    *reinterpret_cast<float4*>(colnorm2 + j0) = make_float4(a0, a1, a2, a3);
  } else {
    for (int u = 0; u < 4 && j0 + u < K; ++u) {
      float a = 0.f;
      for (int c = 0; c < C; ++c) {
        const float v = to_f32(queue[(size_t)c * ldq + j0 + u]);
        a = fmaf(v, v, a);
      }
      colnorm2[j0 + u] = a;
    }
  }
}

template <typename TQ>
__global__ void __launch_bounds__(256) queue_rowsum_kernel(const TQ* __restrict__ queue, long long K, long long ldq,
                                                           const float* __restrict__ colnorm2, float eps,
                                                           float* __restrict__ sum_vec, float* __restrict__ sum_unit) {
  __shared__ float red[2][8];
  const int c = blockIdx.x;
  const TQ* row = queue + (size_t)c * ldq;
  float s0 = 0.f, s1 = 0.f;
  for (long long j = threadIdx.x; j < K; j += 256) {
    const float v = to_f32(row[j]);
    s0 += v;
    s1 = fmaf(v, __fdiv_rn(1.f, fmaxf(sqrtf(colnorm2[j]), eps)), s1);
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
    for (int w = 0; w < 8; ++w) {
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
  const unsigned grid = (unsigned)((K + 4 * 256 - 1) / (4 * 256));
  if (queue_dtype == RMCL_F32) {
    rmcl::queue_colnorm_kernel<float><<<grid, 256, 0, s>>>((const float*)queue, C, K, ldq, colnorm2);
    RMCL_LAUNCH_OK("queue_colnorm_kernel");
    rmcl::queue_rowsum_kernel<float><<<C, 256, 0, s>>>((const float*)queue, K, ldq, colnorm2, cos_eps, sum_vec, sum_unit);
  } else {
    using bf16 = __nv_bfloat16;
    rmcl::queue_colnorm_kernel<bf16><<<grid, 256, 0, s>>>((const bf16*)queue, C, K, ldq, colnorm2);
    RMCL_LAUNCH_OK("queue_colnorm_kernel");
    rmcl::queue_rowsum_kernel<bf16><<<C, 256, 0, s>>>((const bf16*)queue, K, ldq, colnorm2, cos_eps, sum_vec, sum_unit);
  }
  RMCL_LAUNCH_OK("queue_rowsum_kernel");
  return RMCL_OK;
}
