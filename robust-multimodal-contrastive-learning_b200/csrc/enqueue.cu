// K3 — ring-buffer enqueue of the (gathered) keys into the negative queue, pointer on device.
//
// Replaces vilt/modules/objectives.py:244-248:
//   ptr = int(proj_queue_ptr)                       # D2H sync
//   proj_queue[:, ptr:ptr+B] = keys.T               # strided copy
//   proj_queue_ptr[0] = (ptr + B) % num_negative    # host modulo + H2D scalar write
// and MoCo/MoCo_RMCL.py:81-94.
//
// Layout: keys [B,C] row-major, queue [C,K] with K contiguous (vilt_module.py:92), so this is a
// transposing scatter: 32x32 tiles staged through padded shared memory, coalesced on both sides.
// Bytes: B*C*(sizeof(key)+sizeof(queue elem)) — 262 KB at cfg2, i.e. launch-latency bound; see
// DESIGN.md.  The pointer never leaves the device: every CTA reads the low word of *ptr_dev,
// then takes a ticket in the (otherwise zero) high word; the last CTA to arrive stores the
// advanced 64-bit pointer, which also clears the ticket.
//
// rmcl_enqueue_shadow writes the same columns into a second, bf16 copy of the queue in the same launch.
// The reference keeps the queue in fp32 (checkpoint format) and, under Lightning precision=16, lets
// autocast re-cast all C*K elements to half precision for every einsum (objectives.py:272,329).  Keeping
// a bf16 shadow current costs B*C*2 bytes per step here and gives the tcgen05 InfoNCE kernel its operand
// without that per-call conversion pass.
#include "common.cuh"

namespace rmcl {

template <typename TK, typename TQ>
__global__ void __launch_bounds__(256) enqueue_kernel(TQ* __restrict__ queue, const TK* __restrict__ keys,
                                                      long long* ptr_dev, int B, int C, long long K,
                                                      long long ldq, __nv_bfloat16* __restrict__ shadow,
                                                      long long lds) {
  __shared__ float tile[32][33];
  __shared__ long long s_ptr;
  unsigned int* ptr_words = reinterpret_cast<unsigned int*>(ptr_dev);
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_ptr = (long long)(*reinterpret_cast<volatile unsigned int*>(ptr_words));
  const int b0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int b = b0 + ty + 8 * i, c = c0 + tx;
    if (b < B && c < C) tile[ty + 8 * i][tx] = to_f32(keys[(long long)b * C + c]);
  }
  __syncthreads();
  const long long ptr = s_ptr;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = c0 + ty + 8 * i, b = b0 + tx;
    if (b < B && c < C) {
      long long col = ptr + b;
      if (col >= K) col -= K;  // only reachable when ptr is not a multiple of B
      const float v = tile[tx][ty + 8 * i];
      queue[(long long)c * ldq + col] = from_f32<TQ>(v);
      if (shadow) shadow[(long long)c * lds + col] = __float2bfloat16_rn(to_f32(from_f32<TQ>(v)));   // == bf16(queue element)
    }
  }
  if (threadIdx.x == 0) {
    const unsigned int total = gridDim.x * gridDim.y;
    const unsigned int ticket = atomicAdd(ptr_words + 1, 1u);
    if (ticket == total - 1) {
      // everyone has read the old pointer; publish the new one (high word back to zero)
      *reinterpret_cast<volatile long long*>(ptr_dev) = (ptr + B) % K;
    }
  }
}

}  // namespace rmcl

static int enqueue_impl(void* queue, rmcl_dtype queue_dtype, const void* keys, rmcl_dtype keys_dtype, int64_t* ptr_dev,
                        int B, int C, int64_t K, int64_t ldq, void* shadow, int64_t lds, void* stream) {
  RMCL_CHECK_ARG(queue && keys && ptr_dev, "rmcl_enqueue: null pointer");
  RMCL_CHECK_ARG(B > 0 && C > 0 && K > 0 && K < (1ll << 31), "rmcl_enqueue: bad sizes B=%d C=%d K=%lld", B, C,
                 (long long)K);
  RMCL_CHECK_ARG(ldq >= K, "rmcl_enqueue: ldq < K");
  RMCL_CHECK_ARG(!shadow || lds >= K, "rmcl_enqueue_shadow: lds < K");
  RMCL_CHECK_ARG(rmcl::dtype_ok(queue_dtype) && rmcl::dtype_ok(keys_dtype), "rmcl_enqueue: bad dtype");
  RMCL_CHECK_ARG(B <= K && K % B == 0, "rmcl_enqueue: queue length %lld is not a multiple of the batch %d",
                 (long long)K, B);
  RMCL_CHECK_ARG((reinterpret_cast<uintptr_t>(ptr_dev) & 7u) == 0, "rmcl_enqueue: ptr_dev not 8-byte aligned");
  dim3 grid((B + 31) / 32, (C + 31) / 32);
  cudaStream_t s = (cudaStream_t)stream;
  long long* p = reinterpret_cast<long long*>(ptr_dev);
  using bf16 = __nv_bfloat16;
  bf16* sh = reinterpret_cast<bf16*>(shadow);
  if (keys_dtype == RMCL_F32 && queue_dtype == RMCL_F32)
    rmcl::enqueue_kernel<float, float><<<grid, 256, 0, s>>>((float*)queue, (const float*)keys, p, B, C, K, ldq, sh, lds);
  else if (keys_dtype == RMCL_F32 && queue_dtype == RMCL_BF16)
    rmcl::enqueue_kernel<float, bf16><<<grid, 256, 0, s>>>((bf16*)queue, (const float*)keys, p, B, C, K, ldq, sh, lds);
  else if (keys_dtype == RMCL_BF16 && queue_dtype == RMCL_F32)
    rmcl::enqueue_kernel<bf16, float><<<grid, 256, 0, s>>>((float*)queue, (const bf16*)keys, p, B, C, K, ldq, sh, lds);
  else
    rmcl::enqueue_kernel<bf16, bf16><<<grid, 256, 0, s>>>((bf16*)queue, (const bf16*)keys, p, B, C, K, ldq, sh, lds);
  RMCL_LAUNCH_OK("enqueue_kernel");
  return RMCL_OK;
}

extern "C" int rmcl_enqueue(void* queue, rmcl_dtype queue_dtype, const void* keys, rmcl_dtype keys_dtype,
                            int64_t* ptr_dev, int B, int C, int64_t K, int64_t ldq, void* stream) {
  return enqueue_impl(queue, queue_dtype, keys, keys_dtype, ptr_dev, B, C, K, ldq, nullptr, 0, stream);
}

extern "C" int rmcl_enqueue_shadow(void* queue, rmcl_dtype queue_dtype, void* shadow_bf16, int64_t lds, const void* keys,
                                   rmcl_dtype keys_dtype, int64_t* ptr_dev, int B, int C, int64_t K, int64_t ldq,
                                   void* stream) {
  RMCL_CHECK_ARG(shadow_bf16 != nullptr, "rmcl_enqueue_shadow: null shadow");
  return enqueue_impl(queue, queue_dtype, keys, keys_dtype, ptr_dev, B, C, K, ldq, shadow_bf16, lds, stream);
}
