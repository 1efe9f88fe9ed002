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

// x = the value the queue now holds.  Plane 0 (rows [0,C)): bf16(x), the operand of the bf16 tcgen05 InfoNCE;
// plane 1 (rows [C,2C), RMCL_BF16_HILO): bf16(x - hi), the low half of the split-operand fp32-accurate path.
__device__ __forceinline__ void write_shadow(__nv_bfloat16* __restrict__ shadow, long long lds, int C, int planes, int c,
                                             long long col, float x) {
  const __nv_bfloat16 hi = __float2bfloat16_rn(x);
  shadow[(long long)c * lds + col] = hi;
  if (planes == 2) shadow[(long long)(C + c) * lds + col] = __float2bfloat16_rn(x - __bfloat162float(hi));
}

// One pass over an fp32 queue: the hi/lo planes of RMCL_BF16_HILO (rmcl_queue_split).
__global__ void __launch_bounds__(256) queue_split_kernel(const float* __restrict__ queue, int C, long long K, long long ldq,
                                                          __nv_bfloat16* __restrict__ hilo, long long ldh) {
  const long long n4 = K / 4;     // K % 8 == 0 is checked by the caller; rows are 16-byte aligned
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4 * C; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i / n4);
    const long long j = (i - (long long)c * n4) * 4;
    const float4 v = __ldcs(reinterpret_cast<const float4*>(queue + (long long)c * ldq + j));
    const float x[4] = {v.x, v.y, v.z, v.w};
    __nv_bfloat16 h[4], l[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      h[u] = __float2bfloat16_rn(x[u]);
      l[u] = __float2bfloat16_rn(x[u] - __bfloat162float(h[u]));
    }
    *reinterpret_cast<uint2*>(hilo + (long long)c * ldh + j) = *reinterpret_cast<const uint2*>(h);
    *reinterpret_cast<uint2*>(hilo + (long long)(C + c) * ldh + j) = *reinterpret_cast<const uint2*>(l);
  }
}

template <typename TK, typename TQ>
__global__ void __launch_bounds__(256) enqueue_kernel(TQ* __restrict__ queue, const TK* __restrict__ keys,
                                                      long long* ptr_dev, int B, int C, long long K,
                                                      long long ldq, __nv_bfloat16* __restrict__ shadow,
                                                      long long lds, int planes) {
  __shared__ float tile[32][33];
  __shared__ long long s_ptr;
  unsigned int* ptr_words = reinterpret_cast<unsigned int*>(ptr_dev);
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#if RMCL_PDL_EARLY_TRIGGER & 2
  if (threadIdx.x == 0) pdl_trigger();   // a tiny grid: whatever follows may queue up behind it right away
#endif
  pdl_wait();   // programmatic dependent launch: the keys come from the kernel in front of this one
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
      if (shadow) write_shadow(shadow, lds, C, planes, c, col, to_f32(from_f32<TQ>(v)));
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
                        int B, int C, int64_t K, int64_t ldq, void* shadow, int64_t lds, int planes, void* stream) {
  RMCL_CHECK_ARG(queue && keys && ptr_dev, "rmcl_enqueue: null pointer");
  RMCL_CHECK_ARG(B > 0 && C > 0 && K > 0 && K < (1ll << 31), "rmcl_enqueue: bad sizes B=%d C=%d K=%lld", B, C,
                 (long long)K);
  RMCL_CHECK_ARG(ldq >= K, "rmcl_enqueue: ldq < K");
  RMCL_CHECK_ARG(!shadow || lds >= K, "rmcl_enqueue_shadow: lds < K");
  RMCL_CHECK_ARG(!shadow || planes == 1 || planes == 2, "rmcl_enqueue_shadow: shadow_planes must be 1 or 2 (got %d)", planes);
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
    RMCL_CUDA_OK(rmcl::launch_pdl(rmcl::enqueue_kernel<float, float>, grid, dim3(256), 0, s, (float*)queue, (const float*)keys, p, B, C, (long long)K, (long long)ldq, sh,
                                  (long long)lds, planes));
  else if (keys_dtype == RMCL_F32 && queue_dtype == RMCL_BF16)
    RMCL_CUDA_OK(rmcl::launch_pdl(rmcl::enqueue_kernel<float, bf16>, grid, dim3(256), 0, s, (bf16*)queue, (const float*)keys, p, B, C, (long long)K, (long long)ldq, sh,
                                  (long long)lds, planes));
  else if (keys_dtype == RMCL_BF16 && queue_dtype == RMCL_F32)
    RMCL_CUDA_OK(rmcl::launch_pdl(rmcl::enqueue_kernel<bf16, float>, grid, dim3(256), 0, s, (float*)queue, (const bf16*)keys, p, B, C, (long long)K, (long long)ldq, sh,
                                  (long long)lds, planes));
  else
    RMCL_CUDA_OK(rmcl::launch_pdl(rmcl::enqueue_kernel<bf16, bf16>, grid, dim3(256), 0, s, (bf16*)queue, (const bf16*)keys, p, B, C, (long long)K, (long long)ldq, sh,
                                  (long long)lds, planes));
  return RMCL_OK;
}

extern "C" int rmcl_enqueue(void* queue, rmcl_dtype queue_dtype, const void* keys, rmcl_dtype keys_dtype,
                            int64_t* ptr_dev, int B, int C, int64_t K, int64_t ldq, void* stream) {
  return enqueue_impl(queue, queue_dtype, keys, keys_dtype, ptr_dev, B, C, K, ldq, nullptr, 0, 1, stream);
}

extern "C" int rmcl_enqueue_shadow(void* queue, rmcl_dtype queue_dtype, void* shadow_bf16, int64_t lds, int shadow_planes,
                                   const void* keys, rmcl_dtype keys_dtype, int64_t* ptr_dev, int B, int C, int64_t K,
                                   int64_t ldq, void* stream) {
  RMCL_CHECK_ARG(shadow_bf16 != nullptr, "rmcl_enqueue_shadow: null shadow");
  return enqueue_impl(queue, queue_dtype, keys, keys_dtype, ptr_dev, B, C, K, ldq, shadow_bf16, lds, shadow_planes, stream);
}

extern "C" int rmcl_queue_split(const void* queue_f32, int C, int64_t K, int64_t ldq, void* hilo_bf16, int64_t ld_hilo,
                                void* stream) {
  RMCL_CHECK_ARG(queue_f32 && hilo_bf16, "rmcl_queue_split: null pointer");
  RMCL_CHECK_ARG(C > 0 && K > 0 && K % 8 == 0 && ldq >= K && ld_hilo >= K && ldq % 4 == 0 && ld_hilo % 8 == 0,
                 "rmcl_queue_split: bad sizes C=%d K=%lld ldq=%lld ld_hilo=%lld (K %% 8, ldq %% 4, ld_hilo %% 8 must be 0)", C,
                 (long long)K, (long long)ldq, (long long)ld_hilo);
  RMCL_CHECK_ARG(((reinterpret_cast<uintptr_t>(queue_f32) | reinterpret_cast<uintptr_t>(hilo_bf16)) & 15u) == 0,
                 "rmcl_queue_split: pointers must be 16-byte aligned");
  const int sms = rmcl::sm_count();
  if (sms <= 0) return RMCL_E_CUDA;
  rmcl::queue_split_kernel<<<sms * 8, 256, 0, (cudaStream_t)stream>>>((const float*)queue_f32, C, K, ldq,
                                                                       (__nv_bfloat16*)hilo_bf16, ld_hilo);
  RMCL_LAUNCH_OK("queue_split_kernel");
  return RMCL_OK;
}

// ------------------------------------------------------------------------------------------------------------
// Fused key exchange + enqueue over NVLink peer memory (single node): replaces the pair
//   ncclAllGather(keys) -> enqueue_kernel
// (objectives.py:226-235 + 244-248) by ONE launch per rank.  Every rank owns a two-slot staging buffer
// [2][world*B][C] in peer-mapped (symmetric) memory and a flag word:
//   1. push    my keys -> slot (epoch & 1), rows [rank*B, (rank+1)*B) of EVERY rank's staging buffer
//              (16-byte stores straight into the peers' HBM over NVLink / NVSwitch);
//   2. signal  when all my CTAs have pushed (local arrival counter), one system-scope release add on every
//              rank's flag;
//   3. wait    until my own flag shows epoch*world arrivals (acquire, system scope): all keys of this step are here
//              (epoch = number of this call, kept in the rank's own flag words);
//   4. enqueue the world*B staged keys into my replica of the queue (+ bf16 shadow) — the transposing scatter of
//              enqueue_kernel — and advance my pointer.
// Two slots suffice: a peer can push step n+2 into slot n&1 only after it has seen my signal of step n+1, which I
// send after my step-n enqueue has finished reading that slot.  All CTAs are co-resident (grid = RMCL_P2P_MAX_CTAS = 32),
// so the spin in step 3 cannot starve the pushes it waits for.
#ifndef RMCL_P2P_MAX_CTAS
#define RMCL_P2P_MAX_CTAS 32
#endif
namespace rmcl {

__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_sys_add(unsigned int* p, unsigned int v) {
  asm volatile("red.release.sys.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <typename TQ>
__global__ void __launch_bounds__(256) gather_enqueue_p2p_kernel(float* const* __restrict__ stage_ptrs,
                                                                 unsigned int* const* __restrict__ flag_ptrs,
                                                                 const float* __restrict__ keys_local, TQ* __restrict__ queue,
                                                                 long long* ptr_dev, int rank, int world, int B, int C,
                                                                 long long K, long long ldq, __nv_bfloat16* __restrict__ shadow,
                                                                 long long lds, int planes) {
  __shared__ float wtile[8][32][33];
  __shared__ long long s_ptr;
  __shared__ unsigned int s_epoch;
  unsigned int* ptr_words = reinterpret_cast<unsigned int*>(ptr_dev);
  const int tx = threadIdx.x & 31;
  const int Bt = world * B;
  unsigned int* my_flags = flag_ptrs[rank];      // [0] arrivals from all ranks, [1] local CTA counter, [2] completed calls
  if (threadIdx.x == 0) {
    s_ptr = (long long)(*reinterpret_cast<volatile unsigned int*>(ptr_words));
    // the call number lives on the device (advanced by the last CTA at the very end, after every CTA has read it), so the
    // launch carries no host-side counter and can be captured in a CUDA graph
    s_epoch = *reinterpret_cast<volatile unsigned int*>(my_flags + 2) + 1u;
  }
  __syncthreads();
  const unsigned int epoch = s_epoch;
  const size_t slot = (size_t)(epoch & 1u) * Bt * C;

  // 1. push
  const size_t n_local = (size_t)B * C;
  const size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, gthreads = (size_t)gridDim.x * blockDim.x;
  if ((n_local & 3) == 0) {
    const size_t nv = n_local / 4;
    const float4* src = reinterpret_cast<const float4*>(keys_local);
    for (size_t v = gtid; v < nv * world; v += gthreads) {
      const int peer = (int)(v / nv);
      const size_t e = v - (size_t)peer * nv;
      float4* dst = reinterpret_cast<float4*>(stage_ptrs[peer] + slot + (size_t)rank * n_local) + e;
      *dst = __ldg(src + e);
    }
  } else {
    for (size_t v = gtid; v < n_local * world; v += gthreads) {
      const int peer = (int)(v / n_local);
      const size_t e = v - (size_t)peer * n_local;
      stage_ptrs[peer][slot + (size_t)rank * n_local + e] = keys_local[e];
    }
  }
  // 2. signal (the last CTA of this rank to finish pushing tells everyone)
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    if (atomicAdd(my_flags + 1, 1u) == gridDim.x - 1u) {
      my_flags[1] = 0u;
      __threadfence_system();
      for (int peer = 0; peer < world; ++peer) red_release_sys_add(flag_ptrs[peer], 1u);
    }
    // 3. wait for every rank's keys of this step
    const unsigned int target = epoch * (unsigned int)world;
    const long long t0 = clock64();
    while (ld_acquire_sys(my_flags) < target) {
      if (clock64() - t0 > 20000000000ll) __trap();     // ~10 s: a missing peer becomes an error, not a hang
    }
  }
  __syncthreads();

  // 4. enqueue the staged keys: tiles of 32 keys x 32 channels, one tile per WARP at a time (the grid is small on purpose,
  //    so the parallelism has to come from inside the CTA: 8 warps x 32 loads in flight each; with one tile per CTA and two
  //    block-wide barriers per tile the 512 tiles of an 8-rank exchange took 38 us on 16 CTAs)
  const float* staged = stage_ptrs[rank] + slot;
  const long long ptr = s_ptr;
  const int tiles_b = (Bt + 31) / 32, tiles_c = (C + 31) / 32;
  const int warp = threadIdx.x >> 5, lane = tx;
  float (*wt)[33] = wtile[warp];
  for (int t = blockIdx.x * 8 + warp; t < tiles_b * tiles_c; t += gridDim.x * 8) {
    const int b0 = (t % tiles_b) * 32, c0 = (t / tiles_b) * 32;
    __syncwarp();
#pragma unroll 8
    for (int i = 0; i < 32; ++i) {
      const int b = b0 + i, c = c0 + lane;
      if (b < Bt && c < C) wt[i][lane] = __ldcg(staged + (size_t)b * C + c);   // written by peers: bypass L1
    }
    __syncwarp();
#pragma unroll 8
    for (int i = 0; i < 32; ++i) {
      const int c = c0 + i, b = b0 + lane;
      if (b < Bt && c < C) {
        long long col = ptr + b;
        if (col >= K) col -= K;
        const float v = wt[lane][i];
        queue[(long long)c * ldq + col] = from_f32<TQ>(v);
        if (shadow) write_shadow(shadow, lds, C, planes, c, col, to_f32(from_f32<TQ>(v)));
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int ticket = atomicAdd(ptr_words + 1, 1u);
    if (ticket == gridDim.x - 1u) {
      *reinterpret_cast<volatile unsigned int*>(my_flags + 2) = epoch;
      *reinterpret_cast<volatile long long*>(ptr_dev) = (ptr + Bt) % K;
    }
  }
}

}  // namespace rmcl

extern "C" int rmcl_gather_enqueue_p2p(void* const* stage_ptrs_dev, void* const* flag_ptrs_dev, const void* keys_local,
                                       void* queue, rmcl_dtype queue_dtype, void* shadow_bf16, int64_t lds, int shadow_planes,
                                       int64_t* ptr_dev, int rank, int world, int B_local, int C, int64_t K, int64_t ldq,
                                       void* stream) {
  RMCL_CHECK_ARG(stage_ptrs_dev && flag_ptrs_dev && keys_local && queue && ptr_dev, "rmcl_gather_enqueue_p2p: null pointer");
  RMCL_CHECK_ARG(world > 0 && rank >= 0 && rank < world && B_local > 0 && C > 0 && K > 0 && K < (1ll << 31) && ldq >= K,
                 "rmcl_gather_enqueue_p2p: bad sizes rank=%d world=%d B=%d C=%d K=%lld", rank, world, B_local, C, (long long)K);
  RMCL_CHECK_ARG(rmcl::dtype_ok(queue_dtype), "rmcl_gather_enqueue_p2p: bad dtype");
  const long long Bt = (long long)world * B_local;
  RMCL_CHECK_ARG(Bt <= K && K % Bt == 0, "rmcl_gather_enqueue_p2p: queue length %lld is not a multiple of the gathered batch %lld",
                 (long long)K, Bt);
  RMCL_CHECK_ARG(!shadow_bf16 || lds >= K, "rmcl_gather_enqueue_p2p: lds < K");
  RMCL_CHECK_ARG(!shadow_bf16 || shadow_planes == 1 || shadow_planes == 2, "rmcl_gather_enqueue_p2p: shadow_planes must be 1 or 2");
  const int sms = rmcl::sm_count();
  if (sms <= 0) return RMCL_E_CUDA;
  const long long tiles = ((Bt + 31) / 32) * ((C + 31) / 32);
  // All CTAs must be resident at once (they spin on the flag), and while they wait for the slowest rank they hold their
  // SM's thread slots away from whatever runs beside them — in the training step that is the HBM-bound EMA the exchange is
  // meant to hide under.  The work is tiny (world x 256 KB of pushes, world*B*C*6 bytes of scatter), so a small fixed grid
  // loses nothing: 32 CTAs move 8 x 256 KB in ~16 sweeps of 16-byte stores and scatter 8 tiles each at a time.  (Round 1 launched 2 CTAs per SM = 296 spinning
  // CTAs; 1 -> 8 GPU weak scaling of the cfg2 step was 0.93.)
  long long grid = RMCL_P2P_MAX_CTAS;
  if (grid > 2ll * sms) grid = 2ll * sms;
  if (grid > tiles) grid = tiles;
  cudaStream_t s = (cudaStream_t)stream;
  using bf16 = __nv_bfloat16;
  if (queue_dtype == RMCL_F32)
    rmcl::gather_enqueue_p2p_kernel<float><<<(unsigned)grid, 256, 0, s>>>(
        (float* const*)stage_ptrs_dev, (unsigned int* const*)flag_ptrs_dev, (const float*)keys_local, (float*)queue,
        reinterpret_cast<long long*>(ptr_dev), rank, world, B_local, C, K, ldq, (bf16*)shadow_bf16, lds, shadow_planes);
  else
    rmcl::gather_enqueue_p2p_kernel<bf16><<<(unsigned)grid, 256, 0, s>>>(
        (float* const*)stage_ptrs_dev, (unsigned int* const*)flag_ptrs_dev, (const float*)keys_local, (bf16*)queue,
        reinterpret_cast<long long*>(ptr_dev), rank, world, B_local, C, K, ldq, (bf16*)shadow_bf16, lds, shadow_planes);
  RMCL_LAUNCH_OK("gather_enqueue_p2p_kernel");
  return RMCL_OK;
}
