// K1 partial stage, fp32 CUDA-core version: split-K flash pass over the queue for any dtype/shape.
//
// This is the fp32 parity path (the reference runs InfoNCE in fp32 inside PGD, autocast disabled:
// attack/pgd_attack_vilt.py:141, and its queue buffer is fp32: vilt_module.py:92) and the path for
// shapes the tcgen05 kernel does not cover.  Output format: see infonce.cuh.
//
// CTA = 16 query rows x one contiguous range of queue columns, 256 threads, tiles of TK columns:
//   load   queue[:, k0:k0+TK] -> smem as fp32 [C][TK+1]   (coalesced along K, conflict-free pad)
//   S      thread (j, row-group): s = sum_c q^[r][c] * tile[c][j]      -> smem, log2 units
//   stats  16 lanes per row: tile max / argmax, running (m, l), alpha = 2^(m_old-m_new), P = 2^(S-m)
//   O      thread owns columns c = tid + 256*i, all 16 rows: acc = acc*alpha + sum_j P[r][j]*tile[c][j]
#include "infonce.cuh"

namespace rmcl {

constexpr int kSimtThreads = 256;
constexpr int kSimtRows = 16;
constexpr int kSimtMaxCPerThread = 4;  // C <= 1024

template <typename TQ, int TK>
__global__ void __launch_bounds__(kSimtThreads) infonce_simt_kernel(
    const float* __restrict__ q_hat, const TQ* __restrict__ queue, int B, int C, long long K, long long ldq,
    float scale2, bool bf16_mode, long long cols_per_split, float* __restrict__ pm, float* __restrict__ pl,
    float* __restrict__ pav, int* __restrict__ pai, float* __restrict__ po, const float* __restrict__ n2,
    const float* __restrict__ qn2, float* __restrict__ pdist, bool want_o) {
  // want_o == false (no gradient requested: the clean-query argmax call objectives.py:267-275, the greedy attack's
  // candidate losses): the P.queue^T accumulation — half of the arithmetic — and its write-out are skipped.
  extern __shared__ __align__(16) float smem[];
  float* qs = smem;                               // [16][C]
  float* tile = qs + kSimtRows * C;               // [C][TK+1]
  float* sp = tile + (((size_t)C * (TK + 1) + 3) & ~(size_t)3);  // [16][TK], 16B aligned
  float* s_m = sp + kSimtRows * TK;               // [16]
  float* s_l = s_m + kSimtRows;
  float* s_alpha = s_l + kSimtRows;
  float* s_av = s_alpha + kSimtRows;
  int* s_ai = reinterpret_cast<int*>(s_av + kSimtRows);
  float* s_qn2 = reinterpret_cast<float*>(s_ai + kSimtRows);   // diagnostics: |q^|^2 and the distance sums per row
  float* s_dist = s_qn2 + kSimtRows;

  const int tid = threadIdx.x;
  const int split = blockIdx.x;
  const int row0 = blockIdx.y * kSimtRows;
  const long long k_begin = (long long)split * cols_per_split;
  const long long k_end = (k_begin + cols_per_split < K) ? k_begin + cols_per_split : K;

  for (int i = tid; i < kSimtRows * C; i += kSimtThreads) {
    const int r = i / C, c = i - r * C;
    float v = (row0 + r < B) ? q_hat[(size_t)(row0 + r) * C + c] : 0.f;
    if (bf16_mode) v = __bfloat162float(__float2bfloat16_rn(v));
    qs[i] = v;
  }
  if (tid < kSimtRows) {
    s_m[tid] = -INFINITY;
    s_l[tid] = 0.f;
    s_av[tid] = -INFINITY;
    s_ai[tid] = 0;
    s_qn2[tid] = (n2 != nullptr && row0 + tid < B) ? qn2[row0 + tid] : 0.f;
    s_dist[tid] = 0.f;
  }
  float acc[kSimtMaxCPerThread][kSimtRows];
#pragma unroll
  for (int i = 0; i < kSimtMaxCPerThread; ++i)
#pragma unroll
    for (int r = 0; r < kSimtRows; ++r) acc[i][r] = 0.f;
  __syncthreads();

  constexpr int kGroups = kSimtThreads / TK;        // row groups in the S phase
  constexpr int kRpt = kSimtRows / kGroups;         // rows per thread in the S phase
  const int j1 = tid % TK, rg = tid / TK;
  const int srow = tid >> 4, ssub = tid & 15;       // stats phase: 16 lanes per row
  float dacc[kRpt];                                 // diagnostics: sum_j |q^_r - queue_j| over this thread's columns
#pragma unroll
  for (int r = 0; r < kRpt; ++r) dacc[r] = 0.f;

  for (long long k0 = k_begin; k0 < k_end; k0 += TK) {
    // ---- load tile
    for (int i = tid; i < C * TK; i += kSimtThreads) {
      const int c = i / TK, j = i - c * TK;
      const long long col = k0 + j;
      tile[c * (TK + 1) + j] = (col < k_end) ? to_f32(queue[(size_t)c * ldq + col]) : 0.f;
    }
    __syncthreads();
    // ---- S = q^ . tile   (log2 units)
    {
      float s[kRpt];
#pragma unroll
      for (int r = 0; r < kRpt; ++r) s[r] = 0.f;
      const float* qrow = qs + (size_t)(rg * kRpt) * C;
      int c = 0;
      const int c_vec = (C % 4 == 0) ? C : 0;  // float4 reads of q^ rows need 16B-aligned rows
      for (; c + 4 <= c_vec; c += 4) {
        const float t0 = tile[(c + 0) * (TK + 1) + j1], t1 = tile[(c + 1) * (TK + 1) + j1];
        const float t2 = tile[(c + 2) * (TK + 1) + j1], t3 = tile[(c + 3) * (TK + 1) + j1];
#pragma unroll
        for (int r = 0; r < kRpt; ++r) {
          const float4 qv = *reinterpret_cast<const float4*>(qrow + (size_t)r * C + c);
          s[r] = fmaf(qv.x, t0, s[r]);
          s[r] = fmaf(qv.y, t1, s[r]);
          s[r] = fmaf(qv.z, t2, s[r]);
          s[r] = fmaf(qv.w, t3, s[r]);
        }
      }
      for (; c < C; ++c) {
        const float t = tile[c * (TK + 1) + j1];
#pragma unroll
        for (int r = 0; r < kRpt; ++r) s[r] = fmaf(qrow[(size_t)r * C + c], t, s[r]);
      }
      const bool valid = (k0 + j1 < k_end);
      if (n2 != nullptr && valid) {   // |q^ - queue_j|^2 = |q^|^2 - 2 q^.queue_j + |queue_j|^2
        const float nj = n2[k0 + j1];
#pragma unroll
        for (int r = 0; r < kRpt; ++r)
          dacc[r] += sqrtf(fmaxf(fmaf(-2.f, s[r], s_qn2[rg * kRpt + r] + nj), 0.f));
      }
#pragma unroll
      for (int r = 0; r < kRpt; ++r) sp[(rg * kRpt + r) * TK + j1] = valid ? s[r] * scale2 : -INFINITY;
    }
    __syncthreads();
    // ---- row statistics (16 lanes per row; xor-shuffles 8..1 stay inside the 16-lane group)
    {
      float bv = -INFINITY;
      int bi = 0;
      for (int j = ssub; j < TK; j += 16) {
        const float v = sp[srow * TK + j];
        if (v > bv) { bv = v; bi = j; }
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
      }
      const float m_old = s_m[srow];
      const float m_new = fmaxf(m_old, bv);
      float lsum = 0.f;
      for (int j = ssub; j < TK; j += 16) {
        const float p = exp2f(sp[srow * TK + j] - m_new);  // masked columns: 2^(-inf) = 0
        sp[srow * TK + j] = p;
        lsum += p;
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
      __syncwarp();
      if (ssub == 0) {
        const float alpha = exp2f(m_old - m_new);  // first tile: 2^(-inf) = 0
        s_alpha[srow] = alpha;
        s_l[srow] = fmaf(s_l[srow], alpha, lsum);
        s_m[srow] = m_new;
        if (bv > s_av[srow]) {  // strict: the earliest tile keeps ties (first-occurrence argmax)
          s_av[srow] = bv;
          s_ai[srow] = (int)(k0 + bi);
        }
      }
    }
    __syncthreads();
    // ---- O = O*alpha + P . tile^T
#pragma unroll
    for (int i = 0; i < kSimtMaxCPerThread; ++i) {
      const int c = tid + kSimtThreads * i;
      if (want_o && c < C) {
#pragma unroll
        for (int r = 0; r < kSimtRows; ++r) acc[i][r] *= s_alpha[r];
        const float* trow = tile + (size_t)c * (TK + 1);
        for (int j = 0; j < TK; j += 4) {
          const float t0 = trow[j], t1 = trow[j + 1], t2 = trow[j + 2], t3 = trow[j + 3];
#pragma unroll
          for (int r = 0; r < kSimtRows; ++r) {
            const float4 p = *reinterpret_cast<const float4*>(sp + r * TK + j);
            acc[i][r] = fmaf(p.x, t0, acc[i][r]);
            acc[i][r] = fmaf(p.y, t1, acc[i][r]);
            acc[i][r] = fmaf(p.z, t2, acc[i][r]);
            acc[i][r] = fmaf(p.w, t3, acc[i][r]);
          }
        }
      }
    }
    __syncthreads();
  }

  // ---- emit partials
  if (n2 != nullptr) {
    // a row's columns live in TK consecutive threads = TK/32 warps: at most two shared-memory adds per
    // row, and a two-term floating-point sum does not depend on the order => deterministic
#pragma unroll
    for (int r = 0; r < kRpt; ++r) {
      const float v = warp_sum(dacc[r]);
      if ((tid & 31) == 0) atomicAdd(&s_dist[rg * kRpt + r], v);
    }
    __syncthreads();
    if (tid < kSimtRows && row0 + tid < B) pdist[(size_t)(row0 + tid) * gridDim.x + split] = s_dist[tid];
  }
  if (tid < kSimtRows && row0 + tid < B) {
    const size_t o = (size_t)(row0 + tid) * gridDim.x + split;
    pm[o] = s_m[tid];
    pl[o] = s_l[tid];
    pav[o] = s_av[tid];
    pai[o] = s_ai[tid];
  }
#pragma unroll
  for (int i = 0; i < kSimtMaxCPerThread; ++i) {
    const int c = tid + kSimtThreads * i;
    if (want_o && c < C) {
#pragma unroll
      for (int r = 0; r < kSimtRows; ++r)
        if (row0 + r < B) po[((size_t)split * B + row0 + r) * C + c] = acc[i][r];
    }
  }
}

template <typename TQ, int TK>
static int launch_simt(const float* q_hat, const void* queue, int B, int C, long long K, long long ldq, float scale2,
                       bool bf16_mode, const InfoNcePlan& p, InfoNcePartials out, bool want_o, cudaStream_t s) {
  const size_t smem =
      ((size_t)kSimtRows * C + (((size_t)C * (TK + 1) + 3) & ~(size_t)3) + (size_t)kSimtRows * TK + 7 * kSimtRows) * 4;
  auto kern = infonce_simt_kernel<TQ, TK>;
  RMCL_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(p.splits, p.row_blocks);
  kern<<<grid, kSimtThreads, smem, s>>>(q_hat, (const TQ*)queue, B, C, K, ldq, scale2, bf16_mode, p.cols_per_split,
                                        out.m, out.l, out.av, out.ai, out.o, out.n2, out.qn2, out.dist, want_o);
  RMCL_LAUNCH_OK("infonce_simt_kernel");
  return RMCL_OK;
}

int infonce_simt_launch(const float* q_hat, const void* queue, int queue_dtype, int B, int C, long long K,
                        long long ldq, float scale2, const InfoNcePlan& p, InfoNcePartials out, bool want_o, cudaStream_t s) {
  const bool bf = (queue_dtype == RMCL_BF16);
  if (p.row_blocks > 65535) {
    set_error("InfoNCE: too many rows (%d)", B);
    return RMCL_E_UNSUPPORTED_DIM;
  }
  if (p.tile_cols == 64) {
    return bf ? launch_simt<__nv_bfloat16, 64>(q_hat, queue, B, C, K, ldq, scale2, true, p, out, want_o, s)
              : launch_simt<float, 64>(q_hat, queue, B, C, K, ldq, scale2, false, p, out, want_o, s);
  }
  return bf ? launch_simt<__nv_bfloat16, 32>(q_hat, queue, B, C, K, ldq, scale2, true, p, out, want_o, s)
            : launch_simt<float, 32>(q_hat, queue, B, C, K, ldq, scale2, false, p, out, want_o, s);
}

}  // namespace rmcl
