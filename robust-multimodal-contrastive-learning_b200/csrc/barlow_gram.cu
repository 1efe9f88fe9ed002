// Barlow-Twins cross-correlation loss through Gram matrices: any (gathered) batch size, ~D/(1.5 Bg) times fewer
// flops than forming c.  Same reference lines as barlow.cu (vilt/modules/objectives.py:480-486, 506-512,
// 533-539; attack/pgd_attack_vilt.py:219-224).
//
// With c = q^T k / bs  (D x D, D = 8192) and the Gram matrices Gq = q q^T, Gk = k k^T  (Bg x Bg):
//     sum_ij c_ij^2            = <Gq, Gk>_F / bs^2
//     c_ii                     = sum_b q[b,i] k[b,i] / bs                       (column sums, elementwise)
//     on_diag                  = sum_i (c_ii - 1)^2
//     off_diag                 = <Gq, Gk>_F / bs^2 - sum_i c_ii^2
//     d(sum_ij c_ij^2)/dq[b,i] = 2/bs^2 (Gk q)[b,i],     d(c_ii)/dq[b,i] = k[b,i] / bs
// i.e. three GEMMs with the LONG dimension D as contraction (Gq, Gk: 2 Bg^2 D flop each) or as the free
// dimension (Gk q: 2 Bg^2 D) instead of two D x D x Bg contractions (4 Bg D^2 flop): 43x fewer flops at Bg = 128,
// 5x at Bg = 1024, and nothing of size D x D anywhere.  The whole loss is then bound by reading q and k once.
//
//   gram_prep    q, k (fp32/bf16) -> bf16 [BGp, D_ld] zero padded (BGp = Bg rounded up to 128); c_ii column sums
//   gram_kernel  split-K tcgen05 GEMMs Gq, Gk: CTA = one 128 x 128 tile of both x a range of D, operands TMA-staged
//                K-major from the same matrices (diagonal tiles load one operand and use it twice), fp32 partials
//   gram_finish  one CTA per Gram row: sums the split partials in fixed order, emits Gk as a bf16 hi/lo pair (16
//                mantissa bits for the gradient GEMM) and the row's share of <Gq, Gk>; the last CTA reduces the
//                loss terms deterministically
//   gram_grad    dq tile [128 batch rows x 128 features] = (Gk_hi + Gk_lo) q  on the tensor cores (A = Gk K-major,
//                B = q rows, MN-major), epilogue adds the elementwise diagonal terms and writes dq in place
// chained by programmatic dependent launch.
#include "common.cuh"
#include "tc_ptx.cuh"

namespace rmcl {

namespace {

using namespace tcx;

constexpr int kGThreads = 192;   // warps 0-3 epilogue (one TMEM lane quadrant each), warp 4 TMA + TMEM allocator, warp 5 MMA issuer
constexpr int kBox = 128 * 64 * 2;        // 16 KB: 128 rows x 64 bf16 columns, SWIZZLE_128B
constexpr int kGramStages = 3;            // x 64 KB
constexpr int kGradStages = 4;            // x 48 KB
constexpr int kGradTN = 128;

struct GShared {
  uint64_t full[4];
  uint64_t empty[4];
  uint64_t done;
  uint32_t tmem_base;
};

struct GramPlan {
  int BGp, tiles_n, splits;
  long long cols_per_split, d_ld;
  size_t off_qb, off_kb, off_part, off_ghi, off_glo, off_cdiag, off_blk, off_counter, total;
};

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

int barlow_gram_plan(int Bg, int D, GramPlan* p) {
  const int sms = sm_count();
  if (sms <= 0) return RMCL_E_CUDA;
  if (Bg > 4096) {
    set_error("rmcl_barlow_fwd_bwd: gathered batch %d > 4096 is not supported", Bg);
    return RMCL_E_UNSUPPORTED_DIM;
  }
  p->BGp = (Bg + 127) / 128 * 128;
  p->tiles_n = p->BGp / 128;
  p->d_ld = ((long long)D + 7) / 8 * 8;
  const long long tiles = (long long)p->tiles_n * p->tiles_n;
  const long long chunks = ((long long)D + 63) / 64;
  long long splits = sms / tiles;
  if (splits < 1) splits = 1;
  if (splits > 32) splits = 32;            // partials are summed in fixed order by gram_finish: keep them few
  if (splits > chunks) splits = chunks;
  const long long steps = (chunks + splits - 1) / splits;
  p->cols_per_split = steps * 64;
  p->splits = (int)(((long long)D + p->cols_per_split - 1) / p->cols_per_split);
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
  const size_t G = (size_t)p->BGp * p->BGp;
  p->off_qb = take((size_t)p->BGp * p->d_ld * 2);
  p->off_kb = take((size_t)p->BGp * p->d_ld * 2);
  p->off_part = take((size_t)p->splits * 2 * G * 4);
  p->off_ghi = take(G * 2);
  p->off_glo = take(G * 2);
  p->off_cdiag = take((size_t)D * 4);
  p->off_blk = take((size_t)p->BGp * 8);   // double: see gram_finish_kernel
  p->off_counter = take(256);
  p->total = off;
  return RMCL_OK;
}

// ------------------------------------------------------------------------------------------ prep
template <typename TQ, typename TK>
__global__ void __launch_bounds__(256) gram_prep_kernel(const TQ* __restrict__ q, const TK* __restrict__ k, int Bg, int D, int BGp,
                                                        long long d_ld, float inv_bs, __nv_bfloat16* __restrict__ qb,
                                                        __nv_bfloat16* __restrict__ kb, float* __restrict__ cdiag,
                                                        float* __restrict__ cdiag_out, unsigned int* __restrict__ counter) {
  __shared__ float red[8][33];
  if (threadIdx.x == 0) pdl_trigger();
  if (blockIdx.x == 0 && threadIdx.x == 0) *counter = 0u;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + tx;
  float acc = 0.f;
  for (int bb = 0; bb < BGp; bb += 64) {
    float qv[8], kv[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int b = bb + ty + 8 * u;
      const bool ok = (b < Bg) && (i < D);
      qv[u] = ok ? to_f32(q[(size_t)b * D + i]) : 0.f;
      kv[u] = ok ? to_f32(k[(size_t)b * D + i]) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int b = bb + ty + 8 * u;
      const __nv_bfloat16 q16 = __float2bfloat16_rn(qv[u]), k16 = __float2bfloat16_rn(kv[u]);
      if (b < BGp && i < d_ld) {
        qb[(size_t)b * d_ld + i] = q16;
        kb[(size_t)b * d_ld + i] = k16;
      }
      acc = fmaf(__bfloat162float(q16), __bfloat162float(k16), acc);   // the products the tensor cores see
    }
  }
  red[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && i < D) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][tx];
    t *= inv_bs;
    cdiag[i] = t;
    if (cdiag_out) cdiag_out[i] = t;
  }
}

// ---------------------------------------------------------------------------------- Gram matrices
__global__ void __launch_bounds__(kGThreads, 1)
    gram_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k, int D, int BGp, int tiles_n,
                long long cols_per_split, float* __restrict__ part) {
  constexpr int kStageBytes = 4 * kBox;
  constexpr uint32_t kIdesc = make_idesc(128, 128, 0);
  extern __shared__ uint8_t smem_raw[];
  __shared__ GShared sh;
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int split = blockIdx.x;
  const int tm = blockIdx.y / tiles_n, tn = blockIdx.y % tiles_n;
  const bool diag = (tm == tn);
  const long long k_begin = (long long)split * cols_per_split;
  const long long k_end = (k_begin + cols_per_split < D) ? k_begin + cols_per_split : D;
  const int n_steps = (int)((k_end - k_begin + 63) / 64);

  if (tid == 0) {
    for (int i = 0; i < kGramStages; ++i) {
      mbar_init(&sh.full[i], 1);
      mbar_init(&sh.empty[i], 1);
    }
    mbar_init(&sh.done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_q) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_k) : "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh.tmem_base)), "r"(256)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sh.tmem_base;
  pdl_wait();   // the bf16 operands behind the tensor maps are written by the prep kernel

  if (warp == 4) {
    for (int i = 0; i < n_steps; ++i) {
      const int st = i % kGramStages;
      mbar_wait(&sh.empty[st], ((i / kGramStages) & 1) ^ 1);
      if (elect_one()) {
        uint8_t* base = ring + (size_t)st * kStageBytes;
        const int x = (int)(k_begin + (long long)i * 64);
        mbar_expect_tx(&sh.full[st], diag ? 2 * kBox : 4 * kBox);
        tma_load_2d(base, &tmap_q, &sh.full[st], x, tm * 128);
        tma_load_2d(base + kBox, &tmap_k, &sh.full[st], x, tm * 128);
        if (!diag) {
          tma_load_2d(base + 2 * kBox, &tmap_q, &sh.full[st], x, tn * 128);
          tma_load_2d(base + 3 * kBox, &tmap_k, &sh.full[st], x, tn * 128);
        }
      }
      __syncwarp();
    }
  } else if (warp == 5) {
    for (int i = 0; i < n_steps; ++i) {
      const int st = i % kGramStages;
      mbar_wait(&sh.full[st], (i / kGramStages) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t aq = smem_u32(ring + (size_t)st * kStageBytes), ak = aq + kBox;
        const uint32_t bq = diag ? aq : aq + 2 * kBox, bk = diag ? ak : aq + 3 * kBox;
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          // both operands K-major: rows 128 B apart, 8-row groups 1024 B apart, 16 columns = 32 B inside the swizzled row
          tc_mma_ss(tmem, make_sw128_desc(aq + s * 32, 16, 1024), make_sw128_desc(bq + s * 32, 16, 1024), kIdesc,
                    (i > 0 || s > 0) ? 1u : 0u);
          tc_mma_ss(tmem + 128, make_sw128_desc(ak + s * 32, 16, 1024), make_sw128_desc(bk + s * 32, 16, 1024), kIdesc,
                    (i > 0 || s > 0) ? 1u : 0u);
        }
        tc_commit(&sh.empty[st]);
        if (i == n_steps - 1) tc_commit(&sh.done);
      }
      __syncwarp();
    }
  } else {
    // epilogue: partial tile rows -> global (every thread writes whole 128-byte lines of its own row)
    const int r = warp * 32 + lane;
    const uint32_t tlane = tmem + ((uint32_t)(warp * 32) << 16);
    mbar_wait(&sh.done, 0);
    tc_fence_after();
    if (tid == 0) pdl_trigger();
#pragma unroll 1
    for (int g = 0; g < 2; ++g) {
      float* dst = part + (((size_t)split * 2 + g) * BGp + (size_t)tm * 128 + r) * BGp + (size_t)tn * 128;
#pragma unroll 1
      for (int ch = 0; ch < 4; ++ch) {
        uint32_t o[32];
        tc_ld32(tlane + g * 128 + ch * 32, o);
        tc_wait_ld();
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<uint4*>(dst + ch * 32 + 4 * j) = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
  }
}

// ---------------------------------------------------------------------------------------- finish
__device__ __forceinline__ float block_sum_256(float v, float* red /*[8]*/) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) t += red[w];
  return t;
}

__device__ __forceinline__ double block_sum_256_d(double v, double* red /*[8]*/) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
#pragma unroll
  for (int w = 0; w < 8; ++w) t += red[w];
  return t;
}

// off_diag = <Gq, Gk>/bs^2 - sum_i c_ii^2 is a difference of two sums that are both ~D when c is close to the identity
// (a converged model on a gathered batch >= D): in fp32 one ulp of D = 128 is already 2e-3 of an off_diag of 3e-3
// (measured: tests/test_kernels_gpu.py::test_barlow_near_identity_correlation).  The Gram entries themselves are fp32
// tensor-core accumulations (relative error ~1e-7 each, averaged over Bg^2 terms); everything downstream of them — the
// inner product, the two diagonal sums and the subtraction — is therefore carried in double.  That is Bg^2 + 2 D double
// operations per call, nothing next to the GEMMs.
__global__ void __launch_bounds__(256) gram_finish_kernel(int BGp, int splits, int D, float inv_bs, float lambda, float loss_scale,
                                                          const float* __restrict__ part, const float* __restrict__ cdiag,
                                                          __nv_bfloat16* __restrict__ ghi, __nv_bfloat16* __restrict__ glo,
                                                          double* __restrict__ blk, unsigned int* __restrict__ counter,
                                                          float* __restrict__ on_out, float* __restrict__ off_out,
                                                          float* __restrict__ loss_out) {
  __shared__ double red[8];
  __shared__ bool s_last;
  const int a = blockIdx.x, tid = threadIdx.x;
  const size_t G = (size_t)BGp * BGp;
  pdl_wait();
  if (tid == 0) pdl_trigger();
  double acc = 0.0;
  for (int b = tid; b < BGp; b += 256) {
    float gq = 0.f, gk = 0.f;
    for (int s = 0; s < splits; ++s) {
      gq += __ldcs(part + ((size_t)s * 2) * G + (size_t)a * BGp + b);
      gk += __ldcs(part + ((size_t)s * 2 + 1) * G + (size_t)a * BGp + b);
    }
    const __nv_bfloat16 hi = __float2bfloat16_rn(gk);
    ghi[(size_t)a * BGp + b] = hi;
    glo[(size_t)a * BGp + b] = __float2bfloat16_rn(gk - __bfloat162float(hi));
    acc = fma((double)gq, (double)gk, acc);
  }
  acc = block_sum_256_d(acc, red);
  if (tid == 0) {
    blk[a] = acc;
    __threadfence();
    s_last = (atomicAdd(counter, 1u) == (unsigned)BGp - 1u);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  double tot = 0.0, cd2 = 0.0, on = 0.0;
  for (int j = tid; j < BGp; j += 256) tot += __ldcg(blk + j);
  for (int i = tid; i < D; i += 256) {
    const double c = (double)cdiag[i];
    cd2 = fma(c, c, cd2);
    on = fma(c - 1.0, c - 1.0, on);
  }
  tot = block_sum_256_d(tot, red);
  cd2 = block_sum_256_d(cd2, red);
  on = block_sum_256_d(on, red);
  if (tid == 0) {
    const double off = tot * (double)inv_bs * (double)inv_bs - cd2;
    if (on_out) *on_out = (float)on;
    if (off_out) *off_out = (float)off;
    if (loss_out) *loss_out = (float)((double)loss_scale * (on + (double)lambda * off));
  }
}

// ------------------------------------------------------------------------------------------ grad
__global__ void __launch_bounds__(kGThreads, 1)
    gram_grad_kernel(const __grid_constant__ CUtensorMap tmap_ghi, const __grid_constant__ CUtensorMap tmap_glo,
                     const __grid_constant__ CUtensorMap tmap_qrows, int D, int BGp, long long d_ld, int tm0, int b0, int Bl,
                     float g1, float g2, float g3, const __nv_bfloat16* __restrict__ kb, const float* __restrict__ cdiag,
                     float* __restrict__ dq) {
  constexpr int kBBytes = 2 * 64 * 64 * 2;             // q rows: two boxes of 64 batch rows x 64 features
  constexpr int kStageBytes = 2 * kBox + kBBytes;      // Gk_hi tile + Gk_lo tile + q rows = 48 KB
  constexpr uint32_t kIdesc = make_idesc(128, kGradTN, 1);
  constexpr int kRowBytes = kGradTN * 4 + 16;          // epilogue staging row (fp32), padded against bank conflicts
  static_assert(128 * kRowBytes <= kGradStages * kStageBytes, "epilogue staging must fit the ring");
  extern __shared__ uint8_t smem_raw[];
  __shared__ GShared sh;
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int col0 = blockIdx.x * kGradTN;
  const int tm = tm0 + blockIdx.y;
  const int n_steps = BGp / 64;

  if (tid == 0) {
    for (int i = 0; i < kGradStages; ++i) {
      mbar_init(&sh.full[i], 1);
      mbar_init(&sh.empty[i], 1);
    }
    mbar_init(&sh.done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_ghi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_glo) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_qrows) : "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh.tmem_base)), "r"(128)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sh.tmem_base;
  pdl_wait();

  if (warp == 4) {
    for (int i = 0; i < n_steps; ++i) {
      const int st = i % kGradStages;
      mbar_wait(&sh.empty[st], ((i / kGradStages) & 1) ^ 1);
      if (elect_one()) {
        uint8_t* base = ring + (size_t)st * kStageBytes;
        mbar_expect_tx(&sh.full[st], kStageBytes);
        tma_load_2d(base, &tmap_ghi, &sh.full[st], i * 64, tm * 128);
        tma_load_2d(base + kBox, &tmap_glo, &sh.full[st], i * 64, tm * 128);
        tma_load_2d(base + 2 * kBox, &tmap_qrows, &sh.full[st], col0, i * 64);
        tma_load_2d(base + 2 * kBox + 8192, &tmap_qrows, &sh.full[st], col0 + 64, i * 64);
      }
      __syncwarp();
    }
  } else if (warp == 5) {
    for (int i = 0; i < n_steps; ++i) {
      const int st = i % kGradStages;
      mbar_wait(&sh.full[st], (i / kGradStages) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t ahi = smem_u32(ring + (size_t)st * kStageBytes), alo = ahi + kBox, bq = ahi + 2 * kBox;
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          // A = Gk tile [M=128 batch rows][K=16 batch columns], K-major;  B = q rows [K=16 batch rows][N=128 features],
          // MN-major: 8-row groups 1024 B apart, the two 64-feature boxes 8192 B apart
          const uint64_t bd = make_sw128_desc(bq + s * 2048, 8192, 1024);
          tc_mma_ss(tmem, make_sw128_desc(ahi + s * 32, 16, 1024), bd, kIdesc, (i > 0 || s > 0) ? 1u : 0u);
          tc_mma_ss(tmem, make_sw128_desc(alo + s * 32, 16, 1024), bd, kIdesc, 1u);
        }
        tc_commit(&sh.empty[st]);
        if (i == n_steps - 1) tc_commit(&sh.done);
      }
      __syncwarp();
    }
  } else {
    // epilogue: dq[b, i] = g1 (Gk q)[b, i] + (g2 c_ii + g3) k[b, i]
    const int r = warp * 32 + lane;
    const int b = tm * 128 + r, lb = b - b0;
    const bool row_ok = (lb >= 0) && (lb < Bl);
    const uint32_t tlane = tmem + ((uint32_t)(warp * 32) << 16);
    int nvalid = D - col0;
    if (nvalid > kGradTN) nvalid = kGradTN;
    mbar_wait(&sh.done, 0);
    tc_fence_after();
    uint8_t* stage = ring + (size_t)r * kRowBytes;      // every TMA write has been consumed
    const __nv_bfloat16* krow = kb + (size_t)b * d_ld + col0;   // d_ld % 8 == 0, col0 % 128 == 0: 16-byte aligned
#pragma unroll 1
    for (int ch = 0; ch < kGradTN / 32; ++ch) {
      uint32_t o[32];
      tc_ld32(tlane + ch * 32, o);
      tc_wait_ld();
      if (row_ok) {
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const int c = ch * 32 + v * 8;
          uint4 ku = make_uint4(0u, 0u, 0u, 0u);
          if (col0 + c < d_ld) ku = __ldg(reinterpret_cast<const uint4*>(krow + c));
          const __nv_bfloat162* kh = reinterpret_cast<const __nv_bfloat162*>(&ku);
          float out[8];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 kf = __bfloat1622float2(kh[j]);
            const int i0 = col0 + c + 2 * j;
            const float c0 = (i0 < D) ? __ldg(cdiag + i0) : 0.f, c1 = (i0 + 1 < D) ? __ldg(cdiag + i0 + 1) : 0.f;
            out[2 * j] = fmaf(g1, __uint_as_float(o[v * 8 + 2 * j]), fmaf(g2, c0, g3) * kf.x);
            out[2 * j + 1] = fmaf(g1, __uint_as_float(o[v * 8 + 2 * j + 1]), fmaf(g2, c1, g3) * kf.y);
          }
          *reinterpret_cast<float4*>(stage + (size_t)c * 4) = make_float4(out[0], out[1], out[2], out[3]);
          *reinterpret_cast<float4*>(stage + (size_t)c * 4 + 16) = make_float4(out[4], out[5], out[6], out[7]);
        }
      }
    }
    if (row_ok && nvalid > 0) {
      float* dst = dq + (size_t)lb * D + col0;
      if ((D & 3) == 0) {   // 16-byte aligned rows: one bulk async copy per row
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        bulk_store_row(dst, stage, (uint32_t)nvalid * 4u);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      } else {
        const float* src = reinterpret_cast<const float*>(stage);
        for (int j = 0; j < nvalid; ++j) dst[j] = src[j];
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128) : "memory");
  }
}

template <typename TQ, typename TK>
int launch_gram_prep(const void* q, const void* k, int Bg, int D, const GramPlan& p, char* ws, float inv_bs, float* cdiag_out,
                     cudaStream_t s) {
  const int blocks = (int)((p.d_ld + 31) / 32);
  gram_prep_kernel<TQ, TK><<<blocks, 256, 0, s>>>((const TQ*)q, (const TK*)k, Bg, D, p.BGp, p.d_ld, inv_bs,
                                                  (__nv_bfloat16*)(ws + p.off_qb), (__nv_bfloat16*)(ws + p.off_kb),
                                                  (float*)(ws + p.off_cdiag), cdiag_out, (unsigned int*)(ws + p.off_counter));
  RMCL_LAUNCH_OK("gram_prep_kernel");
  return RMCL_OK;
}

}  // namespace

size_t barlow_gram_workspace_bytes(int Bg, int D) {
  GramPlan p;
  if (barlow_gram_plan(Bg, D, &p) != RMCL_OK) return 0;
  return p.total;
}

int barlow_gram_run(const void* q, int q_dtype, const void* k, int k_dtype, int Bg, int D, int b0, int Bl, float inv_bs,
                    float lambda, float w_on, float w_off, float loss_scale, float* on_diag, float* off_diag, float* loss,
                    float* dq, float* cdiag_out, void* workspace, size_t workspace_bytes, cudaStream_t s) {
  GramPlan p;
  int rc = barlow_gram_plan(Bg, D, &p);
  if (rc != RMCL_OK) return rc;
  if (workspace_bytes < p.total) {
    set_error("rmcl_barlow_fwd_bwd: workspace %zu < required %zu", workspace_bytes, p.total);
    return RMCL_E_WORKSPACE;
  }
  char* ws = (char*)workspace;
  using bf16 = __nv_bfloat16;
  if (q_dtype == RMCL_F32 && k_dtype == RMCL_F32) rc = launch_gram_prep<float, float>(q, k, Bg, D, p, ws, inv_bs, cdiag_out, s);
  else if (q_dtype == RMCL_F32) rc = launch_gram_prep<float, bf16>(q, k, Bg, D, p, ws, inv_bs, cdiag_out, s);
  else if (k_dtype == RMCL_F32) rc = launch_gram_prep<bf16, float>(q, k, Bg, D, p, ws, inv_bs, cdiag_out, s);
  else rc = launch_gram_prep<bf16, bf16>(q, k, Bg, D, p, ws, inv_bs, cdiag_out, s);
  if (rc != RMCL_OK) return rc;

  alignas(64) CUtensorMap tq, tk, thi, tlo, tqr;
  if ((rc = make_tmap_bf16(&tq, ws + p.off_qb, (uint64_t)p.BGp, (uint64_t)D, (uint64_t)p.d_ld, 128)) != RMCL_OK) return rc;
  if ((rc = make_tmap_bf16(&tk, ws + p.off_kb, (uint64_t)p.BGp, (uint64_t)D, (uint64_t)p.d_ld, 128)) != RMCL_OK) return rc;
  const size_t gram_smem = (size_t)kGramStages * 4 * kBox + 1024;
  RMCL_CUDA_OK(cudaFuncSetAttribute(gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gram_smem));
  RMCL_CUDA_OK(launch_pdl(gram_kernel, dim3(p.splits, p.tiles_n * p.tiles_n), dim3(kGThreads), gram_smem, s, tq, tk, D, p.BGp,
                          p.tiles_n, p.cols_per_split, (float*)(ws + p.off_part)));

  RMCL_CUDA_OK(launch_pdl(gram_finish_kernel, dim3(p.BGp), dim3(256), (size_t)0, s, p.BGp, p.splits, D, inv_bs, lambda, loss_scale,
                          (const float*)(ws + p.off_part), (const float*)(ws + p.off_cdiag), (bf16*)(ws + p.off_ghi),
                          (bf16*)(ws + p.off_glo), (double*)(ws + p.off_blk), (unsigned int*)(ws + p.off_counter), on_diag,
                          off_diag, loss));
  if (dq == nullptr) return RMCL_OK;

  if ((rc = make_tmap_bf16(&thi, ws + p.off_ghi, (uint64_t)p.BGp, (uint64_t)p.BGp, (uint64_t)p.BGp, 128)) != RMCL_OK) return rc;
  if ((rc = make_tmap_bf16(&tlo, ws + p.off_glo, (uint64_t)p.BGp, (uint64_t)p.BGp, (uint64_t)p.BGp, 128)) != RMCL_OK) return rc;
  if ((rc = make_tmap_bf16(&tqr, ws + p.off_qb, (uint64_t)p.BGp, (uint64_t)D, (uint64_t)p.d_ld, 64)) != RMCL_OK) return rc;
  const int tm0 = b0 / 128, tm1 = (b0 + Bl - 1) / 128;
  const size_t grad_smem = (size_t)kGradStages * (2 * kBox + 2 * 64 * 64 * 2) + 1024;
  RMCL_CUDA_OK(cudaFuncSetAttribute(gram_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)grad_smem));
  const float g1 = 2.f * w_off * inv_bs * inv_bs * loss_scale;
  const float g2 = 2.f * inv_bs * (w_on - w_off) * loss_scale;
  const float g3 = -2.f * inv_bs * w_on * loss_scale;
  RMCL_CUDA_OK(launch_pdl(gram_grad_kernel, dim3((D + kGradTN - 1) / kGradTN, tm1 - tm0 + 1), dim3(kGThreads), grad_smem, s, thi,
                          tlo, tqr, D, p.BGp, p.d_ld, tm0, b0, Bl, g1, g2, g3, (const bf16*)(ws + p.off_kb),
                          (const float*)(ws + p.off_cdiag), dq));
  return RMCL_OK;
}

}  // namespace rmcl
