// K1 partial stage for wide projections (C = 512 or 768: BASELINE cfg5, dim 768 / queue 262144), tcgen05.
//
// The fused single-pass kernel (infonce_tc.cu) keeps Q^ (C/2 columns) and the fp32 accumulator O (C columns)
// of its 128 rows in tensor memory; at C = 768 that is 384 + 768 of the 512 columns an SM has, and no
// split over a CTA pair removes the need for the full-C contraction in S.  Here the two GEMMs run as two
// tensor-core kernels with the (unnormalised, bf16) probabilities as the only intermediate:
//
//   S pass   infonce_s_kernel<C>:  CTA = 128 rows x one split of the queue, tiles of 64 columns.
//            S = Q^ . tile over C in chunks of 256 rows (TMA ring of 32 KB chunk stages, Q^ in TMEM);
//            P~ = 2^(S*log2e/tau - m_ref) with ONE reference per (row, split), fixed after the first tile
//            (its row maximum + 24): no rescaling can be needed downstream.  Emits the split statistics
//            (m_ref, l, argmax) of infonce.cuh and P~ [B_pad, K_pad] in bf16.
//   PV pass  infonce_pv_kernel:    CTA = 128 rows x the same split x one 256-wide slice of C.
//            O[:, slice] = sum_tiles P~_tile . queue[slice, tile]^T — a plain split-K tcgen05 GEMM with both
//            operands staged by TMA (P~ K-major, the queue slice K-major), O (256 columns) in TMEM,
//            bf16 partials written exactly like the fused kernel's.
//
// The logits are still never materialised in fp32 and never leave the chip in a form the reference has
// (B x K bf16 of P~ instead of 3-4 fp32 copies of the logits, objectives.py:272-274,333).  A fixed reference
// cannot follow a row maximum that grows by more than 2^100 inside one split; that cannot happen for
// normalised keys (|logit| <= 1/tau) and is caught, not ignored: the S pass raises a flag that makes the
// finalize kernel return NaN.
//
// SPLIT instantiations — the fp32-accurate InfoNCE on the bf16 tensor cores (queue layout RMCL_BF16_HILO, C in {64,128,256}).
// The reference evaluates the PGD inner loss in fp32 (attack/pgd_attack_vilt.py:141,152-158) against its fp32 queue buffer
// (vilt_module.py:92); every fp32 operand x is carried as a bf16 pair (hi = bf16(x), lo = bf16(x - hi): 16 mantissa bits) and
// every product as three tensor-core products (the lo.lo term, 2^-18 relative, is dropped):
//   S pass   S  = q_hi.Q_hi + q_hi.Q_lo + q_lo.Q_hi     three accumulating MMA chains per tile, Q^ = [q_hi | q_lo] in TMEM,
//                                                       ring stages = the hi plane, the lo plane, the hi plane again (L2 hit)
//            P~ written as a hi/lo pair as well (two planes of [B_pad, K_pad])
//   PV pass  O  = P_hi.Q_hi^T + P_hi.Q_lo^T + P_lo.Q_hi^T   three MMAs per K=16 step into one accumulator; fp32 partials
// i.e. 3x the bf16 flops for ~2^-17 relative error per product — against the fp32 CUDA-core kernel this replaces
// (infonce_simt.cu: 878 us at B256 C256 K65536, the same as eager cuBLAS SGEMM + ATen).
#include "infonce.cuh"
#include "tc_ptx.cuh"

namespace rmcl {

namespace {

using namespace tcx;

constexpr int kThreads = 320;          // warps 0-7 softmax / epilogue, warp 8 TMA producer + TMEM allocator, warp 9 MMA issuer
constexpr int kRows = 128;
constexpr int kTN = 64;                // queue columns per tile
constexpr int kChunkRows = 256;        // rows of C per TMA chunk stage / per PV slice
constexpr int kChunkBytes = kChunkRows * kTN * 2;   // 32 KB
constexpr float kMargin = 24.f;        // initial reference = first tile's row maximum + 2^24 (see infonce_tc.cu)
constexpr float kOverflow = 100.f;     // P~ would exceed 2^100 (bf16 tops out at 2^127): flag it

#ifndef RMCL_TC2_Q_FIRST
#define RMCL_TC2_Q_FIRST 1
#endif

struct SShared {
  uint64_t k_full[8];
  uint64_t k_empty[8];
  uint64_t s_full[2];
  uint64_t s_free[2];
  uint64_t q_full;
  uint32_t tmem_base;
  float m_ref[kRows];
  float xl[kRows];
  float xav[kRows];
  int xai[kRows];
  float xd[kRows];     // diagnostics: the odd-tile warp's distance sum
};

// =========================================================================================== S pass
template <int C, bool DIAG, bool SPLIT>
__global__ void __launch_bounds__(kThreads, 1)
    infonce_s_kernel(const __grid_constant__ CUtensorMap tmap_queue, const __nv_bfloat16* __restrict__ q_hat, int B,
                     long long K, long long k_pad, float scale2, long long cols_per_split, int want_argmax,
                     float* __restrict__ pm, float* __restrict__ pl, float* __restrict__ pav, int* __restrict__ pai,
                     __nv_bfloat16* __restrict__ ptilde, unsigned int* __restrict__ overflow_flag,
                     const float* __restrict__ n2, const float* __restrict__ qn2, float* __restrict__ pdist) {
  // SPLIT: C is the logical width; one chunk = one whole plane of C rows, three chunks per tile (hi, lo, hi)
  constexpr int kChunkRows = SPLIT ? C : rmcl::kChunkRows;
  constexpr int kChunkBytes = kChunkRows * kTN * 2;
  constexpr int kChunks = SPLIT ? 3 : C / kChunkRows;  // ring stages consumed per tile
  constexpr int kQCols = SPLIT ? 2 * C : C;            // width of a Q^ operand row: [q_hi | q_lo] or q
  constexpr int kStages = 5;
  constexpr uint32_t kTmQ = 0, kTmS = kQCols / 2;
  static_assert((SPLIT || C % kChunkRows == 0) && kQCols / 2 + 2 * kTN <= 512, "tensor memory budget");
  constexpr uint32_t kIdescS = make_idesc(128, kTN, 1);

  extern __shared__ uint8_t smem_raw[];
  __shared__ SShared sh;
  uint8_t* stage_buf = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* ring = stage_buf + 8 * 4096;                // 8 warps x 4 KB of P~ / Q^ transposing scratch first

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int split = blockIdx.x;
  const int row0 = blockIdx.y * kRows;
  const long long k_begin = (long long)split * cols_per_split;
  const long long k_end = (k_begin + cols_per_split < K) ? k_begin + cols_per_split : K;
  const int n_tiles = (int)((k_end - k_begin + kTN - 1) / kTN);

  if (tid == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&sh.k_full[i], 1);
      mbar_init(&sh.k_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sh.s_full[i], 1);
      mbar_init(&sh.s_free[i], kRows);
    }
    mbar_init(&sh.q_full, 8 * 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_queue) : "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh.tmem_base)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sh.tmem_base;

  if (warp < 8) {
    // ===================================================================== softmax warps
    const int quad = warp & 3, par = warp >> 2;           // TMEM lane quadrant; even / odd tiles
    const int r = quad * 32 + lane;
    const uint32_t tlane = tmem + ((uint32_t)(quad * 32) << 16);
    uint8_t* scratch = stage_buf + warp * 4096;           // this warp's 32 rows x 128 B

    // ---- Q^ rows -> tensor memory, 64-column chunks transposed through the warp's scratch (see infonce_tc.cu)
    pdl_wait();
    {
      constexpr int kQChunks = kQCols / 64;
      const __nv_bfloat16* qw = q_hat + (size_t)(split % kQhatReplicas) * ((size_t)gridDim.y * kRows * kQCols) +
                                (size_t)(row0 + quad * 32) * kQCols;
#pragma unroll 1
      for (int ch = par; ch < kQChunks; ch += 2) {
        uint4 v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)
          v[j] = __ldg(reinterpret_cast<const uint4*>(qw + (size_t)(4 * j + (lane >> 3)) * kQCols + ch * 64) + (lane & 7));
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int rr = 4 * j + (lane >> 3), pc = lane & 7;
          *reinterpret_cast<uint4*>(scratch + rr * 128 + ((pc ^ (rr & 7)) << 4)) = v[j];
        }
        __syncwarp();
        uint32_t w[32];
#pragma unroll
        for (int pc = 0; pc < 8; ++pc) {
          const uint4 u = *reinterpret_cast<const uint4*>(scratch + lane * 128 + ((pc ^ (lane & 7)) << 4));
          w[4 * pc + 0] = u.x; w[4 * pc + 1] = u.y; w[4 * pc + 2] = u.z; w[4 * pc + 3] = u.w;
        }
        __syncwarp();
        tc_st32(tlane + kTmQ + ch * 32, w);
      }
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(&sh.q_full);
    }

    float m_ref = 0.f, l_run = 0.f, av_raw = -INFINITY;
    int ai = 0;
    bool overflow = false;
    float dsum = 0.f;                                     // diagnostics: sum_j |q^_r - queue_j| (infonce_tc.cu)
    const float qn2_r = (DIAG && row0 + r < B) ? qn2[row0 + r] : 0.f;
    for (int i = par; i < n_tiles; i += 2) {
      const int b = par;
      const uint32_t ts = tlane + kTmS + b * kTN;
      mbar_wait(&sh.s_full[b], (i >> 1) & 1);
      tc_fence_after();
      uint32_t sv[kTN];
      tc_ld32(ts, sv);
      tc_ld32(ts + 32, sv + 32);
      tc_wait_ld();
      tc_fence_before();
      mbar_arrive(&sh.s_free[b]);
      const long long col0 = k_begin + (long long)i * kTN;
      if (DIAG) {
        // mean L2 distance to the negatives (objectives.py:343): |q^ - queue_j|^2 = |q^|^2 - 2 S_j + |queue_j|^2
        const float4* nv = reinterpret_cast<const float4*>(n2 + col0);
#pragma unroll
        for (int c4 = 0; c4 < kTN / 4; ++c4) {
          if (col0 + 4 * c4 < k_end) {                    // k_end is a multiple of 8 on this path
            const float4 nn = __ldg(nv + c4);
            const float d0 = fmaf(-2.f, __uint_as_float(sv[4 * c4 + 0]), qn2_r + nn.x);
            const float d1 = fmaf(-2.f, __uint_as_float(sv[4 * c4 + 1]), qn2_r + nn.y);
            const float d2 = fmaf(-2.f, __uint_as_float(sv[4 * c4 + 2]), qn2_r + nn.z);
            const float d3 = fmaf(-2.f, __uint_as_float(sv[4 * c4 + 3]), qn2_r + nn.w);
            dsum += (sqrtf(fmaxf(d0, 0.f)) + sqrtf(fmaxf(d1, 0.f))) + (sqrtf(fmaxf(d2, 0.f)) + sqrtf(fmaxf(d3, 0.f)));
          }
        }
      }
      if (col0 + kTN > k_end) {
        const int valid = (int)(k_end - col0);
#pragma unroll
        for (int j = 0; j < kTN; ++j)
          if (j >= valid) sv[j] = 0xff800000u;
      }
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < kTN; ++j) mx = fmaxf(mx, __uint_as_float(sv[j]));
      if (want_argmax && mx > av_raw) {
        av_raw = mx;
        int idx = 0;
#pragma unroll
        for (int j = kTN - 1; j >= 0; --j)
          if (__uint_as_float(sv[j]) == mx) idx = j;
        ai = (int)col0 + idx;
      }
      // one reference per (row, split): fixed by tile 0 (even-tile warp), read once by the odd-tile warp
      if (i == 0) {
        m_ref = mx * scale2 + kMargin;
        sh.m_ref[r] = m_ref;
        __threadfence_block();
        if (n_tiles > 1) named_bar_arrive(1 + quad, 64);
      } else if (i == 1) {
        named_bar_sync(1 + quad, 64);
        m_ref = sh.m_ref[r];
      }
      overflow |= (mx * scale2 - m_ref > kOverflow);

      const float neg_m = -m_ref;
      float ls[4] = {0.f, 0.f, 0.f, 0.f};
      uint32_t pw[kTN / 2];
      uint32_t pw_lo[SPLIT ? kTN / 2 : 1];
#pragma unroll
      for (int j = 0; j < kTN / 2; ++j) {
        const float p0 = ex2_ftz(fmaf(__uint_as_float(sv[2 * j]), scale2, neg_m));
        const float p1 = ex2_ftz(fmaf(__uint_as_float(sv[2 * j + 1]), scale2, neg_m));
        ls[j & 3] += p0 + p1;
        const __nv_bfloat162 pk = __floats2bfloat162_rn(p0, p1);
        pw[j] = *reinterpret_cast<const uint32_t*>(&pk);
        if (SPLIT) {   // the low halves: what bf16 rounding took away from p
          const float2 hf = __bfloat1622float2(pk);
          const __nv_bfloat162 lk = __floats2bfloat162_rn(p0 - hf.x, p1 - hf.y);
          pw_lo[j] = *reinterpret_cast<const uint32_t*>(&lk);
        }
      }
      l_run += (ls[0] + ls[1]) + (ls[2] + ls[3]);
      // P~ row r -> scratch (16-byte chunks XOR-swizzled), then the warp stores its 32 rows coalesced:
      // 8 lanes cover the 128-byte segment of one row, 4 rows per instruction
      if (ptilde == nullptr) continue;   // statistics-only call (no gradient requested): no P~, no PV pass
#pragma unroll
      for (int plane = 0; plane < (SPLIT ? 2 : 1); ++plane) {
        const uint32_t* src = (plane == 0) ? pw : pw_lo;
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 8; ++c)
          *reinterpret_cast<uint4*>(scratch + lane * 128 + ((c ^ (lane & 7)) << 4)) =
              make_uint4(src[4 * c], src[4 * c + 1], src[4 * c + 2], src[4 * c + 3]);
        __syncwarp();
        // plane 1 (lo) follows the B_pad rows of plane 0
        __nv_bfloat16* prow = ptilde + ((size_t)plane * gridDim.y * kRows + row0 + quad * 32) * k_pad + col0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int rr = 4 * j + (lane >> 3), pc = lane & 7;
          const uint4 u = *reinterpret_cast<const uint4*>(scratch + rr * 128 + ((pc ^ (rr & 7)) << 4));
          *reinterpret_cast<uint4*>(prow + (size_t)rr * k_pad + pc * 8) = u;
        }
      }
    }

    // ---- split statistics: both warps used the same reference, so the sums just add
    if (__any_sync(0xffffffffu, overflow) && lane == 0) atomicExch(overflow_flag, 1u);
    if (par == 1) {
      sh.xl[r] = l_run;
      sh.xav[r] = av_raw;
      sh.xai[r] = ai;
      if (DIAG) sh.xd[r] = dsum;
    }
    named_bar_sync(9 + quad, 64);
    if (par == 0 && row0 + r < B) {
      const float l_tot = l_run + sh.xl[r];
      const float av1 = sh.xav[r];
      const int ai1 = sh.xai[r];
      if (av1 > av_raw || (av1 == av_raw && ai1 < ai)) { av_raw = av1; ai = ai1; }
      const size_t o = (size_t)(row0 + r) * gridDim.x + split;
      pm[o] = m_ref;
      pl[o] = l_tot;
      pav[o] = av_raw * scale2;
      pai[o] = ai;
      if (DIAG) pdist[o] = dsum + sh.xd[r];
    }
    tc_fence_before();
  } else if (warp == 8) {
    // ===================================================================== TMA producer
    for (int i = 0; i < n_tiles; ++i) {
      const long long col0 = k_begin + (long long)i * kTN;
      // only tile 0 is requested ahead of Q^: the rest of the ring would otherwise sit in front of the Q^ fetch on the way into
      // the SM (infonce_tc.cu, RMCL_TC_TILES_BEFORE_Q; the split path's Q^ rows are twice as wide)
      if (RMCL_TC2_Q_FIRST && i == 1) mbar_wait(&sh.q_full, 0);
      for (int c = 0; c < kChunks; ++c) {
        const int it = i * kChunks + c, st = it % kStages;
        mbar_wait(&sh.k_empty[st], ((it / kStages) & 1) ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&sh.k_full[st], kChunkBytes);
          // SPLIT: the tensor map covers the [2C, K] hi/lo buffer; chunk order hi (row 0), lo (row C), hi (row 0)
          tma_load_2d(ring + (size_t)st * kChunkBytes, &tmap_queue, &sh.k_full[st], (int)col0,
                      SPLIT ? (c == 1 ? C : 0) : c * kChunkRows);
        }
        __syncwarp();
      }
    }
  } else {
    // ===================================================================== MMA issuer
    mbar_wait(&sh.q_full, 0);
    tc_fence_after();
    for (int i = 0; i < n_tiles; ++i) {
      if (i >= 2) mbar_wait(&sh.s_free[i & 1], ((i - 2) >> 1) & 1);
      tc_fence_after();
      const uint32_t d = tmem + kTmS + (i & 1) * kTN;
      for (int c = 0; c < kChunks; ++c) {
        const int it = i * kChunks + c, st = it % kStages;
        mbar_wait(&sh.k_full[st], (it / kStages) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sbase = smem_u32(ring + (size_t)st * kChunkBytes);
#pragma unroll
          for (int s = 0; s < kChunkRows / 16; ++s) {
            // B = chunk as [N=64 columns][K=16 rows of C], MN-major: 8-row groups 1024 B apart
            const uint64_t bd = make_sw128_desc(sbase + s * 2048, kChunkBytes, 1024);
            // SPLIT: A = q_hi for the hi and lo planes, q_lo (TMEM columns C/2...) for the second hi pass
            const uint32_t a_col = SPLIT ? (c == 2 ? (uint32_t)(C / 2) : 0u) : (uint32_t)(c * (kChunkRows / 2));
            tc_mma_ts(d, tmem + kTmQ + a_col + s * 8, bd, kIdescS, (c > 0 || s > 0) ? 1u : 0u);
          }
          tc_commit(&sh.k_empty[st]);                       // this chunk stage may be refilled
          if (c == kChunks - 1) tc_commit(&sh.s_full[i & 1]);
        }
        __syncwarp();
      }
    }
  }

  __syncthreads();
  if (tid == 0) pdl_trigger();
  if (warp == 8) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

// ========================================================================================== PV pass
// ring depth of the PV pass: as many stages as fit in ~192 KB
__host__ __device__ constexpr int pv_stages(int slice, bool split) {
  const int stage = split ? 2 * (kRows * kTN * 2 + slice * kTN * 2) : (kRows * kTN * 2 + slice * kTN * 2);
  const int n = (192 * 1024) / stage;
  return n > 4 ? 4 : n;
}

struct PvShared {
  uint64_t full[4];
  uint64_t empty[4];
  uint64_t o_done;
  uint32_t tmem_base;
};

// SLICE = output columns per CTA (256; SPLIT: the whole logical width C in {64,128,256}).  SPLIT stage =
// [P_hi 16 KB | P_lo 16 KB | Q_hi SLICE x 64 | Q_lo SLICE x 64], three MMAs per K = 16 step, fp32 partials.
template <int SLICE, bool SPLIT>
__global__ void __launch_bounds__(kThreads, 1)
    infonce_pv_kernel(const __grid_constant__ CUtensorMap tmap_p, const __grid_constant__ CUtensorMap tmap_queue, int B, int C,
                      int b_pad, long long K, long long cols_per_split, void* __restrict__ po_raw) {
  constexpr int kSlice = SLICE;
  constexpr int kPBytes = kRows * kTN * 2;               // 16 KB: P~ tile, K-major SWIZZLE_128B rows
  constexpr int kQBytes = kSlice * kTN * 2;              // queue slice tile
  constexpr int kStageBytes = SPLIT ? 2 * (kPBytes + kQBytes) : kPBytes + kQBytes;
  constexpr int kStages = pv_stages(SLICE, SPLIT);
  constexpr int HC = kSlice / 2;
  constexpr int kOutBytes = SPLIT ? 4 : 2;               // fp32 partials on the fp32-accurate path
  constexpr int kRowStage = kOutBytes * kSlice + 16;
  constexpr uint32_t kIdescO = make_idesc(128, kSlice, 0);
  constexpr uint32_t kTmemCols = kSlice < 32 ? 32 : kSlice;
  static_assert(kRows * kRowStage <= kStages * kStageBytes, "epilogue staging must fit the ring");

  extern __shared__ uint8_t smem_raw[];
  __shared__ PvShared sh;
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // slice fastest, then row block, then split: the CTAs resident together share their P~ tiles (C/256
  // slices) and their queue tiles (all row blocks) through L2
  const int slice = blockIdx.x, split = blockIdx.z;
  const int row0 = blockIdx.y * kRows;
  const long long k_begin = (long long)split * cols_per_split;
  const long long k_end = (k_begin + cols_per_split < K) ? k_begin + cols_per_split : K;
  const int n_tiles = (int)((k_end - k_begin + kTN - 1) / kTN);

  if (tid == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&sh.full[i], 1);
      mbar_init(&sh.empty[i], 1);
    }
    mbar_init(&sh.o_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_p) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_queue) : "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh.tmem_base)), "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sh.tmem_base;
  pdl_wait();   // P~ is written by the S pass

  if (warp == 8) {
    for (int i = 0; i < n_tiles; ++i) {
      const int st = i % kStages;
      mbar_wait(&sh.empty[st], ((i / kStages) & 1) ^ 1);
      if (elect_one()) {
        const long long col0 = k_begin + (long long)i * kTN;
        uint8_t* base = ring + (size_t)st * kStageBytes;
        mbar_expect_tx(&sh.full[st], kStageBytes);
        if (SPLIT) {   // planes: P~ lo follows b_pad rows of hi; queue lo follows C rows of hi
          tma_load_2d(base, &tmap_p, &sh.full[st], (int)col0, row0);
          tma_load_2d(base + kPBytes, &tmap_p, &sh.full[st], (int)col0, b_pad + row0);
          tma_load_2d(base + 2 * kPBytes, &tmap_queue, &sh.full[st], (int)col0, 0);
          tma_load_2d(base + 2 * kPBytes + kQBytes, &tmap_queue, &sh.full[st], (int)col0, C);
        } else {
          tma_load_2d(base, &tmap_p, &sh.full[st], (int)col0, row0);
          tma_load_2d(base + kPBytes, &tmap_queue, &sh.full[st], (int)col0, slice * kSlice);
        }
      }
      __syncwarp();
    }
  } else if (warp == 9) {
    for (int i = 0; i < n_tiles; ++i) {
      const int st = i % kStages;
      mbar_wait(&sh.full[st], (i / kStages) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t pa = smem_u32(ring + (size_t)st * kStageBytes);
#pragma unroll
        for (int s = 0; s < kTN / 16; ++s) {
          // A = P~ tile [M=128 rows][K=16 columns], B = queue slice [N=SLICE rows of C][K=16 columns]; both K-major
          if (SPLIT) {
            const uint64_t a_hi = make_sw128_desc(pa + s * 32, 16, 1024), a_lo = make_sw128_desc(pa + kPBytes + s * 32, 16, 1024);
            const uint64_t b_hi = make_sw128_desc(pa + 2 * kPBytes + s * 32, 16, 1024);
            const uint64_t b_lo = make_sw128_desc(pa + 2 * kPBytes + kQBytes + s * 32, 16, 1024);
            tc_mma_ss(tmem, a_hi, b_hi, kIdescO, (i > 0 || s > 0) ? 1u : 0u);
            tc_mma_ss(tmem, a_hi, b_lo, kIdescO, 1u);
            tc_mma_ss(tmem, a_lo, b_hi, kIdescO, 1u);
          } else {
            const uint64_t ad = make_sw128_desc(pa + s * 32, 16, 1024);
            const uint64_t bd = make_sw128_desc(pa + kPBytes + s * 32, 16, 1024);
            tc_mma_ss(tmem, ad, bd, kIdescO, (i > 0 || s > 0) ? 1u : 0u);
          }
        }
        tc_commit(&sh.empty[st]);
        if (i == n_tiles - 1) tc_commit(&sh.o_done);
      }
      __syncwarp();
    }
  } else {
    // epilogue warps: O rows -> bf16 (SPLIT: kept fp32) -> own staging segment -> one bulk copy per row half (as infonce_tc.cu)
    const int quad = warp & 3, par = warp >> 2;
    const int r = quad * 32 + lane;
    const uint32_t tlane = tmem + ((uint32_t)(quad * 32) << 16);
    mbar_wait(&sh.o_done, 0);
    tc_fence_after();
    uint8_t* stage = ring + (size_t)r * kRowStage + par * HC * kOutBytes;
#pragma unroll 1
    for (int ch = 0; ch < HC / 32; ++ch) {
      uint32_t o[32];
      tc_ld32(tlane + par * HC + ch * 32, o);
      tc_wait_ld();
      if (SPLIT) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<uint4*>(stage + ch * 128 + 16 * j) = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
      } else {
        uint32_t h[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const __nv_bfloat162 pk = __floats2bfloat162_rn(__uint_as_float(o[2 * j]), __uint_as_float(o[2 * j + 1]));
          h[j] = *reinterpret_cast<const uint32_t*>(&pk);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<uint4*>(stage + ch * 64 + 16 * j) = make_uint4(h[4 * j], h[4 * j + 1], h[4 * j + 2], h[4 * j + 3]);
      }
    }
    if (row0 + r < B) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      uint8_t* dst = reinterpret_cast<uint8_t*>(po_raw) +
                     ((((size_t)split * B + row0 + r) * C + (size_t)slice * kSlice + par * HC) * kOutBytes);
      bulk_store_row(dst, stage, HC * kOutBytes);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    tc_fence_before();
  }

  __syncthreads();
  if (tid == 0) pdl_trigger();
  if (warp == 8) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
  }
}

template <int C, bool DIAG, bool SPLIT>
int launch_s(const CUtensorMap& tq, const __nv_bfloat16* q_hat, int B, long long K, long long k_pad, float scale2,
             const InfoNcePlan& p, InfoNcePartials out, __nv_bfloat16* ptilde, unsigned int* flag, int want_argmax,
             cudaStream_t s) {
  const size_t chunk_bytes = (size_t)(SPLIT ? C : kChunkRows) * kTN * 2;
  const size_t smem = 8 * 4096 + 5 * chunk_bytes + 1024;
  auto kern = infonce_s_kernel<C, DIAG, SPLIT>;
  RMCL_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  RMCL_CUDA_OK(launch_pdl(kern, dim3(p.splits, p.row_blocks), dim3(kThreads), smem, s, tq, q_hat, B, K, k_pad, scale2,
                          p.cols_per_split, want_argmax, out.m, out.l, out.av, out.ai, ptilde, flag, out.n2, out.qn2, out.dist));
  return RMCL_OK;
}

template <int SLICE, bool SPLIT>
int launch_pv(const CUtensorMap& tp, const CUtensorMap& tq, int B, int C, long long K, const InfoNcePlan& p, void* po,
              cudaStream_t s) {
  constexpr int kStageBytes = SPLIT ? 2 * (kRows * kTN * 2 + SLICE * kTN * 2) : (kRows * kTN * 2 + SLICE * kTN * 2);
  const size_t smem = (size_t)pv_stages(SLICE, SPLIT) * kStageBytes + 1024;
  auto kern = infonce_pv_kernel<SLICE, SPLIT>;
  RMCL_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  RMCL_CUDA_OK(launch_pdl(kern, dim3(C / SLICE, p.row_blocks, p.splits), dim3(kThreads), smem, s, tp, tq, B, C, p.b_pad, K,
                          p.cols_per_split, po));
  return RMCL_OK;
}

}  // namespace

bool infonce_tc2_supports(int C) { return C == 512 || C == 768; }

bool infonce_tc2_split_supports(int C) { return C == 64 || C == 128 || C == 256; }

#define RMCL_S_DISPATCH(CC, SP)                                                                                            \
  (dg ? launch_s<CC, true, SP>(tq, q_hat, B, K, k_pad, scale2, p, out, ptilde_s, overflow_flag, want_argmax, s)          \
      : launch_s<CC, false, SP>(tq, q_hat, B, K, k_pad, scale2, p, out, ptilde_s, overflow_flag, want_argmax, s))

// split = false: queue is bf16 [C, K], q_hat rows of C.  split = true: queue is the [2C, K] hi/lo buffer (RMCL_BF16_HILO),
// q_hat rows are [q_hi | q_lo] (2C wide), P~ has two planes and the partials in out.o are fp32.
int infonce_tc2_launch(const __nv_bfloat16* q_hat, const void* queue, int B, int C, long long K, long long ldq, float scale2,
                       const InfoNcePlan& p, InfoNcePartials out, __nv_bfloat16* ptilde, long long k_pad,
                       unsigned int* overflow_flag, int want_argmax, int want_o, bool split, cudaStream_t s) {
  __nv_bfloat16* ptilde_s = want_o ? ptilde : nullptr;   // statistics-only call: the S pass alone, no P~
  if (p.row_blocks > 65535 || p.splits > 65535) {
    set_error("InfoNCE: too many rows (%d)", B);
    return RMCL_E_UNSUPPORTED_DIM;
  }
  alignas(64) CUtensorMap tq, tp;
  int rc = make_tmap_bf16(&tq, queue, (uint64_t)(split ? 2 * C : C), (uint64_t)K, (uint64_t)ldq, split ? C : kChunkRows);
  if (rc != RMCL_OK) return rc;
  rc = make_tmap_bf16(&tp, ptilde, (uint64_t)(split ? 2 : 1) * p.b_pad, (uint64_t)k_pad, (uint64_t)k_pad, kRows);
  if (rc != RMCL_OK) return rc;
  const bool dg = out.n2 != nullptr;
  if (split) {
    if (C == 256) rc = RMCL_S_DISPATCH(256, true);
    else if (C == 128) rc = RMCL_S_DISPATCH(128, true);
    else if (C == 64) rc = RMCL_S_DISPATCH(64, true);
    else {
      set_error("fp32-accurate tcgen05 InfoNCE supports C in {64, 128, 256} (got %d)", C);
      return RMCL_E_UNSUPPORTED_DIM;
    }
  } else if (C == 768)
    rc = RMCL_S_DISPATCH(768, false);
  else if (C == 512)
    rc = RMCL_S_DISPATCH(512, false);
  else {
    set_error("two-pass tcgen05 InfoNCE supports C in {512, 768} (got %d)", C);
    return RMCL_E_UNSUPPORTED_DIM;
  }
  if (rc != RMCL_OK) return rc;
  if (!want_o) return RMCL_OK;
  if (!split) return launch_pv<256, false>(tp, tq, B, C, K, p, out.o, s);
  if (C == 256) return launch_pv<256, true>(tp, tq, B, C, K, p, out.o, s);
  if (C == 128) return launch_pv<128, true>(tp, tq, B, C, K, p, out.o, s);
  return launch_pv<64, true>(tp, tq, B, C, K, p, out.o, s);
}

}  // namespace rmcl
