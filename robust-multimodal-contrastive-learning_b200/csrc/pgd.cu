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
// HBM-bound, and the per-sample norm makes it a two-phase computation: the gradient is needed once
// for the norm and once for the update.  One persistent launch walks an ordered list of work items
// (sample, 32 KB chunk) handed out by an atomic ticket:
//     round r:   P1(batch r)    read g chunk (L2: also delta) -> partial norm(s) of the chunk
//                P2(batch r-1)  re-read g, read delta         -> write delta'
// The L2 projection needs |delta + a g|_2 (a = lr/|g|_2) before anything is written; it is taken from
//     |delta + a g|^2 = |delta|^2 + 2a <delta,g> + a^2 |g|^2
// so P1 reduces three sums and delta' is written exactly once, already projected (a first version wrote
// the unprojected delta', reduced its norm and rescaled it in a third phase: 16 B/element, 47 % of roof).
// A batch is a group of samples whose gradients fit in a fraction of the 126 MB L2, so that the
// re-read of P2 is served by L2 and DRAM sees the algorithmic minimum of 12 B/element
// (read g, read delta, write delta).  A later-phase item spins on a per-sample arrival counter; all
// the items it waits for have smaller tickets, i.e. are already running, so the wait cannot deadlock
// regardless of how many CTAs are resident.  Per-chunk partials are combined by a fixed-shape tree, so
// the result is deterministic (and the inf-norm, hence REF_LINF, is exact).
//
// Round 2, two alternatives to the L2-resident scheme were built and measured on B200 (profiles/r2_pgd_staged_experiment.txt)
// and dropped: (a) a persistent one-CTA-per-SM pipeline that keeps the chunks in SHARED memory between norm and update
// (bulk-async loads, tagged per-chunk partials polled by a control warp): 152-240 us against 147 us for the pixel ref_linf
// case — the 30 MB of shared memory on the chip hold only ~7 us of input, about the residency a chunk needs (load latency +
// skew between the ~75 CTAs that share a sample + a 2.6 k-cycle L2 poll under DRAM load + update), so every CTA runs in
// lock-step with its neighbours; (b) this kernel with its operands staged through a double-buffered shared-memory stage by
// cp.async.bulk one item ahead: 240 us — one item in flight per CTA is fewer bytes in flight than the register path has, and
// the per-item latency chains (ticket, partial publish, totals) no longer overlap across 5 CTAs x 8 warps.
#include "common.cuh"

namespace rmcl {

#ifndef RMCL_PGD_PREFETCH
#define RMCL_PGD_PREFETCH 1
#endif
#ifndef RMCL_PGD_CHUNK_KB
#define RMCL_PGD_CHUNK_KB 32
#endif
// resident CTAs per SM asked of ptxas for the fp32/fp32 kernel (48 registers): the loads in flight, not the
// arithmetic, set the speed — at 64 registers / 4 CTAs the same code ran 16 % slower
#ifndef RMCL_PGD_MIN_CTAS
#define RMCL_PGD_MIN_CTAS 5
#endif
// independent 16-byte steps per thread in the update phase (2: 4 loads in flight per thread)
// Timing experiments (results are garbage): bit 0 update items do not wait for the sample totals, 1 norm items publish
// nothing, 2 the update is not stored, 3 norm items do nothing, 4 update items do nothing (profiles/r2_pgd_experiments.txt)
#ifndef RMCL_PGD_EXPERIMENT
#define RMCL_PGD_EXPERIMENT 0
#endif
#ifndef RMCL_PGD_P1_UNROLL
#define RMCL_PGD_P1_UNROLL 4
#endif
#ifndef RMCL_PGD_PRELOAD
#define RMCL_PGD_PRELOAD 1
#endif
#ifndef RMCL_PGD_UPDATE_UNROLL
#define RMCL_PGD_UPDATE_UNROLL 2
#endif
constexpr int kPgdThreads = 256;
constexpr int kPgdChunkBytes = RMCL_PGD_CHUNK_KB * 1024;                 // per operand per work item
#ifndef RMCL_PGD_BATCH_MB
#define RMCL_PGD_BATCH_MB 24
#endif
#ifndef RMCL_PGD_BATCH_MB_L2
#define RMCL_PGD_BATCH_MB_L2 40
#endif
// bytes per batch kept L2-resident between the two phases.  Same-box sweep (tools/pgd_time.py): batches smaller than the
// ~740 resident CTAs' worth of work leave CTAs spinning on norms (8 MB: 199 us, 12 MB: 152 us, 24 MB: 136 us, 40 MB: 142 us
// for the pixel ref_linf case); the L2-projection mode keeps g AND delta resident, i.e. half as many samples per byte, and
// wants the larger batch (24 MB: 176 us, 40 MB: 168 us).
constexpr long long kPgdBatchBytes = (long long)RMCL_PGD_BATCH_MB * 1024 * 1024;
constexpr long long kPgdBatchBytesL2 = (long long)RMCL_PGD_BATCH_MB_L2 * 1024 * 1024;

// max|g| with torch.norm(p=inf)'s NaN behaviour: |x| and the running maximum are non-negative floats or NaNs with a clear
// sign bit, and on those bit patterns the unsigned integer order is the float order with every NaN above +inf — so an integer
// maximum propagates a NaN gradient into the sample's norm (fmaxf would drop it), as ATen does.
__device__ __forceinline__ float absmax_nan(float acc, float x) {
  const unsigned a = __float_as_uint(acc), b = __float_as_uint(x) & 0x7fffffffu;
  return __uint_as_float(a > b ? a : b);
}
__device__ __forceinline__ float warp_absmax_nan(float v) {
  return __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(v)));
}
// torch.clamp(x, min=lo) keeps a NaN
__device__ __forceinline__ float clamp_min_nan(float x, float lo) { return (x != x) ? x : fmaxf(x, lo); }

__device__ __forceinline__ float block_reduce(float v, bool is_max, float* red /*[32]*/) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = is_max ? warp_absmax_nan(v) : warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float r = (lane < nw) ? red[lane] : 0.f;
  r = is_max ? warp_absmax_nan(r) : warp_sum(r);
  return r;  // valid in every thread
}

// One element of the update, given the per-sample denominator.
template <typename TD>
__device__ __forceinline__ float pgd_apply(float delta, float g, float lr, float denom, int mode);
template <>
__device__ __forceinline__ float pgd_apply<float>(float delta, float g, float lr, float denom, int mode) {
  float step;
  if (mode == RMCL_PGD_SIGN_LINF) {
    const float sg = (g > 0.f) ? 1.f : ((g < 0.f) ? -1.f : 0.f);   // torch.sign: 0 for +-0 and for NaN
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
  return (v != v) ? v : fminf(fmaxf(v, -eps), eps);   // ATen's clamp keeps a NaN
}
template <> __device__ __forceinline__ float clamp_eps<__nv_bfloat16>(float v, float eps) {
  const float e = __bfloat162float(__float2bfloat16_rn(eps));  // ATen casts the bound to the tensor dtype
  return (v != v) ? v : fminf(fmaxf(v, -e), e);
}

struct PgdPlan {
  long long N;
  int B;
  int mode;
  float lr, eps;
  int chunk_elems;       // elements per work item
  int chunks;            // work items per sample and phase
  int batch;             // samples per batch
  int n_batches;
  int phases;            // 1 (sign: update only) or 2 (norm, update)
  int n_sums;            // partial sums per chunk: 1 (max|g| or sum g^2) or 3 (+ <delta,g>, |delta|^2 for the L2 projection)
  long long total_items;
  // workspace
  unsigned int* ticket;  // [1]
  unsigned int* exits;   // [1] CTAs that have left the work loop
  unsigned int* done;    // [B] chunks of the sample whose norm phase has finished
  float* part;           // [3][B][chunks] per-chunk partials
  unsigned long long* words;  // [B][2] {1 : 32 | fp32 bits : 32}: denominator, projection scale of the sample — flag and value in
                              // one 64-bit word, so a waiter needs ONE L2 round trip (2-2.6 k cycles under DRAM load) instead of
                              // flag-then-totals, and the double-precision combination is done once per sample, not per chunk
};

struct PgdItem {
  int phase;  // 0 = norm, 1 = update
  int sample;
  int chunk;
};

// ticket -> (phase, sample, chunk) in the round order described at the top of the file.
// RMCL_PGD_INTERLEAVE: inside a round the items of the two phases alternate (P1 item, P2 item, P1 item, ...) instead of
// all P1 items followed by all P2 items, so that at any moment the resident CTAs are a mix of DRAM-streaming norm items
// and L2-served update items rather than all in the same phase.  The dependency argument is unchanged: a P2 item of batch
// r-1 only waits for P1 items of batch r-1, which belong to the previous round, i.e. have smaller tickets.
// Same-box sweep (tools/pgd_sweep.py, profiles/r2_pgd_sweep.txt): interleaved 142.8 / 171.0 us against 145.3 / 179.1 us for the
// pixel ref_linf / l2 cases.
#ifndef RMCL_PGD_INTERLEAVE
#define RMCL_PGD_INTERLEAVE 1
#endif
__device__ __forceinline__ PgdItem pgd_decode(const PgdPlan& p, long long t) {
  const int first_phase = (p.phases == 1) ? 1 : 0;  // sign mode has no norm phase
  const int rounds = p.n_batches + p.phases - 1;
  for (int r = 0; r < rounds; ++r) {
    long long cnt[2] = {0, 0};
    int s0[2] = {0, 0};
    for (int k = 0; k < p.phases; ++k) {
      const int b = r - k;
      if (b < 0 || b >= p.n_batches) continue;
      s0[k] = b * p.batch;
      const int ns = (s0[k] + p.batch <= p.B ? p.batch : p.B - s0[k]);
      cnt[k] = (long long)ns * p.chunks;
    }
    const long long tot = cnt[0] + cnt[1];
    if (t >= tot) {
      t -= tot;
      continue;
    }
    int k;
    long long idx;
#if RMCL_PGD_INTERLEAVE
    const long long both = cnt[0] < cnt[1] ? cnt[0] : cnt[1];
    if (t < 2 * both) {
      k = (int)(t & 1);
      idx = t >> 1;
    } else {
      k = cnt[0] > cnt[1] ? 0 : 1;
      idx = t - both;
    }
#else
    if (t < cnt[0]) {
      k = 0;
      idx = t;
    } else {
      k = 1;
      idx = t - cnt[0];
    }
#endif
    return PgdItem{first_phase + k, s0[k] + (int)(idx / p.chunks), (int)(idx % p.chunks)};
  }
  return PgdItem{-1, 0, 0};
}

// Bytes in flight: a CTA's register-held loads (4 x 16 B per thread) come to ~80 KB per SM at 5 CTAs, against the ~45 KB per SM
// that keep 6.5 TB/s busy at ~1 us of DRAM latency — no slack for the update phase, where half of those loads are L2 hits.
// RMCL_PGD_L2_PREFETCH: the CTA already holds its NEXT ticket while it works on the current item (see the work loop), so
// thread 0 asks the L2 for the next item's DRAM-resident operand(s) with one cp.async.bulk.prefetch.L2 per operand: no
// registers, no shared memory, and the item's own loads then hit in L2.
// Measured (profiles/r2_pgd_sweep.txt): it helps the single-phase sign mode (36.5 -> 34.5 us, 97-98 % of the HBM roof) and
// HURTS the two-phase modes (pixel ref_linf 146 -> 182 us: the prefetched lines compete with the evict_last residents the
// second phase depends on), so it is compiled in only for the mode it helps.
#ifndef RMCL_PGD_L2_PREFETCH
#define RMCL_PGD_L2_PREFETCH 1
#endif
__device__ __forceinline__ void l2_prefetch_bulk(const void* p, unsigned bytes, uint64_t policy) {
  asm volatile("cp.async.bulk.prefetch.L2.global.L2::cache_hint [%0], %1, %2;" ::"l"(p), "r"(bytes), "l"(policy) : "memory");
}

__device__ __forceinline__ unsigned long long ld_word(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.b64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_word(unsigned long long* p, float val) {
  const unsigned long long v = (1ull << 32) | (unsigned long long)__float_as_uint(val);
  asm volatile("st.relaxed.gpu.global.b64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Stores this item's partials; the item that arrives last for its sample combines all the sample's
// partials (thread c takes chunks c, c+256, ... then a fixed-shape block tree: deterministic no matter
// which CTA happens to be last) and publishes the sample totals, followed by one more arrival, so that
// waiters need a single load.  Must be called by the whole CTA.
__device__ __forceinline__ void pgd_publish_partials(const PgdPlan& p, const PgdItem& it, float v0, float v1, float v2,
                                                     bool is_max, float* red) {
  __shared__ bool s_last;
  const long long plane = (long long)p.B * p.chunks;
  if (threadIdx.x == 0) {
    float* dst = p.part + (long long)it.sample * p.chunks + it.chunk;
    dst[0] = v0;
    if (p.n_sums == 3) {
      dst[plane] = v1;
      dst[2 * plane] = v2;
    }
    __threadfence();
    s_last = (atomicAdd(p.done + it.sample, 1u) == (unsigned)p.chunks - 1u);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  float tot[3] = {0.f, 0.f, 0.f};
  for (int k = 0; k < p.n_sums; ++k) {
    const float* pp = p.part + k * plane + (long long)it.sample * p.chunks;
    float acc = 0.f;
    for (int c = threadIdx.x; c < p.chunks; c += kPgdThreads) {
      const float v = __ldcg(pp + c);
      acc = is_max ? absmax_nan(acc, v) : acc + v;
    }
    tot[k] = block_reduce(acc, is_max, red);
  }
  if (threadIdx.x == 0) {
    const float dn = clamp_min_nan(is_max ? tot[0] : sqrtf(tot[0]), 1e-8f);
    float pj = 1.f;
    if (p.n_sums == 3) {
      // |delta + a g|^2 = |delta|^2 + 2a <delta,g> + a^2 |g|^2 with a = lr/denom, combined in double (three fp32 totals; the
      // cross term may cancel)
      const double a = (double)p.lr / (double)dn;
      double n2 = (double)tot[2] + 2.0 * a * (double)tot[1] + a * a * (double)tot[0];
      if (n2 < 0.0) n2 = 0.0;
      pj = fminf(__fdiv_rn(p.eps, fmaxf((float)sqrt(n2), 1e-12f)), 1.f);
    }
    st_word(p.words + 2 * (long long)it.sample, dn);
    st_word(p.words + 2 * (long long)it.sample + 1, pj);
  }
}

template <typename TD, typename TG>
__global__ void __launch_bounds__(kPgdThreads, (sizeof(TD) + sizeof(TG) == 8) ? RMCL_PGD_MIN_CTAS : 4) pgd_ticket_kernel(TD* __restrict__ delta, const TG* __restrict__ grad,
                                                                 const PgdPlan p) {
  __shared__ float red[32];
  __shared__ PgdItem s_item;
  __shared__ float s_norm[2];   // denominator, projection scale of the current item's sample
  constexpr int VG = 16 / sizeof(TG), VD = 16 / sizeof(TD);
  constexpr int VE = VG > VD ? VG : VD;  // elements per thread step in the vector path (8 if any bf16, else 4)
  const bool is_max = (p.mode == RMCL_PGD_REF_LINF);
  const bool vec = (p.N % VE == 0) && (p.chunk_elems % VE == 0) &&
                   (((reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(delta)) & 15u) == 0);
  const bool l2proj = (p.n_sums == 3);
  const bool do_clamp = (p.eps > 0.f) && (p.mode != RMCL_PGD_L2);
  // g (and delta under the L2 projection) is read twice a few tens of MB apart: keep it (evict_last)
  // until its second use, then let it go; everything touched once must not push it out of L2.
  const uint64_t keep = l2_policy_evict_last(), stream = l2_policy_evict_first();

  // The next ticket is always requested one item ahead, so the atomic's L2 round trip overlaps the
  // data phase of the current item.  A ticket held this way is still started only after every smaller
  // ticket taken by this CTA has finished, so the "everything I wait for is already running or will
  // run without waiting for me" argument (top of file) is unchanged.
  long long t_next = 0;
  bool first = true;
  (void)first;
  if (threadIdx.x == 0) t_next = (long long)atomicAdd(p.ticket, 1u);
  for (;;) {
    __syncthreads();
#if !RMCL_PGD_PREFETCH
    if (threadIdx.x == 0 && !first) t_next = (long long)atomicAdd(p.ticket, 1u);
    first = false;
#endif
    if (threadIdx.x == 0) s_item = (t_next < p.total_items) ? pgd_decode(p, t_next) : PgdItem{-1, 0, 0};
    __syncthreads();
    const PgdItem it = s_item;
    if (it.phase < 0) {
      // The last CTA to leave returns the control words to zero, so the next call needs no memset launch
      // (every item has completed by then: a CTA only gets here after finishing all its items).
      if (threadIdx.x == 0 && atomicAdd(p.exits, 1u) == gridDim.x - 1u) {
        for (int i = 0; i < p.B; ++i) {
          p.done[i] = 0u;
          p.words[2 * i] = 0ull;
          p.words[2 * i + 1] = 0ull;
        }
        *p.exits = 0u;
        __threadfence();
        *p.ticket = 0u;
      }
      break;
    }
#if RMCL_PGD_PREFETCH
    if (threadIdx.x == 0) {
      t_next = (long long)atomicAdd(p.ticket, 1u);
#if RMCL_PGD_L2_PREFETCH
      // 1: the single-phase (sign) mode only, see above; 2: also the NORM items of the two-phase modes (not their update items)
      if (vec && t_next < p.total_items && (p.phases == 1 || RMCL_PGD_L2_PREFETCH >= 2)) {
        const PgdItem nx = pgd_decode(p, t_next);
        const long long ne0 = (long long)nx.chunk * p.chunk_elems;
        long long nn = p.N - ne0;
        if (nn > p.chunk_elems) nn = p.chunk_elems;
        const TG* ng = grad + (long long)nx.sample * p.N + ne0;
        const TD* nd = delta + (long long)nx.sample * p.N + ne0;
        if (nx.phase == 0) {            // norm item: g comes from DRAM (and delta too under the L2 projection); both are re-read later
          l2_prefetch_bulk(ng, (unsigned)(nn * sizeof(TG)), keep);
          if (l2proj) l2_prefetch_bulk(nd, (unsigned)(nn * sizeof(TD)), keep);
        } else if (!l2proj && p.phases == 1) {   // update item: g is L2-resident already, delta comes from DRAM, used once
          l2_prefetch_bulk(nd, (unsigned)(nn * sizeof(TD)), stream);
        }
      }
#endif
    }
#endif
    const long long e0 = (long long)it.chunk * p.chunk_elems;
    long long n = p.N - e0;
    if (n > p.chunk_elems) n = p.chunk_elems;
    const TG* g = grad + (long long)it.sample * p.N + e0;
    TD* d = delta + (long long)it.sample * p.N + e0;

#if RMCL_PGD_EXPERIMENT & 8    // timing experiment: norm items do nothing (update items then read everything from DRAM)
    if (it.phase == 0) continue;
#endif
#if RMCL_PGD_EXPERIMENT & 16   // timing experiment: update items do nothing
    if (it.phase != 0) continue;
#endif
    if (it.phase == 0) {
      // ------------------------------------------------------------ P1: partial norm(s) of the chunk
      float acc = 0.f, adg = 0.f, add = 0.f;
      if (vec && !l2proj) {
        const uint4* gv = reinterpret_cast<const uint4*>(g);
        const long long nv = n / VG;
        long long i = threadIdx.x;
        constexpr int kU = RMCL_PGD_P1_UNROLL;   // independent 16-byte loads in flight per thread in the norm phase
        for (; i + (kU - 1) * kPgdThreads < nv; i += kU * kPgdThreads) {
          uint4 u[kU];
#pragma unroll
          for (int k = 0; k < kU; ++k) u[k] = ld_u4_hint(gv + i + k * kPgdThreads, keep);
#pragma unroll
          for (int k = 0; k < kU; ++k) {
            const TG* e = reinterpret_cast<const TG*>(&u[k]);
#pragma unroll
            for (int j = 0; j < VG; ++j) {
              const float x = to_f32(e[j]);
              acc = is_max ? absmax_nan(acc, x) : fmaf(x, x, acc);
            }
          }
        }
        for (; i < nv; i += kPgdThreads) {
          const uint4 u = ld_u4_hint(gv + i, keep);
          const TG* e = reinterpret_cast<const TG*>(&u);
#pragma unroll
          for (int j = 0; j < VG; ++j) {
            const float x = to_f32(e[j]);
            acc = is_max ? absmax_nan(acc, x) : fmaf(x, x, acc);
          }
        }
      } else if (vec) {
        // L2 with projection: |g|^2, <delta,g>, |delta|^2 in one sweep; both operands stay in L2 for P2
        const long long nv = n / VE;
        for (long long i = threadIdx.x; i < nv; i += 2 * kPgdThreads) {
          const bool second = (i + kPgdThreads) < nv;
          float gx[2][VE], dx[2][VE];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (h == 1 && !second) break;
            const long long base = (i + h * kPgdThreads) * VE;
#pragma unroll
            for (int q = 0; q < VE / VG; ++q) {
              const uint4 u = ld_u4_hint(reinterpret_cast<const uint4*>(g + base) + q, keep);
              const TG* e = reinterpret_cast<const TG*>(&u);
#pragma unroll
              for (int j = 0; j < VG; ++j) gx[h][q * VG + j] = to_f32(e[j]);
            }
#pragma unroll
            for (int q = 0; q < VE / VD; ++q) {
              const uint4 u = ld_u4_hint(reinterpret_cast<const uint4*>(d + base) + q, keep);
              const TD* e = reinterpret_cast<const TD*>(&u);
#pragma unroll
              for (int j = 0; j < VD; ++j) dx[h][q * VD + j] = to_f32(e[j]);
            }
          }
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (h == 1 && !second) break;
#pragma unroll
            for (int j = 0; j < VE; ++j) {
              acc = fmaf(gx[h][j], gx[h][j], acc);
              adg = fmaf(dx[h][j], gx[h][j], adg);
              add = fmaf(dx[h][j], dx[h][j], add);
            }
          }
        }
      } else {
        for (long long i = threadIdx.x; i < n; i += kPgdThreads) {
          const float x = to_f32(g[i]);
          acc = is_max ? absmax_nan(acc, x) : fmaf(x, x, acc);
          if (l2proj) {
            const float y = to_f32(d[i]);
            adg = fmaf(y, x, adg);
            add = fmaf(y, y, add);
          }
        }
      }
      acc = block_reduce(acc, is_max, red);
      if (l2proj) {
        adg = block_reduce(adg, false, red);
        add = block_reduce(add, false, red);
      }
#if RMCL_PGD_EXPERIMENT & 2     // timing experiment: partials are not published (no global writes, no atomic)
      if (acc == 123.456f && adg == 1.f && add == 2.f) pgd_publish_partials(p, it, acc, adg, add, is_max, red);
#else
      pgd_publish_partials(p, it, acc, adg, add, is_max, red);
#endif
    } else {
      // ------------------------------------------------------------ P2: the update
      float denom = 1.f, proj = 1.f;
      // The first kH steps of the update are requested BEFORE the totals are waited for: they do not depend on them, and the
      // L2 round trip of the wait (2-2.6 k cycles under DRAM load) then overlaps the first DRAM access instead of preceding it.
      constexpr int kH = RMCL_PGD_UPDATE_UNROLL;
      const long long nv = n / VE;
      uint4 pg[kH][VE / VG], pd[kH][VE / VD];
      if (vec && RMCL_PGD_PRELOAD) {
#pragma unroll
        for (int h = 0; h < kH; ++h) {
          if ((long long)threadIdx.x + h * kPgdThreads >= nv) break;
          const long long base = ((long long)threadIdx.x + h * kPgdThreads) * VE;
#pragma unroll
          for (int q = 0; q < VE / VG; ++q) pg[h][q] = ld_u4_hint(reinterpret_cast<const uint4*>(g + base) + q, stream);
#pragma unroll
          for (int q = 0; q < VE / VD; ++q) pd[h][q] = ld_u4_hint(reinterpret_cast<const uint4*>(d + base) + q, stream);
        }
      }
      if (p.mode != RMCL_PGD_SIGN_LINF && !(RMCL_PGD_EXPERIMENT & 1)) {   // experiment bit 0: no wait for the totals
        if (threadIdx.x == 0) {
          const unsigned long long* w = p.words + 2 * (long long)it.sample;
          unsigned long long w0, w1;
          const long long t0 = clock64();
          for (;;) {   // both words in flight together: one round trip once the sample's last chunk has published
            w0 = ld_word(w);
            w1 = ld_word(w + 1);
            if ((w0 >> 32) != 0ull && (w1 >> 32) != 0ull) break;
            if (clock64() - t0 > 4000000000ll) __trap();   // a protocol bug becomes a CUDA error, not a hung GPU
          }
          s_norm[0] = __uint_as_float((unsigned)w0);
          s_norm[1] = __uint_as_float((unsigned)w1);
        }
        __syncthreads();
        denom = s_norm[0];
        proj = s_norm[1];
      }
      if (vec) {
        // one thread step = VE elements = 16 B of the narrower-typed operand; kH independent steps in flight per thread
        bool first = RMCL_PGD_PRELOAD != 0;
        for (long long i = threadIdx.x; i < nv; i += kH * kPgdThreads) {
          if (!first) {
#pragma unroll
            for (int h = 0; h < kH; ++h) {
              if (i + h * kPgdThreads >= nv) break;
              const long long base = (i + h * kPgdThreads) * VE;
#pragma unroll
              for (int q = 0; q < VE / VG; ++q) pg[h][q] = ld_u4_hint(reinterpret_cast<const uint4*>(g + base) + q, stream);  // last use of g
#pragma unroll
              for (int q = 0; q < VE / VD; ++q) pd[h][q] = ld_u4_hint(reinterpret_cast<const uint4*>(d + base) + q, stream);
            }
          }
          first = false;
#pragma unroll
          for (int h = 0; h < kH; ++h) {
            if (i + h * kPgdThreads >= nv) break;
            const long long base = (i + h * kPgdThreads) * VE;
            float gx[VE], dx[VE];
#pragma unroll
            for (int q = 0; q < VE / VG; ++q) {
              const TG* e = reinterpret_cast<const TG*>(&pg[h][q]);
#pragma unroll
              for (int j = 0; j < VG; ++j) gx[q * VG + j] = to_f32(e[j]);
            }
#pragma unroll
            for (int q = 0; q < VE / VD; ++q) {
              const TD* e = reinterpret_cast<const TD*>(&pd[h][q]);
#pragma unroll
              for (int j = 0; j < VD; ++j) dx[q * VD + j] = to_f32(e[j]);
            }
#pragma unroll
            for (int q = 0; q < VE / VD; ++q) {
              uint4 u;
              TD* e = reinterpret_cast<TD*>(&u);
#pragma unroll
              for (int j = 0; j < VD; ++j) {
                float v = pgd_apply<TD>(dx[q * VD + j], gx[q * VD + j], p.lr, denom, p.mode);
                if (do_clamp) v = clamp_eps<TD>(v, p.eps);
                if (l2proj && proj < 1.f) v = __fmul_rn(v, proj);
                e[j] = from_f32<TD>(v);
              }
#if RMCL_PGD_EXPERIMENT & 4     // timing experiment: the update is not stored
              if (u.x == 0x12345678u && u.y == 0x9abcdef0u) st_u4_hint(reinterpret_cast<uint4*>(d + base) + q, u, stream);
#else
              st_u4_hint(reinterpret_cast<uint4*>(d + base) + q, u, stream);
#endif
            }
          }
        }
      } else {
        for (long long i = threadIdx.x; i < n; i += kPgdThreads) {
          float v = pgd_apply<TD>(to_f32(d[i]), to_f32(g[i]), p.lr, denom, p.mode);
          if (do_clamp) v = clamp_eps<TD>(v, p.eps);
          if (l2proj && proj < 1.f) v = __fmul_rn(v, proj);
          d[i] = from_f32<TD>(v);
        }
      }
    }
  }
}


static int pgd_make_plan(int B, long long N, float lr, float eps, int mode, size_t gsize, size_t dsize, PgdPlan* p) {
  p->N = N;
  p->B = B;
  p->mode = mode;
  p->lr = lr;
  p->eps = eps;
  p->phases = (mode == RMCL_PGD_SIGN_LINF) ? 1 : 2;
  p->n_sums = (mode == RMCL_PGD_L2 && eps > 0.f) ? 3 : 1;
  p->chunk_elems = kPgdChunkBytes / (int)gsize;
  const long long chunks = (N + p->chunk_elems - 1) / p->chunk_elems;
  if (chunks > (1ll << 30)) {
    set_error("rmcl_pgd_step: N=%lld is too large", N);
    return RMCL_E_UNSUPPORTED_DIM;
  }
  p->chunks = (int)chunks;
  // bytes per sample that must survive in L2 between the two phases
  const long long resident = N * (long long)(gsize + (p->n_sums == 3 ? dsize : 0));
  long long batch = (p->n_sums == 3 ? kPgdBatchBytesL2 : kPgdBatchBytes) / resident;
  if (batch < 1) batch = 1;
  if (batch > B) batch = B;
  p->batch = (int)batch;
  p->n_batches = (B + p->batch - 1) / p->batch;
  p->total_items = (long long)p->phases * B * p->chunks;
  return RMCL_OK;
}

static size_t pgd_ctrl_bytes(const PgdPlan& p) { return ((size_t)(2 + p.B) * sizeof(unsigned int) + 255) / 256 * 256; }
// [ticket, exits, done[B] | words[B][2] | part[3][B][chunks]]
static size_t pgd_words_bytes(const PgdPlan& p) { return ((size_t)p.B * 16 + 255) / 256 * 256; }
static size_t pgd_ws_bytes(const PgdPlan& p) {
  return pgd_ctrl_bytes(p) + pgd_words_bytes(p) + 3 * (size_t)p.B * p.chunks * sizeof(float);
}

template <typename TD, typename TG>
static int launch_pgd(void* delta, const void* grad, PgdPlan p, void* ws, size_t ws_bytes, cudaStream_t s) {
  const int sms = sm_count();
  if (sms <= 0) return RMCL_E_CUDA;
  const size_t need = pgd_ws_bytes(p);
  if (!ws || ws_bytes < need) {
    set_error("rmcl_pgd_step: workspace %zu < required %zu (rmcl_pgd_workspace_bytes)", ws_bytes, need);
    return RMCL_E_WORKSPACE;
  }
  unsigned int* c = reinterpret_cast<unsigned int*>(ws);
  p.ticket = c;
  p.exits = c + 1;
  p.done = c + 2;
  p.words = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(ws) + pgd_ctrl_bytes(p));
  p.part = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + pgd_ctrl_bytes(p) + pgd_words_bytes(p));
  long long grid = (long long)sms * 8;
  if (grid > p.total_items) grid = p.total_items;
  pgd_ticket_kernel<TD, TG><<<(unsigned)grid, kPgdThreads, 0, s>>>((TD*)delta, (const TG*)grad, p);
  RMCL_LAUNCH_OK("pgd_ticket_kernel");
  return RMCL_OK;
}

}  // namespace rmcl

extern "C" size_t rmcl_pgd_workspace_bytes(int B, int64_t N, rmcl_dtype grad_dtype) {
  if (B <= 0 || N <= 0 || !rmcl::dtype_ok(grad_dtype)) return 0;
  rmcl::PgdPlan p;
  // the layout does not depend on the mode (three partial planes are always reserved)
  if (rmcl::pgd_make_plan(B, N, 0.f, 1.f, RMCL_PGD_L2, rmcl::dtype_size(grad_dtype), 4, &p) != RMCL_OK) return 0;
  return rmcl::pgd_ws_bytes(p);
}

extern "C" int rmcl_pgd_step(void* delta, rmcl_dtype delta_dtype, const void* grad, rmcl_dtype grad_dtype, int B,
                             int64_t N, float lr, float eps, int mode, void* workspace, size_t workspace_bytes,
                             void* stream) {
  RMCL_CHECK_ARG(delta && grad, "rmcl_pgd_step: null pointer");
  RMCL_CHECK_ARG(B > 0 && N > 0 && B <= (1 << 24), "rmcl_pgd_step: bad sizes B=%d N=%lld", B, (long long)N);
  RMCL_CHECK_ARG(mode >= RMCL_PGD_REF_LINF && mode <= RMCL_PGD_L2, "rmcl_pgd_step: bad mode %d", mode);
  RMCL_CHECK_ARG(rmcl::dtype_ok(delta_dtype) && rmcl::dtype_ok(grad_dtype), "rmcl_pgd_step: bad dtype");
  RMCL_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 255u) == 0, "rmcl_pgd_step: workspace must be 256B aligned");
  cudaStream_t s = (cudaStream_t)stream;
  rmcl::PgdPlan p;
  const int rc = rmcl::pgd_make_plan(B, N, lr, eps, mode, rmcl::dtype_size(grad_dtype), rmcl::dtype_size(delta_dtype), &p);
  if (rc != RMCL_OK) return rc;
  using bf16 = __nv_bfloat16;
  if (delta_dtype == RMCL_F32 && grad_dtype == RMCL_F32)
    return rmcl::launch_pgd<float, float>(delta, grad, p, workspace, workspace_bytes, s);
  if (delta_dtype == RMCL_F32 && grad_dtype == RMCL_BF16)
    return rmcl::launch_pgd<float, bf16>(delta, grad, p, workspace, workspace_bytes, s);
  if (delta_dtype == RMCL_BF16 && grad_dtype == RMCL_F32)
    return rmcl::launch_pgd<bf16, float>(delta, grad, p, workspace, workspace_bytes, s);
  return rmcl::launch_pgd<bf16, bf16>(delta, grad, p, workspace, workspace_bytes, s);
}
