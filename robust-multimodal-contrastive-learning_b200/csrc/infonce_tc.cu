// K1 partial stage, tcgen05 / TMEM / TMA version (bf16 queue, C in {64,128,256}).
//
// One CTA = 128 query rows x one contiguous range of queue columns, walked in tiles of TN columns
// ("flash" online softmax over the queue axis; output format in infonce.cuh).  Per tile
//     S  = Q^ . tile            tcgen05.mma  M128 x N=TN x K=C   A = Q^ (TMEM), B = tile (smem, MN-major)
//     P  = 2^(S*scale - m)      128 softmax threads, one row each: tcgen05.ld -> ex2 -> bf16 -> st.shared (SW128)
//     O += P . tile^T           tcgen05.mma  M128 x N=C  x K=TN  A = P  (smem, K-major), B = tile (smem, K-major)
// The queue tile is staged ONCE by TMA (SWIZZLE_128B boxes of C rows x 64 columns) and serves as
// the B operand of both GEMMs: the reference layout [C,K] with K contiguous (vilt_module.py:92)
// is MN-major for the first and K-major for the second, so no transpose is ever materialised.
//
// Tensor memory (512 columns x 128 lanes):   [ Q^ : C/2 | O : C | S0 : TN | S1 : TN ]
//   Q^ is bf16 packed two per column and is the A operand of every S GEMM, so shared memory holds
//   only the TMA ring of queue tiles and a double-buffered P tile.  P deliberately does NOT alias
//   its S buffer in tensor memory: tcgen05.mma instructions with different accumulators/shapes are
//   not ordered against each other, and the S GEMM of tile i+2 was observed overwriting P(i) before
//   the O GEMM of tile i had read it.  For the same reason no hazard is ever covered by a commit
//   issued behind a *different* MMA group: every consumer waits on the commit placed directly
//   behind the group that produces (or last reads) what it needs.
// Warp roles (320 threads): warps 0-7 softmax + epilogue (TMEM lane quadrant = warp % 4; warp w
//   owns the even tiles, warp w+4 the odd tiles of the same 32 rows), warp 8 TMA producer + TMEM
//   allocator, warp 9 MMA issuer (one elected lane each).
// Online softmax uses a lazily updated reference maximum: O and l are only rescaled when a row's
// maximum grows by more than 2^40 (everything is floating point, so a stale reference costs no
// precision until it threatens overflow), which keeps the correction off the critical path while
// the final (m, l, O) triple stays exact up to fp32 rounding.
//
// FUSED instantiations are the whole InfoNCE call in ONE launch (cooperative: every CTA is resident, grid <= SM count):
//   phase A  the CTAs share out the B rows of the prep stage (normalise q and k, positive logit, bf16 Q^ copies) while
//            their TMA warps already fill the queue ring;  grid barrier
//   phase B  the flash pass above (Q^ into tensor memory, tile loop, split partials)            ;  grid barrier
//   phase C  the CTAs share out the B rows of the finalize stage (merge the splits, loss, dq, dk), the last row's team
//            reduces the loss.
// Against the three-launch chain (prep -> partial -> finalize under programmatic dependent launch) this removes two
// launch latencies and the drain/fill between the kernels; the row code is the same (infonce_rows.cuh).
#include <stdlib.h>

#include "infonce.cuh"
#include "infonce_rows.cuh"
#include "tc_ptx.cuh"

namespace rmcl {

bool infonce_tc_built() { return true; }

namespace {

using namespace tcx;

constexpr int kSoftmaxWarps = 8;   // warps w and w+4 share TMEM lane quadrant w and alternate tiles
constexpr int kTmaWarp = 8, kMmaWarp = 9;
constexpr int kTcThreads = 320;
constexpr int kTcRows = 128;
constexpr int kSmemBudget = 208 * 1024;   // P double buffer + TMA ring; leaves room for alignment slack + static smem
constexpr float kRescaleThreshold = 40.f;  // log2 units: P <= 2^40, far inside bf16/fp32 range
// The first tile's row maximum plus this margin is the initial reference: a later tile must exceed the
// first one by 2^(margin+threshold) before anything is rescaled, which for logit distributions without
// such jumps means never (with the raw first-tile maximum as reference, ~20 % of the CTAs of the cfg2
// bench took the ~2.5 k-cycle O read-modify-write once or twice).  P then starts around 2^-24; entries
// more than ~2^-100 below the first maximum flush to zero, i.e. below e^-70 of the row sum.
#ifndef RMCL_TC_INIT_MARGIN
#define RMCL_TC_INIT_MARGIN 24.f
#endif
constexpr float kInitMargin = RMCL_TC_INIT_MARGIN;

struct TcShared {
  uint64_t k_full[8];
  uint64_t k_empty[8];
  uint64_t s_full[2];   // S GEMM of a tile has landed in S[b]
  uint64_t s_free[2];   // S[b] is in registers: the S GEMM two tiles ahead may overwrite it
  uint64_t p_full[2];   // P[b] is in shared memory: the O GEMM of the tile may start
  uint64_t o_done[2];   // o_done[i&1]: O GEMM of tile i has completed (committed directly behind it)
  uint64_t q_full;
  uint64_t dry;         // target of the dry-pass commits; nobody waits on it
  uint32_t tmem_base;
  float m_ref[kTcRows];        // reference maximum (log2 units) of the O accumulator row
  float xm[kTcRows];           // end of kernel: odd-tile warps hand (m, l, argmax) to the even-tile warps
  float xl[kTcRows];
  float xav[kTcRows];
  int xai[kTcRows];
  float xd[kTcRows];           // diagnostics: the odd-tile warp's distance sum
  FinShared fin[2];            // FUSED: scratch of the two finalize teams
};


// Measurement aid (rmcl_debug_tc_timeline): CTA (0,0) stamps clock64() at protocol points into a
// caller-provided buffer; `timeline` is null in normal operation (one uniform predicate per stamp).
//   [0] kernel entry   [1] setup done   [2] PDL wait done   [3] Q^ in TMEM   [4] all O GEMMs done
//   [5] statistics written   [6] partials stored
//   [8 + 6*i + k], tile i < kTimelineTiles:  k=0 S issued   1 S landed (softmax)   2 P stored (softmax)
//                                           3 O issued   4 P buffer free (o_done of tile i-2 seen)   5 S free seen by issuer
//                                           6 S in registers   7 decision taken (exponentials start)
//   [kTimelineHead + 2*cta + {0,1}], cta < kTimelineCtas: %globaltimer (ns) at entry / exit of every CTA
constexpr int kTimelineTiles = 40;
constexpr int kTimelineHead = 8 + 8 * kTimelineTiles;
constexpr int kTimelineCtas = 1024;
constexpr int kTimelineWords = kTimelineHead + 2 * kTimelineCtas;
#if defined(RMCL_TC_TIMELINE) && RMCL_TC_TIMELINE
__device__ __forceinline__ long long global_ns() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#endif
// Compiled out of the product build (RMCL_TC_TIMELINE=0): even never-taken, the per-lane predicate in
// front of the elect-guarded issue blocks cost ~2 us per launch.  tools/tc_timeline.py builds its own
// instrumented copy of the library.
#ifndef RMCL_TC_TIMELINE
#define RMCL_TC_TIMELINE 0
#endif
// Timing experiments (results are garbage): bit 0 no tcgen05.ld of S, 1 no exponentials, 2 no P stores, 3 no decision
// handshake, 4 queue tiles fetched only for the first ring round, 5 no O write-out, 6 no Q^ fetch — which part of the kernel
// the time goes to (profiles/r2_tc_experiments.txt)
#ifndef RMCL_TC_EXPERIMENT
#define RMCL_TC_EXPERIMENT 0
#endif
__device__ __forceinline__ void tl_stamp(long long* tl, int idx) {
#if RMCL_TC_TIMELINE
  if (tl != nullptr && idx < kTimelineHead) tl[idx] = clock64();
#endif
}

// ----------------------------------------------------------------------------------- kernel
template <int C, int TN, bool DIAG, bool WANT_O, bool FUSED>
__global__ void __launch_bounds__(kTcThreads, 1)
    infonce_tc_kernel(const __grid_constant__ CUtensorMap tmap_queue, const __nv_bfloat16* __restrict__ q_hat, int B,
                      long long K, float scale2, long long cols_per_split, int want_argmax, float* __restrict__ pm,
                      float* __restrict__ pl, float* __restrict__ pav, int* __restrict__ pai, __nv_bfloat16* __restrict__ po,
                      long long* __restrict__ timeline, const float* __restrict__ n2, const float* __restrict__ qn2,
                      float* __restrict__ pdist, const FusedArgs fz, const __nv_bfloat16* __restrict__ queue_raw, long long ldq) {
  constexpr bool want_o = WANT_O;   // compile-time: the with-gradient instantiation is the tuned kernel, untouched
  // want_o == false (no gradient requested: the clean-query argmax call, the greedy attack's candidate losses): the statistics
  // pass only — no P tile, no O GEMM, no partial write-out; ring stages are released by the S GEMM's own commit.
  constexpr int kStageBytes = C * TN * 2;
  constexpr int kBoxBytes = C * 128;            // one TMA box: C rows x 64 bf16 columns
  constexpr int kBoxes = TN / 64;
  constexpr int kPBytes = kBoxes * 16384;       // one P tile: 128 rows x TN bf16 as 128-byte-row boxes
  constexpr int kStages = ((kSmemBudget - 2 * kPBytes) / kStageBytes) < 8 ? ((kSmemBudget - 2 * kPBytes) / kStageBytes) : 8;
  constexpr uint32_t kTmQ = 0, kTmO = C / 2, kTmS = C / 2 + C;
  constexpr int HC = C / 2;                     // O columns per thread in the epilogue
  static_assert(C / 2 + C + 2 * TN <= 512, "tensor memory budget");
  static_assert(TN % 64 == 0 && C % 64 == 0 && C <= 256, "tile shape");
  static_assert(kStages >= 4, "three tiles are live (S GEMM runs two ahead of the O GEMM) plus one in flight");
  static_assert(kTcRows * (2 * C + 16) <= kStages * kStageBytes, "epilogue staging must fit the ring");
  static_assert(2 * kPBytes >= kSoftmaxWarps * 4096, "Q^ transpose scratch lives in the P buffers");
  constexpr uint32_t kIdescS = make_idesc(128, TN, 1);
  constexpr uint32_t kIdescO = make_idesc(128, C, 0);

  extern __shared__ uint8_t smem_raw[];
  __shared__ TcShared sh;
  uint8_t* pbuf = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* ring = pbuf + 2 * kPBytes;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int split = blockIdx.x;
  const int row0 = blockIdx.y * kTcRows;
  const long long k_begin = (long long)split * cols_per_split;
  const long long k_end = (k_begin + cols_per_split < K) ? k_begin + cols_per_split : K;
  const int n_tiles = (int)((k_end - k_begin + TN - 1) / TN);
  long long* tl = (blockIdx.x == 0 && blockIdx.y == 0 && lane == 0) ? timeline : nullptr;   // lane 0 of every warp of CTA (0,0)
  if (tid == 0) tl_stamp(tl, 0);
  const int cta_linear = blockIdx.y * gridDim.x + blockIdx.x;
  const int n_ctas = gridDim.x * gridDim.y;
#if RMCL_TC_TIMELINE
  if (timeline != nullptr && tid == 0 && cta_linear < kTimelineCtas) timeline[kTimelineHead + 2 * cta_linear] = global_ns();
#endif

  if (tid == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&sh.k_full[i], 1);
      mbar_init(&sh.k_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sh.s_full[i], 1);
      mbar_init(&sh.s_free[i], kTcRows);
      mbar_init(&sh.p_full[i], kTcRows);
      mbar_init(&sh.o_done[i], 1);
    }
    mbar_init(&sh.q_full, kSoftmaxWarps * 32);
    mbar_init(&sh.dry, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_queue) : "memory");
  }
  if (warp == kTmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh.tmem_base)),
                 "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sh.tmem_base;
  if (tid == 0) tl_stamp(tl, 1);

  if (warp < kSoftmaxWarps) {
    // =============================================================== softmax + epilogue warps
    // Two warps per row quadrant (a warp may only touch TMEM lanes 32*(w%4)..+31): warp w owns the
    // even tiles of rows 32w..32w+31, warp w+4 the odd tiles, so the two work half a period out of
    // phase — one is in its MUFU-bound exponential while the other waits on barriers / moves data —
    // and each has two tile periods to finish one tile.  They share the O accumulator, hence one
    // reference maximum per row (sh.m_ref); rescale decisions are taken strictly in tile order
    // through a named-barrier handshake between the two warps (ids 1..8).
    const int quad = warp & 3, par = warp >> 2;
    const int r = quad * 32 + lane;                       // row inside the block == TMEM lane
    const uint32_t tlane = tmem + ((uint32_t)(quad * 32) << 16);
    const uint32_t dec_mine = 1 + quad * 2 + par;         // "decision of one of my tiles is published"
    const uint32_t dec_other = 1 + quad * 2 + (par ^ 1);

    if (FUSED) {
      // ---- phase A: this CTA's share of the prep rows, one WARP per row (a CTA owns at most a handful of rows: what counts is
      //      the latency of one row, not throughput); padding rows of the bf16 operand included
      if (cta_linear == 0 && tid == 0) {
        fz.fin.counter[0] = 0u;   // rows finalized
        fz.fin.counter[1] = 0u;   // overflow flag (two-pass kernels only; kept clean)
      }
      for (int row = cta_linear + warp * n_ctas; row < fz.prep.b_pad; row += kSoftmaxWarps * n_ctas)
        prep_row_warp_rt(fz.prep, fz.q_bf16, fz.k_bf16, row, lane);
      if (warp == 0) tl_stamp(tl, 320);   // this CTA's prep rows are written (warp 0's row)
    }

    // m_mine: the reference maximum this warp's row sum l_run is expressed in
    float m_mine = -INFINITY, l_run = 0.f, av_raw = -INFINITY;
    int ai = 0;
    float dsum = 0.f, qn2_r = 0.f;                        // diagnostics: sum_j |q^_r - queue_j|, |q^_r|^2
    // Pass it = -1 is a DRY run of the tile body on whatever tensor memory holds, with every side effect
    // (barriers, shared-memory and m_ref writes) switched off.  This kernel is launched early (PDL) and
    // would otherwise idle in griddepcontrol.wait for the ~3 us the prep kernel takes; the first execution
    // of the ~10 KB unrolled softmax body costs ~5 k cycles in instruction-cache misses (measured with the
    // timeline: 5.0 k for tile 0 against 1.7 k for every later tile), so it is paid during that wait
    // instead of on the critical path of the first tile.
    // ---- Q^ rows: the global loads are ISSUED here and consumed at it == 0, so that the optional dry pass of the tile body
    //      (it == -1) runs while they are in flight instead of in front of them
    if (FUSED) grid_barrier(fz.bar_words, (unsigned)n_ctas, 15, kSoftmaxWarps * 32);   // every CTA's prep rows are written
    else pdl_wait();   // launched early (PDL): q_hat is written by the prep kernel that may still be running
    if (tid == 0) tl_stamp(tl, 2);
    constexpr int kChunks = C / 64;                     // 64-column chunks of a row
    constexpr int kMine = (kChunks + 1) / 2;            // chunks handled by this warp: ch = 2*t + par
    uint4 v[kMine][8];
    {
      const __nv_bfloat16* qw = q_hat + (size_t)(split % kQhatReplicas) * ((size_t)gridDim.y * kTcRows * C) +
                                (size_t)(row0 + quad * 32) * C;
#pragma unroll
      for (int t = 0; t < kMine; ++t) {
        const int ch = 2 * t + par;
        if (ch < kChunks) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            v[t][j] = (RMCL_TC_EXPERIMENT & 64) ? make_uint4(0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u) :   // experiment bit 6: no Q^ fetch
                      FUSED ? __ldcg(reinterpret_cast<const uint4*>(qw + (size_t)(4 * j + (lane >> 3)) * C + ch * 64) + (lane & 7))
                            : __ldg(reinterpret_cast<const uint4*>(qw + (size_t)(4 * j + (lane >> 3)) * C + ch * 64) + (lane & 7));
        }
      }
    }
    // Measured again in round 2 with the loads issued in front of the dry pass (profiles/r2_ab_dry.txt): whole call 32.8 us
    // without, 33.6 us with the dry softmax pass, 33.7 us with both dry passes — over consecutive calls the instruction cache
    // is warm anyway, so the dry passes stay compiled out.
#ifndef RMCL_TC_DRY_SOFTMAX
#define RMCL_TC_DRY_SOFTMAX 0
#endif
    int it_first = RMCL_TC_DRY_SOFTMAX ? -1 : 0;
    asm volatile("" : "+r"(it_first));   // opaque: keeps the compiler from peeling the dry pass off and deleting it
    for (int it = it_first;; ++it) {
      const bool live = it >= 0;
      const int i = live ? par + 2 * it : par;
      if (it == 0) {
        // ---- Q^ rows -> tensor memory (bf16 pairs, element 2j in the low half of column j).
        // Coalesced global loads (8 lanes cover one 128-byte row segment), all issued before the first
        // use, transposed through this warp's 4 KB of the (still unused) P buffers with the same
        // 16-byte XOR swizzle that keeps the segment writes and the row-per-lane reads conflict free.
        {
          uint8_t* scratch = pbuf + warp * 4096;
    #pragma unroll
          for (int t = 0; t < kMine; ++t) {
            const int ch = 2 * t + par;
            if (ch < kChunks) {
    #pragma unroll
              for (int j = 0; j < 8; ++j) {
                const int rr = 4 * j + (lane >> 3), pc = lane & 7;
                *reinterpret_cast<uint4*>(scratch + rr * 128 + ((pc ^ (rr & 7)) << 4)) = v[t][j];
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
          }
          tc_wait_st();
          tc_fence_before();
          mbar_arrive(&sh.q_full);
          if (tid == 0) tl_stamp(tl, 3);
        }

        m_mine = -INFINITY; l_run = 0.f; av_raw = -INFINITY; ai = 0;   // discard the dry pass
        dsum = 0.f;
        if (DIAG) qn2_r = (row0 + r < B) ? __ldcg(qn2 + row0 + r) : 0.f;   // after the wait: written by the prep stage
      }
      if (live && i >= n_tiles) break;
      const int b = par;                                  // == i & 1
      const uint32_t ph = (i >> 1) & 1;
      const uint32_t ts = tlane + kTmS + b * TN;
      if (live) {
        mbar_wait(&sh.s_full[b], ph);
        tc_fence_after();
        if (quad == 0) tl_stamp(tl, 8 + 8 * i + 1);
      }
      uint32_t sv[TN];
#if RMCL_TC_EXPERIMENT & 1   // timing experiment: no tcgen05.ld of S
#pragma unroll
      for (int j = 0; j < TN; ++j) sv[j] = __float_as_uint(0.001f * (float)(j + lane));
#else
#pragma unroll
      for (int ch = 0; ch < TN / 32; ++ch) tc_ld32(ts + ch * 32, sv + ch * 32);
      tc_wait_ld();
#endif
      tc_fence_before();
      if (live) {
        mbar_arrive(&sh.s_free[b]);                       // S[b] is in registers
        if (quad == 0) tl_stamp(tl, 8 + 8 * i + 6);
      }
      const long long col0 = k_begin + (long long)i * TN;
      if (DIAG && live) {
        // mean L2 distance to the negatives (objectives.py:343): |q^ - queue_j|^2 = |q^|^2 - 2 S_j + |queue_j|^2.
        // The column norms are the same for every lane: uniform 16-byte loads, one transaction each.
        const float4* nv = reinterpret_cast<const float4*>(n2 + col0);
#pragma unroll
        for (int c4 = 0; c4 < TN / 4; ++c4) {
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
      if (live && col0 + TN > k_end) {  // ragged last tile: TMA zero-filled the columns past K
        const int valid = (int)(k_end - col0);
#pragma unroll
        for (int j = 0; j < TN; ++j)
          if (j >= valid) sv[j] = 0xff800000u;  // -inf
      }
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < TN; ++j) mx = fmaxf(mx, __uint_as_float(sv[j]));
      if (want_argmax && mx > av_raw) {  // strict: the earliest tile keeps ties
        av_raw = mx;
        int idx = 0;
#pragma unroll
        for (int j = TN - 1; j >= 0; --j)
          if (__uint_as_float(sv[j]) == mx) idx = j;
        ai = (int)col0 + idx;
      }
      const float m_tile = mx * scale2;

      // ---- decision point of tile i: everything up to tile i-1 has been decided by the other warp
#if !(RMCL_TC_EXPERIMENT & 8)   // timing experiment: no decision handshake between the two warps of a quadrant
      if (live && i > 0) named_bar_sync(dec_other, 64);
#endif
      float m_now = (i > 0) ? sh.m_ref[r] : m_tile + kInitMargin;
      if (m_now != m_mine) {            // the other warp moved the reference (or this is my first tile)
        l_run *= exp2f(m_mine - m_now); // first tile: l_run == 0
        m_mine = m_now;
      }
      const bool grow = live && (i > 0) && (m_tile > m_now + kRescaleThreshold);
      if (__any_sync(0xffffffffu, grow)) {
        // rare: bring this quadrant's O rows to the new reference maximum.  Every O GEMM up to
        // tile i-1 must have landed (same accumulator => in order, so the latest commit suffices);
        // the one of tile i cannot start before this warp arrives on p_full below.
        const float alpha = grow ? exp2f(m_now - m_tile) : 1.f;
        if (grow) { m_mine = m_tile; l_run *= alpha; }
        if (want_o) {
          mbar_wait(&sh.o_done[(i - 1) & 1], ((i - 1) >> 1) & 1);
          tc_fence_after();
        }
#pragma unroll 1
        for (int ch = 0; want_o && ch < C / 32; ++ch) {
          uint32_t o[32];
          tc_ld32(tlane + kTmO + ch * 32, o);
          tc_wait_ld();
#pragma unroll
          for (int j = 0; j < 32; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * alpha);
          tc_st32(tlane + kTmO + ch * 32, o);
        }
        tc_wait_st();
      }
      if (live && (i == 0 || grow)) sh.m_ref[r] = m_mine;
#if !(RMCL_TC_EXPERIMENT & 8)
      if (live && i + 1 < n_tiles) {    // publish the decision to the warp that owns tile i+1
        __threadfence_block();
        named_bar_arrive(dec_mine, 64);
      }
#endif

      if (live && quad == 0) tl_stamp(tl, 8 + 8 * i + 7);
      // ---- P = 2^(S*scale - m) as bf16 pairs, row sum in fp32
      const float neg_m = -m_mine;
      float ls[4] = {0.f, 0.f, 0.f, 0.f};   // independent partial sums: no long dependent add chain
      uint32_t pw[TN / 2];
#pragma unroll
      for (int j = 0; j < TN / 2; ++j) {
#if RMCL_TC_EXPERIMENT & 2   // timing experiment: no exponentials
        const float p0 = fmaf(__uint_as_float(sv[2 * j]), scale2, neg_m);
        const float p1 = fmaf(__uint_as_float(sv[2 * j + 1]), scale2, neg_m);
#else
        const float p0 = ex2_ftz(fmaf(__uint_as_float(sv[2 * j]), scale2, neg_m));
        const float p1 = ex2_ftz(fmaf(__uint_as_float(sv[2 * j + 1]), scale2, neg_m));
#endif
        ls[j & 3] += p0 + p1;
        const __nv_bfloat162 pk = __floats2bfloat162_rn(p0, p1);
        pw[j] = *reinterpret_cast<const uint32_t*>(&pk);
      }
      l_run += (ls[0] + ls[1]) + (ls[2] + ls[3]);
      // P buffer b was last read by the O GEMM of tile i-2.  That GEMM's own commit is the only
      // thing that may be trusted here: a commit behind a later MMA group with another accumulator
      // does NOT imply it has finished — such groups overlap and complete out of order (observed
      // on B200: fast warps overwrote P while the O GEMM was still reading it).
      if (live && want_o && i >= 2) mbar_wait(&sh.o_done[b], ((i - 2) >> 1) & 1);
      if (live && quad == 0) tl_stamp(tl, 8 + 8 * i + 4);
      // row r of P in the K-major SWIZZLE_128B layout: 16-byte chunk c of a 128-byte row lands at c ^ (r & 7)
      // (dry pass: the P buffers double as other warps' Q^ transpose scratch, so nothing may be stored;
      //  l_run keeps the exponentials alive for the compiler)
      if (live && want_o) {
        uint8_t* prow = pbuf + b * kPBytes + r * 128;
#if RMCL_TC_EXPERIMENT & 4   // timing experiment: P is not stored (one word keeps pw alive)
        if (l_run == 123.456f) *reinterpret_cast<uint32_t*>(prow) = pw[0] ^ pw[TN / 2 - 1];
#else
#pragma unroll
        for (int c = 0; c < TN / 8; ++c) {
          uint8_t* dst = prow + (c >> 3) * 16384 + (((c & 7) ^ (r & 7)) << 4);
          *reinterpret_cast<uint4*>(dst) = make_uint4(pw[4 * c], pw[4 * c + 1], pw[4 * c + 2], pw[4 * c + 3]);
        }
#endif
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> tensor core reads
        mbar_arrive(&sh.p_full[b]);
        if (quad == 0) tl_stamp(tl, 8 + 8 * i + 2);
      }
    }

    // ---- epilogue: merge the two warps' statistics (odd-tile warp -> even-tile warp), then O
    if (tid == 0) pdl_trigger();   // the finalize kernel's CTAs may become resident now
    if (want_o) {
      if (n_tiles >= 2) mbar_wait(&sh.o_done[(n_tiles - 2) & 1], ((n_tiles - 2) >> 1) & 1);
      mbar_wait(&sh.o_done[(n_tiles - 1) & 1], ((n_tiles - 1) >> 1) & 1);
      tc_fence_after();
    }
    if (tid == 0) tl_stamp(tl, 4);
    const bool row_ok = (row0 + r) < B;
    if (par == 1) {
      sh.xm[r] = m_mine;
      sh.xl[r] = l_run;
      sh.xav[r] = av_raw;
      sh.xai[r] = ai;
      if (DIAG) sh.xd[r] = dsum;
    }
    named_bar_sync(9 + quad, 64);
    if (par == 0 && row_ok) {
      const float m_fin = sh.m_ref[r];
      // exp2f(-inf) == 0 covers a warp that never owned a tile (l == 0, m == -inf)
      const float l_tot = l_run * exp2f(m_mine - m_fin) + sh.xl[r] * exp2f(sh.xm[r] - m_fin);
      const float av1 = sh.xav[r];
      const int ai1 = sh.xai[r];
      if (av1 > av_raw || (av1 == av_raw && ai1 < ai)) { av_raw = av1; ai = ai1; }
      const size_t o = (size_t)(row0 + r) * gridDim.x + split;
      pm[o] = m_fin;
      pl[o] = l_tot;
      pav[o] = av_raw * scale2;
      pai[o] = ai;
      if (DIAG) pdist[o] = dsum + sh.xd[r];
    }
    if (tid == 0) tl_stamp(tl, 5);
    // O row r, half of the columns per warp: tensor memory -> bf16 -> this thread's own staging segment
    // (16-byte chunks; row stride 2C+16 bytes keeps the 128-bit stores bank-conflict free) -> one bulk
    // async copy to global.  No cross-thread synchronisation: every thread stores what it staged.
    // The partial accumulators leave the kernel in bf16: their write-out (148 CTAs x 128 rows x C)
    // is L2-write bound and sits on the critical path of every CTA, and the O GEMM's A operand (P)
    // was bf16 already, so this adds one more 2^-9 relative rounding per split to dq.
    uint8_t* stage = ring + (size_t)r * (2 * C + 16) + par * HC * 2;   // every TMA write has been consumed
#pragma unroll 1
    for (int ch = 0; want_o && !(RMCL_TC_EXPERIMENT & 32) && ch < HC / 32; ++ch) {   // experiment bit 5: no O write-out
      uint32_t o[32];
      tc_ld32(tlane + kTmO + par * HC + ch * 32, o);
      tc_wait_ld();
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
    if (row_ok && want_o && !(RMCL_TC_EXPERIMENT & 32)) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      bulk_store_row(po + ((size_t)split * B + row0 + r) * C + par * HC, stage, HC * 2);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      if (FUSED) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // the WRITES must have landed: other CTAs read them in phase C
      else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    if (FUSED) __threadfence();   // statistics and partials of this CTA are visible before it arrives at the grid barrier
    if (tid == 0) tl_stamp(tl, 6);
    tc_fence_before();
  } else if (warp == kTmaWarp) {
    // ========================================================================= TMA producer
    // RMCL_TC_L2_PREFETCH: a tile is C rows x 128 bytes, each row in a different DRAM page (row stride = 2K bytes), i.e. the
    // TMA boxes reach DRAM as 128-byte accesses scattered over C pages.  This CTA's whole share of a queue row is one
    // contiguous segment of cols_per_split x 2 bytes (1.75 KB at cfg2): ask the L2 for it with one bulk prefetch per row before
    // the tile loop, so DRAM sees C long bursts instead of C x n_tiles short ones and the boxes then hit in L2.
    // Measured (profiles/r2_tc_experiments.txt): the flash pass gets SLOWER, 22.0 -> 25.1 us at cfg2 — fetching the tiles is not
    // what limits the kernel (skipping the fetches altogether gains 0.4 us), and 148 x 256 prefetches land in front of the Q^
    // fetch and the first tiles.  Off.
#ifndef RMCL_TC_L2_PREFETCH
#define RMCL_TC_L2_PREFETCH 0
#endif
    if (RMCL_TC_L2_PREFETCH && n_tiles > 1 && queue_raw != nullptr) {
      const uint32_t bytes = (uint32_t)((k_end - k_begin) * 2);   // K % 8 == 0 and k_begin % TN == 0: multiples of 16
      for (int row = lane; row < C; row += 32)
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(queue_raw + (size_t)row * ldq + k_begin), "r"(bytes) : "memory");
    }
    // RMCL_TC_TILES_BEFORE_Q: the first S GEMM needs Q^ AND tile 0; the ring prefetch (kStages x 32 KB per SM, 24 MB over the
    // chip) is issued at kernel entry and the Q^ fetch (64 KB per SM, out of L2) then queues behind it on the way into the
    // SM — measured as 2.6 us of fill that neither more nor fewer Q^ replicas change (profiles/r2_tc_experiments.txt).  Only
    // this many tiles are requested ahead of Q^; the rest of the ring follows once Q^ is in tensor memory.
#ifndef RMCL_TC_TILES_BEFORE_Q
#define RMCL_TC_TILES_BEFORE_Q 2
#endif
    for (int i = 0; i < n_tiles; ++i) {
      const int st = i % kStages;
      if (i == RMCL_TC_TILES_BEFORE_Q && i < kStages) mbar_wait(&sh.q_full, 0);
      mbar_wait(&sh.k_empty[st], ((i / kStages) & 1) ^ 1);
      if (elect_one()) {
#if RMCL_TC_EXPERIMENT & 16   // timing experiment: only the first kStages tiles are fetched, later ones reuse the stale stage
        if (i >= kStages) {
          mbar_arrive(&sh.k_full[st]);
        } else
#endif
        {
          mbar_expect_tx(&sh.k_full[st], kStageBytes);
          const long long col0 = k_begin + (long long)i * TN;
#pragma unroll
          for (int bx = 0; bx < kBoxes; ++bx)
            tma_load_2d(ring + (size_t)st * kStageBytes + (size_t)bx * kBoxBytes, &tmap_queue, &sh.k_full[st],
                        (int)(col0 + bx * 64), 0);
        }
      }
      __syncwarp();
    }
  } else if (warp == kMmaWarp) {
    // =========================================================================== MMA issuer
    // The whole warp walks the protocol (waits are warp-uniform); one elected lane issues.
    // One loop, one code instance of each GEMM: iteration j issues the S GEMM of tile j and the O GEMM of
    // tile j-2 (S runs two tiles ahead: S[b] is free again as soon as the softmax warps hold tile j-2 in
    // registers, long before P(j-2) is ready).  Iteration j = -1 is a DRY pass, issued while the CTA is
    // still waiting for the prep kernel: both GEMMs run once on whatever shared/tensor memory holds
    // (their accumulators are overwritten by the first real GEMMs, which the hardware orders behind them)
    // and commit to a barrier nobody waits on.  It pulls the ~10 KB of straight-line issue code and the
    // descriptors' first use off the critical path of tile 0 (timeline: first S GEMMs 1.4-2.0 k cycles
    // cold against 0.8 k warm; three inlined copies of the S issue cost three cold starts before).
#ifndef RMCL_TC_DRY_MMA
#define RMCL_TC_DRY_MMA 0
#endif
    int j_first = RMCL_TC_DRY_MMA ? -1 : 0;
    asm volatile("" : "+r"(j_first));
    // RMCL_TC_MERGED_ISSUE: the S GEMM of tile j and the O GEMM of tile j-2 are issued as ONE straight-line burst of 20 MMAs
    // behind a single wait phase.  tools/mma_bench.cu (profiles/r2_mma_bench.txt): the identical MMA stream runs at 1042
    // cycles per tile (98 % of the tensor roof) when issued as one burst per tile, beside the kernel's tcgen05.ld / P-store /
    // bulk-load traffic at 1193 (86 %) — the tile loop here took ~1700: the tensor pipe accepts only a few MMAs ahead of
    // execution, so each of the two wait-then-issue phases per tile (barrier probes, fence, elect) drains it.
    //   1: always merged (the burst waits for P(j-2) before S(j) is issued)
    //   2: merged when P(j-2) is already there (non-blocking probe), else S(j) first as before
#ifndef RMCL_TC_MERGED_ISSUE
#define RMCL_TC_MERGED_ISSUE 2
#endif
    for (int j = j_first; j < n_tiles + 2; ++j) {
      const bool live = j >= 0;
      if (j == 0) {
        mbar_wait(&sh.q_full, 0);
        tc_fence_after();
      }
      const bool do_s = !live || j < n_tiles;
      const bool do_o = want_o && (!live || j >= 2);
      const int is = live ? j : 0, io = live ? j - 2 : 0;
      const int st_s = is % kStages, st_o = io % kStages;
      uint64_t* bar_s = live ? &sh.s_full[is & 1] : &sh.dry;
      uint64_t* bar_k = live ? &sh.k_empty[st_o] : &sh.dry;
      uint64_t* bar_o = live ? &sh.o_done[io & 1] : &sh.dry;
      // -------- S GEMM of tile is: B = tile as [N=TN columns][K=16 rows of C], MN-major: 8-row groups 1024 B apart,
      //          64-column boxes kBoxBytes apart
      auto issue_s = [&]() {
        const uint32_t sbase = smem_u32(ring + (size_t)st_s * kStageBytes);
        const uint32_t dd = tmem + kTmS + (is & 1) * TN;
#pragma unroll
        for (int s = 0; s < C / 16; ++s) {
          const uint64_t bd = make_sw128_desc(sbase + s * 2048, kBoxBytes, 1024);
          tc_mma_ts(dd, tmem + kTmQ + s * 8, bd, kIdescS, s > 0);
        }
        tc_commit(bar_s);
        if (!want_o && live) tc_commit(&sh.k_empty[st_s]);   // no O GEMM will read this stage
      };
      // -------- O GEMM of tile io: A = P as [M=128 rows][K=16 columns], B = tile as [N=C rows][K=16 columns]; both K-major:
      //          rows 128 B apart, 8-row groups 1024 B apart, 16 columns = 32 B inside the swizzled row
      auto issue_o = [&]() {
        const uint32_t sbase = smem_u32(ring + (size_t)st_o * kStageBytes);
        const uint32_t pa = smem_u32(pbuf + (io & 1) * kPBytes);
#pragma unroll
        for (int s = 0; s < TN / 16; ++s) {
          const uint64_t ad = make_sw128_desc(pa + (s >> 2) * 16384 + (s & 3) * 32, 16, 1024);
          const uint64_t bd = make_sw128_desc(sbase + (s >> 2) * kBoxBytes + (s & 3) * 32, 16, 1024);
          tc_mma_ss(tmem + kTmO, ad, bd, kIdescO, (io > 0 || s > 0) ? 1u : 0u);
        }
        tc_commit(bar_k);
        tc_commit(bar_o);
      };
      bool merged = false;
      if (live && do_s && do_o && RMCL_TC_MERGED_ISSUE == 2) {
        // steady state: the three barriers of a tile are probed TOGETHER (three try_waits in flight, ~one probe latency
        // instead of three in a row): tcgen05.mma issue blocks while the pipe executes and the pipe runs dry within a few
        // MMAs once this thread stops feeding it, so whatever the issuer does between two bursts is tensor-pipe idle time
        const long long t0 = clock64();
        for (;;) {
          const uint32_t got = mbar_try_wait3(&sh.s_free[is & 1], ((is - 2) >> 1) & 1, &sh.k_full[st_s], (is / kStages) & 1,
                                              &sh.p_full[io & 1], (io >> 1) & 1);
          if ((got & 3u) == 3u) {
            merged = (got & 4u) != 0u;
            break;
          }
          if (clock64() - t0 > 4000000000ll) __trap();
        }
        tl_stamp(tl, 8 + 8 * is + 5);
      } else {
        if (do_s && live) {
          if (is >= 2) {
            mbar_wait(&sh.s_free[is & 1], ((is - 2) >> 1) & 1);
            tl_stamp(tl, 8 + 8 * is + 5);
          }
          mbar_wait(&sh.k_full[st_s], (is / kStages) & 1);
        }
        if (do_s && do_o) {
          if (RMCL_TC_MERGED_ISSUE == 1 || !live) {
            if (live) mbar_wait(&sh.p_full[io & 1], (io >> 1) & 1);
            merged = true;
          }
        }
      }
      // one code instance of each GEMM's issue sequence (instruction cache): pass 0 issues S (and O when merged), pass 1 the
      // O GEMM that still had to wait for its P tile
#pragma unroll 1
      for (int pass = 0; pass < 2; ++pass) {
        const bool s_now = do_s && pass == 0;
        const bool o_now = do_o && (pass == (merged ? 0 : 1));
        if (!s_now && !o_now) continue;
        if (live) {
          if (o_now && !merged) mbar_wait(&sh.p_full[io & 1], (io >> 1) & 1);
          tc_fence_after();
          if (s_now) tl_stamp(tl, 8 + 8 * is + 0);
          if (o_now) tl_stamp(tl, 8 + 8 * io + 3);
        }
        if (elect_one()) {
          if (s_now) issue_s();
          if (o_now) issue_o();
        }
        __syncwarp();
      }
    }
  }

  __syncthreads();
#if RMCL_TC_TIMELINE
  if (timeline != nullptr && tid == 0 && cta_linear < kTimelineCtas) timeline[kTimelineHead + 2 * cta_linear + 1] = global_ns();
#endif
  if (warp == kTmaWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
  if (FUSED) {
    // ---- phase C: every split partial of every row block is in global memory once all CTAs have passed this barrier
    grid_barrier(fz.bar_words, (unsigned)n_ctas, 0, kTcThreads);
    if (tid == 0) tl_stamp(tl, 7);
    if (warp < kSoftmaxWarps) {
      // two teams of 128 threads on alternate rows, each with its own 64 KB of the (now idle) TMA ring
      const int team = warp >> 2;
      float* fsm = reinterpret_cast<float*>(ring + (size_t)team * 65536);
      for (int row = cta_linear + team * n_ctas; row < fz.fin.B; row += 2 * n_ctas) {
#if RMCL_TC_TIMELINE
        finalize_row<__nv_bfloat16, 128>(fz.fin, row, fsm, &sh.fin[team], tid & 127, kBarFin + team,
                                         (timeline != nullptr && cta_linear == 0 && team == 0) ? timeline + 312 : nullptr);
#else
        finalize_row<__nv_bfloat16, 128>(fz.fin, row, fsm, &sh.fin[team], tid & 127, kBarFin + team);
#endif
      }
    }
#if RMCL_TC_TIMELINE
    __syncthreads();
    if (timeline != nullptr && tid == 0 && cta_linear < kTimelineCtas) timeline[kTimelineHead + 2 * cta_linear + 1] = global_ns();
#endif
  }
}

// ------------------------------------------------------------------------------ host side
thread_local long long* g_tc_timeline = nullptr;

template <typename... KArgs, typename... Args>
inline cudaError_t launch_cooperative(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeCooperative;     // all CTAs resident at once: the grid barriers cannot deadlock
  at[0].val.cooperative = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

template <int C, int TN, bool DIAG, bool WANT_O, bool FUSED>
int launch_tc(const __nv_bfloat16* q_hat, const void* queue, int B, long long K, long long ldq, float scale2,
              const InfoNcePlan& p, InfoNcePartials out, int want_argmax, const FusedArgs& fz, cudaStream_t s) {
  alignas(64) CUtensorMap tmap;
  const int trc = make_tmap_bf16(&tmap, queue, (uint64_t)C, (uint64_t)K, (uint64_t)ldq, (uint32_t)C);
  if (trc != RMCL_OK) return trc;
  constexpr int kStageBytes = C * TN * 2;
  constexpr int kPBytes = (TN / 64) * 16384;
  constexpr int kStages = ((kSmemBudget - 2 * kPBytes) / kStageBytes) < 8 ? ((kSmemBudget - 2 * kPBytes) / kStageBytes) : 8;
  const size_t smem = (size_t)kStages * kStageBytes + 2 * kPBytes + 1024;
  auto kern = infonce_tc_kernel<C, TN, DIAG, WANT_O, FUSED>;
  RMCL_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(p.splits, p.row_blocks);
  // RMCL_B200_INFONCE_COOP=0: plain launch of the same kernel (measurement aid: cost of the cooperative attribute).  The grid
  // never exceeds the SM count and a CTA takes a whole SM's shared memory, so the CTAs are co-resident either way.
  static const bool coop = [] { const char* e = getenv("RMCL_B200_INFONCE_COOP"); return !(e && e[0] == '0'); }();
  if (FUSED && !coop) {
    kern<<<grid, dim3(kTcThreads), smem, s>>>(tmap, q_hat, B, K, scale2, p.cols_per_split, want_argmax, out.m, out.l, out.av, out.ai,
                                              reinterpret_cast<__nv_bfloat16*>(out.o), g_tc_timeline, out.n2, out.qn2, out.dist, fz,
                                              reinterpret_cast<const __nv_bfloat16*>(queue), ldq);
    RMCL_LAUNCH_OK("infonce_fused_kernel");
  } else if (FUSED)
    RMCL_CUDA_OK(launch_cooperative(kern, grid, dim3(kTcThreads), smem, s, tmap, q_hat, B, K, scale2, p.cols_per_split, want_argmax,
                                    out.m, out.l, out.av, out.ai, reinterpret_cast<__nv_bfloat16*>(out.o), g_tc_timeline, out.n2,
                                    out.qn2, out.dist, fz, reinterpret_cast<const __nv_bfloat16*>(queue), ldq));
  else
    RMCL_CUDA_OK(launch_pdl(kern, grid, dim3(kTcThreads), smem, s, tmap, q_hat, B, K, scale2, p.cols_per_split, want_argmax,
                            out.m, out.l, out.av, out.ai, reinterpret_cast<__nv_bfloat16*>(out.o), g_tc_timeline, out.n2,
                            out.qn2, out.dist, fz, reinterpret_cast<const __nv_bfloat16*>(queue), ldq));
  return RMCL_OK;
}

template <bool FUSED>
int dispatch_tc(const __nv_bfloat16* q_hat, const void* queue, int B, int C, long long K, long long ldq, float scale2,
                const InfoNcePlan& p, InfoNcePartials out, int want_argmax, int want_o, const FusedArgs& fz, cudaStream_t s) {
  if (p.row_blocks > 65535) {
    set_error("InfoNCE: too many rows (%d)", B);
    return RMCL_E_UNSUPPORTED_DIM;
  }
  const bool dg = out.n2 != nullptr;
#define RMCL_TC_CASE(CC, TT)                                                                                              \
  case CC:                                                                                                                \
    if (want_o)                                                                                                           \
      return dg ? launch_tc<CC, TT, true, true, FUSED>(q_hat, queue, B, K, ldq, scale2, p, out, want_argmax, fz, s)       \
                : launch_tc<CC, TT, false, true, FUSED>(q_hat, queue, B, K, ldq, scale2, p, out, want_argmax, fz, s);     \
    return dg ? launch_tc<CC, TT, true, false, FUSED>(q_hat, queue, B, K, ldq, scale2, p, out, want_argmax, fz, s)        \
              : launch_tc<CC, TT, false, false, FUSED>(q_hat, queue, B, K, ldq, scale2, p, out, want_argmax, fz, s);
  switch (C) {
    RMCL_TC_CASE(256, 64)
    RMCL_TC_CASE(128, 128)
    RMCL_TC_CASE(64, 128)
  }
#undef RMCL_TC_CASE
  set_error("tcgen05 InfoNCE supports C in {64,128,256} (got %d)", C);
  return RMCL_E_UNSUPPORTED_DIM;
}

}  // namespace

int infonce_tc_tile_cols(int C) { return C == 256 ? 64 : 128; }

}  // namespace rmcl

// Measurement aid: the next tcgen05 InfoNCE launches issued by this thread make CTA (0,0) write its
// clock64() protocol timeline (layout above tl_stamp) into `dev_buf` (>= rmcl_debug_tc_timeline_words()
// int64, device memory); pass NULL to switch it off again.
extern "C" int rmcl_debug_tc_timeline(long long* dev_buf) {
#if !RMCL_TC_TIMELINE
  if (dev_buf != nullptr) {
    rmcl::set_error("rmcl_debug_tc_timeline: this build has no timeline instrumentation (build with -DRMCL_TC_TIMELINE=1)");
    return RMCL_E_UNSUPPORTED_DIM;
  }
#endif
  rmcl::g_tc_timeline = dev_buf;
  return RMCL_OK;
}
extern "C" int rmcl_debug_tc_timeline_words(void) { return rmcl::kTimelineWords; }

namespace rmcl {

int infonce_tc_launch(const __nv_bfloat16* q_hat, const void* queue, int B, int C, long long K, long long ldq,
                      float scale2, const InfoNcePlan& p, InfoNcePartials out, int want_argmax, int want_o, cudaStream_t s) {
  FusedArgs none = {};
  return dispatch_tc<false>(q_hat, queue, B, C, K, ldq, scale2, p, out, want_argmax, want_o, none, s);
}

int infonce_tc_fused_launch(const PrepArgs& prep, bool q_bf16, bool k_bf16, const FinArgs& fin, const void* queue, int B, int C,
                            long long K, long long ldq, float scale2, const InfoNcePlan& p, InfoNcePartials out, int want_argmax,
                            int want_o, cudaStream_t s) {
  FusedArgs fz;
  fz.prep = prep;
  fz.fin = fin;
  fz.bar_words = fin.counter + 2;     // [0], [1] of the counter block belong to the finalize stage
  fz.q_bf16 = q_bf16;
  fz.k_bf16 = k_bf16;
  return dispatch_tc<true>(prep.q_hat_bf16, queue, B, C, K, ldq, scale2, p, out, want_argmax, want_o, fz, s);
}

}  // namespace rmcl
