// Row-wise stages of the fused InfoNCE (see infonce.cu for the overall structure), as device functions so that the
// same code runs as stand-alone kernels (prep / finalize around the SIMT and two-pass partial kernels) and as the first
// and last phase of the single-launch tcgen05 kernel (infonce_tc.cu):
//
//   prep_row      one row of q (and k): q^ = q/max(|q|,1e-12) (and k^ when asked), positive logit, bf16 operand copies
//   finalize_row  one row: merge the splits with the positive, emit lse / loss / argmax,
//                 dq^ = (sum_j p_j queue_j + (p_pos-1) k^)/tau * loss_scale/B, then through the normalisation Jacobian:
//                 dq = (dq^ - q^ (q^.dq^)) / max(|q|,1e-12); the caller that completes the last row reduces the per-row
//                 losses in index order (deterministic).
//
// Both are written for a fixed team of threads (128 / NT) inside a possibly larger CTA and synchronise the team with
// named barriers only.
#pragma once
#include "infonce.cuh"

namespace rmcl {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

// named barrier ids (0 is __syncthreads; 1-2 are used inside finalize_row; the tcgen05 kernels use 1..12 in their main
// phase, which never overlaps with the row phases)
constexpr uint32_t kBarPrep = 13, kBarFin = 14;

__device__ __forceinline__ void team_sync(uint32_t id, uint32_t n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

__device__ __forceinline__ float round_if(float x, bool to_bf16) {
  return to_bf16 ? __bfloat162float(__float2bfloat16_rn(x)) : x;
}

__device__ __forceinline__ float team_sum_128(float v, float* red, int tid, uint32_t bar) {
  v = warp_sum(v);
  team_sync(bar, 128);
  if ((tid & 31) == 0) red[tid >> 5] = v;
  team_sync(bar, 128);
  return red[0] + red[1] + red[2] + red[3];
}

// ------------------------------------------------------------------------------------ prep
struct PrepArgs {
  const void* q;               // [B,C] raw projections (fp32 or bf16)
  const void* k;               // [B,C] keys (raw when normalize_k)
  int B, C;
  float scale2;                // log2(e)/tau
  bool normalize_k, bf16_mode; // bf16_mode: q^ and k^ are rounded to bf16 before the positive dot product (autocast semantics)
  float* q_hat;                // [B,C] fp32
  float* k_hat;                // [B,C] fp32
  float* k_hat_out;            // optional user output
  float* inv_norm;             // [B]
  float* pos2;                 // [B] positive logit in log2 units
  float* qn2;                  // [B] |q^|^2
  __nv_bfloat16* q_hat_bf16;   // kQhatReplicas x [b_pad, C] (split: [b_pad, 2C] = [q_hi | q_lo]) or null
  int b_pad;
  bool split;
};

// A team of 128 threads (tid = 0..127 inside the team, synchronising on named barrier `bar`); `red` = the team's 4 floats of
// shared memory.  row may be a padding row (>= B): zero operand rows.
template <typename TQ, typename TKK>
__device__ __forceinline__ void prep_row(const PrepArgs& a, int row, float* red, int tid, uint32_t bar) {
  const int C = a.C;
  const int qw = a.split ? 2 * C : C;
  const size_t rep_stride = (size_t)a.b_pad * qw;   // kQhatReplicas copies of the bf16 operand (infonce.cuh)
  if (row >= a.B) {  // padding rows of the bf16 operand (the tcgen05 kernels read whole 128-row blocks)
    if (a.q_hat_bf16)
      for (int c = tid; c < qw; c += 128)
        for (int rep = 0; rep < kQhatReplicas; ++rep) a.q_hat_bf16[rep * rep_stride + (size_t)row * qw + c] = __float2bfloat16_rn(0.f);
    return;
  }
  const TQ* qr = reinterpret_cast<const TQ*>(a.q) + (size_t)row * C;
  const TKK* kr = reinterpret_cast<const TKK*>(a.k) + (size_t)row * C;
  // rolled loops on purpose (C/128 iterations): every CTA runs this once with a cold instruction cache, and the IEEE
  // divisions below inline to ~40 instructions each — unrolled by 4 the kernel was 3400 instructions for a few hundred flops
  float sq = 0.f, sk = 0.f;
#pragma unroll 1
  for (int c = tid; c < C; c += 128) {
    const float x = to_f32(qr[c]), y = to_f32(kr[c]);
    sq = fmaf(x, x, sq);
    sk = fmaf(y, y, sk);
  }
  sq = team_sum_128(sq, red, tid, bar);
  sk = team_sum_128(sk, red, tid, bar);
  const float qn = fmaxf(sqrtf(sq), 1e-12f);
  const float kn = a.normalize_k ? fmaxf(sqrtf(sk), 1e-12f) : 1.f;
  float dot = 0.f, qq = 0.f;
#pragma unroll 1
  for (int c = tid; c < C; c += 128) {
    const float qh = __fdiv_rn(to_f32(qr[c]), qn);
    qq = fmaf(qh, qh, qq);
    const float kh = a.normalize_k ? __fdiv_rn(to_f32(kr[c]), kn) : to_f32(kr[c]);
    a.q_hat[(size_t)row * C + c] = qh;
    a.k_hat[(size_t)row * C + c] = kh;
    if (a.k_hat_out) a.k_hat_out[(size_t)row * C + c] = kh;
    if (a.q_hat_bf16) {
      const __nv_bfloat16 qb = __float2bfloat16_rn(qh);
      const __nv_bfloat16 ql = __float2bfloat16_rn(qh - __bfloat162float(qb));
#pragma unroll
      for (int rep = 0; rep < kQhatReplicas; ++rep) {
        a.q_hat_bf16[rep * rep_stride + (size_t)row * qw + c] = qb;
        if (a.split) a.q_hat_bf16[rep * rep_stride + (size_t)row * qw + C + c] = ql;
      }
    }
    dot = fmaf(round_if(qh, a.bf16_mode), round_if(kh, a.bf16_mode), dot);
  }
  dot = team_sum_128(dot, red, tid, bar);
  qq = team_sum_128(qq, red, tid, bar);
  if (tid == 0) {
    a.inv_norm[row] = __fdiv_rn(1.f, qn);
    a.pos2[row] = dot * a.scale2;
    a.qn2[row] = qq;
  }
}

// The same row by ONE WARP (the single-launch kernel: a CTA owns one or two rows, latency is what counts): a lane keeps its
// C/32 elements of q and k in flight together, the reductions are shuffles, nothing synchronises beyond the warp.
template <typename TQ, typename TKK>
__device__ __forceinline__ void prep_row_warp(const PrepArgs& a, int row, int lane) {
  const int C = a.C;
  const int qw = a.split ? 2 * C : C;
  const size_t rep_stride = (size_t)a.b_pad * qw;
  if (row >= a.B) {
    if (a.q_hat_bf16)
      for (int c = lane; c < qw; c += 32)
        for (int rep = 0; rep < kQhatReplicas; ++rep) a.q_hat_bf16[rep * rep_stride + (size_t)row * qw + c] = __float2bfloat16_rn(0.f);
    return;
  }
  const TQ* qr = reinterpret_cast<const TQ*>(a.q) + (size_t)row * C;
  const TKK* kr = reinterpret_cast<const TKK*>(a.k) + (size_t)row * C;
  float sq = 0.f, sk = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float x = to_f32(qr[c]), y = to_f32(kr[c]);
    sq = fmaf(x, x, sq);
    sk = fmaf(y, y, sk);
  }
  sq = warp_sum(sq);
  sk = warp_sum(sk);
  const float qn = fmaxf(sqrtf(sq), 1e-12f);
  const float kn = a.normalize_k ? fmaxf(sqrtf(sk), 1e-12f) : 1.f;
  float dot = 0.f, qq = 0.f;
  for (int c = lane; c < C; c += 32) {       // second pass: L1 hits
    const float qh = __fdiv_rn(to_f32(qr[c]), qn);
    qq = fmaf(qh, qh, qq);
    const float kh = a.normalize_k ? __fdiv_rn(to_f32(kr[c]), kn) : to_f32(kr[c]);
    a.q_hat[(size_t)row * C + c] = qh;
    a.k_hat[(size_t)row * C + c] = kh;
    if (a.k_hat_out) a.k_hat_out[(size_t)row * C + c] = kh;
    if (a.q_hat_bf16) {
      const __nv_bfloat16 qb = __float2bfloat16_rn(qh);
      const __nv_bfloat16 ql = __float2bfloat16_rn(qh - __bfloat162float(qb));
#pragma unroll
      for (int rep = 0; rep < kQhatReplicas; ++rep) {
        a.q_hat_bf16[rep * rep_stride + (size_t)row * qw + c] = qb;
        if (a.split) a.q_hat_bf16[rep * rep_stride + (size_t)row * qw + C + c] = ql;
      }
    }
    dot = fmaf(round_if(qh, a.bf16_mode), round_if(kh, a.bf16_mode), dot);
  }
  dot = warp_sum(dot);
  qq = warp_sum(qq);
  if (lane == 0) {
    a.inv_norm[row] = __fdiv_rn(1.f, qn);
    a.pos2[row] = dot * a.scale2;
    a.qn2[row] = qq;
  }
}

__device__ __forceinline__ void prep_row_warp_rt(const PrepArgs& a, bool q_bf16, bool k_bf16, int row, int lane) {
  if (!q_bf16 && !k_bf16) prep_row_warp<float, float>(a, row, lane);
  else if (!q_bf16) prep_row_warp<float, __nv_bfloat16>(a, row, lane);
  else if (!k_bf16) prep_row_warp<__nv_bfloat16, float>(a, row, lane);
  else prep_row_warp<__nv_bfloat16, __nv_bfloat16>(a, row, lane);
}

// dtype dispatch for callers that carry the dtypes at run time (the single-launch kernel)
__device__ __forceinline__ void prep_row_rt(const PrepArgs& a, bool q_bf16, bool k_bf16, int row, float* red, int tid,
                                            uint32_t bar) {
  if (!q_bf16 && !k_bf16) prep_row<float, float>(a, row, red, tid, bar);
  else if (!q_bf16) prep_row<float, __nv_bfloat16>(a, row, red, tid, bar);
  else if (!k_bf16) prep_row<__nv_bfloat16, float>(a, row, red, tid, bar);
  else prep_row<__nv_bfloat16, __nv_bfloat16>(a, row, red, tid, bar);
}

// -------------------------------------------------------------------------------- finalize
struct FinArgs {
  int B, C, splits;
  float inv_tau, grad_scale /* loss_scale / B */, loss_scale;
  bool bf16_mode, want_grad;
  const float* q_hat;
  const float* k_hat;
  const float* inv_norm;
  const float* pos2;
  const float* pm;
  const float* pl;
  const float* pav;
  const int* pai;
  const void* po;              // [splits][B][C] partial accumulators, element type TP
  float* row_loss;
  unsigned int* counter;       // [0] rows finalized (zero on entry of the first row), [1] overflow flag of the two-pass S kernel
  float* loss;
  float* loss_per_row;
  float* lse_out;
  float* pos_out;
  long long* argmax_out;
  float* dq;
  float* dk;
  const float* pdist;
  const float* qn2;
  float inv_K;
  InfoNceDiag diag;
};

__device__ __forceinline__ float ld_partial(const float* p) { return __ldcs(p); }
__device__ __forceinline__ float ld_partial(const __nv_bfloat16* p) {
  return __bfloat162float(__ushort_as_bfloat16(__ldcs(reinterpret_cast<const unsigned short*>(p))));
}

// shared memory a finalize team needs: merge weights + per-group partial rows
__host__ __device__ inline size_t finalize_smem_bytes(int C, int splits, bool bf16_partials, int team_threads) {
  const int vc = bf16_partials ? 8 : 4;                       // columns per 16-byte load
  const int col_threads = team_threads < 256 ? team_threads : 256;
  size_t rows;
  if (C % vc == 0 && C / vc <= team_threads) rows = (size_t)(team_threads / (C / vc));   // 16-byte path: one row per split group
  else rows = (size_t)(team_threads / col_threads - 1);                                  // scalar path
  return ((size_t)((splits + 3) & ~3) + rows * C) * sizeof(float);
}

// weighted accumulation of one 16-byte load of partials into kVC column sums
__device__ __forceinline__ void fma_partial(const uint4& u, float w, float* acc, const __nv_bfloat16*) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = __bfloat1622float2(h[j]);
    acc[2 * j] = fmaf(f.x, w, acc[2 * j]);
    acc[2 * j + 1] = fmaf(f.y, w, acc[2 * j + 1]);
  }
}
__device__ __forceinline__ void fma_partial(const uint4& u, float w, float* acc, const float*) {
  acc[0] = fmaf(__uint_as_float(u.x), w, acc[0]);
  acc[1] = fmaf(__uint_as_float(u.y), w, acc[1]);
  acc[2] = fmaf(__uint_as_float(u.z), w, acc[2]);
  acc[3] = fmaf(__uint_as_float(u.w), w, acc[3]);
}

// per-team scratch of finalize_row
struct FinShared {
  float red[8];
  float dred[16][5];
  float s_stats[4];   // 0: scale applied to O  1: p_pos - 1  2: sum over the queue of |q^ - queue_j|
  int last;
};

// A team of NT threads (128, 256 or 512; tid = 0..NT-1 inside the team) synchronising on named barrier `bar`.  TP: element
// type of the partial accumulators (fp32 from the SIMT and split-operand kernels, bf16 from the bf16 tcgen05 kernels).
// fin_smem: finalize_smem_bytes() of shared memory, fs: the team's FinShared.  Re-entrant: a team may call it for several
// rows in turn, several teams of one CTA may run it concurrently on different rows (distinct bar / fin_smem / fs).
template <typename TP, int NT>
__device__ __forceinline__ void finalize_row(const FinArgs& a, const int row, float* fin_smem, FinShared* fs, const int tid,
                                             const uint32_t bar, long long* dbg = nullptr) {
#define RMCL_FIN_STAMP(i) do { if (dbg != nullptr && tid == 0) dbg[i] = clock64(); } while (0)
  // counter[1]: raised by the two-pass tcgen05 S kernel when a fixed split reference could not hold the
  // row maximum; nothing computed from those partials is meaningful, so every output becomes NaN.
  constexpr int kColThreads = NT < 256 ? NT : 256;   // "group 0": the threads that own output columns (c = ct + kColThreads*i)
  constexpr int kColsPer = 1024 / kColThreads;       // C <= 1024
  constexpr int kGroups = NT / kColThreads;          // split groups of the scalar path
  const uint32_t bar0 = (NT > 256) ? 1u : bar;       // barrier of group 0 alone (the whole team when NT <= 256)
  const uint32_t bar_last = (NT > 256) ? 2u : bar;
  const int B = a.B, C = a.C, splits = a.splits;
  const TP* po = reinterpret_cast<const TP*>(a.po);
  float* sw = fin_smem;                          // [splits] merge weights
  float* part = fin_smem + ((splits + 3) & ~3);  // per-group partial column sums
  const int grp = tid / kColThreads, ct = tid - grp * kColThreads;   // split group, column thread
  team_sync(bar, NT);                            // the previous row of this team is completely done with the shared arrays
  RMCL_FIN_STAMP(0);

  // The partial stream does not depend on the merge weights until the multiply: put the first
  // batch of loads in flight before waiting for the statistics.
  constexpr int kPre = 8;
  float pre[kPre];
  // 16-byte loads of the partials whenever a row is a whole number of them (bf16: 8 columns, fp32: 4 columns per load):
  // C/kVC threads per pass over a row and NT / (C/kVC) split groups; a thread's loads are issued in batches of kPreV, the
  // first batch before the statistics barrier (NT = 512 at cfg2: 5 loads per thread, one batch).  The fp32 partials of
  // the split-operand and CUDA-core kernels used to take the scalar path below: 73 dependent 4-byte loads per thread in
  // batches of four at the cfg4 shape = 19.9 us under ncu, the largest piece of that call.
  constexpr int kVC = 16 / (int)sizeof(TP);
  const bool kVec = (C % kVC == 0) && (C / kVC <= NT);
  constexpr int kPreV = NT >= 512 ? 6 : 20;   // every load of a row in flight at once (registers are free in the row phases)
  uint4 prev[kPreV];
  const int vpr = kVec ? C / kVC : NT;              // threads per row pass
  const int vgroups = kVec ? NT / vpr : 1;          // split groups
  const int vg = tid / vpr, vc = tid - vg * vpr;
  const bool vactive = kVec && a.want_grad && vg < vgroups;
  const size_t sstride4 = (size_t)B * C / kVC;
  const uint4* prow4 = reinterpret_cast<const uint4*>(po + (size_t)row * C) + vc;
  if (kVec) {
#pragma unroll
    for (int u = 0; u < kPreV; ++u) {
      const int s = vg + u * vgroups;
      prev[u] = (vactive && s < splits) ? __ldcs(prow4 + (size_t)s * sstride4) : make_uint4(0u, 0u, 0u, 0u);
    }
  } else {
    const TP* pcol = po + (size_t)row * C + ct;
    const size_t sstride = (size_t)B * C;
#pragma unroll
    for (int u = 0; u < kPre; ++u) {
      const int s = grp + u * kGroups;
      pre[u] = (a.want_grad && ct < C && s < splits) ? ld_partial(pcol + (size_t)s * sstride) : 0.f;
    }
  }

  // row data needed after the merge: in flight now
  const bool own0 = a.want_grad && grp == 0 && ct < C;
  float qh_pre = 0.f, kh_pre = 0.f, inv_pre = 0.f;
  if (own0) {
    qh_pre = __ldcg(a.q_hat + (size_t)row * C + ct);
    kh_pre = __ldcg(a.k_hat + (size_t)row * C + ct);
    inv_pre = __ldcg(a.inv_norm + row);
  }

  if (tid < 32) {
    // merge the split statistics (one warp; a row's statistics are contiguous => coalesced loads).  Rolled loops on purpose:
    // every CTA runs this once, with a cold instruction cache, so code size is latency here
    const float* rm = a.pm + (size_t)row * splits;
    const float* rl = a.pl + (size_t)row * splits;
    const float* rav = a.pav + (size_t)row * splits;
    const int* rai = a.pai + (size_t)row * splits;
    const float p2 = __ldcg(a.pos2 + row);
    const unsigned poison_flag = __ldcg(a.counter + 1);
    float mmax = -INFINITY;
    for (int s = tid; s < splits; s += 32) mmax = fmaxf(mmax, __ldcg(rm + s));
    mmax = warp_max(mmax);
    float lsum = 0.f;
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    if (a.diag.out) {   // the split sums of the L2 distances: plain sum, in fixed lane order
      float dsum = 0.f;
      for (int s = tid; s < splits; s += 32) dsum += __ldcg(a.pdist + (size_t)row * splits + s);
      dsum = warp_sum(dsum);
      if (tid == 0) fs->s_stats[2] = dsum;
    }
    for (int s = tid; s < splits; s += 32) {
      const float ms = __ldcg(rm + s), ls = __ldcg(rl + s), v = __ldcg(rav + s);
      const int i = __ldcg(rai + s);
      const float w = (ms == -INFINITY) ? 0.f : exp2f(ms - mmax);
      sw[s] = w;
      lsum = fmaf(ls, w, lsum);
      if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
    }
    lsum = warp_sum(lsum);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (tid == 0) {
      const float M = fmaxf(mmax, p2);
      const float wneg = exp2f(mmax - M), wpos = exp2f(p2 - M);
      const float L = fmaf(lsum, wneg, wpos);
      const float lse = (M + log2f(L)) * kLn2;
      const float pos = p2 * kLn2;
      // lse - pos cancels catastrophically when the positive dominates (p_pos -> 1); take the
      // difference before the log instead: positive is the max -> log1p of the remaining mass.
      const float poison = (poison_flag != 0u) ? __int_as_float(0x7fc00000) : 0.f;
      const float lrow = ((p2 >= mmax) ? log1pf(lsum * wneg) : fmaf(M - p2, kLn2, logf(L))) + poison;
      a.row_loss[row] = lrow;
      if (a.loss_per_row) a.loss_per_row[row] = lrow;
      if (a.lse_out) a.lse_out[row] = lse;
      if (a.pos_out) a.pos_out[row] = pos;
      if (a.argmax_out) a.argmax_out[row] = (p2 >= bv) ? 0ll : (long long)bi + 1;
      fs->s_stats[0] = wneg / L + poison;
      fs->s_stats[1] = -(lsum * wneg) / L + poison;  // p_pos - 1 without the cancellation of wpos/L - 1
    }
    RMCL_FIN_STAMP(1);
  }
  team_sync(bar, NT);
  RMCL_FIN_STAMP(2);

  if (a.want_grad) {
    // Column sums of the partials: group g streams splits g, g+G, ...; groups are then added in group order (deterministic).
    float acc[kColsPer];
#pragma unroll
    for (int i = 0; i < kColsPer; ++i) acc[i] = 0.f;
    const size_t sstride = (size_t)B * C;
    const TP* prow = po + (size_t)row * C;
    if (kVec) {
      // kVC columns per thread, weighted sum over this group's splits, then one row of partial sums per group in shared memory
      float av[kVC];
#pragma unroll
      for (int j = 0; j < kVC; ++j) av[j] = 0.f;
      if (vactive) {
        for (int base = 0; base < splits; base += kPreV * vgroups) {
          if (base > 0) {   // next batch: all of its loads in flight before the first use
#pragma unroll
            for (int u = 0; u < kPreV; ++u) {
              const int sp = base + vg + u * vgroups;
              prev[u] = (sp < splits) ? __ldcs(prow4 + (size_t)sp * sstride4) : make_uint4(0u, 0u, 0u, 0u);
            }
          }
#pragma unroll
          for (int u = 0; u < kPreV; ++u) {
            const int sp = base + vg + u * vgroups;
            if (sp < splits) fma_partial(prev[u], sw[sp], av, static_cast<const TP*>(nullptr));
          }
        }
        float* dst = part + (size_t)vg * C + vc * kVC;
#pragma unroll
        for (int j = 0; j < kVC; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(av[j], av[j + 1], av[j + 2], av[j + 3]);
      }
    } else {
#pragma unroll
      for (int i = 0; i < kColsPer; ++i) {
        const int c = ct + kColThreads * i;
        if (c < C) {
          const TP* pcol = prow + c;
          int s = grp;
          if (i == 0) {  // the prefetched batch
#pragma unroll
            for (int u = 0; u < kPre; ++u) {
              const int sp = grp + u * kGroups;
              if (sp < splits) acc[0] = fmaf(pre[u], sw[sp], acc[0]);
            }
            s = grp + kPre * kGroups;
          }
          for (; s + 3 * kGroups < splits; s += 4 * kGroups) {
            float v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = ld_partial(pcol + (size_t)(s + u * kGroups) * sstride);
#pragma unroll
            for (int u = 0; u < 4; ++u) acc[i] = fmaf(v[u], sw[s + u * kGroups], acc[i]);
          }
          for (; s < splits; s += kGroups) acc[i] = fmaf(ld_partial(pcol + (size_t)s * sstride), sw[s], acc[i]);
          if (grp > 0) part[(size_t)(grp - 1) * C + c] = acc[i];
        }
      }
    }
    RMCL_FIN_STAMP(3);
    team_sync(bar, NT);
    RMCL_FIN_STAMP(4);
    if (grp == 0) {
      const float o_scale = fs->s_stats[0], pm1 = fs->s_stats[1];
      const float gs = a.grad_scale * a.inv_tau;
      float dqh[kColsPer], qh[kColsPer];
      float dot = 0.f;
#pragma unroll
      for (int i = 0; i < kColsPer; ++i) {
        const int c = ct + kColThreads * i;
        dqh[i] = 0.f;
        qh[i] = 0.f;
        if (c < C) {
          float s = acc[i];
          if (kVec) {
            s = 0.f;
            for (int g = 0; g < vgroups; ++g) s += part[(size_t)g * C + c];     // fixed order: deterministic
          } else {
#pragma unroll
            for (int g = 1; g < kGroups; ++g) s += part[(size_t)(g - 1) * C + c];
          }
          const float kh = round_if(i == 0 ? kh_pre : __ldcg(a.k_hat + (size_t)row * C + c), a.bf16_mode);
          qh[i] = (i == 0) ? qh_pre : __ldcg(a.q_hat + (size_t)row * C + c);
          dqh[i] = gs * fmaf(s, o_scale, pm1 * kh);
          dot = fmaf(qh[i], dqh[i], dot);
          if (a.dk) a.dk[(size_t)row * C + c] = gs * pm1 * round_if(qh[i], a.bf16_mode);
        }
      }
      dot = warp_sum(dot);
      if ((ct & 31) == 0) fs->red[ct >> 5] = dot;
      team_sync(bar0, kColThreads);   // the warps of group 0 only
      dot = 0.f;
#pragma unroll
      for (int w = 0; w < kColThreads / 32; ++w) dot += fs->red[w];
      const float inv = own0 ? inv_pre : __ldcg(a.inv_norm + row);
      if (a.dq) {
#pragma unroll
        for (int i = 0; i < kColsPer; ++i) {
          const int c = ct + kColThreads * i;
          if (c < C) a.dq[(size_t)row * C + c] = (dqh[i] - qh[i] * dot) * inv;
        }
      }
    }
  }

  RMCL_FIN_STAMP(5);
  // ---- diagnostics of this row (objectives.py:337-349): five dot products over C, then closed forms
  if (a.diag.out) {
    float d5[5] = {0.f, 0.f, 0.f, 0.f, 0.f};   // q.k, |q-k|^2, |k|^2, q.sum_vec, q.sum_unit
    for (int c = tid; c < C; c += NT) {
      const float qh = __ldcg(a.q_hat + (size_t)row * C + c), kh = __ldcg(a.k_hat + (size_t)row * C + c);
      const float df = qh - kh;
      d5[0] = fmaf(qh, kh, d5[0]);
      d5[1] = fmaf(df, df, d5[1]);
      d5[2] = fmaf(kh, kh, d5[2]);
      d5[3] = fmaf(qh, a.diag.sum_vec[c], d5[3]);
      d5[4] = fmaf(qh, a.diag.sum_unit[c], d5[4]);
    }
#pragma unroll
    for (int i = 0; i < 5; ++i) d5[i] = warp_sum(d5[i]);
    if ((tid & 31) == 0) {
#pragma unroll
      for (int i = 0; i < 5; ++i) fs->dred[tid >> 5][i] = d5[i];
    }
    team_sync(bar, NT);
    if (tid == 0) {
      float t[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
      for (int w = 0; w < NT / 32; ++w)
#pragma unroll
        for (int i = 0; i < 5; ++i) t[i] += fs->dred[w][i];
      const float qn = sqrtf(__ldcg(a.qn2 + row)), kn = sqrtf(t[2]);
      float* o = a.diag.rows + (size_t)row * kDiagValues;
      o[0] = sqrtf(t[1]);                                                           // |q^ - k^|
      o[1] = t[0] / (fmaxf(qn, a.diag.cos_eps) * fmaxf(kn, a.diag.cos_eps));        // cosine(q^, k^)
      o[2] = t[0];                                                                  // q^ . k^
      o[3] = fs->s_stats[2] * a.inv_K;                                              // mean_j |q^ - queue_j|
      o[4] = t[4] * a.inv_K / fmaxf(qn, a.diag.cos_eps);                            // mean_j cosine(q^, queue_j)
      o[5] = t[3] * a.inv_K;                                                        // mean_j q^ . queue_j
    }
  }

  // deterministic loss (and diagnostics) reduction by whoever finalizes the last row
  if (a.loss || a.diag.out) {
    // thread 0 wrote this row's loss and diagnostics itself: its own fence + the counter is all the ordering the final
    // reduction needs (dq / dk rows of the other threads are not read by it)
    if (tid == 0) {
      __threadfence();
      fs->last = (atomicAdd(a.counter, 1u) == (unsigned)B - 1u) ? 1 : 0;
    }
    team_sync(bar, NT);
    if (fs->last && tid < kColThreads) {
      __threadfence();
      const int n_red = a.diag.out ? 1 + kDiagValues : 1;
      for (int which = a.loss ? 0 : 1; which < n_red; ++which) {
        float acc = 0.f;
        for (int r = tid; r < B; r += kColThreads)
          acc += (which == 0) ? __ldcg(a.row_loss + r) : __ldcg(a.diag.rows + (size_t)r * kDiagValues + (which - 1));
        // fixed-shape tree: warp shuffle then the warps' partials in order
        acc = warp_sum(acc);
        team_sync(bar_last, kColThreads);
        if ((tid & 31) == 0) fs->red[tid >> 5] = acc;
        team_sync(bar_last, kColThreads);
        if (tid == 0) {
          float t = 0.f;
          for (int w = 0; w < kColThreads / 32; ++w) t += fs->red[w];
          if (which == 0) *a.loss = t * (a.loss_scale / (float)B);
          else a.diag.out[which - 1] = t / (float)B;
        }
      }
    }
  }
  RMCL_FIN_STAMP(6);
#undef RMCL_FIN_STAMP
}

// ---- grid-wide barrier of a kernel whose CTAs are all co-resident (cooperative launch).  words[0] = arrival count (zero
// between uses: the last arriver resets it), words[1] = generation.  Called by the `n_threads` threads of the CTA that
// synchronise on named barrier `bar_id` (0 with n_threads = blockDim.x is the whole CTA); thread 0 must be one of them.
__device__ __forceinline__ void grid_barrier(unsigned int* words, unsigned int n_ctas, uint32_t bar_id, uint32_t n_threads) {
  team_sync(bar_id, n_threads);
  if (threadIdx.x == 0) {
    unsigned int gen;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(gen) : "l"(words + 1) : "memory");
    __threadfence();
    if (atomicAdd(words, 1u) == n_ctas - 1u) {
      words[0] = 0u;
      __threadfence();
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(words + 1) : "memory");
    } else {
      unsigned int now;
      const long long t0 = clock64();
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(now) : "l"(words + 1) : "memory");
        if (clock64() - t0 > 8000000000ll) __trap();     // a CTA that never arrives becomes an error, not a hang
      } while (now == gen);
    }
    __threadfence();
  }
  team_sync(bar_id, n_threads);
}

// What the single-launch kernel needs beyond the flash pass itself
struct FusedArgs {
  PrepArgs prep;
  FinArgs fin;
  unsigned int* bar_words;   // grid barrier: [0] count, [1] generation (workspace, zero before the first use)
  bool q_bf16, k_bf16;
};

}  // namespace rmcl
