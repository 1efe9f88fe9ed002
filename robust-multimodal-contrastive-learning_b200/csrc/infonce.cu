// K1 — fused InfoNCE forward/backward: host dispatch + the prep and finalize stages.
//
// Replaces, per call site (vilt/modules/objectives.py:269-274, 326-334+351;
// attack/pgd_attack_vilt.py:147,152-158; MoCo/MoCo_RMCL.py:150-164) the eager chain
//   F.normalize -> queue.clone() -> einsum(nc,nc->n) -> einsum(nc,ck->nk) -> cat -> /T ->
//   CrossEntropyLoss(logits.float(), 0) -> autograd backward
// without ever materialising the B x (K+1) logits.
//
//   prep      one CTA per row: q^ = q/max(|q|,1e-12) (and k^ when asked), positive logit.
//   partial   split-K flash pass over the queue (infonce_simt.cu / infonce_tc.cu).
//   finalize  one CTA per row: merge the splits with the positive, emit lse / loss / argmax,
//             dq^ = (sum_j p_j queue_j + (p_pos-1) k^)/tau * loss_scale/B, then push it through
//             the normalisation Jacobian: dq = (dq^ - q^ (q^.dq^)) / max(|q|,1e-12).
//             The last row-CTA to finish reduces the per-row losses in index order (deterministic).
//
// bf16 mode (queue_dtype == bf16) mirrors the reference under autocast: q^ and k^ are rounded to
// bf16 before the dot products, accumulation is fp32, the Jacobian uses the fp32 q^.
#include "infonce.cuh"

namespace rmcl {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

int infonce_make_plan(int B, int C, long long K, int queue_dtype, int path, bool aligned_for_tc, InfoNcePlan* p) {
  const int sms = sm_count();
  if (sms <= 0) return RMCL_E_CUDA;
  const bool hilo = (queue_dtype == RMCL_BF16_HILO);   // fp32-accurate split-operand path (two-pass kernels, SPLIT)
  const bool wide = infonce_tc2_supports(C);   // two-pass tcgen05 variant (infonce_tc2.cu)
  const bool tc_ok = infonce_tc_built() && aligned_for_tc && (K % 8 == 0) &&
                     (hilo ? infonce_tc2_split_supports(C) : (queue_dtype == RMCL_BF16 && (C == 64 || C == 128 || C == 256 || wide)));
  if (hilo && path == RMCL_INFONCE_SIMT) {
    set_error("a bf16 hi/lo queue (RMCL_BF16_HILO) is an operand of the tcgen05 path only");
    return RMCL_E_UNSUPPORTED_DIM;
  }
  if (path == RMCL_INFONCE_AUTO) path = (tc_ok || hilo) ? RMCL_INFONCE_TCGEN05 : RMCL_INFONCE_SIMT;
  if (path == RMCL_INFONCE_TCGEN05 && !tc_ok) {
    set_error("tcgen05 InfoNCE needs a 16B-aligned bf16 queue with C in {64,128,256,512,768} or a bf16 hi/lo queue with C in "
              "{64,128,256}, and K %% 8 == 0 (got C=%d K=%lld)", C, K);
    return RMCL_E_UNSUPPORTED_DIM;
  }
  if (path == RMCL_INFONCE_SIMT && C > 1024) {
    set_error("SIMT InfoNCE supports C <= 1024 (got %d)", C);
    return RMCL_E_UNSUPPORTED_DIM;
  }
  p->path = path;
  if (path == RMCL_INFONCE_SIMT) {
    p->rows_per_cta = 16;
    p->tile_cols = (C <= 512) ? 64 : 32;
  } else {
    p->rows_per_cta = 128;
    p->tile_cols = (wide || hilo) ? 64 : infonce_tc_tile_cols(C);
  }
  p->split = (path == RMCL_INFONCE_TCGEN05) && hilo;
  p->two_pass = (path == RMCL_INFONCE_TCGEN05) && (wide || hilo);
  p->row_blocks = (B + p->rows_per_cta - 1) / p->rows_per_cta;
  p->b_pad = p->row_blocks * p->rows_per_cta;
  const long long tiles = (K + p->tile_cols - 1) / p->tile_cols;
  // SIMT: ~2 waves of CTAs; TC: one persistent CTA per SM
  long long target = (path == RMCL_INFONCE_SIMT) ? 2ll * sms : sms;
  long long splits = target / p->row_blocks;
  if (splits < 1) splits = 1;
  if (splits > tiles) splits = tiles;
  const long long tiles_per_split = (tiles + splits - 1) / splits;
  p->cols_per_split = tiles_per_split * p->tile_cols;
  p->splits = (int)((K + p->cols_per_split - 1) / p->cols_per_split);
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
  const size_t S = (size_t)p->splits, Bs = (size_t)B;
  p->off_qhat = take(Bs * C * 4);
  p->off_khat = take(Bs * C * 4);
  p->off_inv = take(Bs * 4);
  p->off_pos2 = take(Bs * 4);
  p->off_qhat_bf16 = take((size_t)kQhatReplicas * p->b_pad * C * 2 * (p->split ? 2 : 1));   // split: rows are [q_hi | q_lo]
  p->off_m = take(S * Bs * 4);
  p->off_l = take(S * Bs * 4);
  p->off_av = take(S * Bs * 4);
  p->off_ai = take(S * Bs * 4);
  p->off_o = take(S * Bs * C * 4);
  p->off_rowloss = take(Bs * 4);
  p->off_counter = take(256);
  p->off_qn2 = take(Bs * 4);
  p->off_pdist = take(S * Bs * 4);
  p->off_diagrows = take(Bs * kDiagValues * 4);
  p->k_pad = (K + 63) / 64 * 64;
  p->off_ptilde = take(p->two_pass ? (size_t)p->b_pad * (size_t)p->k_pad * 2 * (p->split ? 2 : 1) : 0);   // split: hi and lo planes
  p->total = off;
  return RMCL_OK;
}

__device__ __forceinline__ float round_if(float x, bool to_bf16) {
  return to_bf16 ? __bfloat162float(__float2bfloat16_rn(x)) : x;
}

__device__ __forceinline__ float block_sum_128(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  return red[0] + red[1] + red[2] + red[3];
}

// ------------------------------------------------------------------------------------ prep
template <typename TQ, typename TKK>
__global__ void __launch_bounds__(128) infonce_prep_kernel(const TQ* __restrict__ q, const TKK* __restrict__ k, int B,
                                                           int C, float scale2, bool normalize_k, bool bf16_mode,
                                                           float* __restrict__ q_hat, float* __restrict__ k_hat,
                                                           float* __restrict__ k_hat_out, float* __restrict__ inv_norm,
                                                           float* __restrict__ pos2, float* __restrict__ qn2,
                                                           __nv_bfloat16* __restrict__ q_hat_bf16, int b_pad, bool split,
                                                           unsigned int* __restrict__ counter) {
  __shared__ float red[4];
  const int row = blockIdx.x;
  if (threadIdx.x == 0) pdl_trigger();   // the partial kernel may set itself up while this one runs
  if (row == 0 && threadIdx.x == 0) {
    counter[0] = 0u;
    counter[1] = 0u;   // overflow flag of the two-pass tcgen05 variant
  }
  // kQhatReplicas copies of the bf16 operand (infonce.cuh); split: a row is [q_hi | q_lo], 2C wide
  const int qw = split ? 2 * C : C;
  const size_t rep_stride = (size_t)b_pad * qw;
  if (row >= B) {  // padding rows of the bf16 operand (the tcgen05 kernel reads whole 128-row blocks)
    for (int c = threadIdx.x; c < qw; c += 128)
      for (int rep = 0; rep < kQhatReplicas; ++rep) q_hat_bf16[rep * rep_stride + (size_t)row * qw + c] = __float2bfloat16_rn(0.f);
    return;
  }
  const TQ* qr = q + (size_t)row * C;
  const TKK* kr = k + (size_t)row * C;
  float sq = 0.f, sk = 0.f;
  for (int c = threadIdx.x; c < C; c += 128) {
    const float a = to_f32(qr[c]), b = to_f32(kr[c]);
    sq = fmaf(a, a, sq);
    sk = fmaf(b, b, sk);
  }
  sq = block_sum_128(sq, red);
  sk = block_sum_128(sk, red);
  const float qn = fmaxf(sqrtf(sq), 1e-12f);
  const float kn = normalize_k ? fmaxf(sqrtf(sk), 1e-12f) : 1.f;
  float dot = 0.f, qq = 0.f;
  for (int c = threadIdx.x; c < C; c += 128) {
    const float qh = __fdiv_rn(to_f32(qr[c]), qn);
    qq = fmaf(qh, qh, qq);
    const float kh = normalize_k ? __fdiv_rn(to_f32(kr[c]), kn) : to_f32(kr[c]);
    q_hat[(size_t)row * C + c] = qh;
    k_hat[(size_t)row * C + c] = kh;
    if (k_hat_out) k_hat_out[(size_t)row * C + c] = kh;
    if (q_hat_bf16) {
      const __nv_bfloat16 qb = __float2bfloat16_rn(qh);
      const __nv_bfloat16 ql = __float2bfloat16_rn(qh - __bfloat162float(qb));
#pragma unroll
      for (int rep = 0; rep < kQhatReplicas; ++rep) {
        q_hat_bf16[rep * rep_stride + (size_t)row * qw + c] = qb;
        if (split) q_hat_bf16[rep * rep_stride + (size_t)row * qw + C + c] = ql;
      }
    }
    dot = fmaf(round_if(qh, bf16_mode), round_if(kh, bf16_mode), dot);
  }
  dot = block_sum_128(dot, red);
  qq = block_sum_128(qq, red);
  if (threadIdx.x == 0) {
    inv_norm[row] = __fdiv_rn(1.f, qn);
    pos2[row] = dot * scale2;
    qn2[row] = qq;
  }
}

// -------------------------------------------------------------------------------- finalize
constexpr int kFinGroups = 2;                 // split groups streaming the partials concurrently (512 threads:
                                              // two CTAs per SM, so B=256 rows are one wave on 148 SMs)
constexpr int kFinThreads = 256 * kFinGroups;

__device__ __forceinline__ float ld_partial(const float* p) { return __ldcs(p); }
__device__ __forceinline__ float ld_partial(const __nv_bfloat16* p) {
  return __bfloat162float(__ushort_as_bfloat16(__ldcs(reinterpret_cast<const unsigned short*>(p))));
}

// TP: element type of the partial accumulators (fp32 from the SIMT kernel, bf16 from the tcgen05 kernel)
template <typename TP>
__global__ void __launch_bounds__(kFinThreads, 2) infonce_finalize_kernel(
    int B, int C, int splits, float inv_tau, float grad_scale /* loss_scale / B */, float loss_scale, bool bf16_mode,
    bool want_grad, const float* __restrict__ q_hat, const float* __restrict__ k_hat, const float* __restrict__ inv_norm,
    const float* __restrict__ pos2, const float* __restrict__ pm, const float* __restrict__ pl,
    const float* __restrict__ pav, const int* __restrict__ pai, const TP* __restrict__ po,
    float* __restrict__ row_loss, unsigned int* __restrict__ counter, float* __restrict__ loss,
    float* __restrict__ loss_per_row, float* __restrict__ lse_out, float* __restrict__ pos_out,
    long long* __restrict__ argmax_out, float* __restrict__ dq, float* __restrict__ dk, const float* __restrict__ pdist,
    const float* __restrict__ qn2, float inv_K, const InfoNceDiag diag) {
  // counter[1]: raised by the two-pass tcgen05 S kernel when a fixed split reference could not hold the
  // row maximum; nothing computed from those partials is meaningful, so every output becomes NaN.
  extern __shared__ float fin_smem[];
  float* sw = fin_smem;                       // [splits] merge weights
  float* part = fin_smem + ((splits + 3) & ~3);  // [kFinGroups-1][C] partial column sums of groups 1..
  __shared__ float red[8];
  __shared__ float dred[kFinThreads / 32][5];
  __shared__ float s_stats[4];   // 0: scale applied to O  1: p_pos - 1  2: sum over the queue of |q^ - queue_j|
  __shared__ bool s_last;
  const int row = blockIdx.x;
  const int tid = threadIdx.x;
  const int grp = tid >> 8, ct = tid & 255;   // split group, column thread
  pdl_wait();                                 // launched early (PDL): the partials must be complete

  // The partial stream does not depend on the merge weights until the multiply: put the first
  // batch of loads (column ct, splits grp, grp+G, ...) in flight before waiting for the statistics.
  constexpr int kPre = 8;
  float pre[kPre];
  // bf16 partials (tcgen05 kernels; C % 8 == 0): 16-byte loads, 8 columns per thread, C/8 threads per pass over a row and
  // kFinThreads / (C/8) split groups, so that a thread needs only ~splits/groups loads (5 at cfg2) and all of them are in
  // flight before the statistics barrier.  (The scalar path below issued 37 two-byte loads per thread in 5 dependent rounds.)
  constexpr bool kVec = (sizeof(TP) == 2);
  constexpr int kPreV = 6;
  uint4 prev[kVec ? kPreV : 1];
  const int vpr = C >> 3;                           // threads per row pass
  const int vgroups = kVec ? kFinThreads / vpr : 1;  // split groups
  const int vg = tid / vpr, vc = tid - vg * vpr;
  const bool vactive = kVec && want_grad && vg < vgroups;
  if (kVec) {
    const uint4* prow4 = reinterpret_cast<const uint4*>(po + (size_t)row * C) + vc;
    const size_t sstride4 = (size_t)B * C / 8;
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
      const int s = grp + u * kFinGroups;
      pre[u] = (want_grad && ct < C && s < splits) ? ld_partial(pcol + (size_t)s * sstride) : 0.f;
    }
  }

  // row data needed after the merge: in flight now
  const bool own0 = want_grad && grp == 0 && ct < C;
  float qh_pre = 0.f, kh_pre = 0.f;
  if (own0) {
    qh_pre = q_hat[(size_t)row * C + ct];
    kh_pre = k_hat[(size_t)row * C + ct];
  }

  if (tid < 32) {
    // merge the split statistics (one warp; a row's statistics are contiguous => coalesced loads)
    const float* rm = pm + (size_t)row * splits;
    const float* rl = pl + (size_t)row * splits;
    const float* rav = pav + (size_t)row * splits;
    const int* rai = pai + (size_t)row * splits;
    float mmax = -INFINITY;
    for (int s = tid; s < splits; s += 32) mmax = fmaxf(mmax, __ldcg(rm + s));
    mmax = warp_max(mmax);
    float lsum = 0.f;
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    if (diag.out) {   // the split sums of the L2 distances: plain sum, in fixed lane order
      float dsum = 0.f;
      for (int s = tid; s < splits; s += 32) dsum += __ldcg(pdist + (size_t)row * splits + s);
      dsum = warp_sum(dsum);
      if (tid == 0) s_stats[2] = dsum;
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
      const float p2 = pos2[row];
      const float M = fmaxf(mmax, p2);
      const float wneg = exp2f(mmax - M), wpos = exp2f(p2 - M);
      const float L = fmaf(lsum, wneg, wpos);
      const float lse = (M + log2f(L)) * kLn2;
      const float pos = p2 * kLn2;
      // lse - pos cancels catastrophically when the positive dominates (p_pos -> 1); take the
      // difference before the log instead: positive is the max -> log1p of the remaining mass.
      const float poison = (__ldcg(counter + 1) != 0u) ? __int_as_float(0x7fc00000) : 0.f;
      const float lrow = ((p2 >= mmax) ? log1pf(lsum * wneg) : fmaf(M - p2, kLn2, logf(L))) + poison;
      row_loss[row] = lrow;
      if (loss_per_row) loss_per_row[row] = lrow;
      if (lse_out) lse_out[row] = lse;
      if (pos_out) pos_out[row] = pos;
      if (argmax_out) argmax_out[row] = (p2 >= bv) ? 0ll : (long long)bi + 1;
      s_stats[0] = wneg / L + poison;
      s_stats[1] = -(lsum * wneg) / L + poison;  // p_pos - 1 without the cancellation of wpos/L - 1
    }
  }
  __syncthreads();

  if (want_grad) {
    // Column sums of the partials: group g streams splits g, g+G, ... with 4 independent loads in
    // flight per owned column; groups are then added in group order (deterministic).
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const size_t sstride = (size_t)B * C;
    const TP* prow = po + (size_t)row * C;
    if (kVec) {
      // 8 columns per thread, weighted sum over this group's splits, then one row of partial sums per group in shared memory
      float a8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (vactive) {
        const uint4* prow4 = reinterpret_cast<const uint4*>(prow) + vc;
        const size_t sstride4 = sstride / 8;
        auto fma8 = [&](const uint4& u, float w) {
          const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 f = __bfloat1622float2(h[j]);
            a8[2 * j] = fmaf(f.x, w, a8[2 * j]);
            a8[2 * j + 1] = fmaf(f.y, w, a8[2 * j + 1]);
          }
        };
#pragma unroll
        for (int u = 0; u < kPreV; ++u) {
          const int sp = vg + u * vgroups;
          if (sp < splits) fma8(prev[u], sw[sp]);
        }
        for (int sp = vg + kPreV * vgroups; sp < splits; sp += vgroups) fma8(__ldcs(prow4 + (size_t)sp * sstride4), sw[sp]);
        float4* dst = reinterpret_cast<float4*>(part + (size_t)vg * C + vc * 8);
        dst[0] = make_float4(a8[0], a8[1], a8[2], a8[3]);
        dst[1] = make_float4(a8[4], a8[5], a8[6], a8[7]);
      }
    } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = ct + 256 * i;
      if (c < C) {
        const TP* pcol = prow + c;
        int s = grp;
        if (i == 0) {  // the prefetched batch
#pragma unroll
          for (int u = 0; u < kPre; ++u) {
            const int sp = grp + u * kFinGroups;
            if (sp < splits) acc[0] = fmaf(pre[u], sw[sp], acc[0]);
          }
          s = grp + kPre * kFinGroups;
        }
        for (; s + 3 * kFinGroups < splits; s += 4 * kFinGroups) {
          float v[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) v[u] = ld_partial(pcol + (size_t)(s + u * kFinGroups) * sstride);
#pragma unroll
          for (int u = 0; u < 4; ++u) acc[i] = fmaf(v[u], sw[s + u * kFinGroups], acc[i]);
        }
        for (; s < splits; s += kFinGroups) acc[i] = fmaf(ld_partial(pcol + (size_t)s * sstride), sw[s], acc[i]);
        if (grp > 0) part[(size_t)(grp - 1) * C + c] = acc[i];
      }
    }
    }
    __syncthreads();
    if (grp == 0) {
      const float o_scale = s_stats[0], pm1 = s_stats[1];
      const float gs = grad_scale * inv_tau;
      float dqh[4], qh[4];
      float dot = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = ct + 256 * i;
        dqh[i] = 0.f;
        qh[i] = 0.f;
        if (c < C) {
          float a = acc[i];
          if (kVec) {
            a = 0.f;
            for (int g = 0; g < vgroups; ++g) a += part[(size_t)g * C + c];     // fixed order: deterministic
          } else {
#pragma unroll
            for (int g = 1; g < kFinGroups; ++g) a += part[(size_t)(g - 1) * C + c];
          }
          const float kh = round_if(i == 0 ? kh_pre : k_hat[(size_t)row * C + c], bf16_mode);
          qh[i] = (i == 0) ? qh_pre : q_hat[(size_t)row * C + c];
          dqh[i] = gs * fmaf(a, o_scale, pm1 * kh);
          dot = fmaf(qh[i], dqh[i], dot);
          if (dk) dk[(size_t)row * C + c] = gs * pm1 * round_if(qh[i], bf16_mode);
        }
      }
      dot = warp_sum(dot);
      if ((ct & 31) == 0) red[ct >> 5] = dot;
      asm volatile("bar.sync 1, 256;" ::: "memory");   // the 8 warps of group 0 only
      dot = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) dot += red[w];
      const float inv = inv_norm[row];
      if (dq) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int c = ct + 256 * i;
          if (c < C) dq[(size_t)row * C + c] = (dqh[i] - qh[i] * dot) * inv;
        }
      }
    }
  }

  // ---- diagnostics of this row (objectives.py:337-349): five dot products over C, then closed forms
  if (diag.out) {
    float a[5] = {0.f, 0.f, 0.f, 0.f, 0.f};   // q.k, |q-k|^2, |k|^2, q.sum_vec, q.sum_unit
    for (int c = tid; c < C; c += kFinThreads) {
      const float qh = q_hat[(size_t)row * C + c], kh = k_hat[(size_t)row * C + c];
      const float df = qh - kh;
      a[0] = fmaf(qh, kh, a[0]);
      a[1] = fmaf(df, df, a[1]);
      a[2] = fmaf(kh, kh, a[2]);
      a[3] = fmaf(qh, diag.sum_vec[c], a[3]);
      a[4] = fmaf(qh, diag.sum_unit[c], a[4]);
    }
#pragma unroll
    for (int i = 0; i < 5; ++i) a[i] = warp_sum(a[i]);
    if ((tid & 31) == 0) {
#pragma unroll
      for (int i = 0; i < 5; ++i) dred[tid >> 5][i] = a[i];
    }
    __syncthreads();
    if (tid == 0) {
      float t[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
      for (int w = 0; w < kFinThreads / 32; ++w)
#pragma unroll
        for (int i = 0; i < 5; ++i) t[i] += dred[w][i];
      const float qn = sqrtf(qn2[row]), kn = sqrtf(t[2]);
      float* o = diag.rows + (size_t)row * kDiagValues;
      o[0] = sqrtf(t[1]);                                                       // |q^ - k^|
      o[1] = t[0] / (fmaxf(qn, diag.cos_eps) * fmaxf(kn, diag.cos_eps));        // cosine(q^, k^)
      o[2] = t[0];                                                              // q^ . k^
      o[3] = s_stats[2] * inv_K;                                                // mean_j |q^ - queue_j|
      o[4] = t[4] * inv_K / fmaxf(qn, diag.cos_eps);                            // mean_j cosine(q^, queue_j)
      o[5] = t[3] * inv_K;                                                      // mean_j q^ . queue_j
    }
  }

  // deterministic loss (and diagnostics) reduction by the last row-CTA
  if (loss || diag.out) {
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(counter, 1u) == (unsigned)B - 1u);
    __syncthreads();
    if (s_last && tid < 256) {
      __threadfence();
      const int n_red = diag.out ? 1 + kDiagValues : 1;
      for (int which = loss ? 0 : 1; which < n_red; ++which) {
        float acc = 0.f;
        for (int r = tid; r < B; r += 256)
          acc += (which == 0) ? __ldcg(row_loss + r) : __ldcg(diag.rows + (size_t)r * kDiagValues + (which - 1));
        // fixed-shape tree: warp shuffle then 8 partials in order
        acc = warp_sum(acc);
        asm volatile("bar.sync 2, 256;" ::: "memory");
        if ((tid & 31) == 0) red[tid >> 5] = acc;
        asm volatile("bar.sync 2, 256;" ::: "memory");
        if (tid == 0) {
          float t = 0.f;
          for (int w = 0; w < 8; ++w) t += red[w];
          if (which == 0) *loss = t * (loss_scale / (float)B);
          else diag.out[which - 1] = t / (float)B;
        }
      }
    }
  }
}

template <typename TQ, typename TKK>
static int launch_prep(const void* q, const void* k, int B, int C, float scale2, bool nk, bool bf16_mode, char* ws,
                       const InfoNcePlan& p, float* k_hat_out, bool want_bf16, cudaStream_t s) {
  const int rows = want_bf16 ? p.b_pad : B;
  infonce_prep_kernel<TQ, TKK><<<rows, 128, 0, s>>>(
      (const TQ*)q, (const TKK*)k, B, C, scale2, nk, bf16_mode, (float*)(ws + p.off_qhat), (float*)(ws + p.off_khat),
      k_hat_out, (float*)(ws + p.off_inv), (float*)(ws + p.off_pos2), (float*)(ws + p.off_qn2),
      want_bf16 ? (__nv_bfloat16*)(ws + p.off_qhat_bf16) : nullptr, p.b_pad, p.split, (unsigned int*)(ws + p.off_counter));
  RMCL_LAUNCH_OK("infonce_prep_kernel");
  return RMCL_OK;
}

}  // namespace rmcl

using namespace rmcl;

// ---- optional stage timing (measurement aid; off by default) --------------------------------
static thread_local bool g_prof_on = false;
static thread_local cudaEvent_t g_prof_ev[4] = {nullptr, nullptr, nullptr, nullptr};
static thread_local bool g_prof_valid = false;

extern "C" int rmcl_profile_enable(int on) {
  if (on && !g_prof_ev[0]) {
    for (int i = 0; i < 4; ++i) RMCL_CUDA_OK(cudaEventCreate(&g_prof_ev[i]));
  }
  g_prof_on = on != 0;
  g_prof_valid = false;
  return RMCL_OK;
}

extern "C" int rmcl_profile_infonce_ms(float* out3) {
  RMCL_CHECK_ARG(out3 != nullptr, "rmcl_profile_infonce_ms: null pointer");
  RMCL_CHECK_ARG(g_prof_valid, "rmcl_profile_infonce_ms: no profiled rmcl_infonce_fwd_bwd call on this thread");
  RMCL_CUDA_OK(cudaEventSynchronize(g_prof_ev[3]));
  for (int i = 0; i < 3; ++i) RMCL_CUDA_OK(cudaEventElapsedTime(out3 + i, g_prof_ev[i], g_prof_ev[i + 1]));
  return RMCL_OK;
}

#define RMCL_PROF_MARK(i)                                        \
  do {                                                           \
    if (g_prof_on) RMCL_CUDA_OK(cudaEventRecord(g_prof_ev[i], s)); \
  } while (0)

static bool tc_alignment_ok(const void* queue, int64_t ldq) {
  return (reinterpret_cast<uintptr_t>(queue) & 15u) == 0 && (ldq % 8 == 0);
}

extern "C" size_t rmcl_infonce_workspace_bytes(int B, int C, int64_t K, rmcl_dtype queue_dtype, int path) {
  if (B <= 0 || C <= 0 || K <= 0) return 0;
  // size for whichever path needs more so that AUTO can pick either at call time
  size_t best = 0;
  for (int pth : {RMCL_INFONCE_SIMT, RMCL_INFONCE_TCGEN05}) {
    if (path != RMCL_INFONCE_AUTO && path != pth) continue;
    if (queue_dtype == RMCL_BF16_HILO && pth == RMCL_INFONCE_SIMT) continue;
    InfoNcePlan p;
    if (infonce_make_plan(B, C, K, queue_dtype, pth, true, &p) == RMCL_OK && p.total > best) best = p.total;
  }
  return best;
}

extern "C" int rmcl_infonce_describe(int B, int C, int64_t K, rmcl_dtype queue_dtype, int path, int need_grad, char* buf,
                                     size_t buf_bytes) {
  RMCL_CHECK_ARG(B > 0 && C > 0 && K > 0 && buf && buf_bytes > 0, "rmcl_infonce_describe: bad arguments");
  InfoNcePlan p;
  int rc = infonce_make_plan(B, C, K, queue_dtype, path, true, &p);
  if (rc != RMCL_OK) return rc;
  const char* names;
  int n;
  if (p.path == RMCL_INFONCE_SIMT) {
    names = "infonce_prep_kernel,infonce_simt_kernel,infonce_finalize_kernel";
    n = 3;
  } else if (p.two_pass && need_grad) {
    names = p.split ? "infonce_prep_kernel,infonce_s_kernel<split>,infonce_pv_kernel<split>,infonce_finalize_kernel"
                    : "infonce_prep_kernel,infonce_s_kernel,infonce_pv_kernel,infonce_finalize_kernel";
    n = 4;
  } else if (p.two_pass) {
    names = "infonce_prep_kernel,infonce_s_kernel,infonce_finalize_kernel";
    n = 3;
  } else {
    names = "infonce_prep_kernel,infonce_tc_kernel,infonce_finalize_kernel";
    n = 3;
  }
  snprintf(buf, buf_bytes, "%s", names);
  return n;
}

static int infonce_impl(const void* q, rmcl_dtype q_dtype, const void* k, rmcl_dtype k_dtype, const void* queue,
                        rmcl_dtype queue_dtype, int B, int C, int64_t K, int64_t ldq, float tau, float loss_scale,
                        unsigned flags, int path, float* loss, float* loss_per_row, float* lse, float* pos,
                        int64_t* argmax, float* dq, float* dk, float* k_hat_out, void* workspace, size_t workspace_bytes,
                        void* stream, InfoNceDiag diag) {
  RMCL_CHECK_ARG(q && k && queue && workspace, "rmcl_infonce_fwd_bwd: null pointer");
  RMCL_CHECK_ARG(B > 0 && C > 0 && K > 0 && K < (1ll << 31) && ldq >= K, "rmcl_infonce_fwd_bwd: bad sizes B=%d C=%d K=%lld ldq=%lld",
                 B, C, (long long)K, (long long)ldq);
  RMCL_CHECK_ARG(tau > 0.f, "rmcl_infonce_fwd_bwd: temperature must be > 0");
  RMCL_CHECK_ARG(dtype_ok(q_dtype) && dtype_ok(k_dtype) && (dtype_ok(queue_dtype) || queue_dtype == RMCL_BF16_HILO),
                 "rmcl_infonce_fwd_bwd: bad dtype");
  RMCL_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 255u) == 0, "rmcl_infonce_fwd_bwd: workspace must be 256B aligned");
  InfoNcePlan p;
  int rc = infonce_make_plan(B, C, K, queue_dtype, path, tc_alignment_ok(queue, ldq), &p);
  if (rc != RMCL_OK) return rc;
  if (workspace_bytes < p.total) {
    set_error("rmcl_infonce_fwd_bwd: workspace %zu < required %zu", workspace_bytes, p.total);
    return RMCL_E_WORKSPACE;
  }
  cudaStream_t s = (cudaStream_t)stream;
  char* ws = (char*)workspace;
  const float scale2 = kLog2e / tau;
  const bool bf16_mode = (queue_dtype == RMCL_BF16);   // hi/lo queues keep fp32 semantics: q^, k^ are not rounded
  const bool nk = (flags & RMCL_INFONCE_NORMALIZE_K) != 0;
  const bool want_grad = (flags & RMCL_INFONCE_NO_GRAD) == 0 && (dq || dk);
  const bool tc = (p.path == RMCL_INFONCE_TCGEN05);
  using bf16 = __nv_bfloat16;
  const bool partial_only = (flags & RMCL_INFONCE_DEBUG_PARTIAL_ONLY) != 0;   // measurement aid, see the header
  RMCL_PROF_MARK(0);
  if (partial_only)
    rc = RMCL_OK;
  else if (q_dtype == RMCL_F32 && k_dtype == RMCL_F32)
    rc = launch_prep<float, float>(q, k, B, C, scale2, nk, bf16_mode, ws, p, k_hat_out, tc, s);
  else if (q_dtype == RMCL_F32)
    rc = launch_prep<float, bf16>(q, k, B, C, scale2, nk, bf16_mode, ws, p, k_hat_out, tc, s);
  else if (k_dtype == RMCL_F32)
    rc = launch_prep<bf16, float>(q, k, B, C, scale2, nk, bf16_mode, ws, p, k_hat_out, tc, s);
  else
    rc = launch_prep<bf16, bf16>(q, k, B, C, scale2, nk, bf16_mode, ws, p, k_hat_out, tc, s);
  if (rc != RMCL_OK) return rc;
  RMCL_PROF_MARK(1);

  diag.rows = diag.out ? (float*)(ws + p.off_diagrows) : nullptr;
  InfoNcePartials parts{(float*)(ws + p.off_m), (float*)(ws + p.off_l), (float*)(ws + p.off_av), (int*)(ws + p.off_ai),
                        (float*)(ws + p.off_o), diag.out ? diag.colnorm2 : nullptr, (const float*)(ws + p.off_qn2),
                        (float*)(ws + p.off_pdist)};
  if (tc && p.two_pass)
    rc = infonce_tc2_launch((const bf16*)(ws + p.off_qhat_bf16), queue, B, C, K, ldq, scale2, p, parts,
                            (bf16*)(ws + p.off_ptilde), p.k_pad, (unsigned int*)(ws + p.off_counter) + 1, argmax != nullptr,
                            (want_grad || partial_only) ? 1 : 0, p.split, s);
  else if (tc)
    rc = infonce_tc_launch((const bf16*)(ws + p.off_qhat_bf16), queue, B, C, K, ldq, scale2, p, parts, argmax != nullptr,
                           (want_grad || partial_only) ? 1 : 0, s);
  else
    rc = infonce_simt_launch((const float*)(ws + p.off_qhat), queue, queue_dtype, B, C, K, ldq, scale2, p, parts, want_grad, s);
  if (rc != RMCL_OK) return rc;
  RMCL_PROF_MARK(2);
  if (partial_only) return RMCL_OK;

  // merge weights + per-group partial rows: (groups-1) rows of the scalar path, kFinThreads/(C/8) rows of the 16-byte path
  const bool bf16_partials = tc && !p.split;   // the split-operand path keeps its partials in fp32
  const size_t fin_rows = bf16_partials ? (size_t)(kFinThreads / (C / 8)) : (size_t)(kFinGroups - 1);
  const size_t fin_smem = ((size_t)((p.splits + 3) & ~3) + fin_rows * C) * sizeof(float);
  if (bf16_partials) {
    RMCL_CUDA_OK(launch_pdl(infonce_finalize_kernel<__nv_bfloat16>, dim3(B), dim3(kFinThreads), fin_smem, s,
        B, C, p.splits, 1.f / tau, loss_scale / (float)B, loss_scale, bf16_mode, want_grad, (const float*)(ws + p.off_qhat),
        (const float*)(ws + p.off_khat), (const float*)(ws + p.off_inv), (const float*)(ws + p.off_pos2), parts.m, parts.l,
        parts.av, parts.ai, reinterpret_cast<const __nv_bfloat16*>(parts.o), (float*)(ws + p.off_rowloss),
        (unsigned int*)(ws + p.off_counter), loss, loss_per_row, lse, pos, reinterpret_cast<long long*>(argmax), dq, dk,
        (const float*)parts.dist, parts.qn2, 1.f / (float)K, diag));
  } else {
    RMCL_CUDA_OK(launch_pdl(infonce_finalize_kernel<float>, dim3(B), dim3(kFinThreads), fin_smem, s,
        B, C, p.splits, 1.f / tau, loss_scale / (float)B, loss_scale, bf16_mode, want_grad, (const float*)(ws + p.off_qhat),
        (const float*)(ws + p.off_khat), (const float*)(ws + p.off_inv), (const float*)(ws + p.off_pos2), parts.m, parts.l,
        parts.av, parts.ai, (const float*)parts.o, (float*)(ws + p.off_rowloss), (unsigned int*)(ws + p.off_counter), loss,
        loss_per_row, lse, pos, reinterpret_cast<long long*>(argmax), dq, dk, (const float*)parts.dist, parts.qn2,
        1.f / (float)K, diag));
  }
  RMCL_PROF_MARK(3);
  if (g_prof_on) g_prof_valid = true;
  return RMCL_OK;
}

extern "C" int rmcl_infonce_fwd_bwd(const void* q, rmcl_dtype q_dtype, const void* k, rmcl_dtype k_dtype,
                                    const void* queue, rmcl_dtype queue_dtype, int B, int C, int64_t K, int64_t ldq,
                                    float tau, float loss_scale, unsigned flags, int path, float* loss,
                                    float* loss_per_row, float* lse, float* pos, int64_t* argmax, float* dq, float* dk,
                                    float* k_hat_out, void* workspace, size_t workspace_bytes, void* stream) {
  InfoNceDiag none{nullptr, nullptr, nullptr, 0.f, nullptr, nullptr};
  return infonce_impl(q, q_dtype, k, k_dtype, queue, queue_dtype, B, C, K, ldq, tau, loss_scale, flags, path, loss,
                      loss_per_row, lse, pos, argmax, dq, dk, k_hat_out, workspace, workspace_bytes, stream, none);
}

extern "C" int rmcl_infonce_fwd_bwd_diag(const void* q, rmcl_dtype q_dtype, const void* k, rmcl_dtype k_dtype,
                                         const void* queue, rmcl_dtype queue_dtype, int B, int C, int64_t K, int64_t ldq,
                                         float tau, float loss_scale, unsigned flags, int path, float* loss,
                                         float* loss_per_row, float* lse, float* pos, int64_t* argmax, float* dq,
                                         float* dk, float* k_hat_out, const float* colnorm2, const float* sum_vec,
                                         const float* sum_unit, float cos_eps, float* diag_out, void* workspace,
                                         size_t workspace_bytes, void* stream) {
  RMCL_CHECK_ARG(colnorm2 && sum_vec && sum_unit && diag_out, "rmcl_infonce_fwd_bwd_diag: null diagnostics pointer");
  RMCL_CHECK_ARG((reinterpret_cast<uintptr_t>(colnorm2) & 15u) == 0, "rmcl_infonce_fwd_bwd_diag: colnorm2 must be 16B aligned");
  InfoNceDiag dg{colnorm2, sum_vec, sum_unit, cos_eps, diag_out, nullptr};
  return infonce_impl(q, q_dtype, k, k_dtype, queue, queue_dtype, B, C, K, ldq, tau, loss_scale, flags, path, loss,
                      loss_per_row, lse, pos, argmax, dq, dk, k_hat_out, workspace, workspace_bytes, stream, dg);
}
