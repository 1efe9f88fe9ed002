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
#include <stdlib.h>

#include "infonce.cuh"
#include "infonce_rows.cuh"

namespace rmcl {

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

// ------------------------------------------------------------------------------------ prep
// RMCL_PREP_WARP_ROWS: one WARP per row (shuffles only, a lane keeps its C/32 elements of q and k in flight together) instead
// of a 128-thread team per row (8 named barriers on the row's critical path).  The whole call waits for this kernel — the
// flash pass sits in griddepcontrol.wait until it has finished —, so its latency, not its throughput, is what counts.
// Measured (profiles/r2_tc_experiments.txt): SLOWER — whole call 36.5 us against 33.2 us at cfg2: a lane then runs 8 elements'
// worth of IEEE divisions and stores in a row where a team thread runs 2.  Off.
#ifndef RMCL_PREP_WARP_ROWS
#define RMCL_PREP_WARP_ROWS 0
#endif
constexpr int kPrepWarps = 2;   // rows per CTA in the warp-per-row variant (B = 256: 128 CTAs, under one per SM)

template <typename TQ, typename TKK>
__global__ void __launch_bounds__(128) infonce_prep_kernel(const PrepArgs a, unsigned int* __restrict__ counter, int rows) {
  __shared__ float red[4];
  if (threadIdx.x == 0) pdl_trigger();   // the partial kernel may set itself up while this one runs
  pdl_wait();                            // itself launched ahead of the end of the kernel in front of it (q, k come from there)
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    counter[0] = 0u;
    counter[1] = 0u;   // overflow flag of the two-pass tcgen05 variant
  }
#if RMCL_PREP_WARP_ROWS
  const int row = blockIdx.x * kPrepWarps + (threadIdx.x >> 5);
  if (row < rows) prep_row_warp<TQ, TKK>(a, row, threadIdx.x & 31);
#else
  prep_row<TQ, TKK>(a, blockIdx.x, red, threadIdx.x, kBarPrep);
#endif
}

// -------------------------------------------------------------------------------- finalize
constexpr int kFinThreads = 512;              // two CTAs per SM, so B=256 rows are one wave on 148 SMs

// TP: element type of the partial accumulators (fp32 from the SIMT / split-operand kernels, bf16 from the bf16 tcgen05 kernels)
template <typename TP>
__global__ void __launch_bounds__(kFinThreads, 2) infonce_finalize_kernel(const FinArgs a) {
  extern __shared__ float fin_smem[];
  __shared__ FinShared fs;
#if RMCL_PDL_EARLY_TRIGGER & 4
  if (threadIdx.x == 0) pdl_trigger();        // the enqueue behind this kernel may queue up
#endif
  pdl_wait();                                 // launched early (PDL): the partials must be complete
  finalize_row<TP, kFinThreads>(a, blockIdx.x, fin_smem, &fs, threadIdx.x, kBarFin);
}

static PrepArgs make_prep_args(const void* q, const void* k, int B, int C, float scale2, bool nk, bool bf16_mode, char* ws,
                               const InfoNcePlan& p, float* k_hat_out, bool want_bf16) {
  PrepArgs a;
  a.q = q;
  a.k = k;
  a.B = B;
  a.C = C;
  a.scale2 = scale2;
  a.normalize_k = nk;
  a.bf16_mode = bf16_mode;
  a.q_hat = (float*)(ws + p.off_qhat);
  a.k_hat = (float*)(ws + p.off_khat);
  a.k_hat_out = k_hat_out;
  a.inv_norm = (float*)(ws + p.off_inv);
  a.pos2 = (float*)(ws + p.off_pos2);
  a.qn2 = (float*)(ws + p.off_qn2);
  a.q_hat_bf16 = want_bf16 ? (__nv_bfloat16*)(ws + p.off_qhat_bf16) : nullptr;
  a.b_pad = p.b_pad;
  a.split = p.split;
  return a;
}

template <typename TQ, typename TKK>
static int launch_prep(const PrepArgs& a, char* ws, const InfoNcePlan& p, cudaStream_t s) {
  const int rows = a.q_hat_bf16 ? p.b_pad : a.B;
#if RMCL_PREP_WARP_ROWS
  infonce_prep_kernel<TQ, TKK><<<(rows + kPrepWarps - 1) / kPrepWarps, 32 * kPrepWarps, 0, s>>>(a, (unsigned int*)(ws + p.off_counter), rows);
#else
  RMCL_CUDA_OK(launch_pdl(infonce_prep_kernel<TQ, TKK>, dim3(rows), dim3(128), 0, s, a, (unsigned int*)(ws + p.off_counter), rows));
#endif
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

// RMCL_B200_INFONCE_FUSED=1 selects the single-launch cooperative kernel (prep rows | flash pass | finalize rows with grid
// barriers in between) for bf16 queues with C <= 256.  It is correct (the whole GPU suite passes with it) but measured
// SLOWER than the three-launch chain under programmatic dependent launch — 39.5 vs 34.5 us per call at cfg2 by CUDA-graph
// replay, 30.8 vs 25.6 us at the cfg4 shape (profiles/r2_ab_fused.txt) — so the chain stays the default.  The in-kernel
// timeline says why (profiles/r2_timeline_fused.txt): a CTA owns one or two rows of each row phase and runs that code once,
// with a cold instruction cache (~7 cycles per instruction the first time through) and every dependent L2 round trip on the
// critical path (prep rows + grid barrier 9 us, finalize rows 10-12 us), while the separate kernels spread the same rows over
// 256 CTAs whose latencies overlap and whose launches are hidden by PDL.
static bool infonce_fused_enabled() {
  static const bool on = [] {
    const char* e = getenv("RMCL_B200_INFONCE_FUSED");
    return e && e[0] == '1';
  }();
  return on;
}

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
  } else if (infonce_fused_enabled() && p.splits * p.row_blocks <= sm_count()) {
    names = "infonce_fused_kernel";
    n = 1;
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
  const PrepArgs pa = make_prep_args(q, k, B, C, scale2, nk, bf16_mode, ws, p, k_hat_out, tc);
  diag.rows = diag.out ? (float*)(ws + p.off_diagrows) : nullptr;
  InfoNcePartials parts{(float*)(ws + p.off_m), (float*)(ws + p.off_l), (float*)(ws + p.off_av), (int*)(ws + p.off_ai),
                        (float*)(ws + p.off_o), diag.out ? diag.colnorm2 : nullptr, (const float*)(ws + p.off_qn2),
                        (float*)(ws + p.off_pdist)};
  FinArgs fa;
  fa.B = B;
  fa.C = C;
  fa.splits = p.splits;
  fa.inv_tau = 1.f / tau;
  fa.grad_scale = loss_scale / (float)B;
  fa.loss_scale = loss_scale;
  fa.bf16_mode = bf16_mode;
  fa.want_grad = want_grad;
  fa.q_hat = pa.q_hat;
  fa.k_hat = pa.k_hat;
  fa.inv_norm = pa.inv_norm;
  fa.pos2 = pa.pos2;
  fa.pm = parts.m;
  fa.pl = parts.l;
  fa.pav = parts.av;
  fa.pai = parts.ai;
  fa.po = parts.o;
  fa.row_loss = (float*)(ws + p.off_rowloss);
  fa.counter = (unsigned int*)(ws + p.off_counter);
  fa.loss = loss;
  fa.loss_per_row = loss_per_row;
  fa.lse_out = lse;
  fa.pos_out = pos;
  fa.argmax_out = reinterpret_cast<long long*>(argmax);
  fa.dq = dq;
  fa.dk = dk;
  fa.pdist = parts.dist;
  fa.qn2 = parts.qn2;
  fa.inv_K = 1.f / (float)K;
  fa.diag = diag;

  // ---- single launch: prep rows, the tcgen05 flash pass and the finalize rows as three phases of ONE cooperative kernel
  //      (grid-wide barriers between them) whenever the whole grid is co-resident
  RMCL_PROF_MARK(0);
  if (tc && !p.two_pass && !partial_only && infonce_fused_enabled() && p.splits * p.row_blocks <= sm_count()) {
    RMCL_PROF_MARK(1);
    rc = infonce_tc_fused_launch(pa, q_dtype == RMCL_BF16, k_dtype == RMCL_BF16, fa, queue, B, C, K, ldq, scale2, p, parts,
                                 argmax != nullptr, want_grad ? 1 : 0, s);
    if (rc != RMCL_OK) return rc;
    RMCL_PROF_MARK(2);
    RMCL_PROF_MARK(3);
    if (g_prof_on) g_prof_valid = true;
    return RMCL_OK;
  }
  if (partial_only)
    rc = RMCL_OK;
  else if (q_dtype == RMCL_F32 && k_dtype == RMCL_F32)
    rc = launch_prep<float, float>(pa, ws, p, s);
  else if (q_dtype == RMCL_F32)
    rc = launch_prep<float, bf16>(pa, ws, p, s);
  else if (k_dtype == RMCL_F32)
    rc = launch_prep<bf16, float>(pa, ws, p, s);
  else
    rc = launch_prep<bf16, bf16>(pa, ws, p, s);
  if (rc != RMCL_OK) return rc;
  RMCL_PROF_MARK(1);

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

  const bool bf16_partials = tc && !p.split;   // the split-operand path keeps its partials in fp32
  const size_t fin_smem = finalize_smem_bytes(C, p.splits, bf16_partials, kFinThreads);
  if (bf16_partials)
    RMCL_CUDA_OK(launch_pdl(infonce_finalize_kernel<__nv_bfloat16>, dim3(B), dim3(kFinThreads), fin_smem, s, fa));
  else
    RMCL_CUDA_OK(launch_pdl(infonce_finalize_kernel<float>, dim3(B), dim3(kFinThreads), fin_smem, s, fa));
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
