// Internal interface between the InfoNCE stages (prep -> partial -> finalize).
//
// The fused InfoNCE is a split-K "flash" computation over the queue axis.  Every CTA of the
// partial stage owns (a block of rows) x (a contiguous range of queue columns) and emits, per row,
//     m  = max_j s_j                (s_j = q^.queue_j * log2(e)/tau, i.e. logits in log2 units)
//     l  = sum_j 2^(s_j - m)
//     O  = sum_j 2^(s_j - m) * queue_j          [C]
//     (av, ai) = (max value, first index attaining it)   for the row argmax
// The finalize stage merges the splits and the positive logit into lse / loss / dq / dk.
// Two partial-stage implementations write this same format:
//   infonce_simt.cu  — fp32 CUDA-core kernel, any dtype/shape (fp32 parity path, C<=1024)
//   infonce_tc.cu    — tcgen05/TMEM/TMA kernel for bf16 queues (C in {64,128,256})
//   infonce_tc2.cu   — two tcgen05 kernels (S pass, PV pass) for bf16 queues with C in {512,768}
#pragma once
#include "common.cuh"

namespace rmcl {

// The bf16 copy of Q^ is the one operand every CTA of the tcgen05 kernel reads in full at the same
// moment (148 CTAs x 64 KB out of a 128 KB footprint at cfg2).  Round 1 wrote kQhatReplicas = 8 copies (CTA `split` reads
// copy split % kQhatReplicas) to spread that over more L2 lines.  Measured again in round 2 with 1 / 2 / 8 / 32 / 74 copies
// (profiles/r2_tc_experiments.txt): the flash pass does not change (20.5-20.6 us at cfg2 for every count — what delayed the
// fetch was the ring prefetch in front of it, RMCL_TC_TILES_BEFORE_Q), and every extra copy is written by the prep kernel on
// the critical path of the call (whole call 33.1 us with 1 copy, 33.4 with 8, 39.3 with 74).  One copy.
#ifndef RMCL_QHAT_REPLICAS
#define RMCL_QHAT_REPLICAS 1
#endif
constexpr int kQhatReplicas = RMCL_QHAT_REPLICAS;

struct InfoNcePlan {
  int path;             // RMCL_INFONCE_SIMT / RMCL_INFONCE_TCGEN05 (never AUTO)
  int splits;           // number of queue-axis splits
  long long cols_per_split;
  int rows_per_cta;     // 16 (simt) / 128 (tc)
  int row_blocks;
  int tile_cols;        // queue columns per inner tile
  int b_pad;            // B rounded up to rows_per_cta
  bool two_pass;        // tcgen05 with C > 256, or the split-operand path: S pass + PV pass through P~ (infonce_tc2.cu)
  bool split;           // fp32-accurate split-operand path: queue is RMCL_BF16_HILO, Q^ and P~ are hi/lo pairs, fp32 partials
  long long k_pad;      // K rounded up to 64: row stride of P~
  // workspace carve-up (byte offsets)
  size_t off_qhat, off_khat, off_inv, off_pos2, off_qhat_bf16, off_m, off_l, off_av, off_ai, off_o, off_rowloss,
      off_counter, off_qn2, off_pdist, off_diagrows, off_ptilde, total;
};

// Optional: the reference's per-view diagnostics (vilt/modules/objectives.py:337-349) from the same
// pass.  dot and cosine against the queue are linear in the queue, so they reduce to two [C] vectors
// (rmcl_queue_stats); the mean L2 distance is not, and is accumulated from S inside the partial
// kernels as sum_j sqrt(|q^|^2 - 2 q^.queue_j + |queue_j|^2).
struct InfoNceDiag {
  const float* colnorm2;   // [K]  |queue_j|^2
  const float* sum_vec;    // [C]  sum_j queue[:, j]
  const float* sum_unit;   // [C]  sum_j queue[:, j] / max(|queue_j|, cos_eps)
  float cos_eps;
  float* out;              // [6]  means over rows: pos_dist, pos_cosine, pos_dot, neg_dist, neg_cosine, neg_dot
  float* rows;             // [B][6] per-row values (workspace)
};
constexpr int kDiagValues = 6;

// Fills plan; returns RMCL_OK or an error (unsupported shape for a forced path).
int infonce_make_plan(int B, int C, long long K, int queue_dtype, int path, bool aligned_for_tc, InfoNcePlan* plan);

struct InfoNcePartials {
  float* m;         // [B][splits]  (a row's statistics are contiguous: one coalesced load in finalize)
  float* l;         // [B][splits]
  float* av;        // [B][splits]
  int* ai;          // [B][splits]
  float* o;         // [splits][B][C]  fp32 (SIMT partial kernel) or bf16 (tcgen05 partial kernel) elements
  // diagnostics (n2 == nullptr: off)
  const float* n2;  // [K] squared column norms of the queue
  const float* qn2; // [B] |q^|^2
  float* dist;      // [B][splits] sum over the split's columns of |q^ - queue_j|
};

// partial stages
int infonce_simt_launch(const float* q_hat, const void* queue, int queue_dtype, int B, int C, long long K,
                        long long ldq, float scale2, const InfoNcePlan& plan, InfoNcePartials out, bool want_o,
                        cudaStream_t s);
int infonce_tc_launch(const __nv_bfloat16* q_hat_bf16, const void* queue, int B, int C, long long K, long long ldq,
                      float scale2, const InfoNcePlan& plan, InfoNcePartials out, int want_argmax, int want_o,
                      cudaStream_t s);
struct PrepArgs;
struct FinArgs;
// single launch: prep rows -> flash pass -> finalize rows in one cooperative kernel (grid <= SM count)
int infonce_tc_fused_launch(const PrepArgs& prep, bool q_bf16, bool k_bf16, const FinArgs& fin, const void* queue, int B, int C,
                            long long K, long long ldq, float scale2, const InfoNcePlan& plan, InfoNcePartials out,
                            int want_argmax, int want_o, cudaStream_t s);
int infonce_tc_tile_cols(int C);
bool infonce_tc2_supports(int C);
int infonce_tc2_launch(const __nv_bfloat16* q_hat_bf16, const void* queue, int B, int C, long long K, long long ldq,
                       float scale2, const InfoNcePlan& plan, InfoNcePartials out, __nv_bfloat16* ptilde, long long k_pad,
                       unsigned int* overflow_flag, int want_argmax, int want_o, bool split, cudaStream_t s);
bool infonce_tc2_split_supports(int C);
bool infonce_tc_built();

}  // namespace rmcl
