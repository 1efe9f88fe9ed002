/* rmcl_b200.h — C-ABI of librmcl_b200.so: the B200 (sm_100a) kernels behind the RMCL
 * contrastive-adversarial training step.
 *
 * The reference (stanFurrer/Robust-Multimodal-Contrastive-Learning) has no FFI of its own:
 * the hot path is a chain of ATen expressions inside two Python functions.  Every entry point
 * below replaces one such chain; the `replaces:` line cites it (paths relative to the
 * reference root).  INTEGRATION.md shows the ctypes binding a reference maintainer adds.
 *
 * Conventions (all entry points)
 *   - return value: RMCL_OK (0) or a negative rmcl_status; rmcl_last_error() gives the text of
 *     the most recent failure on the calling thread.
 *   - every pointer named *_dev / every tensor argument is DEVICE memory owned by the caller.
 *     Nothing is allocated, freed, or retained; nothing synchronises the device or the host.
 *     Work is enqueued on `stream` (a cudaStream_t passed as void*; NULL = legacy default).
 *   - there is no CPU path.  On a machine without an sm_100 device every launch fails with
 *     RMCL_E_CUDA.
 *   - thread-compatible: no global mutable state besides the per-thread error string.
 */
#ifndef RMCL_B200_H
#define RMCL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  RMCL_OK = 0,
  RMCL_E_BADARG = -1,          /* null pointer, non-positive size, K % B != 0 for enqueue, ... */
  RMCL_E_ALIGN = -2,           /* pointer / stride alignment the kernel needs is not met       */
  RMCL_E_UNSUPPORTED_DIM = -3, /* shape outside what the kernels are built for                 */
  RMCL_E_CUDA = -4,            /* a CUDA runtime/driver call failed (see rmcl_last_error)      */
  RMCL_E_WORKSPACE = -5        /* workspace smaller than rmcl_infonce_workspace_bytes()        */
} rmcl_status;

/* RMCL_BF16_HILO is a QUEUE layout only (rmcl_infonce_fwd_bwd*, written by rmcl_queue_split / rmcl_enqueue_shadow):
 * an fp32 [C,K] queue held as two bf16 planes in one [2C,K] buffer — rows [0,C) hi = bf16(x), rows [C,2C)
 * lo = bf16(x - hi) — i.e. 16 mantissa bits per element.  The InfoNCE kernels contract it as
 * q_hi.Q_hi + q_hi.Q_lo + q_lo.Q_hi on the bf16 tensor cores: fp32-accurate logits and gradients
 * (the reference runs the PGD inner loss in fp32, attack/pgd_attack_vilt.py:141, on its fp32 queue buffer). */
typedef enum { RMCL_F32 = 0, RMCL_BF16 = 1, RMCL_BF16_HILO = 2 } rmcl_dtype;

typedef enum {
  RMCL_PGD_REF_LINF = 0,  /* reference rule: delta += lr*g/max(|g|_inf,1e-8); clamp(+-eps) if eps>0 */
  RMCL_PGD_SIGN_LINF = 1, /* delta += lr*sign(g); clamp(+-eps) if eps>0                             */
  RMCL_PGD_L2 = 2         /* delta += lr*g/max(|g|_2,1e-8); project onto the eps-ball (L2)          */
} rmcl_pgd_mode;

/* InfoNCE code paths.  AUTO picks TCGEN05 when the shape/dtype allows it (bf16 queue with
 * C in {64,128,256,512,768}, or a bf16 hi/lo queue with C in {64,128,256}; K % 8 == 0, 16-byte aligned
 * queue) and SIMT otherwise (fp32 queues: CUDA-core kernel, exact fp32 products).  All are device
 * kernels in this library; forcing TCGEN05 on an unsupported shape is RMCL_E_UNSUPPORTED_DIM. */
typedef enum { RMCL_INFONCE_AUTO = 0, RMCL_INFONCE_SIMT = 1, RMCL_INFONCE_TCGEN05 = 2 } rmcl_infonce_path;

/* flags for rmcl_infonce_fwd_bwd */
#define RMCL_INFONCE_NORMALIZE_K 1u /* k is a raw projection: L2-normalise it too (objectives.py:265) */
#define RMCL_INFONCE_NO_GRAD 2u     /* forward only (clean-query call, objectives.py:269-275)        */
/* Measurement aid: launch only the split-K partial kernel, on the q^/k^ a previous full call with the
 * same shapes left in `workspace` (no prep, no finalize, outputs untouched).  bench.py uses it to time
 * the dominant InfoNCE kernel over consecutive launches. */
#define RMCL_INFONCE_DEBUG_PARTIAL_ONLY 4u

const char* rmcl_last_error(void);
int rmcl_version(void);      /* 1000*major + minor */
int rmcl_sm_count(void);     /* SMs of the current device, or a negative rmcl_status */

/* ---------------------------------------------------------------------------------------------
 * Momentum (EMA) update of the key encoder:  k <- k*m + q*(1-m)  over a list of tensor pairs.
 * replaces: vilt/modules/objectives.py:219-224 (called x4 at 257-260) — 161 tensors,
 *           ~483 elementwise launches; MoCo/MoCo_RMCL.py:65-72 (_momentum_update_key_encoder).
 * Arithmetic is the reference's: round(round(k*mf) + round(q*omf)) in the tensor dtype with
 * mf=(float)m and omf=(float)(1.0-m) (1.0-m evaluated in double, like Python), not an FMA.
 *
 * Planning is a host-only helper: it cuts the tensor list into chunks of at most chunk_elems
 * elements so that one launch balances 161 very unequal tensors over all SMs.  The caller uploads
 * the chunk table to the device once (parameter storage does not move between steps) and reuses it.
 */
typedef struct {
  void* k;          /* device pointer into a key-encoder parameter   */
  const void* q;    /* device pointer into the matching query param  */
  uint64_t n;       /* elements in this chunk                         */
} rmcl_ema_chunk;

/* Returns the number of chunks needed (>=0) or a negative status.  If chunks_out is non-NULL it
 * must have room for that many entries (call once with NULL to size it). */
int64_t rmcl_ema_plan(const void* const* k_ptrs, const void* const* q_ptrs, const uint64_t* numels,
                      int n_tensors, rmcl_dtype dtype, uint64_t chunk_elems, rmcl_ema_chunk* chunks_out);

int rmcl_ema_multi(const rmcl_ema_chunk* chunks_dev, int64_t n_chunks, double m, rmcl_dtype dtype,
                   void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fused InfoNCE forward + backward of queries against [key ; queue], logits never materialised.
 * replaces: objectives.py:269-274 (clean), 326-334+351 (image-attacked), 289-297+314, 364-372+389;
 *           attack/pgd_attack_vilt.py:147,152-158 (PGD inner loss); the autograd backward of each;
 *           the queue.clone() at objectives.py:270.  MoCo/MoCo_RMCL.py:150-164.
 *
 *   q        [B,C]   raw projection-head output (normalised inside: q/max(|q|,1e-12))
 *   k        [B,C]   key; already normalised unless RMCL_INFONCE_NORMALIZE_K
 *   queue    [C,K]   row stride ldq elements (reference layout: K contiguous, vilt_module.py:92)
 *   tau              temperature;  loss = loss_scale * mean_i( lse_i - pos_i )
 *   outputs (any may be NULL except workspace):
 *     loss          f32[1]    scaled mean loss
 *     loss_per_row  f32[B]    lse_i - pos_i (unscaled)
 *     lse           f32[B]    log-sum-exp of row i over the K+1 logits
 *     pos           f32[B]    positive logit  q^.k^/tau
 *     argmax        i64[B]    argmax over the K+1 logits (0 = the positive; objectives.py:275,336)
 *     dq            f32[B,C]  d loss / d q  (through the normalisation)
 *     dk            f32[B,C]  d loss / d k^ (the reference keeps k under no_grad; provided for
 *                              symmetric / MoCo-v3 style callers)
 *     k_hat_out     f32[B,C]  normalised key (what gets enqueued)
 *
 *   workspace  rmcl_infonce_workspace_bytes() bytes, 256-byte aligned, caller-owned.  It must be ZERO-FILLED once before
 *              its first use (it holds the arrival words of the grid-wide barriers of the single-launch kernel; every
 *              call leaves them zeroed again).  Calls sharing a workspace must be stream-ordered.
 *   launches   a chain of 3-4 kernels under programmatic dependent launch (prep -> partial pass(es) -> finalize;
 *              rmcl_infonce_describe names them).  With RMCL_B200_INFONCE_FUSED=1 in the environment a bf16 queue with
 *              C in {64,128,256} and splits x row blocks <= SM count takes ONE cooperative kernel instead (prep rows |
 *              flash pass | finalize rows, grid barriers in between): same results, measured ~15 % slower.
 */
size_t rmcl_infonce_workspace_bytes(int B, int C, int64_t K, rmcl_dtype queue_dtype, int path);

/* Writes the comma-separated names of the kernels one rmcl_infonce_fwd_bwd call with these arguments launches, in
 * launch order, into buf (NUL-terminated, truncated to buf_bytes); returns the number of launches or a negative
 * status.  Host-only (no device work): bench.py reports it as `launches_per_step`. */
int rmcl_infonce_describe(int B, int C, int64_t K, rmcl_dtype queue_dtype, int path, int need_grad, char* buf,
                          size_t buf_bytes);

int rmcl_infonce_fwd_bwd(const void* q, rmcl_dtype q_dtype, const void* k, rmcl_dtype k_dtype,
                         const void* queue, rmcl_dtype queue_dtype, int B, int C, int64_t K,
                         int64_t ldq, float tau, float loss_scale, unsigned flags, int path,
                         float* loss, float* loss_per_row, float* lse, float* pos, int64_t* argmax,
                         float* dq, float* dk, float* k_hat_out, void* workspace,
                         size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Per-view diagnostics of compute_moco_contrastive from the same fused pass.
 * replaces: objectives.py:337-349 (and 300-312, 375-387): pos/neg L2 distance, cosine and dot means —
 *           in the reference a Python loop over the B samples, each iteration reducing the whole
 *           [K,C] queue three times.
 * rmcl_queue_stats reduces the queue once per step (it changes by B columns per step):
 *     colnorm2 f32[K]  |queue_j|^2          sum_vec f32[C]  sum_j queue[:,j]
 *     sum_unit f32[C]  sum_j queue[:,j] / max(|queue_j|, cos_eps)      (cos_eps = 1e-6 in the reference)
 * rmcl_infonce_fwd_bwd_diag is rmcl_infonce_fwd_bwd plus
 *     diag_out f32[6]  means over the B rows of
 *        [0] |q^-k^|_2   [1] cos(q^,k^)   [2] q^.k^   [3] mean_j |q^-queue_j|_2   [4] mean_j cos(q^,queue_j)
 *        [5] mean_j q^.queue_j
 *   (the reference's pos_dist, pos_cosine, pos_dot, neg_dist, neg_cosine, neg_dot of one view).
 *   k must be the normalised key here (the reference passes the normalised k).
 */
int rmcl_queue_stats(const void* queue, rmcl_dtype queue_dtype, int C, int64_t K, int64_t ldq,
                     float cos_eps, float* colnorm2, float* sum_vec, float* sum_unit, void* stream);

int rmcl_infonce_fwd_bwd_diag(const void* q, rmcl_dtype q_dtype, const void* k, rmcl_dtype k_dtype,
                              const void* queue, rmcl_dtype queue_dtype, int B, int C, int64_t K,
                              int64_t ldq, float tau, float loss_scale, unsigned flags, int path,
                              float* loss, float* loss_per_row, float* lse, float* pos, int64_t* argmax,
                              float* dq, float* dk, float* k_hat_out, const float* colnorm2,
                              const float* sum_vec, const float* sum_unit, float cos_eps,
                              float* diag_out, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fused key exchange + enqueue over NVLink peer memory (one node, one launch per rank).
 * replaces: _concat_all_gather (objectives.py:226-235) followed by _dequeue_and_enqueue (objectives.py:238-248),
 *           i.e. the pair ncclAllGather -> rmcl_enqueue.
 * stage_ptrs_dev  device array [world] of pointers: every rank's staging buffer f32[2][world*B_local][C], peer mapped
 *                 (e.g. torch.distributed._symmetric_memory: handle.buffer_ptrs_dev)
 * flag_ptrs_dev   device array [world] of pointers: every rank's flag words u32[>=3], zero before the first call
 *                 ([0] arrivals, [1] local CTA counter, [2] completed calls = the epoch that selects the staging slot;
 *                 all kept on the device, so the launch can be captured in a CUDA graph)
 * keys_local      f32[B_local][C], the caller's normalised keys
 * Every rank pushes its keys into every rank's staging slot, signals, waits until all keys of the step have arrived
 * and then performs the identical enqueue into its own replica (queue / optional bf16 shadow / pointer as rmcl_enqueue).
 * All ranks must make the same sequence of calls.
 */
int rmcl_gather_enqueue_p2p(void* const* stage_ptrs_dev, void* const* flag_ptrs_dev, const void* keys_local,
                            void* queue, rmcl_dtype queue_dtype, void* shadow_bf16, int64_t lds, int shadow_planes,
                            int64_t* ptr_dev, int rank, int world, int B_local, int C, int64_t K,
                            int64_t ldq, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fused Barlow-Twins cross-correlation loss, forward + backward.
 * replaces: vilt/modules/objectives.py:480-486 (text view), 506-512 (image view), 533-539 (both) and the
 *           attacker's inner loss attack/pgd_attack_vilt.py:219-224:
 *               c = q.T @ k; c.div_(bs); all_reduce(c)
 *               on_diag = diagonal(c).add_(-1).pow_(2).sum(); off_diag = off_diagonal(c).pow_(2).sum()
 *               loss = on_diag + lambda * off_diag            (+ autograd backward to q)
 *           The D x D matrix c (8192 x 8192 fp32 = 268 MB in the reference, all-reduced across ranks)
 *           is never materialised: bf16 tcgen05 GEMMs with fp32 accumulation, tile by tile.
 * q, k      [Bg, D] row-major (fp32 or bf16): the batch gathered over all ranks in rank order (Bg <= 256);
 *           a single-GPU call passes its own batch, b0 = 0, Bl = Bg.
 * b0, Bl    rows of q owned by the caller: dq is [Bl, D] and holds d(loss)/d(q[b0:b0+Bl]).
 * inv_bs    1 / per_step_bs (objectives.py:481);  lambda = pl_module.adv_lr (objectives.py:486).
 * w_on,w_off weights of the two sums in the gradient: dq = d(w_on*on_diag + w_off*off_diag)/dq * loss_scale.
 *           The reference's loss is w_on = 1, w_off = lambda.
 * outputs   on_diag, off_diag f32[1] (unweighted sums, as the reference logs them), loss f32[1] =
 *           loss_scale * (on_diag + lambda * off_diag), dq f32[Bl, D], cdiag f32[D] = diagonal(c);
 *           any may be NULL.  (cdiag lets a caller whose upstream gradients of the two sums differ
 *           combine dq = g_on * 2/bs (cdiag-1) k + g_off * dq(w_on=0, w_off=1) without a second pass.)
 * path      RMCL_BARLOW_GRAM: through the Gram matrices q q^T, k k^T — sum_ij c_ij^2 = <q q^T, k k^T>/bs^2,
 *           gradient 2/bs^2 (k k^T) q — three tcgen05 GEMMs with D as the long dimension, any Bg <= 4096,
 *           ~D/(1.5 Bg) times fewer flops.  off_diag is the DIFFERENCE <Gq,Gk>/bs^2 - sum_i c_ii^2: its relative error
 *           is ~2e-6 * sum_i c_ii^2 / off_diag.  Since rank(c) <= Bg, D >= 2 Bg guarantees off_diag >= sum_i c_ii^2
 *           (no cancellation, e.g. the reference's D = 8192 with any gathered batch <= 4096); a projector narrower
 *           than twice the batch can reach c ~ I, where the difference cancels.
 *           RMCL_BARLOW_DIRECT: c evaluated tile by tile, off-diagonal squares summed directly (Bg <= 256).
 *           RMCL_BARLOW_AUTO: GRAM if D >= 2 Bg or Bg > 256, else DIRECT.
 */
enum { RMCL_BARLOW_AUTO = 0, RMCL_BARLOW_DIRECT = 1, RMCL_BARLOW_GRAM = 2 };
size_t rmcl_barlow_workspace_bytes(int Bg, int D);
int rmcl_barlow_fwd_bwd(const void* q, rmcl_dtype q_dtype, const void* k, rmcl_dtype k_dtype, int Bg, int D,
                        int b0, int Bl, float inv_bs, float lambda, float w_on, float w_off,
                        float loss_scale, int path, float* on_diag, float* off_diag, float* loss,
                        float* dq, float* cdiag, void* workspace, size_t workspace_bytes, void* stream);

/* Measurement aid (off by default): when enabled on the calling thread, rmcl_infonce_fwd_bwd
 * records CUDA events on its stream around its three launches (prep, split-K partial, finalize);
 * rmcl_profile_infonce_ms waits for the last of them and returns the three durations of the most
 * recent call.  bench.py uses this for the per-kernel roofline; the product path leaves it off. */
int rmcl_profile_enable(int on);
int rmcl_profile_infonce_ms(float* out3);

/* Measurement aid: while `dev_buf` (device memory, rmcl_debug_tc_timeline_words() int64) is set on
 * the calling thread, CTA (0,0) of every tcgen05 InfoNCE launch stamps clock64() at its protocol
 * points into it (layout: csrc/infonce_tc.cu, tl_stamp).  NULL switches it off (the default). */
int rmcl_debug_tc_timeline(long long* dev_buf);
int rmcl_debug_tc_timeline_words(void);

/* ---------------------------------------------------------------------------------------------
 * Ring-buffer enqueue:  queue[:, ptr:ptr+B] = keys^T ;  ptr = (ptr + B) % K, all on the device.
 * replaces: objectives.py:244-248 (int(ptr) D2H sync, strided slice-assign, host modulo, H2D
 *           scalar write); MoCo/MoCo_RMCL.py:81-94.
 *   queue [C,K] (row stride ldq), keys [B,C] contiguous, *ptr_dev int64 in [0,K).
 * K % B != 0 is RMCL_E_BADARG (the reference's commented-out assert, objectives.py:245); with it
 * the slice never wraps.  The "skip unless B == per_step_bs" rule (objectives.py:242-243) lives
 * in the host wrapper.  While the kernel runs, the upper 32 bits of *ptr_dev are used as an
 * arrival counter; they are zero again when it finishes.
 */
int rmcl_enqueue(void* queue, rmcl_dtype queue_dtype, const void* keys, rmcl_dtype keys_dtype,
                 int64_t* ptr_dev, int B, int C, int64_t K, int64_t ldq, void* stream);

/* Splits an fp32 queue [C,K] (row stride ldq) into the bf16 hi/lo layout RMCL_BF16_HILO describes:
 * hilo_bf16 [2C,K] (row stride ld_hilo).  One pass, 4+4 bytes per element; afterwards rmcl_enqueue_shadow
 * (shadow_planes = 2) keeps the copy current at B*C*4 bytes per step.
 * replaces: nothing in the reference — it is what lets the fp32 PGD-inner InfoNCE (pgd_attack_vilt.py:141,
 *           152-158) run on the bf16 tensor cores at fp32 accuracy instead of on cuBLAS SGEMM. */
int rmcl_queue_split(const void* queue_f32, int C, int64_t K, int64_t ldq, void* hilo_bf16, int64_t ld_hilo,
                     void* stream);

/* Same, and the same B columns are also written (rounded to bf16) into `shadow_bf16` [C,K] (row
 * stride lds): a half-precision copy of the queue kept current at B*C*2 bytes per step, so that the
 * tcgen05 InfoNCE path can run against a checkpoint-compatible fp32 queue.
 * replaces: the per-call autocast cast of the whole queue under Lightning precision=16
 *           (objectives.py:270-272, 329: `proj_queue.clone()` + half-precision einsum).
 * shadow_planes = 1: shadow is [C,K] bf16(queue).  shadow_planes = 2: shadow is the [2C,K] hi/lo pair of
 * RMCL_BF16_HILO (its first C rows are the same bf16(queue) plane, so one buffer serves both InfoNCE modes). */
int rmcl_enqueue_shadow(void* queue, rmcl_dtype queue_dtype, void* shadow_bf16, int64_t lds, int shadow_planes,
                        const void* keys, rmcl_dtype keys_dtype, int64_t* ptr_dev, int B, int C,
                        int64_t K, int64_t ldq, void* stream);

/* ---------------------------------------------------------------------------------------------
 * One PGD perturbation update on delta[B,N] given grad[B,N] (per-sample norms over N).
 * replaces: attack/pgd_attack_vilt.py:162-173 (clone/float, inf-norm, clamp, scale, add, clamp:
 *           7 kernels, 5 passes).  Works on pixels [B,3,H,W] or embeddings [B,L,768] viewed [B,N].
 * REF_LINF in f32 is bit-identical to the reference: fadd(delta, fdiv(fmul(lr,g), d)).
 * L2 with eps > 0 takes the projection scale from |delta|^2 + 2a<delta,g> + a^2|g|^2 (a = lr/|g|),
 * so delta is written once, already projected.
 * NaN: a NaN gradient element makes its sample's norm NaN and hence the whole sample's delta NaN (REF_LINF, L2), as
 * torch.norm / torch.clamp do; SIGN_LINF maps it to a zero step like torch.sign.  Other samples are unaffected.
 * One persistent launch; the gradient is re-read from L2, so DRAM sees 12 B/element.
 *   workspace  rmcl_pgd_workspace_bytes(B, N, grad_dtype) bytes, 256-byte aligned, caller-owned
 *              (arrival counters + per-chunk partial norms).  It must be ZERO-FILLED once before
 *              its first use; every call leaves the counters zeroed again, so no memset launch is
 *              needed between calls.  Calls sharing a workspace must be stream-ordered.
 *              delta is updated in place.
 */
size_t rmcl_pgd_workspace_bytes(int B, int64_t N, rmcl_dtype grad_dtype);

int rmcl_pgd_step(void* delta, rmcl_dtype delta_dtype, const void* grad, rmcl_dtype grad_dtype,
                  int B, int64_t N, float lr, float eps, int mode, void* workspace,
                  size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Host-buffer convenience used for the end-to-end measurement: one kernels-only RMCL step
 * (EMA -> InfoNCE fwd+bwd -> enqueue) with q, k in HOST memory and loss/dq returned to HOST
 * memory.  Parameters, queue and pointer stay resident on the device (they are model state).
 * Kernels are issued on `stream`; the host<->device copies ride a library-owned side stream so that
 * they overlap the EMA and the enqueue.  The call returns after everything it issued has completed
 * (this one entry point does synchronise: its results are host-visible on return).
 */
int rmcl_step_host(const rmcl_ema_chunk* chunks_dev, int64_t n_chunks, double m, rmcl_dtype param_dtype,
                   const void* q_host, const void* k_host, rmcl_dtype qk_dtype,
                   void* q_dev, void* k_dev,
                   void* queue, rmcl_dtype queue_dtype, int64_t* ptr_dev, int B, int C, int64_t K,
                   float tau, int path, float* loss_dev, float* dq_dev, float* k_hat_dev,
                   float* loss_host, float* dq_host, void* workspace, size_t workspace_bytes,
                   void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RMCL_B200_H */
