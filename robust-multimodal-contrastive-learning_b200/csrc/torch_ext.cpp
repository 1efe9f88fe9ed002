// rmcl_b200_torch.so — TORCH_LIBRARY(rmcl, ...) operators over the C-ABI of librmcl_b200.so.
//
// north_star: "exposed through a thin C-ABI torch C++ extension, so it drops into the vilt/modules PyTorch Lightning loop
// unchanged".  The seam it serves is vilt/modules/vilt_module.py:463-464 -> objectives.compute_moco_contrastive
// (objectives.py:217): every kernel call the Python facade makes goes through one of these operators.
//
// What this layer does — and nothing else: argument checks, output / workspace allocation through ATen's caching allocator
// (so the calls are CUDA-graph capturable and stream-safe), the current-stream lookup, and the call of the `extern "C"`
// entry point of include/rmcl_b200.h.  No kernel lives here; the raw symbols stay directly callable (tests/test_capi_cpu.py,
// INTEGRATION.md).  Compared with the ctypes binding this removes ~20 us of Python per InfoNCE call (25 marshalled arguments,
// six torch.empty calls), which is as long as the cfg4-shaped kernel chain itself.
//
//   rmcl::ema_multi_        objectives.py:219-224, 257-260        rmcl::enqueue_         objectives.py:244-248
//   rmcl::infonce_fwd_bwd   objectives.py:326-334+351 (+337-349)  rmcl::pgd_step_        attack/pgd_attack_vilt.py:162-173
//   rmcl::infonce_loss      the same with autograd for q          rmcl::queue_stats      objectives.py:337-349 (once per step)
//   rmcl::barlow_fwd_bwd    objectives.py:480-486
#include <ATen/ATen.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/autograd.h>
#include <torch/library.h>

#include <mutex>
#include <tuple>
#include <unordered_map>

#include "../../include/rmcl_b200.h"

namespace {

using at::Tensor;

rmcl_dtype dt_of(const Tensor& t, const char* what) {
  if (t.scalar_type() == at::kFloat) return RMCL_F32;
  if (t.scalar_type() == at::kBFloat16) return RMCL_BF16;
  TORCH_CHECK(false, "rmcl_b200: ", what, " must be float32 or bfloat16, got ", t.scalar_type());
}

void need_cuda(const Tensor& t, const char* what) {
  TORCH_CHECK(t.is_cuda(), "rmcl_b200 ops run on CUDA tensors only (there is no CPU path): ", what);
}

void check(int rc, const char* what) {
  TORCH_CHECK(rc == RMCL_OK, what, " failed (status ", rc, "): ", rmcl_last_error());
}

void* stream_of(const Tensor& t) { return (void*)at::cuda::getCurrentCUDAStream(t.get_device()).stream(); }

// ---- workspaces: one per (kind, device, shape key, stream) from the caching allocator, kept for the life of the process
struct WsKey {
  int kind, dev;
  int64_t a, b, c, d, e;
  bool operator==(const WsKey& o) const {
    return kind == o.kind && dev == o.dev && a == o.a && b == o.b && c == o.c && d == o.d && e == o.e;
  }
};
struct WsHash {
  size_t operator()(const WsKey& k) const {
    size_t h = 1469598103934665603ull;
    for (int64_t v : {(int64_t)k.kind, (int64_t)k.dev, k.a, k.b, k.c, k.d, k.e}) h = (h ^ (size_t)v) * 1099511628211ull;
    return h;
  }
};
std::mutex g_ws_mu;
std::unordered_map<WsKey, Tensor, WsHash> g_ws;

// returns a 256-byte aligned pointer into a cached uint8 tensor of at least `bytes`
std::pair<void*, size_t> workspace(const WsKey& key, size_t bytes, const Tensor& like, bool zero_fill) {
  std::lock_guard<std::mutex> lock(g_ws_mu);
  auto it = g_ws.find(key);
  if (it == g_ws.end() || (size_t)it->second.numel() < bytes + 256) {
    auto opt = at::TensorOptions().dtype(at::kByte).device(like.device());
    Tensor t = zero_fill ? at::zeros({(int64_t)bytes + 256}, opt) : at::empty({(int64_t)bytes + 256}, opt);
    it = g_ws.insert_or_assign(key, t).first;
  }
  auto p = reinterpret_cast<uintptr_t>(it->second.data_ptr());
  const size_t off = (256 - (p & 255)) & 255;
  return {reinterpret_cast<void*>(p + off), (size_t)it->second.numel() - off};
}

template <typename T> T* ptr_or_null(const Tensor& t) { return t.defined() && t.numel() > 0 ? t.data_ptr<T>() : nullptr; }

// ------------------------------------------------------------------------------------------------ EMA
void ema_multi_(const Tensor& table, int64_t n_chunks, double m, int64_t dtype) {
  need_cuda(table, "chunk table");
  TORCH_CHECK(table.scalar_type() == at::kByte && table.is_contiguous(), "chunk table must be a contiguous uint8 tensor");
  TORCH_CHECK((int64_t)table.numel() >= n_chunks * (int64_t)sizeof(rmcl_ema_chunk), "chunk table shorter than n_chunks");
  c10::cuda::CUDAGuard guard(table.device());
  check(rmcl_ema_multi(reinterpret_cast<const rmcl_ema_chunk*>(table.data_ptr()), n_chunks, m, (rmcl_dtype)dtype, stream_of(table)),
        "rmcl_ema_multi");
}

// -------------------------------------------------------------------------------------------- InfoNCE
enum { W_LOSS = 1, W_ROW = 2, W_LSE = 4, W_POS = 8, W_ARGMAX = 16, W_DQ = 32, W_DK = 64, W_KHAT = 128 };

// returns [loss, loss_per_row, lse, pos, argmax, dq, dk, k_hat, diag]; outputs not asked for are empty tensors
std::vector<Tensor> infonce_fwd_bwd(const Tensor& q_in, const Tensor& k_in, const Tensor& queue, double tau, double loss_scale,
                                    bool normalize_k, bool need_grad, int64_t path, const c10::optional<Tensor>& colnorm2,
                                    const c10::optional<Tensor>& sum_vec, const c10::optional<Tensor>& sum_unit, double cos_eps,
                                    int64_t want, bool partial_only) {
  need_cuda(q_in, "q");
  need_cuda(k_in, "k");
  need_cuda(queue, "queue");
  // queue: [C,K] fp32 / bf16, or the [2C,K] bf16 hi/lo planes of an fp32 queue (RMCL_BF16_HILO: fp32-accurate path)
  const bool hilo = queue.dim() == 2 && q_in.dim() == 2 && queue.scalar_type() == at::kBFloat16 && queue.size(0) == 2 * q_in.size(1);
  TORCH_CHECK(q_in.dim() == 2 && k_in.sizes() == q_in.sizes() && queue.dim() == 2 && (hilo || queue.size(0) == q_in.size(1)),
              "shape mismatch: q ", q_in.sizes(), " k ", k_in.sizes(), " queue ", queue.sizes());
  TORCH_CHECK(queue.stride(1) == 1, "queue must be [C,K] with K contiguous (reference layout)");
  c10::cuda::CUDAGuard guard(q_in.device());
  // host overhead matters here (the whole op is ~25 us of GPU time at the cfg4 shape): no view tensors unless needed
  Tensor q = q_in.requires_grad() ? q_in.detach() : q_in, k = k_in.requires_grad() ? k_in.detach() : k_in;
  if (!q.is_contiguous()) q = q.contiguous();
  if (!k.is_contiguous()) k = k.contiguous();
  if (q.scalar_type() != at::kFloat && q.scalar_type() != at::kBFloat16) q = q.to(at::kFloat);   // fp16 under precision=16
  if (k.scalar_type() != at::kFloat && k.scalar_type() != at::kBFloat16) k = k.to(at::kFloat);
  const int64_t B = q.size(0), C = q.size(1), K = queue.size(1), ldq = queue.stride(0);
  const rmcl_dtype qdt = hilo ? RMCL_BF16_HILO : dt_of(queue, "queue");
  const size_t need = rmcl_infonce_workspace_bytes((int)B, (int)C, K, qdt, (int)path);
  TORCH_CHECK(need > 0, "rmcl_infonce_workspace_bytes failed: ", rmcl_last_error());
  auto ws = workspace(WsKey{0, q.get_device(), B, C, K, (int64_t)qdt * 8 + path, (int64_t)(uintptr_t)stream_of(q)}, need, q, true);
  auto f32 = q.options().dtype(at::kFloat);
  Tensor none;
  Tensor loss = (want & W_LOSS) ? at::empty({}, f32) : none;
  Tensor row = (want & W_ROW) ? at::empty({B}, f32) : none;
  Tensor lse = (want & W_LSE) ? at::empty({B}, f32) : none;
  Tensor pos = (want & W_POS) ? at::empty({B}, f32) : none;
  Tensor argmax = (want & W_ARGMAX) ? at::empty({B}, q.options().dtype(at::kLong)) : none;
  Tensor dq = (need_grad && (want & W_DQ)) ? at::empty({B, C}, f32) : none;
  Tensor dk = (need_grad && (want & W_DK)) ? at::empty({B, C}, f32) : none;
  Tensor khat = (want & W_KHAT) ? at::empty({B, C}, f32) : none;
  unsigned flags = (normalize_k ? RMCL_INFONCE_NORMALIZE_K : 0u) | (need_grad ? 0u : RMCL_INFONCE_NO_GRAD) |
                   (partial_only ? RMCL_INFONCE_DEBUG_PARTIAL_ONLY : 0u);
  Tensor diag;
  if (colnorm2.has_value() && colnorm2->defined()) {
    TORCH_CHECK(sum_vec.has_value() && sum_unit.has_value(), "diagnostics need colnorm2, sum_vec and sum_unit");
    TORCH_CHECK(colnorm2->numel() == K && sum_vec->numel() == C && sum_unit->numel() == C, "QueueStats of a different queue");
    diag = at::empty({6}, f32);
    check(rmcl_infonce_fwd_bwd_diag(q.data_ptr(), dt_of(q, "q"), k.data_ptr(), dt_of(k, "k"), queue.data_ptr(), qdt, (int)B, (int)C, K,
                                    ldq, (float)tau, (float)loss_scale, flags, (int)path, ptr_or_null<float>(loss),
                                    ptr_or_null<float>(row), ptr_or_null<float>(lse), ptr_or_null<float>(pos),
                                    ptr_or_null<int64_t>(argmax), ptr_or_null<float>(dq), ptr_or_null<float>(dk),
                                    ptr_or_null<float>(khat), colnorm2->data_ptr<float>(), sum_vec->data_ptr<float>(),
                                    sum_unit->data_ptr<float>(), (float)cos_eps, diag.data_ptr<float>(), ws.first, ws.second,
                                    stream_of(q)),
          "rmcl_infonce_fwd_bwd_diag");
  } else {
    check(rmcl_infonce_fwd_bwd(q.data_ptr(), dt_of(q, "q"), k.data_ptr(), dt_of(k, "k"), queue.data_ptr(), qdt, (int)B, (int)C, K, ldq,
                               (float)tau, (float)loss_scale, flags, (int)path, ptr_or_null<float>(loss), ptr_or_null<float>(row),
                               ptr_or_null<float>(lse), ptr_or_null<float>(pos), ptr_or_null<int64_t>(argmax), ptr_or_null<float>(dq),
                               ptr_or_null<float>(dk), ptr_or_null<float>(khat), ws.first, ws.second, stream_of(q)),
          "rmcl_infonce_fwd_bwd");
  }
  Tensor empty0;   // one zero-size tensor stands in for every output that was not asked for
  auto e = [&](const Tensor& t) {
    if (t.defined()) return t;
    if (!empty0.defined()) empty0 = at::empty({0}, f32);
    return empty0;
  };
  return {e(loss), e(row), e(lse), e(pos), e(argmax), e(dq), e(dk), e(khat), e(diag)};
}

// loss = CE([q^.k^ ; q^.queue]/T, 0) with dq from the same fused pass (objectives.py:326-334+351;
// attack/pgd_attack_vilt.py:147-158).  k and queue get no gradient, as in the reference (no_grad at objectives.py:262,
// .clone().detach() at 270).
struct InfoNceLossFn : public torch::autograd::Function<InfoNceLossFn> {
  static torch::autograd::variable_list forward(torch::autograd::AutogradContext* ctx, const Tensor& q, const Tensor& k,
                                                const Tensor& queue, double tau, int64_t path, bool normalize_k) {
    at::AutoDispatchBelowADInplaceOrView guard;
    auto r = infonce_fwd_bwd(q, k, queue, tau, 1.0, normalize_k, true, path, c10::nullopt, c10::nullopt, c10::nullopt, 1e-6,
                             W_LOSS | W_ARGMAX | W_DQ | (normalize_k ? W_KHAT : 0), false);
    ctx->save_for_backward({r[5]});
    ctx->saved_data["q_dtype"] = (int64_t)q.scalar_type();
    ctx->mark_non_differentiable({r[4], r[7]});
    return {r[0], r[4], r[7]};
  }
  static torch::autograd::variable_list backward(torch::autograd::AutogradContext* ctx, torch::autograd::variable_list grads) {
    auto dq = ctx->get_saved_variables()[0];
    auto dt = (at::ScalarType)ctx->saved_data["q_dtype"].toInt();
    return {(dq * grads[0]).to(dt), Tensor(), Tensor(), Tensor(), Tensor(), Tensor()};
  }
};

std::tuple<Tensor, Tensor, Tensor> infonce_loss(const Tensor& q, const Tensor& k, const Tensor& queue, double tau, int64_t path,
                                                bool normalize_k) {
  auto r = InfoNceLossFn::apply(q, k, queue, tau, path, normalize_k);
  return {r[0], r[1], r[2]};
}

// forward-only kernel behind the CUDA key (inference_mode, where the Autograd key is absent)
std::tuple<Tensor, Tensor, Tensor> infonce_loss_cuda(const Tensor& q, const Tensor& k, const Tensor& queue, double tau, int64_t path,
                                                     bool normalize_k) {
  auto r = infonce_fwd_bwd(q, k, queue, tau, 1.0, normalize_k, false, path, c10::nullopt, c10::nullopt, c10::nullopt, 1e-6,
                           W_LOSS | W_ARGMAX | (normalize_k ? W_KHAT : 0), false);
  return {r[0], r[4], r[7]};
}

std::tuple<Tensor, Tensor, Tensor> queue_stats(const Tensor& queue, double cos_eps) {
  need_cuda(queue, "queue");
  TORCH_CHECK(queue.dim() == 2 && queue.stride(1) == 1, "queue must be [C,K] with K contiguous (reference layout)");
  c10::cuda::CUDAGuard guard(queue.device());
  auto f32 = queue.options().dtype(at::kFloat);
  Tensor colnorm2 = at::empty({queue.size(1)}, f32), sum_vec = at::empty({queue.size(0)}, f32), sum_unit = at::empty({queue.size(0)}, f32);
  check(rmcl_queue_stats(queue.data_ptr(), dt_of(queue, "queue"), (int)queue.size(0), queue.size(1), queue.stride(0), (float)cos_eps,
                         colnorm2.data_ptr<float>(), sum_vec.data_ptr<float>(), sum_unit.data_ptr<float>(), stream_of(queue)),
        "rmcl_queue_stats");
  return {colnorm2, sum_vec, sum_unit};
}

// -------------------------------------------------------------------------------------------- enqueue
void enqueue_(Tensor& queue, const Tensor& keys_in, Tensor& ptr, const c10::optional<Tensor>& shadow) {
  need_cuda(queue, "queue");
  need_cuda(keys_in, "keys");
  need_cuda(ptr, "ptr");
  TORCH_CHECK(ptr.scalar_type() == at::kLong && ptr.numel() == 1, "ptr must be an int64 tensor with one element");
  TORCH_CHECK(queue.dim() == 2 && queue.stride(1) == 1 && keys_in.dim() == 2 && keys_in.size(1) == queue.size(0),
              "shape mismatch: queue ", queue.sizes(), " keys ", keys_in.sizes());
  c10::cuda::CUDAGuard guard(queue.device());
  Tensor keys = keys_in.detach().contiguous();
  if (shadow.has_value() && shadow->defined()) {
    const Tensor& sh = *shadow;
    TORCH_CHECK(sh.is_cuda() && sh.scalar_type() == at::kBFloat16 && sh.dim() == 2 && sh.size(1) == queue.size(1) && sh.stride(1) == 1 &&
                    (sh.size(0) == queue.size(0) || sh.size(0) == 2 * queue.size(0)),
                "shadow must be a bf16 [C,K] or hi/lo [2C,K] CUDA tensor with K contiguous");
    check(rmcl_enqueue_shadow(queue.data_ptr(), dt_of(queue, "queue"), sh.data_ptr(), sh.stride(0), (int)(sh.size(0) / queue.size(0)),
                              keys.data_ptr(), dt_of(keys, "keys"),
                              ptr.data_ptr<int64_t>(), (int)keys.size(0), (int)keys.size(1), queue.size(1), queue.stride(0),
                              stream_of(queue)),
          "rmcl_enqueue_shadow");
    return;
  }
  check(rmcl_enqueue(queue.data_ptr(), dt_of(queue, "queue"), keys.data_ptr(), dt_of(keys, "keys"), ptr.data_ptr<int64_t>(),
                     (int)keys.size(0), (int)keys.size(1), queue.size(1), queue.stride(0), stream_of(queue)),
        "rmcl_enqueue");
}

// ------------------------------------------------------------------------------------------------ PGD
void pgd_step_(Tensor& delta, const Tensor& grad, double lr, double eps, int64_t mode) {
  need_cuda(delta, "delta");
  need_cuda(grad, "grad");
  TORCH_CHECK(delta.sizes() == grad.sizes(), "delta/grad shape mismatch");
  TORCH_CHECK(delta.is_contiguous() && grad.is_contiguous(), "delta and grad must be contiguous");
  c10::cuda::CUDAGuard guard(delta.device());
  const int64_t B = delta.size(0), N = delta.numel() / B;
  const rmcl_dtype gdt = dt_of(grad, "grad");
  const size_t need = rmcl_pgd_workspace_bytes((int)B, N, gdt);
  TORCH_CHECK(need > 0, "rmcl_pgd_workspace_bytes failed: ", rmcl_last_error());
  // zero-filled once; the kernel leaves its control words zeroed.  One workspace per stream (see the header).
  auto ws = workspace(WsKey{1, delta.get_device(), B, N, (int64_t)gdt, 0, (int64_t)(uintptr_t)stream_of(delta)}, need, delta, true);
  check(rmcl_pgd_step(delta.data_ptr(), dt_of(delta, "delta"), grad.data_ptr(), gdt, (int)B, N, (float)lr, (float)eps, (int)mode, ws.first,
                      ws.second, stream_of(delta)),
        "rmcl_pgd_step");
}

// ----------------------------------------------------------------------------------------- Barlow Twins
enum { BW_ON = 1, BW_OFF = 2, BW_LOSS = 4, BW_DQ = 8, BW_CDIAG = 16 };

// returns [on_diag, off_diag, loss, dq, cdiag]
std::vector<Tensor> barlow_fwd_bwd(const Tensor& q, const Tensor& k, double inv_bs, double lam, int64_t b0, int64_t Bl, double w_on,
                                   double w_off, double loss_scale, int64_t path, int64_t want) {
  need_cuda(q, "q");
  need_cuda(k, "k");
  TORCH_CHECK(q.dim() == 2 && q.sizes() == k.sizes() && q.is_contiguous() && k.is_contiguous(),
              "q and k must be contiguous [Bg, D] tensors of the same shape");
  c10::cuda::CUDAGuard guard(q.device());
  const int64_t Bg = q.size(0), D = q.size(1);
  const size_t need = rmcl_barlow_workspace_bytes((int)Bg, (int)D);
  TORCH_CHECK(need > 0, "rmcl_barlow_workspace_bytes failed: ", rmcl_last_error());
  auto ws = workspace(WsKey{2, q.get_device(), Bg, D, 0, 0, (int64_t)(uintptr_t)stream_of(q)}, need, q, false);
  auto f32 = q.options().dtype(at::kFloat);
  Tensor none;
  Tensor on = (want & BW_ON) ? at::empty({}, f32) : none, off = (want & BW_OFF) ? at::empty({}, f32) : none;
  Tensor loss = (want & BW_LOSS) ? at::empty({}, f32) : none, dq = (want & BW_DQ) ? at::empty({Bl, D}, f32) : none;
  Tensor cdiag = (want & BW_CDIAG) ? at::empty({D}, f32) : none;
  check(rmcl_barlow_fwd_bwd(q.data_ptr(), dt_of(q, "q"), k.data_ptr(), dt_of(k, "k"), (int)Bg, (int)D, (int)b0, (int)Bl, (float)inv_bs,
                            (float)lam, (float)w_on, (float)w_off, (float)loss_scale, (int)path, ptr_or_null<float>(on),
                            ptr_or_null<float>(off), ptr_or_null<float>(loss), ptr_or_null<float>(dq), ptr_or_null<float>(cdiag),
                            ws.first, ws.second, stream_of(q)),
        "rmcl_barlow_fwd_bwd");
  auto e = [&](const Tensor& t) { return t.defined() ? t : at::empty({0}, f32); };
  return {e(on), e(off), e(loss), e(dq), e(cdiag)};
}

}  // namespace

TORCH_LIBRARY(rmcl, m) {
  m.def("ema_multi_(Tensor table, int n_chunks, float m, int dtype) -> ()");
  m.def("infonce_fwd_bwd(Tensor q, Tensor k, Tensor queue, float tau, float loss_scale, bool normalize_k, bool need_grad, int path, "
        "Tensor? colnorm2, Tensor? sum_vec, Tensor? sum_unit, float cos_eps, int want, bool partial_only) -> Tensor[]");
  m.def("infonce_loss(Tensor q, Tensor k, Tensor queue, float tau, int path, bool normalize_k) -> (Tensor, Tensor, Tensor)");
  m.def("queue_stats(Tensor queue, float cos_eps) -> (Tensor, Tensor, Tensor)");
  m.def("enqueue_(Tensor(a!) queue, Tensor keys, Tensor(b!) ptr, Tensor? shadow) -> ()");
  m.def("pgd_step_(Tensor(a!) delta, Tensor grad, float lr, float eps, int mode) -> ()");
  m.def("barlow_fwd_bwd(Tensor q, Tensor k, float inv_bs, float lam, int b0, int Bl, float w_on, float w_off, float loss_scale, "
        "int path, int want) -> Tensor[]");
}

TORCH_LIBRARY_IMPL(rmcl, CUDA, m) {
  m.impl("ema_multi_", &ema_multi_);
  m.impl("infonce_fwd_bwd", &infonce_fwd_bwd);
  m.impl("infonce_loss", &infonce_loss_cuda);
  m.impl("queue_stats", &queue_stats);
  m.impl("enqueue_", &enqueue_);
  m.impl("pgd_step_", &pgd_step_);
  m.impl("barlow_fwd_bwd", &barlow_fwd_bwd);
}

// the differentiable loss: the Autograd key wraps the CUDA op in the custom function above
TORCH_LIBRARY_IMPL(rmcl, Autograd, m) { m.impl("infonce_loss", &infonce_loss); }
