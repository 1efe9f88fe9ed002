"""Device-tensor wrappers over the C-ABI (include/rmcl_b200.h).

torch is used for what the task allows it for — device memory, streams and autograd plumbing.
Every function here launches hand-written sm_100a kernels from librmcl_b200.so on the current
CUDA stream; none has a CPU or eager-torch fallback (CPU tensors raise).
"""
import ctypes as C

import torch

from . import _lib
from ._lib import check

_DT = {torch.float32: _lib.RMCL_F32, torch.bfloat16: _lib.RMCL_BF16}


def _dt(t):
    try:
        return _DT[t.dtype]
    except KeyError:
        raise TypeError(f"rmcl_b200 supports float32 and bfloat16 tensors, got {t.dtype}") from None


def _need_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("rmcl_b200 ops run on CUDA tensors only (there is no CPU path)")


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


# ------------------------------------------------------------------------------------ EMA
class EmaPlan:
    """Chunk table for the multi-tensor momentum update (built once, reused every step).

    Mirrors the pairing of vilt/modules/objectives.py:222 —
    ``zip(q_layer.parameters(), k_layer.parameters())`` — for any number of layers.

    The device table holds raw addresses, so the plan remembers the tensors it was built from and a
    fingerprint of their storage (address, size, dtype, device).  ``ema_multi_`` compares it on every call
    and rebuilds the table when a parameter was re-bound in the meantime (``param.data = ...`` — the
    reference's own EMA line does that —, ``.to()``, ``.half()``, FSDP/DeepSpeed flattening); without the
    check the kernel would keep updating the old storage and the key encoder would silently stop tracking.
    """

    def __init__(self, params_k, params_q, chunk_elems=8192):
        self.params_k, self.params_q = list(params_k), list(params_q)
        if len(self.params_k) != len(self.params_q):
            raise ValueError("key/query parameter lists differ in length")
        self.chunk_elems = int(chunk_elems)
        self.rebuilds = -1
        self._build()

    def fingerprint(self):
        """Storage addresses of every tensor (~25 us of host time for the 322 tensors of ViLT-B/32): a re-bound or
        re-allocated parameter shows up as a different address; shapes and dtypes are validated when the table is built."""
        return [t.data_ptr() for t in self.params_k] + [t.data_ptr() for t in self.params_q]

    def _build(self):
        self.groups = []  # (dtype_enum, device_table, n_chunks, keepalive)
        self.n_params = 0
        self.rebuilds += 1
        by_dtype = {}
        for pk, pq in zip(self.params_k, self.params_q):
            pk, pq = getattr(pk, "data", pk), getattr(pq, "data", pq)
            _need_cuda(pk, pq)
            if pk.shape != pq.shape or pk.dtype != pq.dtype:
                raise ValueError(f"parameter pair mismatch: {tuple(pk.shape)}/{pk.dtype} vs {tuple(pq.shape)}/{pq.dtype}")
            if not (pk.is_contiguous() and pq.is_contiguous()):
                raise ValueError("EMA expects contiguous parameters")
            by_dtype.setdefault(pk.dtype, []).append((pk, pq))
            self.n_params += pk.numel()
        L = _lib.lib()
        for dtype, pairs in by_dtype.items():
            n = len(pairs)
            kp = (C.c_void_p * n)(*[p[0].data_ptr() for p in pairs])
            qp = (C.c_void_p * n)(*[p[1].data_ptr() for p in pairs])
            ne = (C.c_uint64 * n)(*[p[0].numel() for p in pairs])
            dt = _DT[dtype]
            cnt = L.rmcl_ema_plan(kp, qp, ne, n, dt, self.chunk_elems, None)
            if cnt < 0:
                check(int(cnt), "rmcl_ema_plan")
            host = (_lib.EmaChunk * max(cnt, 1))()
            cnt2 = L.rmcl_ema_plan(kp, qp, ne, n, dt, self.chunk_elems, host)
            assert cnt2 == cnt
            raw = torch.frombuffer(bytearray(bytes(host)), dtype=torch.uint8)[: cnt * C.sizeof(_lib.EmaChunk)]
            table = raw.to(pairs[0][0].device)
            self.groups.append((dt, table, int(cnt), pairs))
        self._fp = self.fingerprint()

    def refresh(self):
        """Rebuilds the chunk table if any tensor's storage changed since it was built; returns True if it did."""
        if self.fingerprint() == self._fp:
            return False
        self._build()
        return True

    @property
    def n_chunks(self):
        return sum(g[2] for g in self.groups)


def ema_multi_(plan, m, check_storage=True):
    """k <- k*m + q*(1-m) for every pair in the plan, one launch per dtype group.  ``check_storage`` re-validates
    the cached addresses first (a few microseconds of host time per 100 tensors; see :class:`EmaPlan`)."""
    if check_storage:
        plan.refresh()
    if _lib.ffi() == "torch":
        tx = _lib.torch_ops()
        for dt, table, cnt, _ in plan.groups:
            tx.ema_multi_(table, cnt, float(m), dt)
        return
    L = _lib.lib()
    for dt, table, cnt, _ in plan.groups:
        check(L.rmcl_ema_multi(_p(table), cnt, float(m), dt, _stream()), "rmcl_ema_multi")


# -------------------------------------------------------------------------------- InfoNCE
_ws_cache = {}


def _workspace(B, Cdim, K, qdt, path, device):
    key = (B, Cdim, K, qdt, path, device)
    ws = _ws_cache.get(key)
    if ws is None:
        nbytes = _lib.lib().rmcl_infonce_workspace_bytes(B, Cdim, K, qdt, path)
        if nbytes == 0:
            check(-3, "rmcl_infonce_workspace_bytes")
        ws = torch.zeros(nbytes + 256, dtype=torch.uint8, device=device)     # zero-filled once: grid-barrier words (header)
        _ws_cache[key] = ws
    off = (-ws.data_ptr()) % 256
    return ws, off


def _split_ok(queue):
    return queue.shape[0] in (64, 128, 256) and queue.shape[1] % 8 == 0 and queue.stride(0) % 8 == 0 and queue.data_ptr() % 16 == 0


DIAG_NAMES = ("pos_dist", "pos_cosine", "pos_dot", "neg_dist", "neg_cosine", "neg_dot")
_OUT_ORDER = ("loss", "loss_per_row", "lse", "pos", "argmax", "dq", "dk", "k_hat")          # csrc/torch_ext.cpp
_WANT_BITS = {name: 1 << i for i, name in enumerate(_OUT_ORDER)}
_WANT_PLANS = {}   # want tuple -> (output mask, ((index, name), ...))


class QueueStats:
    """|queue_j|^2 per column and the two [C] sums the per-view diagnostics need (include/rmcl_b200.h,
    rmcl_queue_stats) — computed once per step, shared by every view's ``infonce_fwd_bwd(..., diag=...)``."""

    def __init__(self, queue, cos_eps=1e-6):
        _need_cuda(queue)
        if queue.dim() != 2 or queue.stride(1) != 1:
            raise ValueError("queue must be [C,K] with K contiguous (reference layout)")
        Cd, K = queue.shape
        f32 = dict(dtype=torch.float32, device=queue.device)
        self.cos_eps = float(cos_eps)
        self.shape = (Cd, K)
        if _lib.ffi() == "torch":
            self.colnorm2, self.sum_vec, self.sum_unit = _lib.torch_ops().queue_stats(queue, self.cos_eps)
            return
        self.colnorm2 = torch.empty(K, **f32)
        self.sum_vec = torch.empty(Cd, **f32)
        self.sum_unit = torch.empty(Cd, **f32)
        rc = _lib.lib().rmcl_queue_stats(_p(queue), _dt(queue), Cd, K, queue.stride(0), self.cos_eps, _p(self.colnorm2),
                                         _p(self.sum_vec), _p(self.sum_unit), _stream())
        check(rc, "rmcl_queue_stats")


def infonce_fwd_bwd(q, k, queue, temperature, *, loss_scale=1.0, normalize_k=False, need_grad=True,
                    path="auto", want=("loss", "loss_per_row", "lse", "pos", "argmax", "dq", "dk", "k_hat"),
                    diag=None, _partial_only=False):
    """Fused InfoNCE of raw projections ``q`` [B,C] against [``k`` ; ``queue`` [C,K]].

    Returns a dict with the requested outputs (see include/rmcl_b200.h).  ``loss`` is
    ``loss_scale * mean_i(lse_i - pos_i)`` — CrossEntropyLoss against label 0
    (vilt/modules/objectives.py:333-334,351).  With ``diag`` (a :class:`QueueStats` of ``queue`` — or of the
    fp32 buffer a bf16 shadow mirrors) the result also holds ``"diag"``: the six means of
    objectives.py:337-349 in the order of ``DIAG_NAMES``; ``k`` must then be the normalised key.
    """
    _need_cuda(q, k, queue)
    if q.dim() != 2 or k.shape != q.shape or queue.dim() != 2 or queue.shape[0] != q.shape[1]:
        raise ValueError(f"shape mismatch: q {tuple(q.shape)} k {tuple(k.shape)} queue {tuple(queue.shape)}")
    if queue.stride(1) != 1:
        raise ValueError("queue must be [C,K] with K contiguous (reference layout)")
    if queue.dtype not in _DT:
        raise TypeError(f"rmcl_b200 supports float32 and bfloat16 queues, got {queue.dtype}")
    if diag is not None and diag.shape != (q.shape[1], queue.shape[1]):
        raise ValueError(f"QueueStats of a {diag.shape} queue used with a {(q.shape[1], queue.shape[1])} queue")
    hilo = False
    if queue.dtype == torch.float32 and path != "simt" and _split_ok(queue):
        # fp32 queue on the tensor cores: the split-operand path reads its bf16 hi/lo planes (QueueShadow), fp32-accurate
        queue, hilo = hilo_of(queue), True
    elif queue.dtype == torch.float32 and path == "tcgen05":
        raise _lib.RmclError("tcgen05 InfoNCE on an fp32 queue needs C in {64,128,256}, K % 8 == 0 and a 16-byte aligned, "
                             f"contiguous-row queue (got C={queue.shape[0]} K={queue.shape[1]})")
    if _lib.ffi() == "torch":
        # host time matters here (the whole op is ~22 us of GPU time at the cfg4 shape): the output mask and the index list of
        # a ``want`` tuple are computed once
        plan = _WANT_PLANS.get(want) if isinstance(want, tuple) else None
        if plan is None:
            plan = (sum(bit for name, bit in _WANT_BITS.items() if name in want),
                    tuple((i, name) for i, name in enumerate(_OUT_ORDER) if name in want))
            if isinstance(want, tuple):
                _WANT_PLANS[want] = plan
        mask, picks = plan
        r = _lib.torch_ops().infonce_fwd_bwd(        # detached: this op is not differentiable (the loss op is infonce_loss)
            q.detach(), k.detach(), queue, float(temperature), float(loss_scale), bool(normalize_k), bool(need_grad), _lib.INFONCE_PATHS[path],
            None if diag is None else diag.colnorm2, None if diag is None else diag.sum_vec,
            None if diag is None else diag.sum_unit, 1e-6 if diag is None else diag.cos_eps, mask, bool(_partial_only))
        out = {name: r[i] for i, name in picks if r[i].numel() > 0}
        if "loss" in out:
            out["loss"] = out["loss"].reshape(())
        if diag is not None:
            out["diag"] = r[8]
        return out
    q, k = q.detach().contiguous(), k.detach().contiguous()
    if q.dtype not in _DT:      # e.g. fp16 projections under Lightning precision=16: the kernels take fp32 / bf16
        q = q.float()
    if k.dtype not in _DT:
        k = k.float()
    B, Cd = q.shape
    K, ldq = queue.shape[1], queue.stride(0)
    pth = _lib.INFONCE_PATHS[path]
    qdt = _lib.RMCL_BF16_HILO if hilo else _dt(queue)
    ws, off = _workspace(B, Cd, K, qdt, pth, q.device)
    f32 = dict(dtype=torch.float32, device=q.device)
    out = {}
    if "loss" in want:
        out["loss"] = torch.empty((), **f32)
    for name in ("loss_per_row", "lse", "pos"):
        if name in want:
            out[name] = torch.empty(B, **f32)
    if "argmax" in want:
        out["argmax"] = torch.empty(B, dtype=torch.int64, device=q.device)
    if need_grad and "dq" in want:
        out["dq"] = torch.empty(B, Cd, **f32)
    if need_grad and "dk" in want:
        out["dk"] = torch.empty(B, Cd, **f32)
    if "k_hat" in want:
        out["k_hat"] = torch.empty(B, Cd, **f32)
    flags = (_lib.FLAG_NORMALIZE_K if normalize_k else 0) | (0 if need_grad else _lib.FLAG_NO_GRAD)
    if _partial_only:   # measurement aid (bench.py): the partial kernel alone, on the workspace of a previous call
        flags |= _lib.FLAG_DEBUG_PARTIAL_ONLY
    if diag is not None:
        if diag.shape != (Cd, K):
            raise ValueError(f"QueueStats of a {diag.shape} queue used with a {(Cd, K)} queue")
        out["diag"] = torch.empty(len(DIAG_NAMES), **f32)
        rc = _lib.lib().rmcl_infonce_fwd_bwd_diag(
            _p(q), _dt(q), _p(k), _dt(k), _p(queue), qdt, B, Cd, K, ldq, float(temperature), float(loss_scale),
            flags, pth, _p(out.get("loss")), _p(out.get("loss_per_row")), _p(out.get("lse")), _p(out.get("pos")),
            _p(out.get("argmax")), _p(out.get("dq")), _p(out.get("dk")), _p(out.get("k_hat")),
            _p(diag.colnorm2), _p(diag.sum_vec), _p(diag.sum_unit), diag.cos_eps, _p(out["diag"]),
            C.c_void_p(ws.data_ptr() + off), ws.numel() - off, _stream())
        check(rc, "rmcl_infonce_fwd_bwd_diag")
        return out
    rc = _lib.lib().rmcl_infonce_fwd_bwd(
        _p(q), _dt(q), _p(k), _dt(k), _p(queue), qdt, B, Cd, K, ldq, float(temperature), float(loss_scale),
        flags, pth, _p(out.get("loss")), _p(out.get("loss_per_row")), _p(out.get("lse")), _p(out.get("pos")),
        _p(out.get("argmax")), _p(out.get("dq")), _p(out.get("dk")), _p(out.get("k_hat")),
        C.c_void_p(ws.data_ptr() + off), ws.numel() - off, _stream())
    check(rc, "rmcl_infonce_fwd_bwd")
    return out


class InfoNCE(torch.autograd.Function):
    """loss = CE([q^.k^ ; q^.queue]/T, 0) with dq computed in the same fused pass.

    Drop-in for the expression chain at objectives.py:326-334+351 (and its PGD twin,
    attack/pgd_attack_vilt.py:147-158).  ``k`` and ``queue`` get no gradient, as in the reference
    (``no_grad`` at objectives.py:262, ``.clone().detach()`` at 270).
    """

    @staticmethod
    def forward(ctx, q, k, queue, temperature, path="auto", normalize_k=False, diag=None):
        want = ("loss", "dq", "argmax", "pos", "lse") + (("k_hat",) if normalize_k else ())
        res = infonce_fwd_bwd(q, k, queue, temperature, normalize_k=normalize_k, path=path, want=want, diag=diag)
        ctx.save_for_backward(res["dq"])
        ctx.q_dtype = q.dtype
        ctx.extra = res
        dg = res["diag"] if diag is not None else torch.empty(0, device=q.device)
        k_hat = res["k_hat"] if normalize_k else torch.empty(0, device=q.device)
        ctx.mark_non_differentiable(res["argmax"], dg, k_hat)
        return res["loss"], res["argmax"], dg, k_hat

    @staticmethod
    def backward(ctx, grad_loss, _grad_argmax, _grad_diag, _grad_k_hat):
        (dq,) = ctx.saved_tensors
        return (dq * grad_loss).to(ctx.q_dtype), None, None, None, None, None, None


def infonce_loss(q, k, queue, temperature, path="auto", normalize_k=False, diag=None, return_k_hat=False):
    """(loss, argmax) — or (loss, argmax, diag[6]) when ``diag`` (QueueStats) is given — with autograd
    support for ``q``.  ``return_k_hat`` (with ``normalize_k``) appends the normalised key the same launch
    produced, so a caller holding raw key projections needs no separate normalisation pass."""
    if diag is None and _lib.ffi() == "torch":      # C++ autograd function behind torch.ops.rmcl.infonce_loss
        _need_cuda(q, k, queue)
        if queue.dim() != 2 or queue.shape[0] != q.shape[-1]:
            raise ValueError(f"shape mismatch: q {tuple(q.shape)} queue {tuple(queue.shape)}")
        if queue.dtype == torch.float32 and path != "simt" and _split_ok(queue):
            queue = hilo_of(queue)                 # fp32-accurate split-operand path (see infonce_fwd_bwd)
        elif queue.dtype == torch.float32 and path == "tcgen05":
            raise _lib.RmclError("tcgen05 InfoNCE on an fp32 queue needs C in {64,128,256}, K % 8 == 0 and an aligned queue")
        loss, argmax, k_hat = _lib.torch_ops().infonce_loss(q, k, queue, float(temperature), _lib.INFONCE_PATHS[path],
                                                            bool(normalize_k))
        if return_k_hat and not normalize_k:
            raise ValueError("return_k_hat needs normalize_k=True")
        return (loss, argmax, k_hat) if return_k_hat else (loss, argmax)
    loss, argmax, dg, k_hat = InfoNCE.apply(q, k, queue, temperature, path, normalize_k, diag)
    out = (loss, argmax) if diag is None else (loss, argmax, dg)
    if return_k_hat:
        if not normalize_k:
            raise ValueError("return_k_hat needs normalize_k=True")
        out = out + (k_hat,)
    return out


# -------------------------------------------------------------------------------- Barlow Twins
_bt_ws_cache = {}


def barlow_fwd_bwd(q, k, inv_bs, lam, *, b0=0, Bl=None, w_on=1.0, w_off=None, loss_scale=1.0, path="auto",
                   want=("on_diag", "off_diag", "loss", "dq", "cdiag")):
    """Fused Barlow-Twins loss of ``q``, ``k`` [Bg, D] (the batch gathered over all ranks): the D x D
    cross-correlation ``c = q.T @ k * inv_bs`` is evaluated tile by tile on the tensor cores and never stored.

    Returns ``on_diag = sum_i (c_ii-1)^2``, ``off_diag = sum_{i!=j} c_ij^2`` (objectives.py:483-484),
    ``loss = loss_scale*(on_diag + lam*off_diag)``, ``dq`` [Bl, D] = d(w_on*on_diag + w_off*off_diag)/dq[b0:b0+Bl]
    * loss_scale (``w_off`` defaults to ``lam``) and ``cdiag`` = diagonal(c).  ``path``: "gram" (= "auto": through the
    Gram matrices q q^T and k k^T, any batch) or "direct" (c tile by tile, gathered batch <= 256)."""
    _need_cuda(q, k)
    if q.dim() != 2 or q.shape != k.shape or not q.is_contiguous() or not k.is_contiguous():
        raise ValueError("q and k must be contiguous [Bg, D] tensors of the same shape")
    Bg, D = q.shape
    Bl = Bg - b0 if Bl is None else Bl
    w_off = lam if w_off is None else w_off
    if _lib.ffi() == "torch":
        names = ("on_diag", "off_diag", "loss", "dq", "cdiag")
        mask = sum(1 << i for i, n in enumerate(names) if n in want)
        r = _lib.torch_ops().barlow_fwd_bwd(q, k, float(inv_bs), float(lam), int(b0), int(Bl), float(w_on), float(w_off),
                                            float(loss_scale), _lib.BARLOW_PATHS[path], mask)
        return {n: (r[i].reshape(()) if n in ("on_diag", "off_diag", "loss") else r[i]) if n in want else None
                for i, n in enumerate(names)}
    L = _lib.lib()
    key = (Bg, D, q.device)
    ws = _bt_ws_cache.get(key)
    if ws is None:
        nbytes = L.rmcl_barlow_workspace_bytes(Bg, D)
        if nbytes == 0:
            check(-3, "rmcl_barlow_workspace_bytes")
        ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=q.device)
        _bt_ws_cache[key] = ws
    off = (-ws.data_ptr()) % 256
    f32 = dict(dtype=torch.float32, device=q.device)
    out = {"on_diag": torch.empty((), **f32) if "on_diag" in want else None,
           "off_diag": torch.empty((), **f32) if "off_diag" in want else None,
           "loss": torch.empty((), **f32) if "loss" in want else None,
           "dq": torch.empty(Bl, D, **f32) if "dq" in want else None,
           "cdiag": torch.empty(D, **f32) if "cdiag" in want else None}
    ptr = lambda t: None if t is None else _p(t)
    rc = L.rmcl_barlow_fwd_bwd(_p(q), _dt(q), _p(k), _dt(k), Bg, D, int(b0), int(Bl), float(inv_bs), float(lam), float(w_on),
                               float(w_off), float(loss_scale), _lib.BARLOW_PATHS[path], ptr(out["on_diag"]), ptr(out["off_diag"]),
                               ptr(out["loss"]),
                               ptr(out["dq"]), ptr(out["cdiag"]), ws.data_ptr() + off, ws.numel() - off, _stream())
    check(rc, "rmcl_barlow_fwd_bwd")
    return out


class BarlowTwins(torch.autograd.Function):
    """(on_diag, lam * off_diag) of objectives.py:480-486 with autograd support for ``q``.

    The reference all-reduces the D x D matrix between ranks (objectives.py:482); here the [B, D]
    projections are all-gathered instead and every rank evaluates the full matrix from the gathered batch.
    Both returned sums are differentiable: the backward combines the two gradient pieces (the diagonal one
    is elementwise given diagonal(c)) with whatever upstream weights arrive, without a second pass or a sync.
    ``k`` gets no gradient (``no_grad`` at objectives.py:461)."""

    @staticmethod
    def forward(ctx, q, k, inv_bs, lam, gather=None):
        ql, kl = q.detach().contiguous(), k.detach().contiguous()
        Bl = ql.shape[0]
        b0 = 0
        if gather is not None:
            qa, ka, b0 = gather(ql), gather(kl), gather.rank * Bl
        else:
            qa, ka = ql, kl
        res = barlow_fwd_bwd(qa, ka, inv_bs, lam, b0=b0, Bl=Bl, w_on=0.0, w_off=1.0, want=("on_diag", "off_diag", "dq", "cdiag"))
        ctx.save_for_backward(res["dq"], res["cdiag"], kl)
        ctx.inv_bs, ctx.lam, ctx.q_dtype = float(inv_bs), float(lam), q.dtype
        return res["on_diag"], res["off_diag"] * lam

    @staticmethod
    def backward(ctx, g_on, g_offs):
        dq_off, cdiag, kl = ctx.saved_tensors
        coeff = (cdiag - 1.0) * (g_on * (2.0 * ctx.inv_bs))             # [D]: d(on_diag)/dq[b, i] = coeff_i * k[b, i]
        out = dq_off * (g_offs * ctx.lam)
        out.addcmul_(kl.float(), coeff[None, :])
        return out.to(ctx.q_dtype), None, None, None, None


def barlow_twins_loss(q, k, inv_bs, lam, gather=None):
    """(on_diag, lam*off_diag) with autograd support for ``q``; ``gather`` = rmcl_b200.dist.Gather() under DDP."""
    return BarlowTwins.apply(q, k, inv_bs, lam, gather)


# -------------------------------------------------------------------------------- enqueue
class QueueShadow:
    """bf16 hi/lo copy of an fp32 queue buffer, kept current by ``enqueue_``.

    The reference's ``proj_queue`` is an fp32 buffer (vilt_module.py:92).  Its half-precision einsum under Lightning
    ``precision=16`` re-casts all C*K elements on every call (objectives.py:270-272), and its fp32 PGD-inner einsum
    (pgd_attack_vilt.py:141,152-158) is a CUDA-core / SGEMM contraction.  The shadow is one [2C, K] bf16 buffer:
      rows [0, C)   hi = bf16(queue)          -> ``get()``: the operand of the bf16 tcgen05 InfoNCE (main step under autocast)
      rows [C, 2C)  lo = bf16(queue - hi)     -> ``get_hilo()``: with hi, 16 mantissa bits per element — the operand of the
                                                 fp32-accurate split-operand tcgen05 InfoNCE (RMCL_BF16_HILO)
    built by one pass (``rmcl_queue_split``); afterwards each enqueue writes the B new columns into the queue and both
    planes in the same launch (B*C*4 bytes per step).  The copy is rebuilt if the fp32 buffer was replaced or modified
    in place by anything else (``load_state_dict``, ``.to()``).  A queue that is not fp32 gets a one-plane copy."""

    def __init__(self):
        self.tensor, self._key = None, None

    @staticmethod
    def _state(queue):
        return (queue.data_ptr(), tuple(queue.shape), queue._version)

    def _current(self, queue):
        _need_cuda(queue)
        if self.tensor is None or self._key != self._state(queue):
            self.tensor = queue_split(queue)      # one-off (init / checkpoint load)
            self._key = self._state(queue)
            _register_shadow(queue, self)
        return self.tensor

    def get(self, queue):
        """[C, K] bf16(queue): the hi plane (a view of the first C rows)."""
        return self._current(queue)[: queue.shape[0]]

    def get_hilo(self, queue):
        """[2C, K] hi/lo planes, or None when the queue is not an fp32 buffer (nothing to split)."""
        t = self._current(queue)
        return t if t.shape[0] == 2 * queue.shape[0] else None

    def full(self, queue):
        """The whole shadow buffer ([C, K] or [2C, K]) — what the enqueue kernels keep current."""
        return self._current(queue)


def queue_split(queue):
    """[2C, K] bf16 hi/lo planes of an fp32 ``queue`` [C, K] (rmcl_queue_split); a non-fp32 queue yields its plain
    [C, K] bf16 copy."""
    _need_cuda(queue)
    if queue.dim() != 2 or queue.stride(1) != 1:
        raise ValueError("queue must be [C,K] with K contiguous (reference layout)")
    if queue.dtype != torch.float32:
        return queue.detach().to(torch.bfloat16).contiguous()
    Cd, K = queue.shape
    out = torch.empty(2 * Cd, K, dtype=torch.bfloat16, device=queue.device)
    if K % 8 == 0 and queue.stride(0) % 4 == 0 and queue.data_ptr() % 16 == 0:
        rc = _lib.lib().rmcl_queue_split(_p(queue), Cd, K, queue.stride(0), _p(out), out.stride(0), _stream())
        check(rc, "rmcl_queue_split")
    else:       # ragged queue: shapes the tensor-core paths do not take anyway; same arithmetic in torch
        hi = queue.detach().to(torch.bfloat16)
        out[:Cd] = hi
        out[Cd:] = (queue.detach() - hi.float()).to(torch.bfloat16)
    return out


# queue storage -> its shadow: lets the fp32-accurate InfoNCE find the hi/lo planes of a queue it is handed, and lets
# every enqueue keep a shadow current even when the caller did not pass it (the kernels write through raw pointers, so
# torch's version counter would not notice the change).
_shadows = {}


def _register_shadow(queue, shadow):
    import weakref
    for key in [k for k, (wr, _) in _shadows.items() if wr() is None]:
        del _shadows[key]
    _shadows[queue.data_ptr()] = (weakref.ref(queue), shadow)


def _find_shadow(queue):
    ent = _shadows.get(queue.data_ptr())
    if ent is None:
        return None
    wr, shadow = ent
    alive = wr()
    if alive is None or alive.data_ptr() != queue.data_ptr() or alive.shape != queue.shape:
        del _shadows[queue.data_ptr()]
        return None
    return shadow


def hilo_of(queue, create=True):
    """The [2C, K] hi/lo planes of an fp32 ``queue`` from its registered shadow (created and registered on first use)."""
    shadow = _find_shadow(queue)
    if shadow is None:
        if not create:
            return None
        shadow = QueueShadow()
    return shadow.get_hilo(queue)


def enqueue_(queue, keys, ptr, shadow=None):
    """queue[:, ptr:ptr+B] = keys.T; ptr = (ptr+B) % K — on device, no host sync
    (objectives.py:244-248).  ``ptr`` is the int64[1] buffer ``proj_queue_ptr``.
    ``shadow`` (a :class:`QueueShadow`, or a bf16 [C,K] / hi-lo [2C,K] tensor) receives the same columns; a shadow
    registered for this queue (``hilo_of`` / ``QueueShadow.get``) is kept current even when not passed."""
    _need_cuda(queue, keys, ptr)
    if ptr.dtype != torch.int64 or ptr.numel() != 1:
        raise TypeError("ptr must be an int64 tensor with one element")
    if queue.stride(1) != 1 or keys.dim() != 2 or keys.shape[1] != queue.shape[0]:
        raise ValueError(f"shape mismatch: queue {tuple(queue.shape)} keys {tuple(keys.shape)}")
    keys = keys.detach().contiguous()
    registered = _find_shadow(queue)
    if shadow is None:
        shadow = registered
    elif registered is not None and registered is not shadow:
        registered._key = None        # another copy is being kept current instead: rebuild this one on its next use
    sh = None
    if shadow is not None:
        sh = shadow.full(queue) if isinstance(shadow, QueueShadow) else shadow
        _need_cuda(sh)
        if sh.dtype != torch.bfloat16 or sh.shape[1] != queue.shape[1] or sh.stride(1) != 1 or \
                sh.shape[0] not in (queue.shape[0], 2 * queue.shape[0]):
            raise ValueError("shadow must be a bf16 [C,K] or hi/lo [2C,K] tensor with K contiguous")
    if _lib.ffi() == "torch":
        _lib.torch_ops().enqueue_(queue, keys, ptr, sh)
        return
    if sh is not None:
        rc = _lib.lib().rmcl_enqueue_shadow(_p(queue), _dt(queue), _p(sh), sh.stride(0), sh.shape[0] // queue.shape[0], _p(keys),
                                            _dt(keys), _p(ptr), keys.shape[0], keys.shape[1], queue.shape[1], queue.stride(0),
                                            _stream())
        check(rc, "rmcl_enqueue_shadow")
        return
    rc = _lib.lib().rmcl_enqueue(_p(queue), _dt(queue), _p(keys), _dt(keys), _p(ptr), keys.shape[0], keys.shape[1],
                                 queue.shape[1], queue.stride(0), _stream())
    check(rc, "rmcl_enqueue")


# ------------------------------------------------------------------------------------ PGD
def pgd_step_(delta, grad, lr, eps, mode="ref_linf", _scratch={}):
    """In-place perturbation update (attack/pgd_attack_vilt.py:162-173 for ``ref_linf``)."""
    mode_arg = _lib.PGD_MODES[mode]
    _need_cuda(delta, grad)
    if delta.shape != grad.shape:
        raise ValueError("delta/grad shape mismatch")
    if not (delta.is_contiguous() and grad.is_contiguous()):
        raise ValueError("delta and grad must be contiguous")
    if _lib.ffi() == "torch":
        _lib.torch_ops().pgd_step_(delta, grad, float(lr), float(eps), mode_arg)
        return delta
    B = delta.shape[0]
    N = delta.numel() // B
    # the control words of the persistent kernel are self-resetting but not shareable: one workspace per stream
    key = (delta.device, B, N, grad.dtype, torch.cuda.current_stream().cuda_stream)
    ws = _scratch.get(key)
    if ws is None:
        nbytes = _lib.lib().rmcl_pgd_workspace_bytes(B, N, _dt(grad))
        if nbytes == 0:
            check(-1, "rmcl_pgd_workspace_bytes")
        ws = _scratch[key] = torch.zeros(nbytes + 256, dtype=torch.uint8, device=delta.device)   # zero-filled once
    off = (-ws.data_ptr()) % 256
    rc = _lib.lib().rmcl_pgd_step(_p(delta), _dt(delta), _p(grad), _dt(grad), B, N, float(lr), float(eps),
                                  mode_arg, C.c_void_p(ws.data_ptr() + off), ws.numel() - off, _stream())
    check(rc, "rmcl_pgd_step")
    return delta


# ------------------------------------------------------- host-buffer step (end-to-end call)
class HostStep:
    """One kernels-only RMCL step (EMA -> fused InfoNCE fwd+bwd -> enqueue) driven from HOST
    buffers through ``rmcl_step_host``: q/k projections come from pinned host memory, the loss
    and dq go back to pinned host memory; parameters, queue and pointer stay resident in HBM.
    This is the call the end-to-end number in bench.py is measured through."""

    def __init__(self, plan, queue, ptr, B, Cdim, temperature, m, qk_dtype=torch.bfloat16, path="auto"):
        _need_cuda(queue, ptr)
        if len(plan.groups) != 1:
            raise ValueError("HostStep expects a single-dtype EMA plan")
        dev = queue.device
        self.plan, self.queue, self.ptr = plan, queue, ptr
        self.B, self.C, self.K = B, Cdim, queue.shape[1]
        self.tau, self.m, self.path = float(temperature), float(m), _lib.INFONCE_PATHS[path]
        self.qk_dtype = qk_dtype
        self.q_dev = torch.empty(B, Cdim, dtype=qk_dtype, device=dev)
        self.k_dev = torch.empty(B, Cdim, dtype=qk_dtype, device=dev)
        self.loss_dev = torch.empty((), dtype=torch.float32, device=dev)
        self.dq_dev = torch.empty(B, Cdim, dtype=torch.float32, device=dev)
        self.k_hat_dev = torch.empty(B, Cdim, dtype=torch.float32, device=dev)
        self.loss_host = torch.empty((), dtype=torch.float32).pin_memory()
        self.dq_host = torch.empty(B, Cdim, dtype=torch.float32).pin_memory()
        self.ws, self.ws_off = _workspace(B, Cdim, self.K, _dt(queue), self.path, dev)
        self.h2d_bytes = 2 * B * Cdim * self.q_dev.element_size()
        self.d2h_bytes = 4 + B * Cdim * 4

    def __call__(self, q_host, k_host):
        """Blocks until loss_host / dq_host are valid; returns them."""
        if q_host.is_cuda or k_host.is_cuda or q_host.dtype != self.qk_dtype or k_host.dtype != self.qk_dtype:
            raise ValueError("HostStep takes host tensors of the configured dtype")
        dt, table, cnt, _ = self.plan.groups[0]
        rc = _lib.lib().rmcl_step_host(
            _p(table), cnt, self.m, dt, _p(q_host), _p(k_host), _DT[self.qk_dtype], _p(self.q_dev), _p(self.k_dev),
            _p(self.queue), _dt(self.queue), _p(self.ptr), self.B, self.C, self.K, self.tau, self.path,
            _p(self.loss_dev), _p(self.dq_dev), _p(self.k_hat_dev), _p(self.loss_host), _p(self.dq_host),
            C.c_void_p(self.ws.data_ptr() + self.ws_off), self.ws.numel() - self.ws_off, _stream())
        check(rc, "rmcl_step_host")
        return self.loss_host, self.dq_host


def infonce_launch_names(B, Cdim, K, queue_dtype, path="auto", need_grad=True):
    """Names of the kernels one InfoNCE call with these arguments launches (rmcl_infonce_describe)."""
    buf = C.create_string_buffer(512)
    qdt = _lib.RMCL_BF16_HILO if queue_dtype == "hilo" else _DT[queue_dtype]
    n = _lib.lib().rmcl_infonce_describe(B, Cdim, K, qdt, _lib.INFONCE_PATHS[path], 1 if need_grad else 0, buf, 512)
    if n < 0:
        check(n, "rmcl_infonce_describe")
    return tuple(buf.value.decode().split(","))


def tc_timeline(fn):
    """Run ``fn()`` (which must launch the tcgen05 InfoNCE kernel) with the in-kernel timeline of
    CTA (0,0) switched on; returns the raw int64 clock stamps (layout: csrc/infonce_tc.cu)."""
    L = _lib.lib()
    buf = torch.zeros(L.rmcl_debug_tc_timeline_words(), dtype=torch.int64, device="cuda")
    check(L.rmcl_debug_tc_timeline(C.c_void_p(buf.data_ptr())), "rmcl_debug_tc_timeline")
    try:
        fn()
        torch.cuda.synchronize()
    finally:
        L.rmcl_debug_tc_timeline(C.c_void_p(0))
    return buf.cpu()


def profile_enable(on=True):
    check(_lib.lib().rmcl_profile_enable(1 if on else 0), "rmcl_profile_enable")


def profile_infonce_ms():
    out = (C.c_float * 3)()
    check(_lib.lib().rmcl_profile_infonce_ms(out), "rmcl_profile_infonce_ms")
    return {"prep": out[0], "partial": out[1], "finalize": out[2]}
