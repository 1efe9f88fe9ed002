"""TEST INFRASTRUCTURE — CPU restatement of the reference's RMCL hot path.

This file is the *checker*, never the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it.  The product (``rmcl_b200``) never does and has no CPU
path at all.

Reference: stanFurrer/Robust-Multimodal-Contrastive-Learning (pure Python/torch,
no tests, no golden vectors of its own ⇒ the reference's own test-suite pins
nothing).  Parity is pinned instead by *executing the unmodified reference* in the
authoring container (``oracle/ref_harness.py`` + ``oracle/make_golden.py``) and
committing its inputs/outputs under ``tests/golden/``; ``tests/test_oracle.py``
checks every function below against those vectors, and
``tests/test_oracle_vs_reference.py`` re-runs the reference live when
``/root/reference`` is present.

Every function cites the reference lines it restates (paths relative to the
reference root).  Arithmetic is torch-on-CPU in the dtype of the inputs so that
fp32 results are bit-comparable with the reference's ATen expressions;
``dtype=torch.float64`` gives the high-precision ground truth used for the bf16
tolerance checks.

The three PGD modes: ``ref_linf`` restates attack/pgd_attack_vilt.py:162-173.
``sign_linf`` and ``l2`` are the north-star's extensions (BASELINE.json); they have
no reference lines and are specified here:
  sign_linf:  d <- clamp(d + lr*sign(g), -eps, eps)
  l2:         d <- d + lr*g/max(||g_b||_2,1e-8);  d <- d*min(1, eps/max(||d_b||_2,1e-12))
"""
import math

import torch
import torch.nn.functional as F

__all__ = [
    "momentum_update", "l2_normalize", "info_nce", "info_nce_logits",
    "concat_all_gather", "dequeue_and_enqueue", "pgd_update", "queue_diagnostics",
    "rmcl_kernel_step", "greedy_split_forward", "barlow_twins",
]


# --------------------------------------------------------------------------- EMA
def momentum_update(params_k, params_q, em):
    """k <- k*em + q*(1-em), tensor by tensor, two roundings then an add.

    vilt/modules/objectives.py:219-224 (called x4 at 257-260);
    MoCo/MoCo_RMCL.py:65-72.  Returns new tensors (the reference rebinds .data).
    NB ``1.0 - em`` is evaluated in Python double precision and only then rounded
    to the tensor dtype by ATen's scalar multiply.
    """
    return [pk * em + pq * (1.0 - em) for pk, pq in zip(params_k, params_q)]


# ----------------------------------------------------------------------- InfoNCE
def l2_normalize(x, eps=1e-12):
    """F.normalize(x, dim=1): x / max(||x||_2, eps).  objectives.py:265,269,326."""
    return x / x.norm(p=2, dim=1, keepdim=True).clamp_min(eps)


def info_nce_logits(q_hat, k_hat, queue, temperature):
    """[q.k , q.queue] / T   — objectives.py:271-274, 328-331;
    attack/pgd_attack_vilt.py:152-155.  queue is [C, K] (K contiguous)."""
    l_pos = torch.einsum("nc,nc->n", [q_hat, k_hat]).unsqueeze(-1)
    l_neg = torch.einsum("nc,ck->nk", [q_hat, queue])
    return torch.cat([l_pos, l_neg], dim=1) / temperature


def info_nce(q_raw, k_hat, queue, temperature, loss_div=1.0, grad_out=1.0):
    """Full InfoNCE call site as the reference runs it.

    q_raw     [B,C]  projection head output *before* F.normalize (objectives.py:325-326)
    k_hat     [B,C]  already-normalised key (objectives.py:265), no grad
    queue     [C,K]  detached clone of proj_queue (objectives.py:270)
    loss      CrossEntropyLoss(mean) against label 0 (objectives.py:333-334,351),
              divided by ``loss_div`` (= adv_steps_img inside PGD,
              attack/pgd_attack_vilt.py:158)
    Returns logits, loss, per-row loss, lse, argmax (objectives.py:275,336), the
    positive logit, and d(loss*grad_out)/d q_raw via autograd (what
    ``loss.backward()`` hands to the projection head).
    """
    q_raw = q_raw.detach().clone().requires_grad_(True)
    q_hat = F.normalize(q_raw, dim=1)
    logits = info_nce_logits(q_hat, k_hat.detach(), queue.detach(), temperature)
    labels = torch.zeros(logits.shape[0], dtype=torch.long)
    lf = logits.float() if logits.dtype != torch.float64 else logits
    loss = F.cross_entropy(lf, labels) / (1.0 * loss_div)
    (loss * grad_out).backward()
    with torch.no_grad():
        lse = torch.logsumexp(lf, dim=1)
        return {
            "q_hat": q_hat.detach(),
            "logits": logits.detach(),
            "loss": loss.detach(),
            "loss_per_row": (lse - lf[:, 0]),
            "lse": lse,
            "argmax": logits.argmax(-1),
            "pos": lf[:, 0].clone(),
            "dq": q_raw.grad.detach(),
        }


def queue_diagnostics(q_hat, k_hat, queue, cosine_eps=1e-6):
    """pos/neg L2, cosine and dot diagnostics — objectives.py:337-349 (loops F14).

    Restated without the per-sample Python loop; same means.
    """
    neg = queue.T  # [K, C]
    out = {
        "pos_dist": torch.linalg.norm(q_hat - k_hat, dim=1).mean(),
        "pos_cosine": F.cosine_similarity(q_hat, k_hat, dim=1, eps=cosine_eps).mean(),
        "pos_dot": torch.sum(q_hat * k_hat, dim=1).mean(),
    }
    dist = cosine = dot = 0.0
    for sub_q in q_hat:
        dist = dist + torch.linalg.norm(sub_q - neg, dim=1).mean()
        cosine = cosine + F.cosine_similarity(sub_q.unsqueeze(0), neg, dim=1, eps=cosine_eps).mean()
        dot = dot + torch.sum(sub_q.unsqueeze(0) * neg, dim=1).mean()
    n = q_hat.shape[0]
    out.update(neg_dist=dist / n, neg_cosine=cosine / n, neg_dot=dot / n)
    return out


# ------------------------------------------------------------- gather + enqueue
def concat_all_gather(per_rank_tensors):
    """all_gather along dim 0 in rank order — objectives.py:226-235;
    MoCo/MoCo_RMCL.py:268-279.  Takes the list of every rank's local tensor."""
    return torch.cat(list(per_rank_tensors), dim=0)


def dequeue_and_enqueue(queue, ptr, keys, num_negative, per_step_bs=None):
    """queue[:, ptr:ptr+B] = keys.T ; ptr = (ptr+B) % K — objectives.py:238-248.

    ``keys`` is the already-gathered [B_global, C] batch.  Mirrors the silent skip
    when B_global != per_step_bs (242-243).  Like the reference there is no
    wrap-around split: a slice that runs past K is a shape error in torch, which
    this restatement surfaces as ValueError.  Returns (new_queue, new_ptr).
    """
    batch_size = keys.shape[0]
    if per_step_bs is not None and per_step_bs != batch_size:
        return queue, int(ptr)
    ptr = int(ptr)
    if ptr + batch_size > queue.shape[1]:
        raise ValueError("enqueue slice runs past the end of the queue (reference would fail too)")
    queue = queue.clone()
    queue[:, ptr:ptr + batch_size] = keys.T.to(queue.dtype)
    ptr = (ptr + batch_size) % num_negative
    return queue, ptr


# ------------------------------------------------------------------------- PGD
def pgd_update(delta, grad, lr, eps, mode="ref_linf"):
    """One perturbation update on [B, ...] tensors.

    ref_linf — attack/pgd_attack_vilt.py:162-173:
        g = grad.float(); d = clamp(max_b|g|, min=1e-8);
        delta = delta + (lr*g/d).to(delta); clamp(+-eps) only if eps > 0.
    sign_linf / l2 — see module docstring.
    """
    B = grad.shape[0]
    bshape = (B,) + (1,) * (grad.dim() - 1)
    if mode == "ref_linf":
        g = grad.clone().detach().float() if grad.dtype != torch.float64 else grad.clone()
        denorm = torch.norm(g.view(B, -1), dim=1, p=float("inf")).view(bshape)
        denorm = torch.clamp(denorm, min=1e-8)
        step = (lr * g / denorm).to(delta)
        out = (delta + step).detach()
        if eps > 0:
            out = torch.clamp(out, -eps, eps).detach()
        return out
    if mode == "sign_linf":
        out = delta + lr * torch.sign(grad).to(delta)
        if eps > 0:
            out = torch.clamp(out, -eps, eps)
        return out
    if mode == "l2":
        g = grad.float() if grad.dtype != torch.float64 else grad
        gn = torch.norm(g.view(B, -1), dim=1, p=2).view(bshape).clamp(min=1e-8)
        out = delta + (lr * g / gn).to(delta)
        if eps > 0:
            dn = torch.norm(out.view(B, -1).to(g.dtype), dim=1, p=2).view(bshape).clamp(min=1e-12)
            out = out * torch.clamp(eps / dn, max=1.0).to(out)
        return out
    raise ValueError(mode)


# -------------------------------------------------- kernels-only step (cfg2)
def rmcl_kernel_step(params_k, params_q, em, q_raw, k_hat, queue, ptr, temperature,
                     gathered_keys=None):
    """EMA → InfoNCE fwd+bwd → enqueue: the order of objectives.py:257-260,
    324-351, 394-395 with the backbone forwards removed (BASELINE cfg2).

    Returns (new_params_k, infonce dict, new_queue, new_ptr).
    """
    new_k = momentum_update(params_k, params_q, em)
    res = info_nce(q_raw, k_hat, queue, temperature)
    keys = k_hat if gathered_keys is None else gathered_keys
    new_queue, new_ptr = dequeue_and_enqueue(queue, ptr, keys, queue.shape[1])
    return new_k, res, new_queue, new_ptr


def greedy_split_forward(ori_z, cand_z, all_num, k_modality, queue, temperature):
    """attack/greedy_attack_vilt.py:461-484 restated: for every sample i and candidate j put the
    candidate's representation in row i, recompute the whole batch-mean InfoNCE loss, track the first
    candidate that beats everything before it.  Returns [(losses, best_idx)] like the reference."""
    def batch_loss(z):
        logits = info_nce_logits(z, k_modality, queue, temperature)
        return F.cross_entropy(logits.float() if logits.dtype != torch.float64 else logits,
                               torch.zeros(z.shape[0], dtype=torch.long))
    ori_z = ori_z.clone()
    ori_loss = batch_loss(ori_z)
    out = []
    for i, cls_split in enumerate(torch.split(cand_z, all_num)):
        cur, cur_max, cur_idx = [], ori_loss, -1
        saved = ori_z[i].clone()
        for j, cls in enumerate(cls_split):
            ori_z[i] = cls
            loss = batch_loss(ori_z)
            cur.append(loss)
            if loss > cur_max:
                cur_max, cur_idx = loss, j
        out.append((cur, cur_idx))
        ori_z[i] = saved
    return out


def closed_form_single_negative(s_pos, s_neg, temperature):
    """K=1 known answer: loss = log(1 + exp((s_neg - s_pos)/T)).  SURVEY §8(c)."""
    return math.log1p(math.exp((s_neg - s_pos) / temperature))


# ------------------------------------------------------------------ Barlow Twins
def _off_diagonal(x):
    """objectives.py:454-458: flattened view of the off-diagonal elements of a square matrix."""
    n, m = x.shape
    assert n == m
    return x.flatten()[:-1].view(n - 1, n + 1)[:, 1:].flatten()


def barlow_twins(q_per_rank, k_per_rank, per_step_bs, lam, grad_on=1.0, grad_offs=1.0):
    """Barlow-Twins cross-correlation loss of one view, restated from
    vilt/modules/objectives.py:480-486 (text view; 506-512 image view, 533-539 both) and, with one rank and
    ``per_step_bs = B``, attack/pgd_attack_vilt.py:219-224:

        c = q.T @ k;  c.div_(per_step_bs);  all_reduce(c)          # sum over ranks
        on_diag  = diagonal(c).add_(-1).pow_(2).sum()
        off_diag = off_diagonal(c).pow_(2).sum()
        loss     = on_diag + lam * off_diag

    ``q_per_rank`` / ``k_per_rank``: lists of [B_local, D] tensors (one rank: lists of length 1).  The all-reduce
    is not autograd-aware: every rank's backward treats the summed ``c`` as if it were its own product, so
    ``dq_r = k_r @ dC.T / per_step_bs`` with ``dC`` evaluated on the reduced matrix.
    Returns on_diag, off_diag, loss, c and the per-rank gradients of
    ``grad_on * on_diag + grad_offs * lam * off_diag`` with respect to each rank's q.
    """
    c = None
    for q, k in zip(q_per_rank, k_per_rank):
        cr = q.T @ k
        cr = cr / per_step_bs
        c = cr if c is None else c + cr
    eye = torch.eye(c.shape[0], dtype=c.dtype)
    on_diag = (torch.diagonal(c) - 1).pow(2).sum()
    off_diag = _off_diagonal(c).pow(2).sum()
    w = grad_offs * lam * (1 - eye) + grad_on * eye
    dC = 2.0 * w * (c - eye)
    dq = [(k @ dC.T) / per_step_bs for k in k_per_rank]
    return {"on_diag": on_diag, "off_diag": off_diag, "loss": on_diag + lam * off_diag, "c": c, "dq": dq}
