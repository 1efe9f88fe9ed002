"""``compute_moco_contrastive`` / ``compute_pgd`` with the reference's signatures, on fused kernels.

Drop-in for vilt/modules/objectives.py:160-188 and 217-447: same arguments
(``pl_module``, ``batch``), same attributes read from ``pl_module``, same return
keys (exactly one key containing "loss": ``moco_loss``) and the same ``pl_module.log`` names, so
``vilt.modules.objectives.compute_moco_contrastive = rmcl_b200.compute_moco_contrastive`` is the
whole integration.  The backbone forwards (``infer``/``infer_k``/heads) stay on
torch; what changes is everything between them:

  EMA        161-tensor python loop            -> ops.ema_multi_   (1 launch)
  InfoNCE    normalize/clone/einsum/cat/CE     -> ops.infonce_*    (logits never materialised)
  PGD        7-kernel update, deepcopy         -> pgd_attack.PGDAttack_moco / ops.pgd_step_
  enqueue    int(ptr) sync + strided copy      -> ops.enqueue_     (pointer stays on device)
  gather     list all_gather + cat             -> dist.concat_all_gather
"""
import warnings
from copy import deepcopy

import torch

from . import dist as rdist
from . import ops

_KEY_QUERY_LAYERS = (  # objectives.py:257-260, in this order
    ("text_embeddings", "k_text_embeddings"),
    ("token_type_embeddings", "k_token_type_embeddings"),
    ("transformer", "k_transformer"),
    ("moco_head", "k_moco_head"),
)


def shadow_layer(q_layer, k_layer):
    """vilt_module.py:270-273 (``_shadow_layer``): initialise a key layer from its query layer and freeze it."""
    for param_q, param_k in zip(q_layer.parameters(), k_layer.parameters()):
        param_k.data.copy_(param_q.data)
        param_k.requires_grad = False


def momentum_update_key_encoder(pl_module):
    """objectives.py:219-224 + 257-260 as one launch; the chunk table is cached on the module."""
    plan = pl_module.__dict__.get("_rmcl_ema_plan")
    if plan is None:
        pk, pq = [], []
        for qn, kn in _KEY_QUERY_LAYERS:
            q_layer, k_layer = getattr(pl_module, qn), getattr(pl_module, kn)
            for param_q, param_k in zip(q_layer.parameters(), k_layer.parameters()):
                pq.append(param_q)
                pk.append(param_k)
        plan = ops.EmaPlan(pk, pq)
        pl_module.__dict__["_rmcl_ema_plan"] = plan
    ops.ema_multi_(plan, pl_module.momentum)


def _queue_shadow(pl_module):
    """The module's bf16 queue shadow (ops.QueueShadow), or None for the fp32 buffer.

    ``pl_module.rmcl_bf16_queue``: True / False decide explicitly; unset (None) follows the precision the step runs in —
    under autocast (Lightning ``precision=16``, how the reference trains: its einsum then runs in half precision with an
    fp32 cross-entropy, objectives.py:272-274) the shadow and hence the tcgen05 kernels are used, in full precision the
    fp32 buffer and the fp32 parity kernel."""
    opt = getattr(pl_module, "rmcl_bf16_queue", None)
    if opt is None:
        opt = torch.is_autocast_enabled()
        if opt and not pl_module.__dict__.get("_rmcl_warned_bf16"):
            pl_module.__dict__["_rmcl_warned_bf16"] = True
            warnings.warn("rmcl_b200: autocast is on and pl_module.rmcl_bf16_queue is unset -> the main-step InfoNCE reads a "
                          "bf16 shadow of proj_queue with bf16 q^ (8 mantissa bits, fp32 accumulation; the reference's "
                          "precision=16 einsum is fp16, 11 bits). Set pl_module.rmcl_bf16_queue = True to silence this, "
                          "or False to keep the fp32 buffer and the fp32-accurate kernels.", stacklevel=3)
    if not opt:
        return None
    sh = pl_module.__dict__.get("_rmcl_queue_shadow")
    if sh is None:    # adopt the copy an earlier fp32-accurate call may have registered for this buffer
        sh = pl_module.__dict__["_rmcl_queue_shadow"] = ops._find_shadow(pl_module.proj_queue) or ops.QueueShadow()
    return sh


def _enqueue_shadow(pl_module):
    """The shadow the enqueue must keep current: an existing one always (whatever precision this particular call runs
    in — a stale shadow would go unnoticed, the kernels write through raw pointers), else what _queue_shadow decides."""
    sh = pl_module.__dict__.get("_rmcl_queue_shadow")
    return sh if sh is not None else _queue_shadow(pl_module)


def infonce_queue(pl_module):
    """The queue operand of the main-step InfoNCE calls: the fp32 buffer, or the bf16 plane of its shadow when the module
    opted in / runs under autocast.  An fp32 operand is evaluated at fp32 accuracy either way: by the split-operand tcgen05
    kernels on the shadow's hi/lo planes (C in {64,128,256}; ``infonce_path="auto"``) or by the CUDA-core kernel with exact
    fp32 products (``infonce_path="simt"``).  The PGD inner loss always gets the fp32 buffer: the reference runs it under
    ``autocast(False)`` (pgd_attack_vilt.py:141)."""
    sh = _queue_shadow(pl_module)
    return pl_module.proj_queue if sh is None else sh.get(pl_module.proj_queue)


def dequeue_and_enqueue(pl_module, keys):
    """objectives.py:238-248 — gather, skip on a short batch, ring-buffer write, pointer advance.

    ``pl_module.rmcl_p2p_exchange = True`` (single node, world > 1) replaces ncclAllGather + enqueue by the fused
    peer-memory kernel (dist.P2PKeyExchange): every rank pushes its keys straight into the peers' staging slots over
    NVLink, signals, waits and enqueues in one launch.  The short-batch skip is decided from the local batch size
    (every rank sees the same ``per_step_bs`` and the same local batch under DistributedSampler with drop_last)."""
    world = rdist.dist.get_world_size() if (rdist.dist.is_available() and rdist.dist.is_initialized()) else 1
    if world > 1 and getattr(pl_module, "rmcl_p2p_exchange", False):
        if not rdist.gathered_batch_matches(pl_module.per_step_bs, world * keys.shape[0]):
            return
        ex = pl_module.__dict__.get("_rmcl_p2p")
        if ex is None or (ex.B, ex.C) != tuple(keys.shape):
            ex = pl_module.__dict__["_rmcl_p2p"] = rdist.P2PKeyExchange(keys.shape[0], keys.shape[1], keys.device)
        ex.enqueue_(pl_module.proj_queue, keys.detach().float().contiguous(), pl_module.proj_queue_ptr,
                    shadow=_enqueue_shadow(pl_module))
        return
    keys = rdist.concat_all_gather(keys)
    if not rdist.gathered_batch_matches(pl_module.per_step_bs, keys.shape[0]):
        return
    ops.enqueue_(pl_module.proj_queue, keys, pl_module.proj_queue_ptr, shadow=_enqueue_shadow(pl_module))


def compute_pgd(pl_module, batch, loss_name, k_modality=None):
    """objectives.py:160-188: run the attacker, add delta to the image(s), log |delta|.  The ``nlvr2_attacked``
    branch (two images, a pair of perturbations, 167-174 and 180-183) is kept; the embedding-space attacker
    (an extension) leaves the image alone and stores delta, the embeddings it was computed against and their masks."""
    attacker = pl_module.pgd_attacker
    img_delta = attacker.pgd_attack(pl_module, batch, k_modality=k_modality)
    phase = "train" if pl_module.training else "val"
    if loss_name == "nlvr2_attacked":
        batch["image_0"][0] = batch["image_0"][0] + img_delta[0]
        batch["image_1"][0] = batch["image_1"][0] + img_delta[1]
        delta_range = (torch.linalg.norm(img_delta[0], dim=1).mean() + torch.linalg.norm(img_delta[1], dim=1).mean()) \
            / sum(pl_module.attack_idx)
    else:
        if getattr(attacker, "space", "pixel") != "pixel":
            d_txt, d_img = attacker.split_delta(img_delta)
            batch["image_embeds_delta"], batch["text_embeds_delta"] = d_img, d_txt
            batch["image_embeds_base"], batch["image_masks_base"] = attacker.embed_base, attacker.embed_masks
        else:
            # NB the attacker already left img_init+delta_{n-1} in batch; the reference
            # adds delta_n on top and so do we.
            batch["image"][0] = batch["image"][0] + img_delta
        delta_range = torch.linalg.norm(img_delta, dim=1).mean()
    pl_module.log(f"{loss_name}_attack/{phase}/delta", delta_range)
    return batch


def _attacked_view(pl_module, batch, k_hat, prediction_original, suffix, rate_name, ret, stats):
    """One attacked/augmented view: forward, then ONE fused launch chain for the InfoNCE loss, its
    gradient, the row argmax and (``stats``: ops.QueueStats) the six pos/neg diagnostics."""
    if "image_embeds_delta" in batch:
        # embedding-space PGD (extension): the perturbation belongs to the embeddings and masks the attacker saw —
        # visual_embed samples patches randomly, so it must not be run again here (vilt_module.py:294-304)
        from .pgd_attack import text_embeds_delta
        with text_embeds_delta(pl_module.text_embeddings, batch.get("text_embeds_delta")):
            infer = pl_module.infer(batch, mask_text=False, mask_image=False,
                                    image_embeds=batch["image_embeds_base"] + batch["image_embeds_delta"],
                                    image_masks=batch["image_masks_base"])
    else:
        infer = pl_module.infer(batch, mask_text=False, mask_image=False)
    q_raw = pl_module.moco_head(infer["cls_feats"])
    out = ops.infonce_loss(q_raw, k_hat, infonce_queue(pl_module), pl_module.temperature,
                           getattr(pl_module, "infonce_path", "auto"), diag=stats)
    loss, argmax = out[0], out[1]
    if pl_module.training:
        pl_module.log(f"moco_attack/{rate_name}_success_rate",
                      (~(argmax == prediction_original)).sum() / argmax.shape[0])
    if stats is not None:
        for i, name in enumerate(ops.DIAG_NAMES):
            ret[f"{name}_attacked_{suffix}"] = out[2][i]
    pl_module.log(f"moco_loss/attacked_{suffix}_loss", loss)
    return loss


def compute_moco_contrastive(pl_module, batch, diagnostics=True):
    ret = {}
    phase = "train" if pl_module.training else "val"
    loss, loss_num = 0, 0

    momentum_update_key_encoder(pl_module)

    with torch.no_grad():
        infer_k = pl_module.infer_k(batch, mask_text=False, mask_image=False)
        k_raw = pl_module.k_moco_head(infer_k["cls_feats"])

    # clean query: only the row argmax is used (objectives.py:267-275); the same launch
    # normalises the key, so k^ comes out of it for PGD, the losses and the enqueue.
    infer = pl_module.infer(batch, mask_text=False, mask_image=False)
    q_clean = pl_module.moco_head(infer["cls_feats"])
    clean = ops.infonce_fwd_bwd(q_clean.float(), k_raw.float(), infonce_queue(pl_module), pl_module.temperature,
                                normalize_k=True, need_grad=False, path=getattr(pl_module, "infonce_path", "auto"),
                                want=("argmax", "k_hat"))
    prediction_original, k = clean["argmax"], clean["k_hat"]
    # |queue_j|^2 and the two [C] sums behind the neg_* diagnostics: one reduction of the queue per step,
    # shared by every view (the reference re-reduces the queue 3*B times per view, objectives.py:341-346)
    stats = ops.QueueStats(pl_module.proj_queue, cos_eps=getattr(pl_module.cosine, "eps", 1e-6)) if diagnostics else None

    attacked_words = None
    if pl_module.text_view:
        if pl_module.augmentation:
            augmented_batch = pl_module.text_augmentation_fn(pl_module, deepcopy(batch))
        else:
            augmented_batch = compute_geometric(pl_module, deepcopy(batch), "moco", k_modality=k)
            attacked_words = {n: deepcopy(augmented_batch[n]) for n in ("text", "text_ids", "text_masks")}
        loss = loss + _attacked_view(pl_module, augmented_batch, k, prediction_original, "txt", "Geom", ret, stats)
        loss_num += 1

    if pl_module.image_view:
        if pl_module.augmentation:
            augmented_batch = pl_module.image_augmentation_fn(pl_module, deepcopy(batch))
        else:
            augmented_batch = compute_pgd(pl_module, deepcopy(batch), "moco", k_modality=k)
        loss = loss + _attacked_view(pl_module, augmented_batch, k, prediction_original, "img", "PGD", ret, stats)
        loss_num += 1

    if pl_module.image_view and pl_module.text_view and not pl_module.augmentation:
        for n in ("text", "text_ids", "text_masks"):
            augmented_batch[n] = attacked_words[n]
        loss = loss + _attacked_view(pl_module, augmented_batch, k, prediction_original, "both", "Both", ret, stats)
        loss_num += 1

    if pl_module.training:
        dequeue_and_enqueue(pl_module, k)

    ret["moco_loss"] = loss / loss_num
    metric = getattr(pl_module, f"{phase}_moco_loss")(ret["moco_loss"])
    pl_module.log(f"moco_loss/step/{phase}", metric)
    if diagnostics:
        for view, on in (("img", pl_module.image_view), ("txt", pl_module.text_view),
                         ("both", pl_module.image_view and pl_module.text_view and not pl_module.augmentation)):
            if not on:
                continue
            for tag, key in (("L2", "dist"), ("Cosine", "cosine"), ("Dot", "dot")):
                pos, neg = ret[f"pos_{key}_attacked_{view}"], ret[f"neg_{key}_attacked_{view}"]
                pl_module.log(f"moco_dist_{phase}_{tag}/Pos_attacked_{view}", pos)
                pl_module.log(f"moco_dist_{phase}_{tag}/Neg_attacked_{view}", neg)
                pl_module.log(f"moco_dist_{phase}_{tag}/Neg-Pos_attacked_{view}", neg - pos)
    return ret


def compute_geometric(pl_module, batch, loss_name, k_modality=None):
    """objectives.py:190-215: delegate to the module's greedy text attacker (out of scope here —
    a CPU tokenizer loop — but its call site is kept so ``text_view=True`` modules still run)."""
    attack_words = pl_module.greedy_attacker.adv_attack_samples(pl_module, batch, k_modality)
    batch["text"] = attack_words["text"]
    batch["text_ids"] = attack_words["txt_input_ids"]
    batch["text_masks"] = attack_words["text_masks"]
    phase = "train" if pl_module.training else "val"
    pl_module.log(f"{loss_name}_attack/{phase}/num_changes", attack_words["num_changes"])
    pl_module.log(f"{loss_name}_attack/{phase}/change_rate", attack_words["change_rate"])
    return batch


# ------------------------------------------------------------------------------- Barlow Twins
def _barlow_view(pl_module, batch, k, suffix, ret, gather):
    """One attacked/augmented view of objectives.py:476-498 / 502-523 / 529-551: forward, the fused
    cross-correlation loss (c is never materialised, the all-reduce of c becomes an all-gather of the two
    [B, D] projections) and the three positive-pair diagnostics."""
    infer = pl_module.infer(batch, mask_text=False, mask_image=False)
    q = pl_module.barlowtwins_head(infer["cls_feats"])
    on_diag, off_scaled = ops.barlow_twins_loss(q, k, 1.0 / float(pl_module.per_step_bs), float(pl_module.adv_lr), gather)
    ret[f"barlowtwins_loss_invariance_{suffix}"] = on_diag
    ret[f"barlowtwins_loss_redundancy_{suffix}"] = off_scaled
    tag = {"text": "txt"}.get(suffix, suffix)
    ret[f"pos_dist_attacked_{tag}"] = torch.linalg.norm(q - k, dim=1).mean()
    ret[f"pos_cosine_attacked_{tag}"] = pl_module.cosine(q, k).mean()
    ret[f"pos_dot_attacked_{tag}"] = torch.sum(q * k, dim=1).mean()
    return on_diag + off_scaled


def compute_barlowtwins_contrastive(pl_module, batch):
    """Drop-in for vilt/modules/objectives.py:449-602: same arguments, return keys (``barlowtwins_loss`` plus the
    per-view ``barlowtwins_loss_invariance_*`` / ``barlowtwins_loss_redundancy_*`` — all three contain "loss"
    and are summed by training_step, vilt_module.py:475, exactly as in the reference — and ``pos_*_attacked_*``)
    and log names.  Per view the reference builds the 8192 x 8192 matrix ``c = q.T @ k / bs``, all-reduces it
    and differentiates through it; here one fused tcgen05 pass (ops.barlow_twins_loss) emits both sums and dq."""
    ret = {}
    loss, loss_num = 0, 0
    gather = rdist.Gather()

    with torch.no_grad():
        infer = pl_module.infer(batch, mask_text=False, mask_image=False)
        k = pl_module.barlowtwins_head(infer["cls_feats"])

    attacked_words = None
    if pl_module.text_view:
        if pl_module.augmentation:
            augmented_batch = pl_module.text_augmentation_fn(pl_module, deepcopy(batch))
        else:
            augmented_batch = compute_geometric(pl_module, deepcopy(batch), "barlowtwins", k_modality=k)
            attacked_words = {n: deepcopy(augmented_batch[n]) for n in ("text", "text_ids", "text_masks")}
        loss = loss + _barlow_view(pl_module, augmented_batch, k, "text", ret, gather)
        loss_num += 1

    if pl_module.image_view:
        if pl_module.augmentation:
            augmented_batch = pl_module.image_augmentation_fn(pl_module, deepcopy(batch))
        else:
            augmented_batch = compute_pgd(pl_module, deepcopy(batch), "barlowtwins", k_modality=k)
        loss = loss + _barlow_view(pl_module, augmented_batch, k, "img", ret, gather)
        loss_num += 1

    if pl_module.image_view and pl_module.text_view and not pl_module.augmentation:
        for n in ("text", "text_ids", "text_masks"):
            augmented_batch[n] = attacked_words[n]
        loss = loss + _barlow_view(pl_module, augmented_batch, k, "both", ret, gather)
        loss_num += 1

    ret["barlowtwins_loss"] = loss / loss_num

    phase = "train" if pl_module.training else "val"
    metric = getattr(pl_module, f"{phase}_barlowtwins_loss")(ret["barlowtwins_loss"])
    pl_module.log(f"barlowtwins/{phase}/loss", metric)
    for view, tag, on in (("img", "img", pl_module.image_view), ("text", "txt", pl_module.text_view),
                          ("both", "both", pl_module.image_view and pl_module.text_view and not pl_module.augmentation)):
        if not on:
            continue
        pl_module.log(f"barlowtwins_dist_{phase}_L2/Pos_attacked_{tag}", ret[f"pos_dist_attacked_{tag}"])
        pl_module.log(f"barlowtwins_dist_{phase}_Cosine/Pos_attacked_{tag}", ret[f"pos_cosine_attacked_{tag}"])
        pl_module.log(f"barlowtwins_dist_{phase}_Dot/Pos_attacked_{tag}", ret[f"pos_dot_attacked_{tag}"])
        for part in ("invariance", "redundancy"):
            m = getattr(pl_module, f"{phase}_barlowtwins_loss_{part}_{view}")(ret[f"barlowtwins_loss_{part}_{view}"])
            pl_module.log(f"barlowtwins/{phase}/barlowtwins_loss_{part}_{view}", m)
    return ret
