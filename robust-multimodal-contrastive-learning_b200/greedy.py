"""InfoNCE call sites of the greedy text attack (attack/greedy_attack_vilt.py) on the fused kernels.

The attack itself — tokenisation, synonym/embedding candidate search, word substitution — is a CPU
tokenizer loop and stays with the reference (its call site is kept in
``objectives.compute_geometric``).  What it spends GPU time on are two InfoNCE evaluations:

* ``get_grad`` (greedy_attack_vilt.py:437-445): loss + backward of the current text view — this is
  ``ops.infonce_loss`` (autograd) as is.
* ``split_forward`` (461-484): for every sample i and every candidate sentence j the reference
  overwrites row i of the batch of representations with the candidate's, recomputes the **whole**
  B x (K+1) logits and the batch-mean cross-entropy, and compares it with the running maximum on the
  host: sum_i n_i full InfoNCE evaluations plus as many device->host syncs.  Only row i changes, so
      loss_ij = ori_loss + (L(cand_ij, k_i) - L(ori_i, k_i)) / B
  with L the per-row loss.  ``split_forward_losses`` gets every L in TWO fused passes (originals,
  all candidates at once — ``loss_per_row`` of rmcl_infonce_fwd_bwd) and one host transfer.
"""
import torch

from . import ops


def infonce_rowwise_loss(q, k, queue, temperature, path="auto"):
    """Per-row InfoNCE loss ``lse_i - pos_i`` of (already normalised or raw) ``q`` against
    [``k`` ; ``queue``]; no gradient."""
    return ops.infonce_fwd_bwd(q.float(), k.float(), queue, temperature, need_grad=False, path=path,
                               want=("loss_per_row",))["loss_per_row"]


@torch.no_grad()
def split_forward_losses(ori_z, cand_z, all_num, k_modality, queue, temperature, path="auto"):
    """greedy_attack_vilt.py:461-484 without the per-candidate full recomputation.

    ori_z       [B,C]  normalised representations of the current sentences
    cand_z      [sum(all_num),C]  normalised representations of all candidate sentences, sample-major
                (what ``torch.split(q_txt_attack, all_num)`` splits)
    all_num     candidates per sample
    Returns the reference's ``all_loss``: per sample ``(list of candidate losses, index of the first
    candidate that beats every earlier one and the original loss, or -1)``; losses are 0-dim tensors
    on the device of ``ori_z``.
    """
    B = ori_z.shape[0]
    if len(all_num) != B or sum(all_num) != cand_z.shape[0]:
        raise ValueError("all_num must list the number of candidates of every sample")
    row_ori = infonce_rowwise_loss(ori_z, k_modality, queue, temperature, path)            # [B]
    ori_loss = row_ori.mean()
    owner = torch.repeat_interleave(torch.arange(B, device=ori_z.device),
                                    torch.as_tensor(all_num, device=ori_z.device))
    if cand_z.shape[0] == 0:
        return [([], -1) for _ in range(B)]
    row_cand = infonce_rowwise_loss(cand_z, k_modality[owner], queue, temperature, path)   # [sum n_i]
    loss_cand = ori_loss + (row_cand - row_ori[owner]) / B
    host = loss_cand.tolist()                                                               # the one sync
    ori_host = float(ori_loss)
    all_loss, start = [], 0
    for n in all_num:
        cur_max, cur_idx = ori_host, -1
        for j in range(n):
            if host[start + j] > cur_max:
                cur_max, cur_idx = host[start + j], j
        all_loss.append(([loss_cand[start + j] for j in range(n)], cur_idx))
        start += n
    return all_loss
