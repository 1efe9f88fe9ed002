"""Key all-gather — the one exchange step of the data-parallel RMCL path.

Replaces ``_concat_all_gather`` (vilt/modules/objectives.py:226-235) and
``concat_all_gather`` (MoCo/MoCo_RMCL.py:268-279): instead of ``world`` ``ones_like`` buffers,
the list form of ``all_gather`` and a ``torch.cat``, one ``all_gather_into_tensor`` (a single
ncclAllGather over NVLink on the GPU box, gloo in the CPU tests) writes straight into a contiguous
[B_global, C] buffer in rank order.  Every rank then enqueues the identical gathered batch into
its replica of the queue, so queue and pointer stay bit-identical across ranks without a
broadcast (the reference relied on DDP's per-forward ``broadcast_buffers``).
"""
import torch
import torch.distributed as dist


@torch.no_grad()
def concat_all_gather(tensor, group=None):
    """Gathers ``tensor`` [B_local, ...] from every rank along dim 0, rank order. No gradient."""
    if not (dist.is_available() and dist.is_initialized()):
        return tensor
    world = dist.get_world_size(group)
    if world == 1:
        return tensor
    tensor = tensor.contiguous()
    out = torch.empty((world * tensor.shape[0],) + tuple(tensor.shape[1:]), dtype=tensor.dtype, device=tensor.device)
    dist.all_gather_into_tensor(out, tensor, group=group)
    return out


def gathered_batch_matches(per_step_bs, gathered_rows):
    """The reference silently skips the enqueue when the gathered batch differs from
    ``per_step_bs`` (objectives.py:242-243), e.g. on the last, short batch of an epoch."""
    return per_step_bs is None or int(per_step_bs) == int(gathered_rows)


class Gather:
    """Callable all-gather with the caller's rank attached: what ``ops.barlow_twins_loss`` needs to replace the
    reference's all-reduce of the D x D cross-correlation (objectives.py:482) by an all-gather of the [B, D]
    projections.  ``Gather()`` is a no-op (rank 0) outside an initialised process group."""

    def __init__(self, group=None):
        self.group = group
        on = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank(group) if on else 0
        self.world = dist.get_world_size(group) if on else 1

    def __call__(self, tensor):
        return concat_all_gather(tensor, self.group)
