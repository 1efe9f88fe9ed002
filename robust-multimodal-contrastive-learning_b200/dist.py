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


class P2PKeyExchange:
    """Key all-gather + enqueue as ONE kernel per rank over NVLink peer memory (include/rmcl_b200.h,
    rmcl_gather_enqueue_p2p) instead of ncclAllGather followed by the enqueue kernel.

    The staging buffers and flag words live in torch's symmetric memory (plumbing only: allocation, the IPC handle
    exchange and the one-time barrier); the data path — pushes into the peers' HBM, the system-scope signal/wait,
    the transposing scatter into the local queue replica — is the CUDA kernel.  Single node only."""

    def __init__(self, B_local, C, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        group = dist.group.WORLD if group is None else group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.B, self.C = int(B_local), int(C)
        self.stage = symm_mem.empty((2, self.world * self.B, self.C), dtype=torch.float32, device=device)
        self.flags = symm_mem.empty((64,), dtype=torch.int32, device=device)
        self.flags.zero_()
        self.stage_hdl = symm_mem.rendezvous(self.stage, group)
        self.flags_hdl = symm_mem.rendezvous(self.flags, group)
        torch.cuda.synchronize(device)
        self.flags_hdl.barrier()              # every rank's flags are zero before anyone signals

    def enqueue_(self, queue, keys_local, ptr, shadow=None):
        """In place: gathers every rank's ``keys_local`` [B_local, C] (fp32) and enqueues the world*B_local keys into
        this rank's ``queue`` replica, advancing ``ptr`` — the reference's _dequeue_and_enqueue(keys) under DDP."""
        from . import _lib, ops
        if keys_local.dtype != torch.float32 or not keys_local.is_contiguous() or tuple(keys_local.shape) != (self.B, self.C):
            raise ValueError("keys_local must be a contiguous fp32 [B_local, C] tensor")
        if shadow is None:
            shadow = ops._find_shadow(queue)      # a registered shadow is always kept current
        sh = None if shadow is None else (shadow.full(queue) if isinstance(shadow, ops.QueueShadow) else shadow)
        rc = _lib.lib().rmcl_gather_enqueue_p2p(
            self.stage_hdl.buffer_ptrs_dev, self.flags_hdl.buffer_ptrs_dev, keys_local.data_ptr(), queue.data_ptr(), ops._dt(queue),
            None if sh is None else sh.data_ptr(), 0 if sh is None else sh.stride(0),
            1 if sh is None else sh.shape[0] // queue.shape[0], ptr.data_ptr(), self.rank, self.world,
            self.B, self.C, queue.shape[1], queue.stride(0), ops._stream())
        _lib.check(rc, "rmcl_gather_enqueue_p2p")
