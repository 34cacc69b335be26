"""Multi-GPU partitioning of the BBB step: one process per GPU, parameters replicated,
Monte-Carlo samples sharded -- rank r runs global sample indices [r*S, (r+1)*S) (rng.set_sample_base), so
the Philox streams are disjoint by construction and the result is invariant to the number of ranks.
The forward and backward need no communication; the only exchange is one all-reduce of the mu/rho
gradients (NCCL over NVLink on GPUs, gloo in the CPU tests), after which every rank applies the same
optimiser update.
"""
import torch
import torch.distributed as dist


def _flat_view_of(grads):
    """If the gradients are consecutive slices of one storage (the network-level backward allocates them
    that way and autograd installs them without copying), return that storage as one flat tensor --
    no copy; else None."""
    if not grads:
        return None
    st = grads[0].untyped_storage()
    off = grads[0].storage_offset()
    first = off
    for g in grads:
        if (g.dtype != torch.float32 or not g.is_contiguous() or g.storage_offset() != off
                or g.untyped_storage().data_ptr() != st.data_ptr()):
            return None
        off += g.numel()
    flat = torch.empty(0, dtype=torch.float32, device=grads[0].device)
    flat.set_(st, first, (off - first,))
    return flat


def allreduce_gradients(module, world_size=None, group=None):
    """Average .grad over the ranks with ONE collective over the flat [grad_mu | grad_rho ...] bucket."""
    world_size = world_size or dist.get_world_size(group)
    if world_size == 1:
        return
    grads = [p.grad for p in module.parameters() if p.grad is not None]
    flat = _flat_view_of(grads)
    copied = flat is None
    if copied:
        flat = torch.cat([g.reshape(-1) for g in grads])
    if dist.get_backend(group) == 'nccl':
        dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=group)
    else:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.div_(world_size)
    if copied:
        off = 0
        for g in grads:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()


def allreduce_scalars(values, world_size=None, group=None):
    """Average the ELBO scalars (loss, log prior, log posterior, NLL) over the ranks."""
    world_size = world_size or dist.get_world_size(group)
    t = torch.stack([v.detach().reshape(()) for v in values])
    if world_size > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        t = t / world_size
    return t


def shard_samples(total_samples, rank, world_size):
    """Contiguous block of global MC-sample indices for `rank`: (first index, count)."""
    per, rem = divmod(total_samples, world_size)
    count = per + (1 if rank < rem else 0)
    first = rank * per + min(rank, rem)
    return first, count
