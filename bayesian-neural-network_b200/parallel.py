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


class OverlappedAllReduce:
    """Per-layer gradient all-reduce on a side stream, overlapped with the rest of the backward.

        with parallel.OverlappedAllReduce(world_size) as ar:
            loss = net.sample_elbo(...)[0]
            loss.backward()             # layer l's bucket slice is all-reduced while layers < l are differentiated
        ar.join()                       # the current stream waits for the collectives
        optimizer.step()

    Works under CUDA-graph capture (the fork/join become graph edges).  Only the network-level weight-sampling
    backward reports its layers (functional.grad_ready_hook); anything it does not cover is reduced by join()."""

    def __init__(self, world_size=None, group=None):
        self.world = world_size or dist.get_world_size(group)
        self.group = group
        self.stream = None
        self.reduced = 0

    def __enter__(self):
        from . import functional as F
        self.reduced = 0
        if self.world > 1:
            F.grad_ready_hook = self._on_layer
        return self

    def __exit__(self, *exc):
        from . import functional as F
        F.grad_ready_hook = None

    def _reduce(self, flat):
        if dist.get_backend(self.group) == 'nccl':
            dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=self.group)
        else:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
            flat.div_(self.world)

    def _on_layer(self, layer, flat):
        self.reduced += 1
        if not flat.is_cuda:                            # CPU tests (gloo): nothing to overlap with
            self._reduce(flat)
            return
        if self.stream is None:
            self.stream = torch.cuda.Stream()
        cur = torch.cuda.current_stream()
        self.stream.wait_stream(cur)                    # the layer's backward kernel has been launched on `cur`
        with torch.cuda.stream(self.stream):
            self._reduce(flat)
        flat.record_stream(self.stream)

    def join(self, module=None):
        """Make the current stream wait for the collectives; if no layer reported (other estimators / layer-level
        API), fall back to one all-reduce over the whole bucket."""
        if self.world <= 1:
            return
        if self.reduced == 0:
            if module is not None:
                allreduce_gradients(module, self.world, self.group)
            return
        if self.stream is not None:
            torch.cuda.current_stream().wait_stream(self.stream)


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


# ---- the minibatch-row axis ---------------------------------------------------------------------------------------------
# north_star: "Monte-Carlo ELBO samples AND minibatch shards are partitioned across the GPUs".  Ranks that share a sample
# block (same Philox sample indices -> the same sampled weights) split the ROWS of the minibatch.  With
#     L = 1/S sum_s [ beta (log q_s - log p_s) + NLL_s(all rows) ]                                   (networks.py:199-209)
# a rank holding the rows of shard j (of R) computes  l = sample_elbo(x_j, y_j, beta / R, S_local).loss : the rank-SUM of
# l over the R row shards of a sample block is that block's loss, so the MEAN over all ranks of  R * l  is L -- and the mean
# gradient every exchange here computes (all-reduce / R, PeerShardedAdam) is exactly dL.  No kernel knows about it.
class ShardPlan:
    """Where `rank` sits in the (sample blocks) x (row shards) grid of `world` ranks.
    sample_first, sample_count: its block of global MC-sample indices (rng.set_sample_base(sample_first));
    row_lo, row_hi: its rows of the minibatch;  row_shards: R;  beta_scale = 1 / R;  loss_scale = R."""

    def __init__(self, rank, world, samples_total, batch, row_shards=None):
        if row_shards is None:         # rows are only split when there are more ranks than samples
            row_shards = 1
            while world // row_shards > samples_total and world % (row_shards * 2) == 0:
                row_shards *= 2
        if world % row_shards != 0:
            raise ValueError('row_shards must divide the world size')
        self.rank, self.world, self.row_shards = rank, world, row_shards
        self.sample_blocks = world // row_shards
        self.block, self.row_index = rank // row_shards, rank % row_shards
        self.sample_first, self.sample_count = shard_samples(samples_total, self.block, self.sample_blocks)
        self.row_lo = self.row_index * batch // row_shards
        self.row_hi = (self.row_index + 1) * batch // row_shards
        self.beta_scale, self.loss_scale = 1.0 / row_shards, float(row_shards)

    def rows(self, *tensors):
        out = tuple(t[self.row_lo:self.row_hi] for t in tensors)
        return out if len(out) > 1 else out[0]


def sharded_elbo(net, x_rows, y_rows, beta, plan, sigma=1.0):
    """sample_elbo of this rank's shard (x_rows, y_rows = plan.rows(x, y)).  Returns the 4-tuple of sample_elbo with the
    loss already multiplied by plan.loss_scale: call .backward() on it and exchange gradients as usual (rank mean).  The
    caller has set rng.set_sample_base(plan.sample_first).  Reporting: the rank mean of the returned loss is the full
    loss; log prior / posterior are per sample block, the NLL is this shard's."""
    fn = net.sample_elbo_lr if getattr(net, 'local_reparam', False) else net.sample_elbo
    info = fn(x_rows, y_rows, beta * plan.beta_scale, plan.sample_count, sigma)
    return (info[0] * plan.loss_scale,) + tuple(info[1:])
