"""Probe: cross-process peer memory on one node through torch's CUDA IPC handles (torchrun, 2+ ranks)."""
import os
import time

import torch
import torch.distributed as dist


def open_peers(t, group=None):
    """All-gather the CUDA IPC handle of t's storage; return [tensor view of rank k's buffer for k in ranks]."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    h = t.untyped_storage()._share_cuda_()
    hs = [None] * world
    dist.all_gather_object(hs, (h, t.storage_offset(), t.numel()), group=group)
    out = []
    for k, (hk, off, n) in enumerate(hs):
        if k == rank:
            out.append(t)
            continue
        st = torch.UntypedStorage._new_shared_cuda(*hk)
        v = torch.empty(0, dtype=t.dtype, device=st.device)   # the storage lives on the owner's device
        v.set_(st, off, (n,))
        out.append(v)
    return out


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    n = 8 << 20
    t = torch.full((n,), float(rank), device=dev)
    peers = open_peers(t)
    dist.barrier()
    torch.cuda.synchronize()
    other = peers[(rank + 1) % world]
    print(rank, 'peer device', other.device, 'first', float(other[0]), flush=True)
    # peer read bandwidth
    dst = torch.empty_like(t)
    for _ in range(3):
        dst.copy_(other)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        dst.copy_(other)
    e1.record()
    torch.cuda.synchronize()
    print(rank, 'peer read GB/s', 10 * n * 4 / (e0.elapsed_time(e1) * 1e-3) / 1e9, flush=True)
    # peer write
    other[:4] = 100.0 + rank
    torch.cuda.synchronize()
    dist.barrier()
    print(rank, 'after peer write', t[:4].tolist(), flush=True)
    dist.barrier()
    os._exit(0)


if __name__ == '__main__':
    main()
