"""All-reduce latency of the gradient bucket (19.2 MB fp32 at the MNIST-shape config): NCCL vs symmetric-memory ops.
torchrun --nproc-per-node N tools/time_allreduce.py"""
import os, sys, torch, torch.distributed as dist
rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local); dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
n = 2 * 2395210
def timeit(fn, it=30):
    for _ in range(5): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / it
t = torch.randn(n, device=dev)
us = timeit(lambda: dist.all_reduce(t, op=dist.ReduceOp.AVG))
if rank == 0: print(f'nccl all_reduce AVG {n*4/1e6:.1f} MB: {us:.1f} us', flush=True)
try:
    import torch.distributed._symmetric_memory as symm_mem
    g = dist.group.WORLD
    st = symm_mem.empty(n, dtype=torch.float32, device=dev)
    hdl = symm_mem.rendezvous(st, g.group_name)
    st.normal_()
    for name in ('multimem_all_reduce_', 'two_shot_all_reduce_', 'one_shot_all_reduce'):
        try:
            op = getattr(torch.ops.symm_mem, name)
            us = timeit(lambda: op(st, 'sum', g.group_name))
            if rank == 0: print(f'symm_mem.{name}: {us:.1f} us', flush=True)
        except Exception as e:
            if rank == 0: print(f'symm_mem.{name} failed: {type(e).__name__}: {str(e)[:200]}', flush=True)
except Exception as e:
    if rank == 0: print('symmetric memory unavailable:', type(e).__name__, str(e)[:300], flush=True)
dist.barrier(); torch.cuda.synchronize(); os._exit(0)
