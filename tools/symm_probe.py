"""2+ GPU probe (torchrun): does torch's symmetric memory come up on this node, and does it hand out a multicast
(NVLS) address?  Then one multimem.ld_reduce / multimem.st round trip through libbbb (if built with the probe entry)."""
import os
import sys

import torch
import torch.distributed as dist

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
import torch.distributed._symmetric_memory as sm  # noqa: E402
try:
    print(rank, 'has_multicast_support', sm._SymmetricMemory.has_multicast_support(torch._C._autograd.DeviceType.CUDA, local), flush=True)
except Exception as e:
    print(rank, 'has_multicast_support query failed:', type(e).__name__, e, flush=True)
try:
    t = sm.empty(1 << 20, dtype=torch.float32, device=dev)
    t.fill_(rank + 1)
    h = sm.rendezvous(t, dist.group.WORLD)
    print(rank, 'backend', sm._SymmetricMemory.get_backend(dev) if hasattr(sm._SymmetricMemory, 'get_backend') else '?',
          'multicast_ptr', hex(h.multicast_ptr), 'buffer_ptrs', [hex(p) for p in h.buffer_ptrs], 'signal pads',
          len(h.signal_pad_ptrs), flush=True)
    h.barrier()
    peer = h.get_buffer((rank + 1) % world, (4,), torch.float32)
    print(rank, 'peer value', float(peer[0]), flush=True)
except Exception as e:
    print(rank, 'symmetric memory failed:', type(e).__name__, str(e)[:300], flush=True)
dist.barrier()
dist.destroy_process_group()
