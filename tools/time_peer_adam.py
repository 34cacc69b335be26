"""2+ GPU timing aid (torchrun): phase breakdown of the fused peer-memory exchange (bbb_adam_step_peer) inside the
captured MNIST-shape train step.  Six %globaltimer stamps per call (bbb_debug_set_timeline, slot 8): kernel start ->
every rank's gradients complete (entry barrier) -> block 0 has issued its loads / update / stores -> its peer stores are
performed -> this rank's last block has finished -> every rank has finished (exit barrier).
usage: torchrun --nproc-per-node N tools/time_peer_adam.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bnn_b200  # noqa: E402
from bnn_b200 import _lib as L  # noqa: E402


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    torch.manual_seed(0)
    mp = dict(input_shape=784, classes=10, batch_size=128, hidden_units=1200, mode='classification',
              mu_init=[-0.2, 0.2], rho_init=[-5, -4], prior_init=[0.5, 0, -8], mixture_prior=True, tf32=True)
    net = bnn_b200.BayesianNetwork(mp).to(dev).train()
    opt = bnn_b200.PeerShardedAdam(net.parameters(), lr=1e-3)
    g = torch.Generator().manual_seed(1)
    x = torch.rand(128, 784, generator=g).to(dev)
    y = torch.randint(0, 10, (128,), generator=g).to(dev)
    bnn_b200.manual_seed(7, 0)
    bnn_b200.set_sample_base(rank * 2)
    gs = bnn_b200.GraphedTrainStep(net, opt, x, y, 2, beta=0.5, world_size=world)
    tl = torch.zeros(9 * 2560, dtype=torch.int64, device=dev)
    sync = torch.zeros(1, device=dev)
    for _ in range(5):
        gs(x, y)
    torch.cuda.synchronize()
    L.check(L.lib().bbb_debug_set_timeline(tl.data_ptr()), 'timeline')
    # the stamps are written by the captured launch only if the pointer was set at capture time: capture again
    gs = bnn_b200.GraphedTrainStep(net, opt, x, y, 2, beta=0.5, world_size=world)
    rows = []
    for _ in range(40):
        dist.all_reduce(sync)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gs(x, y); e1.record()
        torch.cuda.synchronize()
        t = tl[8 * 2560:8 * 2560 + 6].cpu().numpy().astype(np.int64)
        rows.append(list(np.diff(t) * 1e-3) + [e0.elapsed_time(e1) * 1e3])
    L.check(L.lib().bbb_debug_set_timeline(None), 'timeline')
    r = np.median(np.array(rows[5:]), axis=0)
    names = ['entry barrier', 'block 0 body', 'block 0 store fence', 'until last block', 'exit barrier', 'whole step']
    print(f'rank {rank}/{world}: ' + '  '.join(f'{n} {v:6.1f} us' for n, v in zip(names, r)), flush=True)
    opt.check_health()
    if hasattr(opt, 'release'):
        opt.release()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
