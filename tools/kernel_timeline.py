#!/usr/bin/env python
"""Debug aid (GPU): phase timeline of the network-level kernels (bbb_debug_set_timeline).  For one layer's forward and
backward kernel of the MNIST-shape step, prints per phase the mean / max over CTAs of the time between consecutive
%globaltimer stamps, and the span from the first CTA's entry to the last CTA's exit.
usage: python tools/kernel_timeline.py"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bnn_b200
from bnn_b200 import _lib as L
from bnn_b200 import functional as F

FWD = ['entry', 'setup done', 'pdl_wait done', 'stage0 data', 'stage0 sampled', 'stage1 data', 'stage1 sampled',
       'stage2 data', 'stage2 sampled', 'stage3 data', 'stage3 sampled', 'accumulator ready', 'staged', 'reduce done + fence',
       'pair complete', 'finalised']
BWD = ['entry', 'setup done', 'pdl_wait done', 'mu/rho landed', 'G ready', 'region0 done', 'region1 done', 'region2 done',
       'W complete', 'dX0 ready', 'dX0 drained', 'dX1 ready', 'dX1 drained', '-', '-', 'exit']


def report(name, t, labels):
    t = t.astype(np.int64)
    used = t[:, 0] > 0
    t = t[used]
    print(f'{name}: {len(t)} CTAs, first entry -> last exit {1e-3 * (t[:, 15].max() - t[:, 0].min()):.2f} us; '
          f'mean CTA lifetime {1e-3 * (t[:, 15] - t[:, 0]).mean():.2f} us')
    prev = 0
    for k in range(1, 16):
        if labels[k] == '-' or not (t[:, k] > 0).all():
            continue
        d = 1e-3 * (t[:, k] - t[:, prev])
        print(f'   {labels[prev]:>20s} -> {labels[k]:<22s} mean {d.mean():6.2f}  max {d.max():6.2f}  min {d.min():6.2f} us'
              f'   (at {1e-3 * (t[:, k] - t[:, 0].min()).mean():6.2f} us)')
        prev = k


def main():
    dev = 'cuda'
    torch.manual_seed(0)
    mp = dict(input_shape=784, classes=10, batch_size=128, hidden_units=[1200, 1200], mode='classification',
              mu_init=[-0.2, 0.2], rho_init=[-5, -4], prior_init=[0.5, 0, -8], mixture_prior=True, tf32=True)
    net = bnn_b200.BayesianNetwork(mp).to(dev).train()
    x = torch.rand(128, 784, device=dev)
    y = torch.randint(0, 10, (128,), device=dev)
    bnn_b200.manual_seed(1)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(3):
        net.zero_grad()
        net.sample_elbo(x, y, 0.5, 2)[0].backward()
    torch.cuda.synchronize()
    buf = torch.zeros(8 * 2560, dtype=torch.int64, device=dev)     # one 160 x 16 slice per kernel launch
    lib = L.lib()
    flush.zero_()
    torch.cuda.synchronize()
    net.zero_grad()
    lib.bbb_debug_set_timeline(buf.data_ptr())
    net.sample_elbo(x, y, 0.5, 2)[0].backward()
    torch.cuda.synchronize()
    lib.bbb_debug_set_timeline(None)
    t = buf.cpu().numpy().reshape(8, 160, 16)
    report('forward 784x1200', t[0], FWD)
    report('forward 1200x1200', t[1], FWD)
    report('backward 1200x1200', t[2], BWD)
    report('backward 784x1200 (no dgrad)', t[3], BWD)
    return 0


if __name__ == '__main__':
    sys.exit(main())
