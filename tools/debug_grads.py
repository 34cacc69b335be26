#!/usr/bin/env python
"""Per-tensor gradient errors of one golden case on the GPU (debug aid; prints instead of asserting).
usage: python tools/debug_grads.py small_reg_mix [tf32=1]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bnn_b200  # noqa: E402
from tests import parity_cases as PC  # noqa: E402
from tests.golden_util import Case  # noqa: E402

name = sys.argv[1]
tf32 = len(sys.argv) < 3 or sys.argv[2] == '1'
c = Case(name)
net = PC.build_net(c, 'cuda', fused=True, tf32=tf32).train()
x, y = c.x.cuda(), c.y.cuda()
with bnn_b200.eps_mode('reference'):
    torch.manual_seed(c.meta['seeds'][2])
    info = (net.sample_elbo_lr if c.lr else net.sample_elbo)(x, y, c.beta, c.S, sigma=c.sigma)
info[0].backward()
print(name, 'dims', c.dims, 'B', c.B, 'S', c.S, 'loss', float(info[0].detach()), 'want', float(c.z['loss'][0]))
grads = PC.net_grads(net)
for li, pi, key in c.grad_keys():
    g = np.asarray(grads[li][pi], dtype=np.float64)
    gmax = float(c.z[f'gmax.{key}'][0])
    if c.meta['full']:
        want = c.z[f'grad.{key}'].astype(np.float64)
        d = np.abs(g - want)
    else:
        st = c.meta['slice']
        want = c.z[f'gslice.{key}'].astype(np.float64)
        d = np.abs(g.reshape(-1)[::st] - want)
    print(f'{key:16s} shape {str(g.shape):14s} err {d.max() / gmax:.3e}  nan {int(np.isnan(g).sum())}  argmax {np.unravel_index(d.argmax(), d.shape)}')
    if c.meta['full'] and d.max() / gmax > 1e-2 and g.ndim == 2:
        bad = (d / gmax > 1e-2)
        print('   bad rows', np.where(bad.any(1))[0][:20], 'bad cols', np.where(bad.any(0))[0][:20])
if c.meta['full']:
    np.set_printoptions(precision=4, linewidth=200, suppress=False)
    for key_li in ((2, 'l3.weight_mu'), (1, 'l2.weight_mu')):
        li, key = key_li
        g = np.asarray(grads[li][0], dtype=np.float64)
        want = c.z[f'grad.{key}'].astype(np.float64)
        print(key, 'got[:3,:12]\n', g[:3, :12], '\nwant\n', want[:3, :12])
    print('l2.bias_mu got', np.asarray(grads[1][2]), '\nwant', c.z['grad.l2.bias_mu'])
