#!/usr/bin/env python
"""Debug aid (GPU): the network-level path (bbb_mlp_fwd / bbb_mlp_bwd) against the per-layer exact-fp32 path on the same
Philox coordinates, shape by shape, printing the worst relative error of every scalar and gradient tensor.
usage: python tools/check_mlp.py [--quick]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bnn_b200
from tests.golden_util import PNAMES

DEV = 'cuda'
CASES = [  # dims, B, S, mode
    ([784, 1200, 1200, 10], 128, 2, 'classification'),
    ([64, 64, 10], 128, 1, 'classification'),
    ([36, 100, 72, 5], 64, 3, 'classification'),
    ([400, 400, 400, 1], 128, 5, 'regression'),
    ([128, 256, 136, 12, 4], 7, 2, 'classification'),
]


def run(dims, B, S, mode, tf32, seed=3):
    torch.manual_seed(0)
    mp = dict(input_shape=dims[0], classes=dims[-1], batch_size=B, hidden_units=list(dims[1:-1]), mode=mode,
              mu_init=[-0.2, 0.2], rho_init=[-5, -4], prior_init=[0.5, 0, -6], mixture_prior=True, tf32=tf32)
    net = bnn_b200.BayesianNetwork(mp).to(DEV).train()
    torch.manual_seed(1)
    x = torch.randn(B, dims[0], device=DEV)
    y = torch.randint(0, dims[-1], (B,), device=DEV) if mode == 'classification' else torch.randn(B, dims[-1], device=DEV)
    bnn_b200.manual_seed(seed, 1)
    info = net.sample_elbo(x, y, 0.3, S, sigma=0.5)
    info[0].backward()
    torch.cuda.synchronize()
    grads = [[getattr(l, pn).grad.detach().cpu().numpy() for pn in PNAMES] for l in net.layers()]
    return [float(v.detach()) for v in info], grads


def main():
    cases = CASES[:2] if '--quick' in sys.argv else CASES
    if '--layerwise' in sys.argv:        # the per-layer TF32 kernels against the exact path: the baseline of this check
        from bnn_b200 import functional as F
        F.use_network_level_call = False
    ok = True
    for dims, B, S, mode in cases:
        t0 = time.time()
        i0, g0 = run(dims, B, S, mode, False)
        i1, g1 = run(dims, B, S, mode, True)
        rel = [abs(a - b) / max(1e-30, abs(a)) for a, b in zip(i0, i1)]
        worst = 0.0
        for li, (a, b) in enumerate(zip(g0, g1)):
            for pi, (ga, gb) in enumerate(zip(a, b)):
                e = float(np.abs(ga - gb).max() / max(1e-30, np.abs(ga).max()))
                worst = max(worst, e)
                if e > 5e-3 or not np.isfinite(e):
                    print(f'   l{li + 1}.{PNAMES[pi]}: rel err {e:.3e}')
        good = rel[1] < 1e-5 and rel[2] < 1e-5 and rel[0] < 5e-3 and rel[3] < 5e-3 and worst < 5e-3
        ok &= good
        print(f'{dims} B={B} S={S} {mode}: loss/logp/logq/nll rel err {rel[0]:.2e} {rel[1]:.2e} {rel[2]:.2e} {rel[3]:.2e}; '
              f'worst grad {worst:.2e}  [{"ok" if good else "FAIL"}]  ({time.time() - t0:.1f}s)', flush=True)
    print('ALL OK' if ok else 'FAILED')
    return 0 if ok else 1


if __name__ == '__main__':
    sys.exit(main())
