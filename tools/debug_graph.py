import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bnn_b200
from tests import parity_cases as PC
from tests.golden_util import Case
c = Case('cfg2_mnist_mix'); DEV = 'cuda'
x = c.x.to(DEV); y = c.y.to(DEV)
def eager(n, seed=321, step=102):
    net = PC.build_net(c, DEV, tf32=True).train()
    opt = bnn_b200.FusedAdam(net.parameters(), lr=1e-3)
    bnn_b200.manual_seed(seed, step)
    for k in range(n):
        net.zero_grad(); info = net.sample_elbo(x, y, c.beta, c.S, sigma=c.sigma); info[0].backward(); opt.step()
    return [p.detach().clone() for p in net.parameters()], [p.grad.detach().clone() for p in net.parameters()]
def stats(a, b, tag):
    for i, (p, q) in enumerate(zip(a, b)):
        d = (p - q).abs()
        bad = ~torch.isclose(p, q, rtol=1e-4, atol=2e-6)
        print(tag, i, tuple(p.shape), 'max', float(d.max()), 'frac_bad', float(bad.float().mean()), 'frac>1e-5', float((d > 1e-5).float().mean()))
for n in (1, 2, 3):
    p1, g1 = eager(n); p2, g2 = eager(n)
    stats(p1[:2], p2[:2], f'eager-vs-eager params n={n}')
    stats(g1[:2], g2[:2], f'eager-vs-eager grads  n={n}')
