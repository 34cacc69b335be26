"""Load a tests/golden/*.npz fixture and replay the reference's RNG draws for it."""
import json
import os

import numpy as np
import torch

from oracle import bbb_oracle as O
from oracle.make_golden import make_data, MU_INIT, RHO_INIT

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
SMALL = ['small_reg_mix', 'small_cls_mix', 'small_cls_gauss', 'small_bandit_bcast']
SMALL_LR = ['small_lr_cls', 'small_lr_reg']
BIG = ['cfg1_reg_mix', 'cfg4_bandit', 'cfg2_mnist_mix']
BIG_LR = ['cfg3_mnist_lr']
DEEP_SMALL = ['deep5_small_mix']          # five reference BayesianLinear layers composed by hand, fully stored
DEEP_BIG = ['deep5_cfg5_scaled']          # scaled BASELINE.json config 5: runs the batch-resident kernels
PNAMES = ('weight_mu', 'weight_rho', 'bias_mu', 'bias_rho')


class Case:
    """One fixture with its inputs regenerated from the recorded seeds."""

    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, name + '.npz'))
        self.z = {k: z[k] for k in z.files}
        self.meta = json.loads(str(self.z['meta']))
        m = self.meta
        self.name, self.kind, self.mode = name, m['kind'], m['mode']
        self.dims, self.B, self.S = m['dims'], m['B'], m['S']
        self.beta, self.sigma = m['beta'], m['sigma']
        self.lr = self.kind == 'lr'
        self.prior = O.make_prior(m['prior_init'], m['mixture'])
        self.prior_init, self.mixture = m['prior_init'], m['mixture']
        with torch.random.fork_rng():
            torch.manual_seed(m['seeds'][0])
            self.layers = O.init_layers(self.dims, MU_INIT, RHO_INIT, local_reparam=self.lr)
            self.x, self.y = make_data(m)
            torch.manual_seed(m['seeds'][2])
            self.eps = O.draw_eps(self.dims, self.S, batch=self.B, local_reparam=self.lr)
            if m['full']:
                torch.manual_seed(m['seeds'][2] + 1)
                self.eps_eval = O.draw_eps(self.dims, 1, batch=self.B, local_reparam=self.lr)[0]
        # replay check: parameter checksums recorded from the reference constructor
        for li, layer in enumerate(self.layers):
            for pn, p in zip(PNAMES, layer):
                want = float(self.z[f'psum.l{li + 1}.{pn}'][0])
                got = float(p.double().sum())
                assert abs(want - got) <= 1e-9 * max(1.0, abs(want)), f'{name}: RNG replay mismatch on l{li+1}.{pn}'

    def model_params(self):
        d = self.dims
        hidden = d[1] if len(d) == 4 else list(d[1:-1])
        return dict(input_shape=d[0], classes=d[-1], batch_size=self.B, hidden_units=hidden, mode=self.mode,
                    mu_init=MU_INIT, rho_init=RHO_INIT, prior_init=self.prior_init,
                    mixture_prior=self.mixture, local_reparam=self.lr)

    def grad_keys(self):
        return [(li, pi, f'l{li + 1}.{pn}') for li in range(len(self.layers)) for pi, pn in enumerate(PNAMES)]

    def cancel_floor(self, li, pi):
        """Round-off the REFERENCE's own fp32 autograd leaves in a mu/rho gradient:
        it adds +(beta/S) eps/sigma (d log q / d mu, direct) and -(beta/S) eps/sigma (through w) to the
        other terms, so each element carries up to ~1 ulp of (beta/S)|eps|/sigma per sample.  The analytic
        CUDA/closed-form path cancels those terms exactly.  Weight sampling only; 0 for LR."""
        if self.lr:
            return 0.0
        layer = self.layers[li]
        rho = layer[1] if pi < 2 else layer[3]
        sig = torch.log1p(torch.exp(rho.double()))
        acc = torch.zeros_like(sig)
        for s in range(self.S):
            acc += (self.eps[s][li][0 if pi < 2 else 1].double().abs() / sig)
        return (2.0 ** -22) * (self.beta / self.S) * acc.numpy()

    def check_grads(self, grads, rtol, allow_cancel_floor=False):
        """grads[li][pi] array-likes; max-norm-relative comparison per tensor:
        |g - g_ref| <= rtol * max|g_ref| (+ the reference's own cancellation round-off if allowed)."""
        worst = 0.0
        for li, pi, key in self.grad_keys():
            g = np.asarray(grads[li][pi], dtype=np.float64)
            gmax = float(self.z[f'gmax.{key}'][0])
            floor = self.cancel_floor(li, pi) if allow_cancel_floor else 0.0
            if self.meta['full']:
                want = self.z[f'grad.{key}'].astype(np.float64)
                err = (np.maximum(np.abs(g - want) - floor, 0.0)).max() / gmax
            else:
                st = self.meta['slice']
                want = self.z[f'gslice.{key}'].astype(np.float64)
                fl = floor.reshape(-1)[::st] if allow_cancel_floor and not self.lr else 0.0
                err = np.maximum(np.abs(g.reshape(-1)[::st] - want) - fl, 0.0).max() / gmax
                nrm = float(self.z[f'gnorm.{key}'][0])
                err = max(err, abs(np.sqrt((g ** 2).sum()) - nrm) / nrm)
            worst = max(worst, err)
            assert err <= rtol, f'{self.name} {key}: max-norm relative error {err:.3e} > {rtol:.1e}'
        return worst

    def check_outputs(self, out, rtol):
        out = np.asarray(out, dtype=np.float64)
        if self.meta['full']:
            want = self.z['outputs'].astype(np.float64)
            got = out.reshape(want.shape)
        else:
            want = self.z['outputs_slice'].astype(np.float64)
            got = out.reshape(-1)[::7]
        err = np.abs(got - want).max() / np.abs(want).max()
        assert err <= rtol, f'{self.name} outputs: {err:.3e} > {rtol:.1e}'
        return err

    def check_scalar(self, key, val, rtol):
        want = float(self.z[key][0])
        got = float(np.asarray(val).reshape(-1)[0])
        assert abs(got - want) <= rtol * abs(want), f'{self.name} {key}: got {got!r} want {want!r}'
