#!/usr/bin/env python
"""Generate tests/golden/*.npz by RUNNING THE UNMODIFIED REFERENCE.

    python oracle/make_golden.py [--reference /root/reference] [--out tests/golden]

TEST INFRASTRUCTURE ONLY.  Runs in the build container (the reference tree is
not present on the GPU box); the fixtures it writes are committed and are what
pins both oracle/bbb_oracle.py and the CUDA path.

Every case seeds the global torch CPU generator three times -- parameters
(uniform_ draws inside the reference constructors, networks.py:53-58), data,
eps (Normal(0,1).sample == randn on the global generator, networks.py:35,42) --
so a consumer can replay the reference's draws with oracle.bbb_oracle.init_layers
/ draw_eps.  Small cases also store every tensor outright; the config-sized cases
store seeds, scalars, checksums and strided gradient slices only.
"""
import argparse
import importlib.util
import json
import os
import sys
import tempfile

import numpy as np
import torch

SEED_PARAMS, SEED_DATA, SEED_EPS = 0, 1, 2
SLICE = 257          # stride used for the gradient / output slices of the big cases

# name, kind, mode, dims, B, S, prior_init, mixture, sigma, M(num_batches), idx, store_full, extra
CASES = [
    dict(name='small_reg_mix', kind='ws', mode='regression', dims=[5, 12, 12, 3], B=7, S=3,
         prior_init=[0.5, -0, -6], mixture=True, sigma=0.1, M=8, idx=2, full=True),
    dict(name='small_cls_mix', kind='ws', mode='classification', dims=[16, 12, 12, 5], B=6, S=2,
         prior_init=[0.5, -0, -8], mixture=True, sigma=1.0, M=4, idx=0, full=True, img=(1, 4, 4)),
    dict(name='small_cls_gauss', kind='ws', mode='classification', dims=[16, 12, 12, 5], B=6, S=2,
         prior_init=[1.], mixture=False, sigma=1.0, M=4, idx=1, full=True, img=(1, 4, 4)),
    dict(name='small_bandit_bcast', kind='ws', mode='regression', dims=[9, 10, 10, 1], B=8, S=2,
         prior_init=[0.5, -0, -6], mixture=True, sigma=1.0, M=64, idx=3, full=True, flat_target=True),
    dict(name='small_lr_cls', kind='lr', mode='classification', dims=[16, 12, 12, 5], B=6, S=2,
         prior_init=[1.], mixture=False, sigma=1.0, M=4, idx=0, full=True, img=(1, 4, 4)),
    dict(name='small_lr_reg', kind='lr', mode='regression', dims=[5, 12, 12, 3], B=7, S=3,
         prior_init=[1], mixture=False, sigma=0.1, M=8, idx=1, full=True),
    # BASELINE.json configs 1-4 at full size (seeds + checksums + slices)
    dict(name='cfg1_reg_mix', kind='ws', mode='regression', dims=[1, 400, 400, 1], B=128, S=5,
         prior_init=[0.5, -0, -6], mixture=True, sigma=0.1, M=8, idx=0, full=False, reg_data=True),
    dict(name='cfg2_mnist_mix', kind='ws', mode='classification', dims=[784, 1200, 1200, 10], B=128, S=2,
         prior_init=[0.5, -0, -8], mixture=True, sigma=1.0, M=468, idx=0, full=False, img=(1, 28, 28)),
    dict(name='cfg3_mnist_lr', kind='lr', mode='classification', dims=[784, 1200, 1200, 10], B=128, S=2,
         prior_init=[1.], mixture=False, sigma=1.0, M=468, idx=0, full=False, img=(1, 28, 28)),
    dict(name='cfg4_bandit', kind='ws', mode='regression', dims=[119, 100, 100, 1], B=64, S=2,
         prior_init=[0.5, -0, -6], mixture=True, sigma=1.0, M=64, idx=0, full=False, flat_target=True,
         bandit_data=True),
    # Nets deeper than the reference's fixed three layers (BASELINE.json config 5 is 4 x 4096 hidden): the reference's
    # own BayesianLinear layers (networks.py:48-88) composed by hand, driven by the loop of networks.py:199-208.
    dict(name='deep5_small_mix', kind='ws', mode='classification', dims=[16, 12, 10, 12, 10, 4], B=6, S=3,
         prior_init=[0.5, -0, -6], mixture=True, sigma=1.0, M=4, idx=1, full=True, img=(1, 4, 4)),
    # a scaled config 5 that runs the batch-resident kernels (batch >= 384, widths >= 256) with ragged tails:
    # batch 640 = 512 + 128, widths that are not multiples of 128, three samples (one odd sample group)
    dict(name='deep5_cfg5_scaled', kind='ws', mode='classification', dims=[320, 384, 264, 392, 256, 10], B=640, S=3,
         prior_init=[0.5, -0, -6], mixture=True, sigma=1.0, M=4, idx=0, full=False, randn_x=True),
]
MU_INIT, RHO_INIT = [-0.2, 0.2], [-5, -4]


def load_reference(ref_root):
    """Import the reference's config.py and networks.py under their own names."""
    sys.path.insert(0, ref_root)
    for name in ('config', 'networks'):
        sys.modules.pop(name, None)
    spec = importlib.util.spec_from_file_location('config', os.path.join(ref_root, 'config.py'))
    cfg = importlib.util.module_from_spec(spec); sys.modules['config'] = cfg; spec.loader.exec_module(cfg)
    spec = importlib.util.spec_from_file_location('networks', os.path.join(ref_root, 'networks.py'))
    net = importlib.util.module_from_spec(spec); sys.modules['networks'] = net; spec.loader.exec_module(net)
    return net


def make_data(case):
    """Synthetic inputs of the config's shape, drawn after manual_seed(SEED_DATA)."""
    torch.manual_seed(SEED_DATA)
    B, dims = case['B'], case['dims']
    if case['mode'] == 'classification' and case.get('randn_x'):
        x = torch.randn(B, dims[0])          # config 5: x ~ N(0,1) [B, d_in]
        y = torch.randint(0, dims[-1], (B,))
    elif case['mode'] == 'classification':
        x = torch.rand(B, *case['img'])
        y = torch.randint(0, dims[-1], (B,))
    elif case.get('reg_data'):
        xs = torch.rand(B, 1) * 0.6
        e = torch.randn(B, 1) * 0.02
        y = xs + 0.3 * torch.sin(2 * np.pi * (xs + e)) + 0.3 * torch.sin(4 * np.pi * (xs + e)) + e
        x = xs
    elif case.get('bandit_data'):
        x = (torch.rand(B, dims[0]) < 0.19).float()
        y = torch.tensor([-35., 0., 5.])[torch.randint(0, 3, (B,))]
    else:
        x = torch.randn(B, dims[0])
        y = torch.randn(B) if case.get('flat_target') else torch.randn(B, dims[-1])
    return x, y


def build_deep_reference(refnet, case):
    """A network of len(dims)-1 reference BayesianLinear layers (networks.py:48-88) with ReLU between them and
    the reference's own sample_elbo / get_nll (networks.py:183-209) bound to it: BayesianNetwork.__init__ is
    fixed at three layers (networks.py:160-164), everything else of the class is depth-agnostic once
    forward / log_prior / log_variational_posterior walk the layer list."""
    dims = case['dims']

    class DeepReference(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.mode, self.local_reparam = case['mode'], False
            self.input_shape, self.classes = dims[0], dims[-1]
            self.n_layers = len(dims) - 1
            for i in range(self.n_layers):          # constructor draw order = layer order (networks.py:53-58)
                setattr(self, f'l{i + 1}', refnet.BayesianLinear(dims[i], dims[i + 1], MU_INIT, RHO_INIT,
                                                                 case['prior_init'], case['mixture']))

        def layers(self):
            return [getattr(self, f'l{i + 1}') for i in range(self.n_layers)]

        def forward(self, x, sample=False):          # networks.py:166-172 with the layer list
            if self.mode == 'classification':
                x = x.view(-1, self.input_shape)
            for i, l in enumerate(self.layers()):
                x = l(x, sample)
                if i + 1 < self.n_layers:
                    x = torch.relu(x)
            return x

        def log_prior(self):                         # networks.py:174-175
            return sum(l.log_prior for l in self.layers())

        def log_variational_posterior(self):         # networks.py:177-178
            return sum(l.log_variational_posterior for l in self.layers())

        get_nll = refnet.BayesianNetwork.get_nll              # networks.py:183-190, unmodified
        sample_elbo = refnet.BayesianNetwork.sample_elbo      # networks.py:192-209, unmodified
    return DeepReference()


def run_case(refnet, case):
    dims = case['dims']
    if len(dims) != 4:
        torch.manual_seed(SEED_PARAMS)
        return record_case(case, build_deep_reference(refnet, case))
    model_params = dict(input_shape=dims[0], classes=dims[-1], batch_size=case['B'], hidden_units=dims[1],
                        mode=case['mode'], mu_init=MU_INIT, rho_init=RHO_INIT, prior_init=case['prior_init'],
                        mixture_prior=case['mixture'], local_reparam=(case['kind'] == 'lr'))
    torch.manual_seed(SEED_PARAMS)
    return record_case(case, refnet.BayesianNetwork(model_params))


def record_case(case, net):
    dims = case['dims']
    x, y = make_data(case)
    beta = 2 ** (case['M'] - (case['idx'] + 1)) / (2 ** case['M'] - 1)

    outs = []
    orig_forward = net.forward

    def recording_forward(inp, sample=False):
        o = orig_forward(inp, sample)
        outs.append(o.detach().clone())
        return o
    net.forward = recording_forward

    net.train()
    net.zero_grad()
    torch.manual_seed(SEED_EPS)
    if case['kind'] == 'lr':
        info = net.sample_elbo_lr(x, y, beta, case['S'], sigma=case['sigma'])
    else:
        info = net.sample_elbo(x, y, beta, case['S'], sigma=case['sigma'])
    info[0].backward()

    rec = {}
    meta = {k: v for k, v in case.items()}
    meta.update(beta=beta, seeds=[SEED_PARAMS, SEED_DATA, SEED_EPS], mu_init=MU_INIT, rho_init=RHO_INIT,
                torch=torch.__version__, slice=SLICE)
    rec['loss'] = info[0].detach().numpy().reshape(-1)
    if case['kind'] == 'lr':
        rec['kl'] = info[1].detach().numpy().reshape(-1)
        rec['nll'] = info[2].detach().numpy().reshape(-1)
    else:
        rec['log_prior'] = info[1].detach().numpy().reshape(-1)
        rec['log_post'] = info[2].detach().numpy().reshape(-1)
        rec['nll'] = info[3].detach().numpy().reshape(-1)
    out = torch.stack(outs).numpy()
    names = []
    for li, layer in enumerate([getattr(net, f'l{i + 1}') for i in range(len(dims) - 1)]):
        for pn in ('weight_mu', 'weight_rho', 'bias_mu', 'bias_rho'):
            p = getattr(layer, pn)
            key = f'l{li + 1}.{pn}'
            names.append(key)
            g = p.grad.detach().numpy()
            pv = p.detach().numpy()
            rec[f'psum.{key}'] = np.array([pv.astype(np.float64).sum()])
            rec[f'gsum.{key}'] = np.array([g.astype(np.float64).sum()])
            rec[f'gnorm.{key}'] = np.array([np.sqrt((g.astype(np.float64) ** 2).sum())])
            rec[f'gmax.{key}'] = np.array([np.abs(g).max()])
            if case['full']:
                rec[f'param.{key}'] = pv
                rec[f'grad.{key}'] = g
            else:
                rec[f'gslice.{key}'] = g.reshape(-1)[::SLICE].copy()
    rec['x'] = x.numpy() if case['full'] else x.numpy().reshape(-1)[::SLICE].copy()
    rec['y'] = y.numpy()
    if case['full']:
        rec['outputs'] = out
    else:
        rec['outputs_slice'] = out.reshape(-1)[::7].copy()
        rec['outputs_sum'] = np.array([out.astype(np.float64).sum()])

    # eval-mode behaviour (networks.py:74-86): mean weights, and sampled prediction without log-probs
    if case['full']:
        net.forward = orig_forward
        net.eval()
        with torch.no_grad():
            if case['kind'] == 'ws':
                rec['eval_mean_out'] = net(x).numpy()
                torch.manual_seed(SEED_EPS + 1)
                rec['eval_sampled_out'] = net(x, sample=True).numpy()
                meta['eval_log_prior_type'] = type(net.l1.log_prior).__name__
                # layer-level calculate_log_probs in eval mode without sampling (w = mu)
                h = x.view(-1, dims[0]) if case['mode'] == 'classification' else x
                net.l1(h, False, True)
                rec['l1_eval_logp'] = np.array([float(net.l1.log_prior), float(net.l1.log_variational_posterior)])
            else:
                # the LR eval branch with sample=False raises AttributeError in the reference
                # (networks.py:131), so only the sampled prediction is pinned
                torch.manual_seed(SEED_EPS + 1)
                rec['eval_sampled_out'] = net(x, sample=True).numpy()
    rec['meta'] = np.array(json.dumps(meta))
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--reference', default=os.environ.get('BNN_REFERENCE_PATH', '/root/reference'))
    ap.add_argument('--out', default=os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden'))
    ap.add_argument('--only', nargs='*', default=None, help='case names to (re)generate; default: all')
    args = ap.parse_args()
    out_dir = os.path.abspath(args.out)
    os.makedirs(out_dir, exist_ok=True)
    torch.set_num_threads(1)            # fixed reduction order => reproducible fixtures
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as scratch:
        os.chdir(scratch)               # the reference writes ./runs and ./saved_models
        try:
            refnet = load_reference(args.reference)
            for case in CASES:
                if args.only and case['name'] not in args.only:
                    continue
                rec = run_case(refnet, case)
                path = os.path.join(out_dir, case['name'] + '.npz')
                np.savez_compressed(path, **rec)
                print(f"{case['name']:22s} loss={rec['loss'][0]:.6f}  -> {os.path.relpath(path, cwd)}"
                      f" ({os.path.getsize(path)} bytes)")
        finally:
            os.chdir(cwd)


if __name__ == '__main__':
    main()
