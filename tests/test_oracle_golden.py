"""CPU tests: the oracle (torch restatement + numpy closed form) against fixtures recorded from
the live reference (oracle/make_golden.py).  This is what pins the oracle."""
import numpy as np
import pytest
import torch

from oracle import bbb_oracle as O
from oracle import closed_form as CF
from tests.golden_util import Case, SMALL, SMALL_LR, BIG, BIG_LR, DEEP_SMALL, DEEP_BIG

RTOL_F32 = 2e-6      # same op order as the reference => round-off only
RTOL_CF = 1e-5       # float64 closed form vs the reference's fp32 autograd; gradients additionally get the
                     # reference's own cancellation round-off (golden_util.Case.cancel_floor)


def _leaf(layers):
    return [tuple(p.clone().requires_grad_(True) for p in layer) for layer in layers]


@pytest.mark.parametrize('name', SMALL + BIG + DEEP_SMALL + DEEP_BIG)
def test_torch_oracle_weight_sampling(name):
    c = Case(name)
    torch.set_num_threads(1)
    layers = _leaf(c.layers)
    loss, lp, lq, nll = O.train_step(c.x, c.y, layers, c.prior, c.eps, c.beta, c.mode, c.sigma)
    c.check_scalar('loss', loss.detach(), RTOL_F32)
    c.check_scalar('log_prior', lp.detach(), RTOL_F32)
    c.check_scalar('log_post', lq.detach(), RTOL_F32)
    c.check_scalar('nll', nll.detach(), RTOL_F32 * 10)
    c.check_grads([[p.grad.numpy() for p in layer] for layer in layers], RTOL_F32 * 5)


@pytest.mark.parametrize('name', SMALL_LR + BIG_LR)
def test_torch_oracle_local_reparam(name):
    c = Case(name)
    torch.set_num_threads(1)
    layers = _leaf(c.layers)
    loss, kl, nll = O.train_step(c.x, c.y, layers, c.prior[1], c.eps, c.beta, c.mode, c.sigma, local_reparam=True)
    c.check_scalar('loss', loss.detach(), RTOL_F32)
    c.check_scalar('kl', kl.detach(), RTOL_F32)
    c.check_scalar('nll', nll.detach(), RTOL_F32 * 10)
    c.check_grads([[p.grad.numpy() for p in layer] for layer in layers], RTOL_F32 * 5)


@pytest.mark.parametrize('name', SMALL + DEEP_SMALL)
def test_torch_oracle_eval_modes(name):
    c = Case(name)
    with torch.no_grad():
        out, lp, lq = O.mlp_forward(c.x, c.layers, c.prior, None, c.mode, calc_log_probs=False)
        assert lp == 0 and lq == 0 and c.meta['eval_log_prior_type'] == 'int'
        np.testing.assert_allclose(out.numpy(), c.z['eval_mean_out'], rtol=1e-5, atol=1e-6)
        out, _, _ = O.mlp_forward(c.x, c.layers, c.prior, c.eps_eval, c.mode, calc_log_probs=False)
        np.testing.assert_allclose(out.numpy(), c.z['eval_sampled_out'], rtol=1e-5, atol=1e-6)
        h = c.x.view(-1, c.dims[0])
        _, lp, lq = O.bayes_linear(h, c.layers[0], c.prior, None, None, True)
        np.testing.assert_allclose([float(lp), float(lq)], c.z['l1_eval_logp'], rtol=2e-6)


@pytest.mark.parametrize('name', SMALL_LR)
def test_torch_oracle_lr_sampled_eval(name):
    c = Case(name)
    with torch.no_grad():
        out, _ = O.mlp_forward_lr(c.x, c.layers, c.prior[1], c.eps_eval, c.mode, calc_kl=False)
    np.testing.assert_allclose(out.numpy(), c.z['eval_sampled_out'], rtol=1e-5, atol=1e-6)


def _np_layers(layers):
    return [tuple(p.double().numpy() for p in layer) for layer in layers]


def _np_eps(eps):
    return [[(a.double().numpy(), b.double().numpy()) for a, b in per] for per in eps]


@pytest.mark.parametrize('name', SMALL + DEEP_SMALL + ['cfg4_bandit', 'cfg1_reg_mix'])
def test_closed_form_weight_sampling(name):
    c = Case(name)
    r = CF.elbo_step(c.x.double().numpy(), c.y.numpy() if c.mode == 'classification' else c.y.double().numpy(),
                     _np_layers(c.layers), c.prior, _np_eps(c.eps), c.beta, c.mode, c.sigma)
    c.check_scalar('loss', r['loss'], RTOL_CF)
    c.check_scalar('log_prior', r['log_prior'], RTOL_CF)
    c.check_scalar('log_post', r['log_post'], RTOL_CF)
    c.check_scalar('nll', r['nll'], RTOL_CF)
    c.check_outputs(r['outputs'], RTOL_CF)
    c.check_grads(r['grads'], RTOL_CF, allow_cancel_floor=True)


@pytest.mark.parametrize('name', SMALL_LR)
def test_closed_form_local_reparam(name):
    c = Case(name)
    r = CF.elbo_step_lr(c.x.double().numpy(), c.y.numpy() if c.mode == 'classification' else c.y.double().numpy(),
                        _np_layers(c.layers), c.prior[1], _np_eps(c.eps), c.beta, c.mode, c.sigma)
    c.check_scalar('loss', r['loss'], RTOL_CF)
    c.check_scalar('kl', r['kl'], RTOL_CF)
    c.check_scalar('nll', r['nll'], RTOL_CF)
    c.check_outputs(r['outputs'], RTOL_CF)
    c.check_grads(r['grads'], RTOL_CF, allow_cancel_floor=True)


def test_beta_schedule_sums_to_one():
    for M in (8, 64, 468):
        assert abs(sum(O.elbo_beta(M, i) for i in range(M)) - 1.0) < 1e-12
