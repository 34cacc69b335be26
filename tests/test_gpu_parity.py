"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the public API and
the C ABI, against the fixtures recorded from the reference and against the oracle.

Tolerances: fp32 FMA path <= 1e-5 relative (max-norm per tensor for gradients, plus the reference's
own cancellation round-off -- see golden_util.Case.cancel_floor); stated per test otherwise."""
import math

import numpy as np
import pytest
import torch

import bnn_b200
from bnn_b200 import functional as F
from oracle import bbb_oracle as O
from oracle import closed_form as CF
from tests import parity_cases as PC
from tests.golden_util import Case, SMALL, SMALL_LR, BIG, BIG_LR, DEEP_SMALL, DEEP_BIG, PNAMES

pytestmark = pytest.mark.gpu
DEV = 'cuda'


@pytest.mark.parametrize('name', SMALL + SMALL_LR + BIG + BIG_LR + DEEP_SMALL + DEEP_BIG)
@pytest.mark.parametrize('fused', [True, False])
def test_train_step_matches_reference(name, fused):
    PC.check_train_step(Case(name), DEV, fused=fused)


@pytest.mark.parametrize('name', SMALL + SMALL_LR + DEEP_SMALL + ['cfg4_bandit'])
def test_layer_level_api_matches_reference(name):
    PC.check_layerwise_train_step(Case(name), DEV)


@pytest.mark.parametrize('name', SMALL + SMALL_LR + DEEP_SMALL)
def test_eval_modes(name):
    PC.check_eval_modes(Case(name), DEV)


@pytest.mark.parametrize('name', SMALL + SMALL_LR + DEEP_SMALL + ['cfg4_bandit'])
def test_batched_prediction(name):
    PC.check_batched_prediction(Case(name), DEV)


@pytest.mark.parametrize('name', ['small_cls_mix', 'deep5_small_mix', 'cfg2_mnist_mix'])
def test_snr_pruning_matches_oracle(name):
    PC.check_snr_pruning(Case(name), DEV)


def test_native_library_is_loaded():
    import ctypes
    assert isinstance(bnn_b200._lib.lib(), ctypes.CDLL)
    with open('/proc/self/maps') as f:
        assert 'libbbb.so' in f.read()


# ---------------------------------------------------------------------------------------------
# stand-alone reductions against the oracle
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize('n', [1, 3, 4, 1000, 1200 * 1200 + 3])
@pytest.mark.parametrize('mixture', [True, False])
def test_logprob_reduce_matches_oracle(n, mixture):
    torch.manual_seed(n)
    mu = torch.empty(n).uniform_(-0.2, 0.2)
    rho = torch.empty(n).uniform_(-5, -4)
    eps = torch.randn(n)
    prior_init = [0.5, 0, -6] if mixture else [1.0]
    pri_o = O.make_prior(prior_init, mixture)
    w_ref = mu + O.softplus(rho) * eps
    lp_ref = float(O.prior_log_prob(w_ref.double(), pri_o))
    lq_ref = float(O.posterior_log_prob(w_ref.double(), mu.double(), rho.double()))
    lp, lq, w = F.logprob_reduce(mu.to(DEV), rho.to(DEV), F.make_prior(prior_init, mixture), eps=eps.to(DEV),
                                 return_w=True)
    np.testing.assert_allclose(w.cpu().numpy(), w_ref.numpy(), rtol=1e-6, atol=1e-7)
    assert abs(float(lp) - lp_ref) <= 1e-5 * abs(lp_ref) + 1e-5
    assert abs(float(lq) - lq_ref) <= 1e-5 * abs(lq_ref) + 1e-5


@pytest.mark.parametrize('n', [1, 5, 4096, 784 * 1200])
def test_kl_gauss_matches_oracle(n):
    torch.manual_seed(n)
    mu = torch.empty(n).uniform_(-0.2, 0.2)
    rho = torch.empty(n).uniform_(-5, -4)
    ref = float(O.gaussian_kl(mu.double(), O.softplus(rho.double()), torch.tensor(1.0, dtype=torch.float64)))
    got = float(F.kl_gauss(mu.to(DEV), rho.to(DEV), 1.0))
    assert abs(got - ref) <= 1e-5 * abs(ref)


# ---------------------------------------------------------------------------------------------
# Philox mode: the stream is statistically normal, tiling-independent and regenerated exactly
# ---------------------------------------------------------------------------------------------
def test_philox_normal_moments_and_ks():
    from scipy import stats
    n = 1 << 22
    z = F.philox_normal(n, DEV, seed=1234, step=7, sample_idx=3, tensor_id=5).double().cpu().numpy()
    assert abs(z.mean()) < 5 / math.sqrt(n)
    assert abs(z.var() - 1) < 5 * math.sqrt(2 / n)
    assert abs(stats.skew(z)) < 5 * math.sqrt(6 / n)
    assert abs(stats.kurtosis(z)) < 5 * math.sqrt(24 / n)
    assert stats.kstest(z[: 1 << 20], 'norm').pvalue > 1e-3
    # disjoint coordinates give uncorrelated streams
    z2 = F.philox_normal(n, DEV, seed=1234, step=7, sample_idx=4, tensor_id=5).double().cpu().numpy()
    assert abs(np.corrcoef(z, z2)[0, 1]) < 5 / math.sqrt(n)
    z3 = F.philox_normal(n, DEV, seed=1234, step=8, sample_idx=3, tensor_id=5).double().cpu().numpy()
    assert abs(np.corrcoef(z, z3)[0, 1]) < 5 / math.sqrt(n)


def _philox_eps_for(dims, S, seed, step, sample_base=0, lr_batch=None):
    """Materialise the eps the kernels generate internally (tensor ids 2l / 2l+1)."""
    eps = []
    for s in range(S):
        per = []
        for l, (d_in, d_out) in enumerate(zip(dims[:-1], dims[1:])):
            nw = (lr_batch * d_out) if lr_batch else d_out * d_in
            ew = F.philox_normal(nw, DEV, seed, step, sample_base + s, 2 * l)
            eb = F.philox_normal(d_out, DEV, seed, step, sample_base + s, 2 * l + 1)
            per.append((ew.view(lr_batch, d_out) if lr_batch else ew.view(d_out, d_in), eb))
        eps.append(per)
    return eps


@pytest.mark.parametrize('name', ['small_cls_mix', 'small_bandit_bcast', 'cfg4_bandit', 'cfg2_mnist_mix',
                                  'small_lr_reg', 'cfg3_mnist_lr', 'deep5_small_mix'])
def test_philox_mode_equals_injecting_the_same_stream(name):
    """In-kernel eps (forward AND the backward's regeneration) == the fill kernel's stream injected
    through the parity path.  Covers aligned (vectorised) and ragged (119, 1, 9 wide) rows."""
    c = Case(name)
    seed, step = 99, 5
    x, y = c.x.to(DEV), c.y.to(DEV)
    net = PC.build_net(c, DEV)
    net.train()
    bnn_b200.manual_seed(seed, step)
    with bnn_b200.eps_mode('philox'):
        info = net.sample_elbo_lr(x, y, c.beta, c.S, c.sigma) if c.lr else net.sample_elbo(x, y, c.beta, c.S, c.sigma)
    info[0].backward()
    g1 = PC.net_grads(net)

    eps = _philox_eps_for(c.dims, c.S, seed, step, lr_batch=c.B if c.lr else None)
    # (a) bit-exact: the same stream injected through the parity path gives identical results
    net2 = PC.build_net(c, DEV)
    net2.train()
    bnn_b200.rng.set_injected_eps([t for per in eps for pair in per for t in pair])
    with bnn_b200.eps_mode('injected'):
        info2 = net2.sample_elbo_lr(x, y, c.beta, c.S, c.sigma) if c.lr else net2.sample_elbo(x, y, c.beta, c.S, c.sigma)
    info2[0].backward()
    # (the loss goes through fp64 atomics whose order is not fixed: equal to fp32 round-off; gradients bit-exact)
    assert abs(float(info2[0].detach()) - float(info[0].detach())) <= 1e-6 * abs(float(info[0].detach()))
    for ga, gb in zip(PC.net_grads(net2), g1):
        for a, b in zip(ga, gb):
            assert np.array_equal(a, b)
    # (b) against the float64 closed form.  With the e^-8 mixture component d(w R(w))/dw reaches 1/sigma2^2 =
    # 8.9e6, so the fp32 rounding of w alone (6e-8 relative) moves the gradient by up to ~3e-5 of its max-norm:
    # an fp32-vs-fp64 conditioning effect the reference shares; 5e-5 is the stated bound for this comparison.
    layers = [tuple(p.double().numpy() for p in layer) for layer in c.layers]
    eps_np = [[(a.double().cpu().numpy(), b.double().cpu().numpy()) for a, b in per] for per in eps]
    xx = c.x.double().numpy()
    yy = c.y.numpy() if c.mode == 'classification' else c.y.double().numpy()
    if c.lr:
        r = CF.elbo_step_lr(xx, yy, layers, c.prior[1], eps_np, c.beta, c.mode, c.sigma)
    else:
        r = CF.elbo_step(xx, yy, layers, c.prior, eps_np, c.beta, c.mode, c.sigma)
    assert abs(float(info[0].detach()) - r['loss']) <= 1e-5 * abs(r['loss'])
    for li in range(len(c.layers)):
        for pi in range(4):
            ref = r['grads'][li][pi]
            err = np.abs(g1[li][pi] - ref).max() / np.abs(ref).max()
            assert err <= 5e-5, (name, li, PNAMES[pi], err)


def test_philox_step_is_replayable_and_advances():
    c = Case('cfg4_bandit')
    x, y = c.x.to(DEV), c.y.to(DEV)
    net = PC.build_net(c, DEV)
    net.train()
    losses = []
    for seed_step in ((7, 0), (7, 0), (7, 1)):
        bnn_b200.manual_seed(*seed_step)
        net.zero_grad()
        info = net.sample_elbo(x, y, c.beta, c.S, c.sigma)
        info[0].backward()
        losses.append((float(info[0]), net.l2.weight_rho.grad.clone()))
    assert losses[0][0] == losses[1][0] and torch.equal(losses[0][1], losses[1][1])
    assert losses[0][0] != losses[2][0]


def test_sample_sharding_is_invariant():
    """G logical shards (disjoint global sample indices) reproduce the one-shard sums."""
    c = Case('small_cls_mix')
    x, y = c.x.to(DEV), c.y.to(DEV)
    S = 4

    def run(base, n):
        net = PC.build_net(c, DEV)
        net.train()
        bnn_b200.manual_seed(11, 3)
        bnn_b200.set_sample_base(base)
        try:
            outs, lps, lqs = F.mlp_forward(x.view(-1, c.dims[0]), [l.params() for l in net.layers()], net._prior, n)
            nll = sum(net.get_nll(outs[i], y) for i in range(n))
            (c.beta * (lqs.sum() - lps.sum()) + nll).backward()
        finally:
            bnn_b200.set_sample_base(0)
        return torch.cat([lps.detach(), lqs.detach()]), PC.net_grads(net)

    full_s, full_g = run(0, S)
    a_s, a_g = run(0, 2)
    b_s, b_g = run(2, 2)
    np.testing.assert_allclose(torch.cat([a_s[:2], b_s[:2], a_s[2:], b_s[2:]]).cpu().numpy(), full_s.cpu().numpy(),
                               rtol=1e-6)
    for li in range(len(c.layers)):
        for pi in range(4):
            tot = a_g[li][pi] + b_g[li][pi]
            assert np.abs(tot - full_g[li][pi]).max() <= 2e-6 * np.abs(full_g[li][pi]).max()


def test_philox_elbo_estimate_is_statistically_consistent():
    """Mean ELBO terms under Philox eps agree with torch-RNG eps within 5 standard errors."""
    c = Case('small_reg_mix')
    x, y = c.x.to(DEV), c.y.to(DEV)
    net = PC.build_net(c, DEV)
    net.train()
    S = 512
    params = [l.params() for l in net.layers()]
    with torch.no_grad():
        bnn_b200.manual_seed(2024, 0)
        with bnn_b200.eps_mode('philox'):
            o1, lp1, lq1 = F.mlp_forward(x, params, net._prior, S)
        torch.manual_seed(5)
        with bnn_b200.eps_mode('reference'):
            o2, lp2, lq2 = F.mlp_forward(x, params, net._prior, S)
    for a, b in ((lp1, lp2), (lq1, lq2), (o1.mean((1, 2)), o2.mean((1, 2)))):
        a, b = a.double().cpu().numpy(), b.double().cpu().numpy()
        se = math.sqrt(a.var() / S + b.var() / S)
        assert abs(a.mean() - b.mean()) < 5 * se


# ---------------------------------------------------------------------------------------------
# edge cases: ragged / degenerate shapes through the layer-level API against the oracle
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize('B,d_in,d_out', [(1, 1, 1), (1, 119, 100), (3, 5, 1), (65, 17, 67), (128, 64, 64), (2, 130, 3)])
@pytest.mark.parametrize('lr', [False, True])
def test_ragged_shapes_layer_level(B, d_in, d_out, lr):
    torch.manual_seed(B * 1000 + d_in)
    prior_init, mixture = ([1.0], False) if lr else ([0.5, 0, -6], True)
    cls = bnn_b200.BayesianLinearLR if lr else bnn_b200.BayesianLinear
    layer = cls(d_in, d_out, [-0.2, 0.2], [-5, -4], prior_init, mixture)
    ref_layer = tuple(p.detach().clone().requires_grad_(True) for p in layer.params())
    x = torch.randn(B, d_in)
    xr = x.clone().requires_grad_(True)
    torch.manual_seed(77)
    e1 = torch.randn(B, d_out) if lr else torch.randn(d_out, d_in)
    e2 = torch.randn(d_out)
    if lr:
        y_ref, kl = O.lr_linear(xr, ref_layer, 1.0, e1, e2)
        (y_ref.pow(2).sum() + 0.3 * kl).backward()
    else:
        y_ref, lp, lq = O.bayes_linear(xr, ref_layer, O.make_prior(prior_init, mixture), e1, e2)
        (y_ref.pow(2).sum() + 0.3 * lq - 0.3 * lp).backward()

    layer = layer.to(DEV).train()
    xg = x.to(DEV).requires_grad_(True)
    with bnn_b200.eps_mode('reference'):
        torch.manual_seed(77)
        y = layer(xg)
    if lr:
        (y.pow(2).sum() + 0.3 * layer.kl_cost).backward()
    else:
        (y.pow(2).sum() + 0.3 * layer.log_variational_posterior - 0.3 * layer.log_prior).backward()
    np.testing.assert_allclose(y.detach().cpu().numpy(), y_ref.detach().numpy(), rtol=2e-5, atol=2e-6)
    gx = xg.grad.cpu().numpy()
    assert np.abs(gx - xr.grad.numpy()).max() <= 2e-5 * max(1e-6, np.abs(xr.grad.numpy()).max())
    for p, pr in zip(layer.params(), ref_layer):
        g, gr = p.grad.cpu().numpy(), pr.grad.numpy()
        assert np.abs(g - gr).max() <= 5e-5 * np.abs(gr).max(), (np.abs(g - gr).max(), np.abs(gr).max())


def test_c_abi_rejects_bad_arguments():
    lib = bnn_b200._lib.lib()
    assert lib.bbb_kl_gauss(None, None, 1.0, 4, None, None) == -1
    assert b'null' in lib.bbb_last_error_string()
    t = torch.zeros(8, device=DEV)
    acc = torch.zeros(1, dtype=torch.float64, device=DEV)
    assert lib.bbb_kl_gauss(t.data_ptr(), t.data_ptr(), -1.0, 8, acc.data_ptr(), None) == -1
    assert lib.bbb_kl_gauss(t.data_ptr(), t.data_ptr(), 1.0, 0, acc.data_ptr(), None) == 0   # empty input is fine
    # BBB_F_RELU_OUT is a large-batch tensor-path feature: asked of any other path it is refused, not ignored
    L = bnn_b200._lib
    assert lib.bbb_linear_fwd_relu_out_supported(128, 64, 64, L.F_TF32) == 0
    assert lib.bbb_linear_fwd_relu_out_supported(512, 64, 64, L.F_TF32) == 1
    assert lib.bbb_linear_fwd_relu_out_supported(512, 64, 64, 0) == 0
    x, w, b, y = (torch.zeros(n, device=DEV) for n in (128 * 64, 64 * 64, 64, 128 * 64))
    rc = lib.bbb_linear_fwd(x.data_ptr(), 0, w.data_ptr(), w.data_ptr(), b.data_ptr(), b.data_ptr(), None, None, None, None,
                            1, 128, 64, 64, L.F_TF32 | L.F_RELU_OUT, y.data_ptr(), None, None, None)
    assert rc == -3 and b'BBB_F_RELU_OUT' in lib.bbb_last_error_string()


def test_empty_batch():
    c = Case('small_reg_mix')
    net = PC.build_net(c, DEV).eval()
    out = net(torch.zeros(0, c.dims[0], device=DEV), sample=True)
    assert tuple(out.shape) == (0, c.dims[-1])
