"""Parity checks shared by the CPU host-logic suite (fake C ABI, device='cpu') and the GPU suite
(real libbbb.so, device='cuda').  Each check drives the PUBLIC API the reference's callers use --
BayesianNetwork / BayesianLinear[LR] from the drop-in `networks` module -- in the deterministic
'reference' eps mode and compares with the fixtures recorded from the reference."""
import numpy as np
import torch

import bnn_b200
import networks
from tests.golden_util import PNAMES

# fp32 bound of BASELINE.json north_star: <= 1e-5 relative (max-norm per tensor for gradients, plus the
# reference's own cancellation round-off, golden_util.Case.cancel_floor)
RTOL = 1e-5


def build_net(case, device, **extra):
    mp = case.model_params()
    mp.update(extra)
    net = networks.BayesianNetwork(mp)
    with torch.no_grad():
        for li, layer in enumerate(net.layers()):
            for pn, p in zip(PNAMES, case.layers[li]):
                getattr(layer, pn).copy_(p)
    return net.to(device)


def net_grads(net):
    return [[getattr(l, pn).grad.detach().cpu().numpy() for pn in PNAMES] for l in net.layers()]


def check_train_step(case, device, fused=True, rtol=RTOL, rtol_gemm=None, **extra):
    """zero_grad + sample_elbo[_lr] + backward through the network-level path.
    rtol bounds the log-prob / KL scalars; rtol_gemm (default rtol) bounds everything downstream of a GEMM
    (NLL, loss, gradients) -- looser in the TF32 tensor-core mode."""
    rg = rtol if rtol_gemm is None else rtol_gemm
    net = build_net(case, device, fused=fused, **extra)
    net.train()
    net.zero_grad()
    x, y = case.x.to(device), case.y.to(device)
    with bnn_b200.eps_mode('reference'):
        torch.manual_seed(case.meta['seeds'][2])
        if case.lr:
            info = net.sample_elbo_lr(x, y, case.beta, case.S, sigma=case.sigma)
        else:
            info = net.sample_elbo(x, y, case.beta, case.S, sigma=case.sigma)
    assert len(info) == (3 if case.lr else 4)
    assert tuple(info[0].shape) == (1,) and info[1].dim() == 0 and tuple(info[-1].shape) == (1,)
    info[0].backward()
    case.check_scalar('loss', info[0].detach().cpu(), rg)
    if case.lr:
        case.check_scalar('kl', info[1].detach().cpu(), rtol)
        case.check_scalar('nll', info[2].detach().cpu(), rg * 10)
    else:
        case.check_scalar('log_prior', info[1].detach().cpu(), rtol)
        case.check_scalar('log_post', info[2].detach().cpu(), rtol)
        case.check_scalar('nll', info[3].detach().cpu(), rg * 10)
    return case.check_grads(net_grads(net), rg, allow_cancel_floor=True)


def check_layerwise_train_step(case, device, rtol=RTOL):
    """The reference's own loop (networks.py:199-208) written against the LAYER-level API:
    net(x, sample=True) per sample, net.log_prior() / log_variational_posterior() / kl_cost()."""
    net = build_net(case, device)
    net.train()
    net.zero_grad()
    x, y = case.x.to(device), case.y.to(device)
    S = case.S
    outs = []
    with bnn_b200.eps_mode('reference'):
        torch.manual_seed(case.meta['seeds'][2])
        a = torch.zeros(S, device=device)
        b = torch.zeros(S, device=device)
        nll = torch.zeros(1, device=device)
        for i in range(S):
            out = net(x, sample=True)
            outs.append(out.detach().cpu().numpy())
            if case.lr:
                a[i] = net.kl_cost()
            else:
                a[i] = net.log_prior()
                b[i] = net.log_variational_posterior()
            nll = nll + net.get_nll(out, y, case.sigma)
    if case.lr:
        loss = case.beta * a.mean() + nll / S
    else:
        loss = case.beta * b.mean() - case.beta * a.mean() + nll / S
    loss.backward()
    case.check_scalar('loss', loss.detach().cpu(), rtol)
    case.check_outputs(np.stack(outs), rtol)
    return case.check_grads(net_grads(net), rtol, allow_cancel_floor=True)


def check_eval_modes(case, device):
    """networks.py:74-86: eval -> mean weights and int-0 log-probs; sample=True in eval draws weights
    but skips the log-prob branch; calculate_log_probs=True evaluates them at w = mu."""
    net = build_net(case, device)
    net.eval()
    x = case.x.to(device)
    with torch.no_grad(), bnn_b200.eps_mode('reference'):
        if not case.lr:
            out = net(x)
            np.testing.assert_allclose(out.cpu().numpy(), case.z['eval_mean_out'], rtol=1e-5, atol=1e-6)
            assert net.l1.log_prior == 0 and isinstance(net.l1.log_prior, int)
            assert net.log_prior() == 0
        torch.manual_seed(case.meta['seeds'][2] + 1)
        out = net(x, sample=True)
        np.testing.assert_allclose(out.cpu().numpy(), case.z['eval_sampled_out'], rtol=1e-5, atol=2e-6)
        if not case.lr:
            h = x.view(-1, case.dims[0])
            net.l1(h, False, True)
            got = [float(net.l1.log_prior), float(net.l1.log_variational_posterior)]
            np.testing.assert_allclose(got, case.z['l1_eval_logp'], rtol=2e-6)


def check_state_dict(case, device):
    """State-dict keys/shapes the reference's load_model_utils.py / weight_pruning.py rely on."""
    net = build_net(case, device)
    keys = list(net.state_dict().keys())
    assert keys == [f'l{i + 1}.{pn}' for i in range(len(case.layers)) for pn in PNAMES]
    d = case.dims
    want = (d[0], d[1]) if case.lr else (d[1], d[0])
    assert tuple(net.l1.weight_mu.shape) == want
    assert all(('mu' in n) or ('rho' in n) for n, _ in net.named_parameters())


def check_batched_prediction(case, device, samples=4, rtol=1e-5):
    """net.sample_predict (one launch per layer for all sampled forwards) and net.predict_proba against the ORACLE:
    `samples` sampled forwards of the reference's layers (networks.py:73-88 / 116-138) with the reference's eps draw
    order, softmax-averaged as class_task.py:81-87 does.  Also == the repo's own loop of
    net(x, sample=True) calls in eval mode (the call pattern of the reference's predict / evaluate)."""
    from oracle import bbb_oracle as O
    net = build_net(case, device)
    net.eval()
    x = case.x.to(device)
    with torch.no_grad():
        with torch.random.fork_rng():
            torch.manual_seed(123)
            eps = O.draw_eps(case.dims, samples, batch=case.B, local_reparam=case.lr)
        if case.lr:
            want = torch.stack([O.mlp_forward_lr(case.x, case.layers, case.prior[1], eps[s], case.mode, calc_kl=False)[0]
                                for s in range(samples)])
        else:
            want = torch.stack([O.mlp_forward(case.x, case.layers, case.prior, eps[s], case.mode,
                                              calc_log_probs=False)[0] for s in range(samples)])
    with torch.no_grad(), bnn_b200.eps_mode('reference'):
        torch.manual_seed(123)
        loop = torch.stack([net(x, sample=True) for _ in range(samples)])
        torch.manual_seed(123)
        got = net.sample_predict(x, samples)
    assert tuple(got.shape) == tuple(want.shape)
    np.testing.assert_allclose(got.cpu().numpy(), want.numpy(), rtol=rtol, atol=2e-6)
    np.testing.assert_allclose(loop.cpu().numpy(), want.numpy(), rtol=rtol, atol=2e-6)
    if case.mode == 'classification':
        with torch.no_grad(), bnn_b200.eps_mode('reference'):
            torch.manual_seed(123)
            p = net.predict_proba(x, samples)
        np.testing.assert_allclose(p.sum(-1).cpu().numpy(), 1.0, rtol=1e-5)
        np.testing.assert_allclose(p.cpu().numpy(), torch.softmax(want, -1).mean(0).numpy(), rtol=1e-4, atol=1e-6)


def check_snr_pruning(case, device, drop=0.4):
    """net.snr() / net.prune_weights() (bbb_snr, bbb_snr_prune) against the oracle's restatement of
    weight_pruning.py:81-115.  Elements whose SNR sits within fp32 round-off of the threshold may fall either way."""
    from oracle import bbb_oracle as O
    net = build_net(case, device)
    want = O.network_snrs(case.layers)
    got = net.snr().cpu()
    np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=1e-5, atol=1e-5)
    frac = net.prune_weights(want, drop)
    pruned = O.prune_layers(case.layers, want, drop)
    thr = float(np.percentile(want.numpy(), 100 * drop))
    n_total, n_kept = 0, 0
    for li, layer in enumerate(net.layers()):
        snr_w = O.compute_snr(case.layers[li][0], case.layers[li][1])
        snr_b = O.compute_snr(case.layers[li][2], case.layers[li][3])
        for pi, pn in enumerate(PNAMES):
            g, w = getattr(layer, pn).detach().cpu().numpy(), pruned[li][pi].numpy()
            edge = np.abs((snr_w if pi < 2 else snr_b).numpy() - thr) < 1e-4
            assert np.array_equal(g[~edge], w[~edge]), (li, pn)
            n_total += g.size
        n_kept += int((pruned[li][0] != 0).sum()) + int((pruned[li][2] != 0).sum())
    assert abs(frac - n_kept / (n_total / 2)) < 0.02
    assert abs(frac - (1 - drop)) < 0.05
