"""GPU tests of the fused head kernels (csrc/bbb_head.cu: bbb_head_fwd and the head backward behind
bbb_linear_bwd) against the oracle, at shapes the fixtures do not reach: batches above one cluster block, several
samples per pass / several passes, widths 1..16, both likelihoods, both arithmetic modes.

Tolerance: the head is exact fp32 in both modes, so <= 1e-5 relative on everything it produces; in TF32 mode the
layers below it are tensor-core contractions and the stated TF32 bound (5e-3) applies downstream of them."""
import numpy as np
import pytest
import torch

import bnn_b200
from bnn_b200 import functional as F
from oracle import bbb_oracle as O

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _case(dims, B, S, mode, seed):
    torch.manual_seed(seed)
    layers = O.init_layers(dims, [-0.2, 0.2], [-5, -4])
    x = torch.randn(B, dims[0])
    y = torch.randint(0, dims[-1], (B,)) if mode == 'classification' else torch.randn(B, dims[-1])
    eps = O.draw_eps(dims, S)
    return layers, x, y, eps


def _run_oracle(layers, x, y, eps, prior_init, mixture, beta, mode, sigma):
    ref = [tuple(p.clone().requires_grad_(True) for p in layer) for layer in layers]
    info = O.train_step(x, y, ref, O.make_prior(prior_init, mixture), eps, beta, mode, sigma)
    return [float(t.detach().reshape(-1)[0]) for t in info], [[p.grad.numpy() for p in layer] for layer in ref]


def _run_gpu(layers, x, y, eps, prior_init, mixture, beta, mode, sigma, tf32):
    dev_layers = [tuple(p.to(DEV).requires_grad_(True) for p in layer) for layer in layers]
    S = len(eps)
    bnn_b200.rng.set_injected_eps([t for per in eps for pair in per for t in pair])
    with bnn_b200.eps_mode('injected'):
        info = F.fused_elbo(x.to(DEV), y.to(DEV), beta, S, sigma, mode, F.make_prior(prior_init, mixture), dev_layers,
                            tf32=tf32)
    info[0].backward()
    return ([float(t.detach().reshape(-1)[0]) for t in info],
            [[p.grad.cpu().numpy() for p in layer] for layer in dev_layers])


@pytest.mark.parametrize('dims,B,S,mode', [
    ((8, 12, 1), 1, 1, 'regression'),            # a single row, a single output
    ((16, 20, 3), 5, 3, 'classification'),       # several samples in one pass of the backward
    ((12, 400, 1), 128, 5, 'regression'),        # cfg1's head with its sample count: three passes
    ((20, 64, 16), 300, 2, 'classification'),    # widest head; batch above 256: chunked passes, ragged last block
    ((20, 100, 10), 700, 3, 'classification'),   # several cluster row blocks per sample
    ((24, 36, 7), 130, 2, 'regression'),         # out not a multiple of 4, batch just above one pass of two samples
])
@pytest.mark.parametrize('mixture', [True, False])
def test_head_step_matches_oracle_exact_mode(dims, B, S, mode, mixture):
    prior_init = [0.5, 0, -6] if mixture else [1.0]
    layers, x, y, eps = _case(dims, B, S, mode, 17 * B + S)
    assert F.head_eligible(tuple(layers[-1]))
    want, gw = _run_oracle(layers, x, y, eps, prior_init, mixture, 0.37, mode, 0.7)
    got, gg = _run_gpu(layers, x, y, eps, prior_init, mixture, 0.37, mode, 0.7, tf32=False)
    for a, b in zip(got, want):
        assert abs(a - b) <= 1e-5 * abs(b) + 1e-6, (got, want)
    for lg, lw in zip(gg, gw):
        for a, b in zip(lg, lw):
            assert np.abs(a - b).max() <= 2e-5 * np.abs(b).max() + 1e-7, (np.abs(a - b).max(), np.abs(b).max())


def test_head_step_tf32_mode():
    """In TF32 mode the head itself stays exact fp32; the hidden layers below it are tensor-core contractions."""
    dims, B, S = (64, 128, 10), 128, 2
    layers, x, y, eps = _case(dims, B, S, 'classification', 5)
    want, gw = _run_oracle(layers, x, y, eps, [0.5, 0, -6], True, 0.5, 'classification', 1.0)
    got, gg = _run_gpu(layers, x, y, eps, [0.5, 0, -6], True, 0.5, 'classification', 1.0, tf32=True)
    assert abs(got[1] - want[1]) <= 1e-5 * abs(want[1]) and abs(got[2] - want[2]) <= 1e-5 * abs(want[2])
    assert abs(got[0] - want[0]) <= 5e-3 * abs(want[0])
    for lg, lw in zip(gg, gw):
        for a, b in zip(lg, lw):
            assert np.abs(a - b).max() <= 5e-3 * np.abs(b).max() + 1e-7


def test_head_forward_is_deterministic_and_counter_resets():
    """No atomics on tensors: two runs give bit-identical outputs and gradients; the completion counter in the
    workspace is left at zero (the C ABI promises it), so a caller may reuse it."""
    dims, B, S = (32, 48, 10), 200, 2
    layers, x, y, eps = _case(dims, B, S, 'classification', 3)
    r1 = _run_gpu(layers, x, y, eps, [0.5, 0, -6], True, 0.5, 'classification', 1.0, tf32=False)
    r2 = _run_gpu(layers, x, y, eps, [0.5, 0, -6], True, 0.5, 'classification', 1.0, tf32=False)
    assert r1[0][3] == r2[0][3]          # nll: fp64 accumulation of fp32 CTA sums, rounded to fp32
    for la, lb in zip(r1[1], r2[1]):
        for a, b in zip(la, lb):
            assert np.array_equal(a, b)


def test_head_fwd_rejects_unsupported_shapes():
    lib = bnn_b200._lib.lib()
    L = bnn_b200._lib
    t = torch.zeros(4096, device=DEV)
    acc = torch.zeros(4, dtype=torch.float64, device=DEV)
    rng = L.Rng(1, 0, 0, 0, None)
    prior = F.make_prior([1.0], False)
    import ctypes as C
    # out = 17 is wider than a head
    r = lib.bbb_head_fwd(t.data_ptr(), 0, t.data_ptr(), t.data_ptr(), t.data_ptr(), t.data_ptr(), None, None,
                         C.byref(rng), C.byref(prior), 1, 4, 8, 17, L.F_SAMPLE, L.NLL_NONE, None, 1.0, 1.0,
                         t.data_ptr(), None, None, None, None, 0.5, None, None, None, None)
    assert r == -3 and b'bbb_head_fwd' in lib.bbb_last_error_string()
    # ELBO assembly without a completion counter
    r = lib.bbb_head_fwd(t.data_ptr(), 0, t.data_ptr(), t.data_ptr(), t.data_ptr(), t.data_ptr(), None, None,
                         C.byref(rng), C.byref(prior), 1, 4, 8, 2, L.F_SAMPLE | L.F_LOGPROB, L.NLL_GAUSS, t.data_ptr(),
                         1.0, 1.0, t.data_ptr(), None, acc.data_ptr(), acc[1:].data_ptr(), acc[2:].data_ptr(), 0.5,
                         None, t.data_ptr(), None, None)
    assert r == -1
