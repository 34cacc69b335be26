"""The reference's own task classes, UNMODIFIED, running on top of the drop-in `networks` / `config` modules
(reg_task.py, class_task.py and bandits.py run unchanged).

The reference tree is taken from baseline/_ref/ (staged by __graft_entry__.build(); travels to the GPU box)
or, in the build container only, from /root/reference.  On CPU the C ABI is replaced by the test double;
with -m gpu the same drivers run on the real kernels."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

from tests import fake_bbb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ref_dir():
    for d in (os.path.join(ROOT, 'baseline', '_ref'), os.environ.get('BNN_REFERENCE_PATH', '/root/reference')):
        if os.path.isfile(os.path.join(d, 'regression', 'reg_task.py')):
            return d
    return None


@pytest.fixture
def ref(tmp_path, monkeypatch):
    d = _ref_dir()
    if d is None:
        pytest.skip('reference tree not staged (baseline/_ref)')
    monkeypatch.chdir(tmp_path)                       # the task classes write ./runs and ./saved_models
    monkeypatch.syspath_prepend(d)
    monkeypatch.syspath_prepend(ROOT)                 # our networks.py / config.py shadow the reference's
    for m in [k for k in sys.modules if k.split('.')[0] in ('networks', 'config', 'utils', 'regression',
                                                             'classification', 'reinforcement_learning')]:
        monkeypatch.delitem(sys.modules, m)
    import networks
    assert os.path.dirname(os.path.abspath(networks.__file__)) == ROOT
    return d


def _drive_regression(dev, local_reparam):
    reg = importlib.import_module('regression.reg_task')
    import config
    assert config.DEVICE.type == dev
    params = dict(batch_size=16, num_batches=2, train_samples=3, test_samples=4, x_shape=1, y_shape=1,
                  noise_tolerance=0.1, lr=1e-3, save_dir='./saved_models', local_reparam=local_reparam,
                  hidden_units=24, mode='regression', mixture_prior=not local_reparam, mu_init=[-0.2, 0.2],
                  rho_init=[-5, -4], prior_init=[1.] if local_reparam else [0.5, -0, -6])
    task = reg.BNN_Regression('t', params)
    before = [p.detach().clone() for p in task.net.parameters()]
    g = torch.Generator().manual_seed(0)
    data = [(torch.randn(16, 1, generator=g), torch.randn(16, 1, generator=g)) for _ in range(2)]
    task.train_step(data)
    assert np.isfinite(task.epoch_loss)
    assert len(task.loss_info) == (3 if local_reparam else 4)
    assert all(not torch.equal(a, b) for a, b in zip(before, task.net.parameters()))
    y = task.evaluate(torch.linspace(-1, 1, 10).reshape(-1, 1))
    assert y.shape == (4, 10) and np.isfinite(y).all()
    assert np.abs(y[0] - y[1]).max() > 0              # evaluate() samples the posterior each time


def _drive_classification(dev, local_reparam):
    cls = importlib.import_module('classification.class_task')
    params = dict(lr=1e-3, hidden_units=32, mode='classification', batch_size=8, num_batches=3, train_samples=2,
                  test_samples=3, x_shape=16, classes=5, mu_init=[-0.2, 0.2], rho_init=[-5, -4],
                  prior_init=[1.] if local_reparam else [0.5, -0, -8], mixture_prior=not local_reparam,
                  save_dir='./saved_models', local_reparam=local_reparam)
    task = cls.BNN_Classification('t', params)
    g = torch.Generator().manual_seed(1)
    data = [(torch.rand(8, 1, 4, 4, generator=g), torch.randint(0, 5, (8,), generator=g)) for _ in range(3)]
    task.train_step(data)
    assert torch.isfinite(task.loss_info[0]).all()
    task.net.eval()
    with torch.no_grad():
        preds, probs = task.predict(data[0][0].to(dev))
    assert preds.shape == (8,) and probs.shape == (8, 5)
    np.testing.assert_allclose(probs.sum(1).cpu().numpy(), 1.0, rtol=1e-5)


def _drive_bandit(dev):
    ban = importlib.import_module('reinforcement_learning.bandits')
    rs = np.random.RandomState(0)
    x = (rs.rand(40, 9) < 0.3).astype(np.float32)
    y = (rs.rand(40) < 0.5).astype(np.int64)
    params = dict(n_samples=2, buffer_size=16, batch_size=4, num_batches=4, lr=1e-3, epsilon=0.0, hidden_units=10,
                  mode='regression', mixture_prior=True, mu_init=[-0.2, 0.2], rho_init=[-5, -4],
                  prior_init=[0.5, -0, -6])
    agent = ban.BNN_Bandit('t', params, x, y)       # raises KeyError('local_reparam') on the reference itself (App. B-1)
    for m in range(10):
        agent.update(m)
    assert len(agent.loss_info) == 4 and torch.isfinite(agent.loss_info[0]).all()
    assert len(agent.cumulative_regrets) == 11


def _drive_mc_dropout(dev):
    """The MC-dropout baselines (reg_task.py:142-183, class_task.py:185-240) construct networks.MLP_Dropout with no
    'dropout' key (regression) or 'dropout': True (classification); the reference hard-codes p = 0.5."""
    reg = importlib.import_module('regression.reg_task')
    cls = importlib.import_module('classification.class_task')
    rp = dict(batch_size=16, num_batches=2, test_samples=4, x_shape=1, y_shape=1, lr=1e-3, save_dir='./saved_models',
              hidden_units=24, mode='regression')
    task = reg.MCDropout_Regression('t', rp)
    g = torch.Generator().manual_seed(0)
    data = [(torch.randn(16, 1, generator=g), torch.randn(16, 1, generator=g)) for _ in range(2)]
    task.train_step(data)
    assert torch.isfinite(task.loss_info).all()
    cp = dict(lr=1e-3, hidden_units=32, mode='classification', batch_size=8, num_batches=3, test_samples=3, x_shape=16,
              classes=5, save_dir='./saved_models', dropout=True)
    ctask = cls.MCDropout_Classification('t', cp)
    drops = [m for m in ctask.net.modules() if isinstance(m, torch.nn.Dropout)]
    assert len(drops) == 2 and all(m.p == 0.5 for m in drops)
    ctask.net.train()
    out = ctask.net(torch.rand(8, 1, 4, 4).to(dev))
    assert out.shape == (8, 5) and torch.isfinite(out).all() and out.abs().max() > 0


def _state_dict_round_trip(dev, local_reparam):
    """a model trained by the drop-in loads into the REFERENCE's BayesianNetwork (load_model_utils.py:9-28
    does model.load_state_dict(torch.load(path))) and back: same keys, same [out,in] / [in,out] layouts, and the
    reference's mean-weight forward on the loaded parameters equals the drop-in's."""
    import importlib.util
    import networks as dropin                              # the drop-in (ROOT is first on sys.path)
    d = _ref_dir()
    spec = importlib.util.spec_from_file_location('ref_networks_original', os.path.join(d, 'networks.py'))
    refnet = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(refnet)                        # `from config import *` resolves to the drop-in config: DEVICE only
    mp = dict(input_shape=16, classes=5, batch_size=8, hidden_units=12, mode='classification', mu_init=[-0.2, 0.2],
              rho_init=[-5, -4], prior_init=[1.] if local_reparam else [0.5, -0, -6], mixture_prior=not local_reparam,
              local_reparam=local_reparam)
    torch.manual_seed(3)
    mine = dropin.BayesianNetwork(mp).to(dev)
    path = os.path.join(os.getcwd(), 'model.pt')
    torch.save(mine.state_dict(), path)
    ref = refnet.BayesianNetwork(mp)
    ref.load_state_dict(torch.load(path, map_location='cpu'))          # strict: keys and shapes must match exactly
    assert list(ref.state_dict().keys()) == list(mine.state_dict().keys())
    x = torch.rand(8, 1, 4, 4)
    if not local_reparam:                                  # (the reference's LR eval branch raises AttributeError: App. B-4)
        ref.eval(); mine.eval()
        with torch.no_grad():
            np.testing.assert_allclose(mine(x.to(dev)).cpu().numpy(), ref(x).numpy(), rtol=1e-5, atol=1e-6)
    # and back: a reference checkpoint into the drop-in
    torch.manual_seed(4)
    ref2 = refnet.BayesianNetwork(mp)
    torch.save(ref2.state_dict(), path)
    mine.load_state_dict(torch.load(path, map_location=dev))
    for (k, a), (_, b) in zip(mine.state_dict().items(), ref2.state_dict().items()):
        assert torch.equal(a.cpu(), b), k


def _cpu_double(monkeypatch):
    """No GPU here: the C ABI is the test double, which needs eps injected (the reference's own CPU draws)."""
    import bnn_b200
    fake_bbb.install(monkeypatch)
    monkeypatch.setattr(torch.cuda, 'is_available', lambda: False)
    monkeypatch.setattr(bnn_b200.rng._st(), 'mode', 'reference')


@pytest.mark.parametrize('local_reparam', [False, True])
def test_reference_regression_task_on_dropin_cpu_double(ref, monkeypatch, local_reparam):
    _cpu_double(monkeypatch)
    _drive_regression('cpu', local_reparam)


@pytest.mark.parametrize('local_reparam', [False, True])
def test_reference_classification_task_on_dropin_cpu_double(ref, monkeypatch, local_reparam):
    _cpu_double(monkeypatch)
    _drive_classification('cpu', local_reparam)


def test_reference_bandit_on_dropin_cpu_double(ref, monkeypatch):
    _cpu_double(monkeypatch)
    _drive_bandit('cpu')


@pytest.mark.parametrize('local_reparam', [False, True])
def test_state_dict_round_trip_with_the_reference_cpu_double(ref, monkeypatch, local_reparam):
    _cpu_double(monkeypatch)
    _state_dict_round_trip('cpu', local_reparam)


@pytest.mark.gpu
@pytest.mark.parametrize('local_reparam', [False, True])
def test_state_dict_round_trip_with_the_reference_gpu(ref, local_reparam):
    _state_dict_round_trip('cuda', local_reparam)


def test_reference_mc_dropout_baselines_on_dropin(ref, monkeypatch):
    _cpu_double(monkeypatch)
    _drive_mc_dropout('cpu')


@pytest.mark.gpu
@pytest.mark.parametrize('local_reparam', [False, True])
def test_reference_regression_task_on_dropin_gpu(ref, local_reparam):
    _drive_regression('cuda', local_reparam)


@pytest.mark.gpu
@pytest.mark.parametrize('local_reparam', [False, True])
def test_reference_classification_task_on_dropin_gpu(ref, local_reparam):
    _drive_classification('cuda', local_reparam)


@pytest.mark.gpu
def test_reference_bandit_on_dropin_gpu(ref):
    _drive_bandit('cuda')
