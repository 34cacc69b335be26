"""the bandit replay path (bnn_b200.ReplayRing, GraphedBanditUpdate) against the reference's list-based
replay buffer and minibatch loop (reinforcement_learning/base_bandit.py:73-90, bandits.py:43-51)."""
import numpy as np
import pytest
import torch

import bnn_b200
from tests import fake_bbb, parity_cases as PC
from tests.golden_util import Case


@pytest.fixture
def fake(monkeypatch):
    return fake_bbb.install(monkeypatch)


def test_replay_ring_has_the_reference_list_semantics():
    rs = np.random.RandomState(0)
    ring = bnn_b200.ReplayRing(capacity=16, row_dim=5, device='cpu')
    buf_x, buf_y = [], []
    for t in range(41):
        row, rew = rs.rand(5).astype(np.float32), float(rs.randint(-35, 6))
        ring.append(row, rew)
        buf_x.append(row); buf_y.append(rew)              # base_bandit.py:66-67
        idx = bnn_b200.reference_idx_pool(len(buf_x), 4, 16, rs)
        cx, cy = ring.pool(idx)
        want_x = torch.Tensor(np.array([buf_x[i] for i in idx]))           # base_bandit.py:85-86
        want_y = torch.Tensor([buf_y[i] for i in idx])
        assert torch.equal(cx, want_x) and torch.equal(cy, want_y)
        assert len(idx) % 4 == 0 and len(idx) <= 16
    with pytest.raises(IndexError):
        ring.pool([0])                                     # evicted long ago


def test_idx_pool_follows_base_bandit():
    """l <= B: the last B entries of the repeated index list; B < l < buffer: the newest floor(l/B) B rows; else the
    newest buffer_size rows -- always a permutation (base_bandit.py:75-83)."""
    for l, want in ((1, [0] * 4), (3, None), (4, [0, 1, 2, 3]), (7, [3, 4, 5, 6]), (9, list(range(1, 9))),
                    (16, list(range(16))), (30, list(range(14, 30)))):
        idx = bnn_b200.reference_idx_pool(l, 4, 16, np.random.RandomState(1))
        if want is not None:
            assert sorted(idx.tolist()) == sorted(want), (l, idx)
        assert len(idx) % 4 == 0


def _reference_update_loop(net, opt, pool_x, pool_y, B, M, S):
    """Bandit.update's minibatch loop with BNN_Bandit.loss_step (base_bandit.py:88-89, bandits.py:43-51)."""
    info = None
    for i in range(0, pool_x.shape[0], B):
        beta = 2 ** (M - (i // B + 1)) / (2 ** M - 1)
        net.train()
        net.zero_grad()
        info = net.sample_elbo(pool_x[i:i + B], pool_y[i:i + B], beta, S)
        info[0].backward()
        opt.step()
    return info


def test_bandit_update_equals_the_reference_minibatch_loop(fake):
    c = Case('small_bandit_bcast')                        # 9-10-10-1, flat targets: the [B] vs [B,1] broadcast quirk
    rs = np.random.RandomState(3)
    pool_x = torch.tensor(rs.rand(24, c.dims[0]).astype(np.float32))
    pool_y = torch.tensor(rs.randint(-35, 6, size=24).astype(np.float32))
    ref_net, net = PC.build_net(c, 'cpu'), PC.build_net(c, 'cpu')
    ref_opt = torch.optim.Adam(ref_net.parameters(), lr=1e-3)
    opt, upd = bnn_b200.make_bandit_update(net, 1e-3, batch_size=8, num_batches=4, n_samples=2, buffer_size=32,
                                           capture=False)
    with bnn_b200.eps_mode('reference'):
        torch.manual_seed(11)
        want = _reference_update_loop(ref_net, ref_opt, pool_x, pool_y, 8, 4, 2)
        torch.manual_seed(11)
        got = upd(pool_x, pool_y)
    assert torch.allclose(got[0], want[0], rtol=1e-5)
    for p, q in zip(ref_net.parameters(), net.parameters()):
        assert torch.allclose(p, q, rtol=1e-5, atol=1e-7), float((p - q).abs().max())
    with pytest.raises(ValueError):
        upd(pool_x[:5], pool_y[:5])                       # not whole minibatches


@pytest.mark.gpu
def test_graphed_bandit_update_equals_eager_on_gpu():
    """One graph replay per Bandit.update == the same minibatch steps issued one by one (same Philox coordinates)."""
    c = Case('cfg4_bandit')
    dev = 'cuda'
    rs = np.random.RandomState(5)
    ring = bnn_b200.ReplayRing(256, c.dims[0], dev)
    for _ in range(200):
        ring.append((rs.rand(c.dims[0]) < 0.2).astype(np.float32), float(rs.choice([-35, 0, 5])))
    res = []
    for capture in (False, True):
        net = PC.build_net(c, dev)
        opt, upd = bnn_b200.make_bandit_update(net, 1e-3, batch_size=64, num_batches=4, n_samples=2, buffer_size=256,
                                               capture=capture)
        bnn_b200.manual_seed(9, 0)
        prs = np.random.RandomState(2)
        for l in (200, 130, 200):                         # 3, 2 and 3 minibatches: two graphs, replayed in turn
            idx = bnn_b200.reference_idx_pool(l, 64, 256, prs)
            info = upd(*ring.pool(idx))
        torch.cuda.synchronize()
        res.append(([p.detach().clone() for p in net.parameters()], float(info[0])))
    for a, b in zip(res[0][0], res[1][0]):
        assert torch.isfinite(b).all()
    # eager consumes host Philox steps 0,1,2,... ; the graphs read device counter + baked position: same sequence
    for a, b in zip(res[0][0], res[1][0]):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-7), float((a - b).abs().max())
