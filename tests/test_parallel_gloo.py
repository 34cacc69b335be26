"""Multi-process test of the MC-sample sharding + gradient all-reduce on CPU:
world_size 2, gloo backend, C ABI replaced by the test double.  Two ranks that each take half of the
global MC samples must reproduce the single-process gradients and ELBO scalars."""
import os
import socket
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import bnn_b200
from bnn_b200 import parallel
from oracle import bbb_oracle as O
from tests import fake_bbb, parity_cases as PC
from tests.golden_util import Case

CASE, S_TOTAL = 'small_cls_mix', 4


class _Patch:
    def setattr(self, obj, name, val):
        setattr(obj, name, val)


def _run(case, eps, count):
    net = PC.build_net(case, 'cpu')
    net.train()
    bnn_b200.rng.set_injected_eps([t for per in eps for pair in per for t in pair])
    with bnn_b200.eps_mode('injected'):
        info = net.sample_elbo(case.x, case.y, case.beta, count, sigma=case.sigma)
    info[0].backward()
    return net, info


def _all_eps(case):
    torch.manual_seed(123)
    return O.draw_eps(case.dims, S_TOTAL)


def _worker(rank, world, port, out_path):
    os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        fake_bbb.install(_Patch())
        case = Case(CASE)
        first, count = parallel.shard_samples(S_TOTAL, rank, world)
        if os.environ.get('BBB_TEST_OVERLAP') == '1':            # per-layer all-reduce issued from the backward itself
            with parallel.OverlappedAllReduce(world) as ar:
                net, info = _run(case, _all_eps(case)[first:first + count], count)
            assert ar.reduced == len(case.dims) - 1
            ar.join(net)
        else:
            net, info = _run(case, _all_eps(case)[first:first + count], count)
            grads = [p.grad for p in net.parameters()]
            assert parallel._flat_view_of(grads) is not None        # zero-copy bucket
            parallel.allreduce_gradients(net, world)
        scal = parallel.allreduce_scalars(info, world)
        if rank == 0:
            torch.save(dict(grads=PC.net_grads(net), scal=scal), out_path)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('overlap', ['0', '1'])
def test_two_rank_sample_sharding_matches_single_process(monkeypatch, overlap):
    """overlap=0: one all-reduce over the flat bucket after the backward; overlap=1: OverlappedAllReduce, the
    per-layer collectives issued from inside the backward.  Both must reproduce the one-process gradients."""
    monkeypatch.setenv('BBB_TEST_OVERLAP', overlap)
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    out_path = os.path.join(tempfile.mkdtemp(), 'rank0.pt')
    mp.spawn(_worker, args=(2, port, out_path), nprocs=2, join=True)
    got = torch.load(out_path, weights_only=False)

    fake_bbb.install(monkeypatch)
    case = Case(CASE)
    net, info = _run(case, _all_eps(case), S_TOTAL)
    want = PC.net_grads(net)
    for a, b in zip(got['grads'], want):
        for ga, gb in zip(a, b):
            assert np.abs(ga - gb).max() <= 2e-6 * np.abs(gb).max()
    np.testing.assert_allclose(got['scal'].numpy(), [float(v.detach()) for v in info], rtol=2e-6)


def _row_worker(rank, world, port, out_path):
    os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        fake_bbb.install(_Patch())
        case = Case(CASE)
        plan = parallel.ShardPlan(rank, world, S_TOTAL, case.x.shape[0], row_shards=world)   # every rank: all samples
        assert (plan.sample_first, plan.sample_count) == (0, S_TOTAL)
        net = PC.build_net(case, 'cpu')
        net.train()
        eps = _all_eps(case)
        bnn_b200.rng.set_injected_eps([t for per in eps for pair in per for t in pair])
        with bnn_b200.eps_mode('injected'):
            info = parallel.sharded_elbo(net, *plan.rows(case.x, case.y), case.beta, plan, sigma=case.sigma)
        info[0].backward()
        parallel.allreduce_gradients(net, world)
        loss = parallel.allreduce_scalars(info[:1], world)
        if rank == 0:
            torch.save(dict(grads=PC.net_grads(net), loss=loss), out_path)
    finally:
        dist.destroy_process_group()


def test_two_rank_row_sharding_matches_single_process(monkeypatch):
    """The minibatch-row axis: both ranks run ALL samples (same eps) on half of the rows with beta / 2 and the loss
    doubled; the rank-mean gradients and loss equal the one-process full-batch step."""
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    out_path = os.path.join(tempfile.mkdtemp(), 'rank0.pt')
    mp.spawn(_row_worker, args=(2, port, out_path), nprocs=2, join=True)
    got = torch.load(out_path, weights_only=False)
    fake_bbb.install(monkeypatch)
    case = Case(CASE)
    net, info = _run(case, _all_eps(case), S_TOTAL)
    for a, b in zip(got['grads'], PC.net_grads(net)):
        for ga, gb in zip(a, b):
            assert np.abs(ga - gb).max() <= 2e-6 * np.abs(gb).max()
    np.testing.assert_allclose(got['loss'].numpy(), [float(info[0].detach())], rtol=2e-6)


def test_shard_plan_grid():
    # more ranks than samples: the surplus factor goes to the rows; every (sample, row) pair is covered exactly once
    for world, S, B in ((8, 2, 128), (8, 64, 4096), (4, 1, 64), (2, 2, 128), (8, 3, 100)):
        plans = [parallel.ShardPlan(r, world, S, B) for r in range(world)]
        assert all(p.row_shards * p.sample_blocks == world for p in plans)
        assert plans[0].row_shards == (1 if S >= world else plans[0].row_shards) and (S >= world or plans[0].row_shards > 1)
        seen = {}
        for p in plans:
            for s_ in range(p.sample_first, p.sample_first + p.sample_count):
                for r_ in range(p.row_lo, p.row_hi):
                    seen[(s_, r_)] = seen.get((s_, r_), 0) + 1
        assert len(seen) == S * B and set(seen.values()) == {1}


def test_shard_samples_partitions_exactly():
    for total in (1, 2, 5, 64):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_samples(total, r, world) for r in range(world)]
            covered = [i for first, n in spans for i in range(first, first + n)]
            assert covered == list(range(total))


def test_backward_hands_out_one_flat_bucket(monkeypatch):
    fake_bbb.install(monkeypatch)
    for name in ('small_cls_mix', 'small_lr_cls'):
        case = Case(name)
        net = PC.build_net(case, 'cpu')
        net.train()
        with bnn_b200.eps_mode('reference'):
            fn = net.sample_elbo_lr if case.lr else net.sample_elbo
            fn(case.x, case.y, case.beta, case.S)[0].backward()
        flat = parallel._flat_view_of([p.grad for p in net.parameters()])
        assert flat is not None and flat.numel() == sum(p.numel() for p in net.parameters())
