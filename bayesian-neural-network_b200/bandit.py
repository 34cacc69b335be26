"""Bandit replay path (reference reinforcement_learning/base_bandit.py:73-90, bandits.py:43-51).

The reference keeps (context ++ action, reward) pairs in Python lists, and on every environment step rebuilds
`torch.Tensor([list of np arrays])`, copies it to the device and runs up to buffer_size / batch_size = 64 sequential
`loss_step`s, each a zero_grad + sample_elbo + backward + Adam.step issued from Python.  Here

* ReplayRing keeps the pairs in DEVICE memory (one preallocated [capacity, row] matrix + [capacity] rewards; an append is
  one small H2D copy of the new row), and gathers a minibatch pool with one index_select on the device -- the index
  permutation stays the caller's (np.random.permutation, base_bandit.py:76-83), so the visiting order is the reference's;
* GraphedBanditUpdate runs all minibatch steps of one `update` as ONE CUDA-graph replay: the per-minibatch step of
  bandits.py:43-51 with beta_i = 2^(M-i-1) / (2^M - 1) baked per position, eps from Philox keyed by a device step counter
  the graph advances, Adam through bnn_b200.FusedAdam.  Graphs are captured lazily per number of minibatches (<= 64).
"""
import numpy as np
import torch

from . import _lib as L
from . import rng as R
from .optim import FusedAdam


class ReplayRing:
    """Device-resident replay buffer with the reference's list semantics: rows are appended for ever, `pool(idx)`
    addresses them by their absolute position (as `self.buffer_x[i]` does), and only the last `capacity` rows are kept --
    base_bandit.py:76-83 never indexes further back than buffer_size."""

    def __init__(self, capacity, row_dim, device):
        self.capacity, self.row_dim, self.device = int(capacity), int(row_dim), torch.device(device)
        self.x = torch.zeros(self.capacity, self.row_dim, dtype=torch.float32, device=self.device)
        self.y = torch.zeros(self.capacity, dtype=torch.float32, device=self.device)
        self.count = 0
        pin = self.device.type == 'cuda'
        self._row = torch.zeros(self.row_dim + 1, dtype=torch.float32, pin_memory=pin)

    def __len__(self):
        return self.count

    def append(self, row, reward):
        """buffer_x.append(np.concatenate((context, action))); buffer_y.append(agent_reward)  (base_bandit.py:66-67)"""
        self._row[:self.row_dim] = torch.as_tensor(np.asarray(row, dtype=np.float32))
        self._row[self.row_dim] = float(reward)
        slot = self.count % self.capacity
        self.x[slot].copy_(self._row[:self.row_dim], non_blocking=True)
        self.y[slot:slot + 1].copy_(self._row[self.row_dim:], non_blocking=True)
        if self.device.type == 'cuda':
            torch.cuda.current_stream().synchronize()     # the pinned staging row is reused by the next append
        self.count += 1

    def pool(self, idx_pool):
        """context_pool, value_pool = the rows at absolute positions idx_pool (base_bandit.py:85-86), gathered on the
        device.  Positions must lie within the last `capacity` appended rows."""
        idx = np.asarray(idx_pool, dtype=np.int64)
        if idx.size and (idx.min() < max(0, self.count - self.capacity) or idx.max() >= self.count):
            raise IndexError('ReplayRing.pool: position outside the retained window')
        slots = torch.as_tensor(idx % self.capacity, device=self.device)
        return self.x.index_select(0, slots), self.y.index_select(0, slots)


def reference_idx_pool(l, batch_size, buffer_size, rng=np.random):
    """The reference's choice of replay positions for a buffer of length l (base_bandit.py:75-83), verbatim semantics."""
    if l <= batch_size:
        idx_pool = int(batch_size // l + 1) * list(range(l))
        return rng.permutation(idx_pool[-batch_size:])
    if l < buffer_size:
        n = int(l // batch_size) * batch_size
        return rng.permutation(list(range(l))[-n:])
    return rng.permutation(list(range(l))[-buffer_size:])


class GraphedBanditUpdate:
    """All minibatch steps of Bandit.update (base_bandit.py:88-89 -> BNN_Bandit.loss_step, bandits.py:43-51) as one
    CUDA-graph replay.  `net` is the drop-in BayesianNetwork, `optimizer` a bnn_b200.FusedAdam over its parameters.
    capture=False (or a CPU device) runs the same steps eagerly: the host logic is then testable without a GPU."""

    def __init__(self, net, optimizer, batch_size, num_batches, n_samples, buffer_size, row_dim, capture=True):
        self.net, self.opt = net, optimizer
        self.B, self.M, self.S = int(batch_size), int(num_batches), int(n_samples)
        self.max_steps = int(buffer_size) // self.B
        dev = next(net.parameters()).device
        self.capture = bool(capture) and dev.type == 'cuda'
        self.pool_x = torch.zeros(self.max_steps * self.B, row_dim, dtype=torch.float32, device=dev)
        self.pool_y = torch.zeros(self.max_steps * self.B, dtype=torch.float32, device=dev)
        self.counter = torch.zeros(1, dtype=torch.int32, device=dev)
        self._one = torch.ones(1, dtype=torch.float32, device=dev)
        self.graphs, self.loss_info = {}, None
        if self.capture and hasattr(optimizer, 'use_device_step'):
            optimizer.use_device_step(self.counter)

    def beta(self, batch_id):
        return 2 ** (self.M - (batch_id + 1)) / (2 ** self.M - 1)         # bandits.py:44

    def _steps(self, n):
        info = None
        for i in range(n):
            x, y = self.pool_x[i * self.B:(i + 1) * self.B], self.pool_y[i * self.B:(i + 1) * self.B]
            self.net.train()
            self.net.zero_grad(set_to_none=True)
            info = self.net.sample_elbo(x, y, self.beta(i), self.S)
            info[0].backward(self._one)
            self.opt.step()
        if self.capture:      # one bump for the whole replay: step i of a replay reads counter + (host position i)
            L.check(L.lib().bbb_counter_add(self.counter.data_ptr(), n, L.stream()), 'bbb_counter_add')
        return info

    def _graph(self, n):
        if n in self.graphs:
            return self.graphs[n]
        opt = self.opt
        params = [p for g in opt.param_groups for p in g['params']]
        p_snap = [p.detach().clone() for p in params]
        s_snap = {id(t): t.clone() for st in opt.state.values() for t in st.values() if torch.is_tensor(t)}
        # Step bookkeeping under replay: the DEVICE counter holds the number of minibatch steps taken so far; a graph bakes
        # the host positions 0..n-1 (Philox step) / 1..n (Adam step) of its own steps and adds the counter on the device,
        # so graphs of different lengths can be replayed in any order without reusing an eps stream or an Adam step.
        c_base = int(self.counter) + getattr(opt, '_t', 0)
        seed = R._st().seed
        R.use_device_step(self.counter)
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self._steps(min(n, 2))                      # warm-up: workspaces and kernels exist before the capture
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            self.counter.fill_(c_base)
            if hasattr(opt, '_t'):
                opt._t = 0
            R.manual_seed(seed, 0)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                info = self._steps(n)
            if hasattr(opt, '_t'):
                opt._t = 0
        finally:
            R.use_device_step(None)
        with torch.no_grad():                               # warm-up and capture must not train the model
            for p, snap in zip(params, p_snap):
                p.copy_(snap)
            for st in opt.state.values():
                for t in st.values():
                    if torch.is_tensor(t):      # (state created by the warm-up itself starts from zero, as a fresh one does)
                        t.copy_(s_snap[id(t)]) if id(t) in s_snap else t.zero_()
        self.counter.fill_(c_base)
        torch.cuda.synchronize()
        self.graphs[n] = (g, info)
        return self.graphs[n]

    def __call__(self, context_pool, value_pool):
        """for i in range(0, len(idx_pool), batch_size): loss_step(context_pool[i:i+B], value_pool[i:i+B], i // B)"""
        n_rows = context_pool.shape[0]
        n = (n_rows + self.B - 1) // self.B
        if n_rows != n * self.B or n > self.max_steps or n == 0:
            raise ValueError('GraphedBanditUpdate: the pool must hold 1..buffer_size/batch_size whole minibatches '
                             '(base_bandit.py:75-83 only produces such pools)')
        self.pool_x[:n_rows].copy_(context_pool, non_blocking=True)
        self.pool_y[:n_rows].copy_(value_pool, non_blocking=True)
        if not self.capture:
            self.loss_info = self._steps(n)
            return self.loss_info
        g, info = self._graph(n)
        g.replay()                                          # (the device counter carries the step count across replays)
        self.loss_info = info
        return info


def make_bandit_update(net, lr, batch_size, num_batches, n_samples, buffer_size, capture=True):
    """Convenience: FusedAdam + GraphedBanditUpdate for a BNN_Bandit-shaped net (bandits.py:24-37)."""
    opt = FusedAdam(net.parameters(), lr=lr)
    row_dim = net.input_shape
    return opt, GraphedBanditUpdate(net, opt, batch_size, num_batches, n_samples, buffer_size, row_dim, capture)
