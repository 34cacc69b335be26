"""FusedAdam: torch.optim.Adam's update (same rule, same state names) for all parameter tensors in ONE
kernel launch (csrc/bbb_adam.cu).  Opt-in: the reference's callers keep constructing
torch.optim.Adam; GraphedTrainStep and bench.py use this one because the stock foreach implementation costs
more than the whole fused forward+backward at the MNIST-shape config."""
import ctypes as C

import torch

from . import _lib as L


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self.step_dev = None        # optional device uint32 counter added to the host step (CUDA-graph replay)
        self.lr_scale_dev = None    # optional device fp32 scalar multiplying lr (schedulers under a captured graph)
        self._t = 0

    def use_device_step(self, counter):
        self.step_dev = counter

    def _ensure_state(self, p):
        st = self.state[p]
        if not st:
            st['step'] = torch.tensor(0.0)
            st['exp_avg'] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st['exp_avg_sq'] = torch.zeros_like(p, memory_format=torch.preserve_format)
        return st

    def fuse_descriptor(self, layer_params):
        """bbb_adam_fuse for one layer's (w_mu, w_rho, b_mu, b_rho): lets the fused backward kernel apply THIS
        optimiser's next step in its gradient epilogue (bbb_linear_bwd_adam).  The step() that follows finds no
        gradients, launches nothing and only advances the step count."""
        group = next(g for g in self.param_groups if any(q is layer_params[0] for q in g['params']))
        d = L.AdamFuse()
        for k, p in enumerate(layer_params):
            if not (p.is_contiguous() and p.dtype == torch.float32):
                raise RuntimeError('FusedAdam needs contiguous fp32 parameters')
            st = self._ensure_state(p)
            d.exp_avg[k], d.exp_avg_sq[k] = st['exp_avg'].data_ptr(), st['exp_avg_sq'].data_ptr()
        b1, b2 = group['betas']
        d.lr, d.beta1, d.beta2, d.eps = float(group['lr']), float(b1), float(b2), float(group['eps'])
        d.step = self._t + 1
        d.step_dev = L.ptr(self.step_dev)
        d.lr_scale_dev = L.ptr(self.lr_scale_dev)
        return d

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        self._t += 1
        for group in self.param_groups:
            ps = [p for p in group['params'] if p.grad is not None]
            for i in range(0, len(ps), 32):
                self._launch(group, ps[i:i + 32])
        return loss

    def _launch(self, group, ps):
        n = len(ps)
        if n == 0:
            return
        tabs = [(C.c_void_p * n)() for _ in range(4)]
        sizes = (C.c_int64 * n)()
        for i, p in enumerate(ps):
            st = self._ensure_state(p)
            L.require_cuda(p, p.grad)
            if not (p.is_contiguous() and p.grad.is_contiguous() and p.dtype == torch.float32):
                raise RuntimeError('FusedAdam needs contiguous fp32 parameters and gradients')
            tabs[0][i], tabs[1][i] = p.data_ptr(), p.grad.data_ptr()
            tabs[2][i], tabs[3][i] = st['exp_avg'].data_ptr(), st['exp_avg_sq'].data_ptr()
            sizes[i] = p.numel()
        b1, b2 = group['betas']
        L.check(L.lib().bbb_adam_step(n, tabs[0], tabs[1], tabs[2], tabs[3], sizes, float(group['lr']), float(b1),
                                      float(b2), float(group['eps']), self._t, L.ptr(self.step_dev),
                                      L.ptr(self.lr_scale_dev), L.stream()), 'bbb_adam_step')
