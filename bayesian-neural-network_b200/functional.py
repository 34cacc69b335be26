"""torch.autograd glue between the nn.Module mirror of the reference API and the C-ABI kernels.

Every function here launches libbbb.so kernels on the current CUDA stream through ctypes
(_lib.py); PyTorch only owns the buffers and the autograd tape.  Nothing here computes on
the CPU and nothing imports oracle/.

Layer level  : bayes_linear / lr_linear          (one launch fwd, two bwd; S = 1)
Network level: mlp_forward / mlp_forward_lr      (all S samples per launch, ReLU fused into the
                                                  consumer's operand load, differentiable outputs)
               fused_elbo / fused_elbo_lr        (the above + likelihood + ELBO assembly, loss only
                                                  differentiable: what sample_elbo[_lr] call)
"""
import ctypes as C

import torch

from . import _lib as L
from . import rng as R


def make_prior(prior_init, mixture_prior):
    """prior_init -> bbb_prior (networks.py:61-68; mixture sigmas are exp(log sigma))."""
    import math
    if mixture_prior:
        assert len(prior_init) == 3, "Scale Mixture Prior requires three values in prior initialisation"
        return L.Prior(L.PRIOR_MIXTURE, float(prior_init[0]), math.exp(prior_init[1]), math.exp(prior_init[2]))
    assert len(prior_init) == 1, "Gaussian Prior requires one value in prior initialisation"
    return L.Prior(L.PRIOR_GAUSSIAN, 0.0, float(prior_init[0]), 0.0)


def _rng(layer, seed, step, sample_base, step_dev):
    return L.Rng(seed, step, sample_base, layer, L.ptr(step_dev))


def _f32c(t):
    """fp32, contiguous, detached.  The common case (an fp32 contiguous tensor seen inside autograd.Function.forward /
    backward, where nothing is recorded) costs no torch op at all: the eager call path is host-bound, and three ops per
    parameter tensor per pass were a third of it."""
    if t.dtype is torch.float32 and t.is_contiguous():
        return t if not t.requires_grad or not torch.is_grad_enabled() else t.detach()
    return t.detach().to(torch.float32).contiguous()


# Called as hook(layer_index, flat_gradient_slice) right after a layer's backward kernel has been launched (from the
# thread autograd runs the backward on).  parallel.OverlappedAllReduce installs it so that a layer's mu/rho gradients
# travel over NVLink while the layers below are still being differentiated.
grad_ready_hook = None


# Persistent flat gradient buckets, keyed by the address of the flat parameter buffer they belong to
# (optim.PeerShardedAdam registers one that the other ranks can read over NVLink): the network-level backward of a
# network whose parameters live in that buffer writes its gradients there instead of into a fresh allocation.
# Values are (flat bucket, owner): owner() (a weakref, or None) is the optimiser whose parameters the bucket serves.
grad_buckets = {}


def _bucket_for(params):
    """The registered bucket of this network, or None.  The bucket is handed to autograd as the gradient tensors
    themselves, and autograd installs them as p.grad WITHOUT a copy; if a p.grad from an earlier backward is still
    alive (zero_grad(set_to_none=False), gradient accumulation) it may BE the bucket, and the next backward would
    overwrite it before AccumulateGrad adds the new gradient to it (a silent 2x).  So the bucket is only used when
    every parameter of its owner has p.grad None; otherwise the gradients go to a fresh allocation and the owner's
    step() copies them in."""
    entry = grad_buckets.get(params[0][0].data_ptr())
    if entry is None:
        return None
    bucket, owner = entry
    opt = owner() if owner is not None else None
    if opt is not None and any(q.grad is not None for g in opt.param_groups for q in g['params']):
        return None
    return bucket


def _alloc_grads(params):
    """Gradient buffers for every layer as consecutive views of ONE flat tensor, in parameter order, so a
    multi-GPU step can all-reduce them with a single collective and no copies (parallel.py)."""
    total = sum(t.numel() for p in params for t in p)
    bucket = _bucket_for(params)
    if bucket is not None and bucket.numel() >= total and bucket.device == params[0][0].device:
        flat = bucket[:total]
    else:
        flat = torch.empty(total, dtype=torch.float32, device=params[0][0].device)
    out, off = [], 0
    for p in params:
        g, start = [], off
        for t in p:
            g.append(flat[off:off + t.numel()].view(t.shape))
            off += t.numel()
        out.append(_LayerGrads(g, flat[start:off]))
    return out


class _LayerGrads(tuple):
    """(grad_w_mu, grad_w_rho, grad_b_mu, grad_b_rho) of one layer + the flat slice of the bucket they live in"""

    def __new__(cls, grads, flat):
        self = super().__new__(cls, grads)
        self.flat = flat
        return self


def _split_beta(beta):
    """beta as a python number, or as a device tensor (read on the device: CUDA-graph replay can then follow
    the reference's per-minibatch schedule).  Returns (host multiplier, device scalar or None)."""
    if isinstance(beta, torch.Tensor):
        return 1.0, beta.detach().to(torch.float32).reshape(1)
    return float(beta), None


class _EpsPlan:
    """eps for one call: injected tensors (reference mode) or Philox coordinates."""

    def __init__(self, injected=None, seed=0, step=0):
        self.injected, self.seed, self.step = injected, seed, step
        # every RNG coordinate is pinned when the forward runs: the backward executes on autograd's own thread,
        # where the thread-local settings of rng.py (sample base, device step counter) are not the caller's
        self.sample_base, self.step_dev = R.get_sample_base(), R.device_step()

    def ptrs(self, l):
        if self.injected is None:
            return None, None
        return L.ptr(self.injected[l][0]), L.ptr(self.injected[l][1])

    def rng(self, l):
        return _rng(l, self.seed, self.step, self.sample_base, self.step_dev)

    def tensors(self):
        return [] if self.injected is None else [t for pair in self.injected.values() for t in pair]


def plan_eps(shapes, S, device, sample=True):
    """shapes: per layer (shape of the weight-like eps, shape of the bias eps).  In reference mode the
    draws follow SURVEY App. A-4: for s: for layer: weight eps, bias eps."""
    if not sample:
        return _EpsPlan()
    if R.get_eps_mode() in ('reference', 'injected'):
        draw = torch.randn if R.get_eps_mode() == 'reference' else R.pop_injected
        per = [([], []) for _ in shapes]
        for _ in range(S):
            for l, (sw, sb) in enumerate(shapes):
                per[l][0].append(draw(sw))
                per[l][1].append(draw(sb))
        inj = {l: (torch.stack(a).to(device), torch.stack(b).to(device)) for l, (a, b) in enumerate(per)}
        return _EpsPlan(inj)
    seed, step = R.next_step()
    return _EpsPlan(None, seed, step)


# =========================================================================================
# weight-sampling kernels: host-side launch helpers
# =========================================================================================
def _ws_fwd(x, x_stride, p, eps, l, prior, S, B, flags, y, logp, logq):
    wm, wr, bm, br = p
    out, inn = wm.shape
    ew, eb = eps.ptrs(l)
    rng = eps.rng(l)
    L.check(L.lib().bbb_linear_fwd(L.ptr(x), x_stride, L.ptr(wm), L.ptr(wr), L.ptr(bm), L.ptr(br), ew, eb,
                                   C.byref(rng), C.byref(prior), S, B, inn, out, flags, L.ptr(y),
                                   L.ptr(logp), L.ptr(logq), L.stream()), 'bbb_linear_fwd')


def _ws_bwd(dy, mask, x, x_stride, p, eps, l, prior, S, B, flags, gp, gq, gp_dev, gq_dev, g_stride, out_scale, dx,
            grads, adam=None):
    wm, wr, bm, br = p
    out, inn = wm.shape
    ew, eb = eps.ptrs(l)
    rng = eps.rng(l)
    if adam is not None:      # the optimiser's update is applied inside the kernel; no gradient is written
        L.check(L.lib().bbb_linear_bwd_adam(L.ptr(dy), L.ptr(mask), L.ptr(x), x_stride, L.ptr(wm), L.ptr(wr),
                                            L.ptr(bm), L.ptr(br), ew, eb, C.byref(rng), C.byref(prior), S, B, inn,
                                            out, flags, gp, gq, L.ptr(gp_dev), L.ptr(gq_dev), g_stride,
                                            L.ptr(out_scale), L.ptr(dx), C.byref(adam), L.stream()),
                'bbb_linear_bwd_adam')
        return
    L.check(L.lib().bbb_linear_bwd(L.ptr(dy), L.ptr(mask), L.ptr(x), x_stride, L.ptr(wm), L.ptr(wr), L.ptr(bm),
                                   L.ptr(br), ew, eb, C.byref(rng), C.byref(prior), S, B, inn, out, flags,
                                   gp, gq, L.ptr(gp_dev), L.ptr(gq_dev), g_stride, L.ptr(out_scale), L.ptr(dx),
                                   L.ptr(grads[0]), L.ptr(grads[1]), L.ptr(grads[2]), L.ptr(grads[3]),
                                   L.stream()), 'bbb_linear_bwd')


def _zeroed_views(shapes, device):
    """One zero-filled buffer carved into tensors of the given shapes (a single memset for all of them): the
    split-K tensor-core kernels add partial tiles into their output with red.add (BBB_F_OUT_ZEROED)."""
    sizes = [int(torch.Size(sh).numel()) for sh in shapes]
    padded = [(n + 63) // 64 * 64 for n in sizes]           # every view starts on a 256-byte boundary
    flat = torch.zeros(sum(padded), dtype=torch.float32, device=device)
    out, off = [], 0
    for sh, n, m in zip(shapes, sizes, padded):
        out.append(flat[off:off + n].view(sh))
        off += m
    return out


def head_eligible(p):
    """bbb_head_fwd's shape conditions (include/bbb.h): a narrow last layer whose rows are 16-byte multiples."""
    out, inn = p[0].shape
    return out <= 16 and inn % 4 == 0 and 4 <= inn <= 8192


def _ws_head(x, x_stride, p, eps, l, prior, S, B, flags, y, logp, logq, mode, target, sigma, d_out, nll, beta_h,
             beta_d, out4, done):
    """Last layer + likelihood (+ its gradient) + ELBO assembly in one launch (bbb_head_fwd)."""
    wm, wr, bm, br = p
    out, inn = wm.shape
    ew, eb = eps.ptrs(l)
    rng = eps.rng(l)
    kind = L.NLL_CE if mode == 'classification' else L.NLL_GAUSS
    L.check(L.lib().bbb_head_fwd(L.ptr(x), x_stride, L.ptr(wm), L.ptr(wr), L.ptr(bm), L.ptr(br), ew, eb,
                                 C.byref(rng), C.byref(prior), S, B, inn, out, flags, kind, L.ptr(target),
                                 float(sigma), 1.0 / S, L.ptr(y), L.ptr(d_out), L.ptr(logp), L.ptr(logq), L.ptr(nll),
                                 beta_h, L.ptr(beta_d), L.ptr(out4), L.ptr(done), L.stream()), 'bbb_head_fwd')


def _prezero(B):
    """Only the kernels for batches of at most 128 rows combine split-K partial tiles with red.add into a zero-filled
    output (BBB_F_OUT_ZEROED).  Above that every kernel either writes its output with plain stores or zeroes it
    itself, so the per-step workspace (gigabytes at batch 4096) is allocated uninitialised."""
    return B <= 128


def _workspace(shapes, device, zero):
    if zero:
        return _zeroed_views(shapes, device)
    return [torch.empty(sh, dtype=torch.float32, device=device) for sh in shapes]


_relu_out_cache = {}


def _relu_out_layers(params, B, tf32):
    """Per layer: is its output stored POST-activation (BBB_F_RELU_OUT)?  The large-batch tensor kernels do that for
    hidden layers, so that nothing downstream passes over TMA-landed tiles to apply the ReLU; the consumers of such an
    output are called without BBB_F_RELU_IN (the (x > 0) mask of BBB_F_DX_PREACT is the same on either form)."""
    if not tf32:
        return [False] * len(params)
    key = (B, tuple(tuple(p[0].shape) for p in params))
    ro = _relu_out_cache.get(key)
    if ro is None:
        ro = [l < len(params) - 1 and
              bool(L.lib().bbb_linear_fwd_relu_out_supported(B, p[0].shape[1], p[0].shape[0], L.F_TF32))
              for l, p in enumerate(params)]
        _relu_out_cache[key] = ro
    return ro


def _net_ws_forward(x2, params, prior, S, eps, sample, logprob, tf32, logp, logq, ys=None, head=None):
    """All layers, all S samples.  Returns the list of outputs ys[l] = [S,B,out_l]: pre-activations, except the hidden
    layers _relu_out_layers names (written into the zero-filled `ys` when the caller supplies them).
    `head(inp, stride, flags, y)`, when given, runs the last layer instead of bbb_linear_fwd (the fused ELBO tail)."""
    B = x2.shape[0]
    ro = _relu_out_layers(params, B, tf32)
    base = ((L.F_SAMPLE if sample else 0) | (L.F_LOGPROB if logprob else 0) | (L.F_TF32 if tf32 else 0) |
            (L.F_OUT_ZEROED if _prezero(B) else 0))
    if ys is None:
        ys = _workspace([(S, B, p[0].shape[0]) for p in params], x2.device, _prezero(B))
    inp, stride = x2, 0
    for l, p in enumerate(params):
        out, inn = p[0].shape
        flags = base | (L.F_RELU_IN if l > 0 and not ro[l - 1] else 0) | (L.F_RELU_OUT if ro[l] else 0)
        if head is not None and l == len(params) - 1:
            head(inp, stride, flags, ys[l])
        else:
            _ws_fwd(inp, stride, p, eps, l, prior, S, B, flags, ys[l], logp, logq)
        inp, stride = ys[l], B * out
    return ys


def _net_ws_backward(x2, ys, d_out, params, prior, S, eps, sample, tf32, gp, gq, gp_dev, gq_dev, g_stride,
                     out_scale, need_dx0, fused_opt=None, live_params=None, dxs=None):
    """Backward of _net_ws_forward.  Returns (dx0 or None, [grads per layer]).  Every layer above the first
    hands down the gradient w.r.t. the PRE-activation output of the layer below (BBB_F_DX_PREACT: the ReLU mask
    is applied where dx is produced), so no layer needs a separate mask pass over dy."""
    B = x2.shape[0]
    ro = _relu_out_layers(params, B, tf32)
    base = (L.F_SAMPLE if sample else 0) | (L.F_TF32 if tf32 else 0) | (L.F_OUT_ZEROED if _prezero(B) else 0)
    grads = _alloc_grads(params) if fused_opt is None else [None] * len(params)
    dy, dx0 = d_out, None
    first = 0 if need_dx0 else 1
    if dxs is None:
        dxs = [None] * first + _workspace([(S, B, p[0].shape[1]) for p in params[first:]], x2.device, _prezero(B))
    for l in reversed(range(len(params))):
        p = params[l]
        out, inn = p[0].shape
        g = grads[l]
        flags = base | (L.F_DX_PREACT if l > 0 else 0) | (L.F_RELU_IN if l > 0 and not ro[l - 1] else 0)
        want_dx = l > 0 or need_dx0
        dx = dxs[l] if want_dx else None
        if not want_dx:
            flags |= L.F_NO_DX
        if l == 0 and need_dx0:
            flags |= L.F_SCALE_DX
        x_in, stride = (x2, 0) if l == 0 else (ys[l - 1], B * inn)
        adam = fused_opt.fuse_descriptor(live_params[l]) if fused_opt is not None else None
        _ws_bwd(dy, None, x_in, stride, p, eps, l, prior, S, B, flags, gp, gq, gp_dev, gq_dev, g_stride, out_scale,
                dx, g, adam)
        if grad_ready_hook is not None and fused_opt is None:
            grad_ready_hook(l, g.flat)
        dy = dx
        if l == 0:
            dx0 = dx
    return dx0, grads


# =========================================================================================
# layer level
# =========================================================================================
class _BayesLinear(torch.autograd.Function):
    """BayesianLinear.forward for one sample (networks.py:73-88) -> (y, log_prior, log_post)."""

    @staticmethod
    def forward(ctx, x, wm, wr, bm, br, prior, sample, logprob, layer_id, tf32):
        L.require_cuda(x, wm, wr, bm, br)
        out, inn = wm.shape
        x2 = _f32c(x).reshape(-1, inn)
        B = x2.shape[0]
        p = tuple(_f32c(t) for t in (wm, wr, bm, br))
        eps = plan_eps([((out, inn), (out,))], 1, x2.device, sample)
        if eps.injected is not None:            # re-key as layer `layer_id`
            eps.injected = {layer_id: eps.injected[0]}
        acc = torch.zeros(2, dtype=torch.float64, device=x2.device)
        y = torch.empty((1, B, out), dtype=torch.float32, device=x2.device)
        flags = (L.F_SAMPLE if sample else 0) | (L.F_LOGPROB if logprob else 0) | (L.F_TF32 if tf32 else 0)
        _ws_fwd(x2, 0, p, eps, layer_id, prior, 1, B, flags, y, acc[0:1], acc[1:2])
        ctx.save_for_backward(x2, *p, *eps.tensors())
        ctx.cfg = (prior, sample, logprob, layer_id, tf32, eps, x.shape, x.requires_grad)
        ctx.set_materialize_grads(False)
        lpq = acc.to(torch.float32)
        return y.view(*x.shape[:-1], out), lpq[0], lpq[1]

    @staticmethod
    def backward(ctx, dy, dlp, dlq):
        prior, sample, logprob, layer_id, tf32, eps, xshape, x_rg = ctx.cfg
        x2, wm, wr, bm, br = ctx.saved_tensors[:5]
        out, inn = wm.shape
        B = x2.shape[0]
        if dy is None:
            dy = torch.zeros((B, out), dtype=torch.float32, device=x2.device)
        dy = _f32c(dy).reshape(1, B, out)
        g = tuple(torch.empty_like(t) for t in (wm, wr, bm, br))
        dx = torch.empty((1, B, inn), dtype=torch.float32, device=x2.device) if x_rg else None
        flags = (L.F_SAMPLE if sample else 0) | (L.F_TF32 if tf32 else 0) | (0 if x_rg else L.F_NO_DX)
        gp_dev = _f32c(dlp).reshape(1) if dlp is not None else None
        gq_dev = _f32c(dlq).reshape(1) if dlq is not None else None
        _ws_bwd(dy, None, x2, 0, (wm, wr, bm, br), eps, layer_id, prior, 1, B, flags,
                1.0 if gp_dev is not None else 0.0, 1.0 if gq_dev is not None else 0.0, gp_dev, gq_dev, 0, None, dx, g)
        return (dx.view(xshape) if x_rg else None, g[0], g[1], g[2], g[3], None, None, None, None, None)


def bayes_linear(x, wm, wr, bm, br, prior, sample, logprob, layer_id=0, tf32=False):
    return _BayesLinear.apply(x, wm, wr, bm, br, prior, sample, logprob, layer_id, tf32)


# =========================================================================================
# network level, weight sampling
# =========================================================================================
class _MLPForward(torch.autograd.Function):
    """S sampled forwards of the whole MLP: (x, params) -> outputs [S,B,C], log_prior [S], log_post [S]."""

    @staticmethod
    def forward(ctx, x2, prior, S, sample, logprob, tf32, *flat):
        L.require_cuda(x2, *flat)
        x2 = _f32c(x2)
        params = [tuple(_f32c(t) for t in flat[i:i + 4]) for i in range(0, len(flat), 4)]
        eps = plan_eps([(tuple(p[0].shape), (p[0].shape[0],)) for p in params], S, x2.device, sample)
        acc = torch.zeros(2 * S, dtype=torch.float64, device=x2.device)
        ys = _net_ws_forward(x2, params, prior, S, eps, sample, logprob, tf32, acc[:S], acc[S:])
        ctx.save_for_backward(x2, *flat, *ys, *eps.tensors())
        ctx.cfg = (prior, S, sample, tf32, eps, len(params), x2.requires_grad)
        ctx.set_materialize_grads(False)
        lpq = acc.to(torch.float32)
        # the last layer's buffer is handed out; backward only needs the hidden pre-activations
        return ys[-1], lpq[:S], lpq[S:]

    @staticmethod
    def backward(ctx, d_out, dlp, dlq):
        prior, S, sample, tf32, eps, nl, x_rg = ctx.cfg
        sv = ctx.saved_tensors
        x2 = sv[0]
        params = [tuple(_f32c(t) for t in sv[1 + 4 * i:5 + 4 * i]) for i in range(nl)]
        ys = list(sv[1 + 4 * nl:1 + 5 * nl])
        B = x2.shape[0]
        if d_out is None:
            d_out = torch.zeros_like(ys[-1])
        gp_dev = _f32c(dlp) if dlp is not None else None
        gq_dev = _f32c(dlq) if dlq is not None else None
        dx0, grads = _net_ws_backward(x2, ys, _f32c(d_out), params, prior, S, eps, sample, tf32,
                                      1.0 if gp_dev is not None else 0.0, 1.0 if gq_dev is not None else 0.0,
                                      gp_dev, gq_dev, 1, None, x_rg)
        flat = [g for lg in grads for g in lg]
        return (dx0.sum(0) if x_rg else None, None, None, None, None, None, *flat)


def mlp_forward(x2, layers, prior, S, sample=True, logprob=True, tf32=False):
    flat = [t for layer in layers for t in layer]
    return _MLPForward.apply(x2, prior, S, sample, logprob, tf32, *flat)


# sample_elbo in TF32 mode goes through ONE C call per network pass (bbb_mlp_fwd / bbb_mlp_bwd) when the library covers
# the network (bbb_mlp_supported); False forces the per-layer calls (A/B measurements, tools/check_mlp.py)
use_network_level_call = True
# the head on a full grid (csrc/bbb_head2.cu) instead of the 8-CTA-cluster head: measured SLOWER on B200 (23.6 + 21.0 us
# against 16.0 + 9.4 us by ncu, cfg2), so it stays opt-in -- DESIGN.md 4.7
use_full_grid_head = False


def _layerwise_forward_call(x2, target, params, eps, prior, S, B, sigma, mode, need_grad, tf32, beta_h, beta_d, out4):
    """sample_elbo's forward as one launch per layer (+ the fused head): every shape, both modes.  Returns
    (ys: per-layer PRE-activation outputs, dxs: the backward's zero-filled dx workspace, d_out)."""
    dev = x2.device
    # ONE zero-filled workspace (one memset per step): fp64 accumulators + the head's completion counter |
    # activations of every layer | dx of every layer above the first (the backward's split-K kernels add into them)
    shapes = [(2 * (2 * S + 1) + 2,)] + [(S, B, p[0].shape[0]) for p in params]
    if need_grad:
        shapes += [(S, B, p[0].shape[1]) for p in params[1:]]
    if _prezero(B):
        ws = _zeroed_views(shapes, dev)
    else:                      # large batch: only the accumulators are zeroed
        ws = _zeroed_views(shapes[:1], dev) + _workspace(shapes[1:], dev, False)
    acc = ws[0][:2 * (2 * S + 1)].view(torch.float64)
    done = ws[0][2 * (2 * S + 1):]
    ys, dxs = ws[1:1 + len(params)], ([None] + ws[1 + len(params):] if need_grad else None)
    logp, logq, nll = acc[:S], acc[S:2 * S], acc[2 * S:]
    Cc = params[-1][0].shape[0]
    d_out = torch.empty((S, B, Cc), dtype=torch.float32, device=dev) if need_grad else None
    if head_eligible(params[-1]):
        # the last layer, the likelihood with its gradient and the assembly of the four scalars: one launch
        tgt = target if mode == 'classification' else _f32c(target)
        nl = len(params) - 1

        def head(inp, stride, flags, y):
            _ws_head(inp, stride, params[nl], eps, nl, prior, S, B, flags, y, logp, logq, mode, tgt, sigma, d_out,
                     nll, beta_h, beta_d, out4, done)
        ys = _net_ws_forward(x2, params, prior, S, eps, True, True, tf32, logp, logq, ys, head)
    else:
        ys = _net_ws_forward(x2, params, prior, S, eps, True, True, tf32, logp, logq, ys)
        out = ys[-1]
        if mode == 'classification':
            L.check(L.lib().bbb_nll_ce(L.ptr(out), L.ptr(target), S, B, Cc, 1.0 / S, L.ptr(nll), L.ptr(d_out),
                                       L.stream()), 'bbb_nll_ce')
        else:
            tgt = _f32c(target)
            L.check(L.lib().bbb_nll_gauss(L.ptr(out), L.ptr(tgt), float(sigma), S, B, Cc, 1.0 / S, L.ptr(nll),
                                          L.ptr(d_out), L.stream()), 'bbb_nll_gauss')
        L.check(L.lib().bbb_elbo_finalize(L.ptr(logp), L.ptr(logq), None, L.ptr(nll), S, beta_h, L.ptr(beta_d),
                                          L.ptr(out4), L.stream()), 'bbb_elbo_finalize')
    return ys, dxs, d_out


def _mlp_layer_table(params, eps, ys, dzs, grads, head_scratch=None):
    tab = (L.MlpLayer * len(params))()
    for l, p in enumerate(params):
        t = tab[l]
        t.w_mu, t.w_rho, t.b_mu, t.b_rho = (q.data_ptr() for q in p)
        t.eps_w, t.eps_b = eps.ptrs(l)
        t.out, t.inn = p[0].shape
        t.y, t.dz = L.ptr(ys[l]), L.ptr(dzs[l])
        if grads is not None and grads[l] is not None:
            t.g_w_mu, t.g_w_rho, t.g_b_mu, t.g_b_rho = (g.data_ptr() for g in grads[l])
    tab[len(params) - 1].w_sample = L.ptr(head_scratch)
    return tab


def _mlp_forward_call(x2, target, params, eps, prior, S, B, sigma, mode, need_grad, beta_h, beta_d, out4):
    """sample_elbo's forward as ONE C call (bbb_mlp_fwd): TMA-fed tcgen05 kernels for the hidden layers + the fused
    head.  Returns (ys, dxs, d_out): ys[l] the pre-activation output of layer l (the hidden ones live in the zero-filled
    workspace: their split-K partial tiles are reduce-added into it), dxs[l] the zero-filled gradient buffer w.r.t.
    layer l's pre-activation input."""
    dev = x2.device
    nl = len(params)
    hidden = [p[0].shape[0] for p in params[:-1]]
    # zero-filled (one memset): fp64 accumulators + the head's done counter | pre-activations of every hidden layer |
    # gradient w.r.t. every hidden layer's pre-activation output (the layer above adds its partial dx tiles into it)
    shapes = [(2 * (2 * S + 1) + 2,)] + [(S, B, h) for h in hidden]
    if need_grad:
        shapes += [(S, B, h) for h in hidden]
    ws = _zeroed_views(shapes, dev)
    acc = ws[0][:2 * (2 * S + 1)].view(torch.float64)
    done = ws[0][2 * (2 * S + 1):]
    Cc = params[-1][0].shape[0]
    ys = ws[1:1 + len(hidden)] + [torch.empty((S, B, Cc), dtype=torch.float32, device=dev)]
    dzs = (ws[1 + len(hidden):] if need_grad else [None] * len(hidden))
    logp, logq, nll = acc[:S], acc[S:2 * S], acc[2 * S:]
    d_out = torch.empty((S, B, Cc), dtype=torch.float32, device=dev) if need_grad else None
    # the head's sampled weights (+ biases): written once by its forward, reused by its backward
    hs = (torch.empty(S * Cc * params[-1][0].shape[1] + 16 * S, dtype=torch.float32, device=dev)
          if use_full_grid_head else None)
    tab = _mlp_layer_table(params, eps, ys, list(dzs) + [d_out], None, hs)
    kind = L.NLL_CE if mode == 'classification' else L.NLL_GAUSS
    tgt = target if mode == 'classification' else _f32c(target)
    rng = eps.rng(0)
    L.check(L.lib().bbb_mlp_fwd(tab, nl, L.ptr(x2), S, B, C.byref(rng), C.byref(prior),
                                L.F_SAMPLE | L.F_LOGPROB | L.F_TF32, kind, L.ptr(tgt), float(sigma), 1.0 / S,
                                L.ptr(d_out), L.ptr(logp), L.ptr(logq), L.ptr(nll), beta_h, L.ptr(beta_d), L.ptr(out4),
                                L.ptr(done), L.stream()), 'bbb_mlp_fwd')
    return ys, ([None] + list(dzs) + [hs] if need_grad else None), d_out


class _FusedELBO(torch.autograd.Function):
    """sample_elbo (networks.py:192-209).  TF32 mode, batch <= 128: ONE C call per pass (bbb_mlp_fwd: a TMA-fed tcgen05
    kernel per hidden layer + the fused head; bbb_mlp_bwd: head + one fused wgrad/dgrad kernel per layer).  Otherwise one
    launch per layer (+ likelihood + assembly).  Only `loss` is differentiable."""

    @staticmethod
    def forward(ctx, x2, target, beta, S, sigma, mode, prior, tf32, fused_opt, *flat):
        L.require_cuda(x2, target, *flat)
        x2 = _f32c(x2)
        dev = x2.device
        params = [tuple(_f32c(t) for t in flat[i:i + 4]) for i in range(0, len(flat), 4)]
        eps = plan_eps([(tuple(p[0].shape), (p[0].shape[0],)) for p in params], S, dev, True)
        need_grad = any(t.requires_grad for t in flat)
        B = x2.shape[0]
        Cc = params[-1][0].shape[0]
        out4 = torch.empty(4, dtype=torch.float32, device=dev)
        beta_h, beta_d = _split_beta(beta)
        dims = [params[0][0].shape[1]] + [p[0].shape[0] for p in params]
        # (the optimiser fused into the network-level backward needs one sample group: S <= 2)
        use_mlp = (use_network_level_call and tf32 and (fused_opt is None or S <= 2) and
                   mode in ('classification', 'regression') and L.mlp_supported(dims, S, B, L.F_TF32))
        if use_mlp:
            ys, dxs, d_out = _mlp_forward_call(x2, target, params, eps, prior, S, B, sigma, mode, need_grad, beta_h,
                                               beta_d, out4)
        else:
            ys, dxs, d_out = _layerwise_forward_call(x2, target, params, eps, prior, S, B, sigma, mode, need_grad, tf32,
                                                     beta_h, beta_d, out4)
        if need_grad:
            ctx.save_for_backward(x2, d_out, *flat, *ys[:-1], *eps.tensors())
            ctx.dxs = dxs
        ctx.used_mlp = use_mlp
        ctx.cfg = (prior, S, beta_h, tf32, eps, len(params))
        ctx.beta_dev = beta_d
        ctx.set_materialize_grads(False)      # no zero-filled gradients for the three non-differentiable scalars
        ctx.fused_opt, ctx.live = fused_opt if fused_opt is not None else (None, None)
        loss, lp, lq, nl = out4[0:1], out4[1], out4[2], out4[3:4]
        ctx.mark_non_differentiable(lp, lq, nl)
        return loss, lp, lq, nl

    @staticmethod
    def backward(ctx, g_loss, *unused):
        prior, S, beta, tf32, eps, nl = ctx.cfg
        if g_loss is None:
            return (None,) * (9 + 4 * nl)
        if getattr(ctx, 'consumed', False):
            # the split-K backward kernels ADD into the dx workspace zero-filled by the forward (BBB_F_OUT_ZEROED)
            raise RuntimeError('sample_elbo: this graph was already differentiated (retain_graph=True is not '
                               'supported by the fused ELBO path: its backward workspace is single-use)')
        ctx.consumed = True
        sv = ctx.saved_tensors
        x2, d_out = sv[0], sv[1]
        params = [tuple(_f32c(t) for t in sv[2 + 4 * i:6 + 4 * i]) for i in range(nl)]
        ys = list(sv[2 + 4 * nl:2 + 4 * nl + nl - 1]) + [None]
        scale = _f32c(g_loss).reshape(1)
        bd = ctx.beta_dev
        if ctx.fused_opt is not None and ctx.used_mlp:
            # ONE C call: backward of every layer with the optimiser's update applied in the kernels' gradient write-back
            # (parameters change here, no .grad appears); only the head's gradients are materialised (scratch)
            with torch.no_grad():
                B = x2.shape[0]
                overlap = bool(getattr(ctx.fused_opt, 'overlap_backward', False))
                if overlap:     # ordinary backward kernels + per-layer Adam on a side stream: every layer writes gradients
                    sg = _alloc_grads(params)
                else:           # update in the write-back: only the head writes its gradients
                    sg = [None] * (nl - 1) + [_alloc_grads(params[-1:])[0]]
                tab = _mlp_layer_table(params, eps, ys, list(ctx.dxs[1:nl]) + [d_out], sg, ctx.dxs[nl])
                adam = (L.AdamFuse * nl)()
                for l in range(nl):
                    d = ctx.fused_opt.fuse_descriptor(ctx.live[l])
                    C.memmove(C.byref(adam[l]), C.byref(d), C.sizeof(L.AdamFuse))
                rng = eps.rng(0)
                L.check(L.lib().bbb_mlp_bwd(tab, nl, L.ptr(x2), S, B, C.byref(rng), C.byref(prior),
                                            L.F_SAMPLE | L.F_TF32 | (L.F_ADAM_OVERLAP if overlap else 0), -beta / S,
                                            beta / S, L.ptr(bd), L.ptr(bd), 0, L.ptr(scale), adam, L.stream()),
                        'bbb_mlp_bwd')
            return (None,) * (9 + 4 * nl)
        if ctx.fused_opt is not None:     # Adam rides in the backward kernels: parameters change here, no .grad appears
            with torch.no_grad():
                _net_ws_backward(x2, ys, d_out, params, prior, S, eps, True, tf32, -beta / S, beta / S, bd, bd, 0,
                                 scale, False, ctx.fused_opt, ctx.live, ctx.dxs)
            return (None,) * (9 + 4 * nl)
        if ctx.used_mlp:
            # ONE C call for the whole backward (bbb_mlp_bwd): ys are the stored pre-activations, ctx.dxs[l + 1] the
            # zero-filled gradient w.r.t. hidden layer l's pre-activation output
            grads = _alloc_grads(params)
            B = x2.shape[0]
            dzs = list(ctx.dxs[1:nl]) + [d_out]
            tab = _mlp_layer_table(params, eps, ys, dzs, grads, ctx.dxs[nl])
            rng = eps.rng(0)
            L.check(L.lib().bbb_mlp_bwd(tab, nl, L.ptr(x2), S, B, C.byref(rng), C.byref(prior), L.F_SAMPLE | L.F_TF32,
                                        -beta / S, beta / S, L.ptr(bd), L.ptr(bd), 0, L.ptr(scale), None, L.stream()),
                    'bbb_mlp_bwd')
            if grad_ready_hook is not None:
                for l in reversed(range(nl)):
                    grad_ready_hook(l, grads[l].flat)
        else:
            _, grads = _net_ws_backward(x2, ys, d_out, params, prior, S, eps, True, tf32, -beta / S, beta / S,
                                        bd, bd, 0, scale, False, dxs=ctx.dxs)
        flat = [g for lg in grads for g in lg]
        return (None,) * 9 + tuple(flat)


def fused_elbo(x2, target, beta, S, sigma, mode, prior, layers, tf32=False, fused_opt=None):
    """fused_opt: None, or a bnn_b200.FusedAdam whose next step the backward kernels apply themselves."""
    flat = [t for layer in layers for t in layer]
    fo = (fused_opt, [tuple(layer) for layer in layers]) if fused_opt is not None else None
    return _FusedELBO.apply(x2, target, beta, S, sigma, mode, prior, tf32, fo, *flat)


# =========================================================================================
# local reparameterisation
# =========================================================================================
def _lr_fwd(x, x_stride, p, eps, l, sigma_p, S, B, flags, y, delta, kl):
    wm, wr, bm, br = p
    inn, out = wm.shape
    ea, eb = eps.ptrs(l)
    rng = eps.rng(l)
    L.check(L.lib().bbb_lr_linear_fwd(L.ptr(x), x_stride, L.ptr(wm), L.ptr(wr), L.ptr(bm), L.ptr(br), ea, eb,
                                      C.byref(rng), sigma_p, S, B, inn, out, flags, L.ptr(y), L.ptr(delta),
                                      L.ptr(kl), L.stream()), 'bbb_lr_linear_fwd')


def _lr_bwd(dy, mask, x, x_stride, p, eps, l, delta, sigma_p, S, B, flags, g_kl, g_kl_dev, out_scale, dx, grads):
    wm, wr, bm, br = p
    inn, out = wm.shape
    ea, eb = eps.ptrs(l)
    rng = eps.rng(l)
    L.check(L.lib().bbb_lr_linear_bwd(L.ptr(dy), L.ptr(mask), L.ptr(x), x_stride, L.ptr(wm), L.ptr(wr), L.ptr(bm),
                                      L.ptr(br), ea, eb, C.byref(rng), L.ptr(delta), sigma_p, S, B, inn, out, flags,
                                      g_kl, L.ptr(g_kl_dev), L.ptr(out_scale), L.ptr(dx), L.ptr(grads[0]),
                                      L.ptr(grads[1]), L.ptr(grads[2]), L.ptr(grads[3]), L.stream()),
            'bbb_lr_linear_bwd')


def _net_lr_forward(x2, params, sigma_p, S, eps, sample, calc_kl, kl, tf32=False):
    B = x2.shape[0]
    base = ((L.F_SAMPLE if sample else 0) | (L.F_LOGPROB if calc_kl else 0) | (L.F_TF32 if tf32 else 0) |
            L.F_OUT_ZEROED)
    # one zero-filled buffer for every layer's y and delta: the tensor-core kernels add their split-K partial sums
    # of x mu and x^2 sigma^2 into them before the epilogue turns them into the outputs
    shapes = [(S, B, p[0].shape[1]) for p in params]
    bufs = _zeroed_views(shapes + (shapes if sample else []), x2.device)
    ys, deltas = bufs[:len(params)], (bufs[len(params):] if sample else [None] * len(params))
    inp, stride = x2, 0
    for l, p in enumerate(params):
        inn, out = p[0].shape
        _lr_fwd(inp, stride, p, eps, l, sigma_p, S, B, base | (L.F_RELU_IN if l > 0 else 0), ys[l], deltas[l], kl)
        inp, stride = ys[l], B * out
    return ys, deltas


def _net_lr_backward(x2, ys, deltas, d_out, params, sigma_p, S, eps, sample, calc_kl, g_kl, g_kl_dev, out_scale,
                     need_dx0, tf32=False):
    """Backward of _net_lr_forward.  As in the weight-sampling path every layer above the first hands down the
    gradient w.r.t. the pre-activation output of the layer below (BBB_F_DX_PREACT).  In TF32 mode the kernels
    overwrite the saved delta buffers with dV (they have no other use in the backward)."""
    B = x2.shape[0]
    base = ((L.F_SAMPLE if sample else 0) | (L.F_LOGPROB if calc_kl else 0) | (L.F_TF32 if tf32 else 0) |
            L.F_OUT_ZEROED)
    grads = _alloc_grads(params)
    dy, dx0 = d_out, None
    first = 0 if need_dx0 else 1
    dxs = [None] * first + _zeroed_views([(S, B, p[0].shape[0]) for p in params[first:]], x2.device)
    for l in reversed(range(len(params))):
        p = params[l]
        inn, out = p[0].shape
        g = grads[l]
        flags = base | ((L.F_RELU_IN | L.F_DX_PREACT) if l > 0 else 0)
        want_dx = l > 0 or need_dx0
        dx = dxs[l] if want_dx else None
        if not want_dx:
            flags |= L.F_NO_DX
        if l == 0 and need_dx0:
            flags |= L.F_SCALE_DX
        x_in, stride = (x2, 0) if l == 0 else (ys[l - 1], B * inn)
        _lr_bwd(dy, None, x_in, stride, p, eps, l, deltas[l], sigma_p, S, B, flags, g_kl, g_kl_dev, out_scale, dx, g)
        if grad_ready_hook is not None:
            grad_ready_hook(l, g.flat)
        dy = dx
        if l == 0:
            dx0 = dx
    return dx0, grads


class _LRLinear(torch.autograd.Function):
    """BayesianLinearLR.forward for one sample (networks.py:116-138) -> (y, kl)."""

    @staticmethod
    def forward(ctx, x, wm, wr, bm, br, sigma_p, sample, calc_kl, layer_id):
        L.require_cuda(x, wm, wr, bm, br)
        inn, out = wm.shape
        x2 = _f32c(x).reshape(-1, inn)
        B = x2.shape[0]
        p = tuple(_f32c(t) for t in (wm, wr, bm, br))
        eps = plan_eps([((B, out), (out,))], 1, x2.device, sample)
        if eps.injected is not None:
            eps.injected = {layer_id: eps.injected[0]}
        kl = torch.zeros(1, dtype=torch.float64, device=x2.device)
        y = torch.empty((1, B, out), dtype=torch.float32, device=x2.device)
        delta = torch.empty_like(y) if sample else None
        flags = (L.F_SAMPLE if sample else 0) | (L.F_LOGPROB if calc_kl else 0)
        _lr_fwd(x2, 0, p, eps, layer_id, sigma_p, 1, B, flags, y, delta, kl)
        saved = [x2, *p] + ([delta] if sample else []) + eps.tensors()
        ctx.save_for_backward(*saved)
        ctx.cfg = (sigma_p, sample, calc_kl, layer_id, eps, x.shape, x.requires_grad)
        ctx.set_materialize_grads(False)
        return y.view(*x.shape[:-1], out), kl.to(torch.float32)[0]

    @staticmethod
    def backward(ctx, dy, dkl):
        sigma_p, sample, calc_kl, layer_id, eps, xshape, x_rg = ctx.cfg
        sv = ctx.saved_tensors
        x2, wm, wr, bm, br = sv[:5]
        delta = sv[5] if sample else None
        inn, out = wm.shape
        B = x2.shape[0]
        if dy is None:
            dy = torch.zeros((B, out), dtype=torch.float32, device=x2.device)
        dy = _f32c(dy).reshape(1, B, out)
        g = tuple(torch.empty_like(t) for t in (wm, wr, bm, br))
        dx = torch.empty((1, B, inn), dtype=torch.float32, device=x2.device) if x_rg else None
        use_kl = calc_kl and dkl is not None
        flags = (L.F_SAMPLE if sample else 0) | (L.F_LOGPROB if use_kl else 0) | (0 if x_rg else L.F_NO_DX)
        g_kl_dev = _f32c(dkl).reshape(1) if use_kl else None
        _lr_bwd(dy, None, x2, 0, (wm, wr, bm, br), eps, layer_id, delta, sigma_p, 1, B, flags, 1.0, g_kl_dev, None,
                dx, g)
        return (dx.view(xshape) if x_rg else None, g[0], g[1], g[2], g[3], None, None, None, None)


def lr_linear(x, wm, wr, bm, br, sigma_p, sample, calc_kl, layer_id=0):
    return _LRLinear.apply(x, wm, wr, bm, br, float(sigma_p), sample, calc_kl, layer_id)


def _lr_eps_shapes(params, B):
    return [((B, p[0].shape[1]), (p[0].shape[1],)) for p in params]


class _MLPForwardLR(torch.autograd.Function):
    """S sampled forwards with LR layers -> outputs [S,B,C], kl [] (computed once, SURVEY B-8)."""

    @staticmethod
    def forward(ctx, x2, sigma_p, S, sample, calc_kl, tf32, *flat):
        L.require_cuda(x2, *flat)
        x2 = _f32c(x2)
        params = [tuple(_f32c(t) for t in flat[i:i + 4]) for i in range(0, len(flat), 4)]
        eps = plan_eps(_lr_eps_shapes(params, x2.shape[0]), S, x2.device, sample)
        kl = torch.zeros(1, dtype=torch.float64, device=x2.device)
        ys, deltas = _net_lr_forward(x2, params, sigma_p, S, eps, sample, calc_kl, kl, tf32)
        ctx.save_for_backward(x2, *flat, *ys, *[d for d in deltas if d is not None], *eps.tensors())
        ctx.cfg = (sigma_p, S, sample, calc_kl, eps, len(params), x2.requires_grad)
        ctx.tf32 = tf32
        ctx.set_materialize_grads(False)
        return ys[-1], kl.to(torch.float32)[0]

    @staticmethod
    def backward(ctx, d_out, dkl):
        sigma_p, S, sample, calc_kl, eps, nl, x_rg = ctx.cfg
        sv = ctx.saved_tensors
        x2 = sv[0]
        params = [tuple(_f32c(t) for t in sv[1 + 4 * i:5 + 4 * i]) for i in range(nl)]
        ys = list(sv[1 + 4 * nl:1 + 5 * nl])
        deltas = list(sv[1 + 5 * nl:1 + 6 * nl]) if sample else [None] * nl
        if d_out is None:
            d_out = torch.zeros_like(ys[-1])
        use_kl = calc_kl and dkl is not None
        g_kl_dev = _f32c(dkl).reshape(1) if use_kl else None
        dx0, grads = _net_lr_backward(x2, ys, deltas, _f32c(d_out), params, sigma_p, S, eps, sample, use_kl, 1.0,
                                      g_kl_dev, None, x_rg, ctx.tf32)
        flat = [g for lg in grads for g in lg]
        return (dx0.sum(0) if x_rg else None, None, None, None, None, None, *flat)


def mlp_forward_lr(x2, layers, sigma_p, S, sample=True, calc_kl=True, tf32=False):
    flat = [t for layer in layers for t in layer]
    return _MLPForwardLR.apply(x2, float(sigma_p), S, sample, calc_kl, tf32, *flat)


class _FusedELBOLR(torch.autograd.Function):
    """sample_elbo_lr (networks.py:211-225): loss = beta * KL + NLL / S, KL evaluated once."""

    @staticmethod
    def forward(ctx, x2, target, beta, S, sigma, mode, sigma_p, tf32, *flat):
        L.require_cuda(x2, target, *flat)
        x2 = _f32c(x2)
        dev = x2.device
        params = [tuple(_f32c(t) for t in flat[i:i + 4]) for i in range(0, len(flat), 4)]
        eps = plan_eps(_lr_eps_shapes(params, x2.shape[0]), S, dev, True)
        acc = torch.zeros(2, dtype=torch.float64, device=dev)
        kl, nll = acc[0:1], acc[1:2]
        ys, deltas = _net_lr_forward(x2, params, sigma_p, S, eps, True, True, kl, tf32)
        out = ys[-1]
        B, Cc = out.shape[1], out.shape[2]
        need_grad = any(t.requires_grad for t in flat)
        d_out = torch.empty_like(out) if need_grad else None
        if mode == 'classification':
            L.check(L.lib().bbb_nll_ce(L.ptr(out), L.ptr(target), S, B, Cc, 1.0 / S, L.ptr(nll), L.ptr(d_out),
                                       L.stream()), 'bbb_nll_ce')
        else:
            tgt = _f32c(target)
            L.check(L.lib().bbb_nll_gauss(L.ptr(out), L.ptr(tgt), float(sigma), S, B, Cc, 1.0 / S, L.ptr(nll),
                                          L.ptr(d_out), L.stream()), 'bbb_nll_gauss')
        out4 = torch.empty(4, dtype=torch.float32, device=dev)
        beta_h, beta_d = _split_beta(beta)
        L.check(L.lib().bbb_elbo_finalize(None, None, L.ptr(kl), L.ptr(nll), S, beta_h, L.ptr(beta_d), L.ptr(out4),
                                          L.stream()), 'bbb_elbo_finalize')
        if need_grad:
            ctx.save_for_backward(x2, d_out, *flat, *ys[:-1], *deltas, *eps.tensors())
        ctx.cfg = (sigma_p, S, beta_h, eps, len(params))
        ctx.tf32 = tf32
        ctx.set_materialize_grads(False)
        ctx.beta_dev = beta_d
        loss, klm, nl = out4[0:1], out4[1], out4[2:3]
        ctx.mark_non_differentiable(klm, nl)
        return loss, klm, nl

    @staticmethod
    def backward(ctx, g_loss, *unused):
        sigma_p, S, beta, eps, nl = ctx.cfg
        if g_loss is None:
            return (None,) * (8 + 4 * nl)
        if ctx.tf32 and getattr(ctx, 'consumed', False):
            # the tcgen05 LR backward overwrites the saved delta buffers with dV (they have no other use in it)
            raise RuntimeError('sample_elbo_lr: this graph was already differentiated (retain_graph=True is not supported '
                               'by the TF32 local-reparameterisation path: its saved state is single-use)')
        ctx.consumed = True
        sv = ctx.saved_tensors
        x2, d_out = sv[0], sv[1]
        params = [tuple(_f32c(t) for t in sv[2 + 4 * i:6 + 4 * i]) for i in range(nl)]
        o = 2 + 4 * nl
        ys = list(sv[o:o + nl - 1]) + [None]
        deltas = list(sv[o + nl - 1:o + 2 * nl - 1])
        scale = _f32c(g_loss).reshape(1)
        _, grads = _net_lr_backward(x2, ys, deltas, d_out, params, sigma_p, S, eps, True, True, beta, ctx.beta_dev,
                                    scale, False, ctx.tf32)
        flat = [g for lg in grads for g in lg]
        return (None,) * 8 + tuple(flat)


def fused_elbo_lr(x2, target, beta, S, sigma, mode, sigma_p, layers, tf32=False):
    flat = [t for layer in layers for t in layer]
    return _FusedELBOLR.apply(x2, target, beta, S, sigma, mode, float(sigma_p), tf32, *flat)


# =========================================================================================
# stand-alone reductions and the eps stream (tests, calculate_log_probs, diagnostics)
# =========================================================================================
def logprob_reduce(mu, rho, prior, eps=None, sample=True, seed=0, step=0, sample_idx=0, tensor_id=0,
                   return_w=False):
    L.require_cuda(mu, rho)
    mu, rho = _f32c(mu), _f32c(rho)
    n = mu.numel()
    acc = torch.zeros(2, dtype=torch.float64, device=mu.device)
    w = torch.empty_like(mu) if return_w else None
    e = _f32c(eps) if eps is not None else None
    L.check(L.lib().bbb_logprob_reduce(L.ptr(mu), L.ptr(rho), L.ptr(e), seed, step, sample_idx, tensor_id,
                                       C.byref(prior), n, L.F_SAMPLE if sample else 0, L.ptr(w), L.ptr(acc[0:1]),
                                       L.ptr(acc[1:2]), L.stream()), 'bbb_logprob_reduce')
    return acc[0], acc[1], w


def kl_gauss(mu, rho, sigma_p):
    L.require_cuda(mu, rho)
    mu, rho = _f32c(mu), _f32c(rho)
    acc = torch.zeros(1, dtype=torch.float64, device=mu.device)
    L.check(L.lib().bbb_kl_gauss(L.ptr(mu), L.ptr(rho), float(sigma_p), mu.numel(), L.ptr(acc), L.stream()),
            'bbb_kl_gauss')
    return acc[0]


def philox_normal(n, device, seed=0, step=0, sample_idx=0, tensor_id=0):
    out = torch.empty(n, dtype=torch.float32, device=device)
    L.require_cuda(out)
    L.check(L.lib().bbb_philox_fill_normal(L.ptr(out), n, seed, step, sample_idx, tensor_id, L.stream()),
            'bbb_philox_fill_normal')
    return out


# =========================================================================================
# consumers of the (mu, rho) stream and of the sampled outputs
# =========================================================================================
def compute_snr(mu, rho):
    """10 log10(|mu| / softplus(rho)) in decibels (weight_pruning.py:81-83), one kernel pass."""
    L.require_cuda(mu, rho)
    mu, rho = _f32c(mu), _f32c(rho)
    out = torch.empty_like(mu)
    L.check(L.lib().bbb_snr(L.ptr(mu), L.ptr(rho), mu.numel(), L.ptr(out), L.stream()), 'bbb_snr')
    return out


def snr_prune_(mu, rho, threshold_db):
    """In place: mu, rho *= (snr > threshold_db) (weight_pruning.py:101-115).  Returns the device count of kept entries."""
    L.require_cuda(mu, rho)
    if not (mu.dtype is torch.float32 and mu.is_contiguous() and rho.dtype is torch.float32 and rho.is_contiguous()):
        raise RuntimeError('snr_prune_ works in place on contiguous fp32 tensors')
    kept = torch.zeros(1, dtype=torch.int64, device=mu.device)
    L.check(L.lib().bbb_snr_prune(L.ptr(mu), L.ptr(rho), mu.numel(), float(threshold_db), L.ptr(kept), L.stream()),
            'bbb_snr_prune')
    return kept


def softmax_mean(logits):
    """[S,B,C] sampled logits -> [B,C] mean over the samples of softmax (class_task.py:84-86), one launch."""
    L.require_cuda(logits)
    z = _f32c(logits)
    S, B, Cc = z.shape
    out = torch.empty((B, Cc), dtype=torch.float32, device=z.device)
    L.check(L.lib().bbb_softmax_mean(L.ptr(z), S, B, Cc, L.ptr(out), L.stream()), 'bbb_softmax_mean')
    return out
