#!/usr/bin/env python
"""CUDA-event timings of the batch-resident (wide-layer) kernels through the C ABI: forward, dgrad and wgrad of one
hidden layer, with the GEMM rate each reaches ('post': the input is a stored post-activation, no ReLU on load; 'nolp':
without the log-prob terms).  Also checks the three against a float64 torch restatement (mean weights,
sample=False) so that a change to them is verified and timed in one GPU call.
usage: python tools/time_wide.py [width=4096] [B=4096] [S=2]"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bnn_b200  # noqa: E402,F401
from bnn_b200 import _lib as L, functional as F  # noqa: E402

W = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
S = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dev = 'cuda'
prior = F.make_prior([0.5, 0, -6], True)
inn = out = W
torch.manual_seed(0)
wm = torch.empty(out, inn, device=dev).uniform_(-0.05, 0.05); wr = torch.empty(out, inn, device=dev).uniform_(-5, -4)
bm = torch.empty(out, device=dev).uniform_(-0.2, 0.2); br = torch.empty(out, device=dev).uniform_(-5, -4)
x = torch.randn(S, B, inn, device=dev)
y = torch.zeros(S, B, out, device=dev); dy = torch.randn(S, B, out, device=dev) / B
dx = torch.zeros(S, B, inn, device=dev)
g = [torch.zeros_like(t) for t in (wm, wr, bm, br)]
acc = torch.zeros(2 * S, dtype=torch.float64, device=dev)
rng = L.Rng(1, 0, 0, 0, None)
st = torch.cuda.current_stream().cuda_stream
xs = B * inn


def fwd(flags=L.F_SAMPLE | L.F_LOGPROB, relu=L.F_RELU_IN):
    L.check(L.lib().bbb_linear_fwd(x.data_ptr(), xs, wm.data_ptr(), wr.data_ptr(), bm.data_ptr(), br.data_ptr(), None,
                                   None, C.byref(rng), C.byref(prior), S, B, inn, out, L.F_TF32 | relu | flags,
                                   y.data_ptr(), acc[:S].data_ptr(), acc[S:].data_ptr(), st), 'fwd')


def bwd(extra, flags=L.F_SAMPLE, gp=-0.25, gq=0.25, relu=L.F_RELU_IN, pre=L.F_DX_PREACT):
    L.check(L.lib().bbb_linear_bwd(dy.data_ptr(), None, x.data_ptr(), xs, wm.data_ptr(), wr.data_ptr(), bm.data_ptr(),
                                   br.data_ptr(), None, None, C.byref(rng), C.byref(prior), S, B, inn, out,
                                   L.F_TF32 | relu | pre | flags | extra, gp, gq, None, None, 0, None,
                                   dx.data_ptr(), g[0].data_ptr(), g[1].data_ptr(), g[2].data_ptr(), g[3].data_ptr(), st),
            'bwd')


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


# ---- check (mean weights: the contraction, the ReLU on load, the (x > 0) mask of dx, the bias sums) ------------------
fwd(0)
bwd(0, 0, 0.0, 0.0)
torch.cuda.synchronize()
xr = x.double().clamp_min(0)
ok = True
for s in range(S):
    ref_y = xr[s] @ wm.double().t() + bm.double()
    e_y = float((y[s].double() - ref_y).abs().max() / ref_y.abs().max())
    ref_dx = (dy[s].double() @ wm.double()) * (x[s] > 0)
    e_dx = float((dx[s].double() - ref_dx).abs().max() / ref_dx.abs().max())
    ok &= e_y < 3e-3 and e_dx < 3e-3
    print(f'sample {s}: y err {e_y:.2e}  dx err {e_dx:.2e}')
ref_g = sum(dy[s].double().t() @ xr[s] for s in range(S))
e_g = float((g[0].double() - ref_g).abs().max() / ref_g.abs().max())
ref_gb = sum(dy[s].double().sum(0) for s in range(S))
e_gb = float((g[2].double() - ref_gb).abs().max() / ref_gb.abs().max())
ok &= e_g < 3e-3 and e_gb < 1e-4
print(f'grad_w_mu err {e_g:.2e}  grad_b_mu err {e_gb:.2e}  ->', 'OK' if ok else 'FAIL')

# ---- time ---------------------------------------------------------------------------------------------------------
flops = 2.0 * S * B * inn * out
RO = getattr(L, 'F_RELU_OUT', 0)
for name, fn in (('forward', fwd), ('fwd post', lambda: fwd(L.F_SAMPLE | L.F_LOGPROB | RO, 0)),
                 ('fwd nolp', lambda: fwd(L.F_SAMPLE | RO, 0)), ('dgrad', lambda: bwd(L.F_NO_WGRAD)), ('dgr nomk', lambda: bwd(L.F_NO_WGRAD, relu=0, pre=0)),
                 ('wgrad', lambda: bwd(L.F_NO_DX)), ('wgr post', lambda: bwd(L.F_NO_DX, relu=0))):
    us = timeit(fn)
    print(f'{name:8s} [{inn}x{out}] B={B} S={S}: {us:8.1f} us = {us / S:7.1f} us/sample, {flops / us / 1e6:6.1f} TFLOP/s', flush=True)
if S >= 4:   # the sample-group split of wgrad (two CTAs per tile, partial gradients + combine kernel): off / by shape
    for mode, name in ((-1, 'wgr 1cta'), (0, 'wgr auto')):
        L.check(L.lib().bbb_debug_wgrad_split(mode), 'split')
        us = timeit(lambda: bwd(L.F_NO_DX, relu=0))
        print(f'{name:8s} [{inn}x{out}] B={B} S={S}: {us:8.1f} us = {us / S:7.1f} us/sample, {flops / us / 1e6:6.1f} TFLOP/s', flush=True)
    L.check(L.lib().bbb_debug_wgrad_split(0), 'split')
