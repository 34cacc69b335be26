"""Closed-form numpy restatement (float64 by default) of the BBB hot path with
ANALYTIC gradients.  TEST INFRASTRUCTURE ONLY -- second opinion beside
oracle/bbb_oracle.py (which differentiates by autograd like the reference does).

It states, without autograd, exactly what the CUDA kernels compute:

  forward   sigma = log(1+e^rho), w = mu + sigma*eps, y = x w^T + b   networks.py:39,43,88
  log q     sum(-0.5 log 2pi - log sigma - eps^2/2)                    networks.py:46 with (w-mu)/sigma = eps
  log p     stable log-sum-exp form of networks.py:24-27, or the Gaussian of networks.py:67,82
  loss      (beta/S) sum_s (log q_s - log p_s) + (1/S) sum_s NLL_s     networks.py:205-208
  backward  d/dmu  = c (G + beta w R(w)),  d/drho = c sigmoid(rho) ((G + beta w R(w)) eps - beta/sigma)
            with c = 1/S, G = dNLL/dw, R = r1/s1^2 + r2/s2^2 (mixture responsibilities) or 1/sp^2.

Parity status: PINNED by tests/golden/*.npz (generated from the live reference by
oracle/make_golden.py); checked in tests/test_oracle_golden.py.
"""
import math

import numpy as np

HALF_LOG_2PI = 0.5 * math.log(2 * math.pi)


def softplus(rho):
    return np.log1p(np.exp(rho))


def sigmoid(rho):
    return 1.0 / (1.0 + np.exp(-rho))


def log_q(sigma, eps):
    return float(np.sum(-HALF_LOG_2PI - np.log(sigma) - 0.5 * eps * eps))


def prior_terms(w, prior):
    """Returns (sum log p(w), R(w)) with d log p / dw = -w R(w)."""
    if prior[0] == 'mixture':
        _, pi, s1, s2 = prior
        a = math.log(pi) - math.log(s1) - w * w / (2 * s1 * s1)
        b = math.log1p(-pi) - math.log(s2) - w * w / (2 * s2 * s2)
        m = np.maximum(a, b)
        ea, eb = np.exp(a - m), np.exp(b - m)
        lse = m + np.log(ea + eb)
        r1 = ea / (ea + eb)
        R = r1 / (s1 * s1) + (1.0 - r1) / (s2 * s2)
        return float(np.sum(lse - HALF_LOG_2PI)), R
    sp = prior[1]
    lp = float(np.sum(-w * w / (2 * sp * sp) - math.log(sp) - HALF_LOG_2PI))
    return lp, np.full_like(w, 1.0 / (sp * sp))


def nll_and_grad(out, target, mode, sigma):
    """NLL (networks.py:183-190) and its gradient w.r.t. `out` [B, C]."""
    if mode == 'classification':
        z = out - out.max(axis=1, keepdims=True)
        lse = np.log(np.exp(z).sum(axis=1, keepdims=True))
        logp = z - lse
        idx = np.arange(out.shape[0])
        val = -float(logp[idx, target].sum())
        g = np.exp(logp)
        g[idx, target] -= 1.0
        return val, g
    # regression: -sum log N(target; out, sigma) with torch broadcasting of (out, target)
    diff = out - target                      # broadcasts exactly like Normal(out, s).log_prob(target)
    val = float(np.sum(diff * diff / (2 * sigma * sigma) + math.log(sigma) + HALF_LOG_2PI))
    g_full = diff / (sigma * sigma)
    # reduce the broadcast gradient back onto out's shape
    g = g_full
    while g.ndim > out.ndim:
        g = g.sum(axis=0)
    for ax, n in enumerate(out.shape):
        if g.shape[ax] != n:
            g = g.sum(axis=ax, keepdims=True)
    return val, g


def elbo_step(x, target, layers, prior, eps, beta, mode, sigma=1.0):
    """Weight-sampling ELBO value and analytic gradients.

    layers: list of (w_mu [out,in], w_rho, b_mu [out], b_rho) numpy arrays
    eps:    eps[s][l] = (eps_w, eps_b)
    Returns dict(loss, log_prior, log_post, nll, outputs [S,B,C], grads=[(gwm, gwr, gbm, gbr), ...]).
    """
    S = len(eps)
    L = len(layers)
    dt = layers[0][0].dtype
    if mode == 'classification':
        x = x.reshape(-1, layers[0][0].shape[1])
    grads = [tuple(np.zeros_like(p) for p in layer) for layer in layers]
    lp_s, lq_s, nll_tot, outs = [], [], 0.0, []
    c = 1.0 / S
    for s in range(S):
        acts, ws, sig, Rs = [x.astype(dt)], [], [], []
        lp, lq = 0.0, 0.0
        h = acts[0]
        for l, (wm, wr, bm, br) in enumerate(layers):
            ew, eb = eps[s][l]
            sw, sb = softplus(wr), softplus(br)
            w, b = wm + sw * ew, bm + sb * eb
            lpw, Rw = prior_terms(w, prior)
            lpb, Rb = prior_terms(b, prior)
            lp += lpw + lpb
            lq += log_q(sw, ew) + log_q(sb, eb)
            z = h @ w.T + b
            ws.append((w, b)); sig.append((sw, sb)); Rs.append((Rw, Rb))
            h = np.maximum(z, 0) if l + 1 < L else z
            acts.append(h)
        lp_s.append(lp); lq_s.append(lq); outs.append(h)
        val, dz = nll_and_grad(h, target, mode, sigma)
        nll_tot += val
        for l in reversed(range(L)):
            wm, wr, bm, br = layers[l]
            ew, eb = eps[s][l]
            (w, b), (sw, sb), (Rw, Rb) = ws[l], sig[l], Rs[l]
            Gw = dz.T @ acts[l]
            Gb = dz.sum(axis=0)
            tw = Gw + beta * w * Rw
            tb = Gb + beta * b * Rb
            gwm, gwr, gbm, gbr = grads[l]
            gwm += c * tw
            gwr += c * sigmoid(wr) * (tw * ew - beta / sw)
            gbm += c * tb
            gbr += c * sigmoid(br) * (tb * eb - beta / sb)
            if l > 0:
                dz = (dz @ w) * (acts[l] > 0)
    lp_m, lq_m = float(np.mean(lp_s)), float(np.mean(lq_s))
    nll_m = nll_tot / S
    return dict(loss=beta * lq_m - beta * lp_m + nll_m, log_prior=lp_m, log_post=lq_m,
                nll=nll_m, outputs=np.stack(outs), grads=grads,
                log_prior_s=np.array(lp_s), log_post_s=np.array(lq_s))


def gaussian_kl(mu, sigma, sp):
    return float(0.5 * np.sum(2 * np.log(sp / sigma) - 1 + (sigma / sp) ** 2 + (mu / sp) ** 2))


def elbo_step_lr(x, target, layers, sigma_p, eps, beta, mode, sigma=1.0):
    """Local-reparameterisation ELBO value and analytic gradients.

    layers: list of (w_mu [in,out], w_rho, b_mu [out], b_rho);  eps[s][l] = (eps_a [B,out], eps_b [out]).
    """
    S = len(eps)
    L = len(layers)
    dt = layers[0][0].dtype
    if mode == 'classification':
        x = x.reshape(-1, layers[0][0].shape[0])
    grads = [tuple(np.zeros_like(p) for p in layer) for layer in layers]
    sp2 = sigma_p * sigma_p
    kl = 0.0
    for l, (wm, wr, bm, br) in enumerate(layers):
        sw, sb = softplus(wr), softplus(br)
        kl += gaussian_kl(wm, sw, sigma_p) + gaussian_kl(bm, sb, sigma_p)
        gwm, gwr, gbm, gbr = grads[l]
        gwm += beta * wm / sp2
        gwr += beta * sigmoid(wr) * (-1.0 / sw + sw / sp2)
        gbm += beta * bm / sp2
        gbr += beta * sigmoid(br) * (-1.0 / sb + sb / sp2)
    nll_tot, outs = 0.0, []
    c = 1.0 / S
    for s in range(S):
        acts, deltas = [x.astype(dt)], []
        h = acts[0]
        for l, (wm, wr, bm, br) in enumerate(layers):
            ea, eb = eps[s][l]
            sw, sb = softplus(wr), softplus(br)
            delta = np.sqrt((h * h) @ (sw * sw))
            z = h @ wm + delta * ea + (bm + sb * eb)
            deltas.append(delta)
            h = np.maximum(z, 0) if l + 1 < L else z
            acts.append(h)
        outs.append(h)
        val, dz = nll_and_grad(h, target, mode, sigma)
        nll_tot += val
        for l in reversed(range(L)):
            wm, wr, bm, br = layers[l]
            ea, eb = eps[s][l]
            sw, sb = softplus(wr), softplus(br)
            xl = acts[l]
            with np.errstate(divide='ignore', invalid='ignore'):
                dV = dz * ea / (2 * deltas[l])
            dV = np.where(deltas[l] > 0, dV, 0.0)          # App. B-7 guard
            gwm, gwr, gbm, gbr = grads[l]
            gwm += c * (xl.T @ dz)
            gwr += c * sigmoid(wr) * (2 * sw * ((xl * xl).T @ dV))
            gsum = dz.sum(axis=0)
            gbm += c * gsum
            gbr += c * sigmoid(br) * (gsum * eb)
            if l > 0:
                dz = (dz @ wm.T + 2 * xl * (dV @ (sw * sw).T)) * (xl > 0)
    nll_m = nll_tot / S
    return dict(loss=beta * kl + nll_m, kl=kl, nll=nll_m, outputs=np.stack(outs), grads=grads)
