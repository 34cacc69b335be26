"""BayesianNetwork (networks.py:140-225) on the CUDA kernels, plus the two stock baselines the
reference's callers import by name (MLP, MLP_Dropout: plain torch.nn, not part of the hot path)."""
import os

import torch
from torch import nn

from . import functional as F
from .layers import BayesianLinear, BayesianLinearLR


class BayesianNetwork(nn.Module):
    """Same model_params dict, attributes (l1, l1_act, ...), methods and return tuples as the reference.

    Extensions (strict supersets): 'local_reparam' defaults to False when absent, which is what lets
    reinforcement_learning/bandits.py construct the net; 'hidden_units' may be a list
    (or 'n_hidden' may be given) for nets deeper than the reference's fixed two hidden layers
    (BASELINE.json config 5); 'tf32' opts in to the tcgen05 kind::tf32 tensor path.
    """

    def __init__(self, model_params):
        super().__init__()
        self.input_shape = model_params['input_shape']
        self.classes = model_params['classes']
        self.batch_size = model_params['batch_size']
        self.hidden_units = model_params['hidden_units']
        self.mode = model_params['mode']
        self.mu_init = model_params['mu_init']
        self.rho_init = model_params['rho_init']
        self.prior_init = model_params['prior_init']
        self.mixture_prior = model_params['mixture_prior']
        self.local_reparam = model_params.get('local_reparam', False)
        # 'tf32' key, else the process-wide default BBB_TF32=1 (the switch for callers that build model_params themselves:
        # reg_task.py / class_task.py / bandits.py run unchanged and cannot pass the key)
        self.tf32 = bool(model_params.get('tf32', os.environ.get('BBB_TF32', '0') == '1'))
        self.fused = bool(model_params.get('fused', True))

        if isinstance(self.hidden_units, (list, tuple)):
            hidden = list(self.hidden_units)
        else:
            hidden = [self.hidden_units] * int(model_params.get('n_hidden', 2))
        dims = [self.input_shape] + hidden + [self.classes]
        layer = BayesianLinearLR if self.local_reparam else BayesianLinear
        self.n_layers = len(dims) - 1
        for i in range(self.n_layers):
            l = layer(dims[i], dims[i + 1], self.mu_init, self.rho_init, self.prior_init, self.mixture_prior)
            l.layer_id = i
            if not self.local_reparam:
                l.tf32 = self.tf32
            setattr(self, f'l{i + 1}', l)
            if i + 1 < self.n_layers:
                setattr(self, f'l{i + 1}_act', nn.ReLU())
        self._prior = None if self.local_reparam else F.make_prior(self.prior_init, self.mixture_prior)
        self._fused_opt = None

    def layers(self):
        return [getattr(self, f'l{i + 1}') for i in range(self.n_layers)]

    def fuse_optimizer(self, optimizer, overlap=False):
        """Opt-in (single GPU): let the backward of sample_elbo apply `optimizer`'s (bnn_b200.FusedAdam) next step --
        loss.backward() then UPDATES the parameters and leaves .grad unset, and the optimizer.step() that follows only
        advances the step count.  overlap=False: the update rides in the backward kernels' gradient epilogue (measured
        slower); overlap=True: the ordinary backward kernels run and every layer's Adam update is launched
        on a side stream as soon as that layer's backward has finished, under the backward of the layer below
        (BBB_F_ADAM_OVERLAP; network-level path only).  Returns False (and changes nothing) when this network cannot
        run the fused tcgen05 backward.  fuse_optimizer(None) switches back."""
        if optimizer is None:
            self._fused_opt = None
            return True
        optimizer.overlap_backward = bool(overlap)
        ok = (self.tf32 and self.fused and not self.local_reparam and self.batch_size <= 128 and
              all(l.weight_mu.shape[1] % 4 == 0 for l in self.layers()) and hasattr(optimizer, 'fuse_descriptor'))
        self._fused_opt = optimizer if ok else None
        return ok

    def forward(self, x, sample=False):
        if self.mode == 'classification':
            x = x.view(-1, self.input_shape)  # Flatten images
        for i, l in enumerate(self.layers()):
            x = l(x, sample)
            if i + 1 < self.n_layers:
                x = getattr(self, f'l{i + 1}_act')(x)
        return x

    # ---- prediction: the reference's callers loop `net(x, sample=True)` test_samples times
    # (class_task.py:81-87, reg_task.py:76-83); these run all the sampled forwards in one launch per layer ----
    def sample_predict(self, x, samples):
        """`samples` sampled forward passes at once -> outputs [samples, B, classes].  Same values, in the
        reference eps mode the same draws in the same order, as `samples` calls of net(x, sample=True) in eval
        mode (no log-probs are evaluated)."""
        x2 = self._flat_input(x)
        params = [l.params() for l in self.layers()]
        with torch.no_grad():
            if self.local_reparam:
                outs, _ = F.mlp_forward_lr(x2, params, float(self.prior_init[0]), samples, True, False, self.tf32)
            else:
                outs, _, _ = F.mlp_forward(x2, params, self._prior, samples, True, False, self.tf32)
        return outs

    def predict_proba(self, x, samples):
        """BNN_Classification.predict (class_task.py:81-87): softmax of every sampled forward, averaged."""
        return F.softmax_mean(self.sample_predict(x, samples))

    # ---- SNR pruning (weight_pruning.py:81-115) on the (mu, rho) stream, one kernel pass per tensor ----
    def snr(self):
        """SNR in decibels of every weight and bias, flattened in the order weight_pruning.get_snr... walks the layers:
        per layer weights then biases.  Returns a 1-D device tensor."""
        parts = []
        for l in self.layers():
            parts.append(F.compute_snr(l.weight_mu.data, l.weight_rho.data).reshape(-1))
            parts.append(F.compute_snr(l.bias_mu.data, l.bias_rho.data).reshape(-1))
        return torch.cat(parts)

    def prune_weights(self, snrs, drop_percentage=0.5):
        """prune_weights(model, snrs, drop_percentage) of weight_pruning.py:85-115: the SNR threshold is the
        100 drop_percentage-th percentile of `snrs` (numpy semantics), then mu and rho of every weight and bias whose
        SNR is not above it are multiplied by zero, in place.  Returns the fraction of parameters kept."""
        import numpy as np
        snr_threshold = float(np.percentile(snrs.detach().cpu().numpy() if torch.is_tensor(snrs) else snrs,
                                            100 * drop_percentage))
        kept, total = 0, 0
        with torch.no_grad():
            for l in self.layers():
                for mu, rho in ((l.weight_mu, l.weight_rho), (l.bias_mu, l.bias_rho)):
                    kept = kept + F.snr_prune_(mu.data, rho.data, snr_threshold)
                    total += mu.numel()
        return float(kept) / max(total, 1)

    def log_prior(self):
        return sum(l.log_prior for l in self.layers())

    def log_variational_posterior(self):
        return sum(l.log_variational_posterior for l in self.layers())

    def kl_cost(self):
        return sum(l.kl_cost for l in self.layers())

    def get_nll(self, outputs, target, sigma=1.):
        if self.mode == 'regression':
            # -Normal(outputs, sigma).log_prob(target).sum() (networks.py:185) written out -- same arithmetic and the same
            # broadcasting of target against outputs -- because the distribution object turns a python
            # sigma into a device tensor with a host-to-device copy, which a CUDA-graph capture does not allow
            if torch.is_tensor(sigma):
                nll = -torch.distributions.Normal(outputs, sigma).log_prob(target).sum()
            else:
                import math
                var = float(sigma) ** 2
                nll = -(-((target - outputs) ** 2) / (2 * var) - math.log(float(sigma)) - math.log(math.sqrt(2 * math.pi))).sum()
        elif self.mode == 'classification':
            nll = nn.CrossEntropyLoss(reduction='sum')(outputs, target)
        else:
            raise Exception("Training mode must be either 'regression' or 'classification'")
        return nll

    # ---- fused-path eligibility: the likelihood kernels cover the reference's two regular cases --------
    def _fusable(self, x2, target):
        if not self.fused:
            return False
        B = x2.shape[0]
        if self.mode == 'classification':
            return target.dtype == torch.int64 and target.dim() == 1 and target.shape[0] == B
        if self.mode == 'regression':      # [B, D] targets; the [B] vs [B,1] broadcast (SURVEY B-3) goes the general way
            return target.dim() == 2 and tuple(target.shape) == (B, self.classes) and target.is_floating_point()
        return False

    def _flat_input(self, input):
        return input.view(-1, self.input_shape) if self.mode == 'classification' else input.reshape(-1, self.input_shape)

    def sample_elbo(self, input, target, beta, samples, sigma=1.):
        ''' Sample ELBO for BNN w/o Local Reparameterisation (networks.py:192-209) '''
        assert self.local_reparam == False, 'sample_elbo() method returns loss for BNNs without local reparameterisation, alternatively use sample_elbo_lr()'
        x2 = self._flat_input(input)
        params = [l.params() for l in self.layers()]
        if self._fusable(x2, target):
            fo = self._fused_opt if (self._fused_opt is not None and self.training and x2.shape[0] <= 128) else None
            return F.fused_elbo(x2, target, beta, samples, sigma, self.mode, self._prior, params, self.tf32, fo)
        outs, lps, lqs = F.mlp_forward(x2, params, self._prior, samples, True, True, self.tf32)
        negative_log_likelihood = torch.zeros(1, device=outs.device)
        for i in range(samples):
            negative_log_likelihood = negative_log_likelihood + self.get_nll(outs[i], target, sigma)
        log_prior = beta * lps.mean()
        log_variational_posterior = beta * lqs.mean()
        negative_log_likelihood = negative_log_likelihood / samples
        loss = log_variational_posterior - log_prior + negative_log_likelihood
        return loss, lps.mean(), lqs.mean(), negative_log_likelihood

    def sample_elbo_lr(self, input, target, beta, samples, sigma=1.):
        ''' Sample ELBO for BNN w/ Local Reparameterisation (networks.py:211-225) '''
        assert self.local_reparam == True, 'sample_elbo_lr() method returns loss for BNNs with local reparameterisation, alternatively use sample_elbo()'
        x2 = self._flat_input(input)
        params = [l.params() for l in self.layers()]
        sigma_p = float(self.prior_init[0])
        if self._fusable(x2, target):
            return F.fused_elbo_lr(x2, target, beta, samples, sigma, self.mode, sigma_p, params, self.tf32)
        outs, kl = F.mlp_forward_lr(x2, params, sigma_p, samples, True, True, self.tf32)
        negative_log_likelihood = torch.zeros(1, device=outs.device)
        for i in range(samples):
            negative_log_likelihood = negative_log_likelihood + self.get_nll(outs[i], target, sigma)
        negative_log_likelihood = negative_log_likelihood / samples
        loss = beta * kl + negative_log_likelihood
        return loss, kl, negative_log_likelihood


class MLP(nn.Module):
    """Non-Bayesian baseline (networks.py:227-253): stock torch.nn, kept so `from networks import MLP` works."""

    def __init__(self, model_params):
        super().__init__()
        self.input_shape = model_params['input_shape']
        self.classes = model_params['classes']
        self.batch_size = model_params['batch_size']
        self.hidden_units = model_params['hidden_units']
        self.mode = model_params['mode']
        self.net = nn.Sequential(nn.Linear(self.input_shape, self.hidden_units), nn.ReLU(),
                                 nn.Linear(self.hidden_units, self.hidden_units), nn.ReLU(),
                                 nn.Linear(self.hidden_units, self.classes))

    def forward(self, x):
        assert self.mode in {'regression', 'classification'}, 'MLP Mode must be either regression or classification'
        if self.mode == 'classification':
            x = x.view(-1, self.input_shape)
        return self.net(x)


class MLP_Dropout(nn.Module):
    """MC-dropout baseline (networks.py:255-285): stock torch.nn."""

    def __init__(self, model_params):
        super().__init__()
        self.input_shape = model_params['input_shape']
        self.classes = model_params['classes']
        self.batch_size = model_params['batch_size']
        self.hidden_units = model_params['hidden_units']
        self.mode = model_params['mode']
        self.net = nn.Sequential(nn.Linear(self.input_shape, self.hidden_units), nn.ReLU(),
                                 nn.Dropout(0.5),      # fixed, as the reference (networks.py:268,271); the callers'
                                 nn.Linear(self.hidden_units, self.hidden_units), nn.ReLU(),
                                 nn.Dropout(0.5),      # 'dropout' key is a boolean routing flag, not a probability
                                 nn.Linear(self.hidden_units, self.classes))

    def forward(self, x):
        assert self.mode in {'regression', 'classification'}, 'MLP Mode must be either regression or classification'
        if self.mode == 'classification':
            x = x.view(-1, self.input_shape)
        return self.net(x)

    def enable_dropout(self):
        ''' Enable the dropout layers during test-time '''
        for m in self.modules():
            if m.__class__.__name__.startswith('Dropout'):
                m.train()
