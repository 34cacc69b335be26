"""Where eps comes from.

'philox' (default): eps is generated inside the kernels by Philox4x32-10 keyed by
    (seed; element/4, global MC-sample index, tensor id, step) and never stored.  The host owns
    (seed, step): every sampling call consumes one step, so a run is replayable and shardable.
'reference': deterministic parity mode -- eps is drawn from torch's global CPU generator with
    torch.randn in exactly the reference's order (networks.py:35,42) and injected.
"""
import threading

import torch

_state = threading.local()


def _st():
    if not hasattr(_state, 'mode'):
        _state.mode, _state.seed, _state.step, _state.sample_base = 'philox', 0x5EEDB200, 0, 0
        _state.step_dev = None
    return _state


def set_eps_mode(mode):
    assert mode in ('philox', 'reference', 'injected')
    _st().mode = mode


def set_injected_eps(tensors):
    """'injected' mode: a list of tensors consumed in the reference's draw order (tests)."""
    _st().queue = list(tensors)


def pop_injected(shape):
    t = _st().queue.pop(0)
    assert tuple(t.shape) == tuple(shape), (tuple(t.shape), tuple(shape))
    return t


def get_eps_mode():
    return _st().mode


def manual_seed(seed, step=0):
    st = _st()
    st.seed, st.step = int(seed) & 0xFFFFFFFFFFFFFFFF, int(step)


def set_sample_base(base):
    """Global index of this process's first MC sample (multi-GPU sample sharding)."""
    _st().sample_base = int(base)


def get_sample_base():
    return _st().sample_base


def next_step():
    """Consume one Philox step (host side)."""
    st = _st()
    s = st.step
    st.step = (st.step + 1) & 0xFFFFFFFF
    return st.seed, s


def use_device_step(counter):
    """Route the step through a device uint32 tensor (CUDA-graph replay); None switches back."""
    _st().step_dev = counter


def device_step():
    return _st().step_dev


class eps_mode:
    """Context manager: with eps_mode('reference'): ..."""

    def __init__(self, mode):
        self.mode = mode

    def __enter__(self):
        self.prev = get_eps_mode()
        set_eps_mode(self.mode)

    def __exit__(self, *a):
        set_eps_mode(self.prev)


def draw_reference_eps(shape, device):
    """One reference-order draw: Normal(0,1).sample(shape) on the CPU generator, then to(device)."""
    return torch.randn(shape).to(device)
