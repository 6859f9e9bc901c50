"""Public API for Metropolis Adjusted Langevin kernels — B200 drop-in for
bblackjax/mcmc/mala.py (MALAState :16-28, MALAInfo :31-48, init :51-54, build_kernel :57-120,
mala :123-189).

Differences forced by the platform (see DESIGN.md "Boundary"):
  * the functions operate on a BATCH of chains (what `jax.vmap(kernel)` produced in the reference):
    `position` is [N, d], `rng_key` is uint32[N, 2] (one key per chain, as vmap hands them over);
  * `logdensity_fn` must be a device log-density built by `mfm_b200.distributions`
    (`dist.tempered(beta)`): a CUDA kernel cannot call or differentiate a Python closure, and there
    is no autodiff/CPU fallback, so an unknown callable raises TypeError.
"""
from typing import Callable, NamedTuple, Tuple

import torch

from .. import base
from ... import _lib
from ...distributions import DeviceLogDensity

__all__ = ["MALAState", "MALAInfo", "init", "build_kernel", "mala"]


class MALAState(NamedTuple):
    position: torch.Tensor
    logdensity: torch.Tensor
    logdensity_grad: torch.Tensor


class MALAInfo(NamedTuple):
    acceptance_rate: torch.Tensor
    is_accepted: torch.Tensor
    proposed_position: torch.Tensor
    proposed_weight: torch.Tensor


def _require_device_fn(logdensity_fn) -> DeviceLogDensity:
    if not isinstance(logdensity_fn, DeviceLogDensity):
        raise TypeError(
            "mfm_b200 MALA needs a device log-density (mfm_b200.distributions.<Dist>.tempered(beta)); "
            "arbitrary Python callables cannot be evaluated/differentiated on the GPU and there is no fallback.")
    return logdensity_fn


def init(position: torch.Tensor, logdensity_fn: Callable) -> MALAState:
    fn = _require_device_fn(logdensity_fn)
    logdensity, grad = fn.value_and_grad(position)
    return MALAState(position, logdensity, grad)


def build_kernel():
    """Returns kernel(rng_key[N,2], state, logdensity_fn, step_size) -> (MALAState, MALAInfo)."""

    def kernel(rng_key: torch.Tensor, state: MALAState, logdensity_fn: Callable, step_size: float
               ) -> Tuple[MALAState, MALAInfo]:
        fn = _require_device_fn(logdensity_fn)
        return mala_step(fn, rng_key, state, step_size, per_chain_keys=True)

    return kernel


def mala_step(fn: DeviceLogDensity, rng_key, state: MALAState, step_size: float, per_chain_keys: bool,
              chain_offset: int = 0, n_total: int = None, inplace: bool = False):
    lib = _lib.load()
    x, l, g = state
    n, d = x.shape
    if not inplace:
        x, l, g = x.clone(), l.clone(), g.clone()
    dev = x.device
    acc_rate = torch.empty(n, dtype=torch.float32, device=dev)
    is_acc = torch.empty(n, dtype=torch.uint8, device=dev)
    prop = torch.empty((n, d), dtype=torch.float32, device=dev)
    weight = torch.empty(n, dtype=torch.float32, device=dev)
    desc = fn.desc()
    ws = _lib.workspace(lib.mfm_mala_workspace_bytes(desc, n), dev, "mala")
    if per_chain_keys:
        assert rng_key.shape == (n, 2), "kernel expects one key per chain (uint32[N,2])"
    _lib.check(lib.mfm_mala_step(desc, _lib.ptr(rng_key.contiguous()), 1 if per_chain_keys else 0, n, chain_offset,
                                 n_total if n_total is not None else n, float(step_size),
                                 _lib.ptr(x), _lib.ptr(l), _lib.ptr(g), _lib.ptr(acc_rate), _lib.ptr(is_acc),
                                 _lib.ptr(prop), _lib.ptr(weight), _lib.ptr(ws), ws.numel(), _lib.stream()))
    return MALAState(x, l, g), MALAInfo(acc_rate, is_acc.bool(), prop, weight)


class mala:
    """`mala(logdensity_fn, step_size)` -> SamplingAlgorithm(init, step) (mala.py:123-189)."""

    init = staticmethod(init)
    build_kernel = staticmethod(build_kernel)

    def __new__(cls, logdensity_fn: Callable, step_size: float) -> base.SamplingAlgorithm:
        kernel = cls.build_kernel()

        def init_fn(position):
            return cls.init(position, logdensity_fn)

        def step_fn(rng_key, state):
            return kernel(rng_key, state, logdensity_fn, step_size)

        return base.SamplingAlgorithm(init_fn, step_fn)
