"""jax.random-like API on device (threefry2x32, legacy uint32[2] keys, x64 off).

Mirrors the jax.random calls the reference makes (exe_flow_matching.py:333,433; util.py:81;
distributions.py:70-76,93-97,163-164,313-314).  Keys are torch.uint32 CUDA tensors.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import _lib


def _dev(device=None):
    return torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())


def PRNGKey(seed: int, device=None) -> torch.Tensor:
    """jax.random.PRNGKey(seed) with x64 off: [0, seed & 0xffffffff] (seed < 2**32)."""
    seed = int(seed)
    k = np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], dtype=np.uint32)
    return torch.from_numpy(k).to(_dev(device))


def host_split(key, num: int = 2) -> np.ndarray:
    """jax.random.split on host uint32[2] keys (no device work)."""
    lib = _lib.load()
    k = (C.c_uint32 * 2)(*[int(v) for v in np.asarray(key, dtype=np.uint32)])
    out = (C.c_uint32 * (2 * num))()
    lib.mfm_host_threefry_split(k, num, out)
    return np.frombuffer(out, dtype=np.uint32).reshape(num, 2).copy()


def split(key: torch.Tensor, num: int = 2) -> torch.Tensor:
    """jax.random.split(key, num) -> uint32[num, 2]; a batch of keys [n,2] -> [n, num, 2]."""
    lib = _lib.load()
    assert key.dtype == torch.uint32 and key.shape[-1] == 2
    if key.dim() == 1:
        out = torch.empty((num, 2), dtype=torch.uint32, device=key.device)
        _lib.check(lib.mfm_threefry_split(_lib.ptr(key), num, _lib.ptr(out), _lib.stream()))
    else:
        n = key.shape[0]
        out = torch.empty((n, num, 2), dtype=torch.uint32, device=key.device)
        _lib.check(lib.mfm_threefry_split_batched(_lib.ptr(key.contiguous()), n, num, _lib.ptr(out), _lib.stream()))
    return out


def bits(key: torch.Tensor, shape) -> torch.Tensor:
    lib = _lib.load()
    n = int(math.prod(shape)) if len(shape) else 1
    out = torch.empty(n, dtype=torch.uint32, device=key.device)
    _lib.check(lib.mfm_threefry_bits(_lib.ptr(key), n, _lib.ptr(out), _lib.stream()))
    return out.reshape(tuple(shape))


def uniform(key: torch.Tensor, shape=(), minval=0.0, maxval=1.0) -> torch.Tensor:
    """jax.random.uniform(key, shape, float32, minval, maxval); key [n,2] == vmap over keys."""
    lib = _lib.load()
    if key.dim() == 2:
        n, d = key.shape[0], int(math.prod(shape))
        out = torch.empty((n, d), dtype=torch.float32, device=key.device)
        _lib.check(lib.mfm_threefry_uniform_batched(_lib.ptr(key.contiguous()), n, d, float(minval), float(maxval),
                                                    _lib.ptr(out), _lib.stream()))
        return out.reshape((n,) + tuple(shape))
    n = int(math.prod(shape)) if len(shape) else 1
    out = torch.empty(n, dtype=torch.float32, device=key.device)
    _lib.check(lib.mfm_threefry_uniform(_lib.ptr(key), n, float(minval), float(maxval), _lib.ptr(out), _lib.stream()))
    return out.reshape(tuple(shape))


def normal(key: torch.Tensor, shape=()) -> torch.Tensor:
    """jax.random.normal(key, shape, float32).  key [n,2] with shape (d,) == vmap over keys -> [n,d]."""
    lib = _lib.load()
    if key.dim() == 2:
        n, d = key.shape[0], int(math.prod(shape))
        out = torch.empty((n, d), dtype=torch.float32, device=key.device)
        _lib.check(lib.mfm_threefry_normal_batched(_lib.ptr(key.contiguous()), n, d, _lib.ptr(out), _lib.stream()))
        return out.reshape((n,) + tuple(shape))
    n = int(math.prod(shape)) if len(shape) else 1
    out = torch.empty(n, dtype=torch.float32, device=key.device)
    _lib.check(lib.mfm_threefry_normal(_lib.ptr(key), n, _lib.ptr(out), _lib.stream()))
    return out.reshape(tuple(shape))


def choice(key: torch.Tensor, a: torch.Tensor, shape, p: torch.Tensor, return_index: bool = False):
    """jax.random.choice(key, a, shape, replace=True, p=p) along axis 0 (exe_flow_matching.py:459); p need not be
    normalised.  a: [n_pop, ...] float32 (or an int n_pop: returns indices)."""
    lib = _lib.load()
    n_draw = int(math.prod(shape)) if len(shape) else 1
    n_pop = int(a) if isinstance(a, int) else a.shape[0]
    p = p.contiguous().float()
    assert p.shape == (n_pop,), "p must have shape (a.shape[0],)"
    idx = torch.empty(n_draw, dtype=torch.int32, device=p.device)
    ws = _lib.workspace(lib.mfm_random_choice_workspace_bytes(n_pop), p.device, "choice")
    _lib.check(lib.mfm_random_choice(_lib.ptr(key), n_pop, _lib.ptr(p), n_draw, _lib.ptr(idx), _lib.ptr(ws), ws.numel(), _lib.stream()))
    if isinstance(a, int):
        return idx.reshape(tuple(shape))
    src = a.contiguous().float().reshape(n_pop, -1)
    out = torch.empty((n_draw, src.shape[1]), dtype=torch.float32, device=src.device)
    _lib.check(lib.mfm_gather_rows(_lib.ptr(src), _lib.ptr(idx), n_pop, n_draw, src.shape[1], _lib.ptr(out), _lib.stream()))
    out = out.reshape(tuple(shape) + tuple(a.shape[1:]))
    return (out, idx.reshape(tuple(shape))) if return_index else out
