"""Shared helpers for GPU parity tests (oracle <-> CUDA)."""
import numpy as np
import torch


def to_dev(a, dev, dtype=torch.float32):
    return torch.as_tensor(np.ascontiguousarray(a)).to(dtype).to(dev)


def key_dev(k, dev):
    return torch.from_numpy(np.ascontiguousarray(np.asarray(k, np.uint32))).to(dev)


def rel_err(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def ulp_diff_f32(a, b):
    a = np.asarray(a, np.float32).view(np.int32).astype(np.int64)
    b = np.asarray(b, np.float32).view(np.int32).astype(np.int64)
    a = np.where(a < 0, -(a & 0x7FFFFFFF), a); b = np.where(b < 0, -(b & 0x7FFFFFFF), b)
    return np.abs(a - b)


def make_targets(dev):
    """(name, oracle target, device distribution) triples for every configured target."""
    from oracle import targets as OT
    from mfm_b200 import distributions as D
    out = []
    t4 = OT.four_mode()
    out.append(("4-mode", t4, D.GaussianMixture(t4.modes, t4.covs, t4.weights, device=dev)))
    t16 = OT.gmm16()
    out.append(("gmm16", t16, D.GaussianMixture(t16.modes, t16.covs, t16.weights, device=dev)))
    out.append(("phi-four", OT.PhiFour(64), D.PhiFour(64, device=dev)))
    out.append(("pines", OT.LogGaussianCoxPines(1600), D.LogGaussianCoxPines(1600, device=dev)))
    return out
