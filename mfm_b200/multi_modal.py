"""CLI mirror of the reference's multi_modal.py for the MFM path (same flags, same per-example overrides).

Baseline methods (--do_fab/--do_dds/--do_flowmc/--do_smc/--do_pocomc), wandb and the post-training metric
table are out of scope; the hot loop, tempering and key schedule are the reference's."""
from __future__ import annotations

import argparse
import os

import numpy as np
import torch

from .distributions import GaussianMixture, LogGaussianCoxPines, PhiFour
from .exe_flow_matching import run

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def build(args, device=None):
    """Per-example overrides (multi_modal.py:23-98)."""
    if args.example == "gaussian-mixture":
        args.dim, args.num_modes, args.lim, args.levels, args.step_size = 2, 16, [-16, 16], 20, 0.2
        a = np.loadtxt(os.path.join(_DATA, "gmm16.txt"))      # frozen PRNGKey(0) constants
        return GaussianMixture(a[:, 0:2], a[:, 2:4], a[:, 4], device=device)
    if args.example == "phi-four":
        args.dim, args.lim, args.num_chain, args.eval_iter, args.step_size = 64, [-1.6, 1.6], 1024, 1, 0.0001
        return PhiFour(args.dim, device=device)
    if args.example == "4-mode":
        args.dim, args.lim, args.levels, args.step_size = 2, [-16, 16], 20, 0.2
        modes = 8.0 * np.array([[1, 1], [1, -1], [-1, 1], [-1, -1]])
        return GaussianMixture(modes, np.ones((4, 2)), np.ones(4) / 4, device=device)
    if args.example == "pines":
        args.dim, args.lim, args.num_chain, args.eval_iter, args.step_size = 1600, None, 128, 1, 0.01
        args.hidden_x = args.hidden_t = args.hidden_xt = [1024, 1024]
        return LogGaussianCoxPines(args.dim, device=device)
    raise Exception("Example not found.")


# (flag, type, default) of the reference's command line (multi_modal.py:149-210), kept name for name so its command lines run
_SCALARS = [
    ("seed", int, None), ("dim", int, 64), ("num_modes", int, 16), ("example", str, "pines"), ("sigma", float, 1e-4),
    ("fourier_dim", int, 128), ("fourier_std", float, 1.0), ("ref_dist", str, "stdgauss"), ("num_importance_samples", int, 0),
    ("mcmc_per_flow_steps", float, 10), ("num_chain", int, 128), ("learning_iter", int, 400), ("eval_iter", int, 100),
    ("alpha", float, 0.95), ("anneal_iter", int, 200), ("num_anneal_temp", int, 200), ("non_linearity", str, "relu"),
    ("step_size", float, 0.2), ("learning_rate", float, 1e-3), ("weight_decay", float, 1e-4), ("adam_beta1", float, 0.9),
    ("adam_beta2", float, 0.999), ("adam_epsilon", float, 1e-8), ("gradient_clip", float, 1.0), ("warmup_steps", int, 0),
    ("rtol", float, 1e-5), ("atol", float, 1e-5), ("mxstep", float, 1_000), ("log_every", int, 100),
]
_SWITCHES = {"hutchs": False, "cond_flow": True, "ot_cond_flow": False}          # store_true flags and their defaults
_LISTS = [("hidden_x", int, "+", [128, 128]), ("hidden_t", int, "+", [128, 128]), ("hidden_xt", int, "+", [128, 128]),
          ("lim", float, 2, [-16, 16])]


def parser():
    p = argparse.ArgumentParser(description="Markovian flow matching on the B200 hot path (reference CLI)")
    for name, typ, default in _SCALARS:
        p.add_argument(f"--{name}", type=typ, default=default)
    for name, default in _SWITCHES.items():
        p.add_argument(f"--{name}", dest=name, action="store_true")
    p.set_defaults(**_SWITCHES)
    for name, typ, nargs, default in _LISTS:
        p.add_argument(f"--{name}", type=typ, nargs=nargs, default=default)
    return p


def main(args):
    dist = build(args)
    seeds = [args.seed] if args.seed else [i ** 10 for i in range(10)]       # multi_modal.py:118 (seed 0 is falsy)
    out = []
    for seed in seeds:
        args.seed = seed
        res = run(dist, args, None, log_every=args.log_every)
        out.append(res)
        print(f"seed {seed}: train_time {res['train_time']:.2f}s final_beta {res['final_beta']:.6f} "
              f"last {res['history'][-1] if res['history'] else None}")
    return out


if __name__ == "__main__":
    main(parser().parse_args())
