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


def parser():
    p = argparse.ArgumentParser()
    p.add_argument("--seed", type=int, default=None)
    p.add_argument("--dim", type=int, default=64)
    p.add_argument("--num_modes", type=int, default=16)
    p.add_argument("--example", type=str, default="pines")
    p.add_argument("--sigma", type=float, default=1e-4)
    p.add_argument("--fourier_dim", type=int, default=128)
    p.add_argument("--fourier_std", type=float, default=1.0)
    p.add_argument("--hutchs", dest="hutchs", action="store_true")
    p.set_defaults(hutchs=False)
    p.add_argument("--ref_dist", type=str, default="stdgauss")
    p.add_argument("--cond_flow", dest="cond_flow", action="store_true")
    p.set_defaults(cond_flow=True)
    p.add_argument("--ot_cond_flow", dest="ot_cond_flow", action="store_true")
    p.set_defaults(ot_cond_flow=False)
    p.add_argument("--num_importance_samples", type=int, default=0)
    p.add_argument("--mcmc_per_flow_steps", type=float, default=10)
    p.add_argument("--num_chain", type=int, default=128)
    p.add_argument("--learning_iter", type=int, default=400)
    p.add_argument("--eval_iter", type=int, default=100)
    p.add_argument("--alpha", type=float, default=0.95)
    p.add_argument("--anneal_iter", type=int, default=200)
    p.add_argument("--num_anneal_temp", type=int, default=200)
    p.add_argument("--non_linearity", type=str, default="relu")
    p.add_argument("--hidden_x", type=int, nargs="+", default=[128, 128])
    p.add_argument("--hidden_t", type=int, nargs="+", default=[128, 128])
    p.add_argument("--hidden_xt", type=int, nargs="+", default=[128, 128])
    p.add_argument("--step_size", type=float, default=0.2)
    p.add_argument("--learning_rate", type=float, default=1e-3)
    p.add_argument("--weight_decay", type=float, default=0.0001)
    p.add_argument("--adam_beta1", type=float, default=0.9)
    p.add_argument("--adam_beta2", type=float, default=0.999)
    p.add_argument("--adam_epsilon", type=float, default=1e-8)
    p.add_argument("--gradient_clip", type=float, default=1.0)
    p.add_argument("--warmup_steps", type=int, default=0)
    p.add_argument("--rtol", type=float, default=1e-5)
    p.add_argument("--atol", type=float, default=1e-5)
    p.add_argument("--mxstep", type=float, default=1_000)
    p.add_argument("--lim", type=float, nargs=2, default=[-16, 16])
    p.add_argument("--log_every", type=int, default=100)
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
