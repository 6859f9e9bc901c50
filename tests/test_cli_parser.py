"""The CLI mirror accepts the reference's flags with the reference's defaults (multi_modal.py:149-210) - no GPU needed."""
import argparse


def test_parser_defaults_and_flags_match_the_reference_cli():
    from mfm_b200 import multi_modal as MM
    a = MM.parser().parse_args([])
    ref_defaults = dict(seed=None, dim=64, num_modes=16, example="pines", sigma=1e-4, fourier_dim=128, fourier_std=1.0, hutchs=False,
                        ref_dist="stdgauss", cond_flow=True, ot_cond_flow=False, num_importance_samples=0, mcmc_per_flow_steps=10,
                        num_chain=128, learning_iter=400, eval_iter=100, alpha=0.95, anneal_iter=200, num_anneal_temp=200,
                        non_linearity="relu", hidden_x=[128, 128], hidden_t=[128, 128], hidden_xt=[128, 128], step_size=0.2,
                        learning_rate=1e-3, weight_decay=1e-4, adam_beta1=0.9, adam_beta2=0.999, adam_epsilon=1e-8, gradient_clip=1.0,
                        warmup_steps=0, rtol=1e-5, atol=1e-5, mxstep=1000, lim=[-16, 16], log_every=100)
    assert vars(a) == ref_defaults
    # the four configurations of BASELINE.json as command lines
    for argv, checks in [
        (["--example", "4-mode", "--learning_iter", "1000", "--mcmc_per_flow_steps", "10"], dict(hutchs=False, mcmc_per_flow_steps=10.0)),
        (["--example", "gaussian-mixture", "--learning_iter", "10000", "--mcmc_per_flow_steps", "100", "--hutchs"], dict(hutchs=True)),
        (["--example", "phi-four", "--learning_iter", "10000", "--mcmc_per_flow_steps", "1000"], dict(learning_iter=10000)),
        (["--example", "pines", "--learning_iter", "10000", "--mcmc_per_flow_steps", "100", "--hutchs"], dict(example="pines", hutchs=True)),
    ]:
        ns = MM.parser().parse_args(argv)
        assert all(getattr(ns, k) == v for k, v in checks.items())
    assert isinstance(MM.parser(), argparse.ArgumentParser)


def test_per_example_overrides_need_no_device_until_build():
    from mfm_b200 import multi_modal as MM
    import pytest
    ns = MM.parser().parse_args(["--example", "nope"])
    with pytest.raises(Exception, match="Example not found"):
        MM.build(ns, device="cpu")
