"""KSD / MMD metrics (mcmc_utils.py mirror) on the device vs the float64 oracle."""
import numpy as np
import pytest
import torch

from oracle import metrics as OM, targets as OT
from tests.helpers import make_targets, to_dev

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,T", [("4-mode", 1000), ("gmm16", 333), ("phi-four", 257), ("pines", 48)])
def test_stein_disc_vs_oracle(cuda, lib, name, T):
    from mfm_b200 import mcmc_utils as MU
    ot, dd = next((o, d) for n, o, d in make_targets(cuda) if n == name)
    rng = np.random.default_rng(T)
    if name == "4-mode":
        X = ot.modes[rng.integers(0, 4, T)] + rng.standard_normal((T, 2))
    elif name == "gmm16":
        X = ot.modes[rng.integers(0, 16, T)] + 0.5 * rng.standard_normal((T, 2))
    elif name == "phi-four":
        X = rng.uniform(-1, 1, (T, 64))
    else:
        X = ot.mu + 0.3 * rng.standard_normal((T, 1600))
    X = X.astype(np.float32)
    u_ref, v_ref = OM.stein_disc(X.astype(np.float64), ot.grad)
    for fn in (dd, dd.logprob, dd.tempered(1.0)):
        u, v = MU.stein_disc(to_dev(X, cuda), fn)
        # float32 pair values (scores are ~1e2..1e4 for phi-four / pines), float64 sums: relative to the V-statistic
        scale = max(abs(v_ref), abs(u_ref))
        assert abs(u.item() - u_ref) <= 2e-4 * scale and abs(v.item() - v_ref) <= 2e-4 * scale, (u.item(), u_ref, v.item(), v_ref)
    with pytest.raises(TypeError):
        MU.stein_disc(to_dev(X, cuda), lambda x: -(x * x).sum())


@pytest.mark.parametrize("m,d", [(500, 2), (129, 64), (40, 1600)])
def test_max_mean_disc_vs_oracle(cuda, lib, m, d):
    from mfm_b200 import mcmc_utils as MU
    rng = np.random.default_rng(m + d)
    s = 1.0 / np.sqrt(d)
    X = (rng.standard_normal((m, d)) * s).astype(np.float32)
    Y = (rng.standard_normal((m, d)) * s + 0.5 * s).astype(np.float32)
    ref = OM.max_mean_disc(X, Y)
    got = MU.max_mean_disc(to_dev(X, cuda), to_dev(Y, cuda)).item()
    assert abs(got - ref) <= 1e-5 * max(1.0, abs(ref)) + 2e-4 * abs(ref), (got, ref)
    same = MU.max_mean_disc(to_dev(X, cuda), to_dev(X, cuda)).item()
    assert abs(same - OM.max_mean_disc(X, X)) <= 1e-5


def test_run_reports_metric_table(cuda, lib):
    """run() ends like the reference: flow samples, resampled samples, logpdf / KSD table (+ MMD with a target generator)."""
    from mfm_b200 import multi_modal as MM, random as mr
    args = MM.parser().parse_args(["--example", "4-mode", "--learning_iter", "4", "--mcmc_per_flow_steps", "2", "--seed", "1",
                                   "--eval_iter", "3"])
    dist = MM.build(args, device=cuda)
    modes = torch.tensor(8.0 * np.array([[1, 1], [1, -1], [-1, 1], [-1, -1]]), dtype=torch.float32, device=cuda)

    def target_gn(keys):                       # 4-mode sampler from per-sample keys (multi_modal.py:70-76 analogue)
        ks = mr.split(keys, 2)
        comp = (mr.uniform(ks[:, 0].contiguous(), (1,))[:, 0] * 4).long().clamp(max=3)
        return modes[comp] + mr.normal(ks[:, 1].contiguous(), (2,))

    res = MM.run(dist, args, target_gn)
    n = 3 * 128
    assert res["flow_samples"].shape == (n, 2) and res["exact_samples"].shape == (n, 2)
    t = res["table"]
    assert list(t)[:9] == ["mcmc/flow", "learn iter", "train time", "logpdf", "logpdf*", "KSD U-stat", "KSD U-stat*", "KSD V-stat", "KSD V-stat*"]
    assert "MMD" in t and "MMD*" in t and all(np.isfinite(v) for v in t.values())
    # importance resampling moves the untrained flow's samples towards the target
    assert t["logpdf*"] >= t["logpdf"]
