"""Order of operations of HotLoop.iteration (no GPU: the device layer is replaced by recorders).

Reference loop (exe_flow_matching.py:433-439): generator, then train_step, every iteration.  With several ranks the AdamW
update of iteration k is applied AFTER iteration k+1's MALA step (which never reads the MLP) but BEFORE a flow-MH step."""
from types import SimpleNamespace

import pytest
import torch


def _loop(monkeypatch, m, pipeline, real_sampler=None):
    from mfm_b200 import exe_flow_matching as E
    log = []
    loop = object.__new__(E.HotLoop)
    loop.args = SimpleNamespace(mcmc_per_flow_steps=m)
    loop.real_sampler, loop.graph, loop.pipeline, loop.beta = real_sampler, False, pipeline, 1.0
    loop.n = loop.n_total = 4
    loop.chain_offset, loop.pg, loop.count, loop.P = 0, None, 0, "P"
    loop.key_sample = torch.zeros(2, dtype=torch.int64)
    loop.states = E.MALAState(torch.zeros(4, 2), torch.zeros(4), torch.zeros(4, 2))
    monkeypatch.setattr(E.mrandom, "split", lambda key, num=2: torch.zeros((num, 2), dtype=torch.int64))

    def gen(key, states, count, P, beta, inplace=False, force=None):
        kind = force if force is not None else ("flow" if loop.is_flow_iteration(count) else "mala")
        log.append((kind, count))
        return states, SimpleNamespace(acceptance_rate=torch.zeros(4))

    pending = []

    class State:
        def loss_and_grad(self, key, pos, off, n_total, group=None, allreduce=False, defer=False):
            log.append(("grad", loop.count, "deferred" if defer else "now"))
            if defer:
                pending.append(loop.count)
            return torch.zeros(1), None

        def apply_gradients(self):
            log.append(("adamw", loop.count))

        def apply_pending(self):
            if pending:
                log.append(("adamw", pending.pop()))

    loop.gen, loop.state = gen, State()
    return loop, log


def test_unpipelined_loop_is_the_reference_order(monkeypatch):
    loop, log = _loop(monkeypatch, m=2, pipeline=False)
    for _ in range(4):
        loop.iteration()
    assert log == [("mala", 1), ("grad", 1, "now"), ("adamw", 1), ("mala", 2), ("grad", 2, "now"), ("adamw", 2),
                   ("flow", 3), ("grad", 3, "now"), ("adamw", 3), ("mala", 4), ("grad", 4, "now"), ("adamw", 4)]


def test_pipelined_update_lands_after_mala_but_before_flow(monkeypatch):
    loop, log = _loop(monkeypatch, m=2, pipeline=True)
    for _ in range(4):
        loop.iteration()
    loop.flush()
    assert log == [("mala", 1), ("grad", 1, "deferred"),
                   ("mala", 2), ("adamw", 1), ("grad", 2, "deferred"),           # update 1 under MALA step 2
                   ("adamw", 2), ("flow", 3), ("grad", 3, "deferred"),           # the flow step sees update 2
                   ("mala", 4), ("adamw", 3), ("grad", 4, "deferred"),
                   ("adamw", 4)]                                                 # flush()
    # every gradient is computed with all earlier updates applied
    applied = 0
    for e in log:
        if e[0] == "adamw":
            assert e[1] == applied + 1
            applied = e[1]
        if e[0] == "grad":
            assert applied == e[1] - 1


@pytest.mark.parametrize("m,flows", [(2, [3, 6, 9]), (10, [11]), (0.5, [1, 2, 4, 5, 7, 8, 10, 11])])
def test_flow_iteration_dispatch(monkeypatch, m, flows):
    """count % (int(m)+1) == 0 -> flow step; fractional m inverts the roles (exe_flow_matching.py:304-311)."""
    loop, _ = _loop(monkeypatch, m=m, pipeline=False)
    assert [c for c in range(1, 12) if loop.is_flow_iteration(c)] == flows


def test_real_samples_replace_the_generator(monkeypatch):
    """mcmc_per_flow_steps < 0 (:328,382-386): positions come from the target generator, no MALA / flow step, no tempering."""
    drawn = []

    def sampler(keys):
        drawn.append(keys.shape)
        return torch.ones(keys.shape[0], 2)

    loop, log = _loop(monkeypatch, m=-1, pipeline=False, real_sampler=sampler)
    loop.iteration(); loop.iteration()
    assert drawn == [(4, 2), (4, 2)] and [e[0] for e in log] == ["grad", "adamw", "grad", "adamw"]
    assert loop.states.logdensity is None and torch.equal(loop.states.position, torch.ones(4, 2))
    assert loop.temper() == 1.0 and not loop.is_flow_iteration(3)


def test_graph_path_runs_mala_for_fractional_m(monkeypatch):
    """m = 0.5 inverts the roles (flow unless count % 3 == 0, exe_flow_matching.py:304-309): counts 1 and 2 are BOTH flow
    iterations, so the captured MALA iteration cannot be selected through a fake count - it is forced explicitly.  The
    graph machinery itself needs a GPU; here the eager first pass of the graph path is what is checked."""
    from mfm_b200 import exe_flow_matching as E
    loop, log = _loop(monkeypatch, m=0.5, pipeline=False)
    loop.graph, loop._eager_done, loop._graph = True, False, None
    loop.key_sample = torch.zeros(2, dtype=torch.int64)
    loop.state.step = 0
    for _ in range(3):
        loop.iteration()
    # counts 1, 2: flow (eager path); count 3: MALA through the graph path's eager first pass
    assert [e for e in log if e[0] in ("mala", "flow")] == [("flow", 1), ("flow", 2), ("mala", 3)]
    assert [e[0] for e in log] == ["flow", "grad", "adamw"] * 2 + ["mala", "grad", "adamw"]
