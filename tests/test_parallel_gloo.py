"""N>1 host logic on CPU (gloo, world_size 2): shard ranges, gradient SUM all-reduce, chain gather,
and that the per-shard key rows equal the rows of the global split (what makes sharded runs
draw the single-GPU random numbers)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as tdist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n_total, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    tdist.init_process_group("gloo", rank=rank, world_size=world)
    from mfm_b200 import parallel
    lo, hi = parallel.shard_range(n_total, rank, world)
    # gradient exchange: each rank contributes the sum over its chains
    per_chain = torch.arange(n_total, dtype=torch.float32)[:, None] * torch.ones(1, 5)
    grads = per_chain[lo:hi].sum(0)
    loss = torch.tensor([float(hi - lo)])
    parallel.allreduce_sum_([grads, loss])
    ok = torch.allclose(grads, per_chain.sum(0)) and loss.item() == n_total
    # tempering gather: global chain order restored from ragged shards
    ll = torch.arange(lo, hi, dtype=torch.float32)
    full = parallel.allgather_chains(ll, n_total)
    ok = ok and torch.equal(full, torch.arange(n_total, dtype=torch.float32))
    ret[rank] = (bool(ok), lo, hi)
    tdist.destroy_process_group()


def test_gloo_world2_exchange():
    world, n_total = 2, 11
    mgr = mp.Manager(); ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n_total, ret), nprocs=world, join=True)
    assert all(ret[r][0] for r in range(world))
    assert [ret[r][1:] for r in range(world)] == [(0, 6), (6, 11)]


def test_shard_ranges_cover_everything():
    from mfm_b200 import parallel
    for n_total in (1, 7, 128, 65536):
        for world in (1, 2, 3, 8):
            r = [parallel.shard_range(n_total, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n_total
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))


def test_shard_key_rows_match_global_split(lib):
    """Row j of split(key, N) depends only on (key, j, N): a rank can derive its rows alone."""
    from mfm_b200 import parallel, random as mr
    from oracle import threefry as tf
    key = tf.PRNGKey(59049)
    n_total = 10
    full = tf.split(key, n_total)
    host = mr.host_split(key, n_total)
    for rank in range(3):
        lo, hi = parallel.shard_range(n_total, rank, 3)
        assert host[lo:hi].tolist() == full[lo:hi].tolist()
