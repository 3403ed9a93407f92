"""Sharded (begin / emit) build equals the monolithic build; with >= 2 GPUs also across devices."""
import numpy as np
import pytest

import arroy_b200 as ab
import oracle
from arroy_b200 import parallel

pytestmark = pytest.mark.gpu
SEED = bytes([42] * 32)


def _seeds(n_trees):
    user = oracle.StdRng(SEED)
    r1 = oracle.StdRng(user.gen_seed())
    return [r1.gen_seed() for _ in range(n_trees)]


def test_two_phase_sharded_build_matches_monolithic():
    import torch
    n, d, T = 6000, 64, 7
    data = oracle.synth_rows(SEED, d, 0, n, 0.5)
    ids = np.arange(n, dtype=np.uint32)
    seeds = _seeds(T)
    ctx = ab.Context(0)
    ctx.stage_items_flat("cosine", ids, data)
    want = ctx.build_trees(seeds, list(range(T)), T)
    n_dev = torch.cuda.device_count()
    world = 2
    ctxs = [ctx, ab.Context(1) if n_dev >= 2 else ab.Context(0)]
    ctxs[1].stage_items_flat("cosine", ids, data)
    counts_all = np.zeros(T, dtype=np.int64)
    for r in range(world):  # phase 1 on every "rank" (its own context, its own device when there are two)
        mine = parallel.shard_trees(T, r, world)
        counts_all[mine] = ctxs[r].build_trees_begin([seeds[t] for t in mine])
    base = parallel.node_id_bases(counts_all, T)   # what the ranks derive after the all-gather
    got = {}
    for r in range(world):  # phase 2
        mine = parallel.shard_trees(T, r, world)
        got.update(ctxs[r].build_trees_emit([t for t in mine], [base[t] for t in mine]))
    assert got.keys() == want.keys()
    assert all(got[k] == want[k] for k in want)
    odb = oracle.Db("cosine", d)
    odb.set_items(ids, data)
    odb.build(oracle.StdRng(SEED), n_trees=T, threads=4)
    assert odb.nodes() == got
