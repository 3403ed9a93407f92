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


def _leaf_blob(metric, data, h0, h1=None):
    n, d = data.shape
    hf = 2 if metric == "dot-product" else 1
    stride = 1 + 4 * hf + 4 * d
    blob = np.zeros(n * stride, dtype=np.uint8)
    b2 = blob.reshape(n, stride)
    b2[:, 1:5] = np.ascontiguousarray(h0, dtype=np.float32).view(np.uint8).reshape(n, 4)
    if hf == 2:
        b2[:, 5:9] = np.ascontiguousarray(h1, dtype=np.float32).view(np.uint8).reshape(n, 4)
    b2[:, 1 + 4 * hf:] = np.ascontiguousarray(data).view(np.uint8).reshape(n, 4 * d)
    ptrs = (blob.ctypes.data + np.arange(n, dtype=np.uint64) * np.uint64(stride)).astype(np.uint64)
    return blob, ptrs


@pytest.mark.parametrize("metric,n,d,T", [("cosine", 150_000, 96, 7), ("dot-product", 20_000, 64, 5)])
def test_group_api_stages_broadcasts_and_builds_like_one_device(metric, n, d, T):
    # arroy_b200_create_group / _group_stage_items / _group_build_trees: chunked H2D on the first device
    # pipelined with ncclBroadcast, trees sharded t mod n_dev — node bytes identical to a single-device build and to the oracle.
    # With one visible GPU the group has one member (same code path minus NCCL); with >= 2 it spans two devices.
    import torch
    n_dev = min(2, torch.cuda.device_count())
    data = oracle.synth_rows(SEED, d, 0, n, 0.5, threads=8)
    ids = np.arange(n, dtype=np.uint32)
    seeds = _seeds(T)
    m = oracle.METRICS[metric]
    h0 = np.array([oracle.new_header(m, v)[0] for v in data[:2000]], dtype=np.float32)
    one = ab.Context(0)
    one.stage_items_flat(metric, ids, data)
    hh0, hh1 = one.item_headers()                      # headers as Writer::add_item stores them
    assert hh0[:2000].tobytes() == h0.tobytes()
    blob, ptrs = _leaf_blob(metric, data, hh0, hh1)
    if metric == "dot-product":
        one.dot_preprocess()
    want = one.build_trees(seeds, list(range(T)), T)
    g = ab.Group(list(range(n_dev)))
    assert g.size() == n_dev
    g.stage_items_ptrs(metric, d, ids, ptrs)
    if metric == "dot-product":
        extra, norm = g.dot_preprocess()
        w_extra, w_norm = oracle.dot_preprocess(data)
        assert extra.tobytes() == w_extra.tobytes() and norm.tobytes() == w_norm.tobytes()
    for r in range(n_dev):                             # every member holds the same items and headers
        gh0, gh1 = g.ctx(r).item_headers()
        oh0, oh1 = one.item_headers()
        assert gh0.tobytes() == oh0.tobytes() and gh1.tobytes() == oh1.tobytes()
        side_a, mg_a = g.ctx(r).side_batch(data[3], (0.0, 0.0), np.arange(0, n, 97, dtype=np.uint32))
        side_b, mg_b = one.side_batch(data[3], (0.0, 0.0), np.arange(0, n, 97, dtype=np.uint32))
        assert mg_a.tobytes() == mg_b.tobytes() and side_a.tobytes() == side_b.tobytes()
    got = g.build_trees(seeds, list(range(T)), T)
    assert got.keys() == want.keys() and all(got[k] == want[k] for k in want)
    arena = ab.Arena()
    assert g.build_trees(seeds, list(range(T)), T, arena=arena) == len(want)
    assert arena.stats()[0] == len(want) and all(arena.get(k) == want[k] for k in list(want)[:50])
    bd = g.stage_breakdown()
    assert bd["total_ms"] > 0
    g.close()
    one.close()
    del blob
