"""Incremental Writer::build through the host mirror (device side() routing + device subtree
builds) against the reference's own add/delete scenarios and against the oracle."""
import numpy as np
import pytest

import arroy_b200 as ab
import oracle
from helpers import check_dump, golden
from test_oracle_golden import SCENARIOS

pytestmark = pytest.mark.gpu
G = golden()
SEED = bytes([42] * 32)


@pytest.fixture(scope="module")
def shared_ctx():
    c = ab.Context(0)
    yield c
    c.close()


def make_env(ctx):
    e = ab.Env(0)
    e._ctx = ctx
    return e


@pytest.mark.parametrize("name", sorted(k for k in SCENARIOS if k != "delete_extraneous_tree"))
def test_reference_incremental_scenarios(shared_ctx, name):
    dims, steps = SCENARIOS[name]
    env = make_env(shared_ctx)
    w = ab.Writer(env, 0, dims, "euclidean")
    rng = ab.StdRng.from_seed(SEED)
    for st in steps:
        if st[0] == "add":
            w.add_item(st[1], st[2])
        elif st[0] == "del":
            w.del_item(st[1])
        else:
            b = w.builder(rng)
            if st[1] is not None:
                b.n_trees(st[1])
            if len(st) > 3:
                b.split_after(st[3])
            b.build()
            gold = G["writer_inline"][st[2]]
            r = ab.Reader.open(env, 0, "euclidean")
            check_dump(gold, env.tree_nodes(), r._roots(), oracle.EUCLIDEAN, dims, oracle.decode_node)
    env._ctx = None


def test_lot_of_random_points_then_update_second_snapshot(shared_ctx):
    env = make_env(shared_ctx)
    w = ab.Writer(env, 0, 30, "euclidean")
    rng = ab.StdRng.from_seed(SEED)
    for i in range(100):
        w.add_item(i, rng.fill_f32(30))
    w.builder(rng).n_trees(10).build()
    for i in range(0, 100, 2):
        w.add_item(i, rng.fill_f32(30))
    w.builder(rng).n_trees(10).build()
    r = ab.Reader.open(env, 0, "euclidean")
    check_dump(G["lot_of_random_points_2"], env.tree_nodes(), r._roots(), oracle.EUCLIDEAN, 30, oracle.decode_node)
    env._ctx = None


@pytest.mark.parametrize("metric,d,n_trees", [("cosine", 32, 4), ("euclidean", 48, 3), ("dot-product", 40, 3), ("manhattan", 33, 2)])
def test_random_update_sequences_match_the_oracle(shared_ctx, metric, d, n_trees):
    n0 = 3000
    data = oracle.synth_rows(SEED, d, 0, n0 + 2000, 0.5, threads=4)
    env = make_env(shared_ctx)
    w = ab.Writer(env, 0, d, metric)
    odb = oracle.Db(metric, d)
    prng, orng = ab.StdRng.from_seed(SEED), oracle.StdRng(SEED)
    r = np.random.default_rng(5)
    for i in range(n0):
        w.add_item(i, data[i])
        odb.add_item(i, data[i])
    live = set(range(n0))
    next_new = n0
    for rnd in range(4):
        w.builder(prng).n_trees(n_trees).build()
        odb.build_incremental(orng, n_trees=n_trees)
        assert env.tree_nodes() == odb.nodes(), (metric, rnd)
        # mutate: delete some, overwrite some, add some
        for i in r.choice(sorted(live), size=150, replace=False):
            w.del_item(int(i)); odb.del_item(int(i)); live.discard(int(i))
        for i in r.choice(sorted(live), size=100, replace=False):
            v = data[int(r.integers(0, len(data)))] * np.float32(0.5)
            w.add_item(int(i), v); odb.add_item(int(i), v)
        for _ in range(int(r.integers(50, 700))):
            w.add_item(next_new, data[next_new % len(data)]); odb.add_item(next_new, data[next_new % len(data)])
            live.add(next_new); next_new += 1
    # growing the forest on an existing index
    w.builder(prng).n_trees(n_trees + 2).build()
    odb.build_incremental(orng, n_trees=n_trees + 2)
    assert env.tree_nodes() == odb.nodes()
    # and the updated index answers queries like the oracle's
    rd = ab.Reader.open(env, 0, metric)
    for it in sorted(live)[:40]:
        want = odb.nns_by_item(it, 10)
        got = rd.nns(10).by_item(it)
        assert [g[0] for g in got] == [x[0] for x in want]
    env._ctx = None


def test_incremental_on_degenerate_data_uses_random_sides_like_the_oracle(shared_ctx):
    # identical vectors => "normal: none" nodes; routing new items through them draws Side::random
    # from R::seed_from_u64(seed + root) in depth-first order (writer.rs:1128-1133, :1419-1421)
    n, d = 400, 32
    env = make_env(shared_ctx)
    w = ab.Writer(env, 0, d, "euclidean")
    odb = oracle.Db("euclidean", d)
    prng, orng = ab.StdRng.from_seed(SEED), oracle.StdRng(SEED)
    one = np.ones(d, dtype=np.float32)
    for i in range(n):
        w.add_item(i, one); odb.add_item(i, one)
    w.builder(prng).n_trees(3).split_after(40).build()
    odb.build_incremental(orng, n_trees=3, split_after=40)
    assert env.tree_nodes() == odb.nodes()
    for i in range(n, n + 120):
        w.add_item(i, one); odb.add_item(i, one)
    w.del_item(7); odb.del_item(7)
    w.builder(prng).n_trees(3).split_after(40).build()
    odb.build_incremental(orng, n_trees=3, split_after=40)
    assert env.tree_nodes() == odb.nodes()
    env._ctx = None
