"""GPU tests through the host mirror (Writer / Reader API of the reference), written after the
reference's own tests in src/tests/writer.rs and src/tests/reader.rs, plus oracle cross-checks."""
import numpy as np
import pytest

import arroy_b200 as ab
import oracle
from helpers import check_dump, golden

pytestmark = pytest.mark.gpu
G = golden()
SEED = bytes([42] * 32)


def rng42():
    return ab.StdRng.from_seed(SEED)


@pytest.fixture(scope="module")
def env_factory():
    envs = []
    shared = {"ctx": None}

    def make():
        e = ab.Env(0)
        if shared["ctx"] is None:
            shared["ctx"] = e.ctx
        else:
            e._ctx = shared["ctx"]
        envs.append(e)
        return e

    yield make
    for e in envs:
        e._ctx = None
    if shared["ctx"] is not None:
        shared["ctx"].close()


@pytest.mark.parametrize("line,n,dims,gen", [
    ("280", 4, 3, lambda i: [i, i, i]),     # write_vectors_until_there_is_a_split
    ("403", 6, 2, lambda i: [i, 0.0]),      # overwrite_one_item_incremental (first build)
    ("605", 3, 2, lambda i: [i, 0.0]),      # delete_one_leaf_in_a_split (first build)
])
def test_reference_inline_snapshots_through_the_writer(env_factory, line, n, dims, gen):
    gold = G["writer_inline"][line]
    env = env_factory()
    w = ab.Writer(env, 0, dims, "euclidean")
    for i in range(n):
        w.add_item(i, gen(float(i)))
    w.builder(rng42()).n_trees(1).build()
    r = ab.Reader.open(env, 0, "euclidean")
    check_dump(gold, env.tree_nodes(), r._roots(), oracle.EUCLIDEAN, dims, oracle.decode_node)
    # items dump: header bias 0.0000 + vector
    for key, it in gold["items"].items():
        assert ["%.4f" % v for v in w.item_vector(int(key))] == it["vector"]


def test_write_and_update_lot_of_random_points_first_snapshot(env_factory):
    gold = G["lot_of_random_points"]
    env = env_factory()
    w = ab.Writer(env, 0, 30, "euclidean")
    rng = rng42()
    for i in range(100):
        w.add_item(i, rng.fill_f32(30))
    w.builder(rng).n_trees(10).build()
    r = ab.Reader.open(env, 0, "euclidean")
    check_dump(gold, env.tree_nodes(), r._roots(), oracle.EUCLIDEAN, 30, oracle.decode_node)
    st = r.stats()
    assert st["leaf"] == 100 and len(st["tree_stats"]) == 10
    assert all(t["dummy_normals"] == 0 and t["descendants"] == t["split_nodes"] + 1 for t in st["tree_stats"])


def _line(env_factory, column=False):
    env = env_factory()
    w = ab.Writer(env, 0, 2, "euclidean")
    for i in range(100):
        w.add_item(i, [0.0, float(i)] if column else [float(i), 0.0])
    w.builder(rng42()).n_trees(50).build()
    return ab.Reader.open(env, 0, "euclidean")


def test_two_dimension_on_a_line(env_factory):  # src/tests/reader.rs:101-144
    r = _line(env_factory)
    q = G["reader_inline"]
    assert r.nns(5).search_k(1).by_item(1) == [tuple(x) for x in q["120"]]
    assert r.nns(5).search_k(2**64 - 1).by_item(0) == [tuple(x) for x in q["128"]]
    assert r.nns(5).by_item(0) == [tuple(x) for x in q["137"]]
    assert r.nns(5).by_item(1000) is None


def test_two_dimension_on_a_column_and_filtering(env_factory):  # src/tests/reader.rs:146-227
    r = _line(env_factory, column=True)
    q = G["reader_inline"]
    assert r.nns(5).by_item(0) == [tuple(x) for x in q["168"]]
    assert r.nns(5).candidates(range(0, 2)).by_item(0) == [tuple(x) for x in q["216"]]
    assert r.nns(5).candidates(range(98, 1000)).by_item(0) == [tuple(x) for x in q["223"]]
    assert r.item_ids() == list(range(100))


def test_search_in_db_with_a_single_vector(env_factory):  # src/tests/reader.rs:81-99
    env = env_factory()
    w = ab.Writer(env, 0, 3, "cosine")
    w.add_item(0, [0.00397, 0.553, 0.0])
    w.builder(rng42()).build()
    r = ab.Reader.open(env, 0, "cosine")
    assert r.nns(1).by_item(0) == [tuple(x) for x in G["reader_inline"]["96"]]


@pytest.mark.parametrize("metric,n,d,trees,centre", [
    ("euclidean", 10000, 64, 10, 0.0),    # config C1 (examples/compare_with_hnsw.rs shape)
    ("cosine", 6000, 128, 8, 0.5),
    ("dot-product", 6000, 96, 8, 0.5),
    ("manhattan", 3000, 48, 4, 0.5),
])
def test_forest_and_queries_match_the_oracle_end_to_end(env_factory, metric, n, d, trees, centre):
    data = oracle.synth_rows(SEED, d, 0, n, centre, threads=4)
    ids = np.arange(n, dtype=np.uint32)
    odb = oracle.Db(metric, d)
    odb.set_items(ids, data)
    odb.build(oracle.StdRng(SEED), n_trees=trees, threads=8)
    env = env_factory()
    w = ab.Writer(env, 0, d, metric)
    w.add_items(ids, data)
    w.builder(rng42()).n_trees(trees).build()
    assert env.tree_nodes() == odb.nodes()
    r = ab.Reader.open(env, 0, metric)
    assert r.n_trees() == trees and r.n_items() == n
    qitems = list(range(0, 200, 7))
    for k, search_k in ((5, None), (100, None), (10, 2000)):
        for it in qitems:
            want = odb.nns_by_item(it, k, search_k=search_k)
            got = r.nns(k).search_k(search_k).by_item(it) if search_k else r.nns(k).by_item(it)
            assert [g[0] for g in got] == [x[0] for x in want], (it, k)
            assert np.array([g[1] for g in got], dtype=np.float32).tobytes() == np.array([x[1] for x in want], dtype=np.float32).tobytes()
    # by_vector with a vector that is not in the index
    qv = oracle.synth_rows(SEED, d, n + 3, 1, centre)[0]
    assert r.nns(20).by_vector(qv) == odb.nns_by_vector(qv, 20)
    # batched by_item (device tree walk + re-rank) == one at a time (host walk + device re-rank)
    for k, sk in ((10, None), (100, None), (7, 3000)):
        out_ids, out_dist, out_len, _ = r.nns_batch_by_item(qitems, k, search_k=sk)
        for i, it in enumerate(qitems):
            single = r.nns(k).search_k(sk).by_item(it) if sk else r.nns(k).by_item(it)
            assert out_ids[i, :out_len[i]].tolist() == [s[0] for s in single], (k, sk, it)
            assert out_dist[i, :out_len[i]].tolist() == [s[1] for s in single]
    # and the host-walk batch path gives the same
    import os
    os.environ["ARROY_B200_HOST_WALK"] = "1"
    try:
        h_ids, h_dist, h_len, _ = r.nns_batch_by_item(qitems, 10)
    finally:
        os.environ.pop("ARROY_B200_HOST_WALK")
    d_ids, d_dist, d_len, _ = r.nns_batch_by_item(qitems, 10)
    assert h_len.tolist() == d_len.tolist() and h_ids.tolist() == d_ids.tolist() and h_dist.tobytes() == d_dist.tobytes()


def _oracle_db(metric, d, ids, data, trees):
    odb = oracle.Db(metric, d)
    odb.set_items(ids, data)
    odb.build(oracle.StdRng(SEED), n_trees=trees, threads=4)
    return odb


def test_two_indexes_and_a_rebuild_share_one_device_context(env_factory):
    # the device context is shared by every Reader / Writer of an Env. A reader must never walk or
    # re-rank another owner's resident items / forest.
    env = env_factory()
    d0, n0, d1, n1 = 48, 3000, 64, 1200
    a = oracle.synth_rows(SEED, d0, 0, n0, 0.5)
    b = oracle.synth_rows(bytes([7] * 32), d1, 0, n1, 0.5)
    ids0, ids1 = np.arange(n0, dtype=np.uint32), np.arange(10, 10 + n1, dtype=np.uint32)
    w0, w1 = ab.Writer(env, 0, d0, "euclidean"), ab.Writer(env, 1, d1, "cosine")
    w0.add_items(ids0, a)
    w1.add_items(ids1, b)
    w0.builder(rng42()).n_trees(4).build()
    w1.builder(rng42()).n_trees(3).build()
    o0, o1 = _oracle_db("euclidean", d0, ids0, a, 4), _oracle_db("cosine", d1, ids1, b, 3)
    r0, r1 = ab.Reader.open(env, 0, "euclidean"), ab.Reader.open(env, 1, "cosine")
    q0, q1 = [0, 5, 77, 2999], [10, 11, 500, 1209]

    def check(r, o, qs):
        out_ids, out_dist, out_len, _ = r.nns_batch_by_item(qs, 10)
        for i, it in enumerate(qs):
            want = o.nns_by_item(it, 10)
            assert out_ids[i, :out_len[i]].tolist() == [x[0] for x in want]
            assert out_dist[i, :out_len[i]].tobytes() == np.array([x[1] for x in want], dtype=np.float32).tobytes()
            assert r.nns(10).by_item(it) == want

    for _ in range(2):          # alternate: every switch finds the other index resident on the device
        check(r0, o0, q0)
        check(r1, o1, q1)
    # a third index is built on the same context while the readers stay open: its staging replaces theirs
    w2 = ab.Writer(env, 2, d0, "euclidean")
    w2.add_items(ids0[:500], a[:500])
    w2.builder(rng42()).n_trees(2).build()
    check(r1, o1, q1)
    check(r0, o0, q0)
    # a rebuild of index 0 makes its open reader stale (NeedBuild) instead of silently wrong
    w0.add_item(n0, a[0] * 0.5)
    w0.builder(rng42()).n_trees(4).build()
    with pytest.raises(ab.ArroyError) as ei:
        r0.nns_batch_by_item(q0, 10)
    assert ei.value.kind == "NeedBuild"
    check(r1, o1, q1)


def test_cancel_in_the_middle_of_a_device_build_rolls_back(env_factory):
    # BuildCancelled must leave the table as it was (the reference's RwTxn is dropped)
    env = env_factory()
    d, n = 32, 4000
    data = oracle.synth_rows(SEED, d, 0, n, 0.5)
    w = ab.Writer(env, 0, d, "dot-product")
    w.add_items(np.arange(n - 500, dtype=np.uint32), data[:n - 500])
    w.builder(rng42()).n_trees(3).build()
    w.add_items(np.arange(n - 500, n, dtype=np.uint32), data[n - 500:])
    w.del_item(3)
    before = env.items()
    for after in (1, 2, 4):      # cancel at different polls: before staging, between the host phases, inside the device loop
        calls = {"n": 0}

        def cancel():
            calls["n"] += 1
            return calls["n"] > after

        with pytest.raises(ab.ArroyError) as ei:
            w.builder(rng42()).n_trees(3).cancel(cancel).build()
        assert ei.value.kind == "BuildCancelled"
        assert env.items() == before and w.need_build()
    w.builder(rng42()).n_trees(3).build()
    assert not w.need_build() and env.items() != before


def test_little_memory_golden_through_the_writer(env_factory):
    # src/tests/writer.rs:1377-1391 (write_and_update_lot_of_random_points_with_little_memory): 100 x 3 Cosine, available_memory(0),
    # 2 trees — every tree is built from sampled chunks + routing + re-spawned tasks; all 188 nodes of the reference's snapshot
    gold = G["little_memory"]
    env = env_factory()
    w = ab.Writer(env, 0, 3, "cosine")
    rng = rng42()
    for i in range(100):
        w.add_item(i, rng.fill_f32(3))
    w.builder(rng).available_memory(0).n_trees(2).build()
    r = ab.Reader.open(env, 0, "cosine")
    assert len(env.tree_nodes()) == 188
    check_dump(gold, env.tree_nodes(), r._roots(), oracle.COSINE, 3, oracle.decode_node)


@pytest.mark.parametrize("metric,n,d,trees,pages", [
    ("euclidean", 3000, 16, 3, 8),      # 60 items per page -> chunks of 480 items
    ("cosine", 2500, 40, 2, 20),        # 24 per page -> 480
    ("dot-product", 2000, 24, 2, 6),    # 39 per page -> 234
    ("manhattan", 1500, 1200, 2, 300),  # an item spans 2 pages -> 150 items -> raised to d + 1
])
def test_memory_limited_build_matches_the_oracle(env_factory, metric, n, d, trees, pages):
    data = oracle.synth_rows(SEED, d, 0, n, 0.5, threads=4)
    ids = np.arange(0, 2 * n, 2, dtype=np.uint32)   # sparse ids: rows != ids
    odb = oracle.Db(metric, d)
    odb.set_items(ids, data)
    odb.build_memory_limited(oracle.StdRng(SEED), n_trees=trees, available_memory=pages * 4096)
    env = env_factory()
    w = ab.Writer(env, 0, d, metric)
    w.add_items(ids, data)
    w.builder(rng42()).available_memory(pages * 4096).n_trees(trees).build()
    assert env.tree_nodes() == odb.nodes()
    assert ab.Reader.open(env, 0, metric)._roots() == odb.roots
    # with enough memory the same call is the plain build
    env2 = env_factory()
    w2 = ab.Writer(env2, 0, d, metric)
    w2.add_items(ids, data)
    w2.builder(rng42()).available_memory(1 << 40).n_trees(trees).build()
    full = oracle.Db(metric, d)
    full.set_items(ids, data)
    full.build(oracle.StdRng(SEED), n_trees=trees, threads=4)
    assert env2.tree_nodes() == full.nodes()
