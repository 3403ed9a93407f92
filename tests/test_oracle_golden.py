"""Pins the CPU oracle to every golden vector the reference's tests hold for this path
(SURVEY.md §8c). Runs without a GPU."""
import numpy as np
import pytest

import oracle
from helpers import check_dump, fmt4, golden

G = golden()


def rng42():
    return oracle.StdRng(bytes([42] * 32))  # src/tests/mod.rs:105-107


def test_chacha12_f32_stream_matches_snapshot_items():
    # golden 1: items of write_and_update_lot_of_random_points (100 x 30, rng.gen::<f32>())
    rng = rng42()
    data = rng.fill_f32(100 * 30).reshape(100, 30)
    items = G["lot_of_random_points"]["items"]
    assert len(items) == 100
    for key, it in items.items():
        assert [fmt4(v) for v in data[int(key), :10]] == it["vector"]
    # full precision: src/tests/upgrade.rs:117 (item 25 == draws 750..779); Rust prints the
    # shortest round-trip decimal, so parsing it back must give the identical f32
    want = np.array([np.float32(s) for s in G["upgrade_item25"]], dtype=np.float32)
    assert want.tobytes() == data[25].tobytes()


def test_build_takes_exactly_one_seed_from_the_user_rng():
    # write_and_update_lot_of_random_points, second half (tests/writer.rs:310-316): after the
    # first build the same rng redraws the even ids. Their printed values pin that a fresh
    # Writer::build consumes exactly one `rng.gen::<[u8;32]>()` (32 words, writer.rs:575).
    rng = rng42()
    data = rng.fill_f32(100 * 30).reshape(100, 30).copy()
    rng.gen_seed()
    for i in range(0, 100, 2):
        data[i] = rng.fill_f32(30)
    for key, it in G["lot_of_random_points_2_items"].items():
        assert [fmt4(v) for v in data[int(key), :10]] == it["vector"], key


def _build(metric, dims, vectors, n_trees, rng=None, split_after=None, ids=None):
    db = oracle.Db(metric, dims)
    for i, v in enumerate(vectors):
        db.add_item(i if ids is None else ids[i], v)
    db.build(rng or rng42(), n_trees=n_trees, split_after=split_after)
    return db


@pytest.mark.parametrize("line,n,dims,gen", [
    ("210", 1, 3, lambda i: [0.0, 1.0, 2.0]),          # write_one_vector_in_one_tree
    ("254", 3, 3, lambda i: [i, i, i]),                # write_vectors_until_there_is_a_descendants
    ("280", 4, 3, lambda i: [i, i, i]),                # golden 2
    ("403", 6, 2, lambda i: [i, 0.0]),                 # golden 3
    ("605", 3, 2, lambda i: [i, 0.0]),                 # golden 4
    ("689", 6, 2, lambda i: [i, 0.0]),
    ("880", 6, 2, lambda i: [i, 0.0]),
    ("1057", 6, 2, lambda i: [i, 0.0]),
    ("1135", 6, 2, lambda i: [i, 0.0]),
    ("563", 2, 2, lambda i: [i, 0.0]),                 # delete_one_item_in_a_descendant (first build)
])
def test_inline_writer_snapshots(line, n, dims, gen):
    gold = G["writer_inline"][line]
    assert gold["dimensions"] == dims and gold["distance"] == "euclidean"
    db = _build("euclidean", dims, [gen(float(i)) for i in range(n)], 1)
    check_dump(gold, db.nodes(), db.roots, oracle.EUCLIDEAN, dims, oracle.decode_node)


def test_one_vector_in_multiple_trees_is_a_single_leaf():
    gold = G["writer_inline"]["230"]
    db = _build("euclidean", 3, [[0.0, 1.0, 2.0]], 10)
    check_dump(gold, db.nodes(), db.roots, oracle.EUCLIDEAN, 3, oracle.decode_node)


def test_lot_of_random_points_all_92_nodes():
    # golden 5: 100 x 30 uniform, Euclidean, 10 trees — SSE path with a 14-element tail
    gold = G["lot_of_random_points"]
    rng = rng42()
    data = rng.fill_f32(100 * 30).reshape(100, 30)
    db = _build("euclidean", 30, data, 10, rng=rng)
    assert len(gold["tree"]) == 92
    check_dump(gold, db.nodes(), db.roots, oracle.EUCLIDEAN, 30, oracle.decode_node)


def test_little_memory_cosine_all_188_nodes():
    # golden 6: 100 x 3 Cosine, available_memory(0), 2 trees (src/tests/writer.rs:1377-1391). Pins
    # Cosine two_means / create_split / margin, fit_in_memory's u64 gen_range, the tmp-file routing
    # scan, hashbrown iteration order and rayon's LIFO task order.
    gold = G["little_memory"]
    assert gold["distance"] == "cosine" and len(gold["tree"]) == 188
    rng = rng42()
    data = rng.fill_f32(100 * 3).reshape(100, 3)
    db = oracle.Db("cosine", 3)
    for i in range(100):
        db.add_item(i, data[i])
    db.build_memory_limited(rng, n_trees=2, available_memory=0)
    check_dump(gold, db.nodes(), db.roots, oracle.COSINE, 3, oracle.decode_node)
    for key, it in gold["items"].items():   # stored Cosine headers (norm) as printed by the snapshot
        assert fmt4(db.item_header(int(key))[0]) == it["header"]["norm"]


def test_target_n_trees_table():
    for n_items, dims, want in G["target_n_trees"]:
        assert oracle.target_n_trees(None, dims, n_items) == want, (n_items, dims)
    for t in (1, 10, 100):
        assert oracle.target_n_trees(t, 768, 100, 3) == t


def _line_db(column=False):
    db = oracle.Db("euclidean", 2)
    for i in range(100):
        db.add_item(i, [0.0, float(i)] if column else [float(i), 0.0])
    db.build(rng42(), n_trees=50)
    return db


def test_reader_two_dimension_on_a_line():
    # golden 7: src/tests/reader.rs:101-144
    db = _line_db()
    q = G["reader_inline"]
    assert db.nns_by_item(1, 5, search_k=1) == [tuple(x) for x in q["120"]]
    assert db.nns_by_item(0, 5, search_k=2**63) == [tuple(x) for x in q["128"]]
    assert db.nns_by_item(0, 5) == [tuple(x) for x in q["137"]]


def test_reader_column_and_filtering():
    db = _line_db(column=True)
    q = G["reader_inline"]
    assert db.nns_by_item(0, 5) == [tuple(x) for x in q["168"]]
    assert db.nns_by_item(0, 5, candidates=range(0, 2)) == [tuple(x) for x in q["216"]]
    assert db.nns_by_item(0, 5, candidates=range(98, 1000)) == [tuple(x) for x in q["223"]]


def test_reader_single_vector_cosine_and_empty():
    db = oracle.Db("cosine", 3)
    db.add_item(0, [0.00397, 0.553, 0.0])
    db.build(rng42())
    assert db.nns_by_item(0, 1) == [tuple(x) for x in G["reader_inline"]["96"]]
    empty = oracle.Db("euclidean", 2)
    empty.build(rng42())
    assert empty.nns_by_vector([0.0, 0.0], 10) == []


def test_top_k_matches_sorted_order_incl_nan_and_signed_zero():
    # golden 8: median_based_top_k == ascending (OrderedFloat, id) order (tests/reader.rs:283-299)
    r = np.random.default_rng(7)
    for trial in range(50):
        n = int(r.integers(1, 400))
        vals = r.standard_normal(n).astype(np.float32)
        vals[r.integers(0, n, size=n // 7)] = np.nan
        vals[r.integers(0, n, size=n // 9)] = -0.0
        vals[r.integers(0, n, size=n // 9)] = 0.0
        vals[r.integers(0, n, size=n // 11)] = np.inf
        k = int(r.integers(1, n + 1))
        # rerank with Dot: built = -dot(q, v) with q = [1]; d=1 vectors = -vals => distance = vals
        vecs = (-vals).reshape(n, 1)
        rows, dist = oracle.rerank(oracle.DOT_PRODUCT, [1.0], (0, 0), vecs, None, None, np.arange(n), k)

        def key(i):
            v = vals[i]
            return (1, 0.0, i) if np.isnan(v) else (0, float(v) + 0.0, i)
        want = sorted(range(n), key=key)[:k]
        assert rows.tolist() == want


def test_simd_paths_agree_with_scalar_on_integers():
    # golden 9: simple_avx.rs:112-153 / simple_sse.rs:112-151 (integer valued => exact)
    v1 = np.array(([10 + i for i in range(16)] * 4) + [26, 27, 28, 29, 30, 31], dtype=np.float32)
    v2 = np.array([40 + i for i in range(16)] + ([10 + i for i in range(16)] * 3) + [56, 57, 58, 59, 60, 61], dtype=np.float32)
    for n in (70, 22, 16, 35):
        a, b = v1[:n], v2[:n]
        assert oracle.dot(a, b) == float(np.sum(a.astype(np.float64) * b.astype(np.float64)))
        assert oracle.euclid(a, b) == float(np.sum((a.astype(np.float64) - b.astype(np.float64)) ** 2))


# ---- incremental builds: the reference's own add / delete / rebuild scenarios (src/tests/writer.rs) ----
# each scenario: dims, then a list of steps; ("add", id, vec) / ("del", id) / ("build", n_trees, golden line, split_after)
LINE6 = [("add", i, [float(i), 0.0]) for i in range(6)]
SCENARIOS = {
    "overwrite_one_item_incremental": (2, LINE6 + [("build", 1, "403"), ("add", 3, [6.0, 0.0]), ("build", 1, "432")]),
    "delete_one_item_in_a_one_item_db": (2, [("add", 0, [0.0, 0.0]), ("build", 1, "464"), ("del", 0), ("build", 1, "481")]),
    "delete_one_item_in_a_descendant": (2, [("add", 0, [0.0, 0.0]), ("add", 1, [1.0, 0.0]), ("build", 1, "563"), ("del", 0), ("build", 1, "581")]),
    "delete_one_leaf_in_a_split": (2, [("add", i, [float(i), 0.0]) for i in range(3)] + [("build", 1, "605"), ("del", 1), ("build", 1, "627")]),
    "delete_one_item_in_a_single_document_database": (2, [("add", 0, [0.0, 0.0]), ("build", None, "650"), ("del", 0), ("build", None, "667")]),
    "delete_one_item": (2, LINE6 + [("build", 1, "689"), ("del", 3), ("build", 1, "717"), ("del", 1), ("build", 1, "743")]),
    "add_one_item_incrementally_in_an_empty_db": (2, [("build", 1, "767"), ("add", 0, [0.0, 0.0]), ("build", 1, "780")]),
    "add_one_item_incrementally_in_a_one_item_db": (2, [("add", 0, [0.0, 0.0]), ("build", 1, "800"), ("add", 1, [1.0, 0.0]), ("build", 1, "815")]),
    "add_one_item_incrementally_to_create_a_split_node": (2, [("add", 0, [0.0, 0.0]), ("add", 1, [1.0, 0.0]), ("build", 1, "837"), ("add", 2, [2.0, 0.0]), ("build", 1, "853")]),
    "add_one_item_incrementally": (2, LINE6 + [("build", 1, "880"), ("add", 25, [25.0, 0.0]), ("build", 1, "908"), ("add", 8, [8.0, 0.0]), ("build", 1, "939")]),
    "create_root_split_node_with_empty_child": (2, LINE6 + [("build", 1, "1057"), ("del", 1), ("del", 5), ("build", 1, "1086"), ("del", 0), ("build", 1, "1108")]),
    "reuse_node_id": (2, LINE6 + [("build", 1, "1135"), ("del", 4), ("build", 1, "1163"), ("add", 4, [4.0, 0.0]), ("build", 1, "1188"), ("build", 2, "1215")]),
    # a Writer opened with dimensions = 2 over 4-dimensional items (the test does that): only
    # fit_in_descendant sees the 2, which is what split_after = 2 expresses
    "delete_extraneous_tree": (4, [("add", i, [float(i), 0.0, 0.0, 0.0]) for i in range(5)] + [("build", None, "980"), ("build", 2, "1000", 2), ("build", 1, "1025", 2)]),
}


def run_scenario(make_db, steps, check):
    """Drive a db object exposing add_item / del_item / build(rng, n_trees, split_after)."""
    db = make_db()
    rng = rng42()
    for st in steps:
        if st[0] == "add":
            db.add_item(st[1], st[2])
        elif st[0] == "del":
            db.del_item(st[1])
        else:
            split_after = st[3] if len(st) > 3 else None
            db.build_incremental(rng, n_trees=st[1], split_after=split_after)
            check(db, st[2])
    return db


@pytest.mark.parametrize("name", sorted(SCENARIOS))
def test_incremental_scenarios_match_reference_snapshots(name):
    dims, steps = SCENARIOS[name]

    def check(db, line):
        gold = G["writer_inline"][line]
        check_dump(gold, db.nodes(), db.roots, oracle.EUCLIDEAN, dims, oracle.decode_node)

    run_scenario(lambda: oracle.Db("euclidean", dims), steps, check)


def test_incremental_equals_fresh_build_for_a_fresh_index():
    rng_a, rng_b = rng42(), rng42()
    data = rng_a.fill_f32(300 * 24).reshape(300, 24)
    rng_b.fill_f32(300 * 24)
    a, b = oracle.Db("cosine", 24), oracle.Db("cosine", 24)
    for i in range(300):
        a.add_item(i, data[i])
        b.add_item(i, data[i])
    a.build(rng_a, n_trees=5)
    b.build_incremental(rng_b, n_trees=5)
    assert a.nodes() == b.nodes() and a.roots == b.roots


def test_lot_of_random_points_second_snapshot_after_update():
    # write_and_update_lot_of_random_points, second half (tests/writer.rs:310-319): 50 items
    # overwritten, then an incremental build over 10 existing roots. Node ids depend on the order in
    # which the per-root results are merged (rayon `reduce`, writer.rs:1148-1159).
    gold = G["lot_of_random_points_2"]
    rng = rng42()
    data = rng.fill_f32(100 * 30).reshape(100, 30)
    db = oracle.Db("euclidean", 30)
    for i in range(100):
        db.add_item(i, data[i])
    db.build_incremental(rng, n_trees=10)
    check_dump(G["lot_of_random_points"], db.nodes(), db.roots, oracle.EUCLIDEAN, 30, oracle.decode_node)
    for i in range(0, 100, 2):
        db.add_item(i, rng.fill_f32(30))
    db.build_incremental(rng, n_trees=10)
    check_dump(gold, db.nodes(), db.roots, oracle.EUCLIDEAN, 30, oracle.decode_node)
