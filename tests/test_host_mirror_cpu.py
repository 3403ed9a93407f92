"""CPU-only tests of the C++ host mirror: the parts of Writer / Reader that need no device
(storage byte formats, error behaviour, single-leaf builds, seeds) — mirrored from the reference's
src/tests/writer.rs and src/tests/reader.rs."""
import numpy as np
import pytest

import arroy_b200 as ab
import oracle
from helpers import check_dump, golden

G = golden()


def rng42():
    return ab.StdRng.from_seed(bytes([42] * 32))


def test_product_rng_matches_oracle_rng_stream():
    a, b = rng42(), oracle.StdRng(bytes([42] * 32))
    assert a.fill_f32(1000).tobytes() == b.fill_f32(1000).tobytes()
    assert [a.next_u32() for _ in range(70)] == [b.next_u32() for _ in range(70)]
    a, b = ab.StdRng.seed_from_u64(42), oracle.StdRng.seed_from_u64(42)
    assert [a.next_u32() for _ in range(20)] == [b.next_u32() for _ in range(20)]


@pytest.mark.parametrize("line,n_items,n_trees", [("190", 1, None), ("210", 1, 1), ("230", 1, 10), ("254", 3, 1)])
def test_single_leaf_builds_match_reference_snapshots(line, n_items, n_trees):
    # write_one_vector*, write_vectors_until_there_is_a_descendants (src/tests/writer.rs:181-264)
    gold = G["writer_inline"][line]
    env = ab.Env()
    w = ab.Writer(env, 0, 3, "euclidean")
    for i in range(n_items):
        w.add_item(i, [0.0, 1.0, 2.0] if n_items == 1 else [i, i, i])
    b = w.builder(rng42())
    if n_trees is not None:
        b.n_trees(n_trees)
    b.build()
    check_dump(gold, env.tree_nodes(), ab.Reader.open(env, 0, "euclidean")._roots(), oracle.EUCLIDEAN, 3, oracle.decode_node)
    assert not w.need_build()


def test_use_u32_max_for_a_vec():
    env = ab.Env()
    w = ab.Writer(env, 0, 3, "euclidean")
    w.add_item(2**32 - 1, [0.0, 1.0, 2.0])
    w.builder(rng42()).n_trees(1).build()
    assert oracle.decode_node(env.tree_nodes()[0], 0, 3)["descendants"] == [2**32 - 1]


def test_metadata_version_and_key_bytes():
    env = ab.Env()
    w = ab.Writer(env, 7, 3, "cosine")
    w.add_item(1, [1, 2, 3])
    w.add_item(70000, [3, 2, 1])
    w.builder(rng42()).build()
    kv = dict(env.items())
    meta_key = bytes([0, 7, 0, 0, 0, 0, 0, 0])          # [index u16 BE][mode][item u32 BE][0] — src/key.rs:56-68
    ver_key = bytes([0, 7, 0, 0, 0, 0, 1, 0])
    assert kv[ver_key] == bytes([0, 0, 0, 0, 0, 0, 0, 7, 0, 0, 0, 0])   # Version 0.7.0, 3 x u32 BE
    m = kv[meta_key]
    assert m.startswith(b"cosine\x00" + (3).to_bytes(4, "big"))
    size = int.from_bytes(m[11:15], "big")
    assert oracle.roaring_deserialize(m[15:15 + size]) == [1, 70000]
    assert np.frombuffer(m[15 + size:], dtype=np.uint32).tolist() == [0]
    leaf = kv[bytes([0, 7, 3, 0, 0, 0, 1, 0])]
    assert leaf[0] == 0 and len(leaf) == 1 + 4 + 12
    assert np.frombuffer(leaf[1:5], np.float32)[0] == np.float32(oracle.new_header(oracle.COSINE, [1, 2, 3])[0])
    assert not any(k[2] == 1 for k in kv)  # updated keys are consumed by the build


def test_writer_errors_and_item_bookkeeping():
    env = ab.Env()
    w = ab.Writer(env, 0, 2, "euclidean")
    assert w.is_empty() and w.need_build()
    with pytest.raises(ab.ArroyError) as ei:
        w.add_item(0, [1.0, 2.0, 3.0])
    assert ei.value.kind == "InvalidVecDimension" and str(ei.value) == "Invalid vector dimensions. Got 3 but expected 2"
    w.add_item(0, [0.0, 0.0])
    assert w.contains_item(0) and not w.contains_item(1)
    assert w.item_vector(0).tolist() == [0.0, 0.0] and w.item_vector(9) is None
    w.append_item(5, [1.0, 1.0])
    with pytest.raises(ab.ArroyError) as ei:
        w.append_item(5, [1.0, 1.0])
    assert ei.value.kind == "InvalidItemAppend"
    assert w.del_item(5) and not w.del_item(5)
    w.builder(rng42()).n_trees(1).build()
    assert not w.need_build()
    w.del_item(0)
    assert w.need_build()
    w.clear()
    assert len(env) == 0


def test_reader_open_errors():
    # src/tests/reader.rs:31-79, :245-281
    env = ab.Env()
    with pytest.raises(ab.ArroyError) as ei:
        ab.Reader.open(env, 0, "euclidean")
    assert ei.value.kind == "MissingMetadata"
    assert str(ei.value) == "Metadata are missing on index 0, You must build your database before attempting to read it"
    w = ab.Writer(env, 0, 2, "euclidean")
    w.add_item(0, [0.0, 0.0])
    w.builder(rng42()).build()
    with pytest.raises(ab.ArroyError) as ei:
        ab.Reader.open(env, 0, "cosine")
    assert ei.value.kind == "UnmatchingDistance" and str(ei.value) == "Invalid distance provided. Got cosine but expected euclidean"
    w.del_item(0)
    with pytest.raises(ab.ArroyError) as ei:
        ab.Reader.open(env, 0, "euclidean")
    assert ei.value.kind == "NeedBuild" and str(ei.value) == "The trees have not been built after an update on index 0"


def test_search_in_empty_database_and_cancel():
    env = ab.Env()
    w = ab.Writer(env, 0, 2, "euclidean")
    w.builder(rng42()).build()
    r = ab.Reader.open(env, 0, "euclidean")
    assert r.nns(10).by_vector([0.0, 0.0]) == [] and r.n_items() == 0 and r.n_trees() == 0
    w.add_item(0, [0.0, 0.0])
    with pytest.raises(ab.ArroyError) as ei:
        w.builder(rng42()).cancel(lambda: True).build()
    assert ei.value.kind == "BuildCancelled" and str(ei.value) == "The corresponding build process has been cancelled"
    with pytest.raises(ab.ArroyError) as ei:
        r.nns(5).by_vector([1.0, 2.0, 3.0])
    assert ei.value.kind == "InvalidVecDimension"


def test_progress_reports_main_steps():
    env = ab.Env()
    w = ab.Writer(env, 0, 2, "euclidean")
    w.add_item(0, [0.0, 0.0])
    seen = []
    w.builder(rng42()).progress(seen.append).build()
    assert seen[0] == "PreProcessingTheItems" and "WritingTheDescendantsAndMetadata" in seen


def test_failed_or_cancelled_build_leaves_the_table_untouched():
    # the reference builds inside a RwTxn that is dropped on error (src/writer.rs:487-629): nothing of a failed build may be
    # visible afterwards — in particular the Updated markers, or need_build() would report a built index
    env = ab.Env()
    w = ab.Writer(env, 0, 2, "euclidean")
    w.add_item(0, [0.0, 0.0])
    w.add_item(1, [1.0, 0.0])
    w.builder(rng42()).build()                      # n <= split_after: a single Descendants node, no device needed
    w.add_item(2, [2.0, 0.0])                       # now n > dimensions: the build needs the device, which this box lacks
    w.del_item(0)
    before = env.items()
    assert w.need_build()
    with pytest.raises(ab.ArroyError):
        w.builder(rng42()).cancel(lambda: True).build()
    assert env.items() == before and w.need_build()
    try:
        w.builder(rng42()).build()                  # fails without a CUDA device (no CPU fallback) ...
        built = True
    except (ab.ArroyError, Exception):
        built = False
    if not built:
        assert env.items() == before and w.need_build()   # ... and must not leave a half-written index behind
        with pytest.raises(ab.ArroyError) as ei:
            ab.Reader.open(env, 0, "euclidean")
        assert ei.value.kind == "NeedBuild"


def test_reader_that_outlives_a_write_to_its_index_is_refused():
    # the in-memory table has no RoTxn snapshots: a reader must not silently mix its decoded trees with newer items
    env = ab.Env()
    w = ab.Writer(env, 0, 2, "euclidean")
    w.add_item(0, [0.0, 0.0])
    w.builder(rng42()).build()
    r = ab.Reader.open(env, 0, "euclidean")
    other = ab.Writer(env, 1, 2, "euclidean")       # a write to ANOTHER index does not disturb the reader
    other.add_item(5, [1.0, 1.0])
    other.builder(rng42()).build()
    assert list(r.item_vector(0)) == [0.0, 0.0]
    w.add_item(1, [1.0, 0.0])
    w.builder(rng42()).build()
    with pytest.raises(ab.ArroyError) as ei:
        r.item_vector(0)
    assert ei.value.kind == "NeedBuild"
    with pytest.raises(ab.ArroyError) as ei:
        r.nns(1).by_vector([0.0, 0.0])
    assert ei.value.kind == "NeedBuild"
    assert ab.Reader.open(env, 0, "euclidean").n_items() == 2
