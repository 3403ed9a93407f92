"""Parity at BASELINE.json's sizes ("close the parity holes at BASELINE sizes"):
  C2  1M x 768 Cosine: a few trees node-for-node vs the oracle, forest invariants, determinism, and 100 top-100
      by_item queries (ids and distances) through the device walk + re-rank vs oracle.Db.nns_by_item
      (src/reader.rs:317-401);
  C3  path: DotProduct at 1M x 768 incl. DotProduct::preprocess (src/distance/dot_product.rs:119-165), 2 trees;
  C4  row width: Cosine d = 1536, 500k rows, 2 trees;
  C5  full 4096 queries x 100k shared candidates, 64 sampled queries vs oracle.rerank (src/reader.rs:381-399).
Bit-exact everywhere (integer ids / node bytes; float distances compared as bytes)."""
import hashlib
import os

import numpy as np
import pytest

import arroy_b200 as ab
import oracle

pytestmark = pytest.mark.gpu
SEED = bytes([42] * 32)
THREADS = min(os.cpu_count() or 4, 32)


def digest(nodes):
    return hashlib.sha256(b"".join(k.to_bytes(4, "little") + nodes[k] for k in sorted(nodes))).hexdigest()


def tree_seeds(T):
    user = oracle.StdRng(SEED)
    r1 = oracle.StdRng(user.gen_seed())
    return [r1.gen_seed() for _ in range(T)]


@pytest.fixture(scope="module")
def ctx():
    c = ab.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def c2_data():
    n, d = 1_000_000, 768
    return oracle.synth_rows(SEED, d, 0, n, 0.5, threads=THREADS)


def forest_arrays(nodes, metric, d):
    """{node id: NodeCodec bytes} -> the arrays arroy_b200_load_forest takes (rows == item ids here)."""
    nn = max(nodes) + 1
    kind = np.zeros(nn, np.uint8)
    left, right, nidx = np.zeros(nn, np.uint32), np.zeros(nn, np.uint32), np.full(nn, 0xffffffff, np.uint32)
    nh0, doff, dlen = np.zeros(nn, np.float32), np.zeros(nn, np.uint32), np.zeros(nn, np.uint32)
    normals, desc = [], []
    n_desc = 0
    for i, b in nodes.items():
        nd = oracle.decode_node(b, metric, d)
        if nd["kind"] == "descendants":
            kind[i] = 1
            doff[i], dlen[i] = n_desc, len(nd["descendants"])
            desc.append(np.asarray(nd["descendants"], dtype=np.uint32))
            n_desc += dlen[i]
        else:
            kind[i] = 2
            left[i], right[i] = nd["left"], nd["right"]
            if nd["normal"] is not None:
                nidx[i] = len(normals)
                nh0[i] = nd["header"][0]
                normals.append(nd["normal"])
    normals = np.stack(normals) if normals else np.zeros((0, d), np.float32)
    return dict(kind=kind, left=left, right=right, normal_idx=nidx, normal_hdr0=nh0, desc_off=doff, desc_len=dlen, normals=normals,
                desc_rows=np.concatenate(desc) if desc else np.zeros(0, np.uint32))


def test_c2_full_size_trees_identical_to_oracle_invariants_and_queries(ctx, c2_data):
    n, d, T = 1_000_000, 768, 4
    data = c2_data
    ids = np.arange(n, dtype=np.uint32)
    ctx.stage_items_flat("cosine", ids, data)
    seeds = tree_seeds(T)
    got = ctx.build_trees(seeds, list(range(T)), T)
    st = ctx.build_stats()
    again = ctx.build_trees(seeds, list(range(T)), T)
    assert digest(got) == digest(again), "two builds with the same seeds must be byte-identical"
    # invariants (the reference's assert_validity, src/reader.rs:508-589): per tree every item appears in
    # exactly one Descendants node, no Descendants node exceeds split_after (= dimensions)
    kinds = {k: v[0] for k, v in got.items()}
    children = {k: (int.from_bytes(v[1:5], "big"), int.from_bytes(v[5:9], "big")) for k, v in got.items() if v[0] == 2}
    for root in range(T):
        seen = np.zeros(n, dtype=np.uint8)
        stack = [root]
        while stack:
            node = stack.pop()
            if kinds[node] == 1:
                items = np.array(oracle.roaring_deserialize(got[node][1:]), dtype=np.int64)
                assert items.size <= d
                assert not seen[items].any()
                seen[items] = 1
            else:
                stack.extend(children[node])
        assert seen.all()
    # the oracle builds the same trees on the CPU (a few seconds per tree)
    odb = oracle.Db("cosine", d)
    odb.set_items(ids, data)
    odb.build(oracle.StdRng(SEED), n_trees=T, threads=T)
    want = odb.nodes()
    assert got.keys() == want.keys()
    assert digest(got) == digest(want)
    assert st["scanned_rows"] == odb.scanned_rows
    # ---- queries at size: top-100 re-rank vs the CPU ids (BASELINE configs[1]) ----------------------------------
    ctx.load_forest(roots=np.arange(T, dtype=np.uint32), **forest_arrays(got, oracle.COSINE, d))
    qitems = np.arange(0, n, n // 100, dtype=np.uint32)[:100]
    for k, sk in ((100, 0), (10, 5000)):
        out_rows, out_dist, out_len, status = ctx.search_batch(k, query_rows=qitems, search_k=sk)
        assert not status.any()
        for i, it in enumerate(qitems):
            w = odb.nns_by_item(int(it), k, search_k=sk or None)
            assert out_rows[i, :out_len[i]].tolist() == [x[0] for x in w], (k, sk, int(it))
            assert out_dist[i, :out_len[i]].tobytes() == np.array([x[1] for x in w], dtype=np.float32).tobytes()
    # by_vector: vectors that are not in the index (Cosine header = their norm, reader.rs:72-73)
    qv = oracle.synth_rows(SEED, d, n + 5, 20, 0.5)
    qh = np.array([oracle.new_header(oracle.COSINE, v)[0] for v in qv], dtype=np.float32)
    out_rows, out_dist, out_len, status = ctx.search_batch(100, queries=qv, qhdr0=qh)
    assert not status.any()
    for i in range(20):
        w = odb.nns_by_vector(qv[i], 100)
        assert out_rows[i, :out_len[i]].tolist() == [x[0] for x in w]
        assert out_dist[i, :out_len[i]].tobytes() == np.array([x[1] for x in w], dtype=np.float32).tobytes()


def test_c3_dot_product_1m_x_768_preprocess_and_trees(ctx, c2_data):
    n, d, T = 1_000_000, 768, 2
    data = c2_data
    ids = np.arange(n, dtype=np.uint32)
    ctx.stage_items_flat("dot-product", ids, data)
    extra, norm = ctx.dot_preprocess()
    w_extra, w_norm = oracle.dot_preprocess(data)
    assert extra.tobytes() == w_extra.tobytes() and norm.tobytes() == w_norm.tobytes()
    got = ctx.build_trees(tree_seeds(T), list(range(T)), T)
    st = ctx.build_stats()
    odb = oracle.Db("dot-product", d)
    odb.set_items(ids, data)
    odb.build(oracle.StdRng(SEED), n_trees=T, threads=T)
    want = odb.nodes()
    assert got.keys() == want.keys() and digest(got) == digest(want)
    assert st["scanned_rows"] == odb.scanned_rows
    # queries on the DotProduct forest (by_item uses the preprocessed header: extra_dim enters the margin)
    ctx.load_forest(roots=np.arange(T, dtype=np.uint32), **forest_arrays(got, oracle.DOT_PRODUCT, d))
    qitems = np.arange(7, n, n // 40, dtype=np.uint32)[:40]
    out_rows, out_dist, out_len, status = ctx.search_batch(100, query_rows=qitems, search_k=3000)
    assert not status.any()
    for i, it in enumerate(qitems):
        w = odb.nns_by_item(int(it), 100, search_k=3000)
        assert out_rows[i, :out_len[i]].tolist() == [x[0] for x in w]
        assert out_dist[i, :out_len[i]].tobytes() == np.array([x[1] for x in w], dtype=np.float32).tobytes()


def test_c4_row_width_cosine_d1536_500k(ctx):
    n, d, T = 500_000, 1536, 2
    data = oracle.synth_rows(SEED, d, 0, n, 0.5, threads=THREADS)
    ids = np.arange(n, dtype=np.uint32)
    ctx.stage_items_flat("cosine", ids, data)
    got = ctx.build_trees(tree_seeds(T), list(range(T)), T)
    st = ctx.build_stats()
    odb = oracle.Db("cosine", d)
    odb.set_items(ids, data)
    odb.build(oracle.StdRng(SEED), n_trees=T, threads=T)
    want = odb.nodes()
    assert got.keys() == want.keys() and digest(got) == digest(want)
    assert st["scanned_rows"] == odb.scanned_rows


def test_c5_full_4096_x_100k_sampled_queries_vs_oracle(ctx):
    n, nq, d, k = 100_000, 4096, 768, 100
    data = oracle.synth_rows(SEED, d, 0, n + nq, 0.5, threads=THREADS)
    ctx.stage_items_flat("cosine", np.arange(n + nq, dtype=np.uint32), data)
    h0, _ = ctx.item_headers()
    rows = np.arange(n, dtype=np.uint32)
    before = ctx.rerank_stats()
    out_rows, out_dist, out_len = ctx.rerank_shared(data[n:], h0[n:], rows, k)
    after = ctx.rerank_stats()
    assert after["prefilter_chunks"] > before["prefilter_chunks"] and after["fallback_chunks"] == before["fallback_chunks"]
    assert (out_len == k).all()
    # size-independent property: every query's distances ascend, ties by id
    assert (np.diff(out_dist, axis=1) >= 0).all()
    sample = np.linspace(0, nq - 1, 64).astype(int)
    from concurrent.futures import ThreadPoolExecutor
    m = oracle.COSINE

    def one(i):
        return oracle.rerank(m, data[n + i], (float(h0[n + i]), 0.0), data, h0, None, rows, k)
    with ThreadPoolExecutor(max_workers=THREADS) as ex:
        res = list(ex.map(one, sample))
    for i, (wr, wd) in zip(sample, res):
        assert out_rows[i].tolist() == wr.tolist() and out_dist[i].tobytes() == wd.tobytes(), int(i)
