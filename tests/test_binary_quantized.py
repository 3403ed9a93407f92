"""Binary-quantized distances (src/distance/binary_quantized_{euclidean,cosine,manhattan}.rs,
src/unaligned_vector/binary_quantized.rs, two_means_binary_quantized src/distance/mod.rs:173-223).

Pinned against the reference: the quantization golden of src/tests/binary_quantized.rs:13-45 (signs, +0.0 -> 1, the layout of the
stored Leaf) and the identities its doc comments state (simple.rs:85-120, binary_quantized_euclidean.rs:99-116, _manhattan.rs:103-112:
the popcount kernels equal the dot product / squared distance / L1 distance of the +-1 vectors). Everything else of the BQ path —
two_means_binary_quantized, the three create_split, normalized_distance — has no golden in the reference: the oracle restates the
code ("parity unpinned" for those, as DESIGN.md says) and the GPU path is compared with the oracle bit for bit."""
import numpy as np
import pytest

import arroy_b200 as ab
import oracle

SEED = bytes([42] * 32)
BQ = ["binary quantized euclidean", "binary quantized cosine", "binary quantized manhattan"]


def test_quantization_matches_the_reference_golden():
    v = np.array([-2.0, -1.0, 0.0, -0.1, 2.0, 2.0, -12.4, 21.2, -2.0, -1.0, 0.0, 1.0, 2.0, 2.0, -12.4, 21.2], dtype=np.float32)
    want = [-1.0, -1.0, 1.0, -1.0, 1.0, 1.0, -1.0, 1.0, -1.0, -1.0, 1.0, 1.0, 1.0, 1.0, -1.0, 1.0]      # src/tests/binary_quantized.rs:24-43
    q = oracle.bq_quantize(v)
    assert q.size == 64 and q[:16].tolist() == want and (q[16:] == -1.0).all()      # a word of 64 bits; missing bits are 0 = -1.0
    p = ab._capi.bq_quantize(v)
    assert p.tobytes() == q.tobytes()
    z = np.array([0.0, -0.0, np.nan, -np.nan, np.inf, -np.inf], dtype=np.float32)      # is_sign_positive: the sign bit decides
    assert oracle.bq_quantize(z)[:6].tolist() == ab._capi.bq_quantize(z)[:6].tolist() == [1.0, -1.0, 1.0, -1.0, 1.0, -1.0]


def test_popcount_kernels_equal_the_plus_minus_one_arithmetic():
    # simple.rs:85-120: dot_product_binary_quantized == dot product of the +-1 vectors; euclidean.rs:99-116: 4 * popcount(u ^ v) ==
    # their squared distance; manhattan.rs:103-112: 2 * popcount == their L1 distance — padding bits included
    rng = np.random.default_rng(3)
    for d in (1, 63, 64, 65, 130, 768):
        a, b = oracle.bq_quantize(rng.standard_normal(d).astype(np.float32)), oracle.bq_quantize(rng.standard_normal(d).astype(np.float32))
        dp = a.size
        bits_a, bits_b = a > 0, b > 0
        assert oracle.margin(oracle.BQ_COSINE, a, (0.0, 0.0), b, (0.0, 0.0)) == float(2 * int((bits_a == bits_b).sum()) - dp) == float(np.dot(a.astype(np.float64), b))
        assert oracle.built_distance(oracle.BQ_EUCLIDEAN, a, (0, 0), b, (0, 0)) == 4.0 * int((bits_a != bits_b).sum()) == float(((a - b) ** 2).sum())
        assert oracle.built_distance(oracle.BQ_MANHATTAN, a, (0, 0), b, (0, 0)) == 2.0 * int((bits_a != bits_b).sum()) == float(np.abs(a - b).sum())
        n = oracle.new_header(oracle.BQ_COSINE, a)[0]
        assert n == np.float32(np.sqrt(np.float32(dp)))
        want = (np.float32(1.0) - np.float32(np.dot(a, b)) / (n * n)) / np.float32(2.0) if n * n != 0 else 0.0
        assert oracle.built_distance(oracle.BQ_COSINE, a, (n, 0), b, (n, 0)) == np.float32(want)


@pytest.fixture(scope="module")
def ctx():
    c = ab.Context(0)
    yield c
    c.close()


@pytest.mark.gpu
@pytest.mark.parametrize("metric", BQ)
@pytest.mark.parametrize("n,d,trees", [(3000, 100, 4), (20000, 768, 3), (700, 64, 2)])
def test_bq_forest_sides_and_rerank_match_the_oracle(ctx, metric, n, d, trees):
    m = oracle.METRICS[metric]
    raw = oracle.synth_rows(SEED, d, 0, n, 0.5, threads=4)
    pm1 = oracle.bq_quantize(raw)                         # what the reference stores (as bits) and iterates (as +-1)
    dp = pm1.shape[1]
    ids = np.arange(n, dtype=np.uint32)
    ctx.stage_items_flat(metric, ids, raw)                # the library quantizes f32 vectors on the way (Writer::add_item)
    h0, _ = ctx.item_headers()
    want_h = np.float32(np.sqrt(np.float32(dp))) if metric.endswith("cosine") else np.float32(0)
    assert (h0 == want_h).all()
    # side() / margin against a +-1 normal
    normal = oracle.bq_quantize(oracle.synth_rows(SEED, d, n + 5, 1, 0.5)[0])
    rows = np.arange(0, n, 3, dtype=np.uint32)
    hdr = (2.0, 0.0) if not metric.endswith("cosine") else (0.0, 0.0)
    side, mg = ctx.side_batch(normal, hdr, rows)
    oh = np.full(n, want_h, dtype=np.float32)
    wside, wmg = oracle.side_batch(m, normal, hdr, pm1, oh, None, rows)
    assert side.tobytes() == wside.tobytes() and mg.tobytes() == wmg.tobytes()
    # create_split = two_means_binary_quantized + the sign-bit normal, consuming the rng identically
    user = oracle.StdRng(SEED)
    r1 = oracle.StdRng(user.gen_seed())
    seeds = [r1.gen_seed() for _ in range(trees)]
    for trial in range(3):
        sub = np.sort(np.random.default_rng(trial).choice(n, size=min(n, 500 + 37 * trial), replace=False)).astype(np.uint32)
        org = oracle.StdRng(seeds[0])
        wn, wh = oracle.create_split(m, org, pm1, oh, None, sub)
        key = np.frombuffer(seeds[0], dtype="<u4")
        gn, gh, pos = ctx.create_split(key, 0, sub)
        assert gn.tobytes() == wn.tobytes() and gh[0] == wh[0]
        probe = oracle.StdRng(seeds[0])
        for _ in range(pos):
            probe.next_u32()
        assert probe.next_u32() == org.next_u32()
    # whole forest, node bytes (the normals are emitted as bit strings)
    odb = oracle.Db(metric, dp)
    odb.set_items(ids, pm1)
    odb.set_user_dims(d)
    odb.build(oracle.StdRng(SEED), n_trees=trees, split_after=d, threads=trees)
    got = ctx.build_trees(seeds, list(range(trees)), trees, split_after=d)
    want = odb.nodes()
    assert got.keys() == want.keys()
    assert all(got[k] == want[k] for k in want)
    # re-rank: distances are popcount arithmetic, normalized by the index' dimensions
    oracle.set_rerank_dims(d)
    try:
        cand = np.arange(1, n, 2, dtype=np.uint32)
        for qi in (0, 11):
            wr, wd = oracle.rerank(m, pm1[qi], (float(want_h), 0.0), pm1, oh, None, cand, 25)
            gr, gd = ctx.rerank(pm1[qi], (float(want_h), 0.0), cand, 25)
            assert gr.tolist() == wr.tolist() and gd.tobytes() == wd.tobytes()
    finally:
        oracle.set_rerank_dims(0)
