"""Edge cases of the C-ABI entry points (empty / ragged / limit inputs), checked against the oracle."""
import numpy as np
import pytest

import arroy_b200 as ab
import oracle

pytestmark = pytest.mark.gpu
SEED = bytes([42] * 32)


@pytest.fixture(scope="module")
def ctx():
    c = ab.Context(0)
    yield c
    c.close()


def test_empty_and_tiny_inputs(ctx):
    d = 40
    data = oracle.synth_rows(SEED, d, 0, 50, 0.5)
    ctx.stage_items_flat("euclidean", np.arange(50, dtype=np.uint32), data)
    side, mg = ctx.side_batch(np.ones(d, np.float32), (0.0, 0.0), np.zeros(0, dtype=np.uint32))
    assert side.size == 0 and mg.size == 0
    r, dist = ctx.rerank(data[0], (0.0, 0.0), np.zeros(0, dtype=np.uint32), 5)
    assert r.size == 0
    r, dist = ctx.rerank(data[0], (0.0, 0.0), np.arange(3, dtype=np.uint32), 10)     # k > candidates
    wr, wd = oracle.rerank(oracle.EUCLIDEAN, data[0], (0, 0), data, np.zeros(50, np.float32), None, np.arange(3, dtype=np.uint32), 10)
    assert r.tolist() == wr.tolist() and dist.tobytes() == wd.tobytes()
    r, dist = ctx.rerank(data[0], (0.0, 0.0), np.arange(50, dtype=np.uint32), 0)      # k == 0
    assert r.size == 0
    out_rows, out_dist, out_len = ctx.rerank_shared(data[:3], None, np.arange(50, dtype=np.uint32), 50)
    for i in range(3):
        wr, wd = oracle.rerank(oracle.EUCLIDEAN, data[i], (0, 0), data, np.zeros(50, np.float32), None, np.arange(50, dtype=np.uint32), 50)
        assert out_rows[i, :out_len[i]].tolist() == wr.tolist() and out_dist[i, :out_len[i]].tobytes() == wd.tobytes()
    # staging zero items is allowed; building on it is not
    ctx.stage_items_flat("euclidean", np.zeros(0, dtype=np.uint32), np.zeros((0, d), dtype=np.float32))
    with pytest.raises(ab.ArroyB200Error):
        ctx.build_trees([SEED], [0], 1)


def test_two_item_nodes_and_split_after_one(ctx):
    # the smallest possible splits: K = 1 forces every node down to single items
    n, d = 64, 32
    data = oracle.synth_rows(SEED, d, 0, n, 0.5)
    ids = np.arange(n, dtype=np.uint32)
    odb = oracle.Db("cosine", d)
    odb.set_items(ids, data)
    rng = oracle.StdRng(SEED)
    user = rng.clone()
    odb.build(rng, n_trees=3, split_after=1)
    ctx.stage_items_flat("cosine", ids, data)
    r1 = oracle.StdRng(user.gen_seed())
    seeds = [r1.gen_seed() for _ in range(3)]
    got = ctx.build_trees(seeds, [0, 1, 2], 3, split_after=1)
    assert got == odb.nodes()


def test_many_trees_run_in_several_waves(ctx, monkeypatch):
    # more trees than one wave holds: results must not depend on the wave size
    n, d, T = 1500, 48, 9
    data = oracle.synth_rows(SEED, d, 0, n, 0.5)
    ids = np.arange(n, dtype=np.uint32)
    ctx.stage_items_flat("euclidean", ids, data)
    user = oracle.StdRng(SEED)
    r1 = oracle.StdRng(user.gen_seed())
    seeds = [r1.gen_seed() for _ in range(T)]
    one_wave = ctx.build_trees(seeds, list(range(T)), T)
    monkeypatch.setenv("ARROY_B200_MAX_WAVE", "4")
    three_waves = ctx.build_trees(seeds, list(range(T)), T)
    monkeypatch.setenv("ARROY_B200_LOCKSTEP", "1")
    lockstep = ctx.build_trees(seeds, list(range(T)), T)
    assert one_wave == three_waves == lockstep


def test_search_batch_by_vector_and_status(ctx):
    n, d, T = 4000, 64, 5
    data = oracle.synth_rows(SEED, d, 0, n, 0.5)
    env = ab.Env(0)
    env._ctx = ctx
    w = ab.Writer(env, 0, d, "cosine")
    w.add_items(np.arange(n, dtype=np.uint32), data)
    w.builder(ab.StdRng.from_seed(SEED)).n_trees(T).build()
    r = ab.Reader.open(env, 0, "cosine")
    out_ids, out_dist, out_len, _ = r.nns_batch_by_item(np.arange(30, dtype=np.uint32), 2000)   # large k, still <= 2048
    odb = oracle.Db("cosine", d)
    odb.set_items(np.arange(n, dtype=np.uint32), data)
    odb.build(oracle.StdRng(SEED), n_trees=T, threads=4)
    for i in range(30):
        want = odb.nns_by_item(i, 2000)
        assert out_ids[i, :out_len[i]].tolist() == [x[0] for x in want]
    env._ctx = None


def test_restage_invalidates_the_device_forest(ctx):
    # a forest validated against n items must not be walked after a restage with fewer items
    # (walk_kernel indexes the per-query bitmap and the item matrix by the forest's rows)
    n, d = 3000, 32
    data = oracle.synth_rows(SEED, d, 0, n, 0.5)
    ctx.stage_items_flat("euclidean", np.arange(n, dtype=np.uint32), data)
    # a one-split forest over all rows: node 0 = split (no normal), nodes 1 / 2 = the two halves
    rows = np.arange(n, dtype=np.uint32)
    ctx.load_forest(kind=[2, 1, 1], left=[1, 0, 0], right=[2, 0, 0], normal_idx=[0xffffffff, 0, 0], normal_hdr0=[0, 0, 0],
                    desc_off=[0, 0, n // 2], desc_len=[0, n // 2, n - n // 2], normals=np.zeros((0, d), np.float32), desc_rows=rows, roots=[0])
    assert ctx.epochs()[1] != 0
    out_rows, out_dist, out_len, status = ctx.search_batch(5, query_rows=[0, 1, 2], search_k=n)
    assert status.tolist() == [0, 0, 0] and out_len.tolist() == [5, 5, 5] and out_rows[:, 0].tolist() == [0, 1, 2]
    e0 = ctx.epochs()
    ctx.stage_items_flat("euclidean", np.arange(100, dtype=np.uint32), data[:100])
    e1 = ctx.epochs()
    assert e1[0] != e0[0] and e1[1] == 0
    with pytest.raises(ab.ArroyB200Error) as ei:
        ctx.search_batch(5, query_rows=[0, 1, 2], search_k=n)
    assert ei.value.code == 5   # ARROY_B200_ERR_NOT_STAGED


@pytest.mark.parametrize("metric", ["euclidean", "cosine", "manhattan"])
def test_count_beyond_the_topk_buffer(ctx, metric):
    # the reference has no limit on count (reader.rs:396-399); k > 2048 takes the full segmented sort
    n, d, k = 6000, 40, 3000
    data = oracle.synth_rows(SEED, d, 0, n, 0.5)
    m = oracle.METRICS[metric]
    ctx.stage_items_flat(metric, np.arange(n, dtype=np.uint32), data)
    h0, _ = ctx.item_headers()
    rows = np.arange(n, dtype=np.uint32)
    for qi in (0, 17):
        qh = oracle.new_header(m, data[qi])
        wr, wd = oracle.rerank(m, data[qi], qh, data, h0, None, rows, k)
        gr, gd = ctx.rerank(data[qi], qh, rows, k)
        assert gr.tolist() == wr.tolist() and gd.tobytes() == wd.tobytes()
    out_rows, out_dist, out_len = ctx.rerank_shared(data[:3], h0[:3], rows, k)
    for i in range(3):
        wr, wd = oracle.rerank(m, data[i], oracle.new_header(m, data[i]), data, h0, None, rows, k)
        assert out_len[i] == k and out_rows[i].tolist() == wr.tolist() and out_dist[i].tobytes() == wd.tobytes()


def test_branch_free_division_equals_div_rn(ctx):
    # create_split / two_means divide by a loop-invariant norm or count; the library does that with the refinement of div.rn's own
    # fast path hoisted out of the loop (exact.cuh UDiv). 2^28 quotients over every operand class (normal, +-0, denormal, huge,
    # NaN / Inf patterns; divisors 2..11, around 1, any exponent) must equal div.rn.f32 bit for bit.
    total_fb = 0
    for seed in (1, 2, 3, 4):
        mism, fb = ctx.selftest_udiv(1 << 24, seed * 0x9E3779B97F4A7C15 % (1 << 63))
        assert mism == 0
        total_fb += fb
    assert 0 < total_fb < 4 * (1 << 24)   # both paths were exercised
