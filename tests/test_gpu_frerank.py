"""Fused re-rank (bf16 shadow pre-filter + exact re-score, frerank.cuh) must return exactly what the
plain distance + top-k kernels and the oracle (reader.rs:381-399) return."""
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


def _both(ctx, monkeypatch, q, qh, rows_per_query, k):
    offs = np.zeros(len(rows_per_query) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(r) for r in rows_per_query])
    rows = np.concatenate(rows_per_query).astype(np.uint32)
    monkeypatch.setenv("ARROY_B200_FRERANK", "0")
    plain = ctx.rerank_batch(q, qh, rows, offs, k)
    monkeypatch.setenv("ARROY_B200_FRERANK", "1")
    c0 = ctx.counters()
    fused = ctx.rerank_batch(q, qh, rows, offs, k)
    c1 = ctx.counters()
    assert c1["fused_rerank_batches"] > c0["fused_rerank_batches"]
    assert plain[2].tolist() == fused[2].tolist()
    for i in range(q.shape[0]):
        n = plain[2][i]
        assert plain[0][i, :n].tolist() == fused[0][i, :n].tolist()
        assert plain[1][i, :n].tobytes() == fused[1][i, :n].tobytes()
    return fused, c1["fused_rerank_fallbacks"] - c0["fused_rerank_fallbacks"]


@pytest.mark.parametrize("metric,d", [("euclidean", 96), ("cosine", 768), ("dot-product", 200), ("cosine", 33), ("euclidean", 20)])
def test_fused_rerank_matches_plain_kernels_and_oracle(ctx, monkeypatch, metric, d):
    n, nq, k = 20_000, 24, 100
    data = oracle.synth_rows(SEED, d, 0, n + nq, 0.5)
    ctx.stage_items_flat(metric, np.arange(n + nq, dtype=np.uint32), data)
    h0, _ = ctx.item_headers()
    q, qh = data[n:], h0[n:]
    rng = np.random.default_rng(3)
    lists = [np.sort(rng.choice(n, int(rng.integers(600, 7000)), replace=False)) for _ in range(nq)]   # ragged candidate lists
    (out_rows, out_dist, out_len), fallbacks = _both(ctx, monkeypatch, q, qh, lists, k)
    assert fallbacks == 0
    m = oracle.METRICS[metric]
    hdr = h0 if metric == "cosine" else np.zeros(n + nq, np.float32)
    for i in (0, 11, nq - 1):
        wr, wd = oracle.rerank(m, q[i], (float(qh[i]), 0.0), data, hdr, None, lists[i].astype(np.uint32), k)
        assert out_rows[i, :out_len[i]].tolist() == wr.tolist() and out_dist[i, :out_len[i]].tobytes() == wd.tobytes()
    _both(ctx, monkeypatch, q, qh, lists, 1)
    _both(ctx, monkeypatch, q, qh, [l[:700] for l in lists], 650)   # k close to the number of candidates


def test_fused_rerank_with_ties_and_degenerate_rows(ctx, monkeypatch):
    n, d, nq = 9_000, 64, 16
    rng = np.random.default_rng(5)
    base = rng.standard_normal((40, d)).astype(np.float32)
    data = base[rng.integers(0, 40, n)]           # many exactly equal candidates: ties are broken by id
    data[::7] = 0.0
    data[1::11] *= np.float32(1e18)
    data[2::13] *= np.float32(1e-18)
    data[5, 3] = np.nan
    data[77, 0] = np.inf
    q = np.concatenate([base[:8], rng.standard_normal((nq - 8, d)).astype(np.float32)])
    lists = [np.arange(0, n, 1 + (i % 3)) for i in range(nq)]
    lists = [l[:8000] for l in lists]
    for metric in ("cosine", "euclidean", "dot-product"):
        ctx.stage_items_flat(metric, np.arange(n, dtype=np.uint32), data)
        qh = np.sqrt((q.astype(np.float64) ** 2).sum(1)).astype(np.float32) if metric == "cosine" else None
        _both(ctx, monkeypatch, q, qh, lists, 50)   # too many survivors -> falls back inside, results still identical
