"""side() through the bf16 shadow (kernels.cuh scan_claim_shadow) for the bias metrics and the many-trees regime.
The at-size tests of test_gpu_fullsize.py run Cosine and DotProduct through it with few trees; here: Euclidean and Manhattan
(margin = bias + dot, src/distance/euclidean.rs:79-81, manhattan.rs:82-84), and a wave of 32 trees, where EVERY node goes through
the shadow and the fused root pass covers four batches of normals. Node bytes must equal the oracle's; the scan statistics prove
that the pre-filter really ran and that it left rows for the exact re-score."""
import os

import numpy as np
import pytest

import arroy_b200 as ab
import oracle

pytestmark = pytest.mark.gpu
SEED = bytes([42] * 32)
THREADS = min(os.cpu_count() or 4, 32)


def tree_seeds(T):
    user = oracle.StdRng(SEED)
    r1 = oracle.StdRng(user.gen_seed())
    return [r1.gen_seed() for _ in range(T)]


@pytest.fixture(scope="module")
def ctx():
    c = ab.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("metric,n,d,T", [
    ("euclidean", 70_000, 64, 2),      # 4.5 M elements: shadow-eligible; nodes above 8192 rows take it
    ("manhattan", 70_000, 64, 2),
    ("euclidean", 40_000, 128, 32),    # 32 trees on the GPU: every node takes it
    ("cosine", 40_000, 128, 33),       # five batches of root normals, the last one with a single normal
])
def test_shadow_scans_match_the_oracle(ctx, metric, n, d, T):
    data = oracle.synth_rows(SEED, d, 0, n, 0.5, threads=THREADS)
    ids = np.arange(n, dtype=np.uint32)
    ctx.stage_items_flat(metric, ids, data)
    got = ctx.build_trees(tree_seeds(T), list(range(T)), T)
    sh = ctx.build_shadow_stats()
    odb = oracle.Db(metric, d)
    odb.set_items(ids, data)
    odb.build(oracle.StdRng(SEED), n_trees=T, threads=min(T, THREADS))
    want = odb.nodes()
    assert got.keys() == want.keys()
    bad = [k for k in want if want[k] != got[k]]
    assert not bad, "node bytes differ for ids %s" % bad[:10]
    assert ctx.build_stats()["scanned_rows"] == odb.scanned_rows
    assert sh["rows_via_bf16_shadow"] > n            # the pre-filter ran on more than one level
    assert 0 < sh["rows_rescored_f32"] < sh["rows_via_bf16_shadow"] // 4
    # (the fused root pass belongs to the persistent schedule; whichever schedule the library picked, its accounting must be consistent)
    assert (sh["rows_in_fused_root_pass"], sh["fused_root_rows_read"]) in ((T * n, n), (0, 0))
