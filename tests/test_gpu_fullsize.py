"""Parity at BASELINE.json's full C2 size (1M x 768 Cosine): a few trees compared node-for-node with
the oracle, plus size-independent forest invariants and determinism."""
import hashlib
import os

import numpy as np
import pytest

import arroy_b200 as ab
import oracle

pytestmark = pytest.mark.gpu
SEED = bytes([42] * 32)


def test_c2_full_size_trees_identical_to_oracle_and_invariants():
    n, d, T = 1_000_000, 768, 4
    threads = min(os.cpu_count() or 4, 32)
    data = oracle.synth_rows(SEED, d, 0, n, 0.5, threads=threads)
    ids = np.arange(n, dtype=np.uint32)
    ctx = ab.Context(0)
    ctx.stage_items_flat("cosine", ids, data)
    user = oracle.StdRng(SEED)
    r1 = oracle.StdRng(user.gen_seed())
    seeds = [r1.gen_seed() for _ in range(T)]
    got = ctx.build_trees(seeds, list(range(T)), T)
    st = ctx.build_stats()
    again = ctx.build_trees(seeds, list(range(T)), T)
    digest = lambda nodes: hashlib.sha256(b"".join(k.to_bytes(4, "little") + nodes[k] for k in sorted(nodes))).hexdigest()
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
    ctx.close()
