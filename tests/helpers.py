"""Shared helpers for the parity tests (golden comparison, node decoding)."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_golden.json")


def golden():
    with open(GOLDEN) as f:
        return json.load(f)


def fmt4(x):
    """Rust `{:0.4}` of an f32 (exact value, correctly rounded) == C printf of the widened double."""
    return "%.4f" % float(np.float32(x))


def check_dump(gold, nodes, roots, metric, dims, decode_node):
    """Compare a forest ({node id: NodeCodec bytes}) with a parsed reference snapshot.

    Everything the reference's DatabaseHandle dump prints is compared: roots, node ids, node
    kinds, child ids, header values and the first 10 normal components at 4 decimals, and the
    full descendant lists.
    """
    assert list(roots) == gold["roots"]
    assert sorted(nodes.keys()) == sorted(int(k) for k in gold["tree"].keys())
    for key, g in gold["tree"].items():
        n = decode_node(nodes[int(key)], metric, dims)
        assert n["kind"] == g["kind"], (key, n, g)
        if g["kind"] == "descendants":
            assert n["descendants"] == g["descendants"], key
            continue
        assert (n["left"], n["right"]) == (g["left"], g["right"]), key
        if g["normal"] is None:
            assert n["normal"] is None, key
            continue
        hdr_vals = [fmt4(v) for v in n["header"]]
        assert hdr_vals == list(g["header"].values()), (key, hdr_vals, g["header"])
        got = [fmt4(v) for v in n["normal"][:10]]
        assert got == g["normal"], (key, got, g["normal"])
        if not g["truncated"]:
            assert len(n["normal"]) == len(g["normal"])
