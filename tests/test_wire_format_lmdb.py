"""The wire format pinned against LMDB files the REFERENCE wrote (VERDICT r1 #7 / SURVEY.md 8f-2).

tests/golden/v0_6_{smol,large}.mdb are the two databases the reference ships for its upgrade tests (src/tests/upgrade.rs): written
by arroy v0.6 through heed / LMDB / the roaring crate. A read-only LMDB page walker (tests/lmdb_ro.py) lists their key/value pairs;
then, with the PRODUCT's codecs (host mirror, arroy_b200/csrc/host.hpp):
  * every key has the layout of src/key.rs:56-68 and sorts like the reference's KeyCodec;
  * every Descendants node (RoaringBitmap::serialize_into bytes of the real `roaring` crate), the Metadata value and every Leaf
    value decode, and re-encode BYTE-IDENTICALLY — this pins roaring_serialize / roaring_deserialize, which round 1 could only
    check against the published format description;
  * after the v0.6 -> v0.7 node rewrite of src/upgrade.rs:183-270 (restated below, checked node for node against the
    reference's post-upgrade dumps) the tables import into the host mirror, Reader::open accepts them, and (on the GPU)
    the reference's query goldens come out: ids 92 / 24 / 78 with distances 2.4881108 / 2.5068686 / 2.5809734
    (src/tests/upgrade.rs:119-128) and ids 1 / 0 / 2 (upgrade.rs:58-67)."""
import os
import struct

import numpy as np
import pytest

import arroy_b200 as ab
from helpers import check_dump, golden
from lmdb_ro import read_main_db

HERE = os.path.dirname(os.path.abspath(__file__))
G = golden()
MODE_METADATA, MODE_UPDATED, MODE_TREE, MODE_ITEM = 0, 1, 2, 3


def load(name):
    kv, meta = read_main_db(os.path.join(HERE, "golden", "v0_6_%s.mdb" % name))
    md, trees, items = None, {}, {}
    for k, v in kv:
        assert len(k) == 8 and k[7] == 0, k.hex()                       # index:u16 BE | mode:u8 | item:u32 BE | 0
        index, mode, item = int.from_bytes(k[0:2], "big"), k[2], int.from_bytes(k[3:7], "big")
        assert index == 0
        if mode == MODE_METADATA:
            assert item == 0
            md = v
        elif mode == MODE_TREE:
            trees[item] = v
        else:
            assert mode == MODE_ITEM
            items[item] = v
    assert [k for k, _ in kv] == sorted(k for k, _ in kv)
    return md, trees, items, meta


def roaring_one(item):   # RoaringBitmap::from_iter(Some(item)).serialize_into: one array container with one value
    return struct.pack("<IIHHIH", 12346, 1, item >> 16, 0, 16, item & 0xffff)


def upgrade_from_0_6(trees, dims):
    """from_0_6_to_current (src/upgrade.rs:183-270): children stored as NodeId (mode byte + u32 BE) become plain tree ids, a child
    that pointed at an ITEM becomes a new one-item Descendants node, a zero normal becomes `None`, a non-zero one gets its header."""
    out = dict(trees)
    last = max(trees)
    for nid in sorted(trees):
        v = trees[nid]
        if v[0] != 2:
            continue
        kids = []
        for o in (1, 6):
            mode, item = v[o], int.from_bytes(v[o + 1:o + 5], "big")
            if mode == MODE_ITEM:
                last += 1
                out[last] = b"\x01" + roaring_one(item)
                kids.append(last)
            else:
                assert mode == MODE_TREE
                kids.append(item)
        vec = v[11:]
        assert len(vec) == 4 * dims
        zero = not np.frombuffer(vec, dtype="<f4").any()
        normal = b"" if zero else struct.pack("<f", 0.0) + vec          # Euclidean::new_header: bias 0.0
        out[nid] = b"\x02" + kids[0].to_bytes(4, "big") + kids[1].to_bytes(4, "big") + normal
    return out


def decode_node_v07(b, metric, dims):
    import oracle
    return oracle.decode_node(b, metric, dims)


@pytest.mark.parametrize("name,n_items,dims,n_trees", [("smol", 6, 2, 1), ("large", 100, 30, 10)])
def test_reference_lmdb_files_decode_and_reencode_byte_identically(name, n_items, dims, n_trees):
    md, trees, items, meta = load(name)
    assert meta["entries"] == 1 + len(trees) + len(items) and len(items) == n_items
    # Metadata: "euclidean\0" | dims u32 BE | bitmap len u32 BE | RoaringBitmap | roots (u32 native endian)
    assert md.startswith(b"euclidean\0") and int.from_bytes(md[10:14], "big") == dims
    assert ab.reencode("metadata", md) == md
    n_desc = 0
    for nid, v in trees.items():
        if v[0] == 1:                                   # Descendants: tag | RoaringBitmap::serialize_into
            assert ab.reencode("node", v) == v, nid
            n_desc += 1
    assert n_desc >= n_trees
    # Leaf values: tag 0 | Header (Euclidean: bias) | dims x f32 — the bytes Writer::add_item stores for the same vector
    env = ab.Env()
    w = ab.Writer(env, 7, dims, "euclidean")
    for it, v in items.items():
        assert len(v) == 1 + 4 + 4 * dims and v[0] == 0
        w.add_item(it, np.frombuffer(v[5:], dtype="<f4"))
    stored = {int.from_bytes(k[3:7], "big"): val for k, val in env.items() if k[2] == MODE_ITEM}
    assert stored == items
    # the upgraded forest equals the reference's own post-upgrade dump, node for node
    up = upgrade_from_0_6(trees, dims)
    gold = G["upgrade_%s_dump" % name]
    import oracle
    roots = np.frombuffer(md[len(md) - 4 * n_trees:], dtype="<u4").tolist()
    check_dump(gold, up, roots, oracle.EUCLIDEAN, dims, oracle.decode_node)
    for nid, v in up.items():                           # and every v0.7 node survives the product's codec unchanged
        assert ab.reencode("node", v) == v, nid


def import_upgraded(name, dims):
    md, trees, items, _ = load(name)
    env = ab.Env(0)
    key = lambda mode, item: (0).to_bytes(2, "big") + bytes([mode]) + item.to_bytes(4, "big") + b"\0"
    env.put_raw(key(MODE_METADATA, 0), md)
    for nid, v in upgrade_from_0_6(trees, dims).items():
        env.put_raw(key(MODE_TREE, nid), v)
    for it, v in items.items():
        env.put_raw(key(MODE_ITEM, it), v)
    return env, items


@pytest.mark.parametrize("name,n_items,dims,n_trees", [("smol", 6, 2, 1), ("large", 100, 30, 10)])
def test_reader_opens_the_imported_reference_database(name, n_items, dims, n_trees):
    env, items = import_upgraded(name, dims)
    with pytest.raises(ab.ArroyError) as ei:
        ab.Reader.open(env, 0, "cosine")
    assert ei.value.kind == "UnmatchingDistance"
    r = ab.Reader.open(env, 0, "euclidean")
    assert (r.n_items(), r.n_trees(), r.dimensions()) == (n_items, n_trees, dims)
    assert list(r.item_ids()) == sorted(items)
    assert np.asarray(r.item_vector(0), dtype=np.float32).tobytes() == items[0][5:]
    if name == "large":   # src/tests/upgrade.rs:117: item 25 printed at full precision
        assert [repr(float(np.float32(x))) for x in G["upgrade_item25"]] == [repr(float(x)) for x in r.item_vector(25)]


@pytest.mark.gpu
def test_reference_query_goldens_on_the_imported_databases():
    env, _ = import_upgraded("large", 30)
    r = ab.Reader.open(env, 0, "euclidean")
    got = r.nns(3).search_k(100).by_vector(np.zeros(30, dtype=np.float32))
    assert [i for i, _ in got] == [i for i, _ in G["upgrade_nns_zero"]] == [92, 24, 78]
    for (_, d), (_, want) in zip(got, G["upgrade_nns_zero"]):
        assert np.float32(d) == np.float32(float(want)), (d, want)       # Rust prints the shortest round-trip repr of the f32
    env2, _ = import_upgraded("smol", 2)
    env2._ctx = env.ctx
    r2 = ab.Reader.open(env2, 0, "euclidean")
    got = r2.nns(3).search_k(100).by_vector([1.0, 0.0])
    assert [(i, float(d)) for i, d in got] == [(i, float(d)) for i, d in G["upgrade_smol_nns"]]
    env2._ctx = None
