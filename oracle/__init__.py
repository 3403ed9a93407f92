"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes wrapper over oracle/liboracle.so, the CPU restatement of the reference's
distance / split / re-rank path (see the headers of rng.hpp, distance.hpp, build.hpp for
the reference file:line each function follows). Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this package; the product
package (arroy_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

EUCLIDEAN, COSINE, DOT_PRODUCT, MANHATTAN = 0, 1, 2, 3
BQ_EUCLIDEAN, BQ_COSINE, BQ_MANHATTAN = 4, 5, 6
METRICS = {"euclidean": 0, "cosine": 1, "dot-product": 2, "manhattan": 3,
           "binary quantized euclidean": 4, "binary quantized cosine": 5, "binary quantized manhattan": 6}

_f32p = C.POINTER(C.c_float)
_u32p = C.POINTER(C.c_uint32)
_u8p = C.POINTER(C.c_uint8)
NODE_SINK = C.CFUNCTYPE(None, C.c_void_p, C.c_uint32, _u8p, C.c_uint64)


def build_lib(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("oracle_capi.cpp", "build.hpp", "distance.hpp", "rng.hpp")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    L = C.CDLL(build_lib())
    vp, u8, u32, u64, i64, f32, i32 = C.c_void_p, C.c_uint8, C.c_uint32, C.c_uint64, C.c_int64, C.c_float, C.c_int

    def sig(name, res, *args):
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = list(args)

    sig("oracle_rng_from_seed", vp, _u8p)
    sig("oracle_rng_seed_from_u64", vp, u64)
    sig("oracle_rng_clone", vp, vp)
    sig("oracle_rng_free", None, vp)
    sig("oracle_rng_next_u32", u32, vp)
    sig("oracle_rng_next_u64", u64, vp)
    sig("oracle_rng_gen_f32", f32, vp)
    sig("oracle_rng_fill_f32", None, vp, _f32p, u64)
    sig("oracle_rng_gen_bool", i32, vp)
    sig("oracle_rng_gen_seed", None, vp, _u8p)
    sig("oracle_rng_gen_range_u32_incl", u32, vp, u32, u32)
    sig("oracle_rng_gen_range_u64", u64, vp, u64, u64)
    sig("oracle_rng_sample2", None, vp, u32, _u32p)
    sig("oracle_synth_rows", None, _u8p, u64, u64, u64, f32, _f32p, i32)
    sig("oracle_dot", f32, _f32p, _f32p, u64)
    sig("oracle_euclid", f32, _f32p, _f32p, u64)
    sig("oracle_margin", f32, i32, _f32p, f32, f32, _f32p, f32, f32, u64)
    sig("oracle_built_distance", f32, i32, _f32p, f32, f32, _f32p, f32, f32, u64)
    sig("oracle_normalized_distance", f32, i32, f32)
    sig("oracle_new_header", None, i32, _f32p, u64, _f32p)
    sig("oracle_side_batch", None, i32, _f32p, f32, f32, _f32p, _f32p, _f32p, u64, _u32p, u64, _u8p, _f32p)
    sig("oracle_create_split", None, i32, vp, _f32p, _f32p, _f32p, u64, _u32p, u32, _f32p, _f32p)
    sig("oracle_rerank", u32, i32, _f32p, f32, f32, _f32p, _f32p, _f32p, u64, _u32p, u64, u32, _u32p, _f32p)
    sig("oracle_dot_preprocess", None, _f32p, u64, u64, _f32p, _f32p)
    sig("oracle_set_rerank_dims", None, u64)
    sig("oracle_bq_quantize", u64, _f32p, u64, _f32p)
    sig("oracle_db_set_user_dims", None, vp, u64)
    sig("oracle_target_n_trees", u64, i64, u64, u64, u64)
    sig("oracle_split_imbalance", C.c_double, u64, u64)
    sig("oracle_db_new", vp, i32, u64)
    sig("oracle_db_free", None, vp)
    sig("oracle_db_add_item", None, vp, u32, _f32p)
    sig("oracle_db_del_item", i32, vp, u32)
    sig("oracle_db_set_items", None, vp, u64, _u32p, _f32p)
    sig("oracle_db_build", i32, vp, vp, i64, u64, i32)
    sig("oracle_db_build_memory_limited", i32, vp, vp, i64, u64, u64)
    sig("oracle_db_build_incremental", i32, vp, vp, i64, u64)
    sig("oracle_db_n_nodes", u64, vp)
    sig("oracle_db_n_roots", u64, vp)
    sig("oracle_db_roots", None, vp, _u32p)
    sig("oracle_db_scanned_rows", u64, vp)
    sig("oracle_db_item_header", None, vp, u32, _f32p)
    sig("oracle_db_emit_nodes", None, vp, NODE_SINK, vp)
    sig("oracle_db_nns_by_item", i64, vp, u32, u64, u64, u64, _u32p, i64, _u32p, _f32p, _u32p, u64, C.POINTER(u64))
    sig("oracle_db_nns_by_vector", i64, vp, _f32p, u64, u64, u64, _u32p, i64, _u32p, _f32p, _u32p, u64, C.POINTER(u64))
    _LIB = L
    return L


def _fp(a):
    return a.ctypes.data_as(_f32p) if a is not None else None


def _up(a):
    return a.ctypes.data_as(_u32p) if a is not None else None


def f32c(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class StdRng:
    """rand 0.8.5 StdRng (ChaCha12)."""

    def __init__(self, seed=None, *, handle=None):
        L = lib()
        if handle is not None:
            self.h = handle
        else:
            seed = bytes(seed) if seed is not None else bytes([42] * 32)
            assert len(seed) == 32
            self.h = L.oracle_rng_from_seed((C.c_uint8 * 32)(*seed))

    @classmethod
    def seed_from_u64(cls, s):
        return cls(handle=lib().oracle_rng_seed_from_u64(s))

    def clone(self):
        return StdRng(handle=lib().oracle_rng_clone(self.h))

    def __del__(self):
        try:
            lib().oracle_rng_free(self.h)
        except Exception:
            pass

    def next_u32(self):
        return lib().oracle_rng_next_u32(self.h)

    def next_u64(self):
        return lib().oracle_rng_next_u64(self.h)

    def gen_f32(self):
        return lib().oracle_rng_gen_f32(self.h)

    def fill_f32(self, n):
        out = np.empty(n, dtype=np.float32)
        lib().oracle_rng_fill_f32(self.h, _fp(out), n)
        return out

    def gen_bool(self):
        return bool(lib().oracle_rng_gen_bool(self.h))

    def gen_seed(self):
        out = (C.c_uint8 * 32)()
        lib().oracle_rng_gen_seed(self.h, out)
        return bytes(out)

    def gen_range_u32_incl(self, lo, hi):
        return lib().oracle_rng_gen_range_u32_incl(self.h, lo, hi)

    def gen_range_u64(self, lo, hi):
        return lib().oracle_rng_gen_range_u64(self.h, lo, hi)

    def sample2(self, length):
        out = (C.c_uint32 * 2)()
        lib().oracle_rng_sample2(self.h, length, out)
        return out[0], out[1]


def synth_rows(seed, d, row0, rows, centre=0.0, threads=1):
    out = np.empty((rows, d), dtype=np.float32)
    lib().oracle_synth_rows((C.c_uint8 * 32)(*bytes(seed)), d, row0, rows, centre, _fp(out), threads)
    return out


def dot(a, b):
    a, b = f32c(a), f32c(b)
    return lib().oracle_dot(_fp(a), _fp(b), a.size)


def euclid(a, b):
    a, b = f32c(a), f32c(b)
    return lib().oracle_euclid(_fp(a), _fp(b), a.size)


def margin(metric, normal, nh, item, ih):
    normal, item = f32c(normal), f32c(item)
    return lib().oracle_margin(metric, _fp(normal), nh[0], nh[1], _fp(item), ih[0], ih[1], normal.size)


def built_distance(metric, p, ph, q, qh):
    p, q = f32c(p), f32c(q)
    return lib().oracle_built_distance(metric, _fp(p), ph[0], ph[1], _fp(q), qh[0], qh[1], p.size)


def normalized_distance(metric, d):
    return lib().oracle_normalized_distance(metric, d)


def new_header(metric, v):
    v = f32c(v)
    out = np.zeros(2, dtype=np.float32)
    lib().oracle_new_header(metric, _fp(v), v.size, _fp(out))
    return float(out[0]), float(out[1])


def side_batch(metric, normal, nh, vectors, h0, h1, rows, want_margin=True):
    normal, vectors = f32c(normal), f32c(vectors)
    rows = np.ascontiguousarray(rows, dtype=np.uint32)
    side = np.empty(rows.size, dtype=np.uint8)
    mg = np.empty(rows.size, dtype=np.float32) if want_margin else None
    lib().oracle_side_batch(metric, _fp(normal), nh[0], nh[1], _fp(vectors), _fp(h0), _fp(h1), vectors.shape[1],
                            _up(rows), rows.size, side.ctypes.data_as(_u8p), _fp(mg))
    return side, mg


def create_split(metric, rng, vectors, h0, h1, rows):
    vectors = f32c(vectors)
    rows = np.ascontiguousarray(rows, dtype=np.uint32)
    d = vectors.shape[1]
    normal = np.empty(d, dtype=np.float32)
    hdr = np.zeros(2, dtype=np.float32)
    lib().oracle_create_split(metric, rng.h, _fp(vectors), _fp(h0), _fp(h1), d, _up(rows), rows.size, _fp(normal), _fp(hdr))
    return normal, (float(hdr[0]), float(hdr[1]))


def rerank(metric, query, qh, vectors, h0, h1, rows, count):
    query, vectors = f32c(query), f32c(vectors)
    rows = np.ascontiguousarray(rows, dtype=np.uint32)
    out_rows = np.empty(max(count, 1), dtype=np.uint32)
    out_dist = np.empty(max(count, 1), dtype=np.float32)
    k = lib().oracle_rerank(metric, _fp(query), qh[0], qh[1], _fp(vectors), _fp(h0), _fp(h1), vectors.shape[1],
                            _up(rows), rows.size, count, _up(out_rows), _fp(out_dist))
    return out_rows[:k].copy(), out_dist[:k].copy()


def bq_quantize(vectors):
    """BinaryQuantized::from_slice + ::iter of every row: the +-1.0 values of the quantized vectors, 64 * ceil(d / 64) columns."""
    vectors = f32c(vectors)
    squeeze = vectors.ndim == 1
    v2 = vectors.reshape(1, -1) if squeeze else vectors
    n, d = v2.shape
    dp = (d + 63) // 64 * 64
    out = np.empty((n, dp), dtype=np.float32)
    for i in range(n):
        row = np.ascontiguousarray(v2[i])
        lib().oracle_bq_quantize(_fp(row), d, out[i].ctypes.data_as(_f32p))
    return out[0] if squeeze else out


def set_rerank_dims(dims):
    """binary-quantized indexes: the `dimensions` normalized_distance divides by in the next rerank() calls (0 = vector length)"""
    lib().oracle_set_rerank_dims(dims)


def dot_preprocess(vectors):
    vectors = f32c(vectors)
    n, d = vectors.shape
    extra = np.empty(n, dtype=np.float32)
    norm = np.empty(n, dtype=np.float32)
    lib().oracle_dot_preprocess(_fp(vectors), n, d, _fp(extra), _fp(norm))
    return extra, norm


def target_n_trees(n_trees, dims, n_items, n_roots=0):
    return lib().oracle_target_n_trees(-1 if n_trees is None else n_trees, dims, n_items, n_roots)


class Db:
    """In-memory stand-in for one arroy index (Writer + Reader of the reference)."""

    def __init__(self, metric, dims):
        self.metric = METRICS[metric] if isinstance(metric, str) else metric
        self.dims = dims
        self.h = lib().oracle_db_new(self.metric, dims)
        self._keep = None

    def __del__(self):
        try:
            lib().oracle_db_free(self.h)
        except Exception:
            pass

    def add_item(self, item, vector):
        v = f32c(vector)
        assert v.size == self.dims
        lib().oracle_db_add_item(self.h, item, _fp(v))

    def del_item(self, item):
        return bool(lib().oracle_db_del_item(self.h, item))

    def set_items(self, ids, vectors):
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        vectors = f32c(vectors)
        self._keep = (ids, vectors)
        lib().oracle_db_set_items(self.h, ids.size, _up(ids), _fp(vectors))

    def set_user_dims(self, dims):
        """binary-quantized indexes: Reader::dimensions (the vectors handed to set_items are the padded +-1 form)"""
        lib().oracle_db_set_user_dims(self.h, dims)

    def build(self, rng, n_trees=None, split_after=None, threads=1):
        rc = lib().oracle_db_build(self.h, rng.h, -1 if n_trees is None else n_trees, split_after or 0, threads)
        if rc != 0:
            raise RuntimeError("oracle build failed")

    def build_incremental(self, rng, n_trees=None, split_after=None):
        """Writer::build in general (fresh or on top of existing trees), 1-thread execution order."""
        rc = lib().oracle_db_build_incremental(self.h, rng.h, -1 if n_trees is None else n_trees, split_after or 0)
        if rc != 0:
            raise RuntimeError("oracle build failed")

    def build_memory_limited(self, rng, n_trees=None, split_after=None, available_memory=0):
        """Writer::build with available_memory set, on a 1-thread rayon pool (golden pinning only)."""
        rc = lib().oracle_db_build_memory_limited(self.h, rng.h, -1 if n_trees is None else n_trees, split_after or 0, available_memory)
        if rc != 0:
            raise RuntimeError("oracle build failed")

    @property
    def roots(self):
        n = lib().oracle_db_n_roots(self.h)
        out = np.empty(n, dtype=np.uint32)
        lib().oracle_db_roots(self.h, _up(out))
        return out.tolist()

    @property
    def scanned_rows(self):
        return lib().oracle_db_scanned_rows(self.h)

    def item_header(self, item):
        out = np.zeros(2, dtype=np.float32)
        lib().oracle_db_item_header(self.h, item, _fp(out))
        return float(out[0]), float(out[1])

    def nodes(self):
        """{tree node id: NodeCodec bytes} — what TmpNodes::put would have received."""
        out = {}

        def sink(_arg, node_id, ptr, length):
            out[node_id] = C.string_at(ptr, length)

        cb = NODE_SINK(sink)
        lib().oracle_db_emit_nodes(self.h, cb, None)
        return out

    def _nns(self, fn, key, count, search_k, oversampling, candidates, want_candidates):
        out_ids = np.empty(max(count, 1), dtype=np.uint32)
        out_dist = np.empty(max(count, 1), dtype=np.float32)
        cand = None if candidates is None else np.ascontiguousarray(sorted(candidates), dtype=np.uint32)
        ncap = 1 << 22 if want_candidates else 0
        out_c = np.empty(ncap, dtype=np.uint32) if want_candidates else None
        n_c = C.c_uint64(0)
        k = fn(self.h, key, count, search_k or 0, oversampling or 0, _up(cand), -1 if cand is None else cand.size,
               _up(out_ids), _fp(out_dist), _up(out_c), ncap, C.byref(n_c))
        if k < 0:
            return None
        res = list(zip(out_ids[:k].tolist(), out_dist[:k].tolist()))
        if want_candidates:
            return res, out_c[: n_c.value].copy()
        return res

    def nns_by_item(self, item, count, search_k=None, oversampling=None, candidates=None, want_candidates=False):
        return self._nns(lib().oracle_db_nns_by_item, item, count, search_k, oversampling, candidates, want_candidates)

    def nns_by_vector(self, vector, count, search_k=None, oversampling=None, candidates=None, want_candidates=False):
        v = f32c(vector)
        return self._nns(lib().oracle_db_nns_by_vector, _fp(v), count, search_k, oversampling, candidates, want_candidates)


# ---- NodeCodec decoding helpers (for tests) ------------------------------------------------

def roaring_deserialize(b):
    """Portable RoaringBitmap format without run containers (cookie 12346)."""
    cookie, n = np.frombuffer(b[:8], dtype="<u4")
    assert cookie == 12346, cookie
    n = int(n)
    desc = np.frombuffer(b[8:8 + 4 * n], dtype="<u2").reshape(n, 2)
    offs = np.frombuffer(b[8 + 4 * n:8 + 8 * n], dtype="<u4")
    out = []
    for i in range(n):
        key, card = int(desc[i, 0]), int(desc[i, 1]) + 1
        o = int(offs[i])
        if card > 4096:
            words = np.frombuffer(b[o:o + 8192], dtype="<u8")
            bits = np.unpackbits(words.view(np.uint8), bitorder="little")
            lows = np.nonzero(bits)[0]
        else:
            lows = np.frombuffer(b[o:o + 2 * card], dtype="<u2")
        out.extend(((key << 16) | int(x)) for x in lows)
    return out


def decode_node(b, metric, dims):
    """NodeCodec bytes -> dict (src/node.rs:246-282)."""
    tag = b[0]
    if tag == 1:
        return {"kind": "descendants", "descendants": roaring_deserialize(b[1:])}
    assert tag == 2, tag
    left = int.from_bytes(b[1:5], "big")
    right = int.from_bytes(b[5:9], "big")
    rest = b[9:]
    if len(rest) == 0:
        return {"kind": "split", "left": left, "right": right, "normal": None}
    nh = 2 if metric == DOT_PRODUCT else 1
    hdr = np.frombuffer(rest[:4 * nh], dtype=np.float32)
    vec = np.frombuffer(rest[4 * nh:], dtype=np.float32)
    assert vec.size == dims
    return {"kind": "split", "left": left, "right": right, "header": hdr.copy(), "normal": vec.copy()}
