"""Read-only LMDB page walker (test infrastructure): enough of the on-disk format of LMDB 0.9 (lmdb.h / mdb.c, MDB_DATA_VERSION 1)
to list the (key, value) pairs of the unnamed main database of a data.mdb file — what heed's `env.open_database(&rtxn, None)`
plus `database.iter()` yield. Branch, leaf and overflow pages; no sub-databases, no DUPSORT."""
import struct

P_BRANCH, P_LEAF, P_OVERFLOW, P_META = 0x01, 0x02, 0x04, 0x08
F_BIGDATA, F_SUBDATA, F_DUPDATA = 0x01, 0x02, 0x04
MAGIC = 0xBEEFC0DE
PAGEHDR = 16


def _meta(buf, off):
    pgno, pad, flags, lower, upper = struct.unpack_from("<QHHHH", buf, off)
    assert flags & P_META, "not a meta page"
    magic, version, address, mapsize = struct.unpack_from("<IIQQ", buf, off + PAGEHDR)
    assert magic == MAGIC, hex(magic)
    dbs = []
    o = off + PAGEHDR + 24
    for _ in range(2):   # FREE_DBI, MAIN_DBI
        md_pad, md_flags, md_depth, branch, leaf, overflow, entries, root = struct.unpack_from("<IHHQQQQQ", buf, o)
        dbs.append(dict(pad=md_pad, flags=md_flags, depth=md_depth, branch_pages=branch, leaf_pages=leaf, overflow_pages=overflow, entries=entries, root=root))
        o += 48
    last_pg, txnid = struct.unpack_from("<QQ", buf, o)
    return dict(version=version, dbs=dbs, last_pg=last_pg, txnid=txnid)


def read_main_db(path):
    """[(key bytes, value bytes)] of the main database, in key order, and the meta information used."""
    buf = open(path, "rb").read()
    m0 = _meta(buf, 0)
    psize = m0["dbs"][0]["pad"] or 4096        # the FREE_DBI record's md_pad holds the page size
    m1 = _meta(buf, psize)
    meta = m1 if m1["txnid"] > m0["txnid"] else m0
    main = meta["dbs"][1]
    out = []

    def page(pgno):
        return pgno * psize

    def walk(pgno):
        off = page(pgno)
        _pg, _pad, flags, lower, upper = struct.unpack_from("<QHHHH", buf, off)
        nkeys = (lower - PAGEHDR) // 2
        ptrs = struct.unpack_from("<%dH" % nkeys, buf, off + PAGEHDR)
        if flags & P_BRANCH:
            for p in ptrs:
                lo, hi, nflags, ksize = struct.unpack_from("<HHHH", buf, off + p)
                walk(lo | (hi << 16) | (nflags << 32))
        elif flags & P_LEAF:
            for p in ptrs:
                lo, hi, nflags, ksize = struct.unpack_from("<HHHH", buf, off + p)
                dsize = lo | (hi << 16)
                key = buf[off + p + 8: off + p + 8 + ksize]
                assert not (nflags & (F_SUBDATA | F_DUPDATA)), "sub-databases / dupsort are not supported"
                d0 = off + p + 8 + ksize
                if nflags & F_BIGDATA:
                    (opg,) = struct.unpack_from("<Q", buf, d0)
                    ooff = page(opg)
                    _opgno, _opad, oflags, npages = struct.unpack_from("<QHHI", buf, ooff)
                    assert oflags & P_OVERFLOW
                    val = buf[ooff + PAGEHDR: ooff + PAGEHDR + dsize]
                else:
                    val = buf[d0: d0 + dsize]
                out.append((bytes(key), bytes(val)))
        else:
            raise ValueError("unexpected page flags 0x%x at page %d" % (flags, pgno))

    if main["root"] != 0xFFFFFFFFFFFFFFFF:
        walk(main["root"])
    assert len(out) == main["entries"], (len(out), main["entries"])
    return out, dict(psize=psize, txnid=meta["txnid"], **main)
