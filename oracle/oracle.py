"""CPU ORACLE (test infrastructure, not product code) -- Python face of oracle/dj_oracle.c.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  It offers

* ctypes bindings to the C restatement (``libdj_oracle.so``, built by ``oracle/Makefile``);
* an independent numpy restatement of the same steps (vectorised murmur3, counting-sort
  partition, sort+searchsorted inner join) used to cross-check the C code;
* helpers that simulate the reference's N-rank pipeline in one process
  (src/distributed_join.cpp:134-340: hash_partition -> all-to-all -> per-rank inner join).

Reference citations live in dj_oracle.c's header.  Parity status: join cardinality / row
multiset pinned by the reference's analytical tests (G1) and generator invariants (G5);
key -> partition-id assignment under MURMUR3 is "parity unpinned" (cuDF 0.19 not vendored).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libdj_oracle.so")

HASH_IDENTITY = 0
HASH_MURMUR3 = 1
SEED_NVLINK = 12345678  # src/distributed_join.cpp:211
SEED_IB = 87654321  # src/distributed_join.cpp:160
DEFAULT_HASH_SEED = 0  # cudf::DEFAULT_HASH_SEED in cuDF 0.19 (src/shuffle_on.hpp:50)
GEN_SEED = 1234  # generate_dataset/generate_dataset.cuh:44


class GenParams(C.Structure):
    _fields_ = [
        ("nb", C.c_int64),
        ("np", C.c_int64),
        ("rand_max", C.c_int64),
        ("selectivity", C.c_double),
        ("seed", C.c_uint64),
        ("unique", C.c_int32),
        ("pad", C.c_int32),
    ]


def build(force: bool = False) -> str:
    """Compile the C oracle (gcc, OpenMP).  Building the checker is not using it."""
    src = os.path.join(_HERE, "dj_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libdj_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        i64p = C.POINTER(C.c_int64)
        L.oracle_murmur3_i64.restype = C.c_uint32
        L.oracle_murmur3_i64.argtypes = [C.c_int64, C.c_uint32]
        L.oracle_row_hash_i64.restype = C.c_uint32
        L.oracle_row_hash_i64.argtypes = [C.c_int64, C.c_uint32, C.c_int]
        L.oracle_partition_ids_i64.restype = None
        L.oracle_partition_ids_i64.argtypes = [i64p, C.c_int64, C.c_uint32, C.c_int, C.c_int, C.POINTER(C.c_int32)]
        L.oracle_hash_partition_i64.restype = None
        L.oracle_hash_partition_i64.argtypes = [i64p, i64p, C.c_int64, C.c_uint32, C.c_int, C.c_int, i64p, i64p, i64p]
        L.oracle_inner_join_i64.restype = C.c_int64
        L.oracle_inner_join_i64.argtypes = [i64p, i64p, C.c_int64, i64p, i64p, C.c_int64, i64p, i64p, i64p, i64p, C.c_int64]
        L.oracle_multiset_checksum4.restype = None
        L.oracle_multiset_checksum4.argtypes = [i64p, i64p, i64p, i64p, C.c_int64, C.POINTER(C.c_uint64)]
        L.oracle_build_bitmap.restype = None
        L.oracle_build_bitmap.argtypes = [C.POINTER(GenParams), C.c_int, C.POINTER(C.c_uint32)]
        L.oracle_generate_rows.restype = C.c_int64
        L.oracle_generate_rows.argtypes = [C.POINTER(GenParams), C.c_int, C.c_int, C.c_int64, C.c_int64,
                                           C.POINTER(C.c_uint32), i64p, i64p]
        L.oracle_generate_analytical.restype = None
        L.oracle_generate_analytical.argtypes = [C.c_int64, C.c_int64, C.c_int64, i64p, i64p]
        L.oracle_partitioned_join_omp.restype = C.c_int64
        L.oracle_partitioned_join_omp.argtypes = [i64p, i64p, C.c_int64, i64p, i64p, C.c_int64, C.c_int, C.c_uint32,
                                                  C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.POINTER(C.c_int)]
        L.oracle_num_threads.restype = C.c_int
        L.oracle_set_threads.restype = None
        L.oracle_set_threads.argtypes = [C.c_int]
        _lib = L
    return _lib


def _p(a: np.ndarray, t=C.c_int64):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def _i64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int64)


# ----------------------------------------------------------------------------- C oracle API
def murmur3_i64(key: int, seed: int) -> int:
    return int(lib().oracle_murmur3_i64(C.c_int64(int(np.int64(key))), C.c_uint32(seed)))


def row_hash_i64(key: int, seed: int, hash_id: int = HASH_MURMUR3) -> int:
    return int(lib().oracle_row_hash_i64(C.c_int64(int(np.int64(key))), C.c_uint32(seed), hash_id))


def partition_ids(keys, seed, nparts, hash_id=HASH_MURMUR3) -> np.ndarray:
    keys = _i64(keys)
    out = np.empty(keys.size, dtype=np.int32)
    lib().oracle_partition_ids_i64(_p(keys), keys.size, seed, hash_id, nparts, _p(out, C.c_int32))
    return out


def hash_partition(keys, payload, nparts, seed, hash_id=HASH_MURMUR3):
    """cudf::hash_partition restated.  Returns (keys_out, payload_out, offsets[nparts+1])."""
    keys, payload = _i64(keys), _i64(payload)
    ko, po = np.empty_like(keys), np.empty_like(payload)
    off = np.empty(nparts + 1, dtype=np.int64)
    lib().oracle_hash_partition_i64(_p(keys), _p(payload), keys.size, seed, hash_id, nparts, _p(ko), _p(po), _p(off))
    return ko, po, off


def inner_join(lk, lp, rk, rp, count_only=False):
    """cudf::inner_join restated.  Returns (n_out, (lk, lp, rk, rp) or None)."""
    lk, lp, rk, rp = _i64(lk), _i64(lp), _i64(rk), _i64(rp)
    n = lib().oracle_inner_join_i64(_p(lk), _p(lp), lk.size, _p(rk), _p(rp), rk.size, None, None, None, None, 0)
    if count_only:
        return int(n), None
    outs = [np.empty(n, dtype=np.int64) for _ in range(4)]
    n2 = lib().oracle_inner_join_i64(_p(lk), _p(lp), lk.size, _p(rk), _p(rp), rk.size, *[_p(o) for o in outs], n)
    assert n2 == n
    return int(n), tuple(outs)


def multiset_checksum4(c0, c1, c2, c3):
    cols = [_i64(c) for c in (c0, c1, c2, c3)]
    out = (C.c_uint64 * 2)()
    lib().oracle_multiset_checksum4(*[_p(c) for c in cols], cols[0].size, out)
    return int(out[0]), int(out[1])


def gen_params(nb, np_, selectivity, rand_max, unique=True, seed=GEN_SEED) -> GenParams:
    return GenParams(int(nb), int(np_), int(rand_max), float(selectivity), int(seed), 1 if unique else 0, 0)


def build_bitmap(g: GenParams, src: int) -> np.ndarray:
    bm = np.zeros((g.rand_max + 1 + 31) // 32, dtype=np.uint32)
    lib().oracle_build_bitmap(C.byref(g), src, _p(bm, C.c_uint32))
    return bm


def generate_rows(g: GenParams, which: int, src: int, row_begin: int, count: int, bitmap=None, materialize=True):
    """Rows [row_begin, row_begin+count) of source rank src's local build(0)/probe(1) table.
    Returns (keys, payload, hits)."""
    if which == 1 and not g.unique and bitmap is None:
        bitmap = build_bitmap(g, src)
    keys = np.empty(count, dtype=np.int64) if materialize else None
    pay = np.empty(count, dtype=np.int64) if materialize else None
    hits = lib().oracle_generate_rows(C.byref(g), which, src, row_begin, count,
                                      _p(bitmap, C.c_uint32) if bitmap is not None else None, _p(keys), _p(pay))
    return keys, pay, int(hits)


def generate_tables_distributed(g: GenParams, rank: int, world: int):
    """What rank `rank` holds after src/generate_table.cuh:155-272: every source rank s deals
    rows [n/N*rank, n/N*(rank+1)) of its local tables to `rank`, received in source order."""
    out = []
    for which, n in ((0, g.nb), (1, g.np)):
        chunk = n // world
        ks, ps = [], []
        for s in range(world):
            bm = build_bitmap(g, s) if (which == 1 and not g.unique) else None
            k, p, _ = generate_rows(g, which, s, chunk * rank, chunk, bm)
            ks.append(k)
            ps.append(p)
        out.append((np.concatenate(ks), np.concatenate(ps)))
    return out[0], out[1]


def generate_global_tables(g: GenParams, world: int):
    """The union of every rank's tables after generate_tables_distributed (src/generate_table.cuh:155-272):
    source rank s contributes rows [0, (n // world) * world) of its local build / probe table.  Generated
    straight into two global SoA tables (row order is irrelevant to a join).  Returns
    ((bk, bp), (pk, pp), hits) where hits counts the probe rows drawn as matches."""
    out, hits = [], 0
    for which, n in ((0, g.nb), (1, g.np)):
        per_src = (n // world) * world
        keys = np.empty(per_src * world, dtype=np.int64)
        pay = np.empty(per_src * world, dtype=np.int64)
        for s in range(world):
            bm = build_bitmap(g, s) if (which == 1 and not g.unique) else None
            k, p = keys[s * per_src:(s + 1) * per_src], pay[s * per_src:(s + 1) * per_src]
            h = lib().oracle_generate_rows(C.byref(g), which, s, 0, per_src,
                                           _p(bm, C.c_uint32) if bm is not None else None, _p(k), _p(p))
            hits += int(h) if which == 1 else 0
        out.append((keys, pay))
    return out[0], out[1], hits


def generate_analytical(mult: int, row_begin: int, count: int):
    k = np.empty(count, dtype=np.int64)
    p = np.empty(count, dtype=np.int64)
    lib().oracle_generate_analytical(mult, row_begin, count, _p(k), _p(p))
    return k, p


def partitioned_join_omp(bk, bp, pk, pp, nparts=1024, seed=SEED_NVLINK, checksum=False):
    """CPU baseline: OpenMP murmur3 hash partition + per-partition hash join.
    Returns dict(n_out, seconds, threads, checksum)."""
    bk, bp, pk, pp = _i64(bk), _i64(bp), _i64(pk), _i64(pp)
    sec = C.c_double(0)
    thr = C.c_int(0)
    ck = (C.c_uint64 * 2)()
    n = lib().oracle_partitioned_join_omp(_p(bk), _p(bp), bk.size, _p(pk), _p(pp), pk.size, nparts, seed,
                                          C.byref(sec), ck if checksum else None, C.byref(thr))
    return {"n_out": int(n), "seconds": sec.value, "threads": thr.value,
            "checksum": (int(ck[0]), int(ck[1])) if checksum else None}


def set_threads(n: int) -> None:
    lib().oracle_set_threads(int(n))


def num_threads() -> int:
    return int(lib().oracle_num_threads())


# ------------------------------------------------------------------ numpy restatement
def np_murmur3_i64(keys, seed: int) -> np.ndarray:
    """Vectorised MurmurHash3_x86_32 over the 8 LE bytes of each int64 key."""
    u = _i64(keys).view(np.uint64)
    c1, c2 = np.uint32(0xCC9E2D51), np.uint32(0x1B873593)
    h = np.full(u.shape, seed, dtype=np.uint32)

    def rotl(x, r):
        return (x << np.uint32(r)) | (x >> np.uint32(32 - r))

    with np.errstate(over="ignore"):
        for i in range(2):
            k = (u >> np.uint64(32 * i)).astype(np.uint32)
            k = k * c1
            k = rotl(k, 15)
            k = k * c2
            h = h ^ k
            h = rotl(h, 13)
            h = h * np.uint32(5) + np.uint32(0xE6546B64)
        h = h ^ np.uint32(8)
        h ^= h >> np.uint32(16)
        h = h * np.uint32(0x85EBCA6B)
        h ^= h >> np.uint32(13)
        h = h * np.uint32(0xC2B2AE35)
        h ^= h >> np.uint32(16)
    return h


def np_partition_ids(keys, seed, nparts, hash_id=HASH_MURMUR3) -> np.ndarray:
    with np.errstate(over="ignore"):
        h = np_murmur3_i64(keys, seed) if hash_id == HASH_MURMUR3 else _i64(keys).astype(np.uint32)
        rh = h + np.uint32(0x9E3779B9)
    return (rh % np.uint32(nparts)).astype(np.int32)


def np_inner_join(lk, lp, rk, rp):
    """Sort + searchsorted inner join (independent of the C hash join).  Returns 4 columns."""
    lk, lp, rk, rp = _i64(lk), _i64(lp), _i64(rk), _i64(rp)
    if lk.size == 0 or rk.size == 0:
        e = np.empty(0, dtype=np.int64)
        return e, e, e, e
    order = np.argsort(lk, kind="stable")
    ls, lps = lk[order], lp[order]
    lo = np.searchsorted(ls, rk, side="left")
    hi = np.searchsorted(ls, rk, side="right")
    cnt = hi - lo
    ridx = np.repeat(np.arange(rk.size), cnt)
    starts = np.repeat(lo, cnt)
    within = np.arange(cnt.sum()) - np.repeat(np.cumsum(cnt) - cnt, cnt)
    lidx = starts + within
    return ls[lidx], lps[lidx], rk[ridx], rp[ridx]


def sort_rows(*cols):
    """Canonical order of a row multiset (lexicographic over all columns)."""
    cols = [_i64(c) for c in cols]
    if cols[0].size == 0:
        return tuple(cols)
    order = np.lexsort(tuple(reversed(cols)))
    return tuple(c[order] for c in cols)


# ------------------------------------------------------------------ N-rank pipeline simulation
def bucket_to_rank(bucket: int, group_size: int) -> int:
    """src/distributed_join.cpp:247-266: bucket b*G + i of batch b goes to group-local rank i."""
    return bucket % group_size


def simulate_distributed_inner_join(lefts, rights, odf=1, seed=SEED_NVLINK):
    """lefts/rights: per-rank (keys, payload).  Returns per-rank list of 4-column outputs,
    following src/distributed_join.cpp:211-339 (partition into G*odf buckets, batch b's bucket
    b*G+i to rank i, per-batch local join, concatenate)."""
    G = len(lefts)
    if G == 1:
        n, cols = inner_join(*lefts[0], *rights[0])
        return [cols]
    nparts = G * odf
    parts_l = [hash_partition(k, p, nparts, seed) for k, p in lefts]
    parts_r = [hash_partition(k, p, nparts, seed) for k, p in rights]
    results = []
    for dst in range(G):
        outs = []
        for b in range(odf):
            q = b * G + dst

            def gather(parts):
                ks = [pk[off[q]:off[q + 1]] for pk, _, off in parts]
                ps = [pp[off[q]:off[q + 1]] for _, pp, off in parts]
                return np.concatenate(ks), np.concatenate(ps)

            lk, lp = gather(parts_l)
            rk, rp = gather(parts_r)
            _, cols = inner_join(lk, lp, rk, rp)
            outs.append(cols)
        results.append(tuple(np.concatenate([o[c] for o in outs]) for c in range(4)))
    return results
