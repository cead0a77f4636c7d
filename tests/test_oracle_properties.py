"""Property tests of the CPU oracle (hypothesis): the C restatement against the independent numpy
restatement on arbitrary small tables -- duplicates, negative keys, empty sides, any partition count."""
import numpy as np
from hypothesis import given, settings
from hypothesis import strategies as st

keys = st.lists(st.integers(min_value=-(2**63), max_value=2**63 - 1), max_size=60)
small_keys = st.lists(st.integers(min_value=-5, max_value=5), max_size=60)


@settings(max_examples=60, deadline=None)
@given(lk=small_keys, rk=small_keys)
def test_inner_join_is_the_multiset_of_equal_key_pairs(oracle, lk, rk):
    lk, rk = np.array(lk, dtype=np.int64), np.array(rk, dtype=np.int64)
    lp, rp = np.arange(lk.size, dtype=np.int64), np.arange(rk.size, dtype=np.int64) + 1000
    n, cols = oracle.inner_join(lk, lp, rk, rp)
    want = sorted((int(a), i, int(b), 1000 + j) for i, a in enumerate(lk) for j, b in enumerate(rk) if a == b)
    got = sorted(zip(*[c.tolist() for c in cols])) if n else []
    assert n == len(want) and got == want
    ref = oracle.np_inner_join(lk, lp, rk, rp)
    assert oracle.multiset_checksum4(*cols) == oracle.multiset_checksum4(*ref)


@settings(max_examples=60, deadline=None)
@given(ks=keys, nparts=st.integers(min_value=1, max_value=40), seed=st.sampled_from([0, 1, 12345678, 87654321]))
def test_hash_partition_is_a_permutation_grouped_by_partition_id(oracle, ks, nparts, seed):
    ks = np.array(ks, dtype=np.int64)
    pay = np.arange(ks.size, dtype=np.int64)
    ko, po, off = oracle.hash_partition(ks, pay, nparts, seed)
    assert off[0] == 0 and off[-1] == ks.size and (np.diff(off) >= 0).all()
    assert sorted(po.tolist()) == pay.tolist() and (ko == ks[po]).all()  # rows stay intact
    ids = oracle.partition_ids(ko, seed, nparts)
    assert (ids == oracle.np_partition_ids(ko, seed, nparts)).all()
    for p in range(nparts):
        assert (ids[off[p]:off[p + 1]] == p).all()


@settings(max_examples=30, deadline=None)
@given(ranks=st.integers(min_value=1, max_value=4), odf=st.integers(min_value=1, max_value=3),
       data=st.lists(st.tuples(st.integers(min_value=0, max_value=30), st.integers(min_value=0, max_value=30)),
                     max_size=80))
def test_partition_exchange_join_equals_global_join(oracle, ranks, odf, data):
    """The N-rank pipeline (partition -> exchange -> per-batch join, src/distributed_join.cpp:211-339) returns the
    same row multiset as one global join, for any distribution of rows over ranks."""
    lk = np.array([d[0] for d in data], dtype=np.int64)
    rk = np.array([d[1] for d in data], dtype=np.int64)
    lp, rp = np.arange(lk.size, dtype=np.int64), np.arange(rk.size, dtype=np.int64) + 500
    cut = lambda a: np.array_split(a, ranks)
    lefts, rights = list(zip(cut(lk), cut(lp))), list(zip(cut(rk), cut(rp)))
    per_rank = oracle.simulate_distributed_inner_join(lefts, rights, odf=odf)
    got = tuple(np.concatenate([r[c] for r in per_rank]) for c in range(4))
    n_ref, ref = oracle.inner_join(lk, lp, rk, rp)
    assert got[0].size == n_ref
    assert oracle.multiset_checksum4(*got) == oracle.multiset_checksum4(*ref)
