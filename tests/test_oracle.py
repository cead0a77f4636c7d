"""CPU tests: the oracle (oracle/dj_oracle.c + numpy restatement) against the golden vectors and
the analytical answers the reference's own tests assert (G1, G3, G4, G5)."""
import json
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_murmur3_known_answers(oracle):
    kat = json.load(open(os.path.join(GOLD, "murmur3_kat.json")))
    for seed, table in kat["murmur3"].items():
        keys = np.array([int(k) for k in table], dtype=np.int64)
        want = np.array(list(table.values()), dtype=np.uint32)
        got_c = np.array([oracle.murmur3_i64(int(k), int(seed)) for k in keys], dtype=np.uint32)
        assert (got_c == want).all()
        assert (oracle.np_murmur3_i64(keys, int(seed)) == want).all()


def test_survey_known_answers(oracle):
    # SURVEY.md 8(c) self-generated KATs (murmur part verified against sklearn there)
    assert oracle.murmur3_i64(0, 12345678) == 0x1B5A1A0B
    assert oracle.murmur3_i64(-1, 12345678) == 0x36E722EF
    assert oracle.murmur3_i64(799999999, 12345678) == 0x6C2631A9
    assert oracle.murmur3_i64(42, 87654321) == 0xD0985B41
    assert oracle.murmur3_i64(-1, 0) == 0x627564E8
    assert oracle.row_hash_i64(0, 12345678) == 0xB99193C4
    assert oracle.row_hash_i64(42, 12345678) == 0xE8FD1B95


def test_partition_ids_match_fixture_and_numpy(oracle):
    kat = json.load(open(os.path.join(GOLD, "murmur3_kat.json")))
    for seed, table in kat["partition_assumed"].items():
        keys = np.array([int(k) for k in table], dtype=np.int64)
        for nparts, name in ((8, "p8"), (32, "p32"), (7, "p7")):
            want = np.array([v[name] for v in table.values()], dtype=np.int32)
            assert (oracle.partition_ids(keys, int(seed), nparts) == want).all()
            assert (oracle.np_partition_ids(keys, int(seed), nparts) == want).all()
    rng = np.random.default_rng(3)
    keys = rng.integers(-(1 << 63), (1 << 63) - 1, 100_000, dtype=np.int64)
    for nparts in (2, 3, 8, 32, 100):
        assert (oracle.partition_ids(keys, 12345678, nparts) == oracle.np_partition_ids(keys, 12345678, nparts)).all()


def test_hash_partition_is_stable_counting_sort(oracle):
    rng = np.random.default_rng(4)
    keys = rng.integers(0, 1 << 40, 50_000, dtype=np.int64)
    pay = np.arange(keys.size, dtype=np.int64)
    ko, po, off = oracle.hash_partition(keys, pay, 8, oracle.SEED_NVLINK)
    ids = oracle.partition_ids(keys, oracle.SEED_NVLINK, 8)
    order = np.argsort(ids, kind="stable")
    assert (ko == keys[order]).all() and (po == pay[order]).all()
    assert off[0] == 0 and off[-1] == keys.size
    assert (np.diff(off) == np.bincount(ids, minlength=8)).all()


@pytest.mark.parametrize("case", json.load(open(os.path.join(GOLD, "analytical.json")))["cases"][:3])
def test_analytical_join_cardinality(oracle, case):
    """G1: keys 3i JOIN 5j, i,j in [0,size) -> size/5 rows, predicate of
    test/compare_against_analytical.cu:44-54."""
    size = case["size"]
    lk, lp = oracle.generate_analytical(3, 0, size)
    rk, rp = oracle.generate_analytical(5, 0, size)
    n, (c0, c1, c2, c3) = oracle.inner_join(lk, lp, rk, rp)
    assert n == case["rows"] == size // 5
    assert (c0 % 15 == 0).all() and (c1 == c0 // 3).all() and (c2 % 15 == 0).all()
    assert (c3 == c2 // 5).all() and (c0 == c2).all()


def test_c_join_equals_numpy_join_with_duplicates(oracle):
    rng = np.random.default_rng(5)
    lk = rng.integers(0, 2000, 5000, dtype=np.int64)
    rk = rng.integers(0, 2000, 7000, dtype=np.int64)
    lp, rp = np.arange(5000, dtype=np.int64), np.arange(7000, dtype=np.int64) + 10**9
    n, cols = oracle.inner_join(lk, lp, rk, rp)
    ref = oracle.np_inner_join(lk, lp, rk, rp)
    assert n == ref[0].size
    for a, b in zip(oracle.sort_rows(*cols), oracle.sort_rows(*ref)):
        assert (a == b).all()
    assert oracle.multiset_checksum4(*cols) == oracle.multiset_checksum4(*ref)
    # empty side -> empty result (src/distributed_join.cpp:76-82)
    e = np.empty(0, dtype=np.int64)
    assert oracle.inner_join(e, e, rk, rp)[0] == 0 and oracle.inner_join(lk, lp, e, e)[0] == 0


@pytest.mark.parametrize("unique", [True, False])
def test_generator_invariants(oracle, unique):
    """G5: unique build keys -> cardinality == #hit draws; miss keys are absent from build."""
    g = oracle.gen_params(100_000, 250_000, 0.3, 500_000, unique)
    bk, bp, _ = oracle.generate_rows(g, 0, 0, 0, g.nb)
    pk, pp, hits = oracle.generate_rows(g, 1, 0, 0, g.np)
    assert bk.min() >= 0 and bk.max() <= g.rand_max and pk.min() >= 0 and pk.max() <= g.rand_max
    assert (bp == np.arange(g.nb)).all() and (pp == np.arange(g.np)).all()
    assert abs(hits / g.np - 0.3) < 0.01
    n, _ = oracle.inner_join(bk, bp, pk, pp, count_only=True)
    in_build = np.isin(pk, bk)
    assert in_build.sum() == hits  # misses never collide with a build key
    if unique:
        assert np.unique(bk).size == bk.size and n == hits
    else:
        mult = np.bincount(bk, minlength=g.rand_max + 1)
        assert n == mult[pk].sum()
    # chunked generation == whole generation (counter-based)
    k2, p2, _ = oracle.generate_rows(g, 1, 0, 1000, 5000, oracle.build_bitmap(g, 0) if not unique else None)
    assert (k2 == pk[1000:6000]).all() and (p2 == pp[1000:6000]).all()


def test_generator_rank_offsets(oracle):
    """src/generate_table.cuh:192-202: key += rand_max*rank, payload += n_rank*rank."""
    g = oracle.gen_params(1000, 2000, 0.5, 4000, True)
    k0, p0, _ = oracle.generate_rows(g, 0, 0, 0, 1000)
    k3, p3, _ = oracle.generate_rows(g, 0, 3, 0, 1000)
    assert k3.min() >= 3 * 4000 and k3.max() <= 4 * 4000 and (p3 == p0 + 3000).all()
    (bk, bp), (pk, pp) = oracle.generate_tables_distributed(g, 1, 4)
    assert bk.size == 1000 and pk.size == 2000


def test_config1_two_rank_pipeline_matches_single_join(oracle):
    """BASELINE config 1 restated on CPU: 1M x 1M rows, selectivity 0.3, two ranks: partition ->
    exchange -> local join equals one global join (G3, test/compare_against_single_gpu.cu:163-205)."""
    world, n_rank = 2, 500_000
    g = oracle.gen_params(n_rank, n_rank, 0.3, 2 * n_rank, True)
    tabs = [oracle.generate_tables_distributed(g, r, world) for r in range(world)]
    lefts, rights = [t[0] for t in tabs], [t[1] for t in tabs]
    for odf in (1, 4):
        per_rank = oracle.simulate_distributed_inner_join(lefts, rights, odf=odf)
        got = tuple(np.concatenate([r[c] for r in per_rank]) for c in range(4))
        gl = (np.concatenate([l[0] for l in lefts]), np.concatenate([l[1] for l in lefts]))
        gr = (np.concatenate([r[0] for r in rights]), np.concatenate([r[1] for r in rights]))
        n, ref = oracle.inner_join(*gl, *gr)
        hits = sum(oracle.generate_rows(g, 1, s, 0, n_rank, materialize=False)[2] for s in range(world))
        assert got[0].size == n
        # per-rank key ranges [r*rand_max, (r+1)*rand_max] share their end points
        # (src/generate_table.cuh:192-202 adds rand_max*rank to keys in [0, rand_max]), so the
        # global cardinality may exceed the hit count by at most one match per boundary key copy
        assert 0 <= n - hits <= 4 * world
        assert oracle.multiset_checksum4(*got) == oracle.multiset_checksum4(*ref)
        # co-location (G4): every key of rank r's output hashes to r mod G
        for r, cols in enumerate(per_rank):
            assert (oracle.partition_ids(cols[0], oracle.SEED_NVLINK, world * odf) % world == r).all()


def test_omp_baseline_agrees(oracle):
    g = oracle.gen_params(200_000, 300_000, 0.3, 600_000, False)
    bk, bp, _ = oracle.generate_rows(g, 0, 0, 0, g.nb)
    pk, pp, _ = oracle.generate_rows(g, 1, 0, 0, g.np)
    r = oracle.partitioned_join_omp(bk, bp, pk, pp, nparts=64, checksum=True)
    n, cols = oracle.inner_join(bk, bp, pk, pp)
    assert r["n_out"] == n and r["checksum"] == oracle.multiset_checksum4(*cols)


def test_generate_global_tables_is_the_union_of_every_ranks_tables(oracle):
    """bench.py's parity oracle regenerates the GLOBAL tables in one go: they must be exactly the rows the
    ranks hold after generate_tables_distributed (src/generate_table.cuh:155-272), for unique and
    duplicate build keys, and their join must equal the join of the concatenated per-rank tables."""
    import numpy as np

    for unique, sel, world in ((True, 0.3, 4), (False, 0.9, 2)):
        g = oracle.gen_params(20_000, 30_000, sel, 60_000, unique)
        (bk, bp), (pk, pp), hits = oracle.generate_global_tables(g, world)
        tabs = [oracle.generate_tables_distributed(g, r, world) for r in range(world)]
        cat = [np.concatenate([t[i][j] for t in tabs]) for i in (0, 1) for j in (0, 1)]
        for got, want in zip((bk, bp, pk, pp), cat):
            assert (np.sort(got) == np.sort(want)).all()
        n_ref, ref = oracle.inner_join(*cat)
        r = oracle.partitioned_join_omp(bk, bp, pk, pp, nparts=64, checksum=True)
        assert r["n_out"] == n_ref and r["checksum"] == oracle.multiset_checksum4(*ref)
        if unique:
            assert n_ref == hits
