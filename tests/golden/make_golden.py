"""Generates tests/golden/*.json.  Run in the build container: python tests/golden/make_golden.py

The reference (rapidsai/distributed-join) is GPU-only C++ on cuDF 0.19 / UCX / MPI and cannot be
built or imported here, so the golden vectors come from
  * an independent implementation of the published hash the path uses
    (sklearn.utils.murmurhash3_32 == MurmurHash3_x86_32, the algorithm behind
    cudf::hash_id::HASH_MURMUR3, call sites src/distributed_join.cpp:211-225), and
  * the analytical answers the reference's own tests assert
    (test/compare_against_analytical.cu:44-54,152,194-201: 3i JOIN 5j over [0,size) has size/5 rows,
     each with key%15==0, left payload key/3, right payload key/5).
"""
import json
import os

import numpy as np
from sklearn.utils import murmurhash3_32

HERE = os.path.dirname(os.path.abspath(__file__))


def murmur(key: int, seed: int) -> int:
    return int(murmurhash3_32(int(key).to_bytes(8, "little", signed=True), seed=seed, positive=True))


def main():
    rng = np.random.default_rng(20260921)
    keys = [0, 1, 2, 42, -1, 799999999, (1 << 40) + 7, -(1 << 63), (1 << 63) - 1]
    keys += [int(k) for k in rng.integers(-(1 << 63), (1 << 63) - 1, 64, dtype=np.int64)]
    seeds = [12345678, 87654321, 0, 1]
    kat = {str(s): {str(k): murmur(k, s) for k in keys} for s in seeds}
    # row hash / partition id under the cuDF 0.19 row_hasher assumption:
    part = {str(s): {str(k): {"row_hash": (murmur(k, s) + 0x9E3779B9) & 0xFFFFFFFF,
                              "p8": ((murmur(k, s) + 0x9E3779B9) & 0xFFFFFFFF) % 8,
                              "p32": ((murmur(k, s) + 0x9E3779B9) & 0xFFFFFFFF) % 32,
                              "p7": ((murmur(k, s) + 0x9E3779B9) & 0xFFFFFFFF) % 7}
                     for k in keys[:9]} for s in seeds[:2]}
    with open(os.path.join(HERE, "murmur3_kat.json"), "w") as f:
        json.dump({"source": "sklearn.utils.murmurhash3_32 over 8 LE key bytes", "murmur3": kat,
                   "partition_assumed": part}, f, indent=1)
    analytical = {"source": "test/compare_against_analytical.cu:152,194-201",
                  "cases": [{"size": s, "odf": o, "nvl": n, "rows": s // 5}
                            for s, o, n in [(30000, 1, 1), (300000, 1, 1), (300000, 4, 1), (3000000, 1, 1),
                                            (3000000, 4, 1), (3000000, 4, 2)]]}
    with open(os.path.join(HERE, "analytical.json"), "w") as f:
        json.dump(analytical, f, indent=1)


if __name__ == "__main__":
    main()
