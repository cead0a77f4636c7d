"""Worker for tests/test_distributed_cpu.py: BASELINE config 0 -- radix hash-partition + local hash join of
1M build x 1M probe int64 rows (selectivity 0.3) through the C/numpy oracle with a world_size=2 `gloo`
exchange of the partitions (no GPU).  Mirrors src/distributed_join.cpp:211-339 step by step:
hash_partition (murmur3, seed 12345678) -> communicate_sizes -> per-column all-to-all by offsets ->
per-batch local inner join -> concatenation."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle as O  # noqa: E402


def exchange(col: np.ndarray, send_counts, recv_counts):
    out = torch.empty(int(sum(recv_counts)), dtype=torch.int64)
    dist.all_to_all_single(out, torch.from_numpy(np.ascontiguousarray(col)), [int(c) for c in recv_counts],
                           [int(c) for c in send_counts])
    return out.numpy()


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    total, odf = int(sys.argv[1]), int(sys.argv[2])
    n_rank = total // world
    g = O.gen_params(n_rank, n_rank, 0.3, 2 * n_rank, True)
    (lk, lp), (rk, rp) = O.generate_tables_distributed(g, rank, world)
    nparts = world * odf
    parts = [O.hash_partition(k, p, nparts, O.SEED_NVLINK) for k, p in ((lk, lp), (rk, rp))]
    outs = []
    for b in range(odf):
        recv = []
        for pk, pp, off in parts:
            send_counts = np.diff(off)[b * world:(b + 1) * world]
            # communicate_sizes (src/all_to_all_comm.cpp:54-111)
            rc = torch.empty(world, dtype=torch.int64)
            dist.all_to_all_single(rc, torch.from_numpy(send_counts.astype(np.int64)))
            lo, hi = off[b * world], off[(b + 1) * world]
            recv.append((exchange(pk[lo:hi], send_counts, rc.tolist()), exchange(pp[lo:hi], send_counts, rc.tolist())))
        _, cols = O.inner_join(*recv[0], *recv[1])
        outs.append(cols)
    mine = tuple(np.concatenate([o[c] for o in outs]) for c in range(4))
    # co-location (G4): all keys on this rank belong to this rank's buckets
    assert (O.partition_ids(mine[0], O.SEED_NVLINK, nparts) % world == rank).all()
    ck = O.multiset_checksum4(*mine)
    stats = torch.tensor([mine[0].size, ck[0] >> 32, ck[0] & 0xFFFFFFFF, ck[1] >> 32, ck[1] & 0xFFFFFFFF],
                         dtype=torch.int64)
    allst = [torch.zeros_like(stats) for _ in range(world)]
    dist.all_gather(allst, stats)
    if rank == 0:
        n_total = sum(int(s[0]) for s in allst)
        c0 = sum((int(s[1]) << 32) | int(s[2]) for s in allst) & 0xFFFFFFFFFFFFFFFF
        c1 = sum((int(s[3]) << 32) | int(s[4]) for s in allst) & 0xFFFFFFFFFFFFFFFF
        tabs = [O.generate_tables_distributed(g, r, world) for r in range(world)]
        gl = (np.concatenate([t[0][0] for t in tabs]), np.concatenate([t[0][1] for t in tabs]))
        gr = (np.concatenate([t[1][0] for t in tabs]), np.concatenate([t[1][1] for t in tabs]))
        n_ref, ref = O.inner_join(*gl, *gr)
        assert n_total == n_ref, (n_total, n_ref)
        assert (c0, c1) == O.multiset_checksum4(*ref)
        print(f"DIST_OK rows={n_total} world={world} odf={odf}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
