"""The multi-GPU plumbing (query partition, tree broadcast, CSR gather) exercised on CPU with
world_size 2 over gloo.  The per-rank "engine" here is the oracle -- in a test, as the checker
stand-in; on GPUs each rank calls rt_query_boxes on its shard instead (bench.py --gpus N)."""
import os
import socket
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def test_shard_range_and_concat(rtb):
    sh = rtb.sharding
    for n in (0, 1, 7, 8, 1000, 1001):
        for world in (1, 2, 3, 8):
            ranges = [sh.shard_range(n, world, r) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            sizes = [e - b for b, e in ranges]
            assert max(sizes) <= -(-n // world) if n else True
    parts = [(np.array([0, 2, 2], np.uint64), np.array([5, 9], np.uint32)),
             (np.array([0, 1], np.uint64), np.array([3], np.uint32)),
             (np.array([0], np.uint64), np.zeros(0, np.uint32))]
    off, ids = sh.concat_csr(parts)
    assert off.tolist() == [0, 2, 2, 3] and ids.tolist() == [5, 9, 3]


def _worker(rank, world, port, out_path):
    import importlib
    sys.path.insert(0, os.path.dirname(HERE))
    sys.path.insert(0, HERE)
    import torch.distributed as dist
    import oracles as O
    rtb = importlib.import_module("r-star-tree_b200")
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    n, nq = 4000, 1501
    queries = rtb.workloads.cube_queries(nq, n, seed=3)
    # rank 0 builds the tree; the image travels by ONE broadcast
    image = None
    if rank == 0:
        tree = rtb.RTree(4).insert(rtb.workloads.uniform_points(n, seed=2))
        image = tree.flatten_device()
        blocks = image.blocks()
        meta = [tree.flatten()]
    else:
        blocks, meta = None, [None]
    got = rtb.sharding.broadcast_bytes(blocks, 0)
    dist.broadcast_object_list(meta, 0)
    flat = meta[0]
    if rank == 0:
        assert np.array_equal(got, blocks)
    assert got.dtype == np.uint8 and got.nbytes > 0
    # every rank answers its contiguous shard
    begin, end = rtb.sharding.shard_range(nq, world, rank)
    F = O.Flat(O.F32, 3, False, 8, flat["leaf_level"], 0, flat["nodes"], flat["children_bound"],
               flat["children"], flat["data"])
    off, ids = O.oracle_query_boxes(F, queries[begin:end])
    merged = rtb.sharding.gather_csr(off, ids, 0)
    # per-rank parity as bench.py does it: every rank answers a batch of ITS OWN (seeded by its rank),
    # sends the CSR of a prefix, and rank 0 regenerates each rank's queries and checks them all
    def my_batch(r, count):
        return rtb.workloads.cube_queries(count, n, seed=43 + 1000 * r)
    mine = O.oracle_query_boxes(F, my_batch(rank, 300))
    if rank == 1:
        mine[1][5] ^= 1          # rank 1 "computes" one wrong id: it must be flagged, rank 0 must not
    parts = [None] * world
    dist.all_gather_object(parts, (200, mine[0][:201], mine[1][:int(mine[0][200])]))
    if rank == 0:
        full = O.oracle_query_boxes(F, queries)
        ok = np.array_equal(merged[0], full[0]) and np.array_equal(merged[1], full[1])
        verdict = rtb.sharding.check_rank_prefixes(parts, lambda q: O.oracle_query_boxes(F, q), my_batch)
        flags = [int(v["bit_exact"]) for v in verdict]
        np.save(out_path, np.array([int(ok), int(got.nbytes), len(full[1])] + flags))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_query_sharding_over_gloo(tmp_path):
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "result.npy")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    ok, nbytes, hits, rank0_exact, rank1_exact = np.load(out)
    assert ok == 1 and nbytes > 0 and hits > 0
    assert rank0_exact == 1 and rank1_exact == 0     # the corrupted rank is caught, the clean one passes
