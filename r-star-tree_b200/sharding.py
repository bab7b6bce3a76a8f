"""Multi-GPU plumbing for the query path.

Queries are independent and the tree is read-only, so the path shards by QUERY:
the tree is replicated on every GPU (one broadcast of the block buffer), rank r answers
the contiguous query range ``shard_range(n, world, r)``, and per-rank CSR results
concatenate in rank order after rebasing the offsets.  There is no data-path collective.

The functions here are pure host logic (numpy + an optional torch.distributed process
group for the broadcast / gather), so they are tested on CPU with the gloo backend.
"""
import numpy as np


def shard_range(n, world_size, rank):
    """Contiguous, even split: query i goes to rank i // ceil(n / world_size)."""
    per = -(-n // world_size) if world_size > 0 else n
    begin = min(n, rank * per)
    end = min(n, begin + per)
    return begin, end


def concat_csr(parts):
    """[(offsets_r, ids_r)] in rank order -> one (offsets, ids) over all queries."""
    offsets = [np.zeros(1, np.uint64)]
    ids = []
    base = np.uint64(0)
    for off, idv in parts:
        off = np.asarray(off, np.uint64)
        offsets.append(off[1:] + base)
        ids.append(np.asarray(idv, np.uint32))
        base = base + off[-1]
    return np.concatenate(offsets), (np.concatenate(ids) if ids else np.zeros(0, np.uint32))


def broadcast_bytes(buf, src, group=None, device=None):
    """Broadcast a uint8 numpy buffer (the tree image) from rank `src` to every rank.

    With a CUDA `device` the payload travels GPU to GPU (NCCL over NVLink); on CPU groups
    (gloo) it travels as a host tensor.  Returns the received numpy array (host) or the
    torch tensor (device)."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank(group)
    n = torch.tensor([0 if buf is None else int(buf.nbytes)], dtype=torch.int64,
                     device=device if device is not None else "cpu")
    if rank != src:
        n.zero_()
    dist.broadcast(n, src, group=group)
    nbytes = int(n.item())
    if rank == src:
        t = torch.from_numpy(np.ascontiguousarray(buf).view(np.uint8).reshape(-1))
        if device is not None:
            t = t.to(device)
    else:
        t = torch.empty(nbytes, dtype=torch.uint8, device=device if device is not None else "cpu")
    dist.broadcast(t, src, group=group)
    return t if device is not None else t.numpy()


def gather_csr(offsets, ids, dst=0, group=None):
    """Gather per-rank CSR results on rank `dst` (host tensors) and concatenate them in rank
    order.  Returns (offsets, ids) on dst, None elsewhere."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    parts = [None] * world if rank == dst else None
    dist.gather_object((np.asarray(offsets), np.asarray(ids)), parts, dst=dst, group=group)
    if rank != dst:
        return None
    return concat_csr(parts)


def check_rank_prefixes(parts, reference_search, regenerate):
    """Per-rank parity of a sharded run, on the rank that holds the checker.

    parts[r] = (count, offsets[count + 1], ids) is what rank r's engine answered for the first
    `count` queries of ITS batch (gathered with all_gather_object); `regenerate(r, count)` rebuilds
    those queries from rank r's seed, `reference_search(queries)` -> (offsets, ids) is the checker
    (the reference's RTree::search()).  Returns one dict per rank: bit_exact is true only when both
    the offsets and the ids are identical."""
    out = []
    for r, (count, off, ids) in enumerate(parts):
        want_off, want_ids = reference_search(regenerate(r, int(count)))
        ok = bool(np.array_equal(np.asarray(off), want_off) and np.array_equal(np.asarray(ids), want_ids))
        out.append({"rank": r, "checked_queries": int(count), "hits": int(len(want_ids)), "bit_exact": ok})
    return out
