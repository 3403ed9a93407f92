"""Multi-GPU forest build: trees are the independent units (src/writer.rs:795), so tree t is
built by rank t mod world; the only collective on the data path is ONE broadcast of the staged
item buffer from the rank that did the host->device copy. A small all-gather of
per-tree node counts afterwards lets every rank number its nodes exactly as a single-GPU build
(and a 1-thread reference build) would.

`dist` is torch.distributed (NCCL on GPUs; the pure host logic below is exercised with gloo on CPU).
"""
import numpy as np


def shard_trees(n_trees, rank, world):
    """Global indices of the trees built by `rank`."""
    return list(range(rank, n_trees, world))


def node_id_bases(counts_all, first_free):
    """counts_all[t] = nodes of global tree t (root included). Non-root ids are handed out by one
    counter starting at first_free, last tree first, post-order inside a tree (the order a 1-thread
    rayon pool produces). Returns base[t] for every tree."""
    n_trees = len(counts_all)
    base = np.zeros(n_trees, dtype=np.uint64)
    counter = int(first_free)
    for t in range(n_trees - 1, -1, -1):
        base[t] = counter
        counter += int(counts_all[t]) - 1
    if counter > 2**32:
        raise OverflowError("Database full. Arroy cannot generate enough internal IDs for your items")
    return base


def gather_counts(dist, local_counts, n_trees, rank, world, device=None):
    """All-gather the per-tree node counts (tiny) into the global, tree-indexed array."""
    import torch
    mine = shard_trees(n_trees, rank, world)
    per_rank = (n_trees + world - 1) // world
    buf = torch.zeros(per_rank, dtype=torch.int64, device=device)
    buf[:len(mine)] = torch.as_tensor(np.asarray(local_counts, dtype=np.int64), device=device)
    if world > 1:
        out = [torch.zeros_like(buf) for _ in range(world)]
        dist.all_gather(out, buf)
    else:
        out = [buf]
    counts_all = np.zeros(n_trees, dtype=np.int64)
    for r in range(world):
        trees_r = shard_trees(n_trees, r, world)
        counts_all[trees_r] = out[r].cpu().numpy()[:len(trees_r)]
    return counts_all


def broadcast_items(dist, items_tensor, src=0):
    """The single data-path collective: NCCL broadcast of the item matrix over NVLink."""
    dist.broadcast(items_tensor, src=src)


def sharded_build(ctx, dist, rank, world, tree_seeds, root_ids, first_free, split_after=0, arena=None, device=None):
    """Build this rank's share of the forest on `ctx` (items already staged on every rank) and emit
    its nodes with globally consistent ids. Returns ({node id: bytes} or None with an arena, counts_all)."""
    n_trees = len(tree_seeds)
    mine = shard_trees(n_trees, rank, world)
    local_counts = ctx.build_trees_begin([tree_seeds[t] for t in mine], split_after)
    counts_all = gather_counts(dist, local_counts, n_trees, rank, world, device=device)
    base = node_id_bases(counts_all, first_free)
    nodes = ctx.build_trees_emit([root_ids[t] for t in mine], [base[t] for t in mine], arena=arena)
    return nodes, counts_all


class _DevView:
    """zero-copy torch view of a device buffer of the library"""

    def __init__(self, ptr, shape, typestr="<f4"):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (int(ptr), False), "version": 2}


def stage_and_broadcast(ctx, dist, rank, metric, dim, ids, leaf_ptrs, device, chunk_mb=256):
    """Item staging for a multi-process (one rank per GPU) build: rank 0 decodes + uploads its host leaf values chunk by
    chunk (arroy_b200_stage_begin / _rows / _end) and every chunk is broadcast straight out of / into the library's item
    buffers while the next one is crossing PCIe. Still ONE logical broadcast of the item buffer, issued in
    pieces so that only the last piece is not hidden behind the H2D copy. leaf_ptrs is only read on rank 0."""
    import torch
    n = len(ids)
    ctx.stage_begin(metric, dim, ids)
    (p_items, p_h0, p_h1), ld = ctx.device_ptrs()
    items = torch.as_tensor(_DevView(p_items, (max(n, 1), ld)), device=device)
    chunk_rows = max(1, (chunk_mb << 20) // (ld * 4))
    works = []
    for a in range(0, n, chunk_rows):
        b = min(n, a + chunk_rows)
        if rank == 0:
            ctx.stage_rows(a, leaf_ptrs[a:b])       # returns when the chunk is in rank 0's HBM
        works.append(dist.broadcast(items[a:b], src=0, async_op=True))
    if rank == 0:
        ctx.stage_end(headers_on_device=False)      # headers decoded from the leaf values
    if n:
        h0 = torch.as_tensor(_DevView(p_h0, (n,)), device=device)
        h1 = torch.as_tensor(_DevView(p_h1, (n,)), device=device)
        works.append(dist.broadcast(h0, src=0, async_op=True))
        works.append(dist.broadcast(h1, src=0, async_op=True))
    for w in works:
        w.wait()
    torch.cuda.synchronize()
    if rank != 0:
        ctx.stage_end(headers_on_device=True)
