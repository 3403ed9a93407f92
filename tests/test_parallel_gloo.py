"""World-size-2 gloo test (CPU) of the multi-GPU host logic: tree sharding, count exchange and
node-id bases must reproduce the single-process numbering."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from arroy_b200 import parallel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_trees, counts_all, first_free, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = parallel.shard_trees(n_trees, rank, world)
    local = [counts_all[t] for t in mine]
    got = parallel.gather_counts(dist, local, n_trees, rank, world)
    base = parallel.node_id_bases(got, first_free)
    # the single data-path collective, on a small stand-in buffer
    items = torch.arange(12, dtype=torch.float32).reshape(3, 4) if rank == 0 else torch.zeros(3, 4)
    parallel.broadcast_items(dist, items, src=0)
    np.save(os.path.join(out_dir, "r%d.npy" % rank), np.concatenate([got, base.astype(np.int64), items.numpy().astype(np.int64).ravel()]))
    dist.destroy_process_group()


def test_sharding_and_id_bases_world2(tmp_path):
    n_trees, first_free = 7, 7
    counts_all = [5, 9, 3, 11, 7, 1, 13]
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n_trees, counts_all, first_free, str(tmp_path)), nprocs=world, join=True)
    want_base = parallel.node_id_bases(counts_all, first_free)
    # single-process reference numbering: last tree first, non-root nodes only
    counter = first_free
    for t in range(n_trees - 1, -1, -1):
        assert want_base[t] == counter
        counter += counts_all[t] - 1
    for r in range(world):
        a = np.load(os.path.join(str(tmp_path), "r%d.npy" % r))
        assert a[:n_trees].tolist() == counts_all
        assert a[n_trees:2 * n_trees].tolist() == want_base.astype(np.int64).tolist()
        assert a[2 * n_trees:].tolist() == list(range(12))
    # shards partition the trees
    all_t = sorted(parallel.shard_trees(n_trees, 0, 2) + parallel.shard_trees(n_trees, 1, 2))
    assert all_t == list(range(n_trees))


def test_id_ranges_do_not_overlap():
    counts = [4, 1, 6, 2]
    base = parallel.node_id_bases(counts, 4)
    used = set(range(4))  # roots
    for t, c in enumerate(counts):
        ids = set(int(base[t]) + li for li in range(c - 1))
        assert not (ids & used)
        used |= ids
    assert used == set(range(4 + sum(c - 1 for c in counts)))
