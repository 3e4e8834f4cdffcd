"""N > 1 host logic on CPU: world_size-2 gloo run of the sharding + gather plumbing.  Each rank evaluates its contiguous
shard of an edge batch (with the CPU oracle standing in for the device call) and the shards are all-gathered; the result must
equal the unsharded evaluation."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from po_rrt_b200 import shard, synth


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 32, 1000003):
        for world in (1, 2, 3, 8):
            r = [shard.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n_edges, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import pyoracle as O
    occ, zones = synth.door_map(size=256, n_zones=3, seed=1)
    omap = O.GridMap(occ, zones, [-1.0, -1.0], [1.0, 1.0], O.DOOR, 0.3)
    a, b = synth.edges(n_edges, seed=2, max_len=0.2)
    lo, hi = shard.shard_range(n_edges, rank, world)
    local = torch.from_numpy(omap.edge_validity(a[lo:hi], b[lo:hi]))        # this rank's slice only
    full = shard.all_gather_shards(local, n_edges)
    # the 128-byte ncclUniqueId travels from rank 0 to everybody (shard.init_comm's host-side step); rank 0 really asks the library
    from po_rrt_b200 import _lib
    import ctypes as C

    def make_id():
        buf = np.zeros(128, np.uint8)
        assert _lib.load().porrt_comm_unique_id(buf.ctypes.data_as(C.c_void_p)) == 0
        return buf.tolist()
    uid = shard.exchange_unique_id(make_id)
    assert len(uid) == 128
    gathered = [None] * world
    dist.all_gather_object(gathered, uid)
    assert all(g == gathered[0] for g in gathered)
    lo_c, hi_c = C.c_int64(), C.c_int64()
    assert _lib.load().porrt_shard_range(n_edges, rank, world, C.byref(lo_c), C.byref(hi_c)) == 0
    assert (lo_c.value, hi_c.value) == (lo, hi)          # the library shards exactly like shard.py
    t = shard.max_over_ranks(1.0 + rank)
    if rank == 0:
        np.save(os.path.join(out_dir, "gathered.npy"), full.numpy())
        np.save(os.path.join(out_dir, "tmax.npy"), np.array([t]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_edges_gloo(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    n_edges = 20_001                     # odd: ragged shards
    mp.spawn(_worker, args=(2, port, n_edges, str(tmp_path)), nprocs=2, join=True)
    from oracle import pyoracle as O
    occ, zones = synth.door_map(size=256, n_zones=3, seed=1)
    omap = O.GridMap(occ, zones, [-1.0, -1.0], [1.0, 1.0], O.DOOR, 0.3)
    a, b = synth.edges(n_edges, seed=2, max_len=0.2)
    np.testing.assert_array_equal(np.load(tmp_path / "gathered.npy"), omap.edge_validity(a, b))
    assert np.load(tmp_path / "tmax.npy")[0] == 2.0
