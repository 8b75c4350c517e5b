"""CPU, world_size 2, gloo: the tile sharding and the end-of-batch histogram gather (the only
collective of the path).  The per-tile work is stood in for by the oracle's label histogram
-- the test covers the host-side N>1 logic, not the kernels."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from urban_pointcloud_processing_b200 import batch


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close()
    return p


def _tile_hist(i):
    rng = np.random.default_rng(1000 + i)
    labels = rng.choice(np.array([0, 1, 9, 10, 40, 99]), size=500 + 37 * i)
    return torch.from_numpy(np.bincount(labels, minlength=256).astype(np.int64))


def _worker(rank, world, port, n_tiles, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        hist, ids = batch.run_batch(n_tiles, _tile_hist, torch.device('cpu'))
        out[rank] = (hist.numpy().copy(), ids)
    finally:
        dist.destroy_process_group()


def test_shard_tiles_partitions_the_batch():
    for n, w in ((7, 2), (512, 8), (3, 4), (0, 2)):
        shards = [batch.shard_tiles(n, r, w) for r in range(w)]
        assert sorted(sum(shards, [])) == list(range(n))
        assert max(len(s) for s in shards) - min(len(s) for s in shards) <= 1


def test_two_rank_batch_equals_single_process():
    n_tiles = 7
    want = np.stack([_tile_hist(i).numpy() for i in range(n_tiles)])
    single, ids = batch.run_batch(n_tiles, _tile_hist, torch.device('cpu'), rank=0, world_size=1)
    assert np.array_equal(single.numpy(), want) and ids == list(range(n_tiles))
    mgr = mp.Manager()
    out = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, n_tiles, out), nprocs=2, join=True)
    assert set(out.keys()) == {0, 1}
    assert out[0][1] == [0, 2, 4, 6] and out[1][1] == [1, 3, 5]
    for r in (0, 1):
        assert np.array_equal(out[r][0], want)
