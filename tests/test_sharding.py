"""CPU tests of the multi-GPU plumbing: tile sharding across ranks and the final gather, with
torch.distributed (gloo, world_size 2).  The data path has no collective (tiles are independent,
SURVEY 8e); only the result gather communicates."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import fpm_testlib as T

sys.path.insert(0, os.path.join(T.ROOT, "fpm-opencv_b200"))
import sharding  # noqa: E402


def test_shard_ranges_cover_all_tiles_once():
    for n in (1, 2, 7, 148, 320, 321, 1184):
        for w in (1, 2, 3, 4, 8):
            seen = []
            for r in range(w):
                a, b = sharding.shard_range(n, r, w)
                assert 0 <= a <= b <= n
                seen += list(range(a, b))
            assert seen == list(range(n))
            sizes = [sharding.shard_range(n, r, w)[1] - sharding.shard_range(n, r, w)[0] for r in range(w)]
            assert max(sizes) - min(sizes) <= 1


def test_tile_grid_covers_frame():
    tiles = sharding.tile_grid(2560, 2160, 128)          # BASELINE config 4: 20 x 16 = 320 tiles
    assert len(tiles) == 320 and tiles[0] == (0, 0) and tiles[-1] == (2432, 1920)
    assert len(set(tiles)) == 320


def _worker(rank, world, port, n_tiles, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a, b = sharding.shard_range(n_tiles, rank, world)
    # stand-in for the per-tile reconstruction: a deterministic function of the tile index only
    local = torch.stack([torch.full((4, 4), float(t)) + torch.arange(16.).reshape(4, 4) for t in range(a, b)]) if b > a \
        else torch.zeros((0, 4, 4))
    full = sharding.gather_tiles(local, n_tiles, rank, world)
    tmax = sharding.max_over_ranks(float(rank + 1))
    if rank == 0:
        np.save(out, full.numpy())
        assert tmax == float(world)
    else:
        assert full is None
    dist.destroy_process_group()


@pytest.mark.parametrize("n_tiles", [5, 8])
def test_gather_over_gloo_world2(tmp_path, n_tiles):
    out = str(tmp_path / "g.npy")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, n_tiles, out), nprocs=2, join=True)
    full = np.load(out)
    want = np.stack([np.full((4, 4), float(t)) + np.arange(16.).reshape(4, 4) for t in range(n_tiles)])
    assert np.array_equal(full, want)
