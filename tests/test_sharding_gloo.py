"""world_size-2 (and 3) gloo runs of the host-side multi-GPU logic on CPU: the tile partition covers the frame exactly
once, one all-gather of equal-sized block arrays carries every rank's tiles, and rank 0 reassembles the frame.
The tile blocks are synthetic here (pixel value = a function of the pixel coordinates); on the GPU box the same
functions move blocks rendered by fjgpu_render_tiles_device (bench.py --gpus N)."""
import os
import socket
import sys

import numpy as np
import pytest

import scenekit as sk

sk.pkg()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _expected(xres, yres):
    yy, xx = np.mgrid[0:yres, 0:xres].astype(np.float32)
    return np.stack([xx, yy, xx * 1000 + yy, np.ones_like(xx)], -1)


def _worker(rank, world, port, xres, yres, tile, out):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import scenekit
    scenekit.pkg()
    from fujiyama_renderer_b200 import sharding
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    tiles = sharding.make_tiles(xres, yres, tile)
    mine = sharding.rank_tiles(tiles, rank, world)
    per = sharding.blocks_per_rank(len(tiles), world)
    exp = _expected(xres, yres)
    blocks = torch.zeros((per, tile, tile, 4), dtype=torch.float32)
    for k, (_, x0, y0, x1, y1) in enumerate(mine):
        blocks[k, : y1 - y0, : x1 - x0] = torch.from_numpy(exp[y0:y1, x0:x1])
    g = sharding.all_gather_blocks(blocks, world, dist)
    if rank == 0:
        frame = sharding.assemble_frame(g.numpy(), tiles, world, xres, yres)
        np.save(out, frame)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,res", [(2, (100, 70)), (3, (64, 64)), (2, (33, 31))])
def test_partition_gather_assemble(tmp_path, world, res):
    import torch.multiprocessing as mp
    out = str(tmp_path / "frame.npy")
    port = _free_port()
    mp.spawn(_worker, args=(world, port, res[0], res[1], 32, out), nprocs=world, join=True)
    frame = np.load(out)
    assert np.array_equal(frame, _expected(*res))


def test_partition_is_exact_and_matches_reference_tiler():
    from fujiyama_renderer_b200 import sharding
    tiles = sharding.make_tiles(1920, 1080, 32)
    assert len(tiles) == 60 * 34
    assert tiles == sk.make_tiles(1920, 1080, 32)
    for world in (1, 2, 4, 8):
        seen = sorted(t[0] for r in range(world) for t in sharding.rank_tiles(tiles, r, world))
        assert seen == list(range(len(tiles)))
        sizes = [len(sharding.rank_tiles(tiles, r, world)) for r in range(world)]
        assert max(sizes) - min(sizes) <= 1 and max(sizes) == sharding.blocks_per_rank(len(tiles), world)
    # ragged region
    t = sharding.make_tiles(100, 70, 32, region=(10, 5, 90, 64))
    assert t[0][1:] == (10, 5, 32, 32) and t[-1][1:] == (64, 32, 90, 64)
