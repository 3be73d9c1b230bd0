"""Multi-GPU sharding of a frame: tiles are the independent units of the path (each tile generates its own
samples incl. filter margins, traces, and resolves only its own pixels: src/fj_renderer.cc:1098-1121), so the
tile list is dealt round-robin to the ranks, every rank renders its tiles with the full scene replicated in its
HBM, and ONE all-gather of the packed tile blocks ends the frame.  No collective runs while
tracing.  Works on any torch.distributed backend (NCCL on the GPUs, gloo in the CPU tests).
"""
import numpy as np


def make_tiles(xres, yres, tile=32, region=None):
    """Tiler::GenerateTiles (src/fj_tiler.cc:56-113): row-major (id, xmin, ymin, xmax, ymax) clipped to the region."""
    x0, y0, x1, y1 = region if region else (0, 0, xres, yres)
    X0, Y0 = max(0, x0) // tile, max(0, y0) // tile
    X1, Y1 = -(-min(xres, x1) // tile), -(-min(yres, y1) // tile)
    out = []
    for y in range(Y0, Y1):
        for x in range(X0, X1):
            out.append((len(out), max(x * tile, x0), max(y * tile, y0), min((x + 1) * tile, x1), min((y + 1) * tile, y1)))
    return out


def rank_tiles(tiles, rank, world):
    """Tiles of `rank`: index i goes to rank i % world (interleaved: cheap static balance against spatially
    clustered cost — the same rule libfjscene's fjscene_set_device applies)."""
    return tiles[rank::world]


def blocks_per_rank(ntiles, world):
    """Every rank contributes the same number of blocks to the all-gather (short ranks pad with zero blocks)."""
    return -(-ntiles // world)


def all_gather_blocks(local_blocks, world, dist=None):
    """local_blocks: [blocks_per_rank, th, tw, 4] tensor (device or CPU).  Returns [world * blocks_per_rank, th, tw, 4]."""
    import torch
    if world == 1:
        return local_blocks
    out = torch.empty((world * local_blocks.shape[0],) + tuple(local_blocks.shape[1:]), dtype=local_blocks.dtype,
                      device=local_blocks.device)
    dist.all_gather_into_tensor(out, local_blocks.contiguous())
    return out


def assemble_frame(gathered, tiles, world, xres, yres):
    """Un-permutes the gathered blocks ([world * blocks_per_rank, th, tw, 4], numpy) into the row-major RGBA frame."""
    per = blocks_per_rank(len(tiles), world)
    frame = np.zeros((yres, xres, 4), np.float32)
    for i, (_, x0, y0, x1, y1) in enumerate(tiles):
        r, k = i % world, i // world
        frame[y0:y1, x0:x1] = gathered[r * per + k, : y1 - y0, : x1 - x0]
    return frame
