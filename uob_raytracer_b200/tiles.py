"""Row-tile partitioning of a frame across ranks and the frame gather.

Pixels are independent and the RNG is keyed on the frame-global pixel id, so rank g of N renders
rows [g*H/N, (g+1)*H/N) with frame-global y and the N tiles concatenate to exactly the 1-GPU
frame.  The tiles are gathered with ONE in-place all-gather per frame (NCCL over NVLink on GPUs;
gloo in the CPU tests): every rank passes the whole-frame buffer as output and its own tile — a
view into that same buffer — as input.
"""
from __future__ import annotations


def row_tile(height: int, world_size: int, rank: int) -> tuple[int, int]:
    """(row0, rows) of `rank`.  The all-gather needs equal tiles, so height % world_size must be 0
    (true for every BASELINE config: 1080/8, 2160/8, 4320/8)."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank {rank} of {world_size}")
    if height % world_size:
        raise ValueError(f"frame height {height} is not divisible by {world_size} ranks")
    rows = height // world_size
    return rank * rows, rows


def tile_view(frame_flat, width: int, height: int, world_size: int, rank: int):
    """The slice of a flat whole-frame tensor that `rank` owns."""
    row0, rows = row_tile(height, world_size, rank)
    return frame_flat[row0 * width:(row0 + rows) * width]


def gather_frame(frame_flat, width: int, height: int, dist, group=None):
    """In-place all-gather of the row tiles into frame_flat on every rank."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if world == 1:
        return frame_flat
    dist.all_gather_into_tensor(frame_flat, tile_view(frame_flat, width, height, world, rank), group=group)
    return frame_flat
