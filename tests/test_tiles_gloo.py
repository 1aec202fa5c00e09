"""Multi-rank host logic on CPU: world_size-2 (and 4) gloo process groups.  Each rank produces its
row tile (with the oracle as a stand-in renderer: this is a test of the partition + gather plumbing,
not of the CUDA kernel) and the in-place all-gather must rebuild the 1-rank frame bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, W, H, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import bind as ob
        from uob_raytracer_b200 import host, tiles
        scene = host.load_test_model()
        cam = host.Camera()
        A, S, B = 2, 4, 3
        f = host.fitted_focal(A, H)
        row0, rows = tiles.row_tile(H, world, rank)
        part, _ = ob.oracle_render(W, H, A, S, B, f, scene.verts, scene.normals, scene.colors, cam.rot(), cam.position,
                                   cam.light, y0=row0, y1=row0 + rows, threads=1)
        frame = torch.zeros(W * H, dtype=torch.int32)
        tiles.tile_view(frame, W, H, world, rank).copy_(torch.from_numpy(part[row0:row0 + rows].view(np.int32).reshape(-1)))
        tiles.gather_frame(frame, W, H, dist)
        np.save(os.path.join(out_dir, f"rank{rank}.npy"), frame.numpy().view(np.uint32).reshape(H, W))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_row_tiles_allgather_gloo(tmp_path, ob, golden_scene, world):
    W, H = 96, 64
    mp.spawn(_worker, args=(world, _free_port(), W, H, str(tmp_path)), nprocs=world, join=True)
    v, n, c = golden_scene
    want, _ = ob.oracle_render(W, H, 2, 4, 3, 1100.0 * 2 * H / 1024, v, n, c, ob.oracle_rot_matrix(0, 0), [0, 0, -3.2],
                               [0, -0.5, -0.7])
    for r in range(world):
        got = np.load(tmp_path / f"rank{r}.npy")
        assert (got == want).all(), f"rank {r}: gathered frame differs from the 1-rank frame"


def test_row_tile_math():
    from uob_raytracer_b200 import tiles
    assert [tiles.row_tile(1080, 8, g) for g in (0, 7)] == [(0, 135), (945, 135)]
    assert tiles.row_tile(4320, 1, 0) == (0, 4320)
    with pytest.raises(ValueError):
        tiles.row_tile(1080, 7, 0)
    with pytest.raises(ValueError):
        tiles.row_tile(1080, 4, 4)
