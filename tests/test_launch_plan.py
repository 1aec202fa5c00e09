"""CPU tests of the host-side launch planning in libuob_rt.so (pure functions, no device): the visible rectangle that lets
tiles outside the projected scene box skip everything, and the tile lists of a mixed launch (ordinary 16x16 tiles +
8x8 sub-tiles of the tiles that see a sphere).  Checked against brute-force numpy geometry with the reference's own
camera model (kernels.cl:384-400: ray through sub-pixel (vx, vy) = R (vx, vy, f); vx = x*A + dx - W*A/2)."""
import ctypes

import numpy as np
import pytest

import uob_raytracer_b200 as u
from uob_raytracer_b200._lib import RtConfig, c_float_p, rt_lib

SPHERES = [((0.3, 0.1, -0.5), 0.075), ((-0.4, 0.8, -0.5), 0.05)]  # kernels.cl:7-10 (centre, radius^2)


def _cfg(W, H, A, row0=0, rows=0):
    return RtConfig(W, H, A, 8, 10, 0, row0, rows, 0, 0, 0)


def _fp(a):
    return np.ascontiguousarray(a, np.float32).ctypes.data_as(c_float_p)


def _rays(W, H, A, rot, focal, xs, ys, dx=0, dy=0):
    """Un-normalised primary ray directions of pixels (xs, ys), sub-sample (dx, dy)."""
    R = np.asarray(rot, np.float64).reshape(3, 4)[:, :3]
    v = np.stack([xs * A + dx - W * A / 2.0, ys * A + dy - H * A / 2.0, np.full(xs.shape, float(focal))], -1)
    return v @ R.T


def _visible_rect(cfg, lo, hi, rot, cam, focal):
    rect = (ctypes.c_int * 4)()
    rc = rt_lib().rt_debug_visible_rect(ctypes.byref(cfg), _fp(lo), _fp(hi), _fp(rot), _fp(list(cam) + [0.0]), float(focal), rect)
    assert rc == 0
    return list(rect)


def _tile_lists(cfg, rot, cam, focal):
    cap = ((cfg.width + 7) // 8) * ((cfg.height + 7) // 8) + 16
    tiles = (ctypes.c_int * cap)()
    nl, ns = ctypes.c_int(0), ctypes.c_int(0)
    rc = rt_lib().rt_debug_tile_lists(ctypes.byref(cfg), _fp(rot), _fp(list(cam) + [0.0]), float(focal), tiles, cap, ctypes.byref(nl),
                                      ctypes.byref(ns))
    assert rc == 0
    t = np.array(tiles[:nl.value + ns.value])
    return t[:nl.value], t[nl.value:]


CAMERAS = [  # (W, H, A, yaw, pitch, cam)
    (1920, 1080, 2, 0.0, 0.0, (0.0, 0.0, -3.2)),
    (1024, 1024, 2, 0.0, 0.0, (0.0, 0.0, -3.2)),
    (3840, 2160, 4, 0.0, 0.0, (0.0, 0.0, -3.2)),
    (480, 270, 2, 0.3, -0.2, (0.4, -0.3, -3.0)),
    (333, 205, 2, -0.5, 0.35, (-0.5, 0.2, -2.6)),
]


@pytest.mark.parametrize("W,H,A,yaw,pitch,cam", CAMERAS)
def test_visible_rectangle_contains_every_ray_that_can_hit_the_box(W, H, A, yaw, pitch, cam):
    rot = u.rot_matrix(yaw, pitch)
    focal = 1100.0 * A * H / 1024
    lo, hi = (-1.0, -1.0, -1.0), (1.0, 1.0, 1.0)
    x0, y0, x1, y1 = _visible_rect(_cfg(W, H, A), lo, hi, rot, cam, focal)
    assert 0 <= x0 <= x1 <= W and 0 <= y0 <= y1 <= H
    # slab test of every pixel's corner sub-samples against the box, in float64
    ys, xs = np.mgrid[0:H, 0:W]
    hit = np.zeros((H, W), bool)
    o = np.asarray(cam, np.float64)
    for dx, dy in ((0, 0), (A - 1, 0), (0, A - 1), (A - 1, A - 1)):
        d = _rays(W, H, A, rot, focal, xs.astype(np.float64), ys.astype(np.float64), dx, dy)
        with np.errstate(divide="ignore", invalid="ignore"):
            t0 = (np.asarray(lo) - o) / d
            t1 = (np.asarray(hi) - o) / d
        tn = np.minimum(t0, t1).max(-1)
        tf = np.maximum(t0, t1).min(-1)
        hit |= (tn <= tf) & (tf >= 0)
    assert hit.any()
    hy, hx = np.nonzero(hit)
    assert hx.min() >= x0 and hx.max() < x1 and hy.min() >= y0 and hy.max() < y1
    # and it is tight: not more than a few pixels around the hit region
    assert x0 >= hx.min() - 4 and x1 <= hx.max() + 5 and y0 >= hy.min() - 4 and y1 <= hy.max() + 5


def test_visible_rectangle_degenerate_cameras_give_the_whole_frame():
    cfg = _cfg(640, 480, 2)
    lo, hi = (-1.0, -1.0, -1.0), (1.0, 1.0, 1.0)
    full = [0, 0, 640, 480]
    assert _visible_rect(cfg, lo, hi, u.rot_matrix(), (0.0, 0.0, 0.0), 600.0) == full           # camera inside the box
    assert _visible_rect(cfg, lo, hi, u.rot_matrix(3.0, 0.0), (0.0, 0.0, -3.2), 600.0) == full   # box behind the camera
    assert _visible_rect(cfg, lo, hi, np.zeros(12, np.float32), (0.0, 0.0, -3.2), 600.0) == full  # singular rotation
    assert _visible_rect(cfg, lo, hi, u.rot_matrix(), (0.0, 0.0, -3.2), 0.0) == full             # no focal length
    assert _visible_rect(cfg, (-3e38,) * 3, (3e38,) * 3, u.rot_matrix(), (0.0, 0.0, -3.2), 600.0) == full  # unbounded scene
    # the default 16:9 framing: the box front face fills 1000/1024 of the height and leaves the side bands out
    x0, y0, x1, y1 = _visible_rect(_cfg(1920, 1080, 2), lo, hi, u.rot_matrix(), (0.0, 0.0, -3.2), 2320.3125)
    assert 5 <= y0 <= 14 and 1066 <= y1 <= 1075 and 420 <= x0 <= 440 and 1480 <= x1 <= 1500


@pytest.mark.parametrize("W,H,A,yaw,pitch,cam", CAMERAS)
@pytest.mark.parametrize("part", [None, (1, 3), (2, 3)])
def test_mixed_launch_lists_cover_every_tile_once_and_split_what_sees_a_sphere(W, H, A, yaw, pitch, cam, part):
    rot = u.rot_matrix(yaw, pitch)
    focal = 1100.0 * A * H / 1024
    row0, rows = 0, H
    if part:  # a row tile of a multi-GPU partition, not aligned to the tile height
        k, n = part
        row0, rows = (H * k) // n, (H * (k + 1)) // n - (H * k) // n
    cfg = _cfg(W, H, A, row0, rows)
    light, split = _tile_lists(cfg, rot, cam, focal)
    gx, gy, sgx = (W + 15) // 16, (rows + 15) // 16, (W + 7) // 8
    assert len(set(light.tolist())) == len(light) and len(set(split.tolist())) == len(split)
    cover = np.zeros((rows, W), np.int32)  # how often each pixel of the row range is rendered
    for t in light:
        by, bx = divmod(int(t), gx)
        assert 0 <= by < gy
        cover[by * 16:(by + 1) * 16, bx * 16:(bx + 1) * 16] += 1
    for t in split:
        by, bx = divmod(int(t), sgx)
        assert by * 8 < rows and bx * 8 < W
        cover[by * 8:(by + 1) * 8, bx * 8:(bx + 1) * 8] += 1
    assert (cover == 1).all()
    # every pixel with a sub-sample ray that hits a sphere lies in a split tile
    ys, xs = np.mgrid[row0:row0 + rows, 0:W]
    o = np.asarray(cam, np.float64)
    sees = np.zeros((rows, W), bool)
    for dx, dy in ((0, 0), (A - 1, A - 1)):
        d = _rays(W, H, A, rot, focal, xs.astype(np.float64), ys.astype(np.float64), dx, dy)
        for c, r2 in SPHERES:
            L = o - np.asarray(c)
            a, b, cc = (d * d).sum(-1), 2.0 * (d @ L), L @ L - r2
            disc = b * b - 4 * a * cc
            sees |= (disc >= 0) & (-b + np.sqrt(np.maximum(disc, 0)) >= 0)
    in_split = np.zeros((rows, W), bool)
    for t in split:
        by, bx = divmod(int(t), sgx)
        in_split[by * 8:(by + 1) * 8, bx * 8:(bx + 1) * 8] = True
    assert not (sees & ~in_split).any()
    if sees.any():  # and the classification is not trivially "split everything"
        assert in_split.mean() < min(1.0, 6.0 * sees.mean() + 0.2)


def test_tile_lists_are_deterministic_and_deal_evenly():
    """Every rank plans the launch on its own host: the result must not depend on anything but the inputs.  The ranks of
    an N-way interleave deal the split sub-tiles (entries phase, phase + N, ...: +- 1 each) and ALL tiles of the static
    launch-order table, skipping those inside a sphere rectangle — so the ordinary tiles a rank really renders may differ
    by a few between ranks, not more."""
    cfg = _cfg(1920, 1080, 2)
    a = _tile_lists(cfg, u.rot_matrix(), (0.0, 0.0, -3.2), 2320.3125)
    b = _tile_lists(cfg, u.rot_matrix(), (0.0, 0.0, -3.2), 2320.3125)
    assert (a[0] == b[0]).all() and (a[1] == b[1]).all()
    assert 500 < len(a[1]) < 2500  # two spheres of ~110 px radius: some hundred tiles, not the frame
    for n in (2, 4, 8):
        sizes = [len(a[1][p::n]) for p in range(n)]
        assert max(sizes) - min(sizes) <= 1
    # the ordinary tiles: reconstruct the full launch-order table (ordinary tiles keep their relative order in it)
    gx, gy = 120, 68
    in_rect = np.ones(gx * gy, bool)
    in_rect[a[0]] = False
    # the table itself is not exported; any order that interleaves the rectangle tiles evenly deals evenly — check the
    # worst case instead: rectangle tiles are < 6 % of the table, so no rank can lose more than that share
    assert in_rect.mean() < 0.06
    # the side of the frame a rank gets is not tied to its phase: both halves of the frame in both phases of a 2-way deal
    gx = 120
    for p in range(2):
        bx = a[0][p::2] % gx
        left = (bx < gx // 2).mean()
        assert 0.4 < left < 0.6
