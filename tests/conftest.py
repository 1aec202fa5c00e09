import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    # GPU tests must never silently pass on a box without a GPU: skip them only when
    # not explicitly selected with -m gpu.
    selected = config.getoption("-m") or ""
    if "gpu" in selected and "not gpu" not in selected:
        return
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the native libraries (product + oracle) once per session."""
    from uob_raytracer_b200 import build as b
    b.build()
    from oracle import bind as ob
    ob.oracle_lib()


@pytest.fixture(scope="session")
def ob():
    from oracle import bind
    return bind


@pytest.fixture(scope="session")
def golden_scene():
    z = np.load(os.path.join(GOLDEN, "scene_cornell.npz"))
    return z["verts"], z["normals"], z["colors"]


@pytest.fixture(scope="session")
def golden_ico2():
    z = np.load(os.path.join(GOLDEN, "scene_ico2.npz"))
    return z["verts"], z["normals"], z["colors"]


@pytest.fixture(scope="session")
def golden_frames():
    return dict(np.load(os.path.join(GOLDEN, "frames_small.npz")))


@pytest.fixture(scope="session")
def golden_meta():
    with open(os.path.join(GOLDEN, "frame_hashes.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def ray_counts():
    with open(os.path.join(GOLDEN, "ray_counts.json")) as f:
        return json.load(f)


def channel_diff(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Per-pixel max |difference| over the R, G, B channels of two ARGB frames."""
    d = np.zeros(a.shape, np.int32)
    for sh in (16, 8, 0):
        d = np.maximum(d, np.abs(((a >> sh) & 255).astype(np.int32) - ((b >> sh) & 255).astype(np.int32)))
    return d


# north_star tolerance: within 1/255 per RGB channel on at least 99.9 % of pixels
TOL_LEVELS = 1
TOL_FRACTION = 0.999


def assert_within_tolerance(a: np.ndarray, b: np.ndarray, what: str = ""):
    d = channel_diff(a, b)
    ok = float((d <= TOL_LEVELS).mean())
    assert ok >= TOL_FRACTION, f"{what}: only {ok * 100:.4f}% of pixels within {TOL_LEVELS}/255 (need {TOL_FRACTION * 100}%)"
    assert ((a >> 24) == 255).all(), f"{what}: alpha must be 0xFF"
    return ok
