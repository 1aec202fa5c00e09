"""CPU tests of the host side (libuob_host.so) and of the C-ABI surface (no compute calls)."""
import ctypes
import os
import re

import numpy as np
import pytest

import uob_raytracer_b200 as u
from uob_raytracer_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b((?:rt|uob)_[a-z0-9_]+)\s*\(", src)))


def test_rt_library_exports_every_declared_symbol():
    names = _declared("uob_rt.h")
    assert set(names) == set(_lib.RT_SYMBOLS), "binding table and header disagree"
    lib = ctypes.CDLL(_lib.RT_LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"libuob_rt.so does not export {n}"
    assert b"sm_100a" in _lib.rt_lib().rt_version()


def test_host_library_exports_every_declared_symbol():
    names = _declared("uob_host.h")
    assert set(names) == set(_lib.HOST_SYMBOLS)
    lib = ctypes.CDLL(_lib.HOST_LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"libuob_host.so does not export {n}"


def test_default_config_is_reference_head():
    cfg = _lib.RtConfig()
    _lib.rt_lib().rt_default_config(ctypes.byref(cfg))
    assert (cfg.width, cfg.height, cfg.aa, cfg.shadow_samples, cfg.max_bounces) == (1024, 1024, 2, 10, 10)


def test_create_rejects_bad_config_and_reports():
    lib = _lib.rt_lib()
    cfg = _lib.RtConfig(0, 10, 2, 10, 10, 0, 0, 0, 0)
    assert not lib.rt_create(ctypes.byref(cfg))
    assert b"width" in lib.rt_last_error(None)
    assert not lib.rt_create(None)


def test_no_cpu_fallback_without_gpu():
    """On a box without a CUDA device the product refuses to run instead of falling back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(u.RtError, match="no CUDA device|no CPU fallback|CUDA"):
        u.Renderer(64, 64)


def test_product_does_not_import_oracle():
    """The product path must not route through the oracle (or any CPU fallback)."""
    pkg = os.path.join(ROOT, "uob_raytracer_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(root, f)).read()
                assert "oracle.bind" not in text and "from oracle" not in text and "import oracle" not in text, f
                assert "liboracle" not in text and "libref_" not in text, f


def test_cornell_box_matches_reference_golden(golden_scene):
    s = u.load_test_model()
    assert s.n == 26 == _lib.host_lib().uob_test_model_count()
    for got, want in zip((s.verts, s.normals, s.colors), golden_scene):
        assert got.tobytes() == want.tobytes()
    assert (s.colors[:, 3] == 1.0).all()
    assert np.abs(s.verts[:, :3]).max() == 1.0


def test_load_obj_matches_reference_golden(golden_ico2):
    s = u.load_obj(os.path.join(ROOT, "tests", "golden", "ico2.obj"))
    assert s.n == 320
    for got, want in zip((s.verts, s.normals, s.colors), golden_ico2):
        assert got.tobytes() == want.tobytes()
    assert (s.colors == np.array([0.0, 0.2, 0.4, 0.5], np.float32)).all()


def test_load_obj_quirks(tmp_path):
    """x1.5 scale, v <- -v + (-0.4,1.15,-0.7) after the normal was computed, non v/f lines ignored."""
    p = tmp_path / "t.obj"
    p.write_text("# comment\nvn 0 0 1\nv 0 0 0\nv 1 0 0\nv 0 1 0\n\nvt 0 0\nf 1 2 3\n")
    s = u.load_obj(str(p))
    assert s.n == 1
    t = np.array([-0.4, 1.15, -0.7], np.float32)
    assert np.array_equal(s.verts[0, :3], t)
    assert np.array_equal(s.verts[1, :3], np.float32(-1.5) * np.array([1, 0, 0], np.float32) + t)
    assert np.array_equal(s.verts[2, :3], np.float32(-1.5) * np.array([0, 1, 0], np.float32) + t)
    # normal = normalize(cross(e2, e1)) of the scaled, untransformed vertices = (0,0,-1); kept although v flipped
    assert np.array_equal(s.normals[0], np.array([0, 0, -1, 0], np.float32))


def test_load_obj_errors(tmp_path):
    with pytest.raises(IOError):
        u.load_obj(str(tmp_path / "missing.obj"))
    p = tmp_path / "bad.obj"
    p.write_text("v 0 0 0\nf 1 2 3\n")
    with pytest.raises(IOError):
        u.load_obj(str(p))
    e = tmp_path / "empty.obj"
    e.write_text("")
    assert u.load_obj(str(e)).n == 0


def test_scene_append_like_reference_call_site(golden_scene):
    box = u.load_test_model()
    ico = u.load_obj(os.path.join(ROOT, "tests", "golden", "ico2.obj"))
    both = box + ico
    assert both.n == 346 and np.array_equal(both.verts[:78], box.verts) and np.array_equal(both.colors[26:], ico.colors)


def test_rot_matrix_and_light_animation_match_oracle(ob):
    for yaw, pitch in [(0, 0), (0.3, -0.2), (-1.1, 0.7), (3.0, 1.5)]:
        assert u.rot_matrix(yaw, pitch).tobytes() == ob.oracle_rot_matrix(yaw, pitch).tobytes()
    cam = u.Camera()
    assert cam.focal == 2200.0 and list(cam.position[:3]) == [0.0, 0.0, np.float32(-3.2)]
    xs = []
    for _ in range(300):
        cam.update()
        xs.append(cam.light[0])
    assert np.array(xs, np.float32).tobytes() == ob.oracle_light_sequence(300).tobytes()


def test_fitted_focal():
    assert u.fitted_focal(2, 1024) == 2200.0
    assert u.fitted_focal(2, 1080) == 2320.3125
    assert u.CONFIGS["cfg3"].focal == 9281.25 == u.fitted_focal(4, 2160)


def test_framebuffer_dump(tmp_path):
    frame = np.zeros((3, 5), np.uint32)
    frame[:] = 0xFF000000
    frame[0, 0] = 0xFFFF0000  # red, top-left
    frame[2, 4] = 0xFF0000FF  # blue, bottom-right
    u.save_ppm(str(tmp_path / "f.ppm"), frame)
    raw = (tmp_path / "f.ppm").read_bytes()
    assert raw.startswith(b"P6\n5 3\n255\n")
    px = np.frombuffer(raw[len(b"P6\n5 3\n255\n"):], np.uint8).reshape(3, 5, 3)
    assert list(px[0, 0]) == [255, 0, 0] and list(px[2, 4]) == [0, 0, 255]
    u.save_bmp(str(tmp_path / "f.bmp"), frame)
    from PIL import Image
    im = np.asarray(Image.open(str(tmp_path / "f.bmp")).convert("RGB"))
    assert im.shape == (3, 5, 3) and list(im[0, 0]) == [255, 0, 0] and list(im[2, 4]) == [0, 0, 255]


def test_icosphere_generator(tmp_path):
    p = tmp_path / "i.obj"
    assert u.write_icosphere_obj(str(p), 3, 0.2, 0.05) == 20 * 4 ** 3
    s = u.load_obj(str(p))
    assert s.n == 1280
    # after load_obj's transform the mesh sits inside the box
    assert np.abs(s.verts[:, :3]).max() < 1.5
    a = p.read_bytes()
    u.write_icosphere_obj(str(p), 3, 0.2, 0.05)
    assert a == p.read_bytes(), "generator must be deterministic"


def test_parallel_obj_ingest_matches_reference_loader(ob, tmp_path):
    """A mesh big enough for the multi-threaded parser (81,920 faces, > 1 MB of text) against the reference's
    own load_obj compiled from Loader.cpp: identical bytes."""
    if not ob.ref_available():
        pytest.skip("oracle/_ref not built")
    p = tmp_path / "ico6.obj"
    assert u.write_icosphere_obj(str(p), 6, 0.2, 0.05) == 81920
    assert p.stat().st_size > (1 << 20)
    s = u.load_obj(str(p))
    v, n, c = ob.ref_load_obj(str(p))
    assert s.n == 81920 == c.shape[0]
    assert s.verts.tobytes() == v.tobytes() and s.normals.tobytes() == n.tobytes() and s.colors.tobytes() == c.tobytes()


def test_obj_number_spellings_match_reference_loader(ob, tmp_path):
    """Spellings the fast path (std::from_chars) does not take itself — leading '+', exponent forms, tabs, indented
    lines, CR line ends, extra tokens — still read as the reference's stream extractors read them: identical bytes."""
    if not ob.ref_available():
        pytest.skip("oracle/_ref/libref_scene.so not built")
    txt = ("# odd but valid\nv 1 2 3\nv\t+1.5  -2.5e-1 .5\n  v 1e2 2E+1 3.\nvn 0 0 1\nvt 0 0\n"
           "v 0.1000000015 123456.789e-3 -0\ng grp\nf 1 2 3\r\nf 2 3 4\nf   4 1\t2  \ns off\nf 3 1 4 2\n")
    p = tmp_path / "q.obj"
    p.write_text(txt)
    s = u.load_obj(str(p))
    v, n, c = ob.ref_load_obj(str(p))
    assert s.n == 4
    assert np.array_equal(s.verts.view(np.uint32), v.view(np.uint32))
    assert np.array_equal(s.normals.view(np.uint32), n.view(np.uint32))
    assert np.array_equal(s.colors.view(np.uint32), c.view(np.uint32))
