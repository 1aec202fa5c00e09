"""TEST INFRASTRUCTURE — NOT PRODUCT CODE.

ctypes bindings for the CPU oracle (oracle/liboracle.so, the C restatement) and
for the reference's own sources compiled into oracle/_ref/ (build_ref.py).
Imported only by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / `--impl reference` legs.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

_fp = ctypes.POINTER(ctypes.c_float)
_u32p = ctypes.POINTER(ctypes.c_uint32)
_u64p = ctypes.POINTER(ctypes.c_uint64)
_ip = ctypes.POINTER(ctypes.c_int)

COUNTER_NAMES = ("primary_rays", "shadow_rays", "bounce_rays", "closest_tri_tests", "shadow_stage1_tests",
                 "shadow_stage2_tests", "sphere_tests", "pixels")


class OracleParams(ctypes.Structure):
    _fields_ = [("W", ctypes.c_int), ("H", ctypes.c_int), ("aa", ctypes.c_int), ("shadow_samples", ctypes.c_int),
                ("max_bounces", ctypes.c_int), ("y0", ctypes.c_int), ("y1", ctypes.c_int), ("row_step", ctypes.c_int),
                ("threads", ctypes.c_int), ("focal", ctypes.c_float)]


def _f32(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if shape is not None:
        a = a.reshape(shape)
    return a


def build_oracle(fast: bool = False) -> str:
    """Compile oracle/cornell_oracle.c (strict IEEE build unless fast)."""
    name = "liboracle_fast.so" if fast else "liboracle.so"
    subprocess.check_call(["make", "-s", "-C", HERE, name])
    return os.path.join(HERE, name)


_oracle_cache: dict[bool, ctypes.CDLL] = {}


def oracle_lib(fast: bool = False) -> ctypes.CDLL:
    if fast not in _oracle_cache:
        path = os.path.join(HERE, "liboracle_fast.so" if fast else "liboracle.so")
        src = os.path.join(HERE, "cornell_oracle.c")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
            build_oracle(fast)
        lib = ctypes.CDLL(path)
        lib.oracle_render.argtypes = [ctypes.POINTER(OracleParams), _fp, _fp, _fp, ctypes.c_int, _fp, _fp, _fp, _u32p,
                                      _u64p, _u64p]
        lib.oracle_render.restype = ctypes.c_int
        lib.oracle_global_id.argtypes = [ctypes.c_int] * 3
        lib.oracle_global_id.restype = ctypes.c_int
        _oracle_cache[fast] = lib
    return _oracle_cache[fast]


def oracle_render(W, H, aa, shadow_samples, max_bounces, focal, verts, normals, colors, rot12, cam, light,
                  y0=0, y1=None, row_step=1, threads=0, fast=False, want_row_rays=False):
    """Returns (frame uint32[H,W], counters dict[, row_rays uint64[H]])."""
    lib = oracle_lib(fast)
    verts, normals, colors = _f32(verts, (-1, 4)), _f32(normals, (-1, 4)), _f32(colors, (-1, 4))
    n = colors.shape[0]
    assert verts.shape[0] == 3 * n and normals.shape[0] == n
    rot12, cam, light = _f32(rot12, (12,)), _f32(cam), _f32(light)
    p = OracleParams(W, H, aa, shadow_samples, max_bounces, y0, H if y1 is None else y1, row_step, threads, focal)
    out = np.zeros((H, W), np.uint32)
    ctr = np.zeros(8, np.uint64)
    rows = np.zeros(H, np.uint64)
    rc = lib.oracle_render(ctypes.byref(p), verts.ctypes.data_as(_fp), normals.ctypes.data_as(_fp),
                           colors.ctypes.data_as(_fp), n, rot12.ctypes.data_as(_fp), cam.ctypes.data_as(_fp),
                           light.ctypes.data_as(_fp), out.ctypes.data_as(_u32p), ctr.ctypes.data_as(_u64p),
                           rows.ctypes.data_as(_u64p))
    if rc != 0:
        raise RuntimeError(f"oracle_render failed: {rc}")
    counters = {k: int(v) for k, v in zip(COUNTER_NAMES, ctr)}
    counters["rays"] = counters["primary_rays"] + counters["shadow_rays"] + counters["bounce_rays"]
    return (out, counters, rows) if want_row_rays else (out, counters)


# ---- per-function entry points ------------------------------------------------

def oracle_seed(global_id: int) -> np.ndarray:
    out = np.zeros(3, np.uint32)
    oracle_lib().oracle_seed(ctypes.c_int(global_id), out.ctypes.data_as(_u32p))
    return out


def oracle_xorshift3(v) -> np.ndarray:
    v = np.ascontiguousarray(v, np.uint32)
    out = np.zeros(3, np.uint32)
    oracle_lib().oracle_xorshift3(v.ctypes.data_as(_u32p), out.ctypes.data_as(_u32p))
    return out


def oracle_crush(v, rng: float) -> np.ndarray:
    v = np.ascontiguousarray(v, np.uint32)
    out = np.zeros(3, np.float32)
    oracle_lib().oracle_crush(v.ctypes.data_as(_u32p), ctypes.c_float(rng), out.ctypes.data_as(_fp))
    return out


def oracle_global_id(x: int, y: int, W: int) -> int:
    return int(oracle_lib().oracle_global_id(x, y, W))


def oracle_closest_hits(start, direction, verts, normals, colors):
    start, direction = _f32(start, (-1, 3)), _f32(direction, (-1, 3))
    verts, normals, colors = _f32(verts, (-1, 4)), _f32(normals, (-1, 4)), _f32(colors, (-1, 4))
    m, n = start.shape[0], colors.shape[0]
    ids = np.full(m, -1, np.int32)
    pt, nr, col = np.zeros((m, 3), np.float32), np.zeros((m, 3), np.float32), np.zeros((m, 4), np.float32)
    oracle_lib().oracle_closest_hits(start.ctypes.data_as(_fp), direction.ctypes.data_as(_fp), m,
                                     verts.ctypes.data_as(_fp), normals.ctypes.data_as(_fp), colors.ctypes.data_as(_fp),
                                     n, ids.ctypes.data_as(_ip), pt.ctypes.data_as(_fp), nr.ctypes.data_as(_fp),
                                     col.ctypes.data_as(_fp))
    return ids, pt, nr, col


def oracle_in_shadow(start, direction, radius_sq, verts, colors):
    start, direction = _f32(start, (-1, 3)), _f32(direction, (-1, 3))
    radius_sq = _f32(radius_sq, (-1,))
    verts, colors = _f32(verts, (-1, 4)), _f32(colors, (-1, 4))
    m, n = start.shape[0], colors.shape[0]
    out = np.zeros(m, np.int32)
    oracle_lib().oracle_in_shadow(start.ctypes.data_as(_fp), direction.ctypes.data_as(_fp),
                                  radius_sq.ctypes.data_as(_fp), m, verts.ctypes.data_as(_fp),
                                  colors.ctypes.data_as(_fp), n, out.ctypes.data_as(_ip))
    return out


def oracle_bounce(kind: int, direction, normal, point, medium: float):
    d, nrm, pt = _f32(direction, (3,)), _f32(normal, (3,)), _f32(point, (3,))
    s, o = np.zeros(3, np.float32), np.zeros(3, np.float32)
    med = ctypes.c_float(0)
    oracle_lib().oracle_bounce(kind, d.ctypes.data_as(_fp), nrm.ctypes.data_as(_fp), pt.ctypes.data_as(_fp),
                               ctypes.c_float(medium), s.ctypes.data_as(_fp), o.ctypes.data_as(_fp), ctypes.byref(med))
    return s, o, med.value


def oracle_rot_matrix(yaw: float, pitch: float) -> np.ndarray:
    out = np.zeros(12, np.float32)
    oracle_lib().oracle_rot_matrix(ctypes.c_float(yaw), ctypes.c_float(pitch), out.ctypes.data_as(_fp))
    return out


def oracle_light_sequence(frames: int, x0: float = 0.0) -> np.ndarray:
    """light_position.x after each of `frames` calls of update() (skeleton.cpp:290-298)."""
    lib = oracle_lib()
    x, lor = ctypes.c_float(x0), ctypes.c_int(1)
    out = np.zeros(frames, np.float32)
    for i in range(frames):
        lib.oracle_light_step(ctypes.byref(x), ctypes.byref(lor))
        out[i] = x.value
    return out


# ---- the reference's own sources (oracle/_ref) ---------------------------------

def ref_available(aa=None, shadow_samples=None, max_bounces=None) -> bool:
    if aa is None:
        return os.path.exists(os.path.join(REF_DIR, "libref_scene.so"))
    return os.path.exists(os.path.join(REF_DIR, f"libref_a{aa}_s{shadow_samples}_b{max_bounces}.so"))


_ref_cache: dict[str, ctypes.CDLL] = {}


def _ref(name: str) -> ctypes.CDLL:
    if name not in _ref_cache:
        _ref_cache[name] = ctypes.CDLL(os.path.join(REF_DIR, name))
    return _ref_cache[name]


def ref_speed_available(aa, shadow_samples, max_bounces) -> bool:
    """The -O3 x86-64-v3 build of the reference kernel exists and this CPU can run it (AVX2 + FMA)."""
    if not os.path.exists(os.path.join(REF_DIR, f"libref_a{aa}_s{shadow_samples}_b{max_bounces}_speed.so")):
        return False
    try:
        with open("/proc/cpuinfo") as f:
            flags = next((line for line in f if line.startswith("flags")), "")
    except OSError:
        return False
    return all(f" {w}" in flags for w in ("avx2", "fma", "bmi2"))


def ref_render(W, H, aa, shadow_samples, max_bounces, focal, verts, normals, colors, rot12, cam, light,
               y0=0, y1=None, row_step=1, threads=0, speed=False) -> np.ndarray:
    """The verbatim kernels.cl `draw` on host threads. Returns uint32[H,W].  speed=True: the -O3/AVX2/FMA build
    (timing only, not bit-reproducible)."""
    suffix = "_speed" if speed else ""
    lib = _ref(f"libref_a{aa}_s{shadow_samples}_b{max_bounces}{suffix}.so")
    lib.ref_render.argtypes = [ctypes.c_int] * 5 + [_fp, _fp, _fp, ctypes.c_int, _fp, _fp, _fp, ctypes.c_float, _u32p,
                                                    ctypes.c_int]
    verts, normals, colors = _f32(verts, (-1, 4)), _f32(normals, (-1, 4)), _f32(colors, (-1, 4))
    n = colors.shape[0]
    rot12, cam, light = _f32(rot12, (12,)), _f32(cam), _f32(light)
    cam4, light4 = np.zeros(4, np.float32), np.zeros(4, np.float32)
    cam4[:3], light4[:3] = cam[:3], light[:3]
    out = np.zeros((H, W), np.uint32)
    lib.ref_render(W, H, y0, H if y1 is None else y1, row_step, verts.ctypes.data_as(_fp), normals.ctypes.data_as(_fp),
                   colors.ctypes.data_as(_fp), n, rot12.ctypes.data_as(_fp), cam4.ctypes.data_as(_fp),
                   light4.ctypes.data_as(_fp), ctypes.c_float(focal), out.ctypes.data_as(_u32p), threads)
    return out


def _ref_scene_call(fn, *pre, capacity: int):
    v = np.zeros((3 * capacity, 4), np.float32)
    nr = np.zeros((capacity, 4), np.float32)
    c = np.zeros((capacity, 4), np.float32)
    n = fn(*pre, v.ctypes.data_as(_fp), nr.ctypes.data_as(_fp), c.ctypes.data_as(_fp), capacity)
    if n < 0:
        return _ref_scene_call(fn, *pre, capacity=-n)
    return v[: 3 * n].copy(), nr[:n].copy(), c[:n].copy()


def ref_load_test_model():
    """LoadTestModel (TestModelH.h:44-219) flattened as skeleton.cpp:474-484."""
    lib = _ref("libref_scene.so")
    lib.ref_load_test_model.argtypes = [_fp, _fp, _fp, ctypes.c_int]
    return _ref_scene_call(lib.ref_load_test_model, capacity=64)


def ref_load_obj(path: str):
    """load_obj (Loader.cpp:11-59) flattened as skeleton.cpp:474-484."""
    lib = _ref("libref_scene.so")
    lib.ref_load_obj.argtypes = [ctypes.c_char_p, _fp, _fp, _fp, ctypes.c_int]
    return _ref_scene_call(lib.ref_load_obj, path.encode(), capacity=4096)


def ref_ocl_available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "libref_ocl.so"))


def ref_ocl_source(aa: int = 2, shadow_samples: int = 10, max_bounces: int = 10, W: int = 1024, H: int = 1024) -> str:
    """kernels.cl text for one config.  HEAD values return the embedded text UNMODIFIED; anything else gets the same
    parameter-token substitutions build_ref.py rewrite (6) applies (the reference's only way to change them is a source
    edit).  NVIDIA's OpenCL compiler accepts the rest of the file as is (excess initialisers and the &rays pointer are
    warnings there)."""
    import re
    lib = _ref("libref_ocl.so")
    lib.ref_ocl_source.restype = ctypes.c_char_p
    src = lib.ref_ocl_source().decode()
    if (aa, shadow_samples, max_bounces, W, H) == (2, 10, 10, 1024, 1024):
        return src
    def sub1(pat, rep):
        nonlocal src
        src, k = re.subn(pat, rep, src, flags=re.M)
        if k != 1:
            raise RuntimeError(f"rewrite {pat!r} matched {k} times")
    sub1(r"^#define SCREEN_WIDTH .*$", f"#define SCREEN_WIDTH {W}")
    sub1(r"^#define SCREEN_HEIGHT .*$", f"#define SCREEN_HEIGHT {H}")
    sub1(r"^constant char rays_x = \d+;", f"constant char rays_x = {aa};")
    sub1(r"^constant char rays_y = \d+;", f"constant char rays_y = {aa};")
    sub1(r"^#define aa_rays \d+", f"#define aa_rays {aa * aa}")
    sub1(r"const short light_sources = \d+;", f"const short light_sources = {shadow_samples};")
    sub1(r"const int bounces = \d+;", f"const int bounces = {max_bounces};")
    return src


def ref_ocl_render(W, H, aa, shadow_samples, max_bounces, focal, verts, normals, colors, rot12, cam, light, frames=3):
    """The reference's own OpenCL kernel on whatever OpenCL GPU the box has (oracle/ref_ocl.c).
    Returns (frame uint32[H, W], {"kernel_ms", "total_ms", "device"}); raises RuntimeError with the driver's message."""
    lib = _ref("libref_ocl.so")
    lib.ref_ocl_last_error.restype = ctypes.c_char_p
    verts, normals, colors = _f32(verts), _f32(normals), _f32(colors)
    n = normals.size // 4
    out = np.zeros((H, W), np.uint32)
    k_ms, t_ms = ctypes.c_double(0), ctypes.c_double(0)
    name = ctypes.create_string_buffer(256)
    src = ref_ocl_source(aa, shadow_samples, max_bounces, W, H).encode()
    rot12, cam, light = _f32(rot12, (12,)), _f32(cam), _f32(light)
    cam4, light4 = np.zeros(4, np.float32), np.zeros(4, np.float32)
    cam4[:3], light4[:3] = cam[:3], light[:3]
    lib.ref_ocl_render.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, _fp, _fp, _fp, ctypes.c_int, _fp, _fp, _fp,
                                   ctypes.c_float, ctypes.c_int, _u32p, ctypes.POINTER(ctypes.c_double),
                                   ctypes.POINTER(ctypes.c_double), ctypes.c_char_p]
    rc = lib.ref_ocl_render(src, W, H, verts.ctypes.data_as(_fp), normals.ctypes.data_as(_fp), colors.ctypes.data_as(_fp), n,
                            rot12.ctypes.data_as(_fp), cam4.ctypes.data_as(_fp), light4.ctypes.data_as(_fp),
                            float(focal), int(frames), out.ctypes.data_as(_u32p), ctypes.byref(k_ms), ctypes.byref(t_ms), name)
    if rc != 0:
        raise RuntimeError(f"ref_ocl_render rc={rc}: {lib.ref_ocl_last_error().decode(errors='replace')}")
    return out, {"kernel_ms": k_ms.value, "total_ms": t_ms.value, "device": name.value.decode(errors="replace")}


def frame_hash(frame: np.ndarray) -> str:
    """sha256 over the little-endian ARGB bytes of a frame (first 16 hex digits)."""
    import hashlib
    return hashlib.sha256(np.ascontiguousarray(frame, dtype="<u4").tobytes()).hexdigest()[:16]
