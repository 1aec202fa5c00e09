"""Host side of the render path: scene sources, camera / light state, headless dump.

Thin numpy wrappers over libuob_host.so (include/uob_host.h), which restates —
GLM-free, in C++ like the reference's host — LoadTestModel (TestModelH.h:44-219),
load_obj (Loader.cpp:11-59), the scene flatten (skeleton.cpp:474-484), the
rotation matrix (skeleton.cpp:149-151) and update()'s light animation
(skeleton.cpp:290-298).
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass, field

import numpy as np

from ._lib import c_float_p, c_int_p, c_u32_p, host_lib

INT_MIN = -(2 ** 31)


@dataclass
class Scene:
    """The three float4 arrays the reference uploads (skeleton.cpp:474-496)."""
    verts: np.ndarray    # (3n, 4) float32, w = 0
    normals: np.ndarray  # (n, 4)  float32, w = 0
    colors: np.ndarray   # (n, 4)  float32, w = material (>0 diffuse, 0 mirror, <0 glass)

    @property
    def n(self) -> int:
        return int(self.colors.shape[0])

    def __add__(self, other: "Scene") -> "Scene":
        # triangles.insert(triangles.end(), ...) of skeleton.cpp:103
        return Scene(np.concatenate([self.verts, other.verts]), np.concatenate([self.normals, other.normals]),
                     np.concatenate([self.colors, other.colors]))


def _scene_call(fn, *pre) -> Scene:
    n = fn(*pre, None, None, None, 0)
    if n == INT_MIN:
        raise IOError("scene source could not be read (missing file, or a face index out of range)")
    n = -n if n < 0 else n
    v = np.zeros((3 * n, 4), np.float32)
    nr = np.zeros((n, 4), np.float32)
    c = np.zeros((n, 4), np.float32)
    if n:
        got = fn(*pre, v.ctypes.data_as(c_float_p), nr.ctypes.data_as(c_float_p), c.ctypes.data_as(c_float_p), n)
        if got != n:
            raise RuntimeError(f"scene source returned {got}, expected {n}")
    return Scene(v, nr, c)


def load_test_model() -> Scene:
    """The Cornell Box of TestModelH.h: 26 triangles, order significant."""
    return _scene_call(host_lib().uob_load_test_model)


def load_obj(path: str) -> Scene:
    """Loader.cpp's load_obj with all its quirks (x1.5, point-reflect + translate, stale normals)."""
    return _scene_call(host_lib().uob_load_obj, str(path).encode())


def rot_matrix(yaw: float = 0.0, pitch: float = 0.0) -> np.ndarray:
    out = np.zeros(12, np.float32)
    host_lib().uob_rot_matrix(yaw, pitch, out.ctypes.data_as(c_float_p))
    return out


def fitted_focal(aa: int, height: int) -> float:
    return float(host_lib().uob_fitted_focal(aa, height))


@dataclass
class Camera:
    """Process globals of skeleton.cpp:61-67 plus the animation flag `lor` (:74)."""
    focal: float = 2200.0
    position: np.ndarray = field(default_factory=lambda: np.array([0.0, 0.0, -3.2, 1.0], np.float32))
    light: np.ndarray = field(default_factory=lambda: np.array([0.0, -0.5, -0.7, 1.0], np.float32))
    yaw: float = 0.0
    pitch: float = 0.0
    lor: bool = True

    def rot(self) -> np.ndarray:
        return rot_matrix(self.yaw, self.pitch)

    def update(self) -> None:
        """The deterministic part of update(): the light ping-pongs in x (skeleton.cpp:290-298)."""
        x = ctypes.c_float(float(self.light[0]))
        lor = ctypes.c_int(1 if self.lor else 0)
        host_lib().uob_light_step(ctypes.byref(x), ctypes.byref(lor))
        self.light[0] = x.value
        self.lor = bool(lor.value)


def save_bmp(path: str, frame: np.ndarray) -> None:
    frame = np.ascontiguousarray(frame, np.uint32)
    rc = host_lib().uob_save_bmp(str(path).encode(), frame.ctypes.data_as(c_u32_p), frame.shape[1], frame.shape[0])
    if rc:
        raise IOError(f"uob_save_bmp failed: {rc}")


def save_ppm(path: str, frame: np.ndarray) -> None:
    frame = np.ascontiguousarray(frame, np.uint32)
    rc = host_lib().uob_save_ppm(str(path).encode(), frame.ctypes.data_as(c_u32_p), frame.shape[1], frame.shape[0])
    if rc:
        raise IOError(f"uob_save_ppm failed: {rc}")


def write_icosphere_obj(path: str, subdiv: int, radius: float = 0.2, noise: float = 0.05) -> int:
    n = host_lib().uob_write_icosphere_obj(str(path).encode(), subdiv, radius, noise)
    if n < 0:
        raise IOError(f"uob_write_icosphere_obj failed: {n}")
    return n
