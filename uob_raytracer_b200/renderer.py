"""Python mirror of the reference's device boundary over the C ABI (include/uob_rt.h).

    Renderer(...)            ~ opencl_initialise   (skeleton.cpp:366-497)
    Renderer.upload_scene    ~ the flatten + three buffer writes (:474-496)
    Renderer.render          ~ offload_rendering   (:146-182): blocking, frame valid on return
    Renderer.render_device   ~ kernel only (no read-back), for device-side timing / multi-GPU

Everything executes in libuob_rt.so (CUDA, sm_100a).  No fallback.
"""
from __future__ import annotations

import ctypes

import numpy as np

from ._lib import (RT_FLAG_COUNT_RAYS, RT_FLAG_REFERENCE_LOOPS, RT_FLAG_SPLIT_PIXELS, RT_FLAG_NO_SPLIT, RT_FLAG_SPLIT_HEAVY, RT_FLAG_FORCE_BRUTE, RT_FLAG_FORCE_BVH, RT_FLAG_STRICT_IEEE, RT_OK, RtConfig, c_float_p,
                   rt_lib)
from .host import Camera, Scene


class RtError(RuntimeError):
    pass


def _f32(a, n: int) -> np.ndarray:
    a = np.ascontiguousarray(a, np.float32).reshape(-1)
    if a.size < n:
        raise ValueError(f"expected at least {n} floats, got {a.size}")
    return a


class Renderer:
    def __init__(self, width: int = 1024, height: int = 1024, aa: int = 2, shadow_samples: int = 10,
                 max_bounces: int = 10, device: int = 0, row0: int = 0, rows: int = 0, strict: bool = False,
                 force_bvh: bool = False, force_brute: bool = False, block_stride: int = 0, block_phase: int = 0, count_rays: bool = False,
                 reference_loops: bool = False, split_pixels: bool | str | None = None):
        self._lib = rt_lib()
        flags = (RT_FLAG_STRICT_IEEE if strict else 0) | (RT_FLAG_FORCE_BVH if force_bvh else 0) | \
                (RT_FLAG_FORCE_BRUTE if force_brute else 0) | (RT_FLAG_COUNT_RAYS if count_rays else 0) | \
                (RT_FLAG_REFERENCE_LOOPS if reference_loops else 0) | \
                (0 if split_pixels is None else RT_FLAG_SPLIT_HEAVY if split_pixels == "heavy" else
                 (RT_FLAG_SPLIT_PIXELS if split_pixels else RT_FLAG_NO_SPLIT))
        self.cfg = RtConfig(width, height, aa, shadow_samples, max_bounces, device, row0, rows, flags, block_stride, block_phase)
        self._ctx = self._lib.rt_create(ctypes.byref(self.cfg))
        if not self._ctx:
            raise RtError(self._lib.rt_last_error(None).decode())
        self.width, self.height = width, height
        self.row0 = row0 if rows > 0 else 0
        self.rows = rows if rows > 0 else height

    # -- lifetime ------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_ctx", None):
            self._lib.rt_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc: int) -> None:
        if rc != RT_OK:
            raise RtError(f"[{rc}] " + self._lib.rt_last_error(self._ctx).decode())

    # -- scene ---------------------------------------------------------------
    def upload_scene(self, scene: Scene) -> None:
        n = scene.n
        v, nr, c = _f32(scene.verts, 12 * n), _f32(scene.normals, 4 * n), _f32(scene.colors, 4 * n)
        self._check(self._lib.rt_upload_scene(self._ctx, v.ctypes.data_as(c_float_p), nr.ctypes.data_as(c_float_p),
                                              c.ctypes.data_as(c_float_p), n))

    @property
    def scene_mode(self) -> str:
        return self._lib.rt_scene_mode(self._ctx).decode()

    @property
    def last_kernel_name(self) -> str:
        """The draw kernel instantiation the last render call launched (rt_last_kernel_name)."""
        return self._lib.rt_last_kernel_name(self._ctx).decode()

    # -- rendering -----------------------------------------------------------
    def render(self, rot12, cam, light, focal: float, out: np.ndarray | None = None) -> np.ndarray:
        """Blocking render + read-back of this context's rows. Returns uint32 [rows, width]."""
        if out is None:
            out = np.empty((self.rows, self.width), np.uint32)
        assert out.dtype == np.uint32 and out.flags.c_contiguous and out.size == self.rows * self.width
        r, c, l = _f32(rot12, 12), _f32(cam, 3), _f32(light, 3)
        c4, l4 = np.zeros(4, np.float32), np.zeros(4, np.float32)
        c4[:3], l4[:3] = c[:3], l[:3]
        self._check(self._lib.rt_render(self._ctx, r.ctypes.data_as(c_float_p), c4.ctypes.data_as(c_float_p),
                                        l4.ctypes.data_as(c_float_p), focal, out.ctypes.data))
        return out

    def render_camera(self, camera: Camera, out: np.ndarray | None = None) -> np.ndarray:
        return self.render(camera.rot(), camera.position, camera.light, camera.focal, out)

    def render_host_ptr(self, rot12, cam4, light4, focal: float, host_ptr: int) -> None:
        """Blocking render into a raw host pointer (e.g. pinned memory owned by the caller)."""
        self._check(self._lib.rt_render(self._ctx, rot12.ctypes.data_as(c_float_p), cam4.ctypes.data_as(c_float_p),
                                        light4.ctypes.data_as(c_float_p), focal, host_ptr))

    def render_begin(self, rot12, cam4, light4, focal: float, host_ptr: int) -> None:
        """Pipelined rt_render: enqueue the frame and its read-back into host_ptr, return at once (<= 2 in flight)."""
        r, c, l = _f32(rot12, 12), _f32(cam4, 3), _f32(light4, 3)
        c4, l4 = np.zeros(4, np.float32), np.zeros(4, np.float32)
        c4[:3], l4[:3] = c[:3], l[:3]
        self._check(self._lib.rt_render_begin(self._ctx, r.ctypes.data_as(c_float_p), c4.ctypes.data_as(c_float_p),
                                              l4.ctypes.data_as(c_float_p), focal, host_ptr))

    def render_end(self) -> None:
        """Block until the oldest frame begun with render_begin is complete in its host buffer."""
        self._check(self._lib.rt_render_end(self._ctx))

    def render_device(self, rot12, cam4, light4, focal: float, dev_ptr: int = 0, stream: int = 0) -> None:
        """Asynchronous kernel launch only; dev_ptr = whole-frame device buffer (0 = the context's own)."""
        self._check(self._lib.rt_render_device(self._ctx, rot12.ctypes.data_as(c_float_p),
                                               cam4.ctypes.data_as(c_float_p), light4.ctypes.data_as(c_float_p), focal,
                                               dev_ptr or None, stream or None))

    # -- peer-written frames (multi-GPU without a collective) -----------------
    def read_frame(self, out: np.ndarray | None = None) -> np.ndarray:
        """Blocking read-back of the context's whole frame buffer."""
        if out is None:
            out = np.empty((self.height, self.width), np.uint32)
        self._check(self._lib.rt_read_frame(self._ctx, out.ctypes.data))
        return out

    def read_frame_host_ptr(self, host_ptr: int) -> None:
        self._check(self._lib.rt_read_frame(self._ctx, host_ptr))

    def enable_peer(self, peer_device: int) -> None:
        self._check(self._lib.rt_enable_peer(self._ctx, peer_device))

    def ipc_export_frame(self) -> bytes:
        buf = ctypes.create_string_buffer(64)
        self._check(self._lib.rt_ipc_export_frame(self._ctx, buf))
        return buf.raw

    def ipc_open_frame(self, handle: bytes) -> int:
        out = ctypes.c_void_p()
        self._check(self._lib.rt_ipc_open_frame(self._ctx, ctypes.c_char_p(handle), ctypes.byref(out)))
        return int(out.value)

    def ipc_close_frame(self, dev_ptr: int) -> None:
        self._check(self._lib.rt_ipc_close_frame(self._ctx, dev_ptr))

    # -- frame hand-over flags between GPUs ------------------------------------
    PEER_FLAGS = 64

    @property
    def peer_flags_ptr(self) -> int:
        """Device address of this context's hand-over flags (behind the pixels of its frame buffer)."""
        return int(self._lib.rt_peer_flags(self._ctx) or 0)

    def peer_signal(self, flag_ptr: int, value: int, stream: int = 0) -> None:
        self._check(self._lib.rt_peer_signal(self._ctx, flag_ptr, value, stream or None))

    def peer_wait(self, flags_ptr: int, n: int, value: int, stream: int = 0) -> None:
        self._check(self._lib.rt_peer_wait(self._ctx, flags_ptr, n, value, stream or None))

    def gate_next_frame(self, flag_ptr: int, value: int) -> None:
        """The next render_device stores no pixel before *flag_ptr >= value (the wait folded into the draw kernel)."""
        self._check(self._lib.rt_gate_next_frame(self._ctx, flag_ptr, value))

    @property
    def frame_slot_bytes(self) -> int:
        """Distance in bytes between the two frame slots of this context (pixels + hand-over flags)."""
        return 4 * int(self._lib.rt_frame_slot_words(self._ctx))

    def signal_after_frame(self, counter_ptr: int) -> None:
        """The next render_device ends with *counter_ptr += 1 once all its pixels are visible system-wide (no extra launch)."""
        self._check(self._lib.rt_signal_after_frame(self._ctx, counter_ptr))

    def stream_wait_geq(self, word_ptr: int, value: int, stream: int = 0) -> None:
        self._check(self._lib.rt_stream_wait_geq(self._ctx, word_ptr, value, stream or None))

    def stream_write(self, word_ptr: int, value: int, stream: int = 0) -> None:
        self._check(self._lib.rt_stream_write(self._ctx, word_ptr, value, stream or None))

    def read_frame_slot(self, slot: int, out: np.ndarray | None = None) -> np.ndarray:
        if out is None:
            out = np.empty((self.height, self.width), np.uint32)
        self._check(self._lib.rt_read_frame_slot(self._ctx, slot, out.ctypes.data))
        return out

    def read_frame_slot_host_ptr(self, slot: int, host_ptr: int) -> None:
        self._check(self._lib.rt_read_frame_slot(self._ctx, slot, host_ptr))

    def set_strip_targets(self, frame_ptrs, strip_rows: int) -> None:
        """Deal the rows out over several whole-frame buffers (parallel egress): row y -> frame_ptrs[(y // strip_rows) % n]."""
        n = len(frame_ptrs)
        arr = (ctypes.c_void_p * max(n, 1))(*[int(p) for p in frame_ptrs])
        self._check(self._lib.rt_set_strip_targets(self._ctx, arr, n, strip_rows))

    def peer_add(self, counter_ptrs, stream: int = 0) -> None:
        """One small kernel: everything queued so far is visible system-wide, then each counter += 1."""
        n = len(counter_ptrs)
        arr = (ctypes.c_void_p * n)(*[int(p) for p in counter_ptrs])
        self._check(self._lib.rt_peer_add(self._ctx, arr, n, stream or None))

    def render_strips(self, rot12, cam4, light4, focal: float, frame_bases, rank: int, strip_rows: int, slot: int, deliveries_expected: int,
                      host_ptr: int) -> None:
        """One parallel-egress frame (rt_render_strips), asynchronous on the context's stream."""
        n = len(frame_bases)
        if getattr(self, "_strip_bases_key", None) != tuple(frame_bases):
            self._strip_bases = (ctypes.c_void_p * n)(*[int(p) for p in frame_bases])
            self._strip_bases_key = tuple(frame_bases)
        self._check(self._lib.rt_render_strips(self._ctx, rot12.ctypes.data_as(c_float_p), cam4.ctypes.data_as(c_float_p),
                                               light4.ctypes.data_as(c_float_p), focal, self._strip_bases, n, rank, strip_rows, slot,
                                               deliveries_expected, host_ptr))

    def host_register(self, ptr: int, nbytes: int) -> None:
        self._check(self._lib.rt_host_register(ptr, nbytes))

    def host_unregister(self, ptr: int) -> None:
        self._lib.rt_host_unregister(ptr)

    def read_strips(self, slot: int, strip_rows: int, n: int, phase: int, host_ptr: int, stream: int = 0) -> None:
        """Asynchronous copy of strips phase, phase + n, ... of frame slot `slot` into the same rows of a host frame."""
        self._check(self._lib.rt_read_strips(self._ctx, slot, strip_rows, n, phase, host_ptr, stream or None))

    def set_stream(self, stream: int) -> None:
        """Use the caller's cudaStream_t (0 = back to the context's own) for everything that follows."""
        self._check(self._lib.rt_set_stream(self._ctx, stream or None))

    def synchronize(self) -> None:
        self._check(self._lib.rt_synchronize(self._ctx))

    @property
    def device_frame_ptr(self) -> int:
        return int(self._lib.rt_device_frame(self._ctx) or 0)

    @property
    def last_kernel_ms(self) -> float:
        return float(self._lib.rt_last_kernel_ms(self._ctx))

    def ray_counts(self) -> dict:
        """Rays of the last frame (count_rays=True contexts): primary / shadow / bounce / total."""
        out = (ctypes.c_uint64 * 3)()
        self._check(self._lib.rt_get_ray_counts(self._ctx, out))
        d = {"primary_rays": int(out[0]), "shadow_rays": int(out[1]), "bounce_rays": int(out[2])}
        d["rays"] = sum(d.values())
        return d

    def measure_fp32_peak(self) -> float:
        """FFMA microbenchmark on this device, TFLOP/s (roofline denominator)."""
        out = ctypes.c_float(0)
        self._check(self._lib.rt_measure_fp32_peak(self._ctx, ctypes.byref(out)))
        return float(out.value)

    @property
    def kernel_launches(self) -> int:
        return int(self._lib.rt_kernel_launches(self._ctx))
