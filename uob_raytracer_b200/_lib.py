"""ctypes loaders for the two in-tree native libraries.

libuob_rt.so   — the CUDA render path behind include/uob_rt.h
libuob_host.so — host-side scene sources / camera state / framebuffer dump (include/uob_host.h)

There is no Python or CPU fallback for the render path: if libuob_rt.so is
missing, loading raises with the build command to run.
"""
from __future__ import annotations

import ctypes
import os

PKG = os.path.dirname(os.path.abspath(__file__))
RT_LIB_PATH = os.environ.get("UOB_RT_LIB") or os.path.join(PKG, "libuob_rt.so")  # override: kernel-variant experiments only
HOST_LIB_PATH = os.path.join(PKG, "libuob_host.so")

c_float_p = ctypes.POINTER(ctypes.c_float)
c_u32_p = ctypes.POINTER(ctypes.c_uint32)
c_int_p = ctypes.POINTER(ctypes.c_int)

RT_OK, RT_ERR_INVALID, RT_ERR_CUDA, RT_ERR_NO_SCENE, RT_ERR_NO_DEVICE = range(5)
RT_FLAG_STRICT_IEEE = 1 << 0
RT_FLAG_FORCE_BRUTE = 1 << 1
RT_FLAG_FORCE_BVH = 1 << 2
RT_FLAG_COUNT_RAYS = 1 << 3
RT_FLAG_REFERENCE_LOOPS = 1 << 4
RT_FLAG_SPLIT_PIXELS = 1 << 5
RT_FLAG_NO_SPLIT = 1 << 6
RT_FLAG_SPLIT_HEAVY = 1 << 7
RT_FLAG_NO_STREAM_MEMOPS = 1 << 8


class RtConfig(ctypes.Structure):
    """struct rt_config of include/uob_rt.h."""
    _fields_ = [("width", ctypes.c_int), ("height", ctypes.c_int), ("aa", ctypes.c_int),
                ("shadow_samples", ctypes.c_int), ("max_bounces", ctypes.c_int), ("device", ctypes.c_int),
                ("row0", ctypes.c_int), ("rows", ctypes.c_int), ("flags", ctypes.c_uint32),
                ("block_stride", ctypes.c_int), ("block_phase", ctypes.c_int)]


# every symbol include/uob_rt.h declares: name -> (restype, argtypes)
RT_SYMBOLS = {
    "rt_default_config": (None, [ctypes.POINTER(RtConfig)]),
    "rt_create": (ctypes.c_void_p, [ctypes.POINTER(RtConfig)]),
    "rt_upload_scene": (ctypes.c_int, [ctypes.c_void_p, c_float_p, c_float_p, c_float_p, ctypes.c_int]),
    "rt_render": (ctypes.c_int, [ctypes.c_void_p, c_float_p, c_float_p, c_float_p, ctypes.c_float, ctypes.c_void_p]),
    "rt_host_alloc": (ctypes.c_void_p, [ctypes.c_size_t]),
    "rt_host_free": (None, [ctypes.c_void_p]),
    "rt_render_begin": (ctypes.c_int, [ctypes.c_void_p, c_float_p, c_float_p, c_float_p, ctypes.c_float, ctypes.c_void_p]),
    "rt_render_end": (ctypes.c_int, [ctypes.c_void_p]),
    "rt_render_device": (ctypes.c_int, [ctypes.c_void_p, c_float_p, c_float_p, c_float_p, ctypes.c_float,
                                        ctypes.c_void_p, ctypes.c_void_p]),
    "rt_set_stream": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "rt_synchronize": (ctypes.c_int, [ctypes.c_void_p]),
    "rt_device_frame": (ctypes.c_void_p, [ctypes.c_void_p]),
    "rt_last_kernel_ms": (ctypes.c_float, [ctypes.c_void_p]),
    "rt_kernel_launches": (ctypes.c_uint64, [ctypes.c_void_p]),
    "rt_scene_mode": (ctypes.c_char_p, [ctypes.c_void_p]),
    "rt_last_kernel_name": (ctypes.c_char_p, [ctypes.c_void_p]),
    "rt_destroy": (None, [ctypes.c_void_p]),
    "rt_last_error": (ctypes.c_char_p, [ctypes.c_void_p]),
    "rt_get_ray_counts": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint64)]),
    "rt_measure_fp32_peak": (ctypes.c_int, [ctypes.c_void_p, c_float_p]),
    "rt_enable_peer": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "rt_ipc_export_frame": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "rt_ipc_open_frame": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p)]),
    "rt_ipc_close_frame": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "rt_peer_flags": (ctypes.c_void_p, [ctypes.c_void_p]),
    "rt_peer_signal": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p]),
    "rt_peer_wait": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_uint32, ctypes.c_void_p]),
    "rt_gate_next_frame": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32]),
    "rt_frame_slot_words": (ctypes.c_size_t, [ctypes.c_void_p]),
    "rt_signal_after_frame": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "rt_stream_wait_geq": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p]),
    "rt_stream_write": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p]),
    "rt_read_frame_slot": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]),
    "rt_set_strip_targets": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.c_int]),
    "rt_read_strips": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]),
    "rt_peer_add": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.c_void_p]),
    "rt_render_strips": (ctypes.c_int, [ctypes.c_void_p, c_float_p, c_float_p, c_float_p, ctypes.c_float, ctypes.POINTER(ctypes.c_void_p), ctypes.c_int,
                                        ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_uint32, ctypes.c_void_p]),
    "rt_host_register": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_size_t]),
    "rt_host_unregister": (ctypes.c_int, [ctypes.c_void_p]),
    "rt_debug_visible_rect": (ctypes.c_int, [ctypes.c_void_p, c_float_p, c_float_p, c_float_p, c_float_p, ctypes.c_float,
                                             ctypes.POINTER(ctypes.c_int)]),
    "rt_debug_tile_lists": (ctypes.c_int, [ctypes.c_void_p, c_float_p, c_float_p, ctypes.c_float, ctypes.POINTER(ctypes.c_int), ctypes.c_int,
                                           ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]),
    "rt_read_frame": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "rt_version": (ctypes.c_char_p, []),
}

# every symbol include/uob_host.h declares
HOST_SYMBOLS = {
    "uob_test_model_count": (ctypes.c_int, []),
    "uob_load_test_model": (ctypes.c_int, [c_float_p, c_float_p, c_float_p, ctypes.c_int]),
    "uob_load_obj": (ctypes.c_int, [ctypes.c_char_p, c_float_p, c_float_p, c_float_p, ctypes.c_int]),
    "uob_rot_matrix": (None, [ctypes.c_float, ctypes.c_float, c_float_p]),
    "uob_light_step": (None, [c_float_p, c_int_p]),
    "uob_default_camera": (None, [c_float_p, c_float_p, c_float_p]),
    "uob_fitted_focal": (ctypes.c_float, [ctypes.c_int, ctypes.c_int]),
    "uob_save_bmp": (ctypes.c_int, [ctypes.c_char_p, c_u32_p, ctypes.c_int, ctypes.c_int]),
    "uob_save_ppm": (ctypes.c_int, [ctypes.c_char_p, c_u32_p, ctypes.c_int, ctypes.c_int]),
    "uob_write_icosphere_obj": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_int, ctypes.c_float, ctypes.c_float]),
}


def _load(path: str, symbols: dict) -> ctypes.CDLL:
    if not os.path.exists(path):
        raise RuntimeError(
            f"{os.path.basename(path)} is not built — run `python -m uob_raytracer_b200.build` "
            "(or __graft_entry__.build()). The render path has no CPU fallback.")
    lib = ctypes.CDLL(path)
    for name, (restype, argtypes) in symbols.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.restype = restype
        fn.argtypes = argtypes
    return lib


_rt = None
_host = None


def rt_lib() -> ctypes.CDLL:
    global _rt
    if _rt is None:
        _rt = _load(RT_LIB_PATH, RT_SYMBOLS)
    return _rt


def host_lib() -> ctypes.CDLL:
    global _host
    if _host is None:
        _host = _load(HOST_LIB_PATH, HOST_SYMBOLS)
    return _host
