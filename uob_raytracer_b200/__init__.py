"""uob_raytracer_b200 — B200-native render path of harrywaugh/UOB_Raytracer.

The package holds only what the hot path needs: csrc/ (CUDA kernels + the C ABI
of include/uob_rt.h, and the C++ host pieces of include/uob_host.h) and this
thin Python mirror used by tests/ and bench.py.
"""
from .configs import CONFIGS, RenderConfig  # noqa: F401
from .host import (Camera, Scene, fitted_focal, load_obj, load_test_model, rot_matrix, save_bmp, save_ppm,  # noqa: F401
                   write_icosphere_obj)
from .renderer import Renderer, RtError  # noqa: F401
