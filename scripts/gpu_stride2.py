import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import uob_raytracer_b200 as u
cfg = u.CONFIGS["cfg2"]
scene = u.load_test_model(); cam = u.Camera()
for B in (10, 1, 0):
    for n in (1, 8, 32):
        with u.Renderer(cfg.width, cfg.height, cfg.aa, cfg.shadow_samples, B, block_stride=n, block_phase=0) as r:
            r.upload_scene(scene)
            ms = []
            for i in range(10):
                r.render_device(cam.rot(), cam.position, cam.light, cfg.focal); ms.append(r.last_kernel_ms)
            print(f"bounces {B}: 1/{n} of the frame: kernel {min(ms)*1e3:.1f} us", flush=True)
