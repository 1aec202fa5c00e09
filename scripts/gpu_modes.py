"""Whole-frame kernel time per lane mapping (plain / mixed / all split), one GPU."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import uob_raytracer_b200 as u
scene = u.load_test_model(); cam = u.Camera()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
for name in sys.argv[1:] or ["head", "cfg2", "cfg3"]:
    cfg = u.CONFIGS[name]
    for mode in (False, "heavy", True):
        with u.Renderer(cfg.width, cfg.height, cfg.aa, cfg.shadow_samples, cfg.max_bounces, split_pixels=mode) as r:
            r.upload_scene(scene)
            ms = []
            for i in range(12):
                flush.zero_(); torch.cuda.synchronize()
                r.render_device(cam.rot(), cam.position, cam.light, cfg.focal); ms.append(r.last_kernel_ms)
            print(f"{name} {r.last_kernel_name}: {np.median(ms[2:])*1e3:.1f} us (min {min(ms)*1e3:.1f})", flush=True)
