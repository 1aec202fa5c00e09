import sys, os
sys.path.insert(0, os.getcwd())
import uob_raytracer_b200 as u
scene = u.load_test_model(); cam = u.Camera()
for name in ("cfg2", "head"):
    cfg = u.CONFIGS[name]
    for split in (False, True):
        with u.Renderer(cfg.width, cfg.height, cfg.aa, cfg.shadow_samples, cfg.max_bounces, split_pixels=split) as r:
            r.upload_scene(scene)
            ms = []
            for i in range(8):
                r.render_device(cam.rot(), cam.position, cam.light, cfg.focal); ms.append(r.last_kernel_ms)
            print(name, "split" if split else "default", "kernel us", round(min(ms) * 1e3, 1), flush=True)
