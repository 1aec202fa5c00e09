"""cfg4 timing: Cornell box + 1.31 M-triangle icosphere through the BVH path (no oracle involved)."""
import sys, os, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import uob_raytracer_b200 as u
cfg = u.CONFIGS["cfg4"]
t = time.time()
path = os.path.join(tempfile.gettempdir(), "ico8.obj")
u.write_icosphere_obj(path, 8, 0.2, 0.05)
t1 = time.time()
mesh = u.load_obj(path)
t2 = time.time()
scene = u.load_test_model() + mesh
cam = u.Camera()
print(f"obj write {t1 - t:.2f}s, load_obj {t2 - t1:.2f}s, triangles {scene.n}", flush=True)
for strict in (False, True):
    with u.Renderer(cfg.width, cfg.height, cfg.aa, cfg.shadow_samples, cfg.max_bounces, strict=strict) as r:
        t3 = time.time(); r.upload_scene(scene); t4 = time.time()
        ms = []
        for _ in range(5):
            r.render_device(cam.rot(), cam.position, cam.light, cfg.focal); ms.append(r.last_kernel_ms)
        print(("strict" if strict else "fast"), "upload+build %.3fs" % (t4 - t3), "kernel ms", round(min(ms), 3), flush=True)
