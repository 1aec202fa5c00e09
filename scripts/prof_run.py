"""Minimal driver for ncu: render one workload a few times through the C ABI."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import uob_raytracer_b200 as u
name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4
strict = len(sys.argv) > 3 and sys.argv[3] == "strict"
cfg = u.CONFIGS[name]
scene = u.load_test_model()
if name == "cfg4":  # the Loader.cpp mesh case: 1.31 M-triangle icosphere in the box (BVH path)
    import tempfile
    path = os.path.join(tempfile.gettempdir(), "ico8.obj")
    u.write_icosphere_obj(path, 8, 0.2, 0.05)
    scene = scene + u.load_obj(path)
cam = u.Camera()
with u.Renderer(cfg.width, cfg.height, cfg.aa, cfg.shadow_samples, cfg.max_bounces, strict=strict) as r:
    r.upload_scene(scene)
    for _ in range(n):
        r.render_device(cam.rot(), cam.position, cam.light, cfg.focal)
    r.synchronize()
    print(name, "kernel ms", r.last_kernel_ms)
