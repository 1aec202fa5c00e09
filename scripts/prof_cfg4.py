"""ncu driver for the BVH kernel: cfg4 scene, a few frames."""
import sys, os, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import uob_raytracer_b200 as u
cfg = u.CONFIGS["cfg4"]
path = os.path.join(tempfile.gettempdir(), "ico8.obj")
if not os.path.exists(path): u.write_icosphere_obj(path, 8, 0.2, 0.05)
scene = u.load_test_model() + u.load_obj(path)
cam = u.Camera()
with u.Renderer(cfg.width, cfg.height, cfg.aa, cfg.shadow_samples, cfg.max_bounces) as r:
    r.upload_scene(scene)
    for _ in range(3):
        r.render_device(cam.rot(), cam.position, cam.light, cfg.focal)
    r.synchronize()
    print("cfg4 kernel ms", r.last_kernel_ms)
