import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import uob_raytracer_b200 as u
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cfg = u.CONFIGS["cfg2"]
scene = u.load_test_model(); cam = u.Camera()
with u.Renderer(cfg.width, cfg.height, cfg.aa, cfg.shadow_samples, cfg.max_bounces, block_stride=n, block_phase=0) as r:
    r.upload_scene(scene)
    for _ in range(4):
        r.render_device(cam.rot(), cam.position, cam.light, cfg.focal)
    r.synchronize()
    print("stride", n, "ms", r.last_kernel_ms)
