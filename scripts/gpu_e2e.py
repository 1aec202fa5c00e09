import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import uob_raytracer_b200 as u
cfg = u.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
scene = u.load_test_model(); cam = u.Camera()
host = torch.empty(cfg.width * cfg.height, dtype=torch.int32).pin_memory()
rot, c4, l4 = cam.rot(), cam.position.copy(), cam.light.copy()
with u.Renderer(cfg.width, cfg.height, cfg.aa, cfg.shadow_samples, cfg.max_bounces) as r:
    r.upload_scene(scene)
    for _ in range(5): r.render_host_ptr(rot, c4, l4, cfg.focal, host.data_ptr())
    t0 = time.perf_counter(); n = 200
    for _ in range(n): r.render_host_ptr(rot, c4, l4, cfg.focal, host.data_ptr())
    t = (time.perf_counter() - t0) / n
    r.render_device(rot, c4, l4, cfg.focal); k = r.last_kernel_ms
    print(f"{cfg.name}: rt_render e2e {t*1e3:.4f} ms/frame, kernel alone {k:.4f} ms")
