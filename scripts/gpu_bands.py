import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import uob_raytracer_b200 as u
cfg = u.CONFIGS["cfg2"]
scene = u.load_test_model(); cam = u.Camera()
rot, c4, l4 = cam.rot(), cam.position.copy(), cam.light.copy()
host = torch.zeros(cfg.width * cfg.height, dtype=torch.int32).pin_memory()
with u.Renderer(cfg.width, cfg.height, cfg.aa, cfg.shadow_samples, cfg.max_bounces) as r:
    r.upload_scene(scene)
    ref = r.render(rot, c4, l4, cfg.focal)
    for _ in range(10): r.render_host_ptr(rot, c4, l4, cfg.focal, host.data_ptr())
    best = 1e9
    for rep in range(3):
        t0 = time.perf_counter()
        for _ in range(200): r.render_host_ptr(rot, c4, l4, cfg.focal, host.data_ptr())
        best = min(best, (time.perf_counter() - t0) / 200)
    ok = bool((host.numpy().view(np.uint32).reshape(cfg.height, cfg.width) == ref).all())
    print(os.environ.get("UOB_RT_LIB", "default").split("/")[-1], "e2e ms", round(best * 1e3, 4), "frame ok", ok, flush=True)
