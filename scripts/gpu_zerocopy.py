"""Experiment: the draw kernel storing its pixels straight into pinned host memory (zero-copy over PCIe)
versus kernel + banded read-back (rt_render).  Run under gpurun."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import uob_raytracer_b200 as u

cfg = u.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
scene = u.load_test_model()
cam = u.Camera()
rot, cam4, light4 = cam.rot(), cam.position.copy(), cam.light.copy()
W, H = cfg.width, cfg.height
host = torch.zeros(H * W, dtype=torch.int32).pin_memory()
host2 = torch.zeros(H * W, dtype=torch.int32).pin_memory()
with u.Renderer(W, H, cfg.aa, cfg.shadow_samples, cfg.max_bounces) as r:
    r.upload_scene(scene)
    ref = r.render(rot, cam4, light4, cfg.focal)
    for name, fn in (("zero-copy kernel", lambda: (r.render_device(rot, cam4, light4, cfg.focal, dev_ptr=host.data_ptr()), r.synchronize())),
                     ("rt_render pinned", lambda: r.render_host_ptr(rot, cam4, light4, cfg.focal, host2.data_ptr()))):
        for _ in range(5):
            fn()
        ts = []
        for _ in range(50):
            t0 = time.perf_counter()
            fn()
            ts.append(time.perf_counter() - t0)
        ts.sort()
        print(name, "median ms", round(ts[len(ts) // 2] * 1e3, 4), "min", round(ts[0] * 1e3, 4), "kernel ms", round(r.last_kernel_ms, 4), flush=True)
    a = host.numpy().view(np.uint32).reshape(H, W)
    b = host2.numpy().view(np.uint32).reshape(H, W)
    print("zero-copy frame equal:", bool((a == ref).all()), " rt_render frame equal:", bool((b == ref).all()))
