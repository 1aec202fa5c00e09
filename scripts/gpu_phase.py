"""Kernel time of every phase of an N-way block interleave on one GPU (is the deal balanced?)."""
import sys, os
sys.path.insert(0, os.getcwd())
import uob_raytracer_b200 as u
scene = u.load_test_model(); cam = u.Camera()
cfg = u.CONFIGS["cfg2"]
for n in (2, 4, 8):
    out = []
    for ph in range(n):
        with u.Renderer(cfg.width, cfg.height, cfg.aa, cfg.shadow_samples, cfg.max_bounces, block_stride=n, block_phase=ph) as r:
            r.upload_scene(scene)
            ms = []
            for i in range(8):
                r.render_device(cam.rot(), cam.position, cam.light, cfg.focal); ms.append(r.last_kernel_ms)
            out.append(round(min(ms) * 1e3, 1))
    print(f"stride {n}: per-phase kernel us {out}", flush=True)
