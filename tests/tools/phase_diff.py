"""Which pixels differ between the 1-GPU plain frame and the N interleaved shares (one GPU renders them one after the other)?"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import uob_raytracer_b200 as u
name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 8
cfg = u.CONFIGS[name]
scene = u.load_test_model(); cam = u.Camera()
with u.Renderer(cfg.width, cfg.height, cfg.aa, cfg.shadow_samples, cfg.max_bounces) as r:
    r.upload_scene(scene)
    r.render_device(cam.rot(), cam.position, cam.light, cfg.focal)
    plain = r.read_frame().copy()
    print("plain:", r.last_kernel_name)
for strict in (False,):
    total = 0
    for ph in range(N):
        with u.Renderer(cfg.width, cfg.height, cfg.aa, cfg.shadow_samples, cfg.max_bounces, block_stride=N, block_phase=ph) as r:
            r.upload_scene(scene)
            r.render_device(cam.rot(), cam.position, cam.light, cfg.focal)
            part = r.read_frame().copy()
            k = r.last_kernel_name
        mask = part != 0
        bad = mask & (part != plain)
        total += int(bad.sum())
        for y, x in zip(*np.nonzero(bad)):
            print(f"phase {ph} ({k}): pixel ({x},{y}) tile ({x//16},{y//16}) share {hex(part[y,x])} plain {hex(plain[y,x])}")
    print("pixels differing:", total)
