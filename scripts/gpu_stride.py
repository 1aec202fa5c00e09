"""Kernel time of 1/N of the frame on ONE GPU (block-interleaved share), to separate kernel scaling from multi-process effects."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import uob_raytracer_b200 as u
cfg = u.CONFIGS["cfg2"]
scene = u.load_test_model(); cam = u.Camera()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
for n in (1, 2, 4, 8):
    with u.Renderer(cfg.width, cfg.height, cfg.aa, cfg.shadow_samples, cfg.max_bounces, block_stride=n, block_phase=0) as r:
        r.upload_scene(scene)
        ms, msf = [], []
        for i in range(10):
            r.render_device(cam.rot(), cam.position, cam.light, cfg.focal); ms.append(r.last_kernel_ms)
        for i in range(10):
            flush.zero_(); torch.cuda.synchronize()
            r.render_device(cam.rot(), cam.position, cam.light, cfg.focal); msf.append(r.last_kernel_ms)
        print(f"1/{n} of the frame: kernel {min(ms)*1e3:.1f} us warm, {np.median(msf)*1e3:.1f} us after L2 flush", flush=True)
