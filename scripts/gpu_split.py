"""Four-lanes-per-pixel (SPLIT) launches vs the default mapping: kernel time of a 1/N block-interleaved share of the frame
on one GPU, and frame equality.  Run under gpurun."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import uob_raytracer_b200 as u
scene = u.load_test_model(); cam = u.Camera()
rot, c4, l4 = cam.rot(), cam.position.copy(), cam.light.copy()
for name in (sys.argv[1:] or ["cfg2", "cfg3"]):
    cfg = u.CONFIGS[name]
    frames = {}
    for strict in (False, True):
        for split in (False, True, "heavy"):
            with u.Renderer(cfg.width, cfg.height, cfg.aa, cfg.shadow_samples, cfg.max_bounces, strict=strict, split_pixels=split) as r:
                r.upload_scene(scene)
                frames[(strict, split)] = r.render(rot, c4, l4, cfg.focal)
        a, b, c = frames[(strict, False)], frames[(strict, True)], frames[(strict, "heavy")]
        print(name, "strict" if strict else "fast", "pixels differing from default: split", int((a != b).sum()), " mixed", int((a != c).sum()), flush=True)
    for n in (1, 2, 4, 8):
        row = []
        for strict in (False, True):
            for split in (False, True, "heavy"):
                with u.Renderer(cfg.width, cfg.height, cfg.aa, cfg.shadow_samples, cfg.max_bounces, strict=strict, split_pixels=split,
                                block_stride=n, block_phase=0) as r:
                    r.upload_scene(scene)
                    ms = []
                    for i in range(8):
                        r.render_device(rot, c4, l4, cfg.focal); ms.append(r.last_kernel_ms)
                    row.append(f"{'strict' if strict else 'fast'}{'/mixed' if split == 'heavy' else '/split' if split else ''} {min(ms) * 1e3:.1f}")
        print(f"{name} 1/{n}: " + "  ".join(row) + "  (us)", flush=True)
