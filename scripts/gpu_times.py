import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import uob_raytracer_b200 as u
cfg = u.CONFIGS["cfg2"]
scene = u.load_test_model(); cam = u.Camera()
for B in (10, 0):
    with u.Renderer(cfg.width, cfg.height, cfg.aa, cfg.shadow_samples, B) as r:
        r.upload_scene(scene)
        for _ in range(3): f = r.render(cam.rot(), cam.position, cam.light, cfg.focal)
        d = (f & 0xffffff).astype(np.float64) * 0.064  # us per thread
        print("B", B, "kernel ms", r.last_kernel_ms, "thread time us: mean %.1f  p50 %.1f p99 %.1f max %.1f" % (d.mean(), np.median(d), np.percentile(d, 99), d.max()))
        ys, xs = np.unravel_index(np.argsort(d.ravel())[-5:], d.shape)
        print("  slowest pixels (x,y,us):", [(int(x), int(y), round(float(d[y, x]), 1)) for x, y in zip(xs, ys)])
        # per 16x16 tile max
        H, W = d.shape
        t = d[:H // 16 * 16, :W // 16 * 16].reshape(H // 16, 16, W // 16, 16).max(axis=(1, 3))
        print("  tile max us: mean %.1f p99 %.1f max %.1f; tiles > 40us: %d" % (t.mean(), np.percentile(t, 99), t.max(), (t > 40).sum()))
