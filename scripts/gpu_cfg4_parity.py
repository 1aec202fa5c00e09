import sys, os, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import uob_raytracer_b200 as u
cfg = u.CONFIGS["cfg4"]
path = os.path.join(tempfile.gettempdir(), "ico8.obj")
if not os.path.exists(path): u.write_icosphere_obj(path, 8, 0.2, 0.05)
scene = u.load_test_model() + u.load_obj(path)
cam = u.Camera()
frames = {}
for strict in (True, False):
    with u.Renderer(cfg.width, cfg.height, cfg.aa, cfg.shadow_samples, cfg.max_bounces, strict=strict) as r:
        r.upload_scene(scene)
        frames[strict] = r.render(cam.rot(), cam.position, cam.light, cfg.focal)
        r.render_device(cam.rot(), cam.position, cam.light, cfg.focal); ms = r.last_kernel_ms
    print("strict" if strict else "fast", "kernel ms", round(ms, 3))
a, b = frames[True], frames[False]
d = np.zeros(a.shape, np.int32)
for sh in (16, 8, 0):
    d = np.maximum(d, np.abs(((a >> sh) & 255).astype(np.int32) - ((b >> sh) & 255).astype(np.int32)))
print("cfg4 fast vs strict: neq", int((a != b).sum()), "gt1", int((d > 1).sum()), "frac within 1/255: %.5f%%" % (100 * (d <= 1).mean()), "max", int(d.max()))
