"""Small workload for compute-sanitizer: every kernel family once at tiny sizes."""
import sys, os, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import uob_raytracer_b200 as u
scene = u.load_test_model(); cam = u.Camera(focal=u.fitted_focal(2, 96))
p = os.path.join(tempfile.gettempdir(), "ico3.obj"); u.write_icosphere_obj(p, 3, 0.2, 0.05)
mesh = scene + u.load_obj(p)
for kw in (dict(), dict(strict=True), dict(shadow_samples=7), dict(aa=3, shadow_samples=4), dict(block_stride=3, block_phase=1), dict(count_rays=True)):
    args = dict(width=101, height=96, aa=2, shadow_samples=8, max_bounces=4); args.update(kw)
    with u.Renderer(**args) as r:
        r.upload_scene(scene)
        f = r.render(cam.rot(), cam.position, cam.light, u.fitted_focal(args["aa"], 96))
        print(kw, hex(int(f.sum()) & 0xffffffff))
for kw in (dict(), dict(strict=True)):
    with u.Renderer(96, 96, 2, 8, 4, **kw) as r:
        r.upload_scene(mesh)
        f = r.render(cam.rot(), cam.position, cam.light, cam.focal)
        print("bvh", r.scene_mode, kw, hex(int(f.sum()) & 0xffffffff))
