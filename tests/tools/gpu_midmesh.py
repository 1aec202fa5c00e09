import sys, os, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import uob_raytracer_b200 as u
from oracle import bind as ob
p = os.path.join(tempfile.gettempdir(), "ico4.obj"); u.write_icosphere_obj(p, 4, 0.2, 0.05)
scene = u.load_test_model() + u.load_obj(p)
W, H, A, S, B = 96, 96, 2, 10, 4
f = 1100.0 * A * H / 1024
cam, light = [0, 0, -3.2], [0, -0.5, -0.7]
want, _ = ob.oracle_render(W, H, A, S, B, f, scene.verts, scene.normals, scene.colors, ob.oracle_rot_matrix(0, 0), cam, light)
for strict in (True, False):
    with u.Renderer(W, H, A, S, B, strict=strict) as r:
        r.upload_scene(scene)
        got = r.render(u.rot_matrix(), cam, light, f)
    d = np.zeros(got.shape, np.int32)
    for sh in (16, 8, 0):
        d = np.maximum(d, np.abs(((got >> sh) & 255).astype(np.int32) - ((want >> sh) & 255).astype(np.int32)))
    print("strict" if strict else "fast", "neq", int((got != want).sum()), "gt1", int((d > 1).sum()), "max", int(d.max()))
