"""Ad-hoc GPU check: strict/fast parity against the oracle + kernel timings. Run under gpurun."""
import sys, time, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import uob_raytracer_b200 as u
from oracle import bind as ob

scene = u.load_test_model()
cam = u.Camera()
rot = cam.rot()
cam4, light4 = cam.position.copy(), cam.light.copy()

def stats(a, b):
    d = np.zeros(a.shape, np.int32)
    for sh in (16, 8, 0):
        d = np.maximum(d, np.abs(((a >> sh) & 255).astype(np.int32) - ((b >> sh) & 255).astype(np.int32)))
    return dict(neq=int((a != b).sum()), gt1=int((d > 1).sum()), maxdiff=int(d.max()), frac_gt1=float((d > 1).mean()))

cases = [("head", 1024, 1024, 2, 10, 10), ("cfg1", 1024, 1024, 1, 1, 0), ("cfg2", 1920, 1080, 2, 8, 10), ("cfg3q", 960, 540, 4, 10, 4)]
big = {"cfg3": (3840, 2160, 4, 10, 4), "cfg5": (7680, 4320, 2, 10, 10)}
if len(sys.argv) > 1:
    cases = [c for c in cases if c[0] in sys.argv[1:]]
import json as _json
counts = _json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests", "golden", "ray_counts.json")))
for name in [a for a in sys.argv[1:] if a in big]:
    W, H, A, S, B = big[name]
    f = 1100.0 * A * H / 1024
    with u.Renderer(W, H, A, S, B) as r:
        r.upload_scene(scene)
        ms = []
        for _ in range(5):
            r.render_device(rot, cam4, light4, f)
            ms.append(r.last_kernel_ms)
    print(name, "fast kernel ms", round(min(ms), 3), "Mrays/s", round(counts[name]["rays"] / min(ms) / 1e3, 1), flush=True)
for name, W, H, A, S, B in cases:
    f = 1100.0 * A * H / 1024
    t = time.time()
    o, ctr = ob.oracle_render(W, H, A, S, B, f, scene.verts, scene.normals, scene.colors, rot, cam4, light4)
    t_or = time.time() - t
    res = {"case": name, "oracle_s": round(t_or, 3), "rays": ctr["rays"]}
    for strict in (True, False):
        with u.Renderer(W, H, A, S, B, strict=strict) as r:
            r.upload_scene(scene)
            g = r.render(rot, cam4, light4, f)
            ms = []
            for _ in range(5):
                r.render_device(rot, cam4, light4, f)
                ms.append(r.last_kernel_ms)
            key = "strict" if strict else "fast"
            res[key] = stats(g, o)
            res[key]["kernel_ms"] = round(min(ms), 4)
            res[key]["Mrays_s"] = round(ctr["rays"] / min(ms) / 1e3, 1)
    print(json.dumps(res), flush=True)
