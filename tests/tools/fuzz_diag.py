"""Diagnose fast-vs-strict disagreement on fuzz cases of tests/test_gpu_parity.py (GPU box).

    python tests/tools/fuzz_diag.py 108 41 ...     # seeds; no arguments = scan all 300 and report the worst 12

Per seed: share of pixels within 1/255, pixels off by > 16 levels, and where they sit: pixels whose colour involves a
bounce (strict frame with B bounces != strict frame with 0 bounces) or direct shading only."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import uob_raytracer_b200 as u  # noqa: E402
import test_gpu_parity as T  # noqa: E402
from conftest import channel_diff  # noqa: E402

z = np.load(os.path.join(ROOT, "tests", "golden", "scene_cornell.npz"))
golden = (z["verts"], z["normals"], z["colors"])


def case(seed):
    """Same draw sequence as test_fuzz_culled_strict_equals_reference_loops."""
    rng = np.random.default_rng(77000 + seed)
    v, nn, c, n_extra, coplanar = T._fuzz_scene(rng, golden)
    W, H = int(rng.integers(33, 130)), int(rng.integers(20, 90))
    A = int(rng.choice([1, 2, 2, 3, 4]))
    S = int(rng.choice([1, 2, 4, 5, 8, 10, 3]))
    B = int(rng.integers(0, 11))
    f = 1100.0 * A * H / 1024 * float(rng.uniform(0.4, 1.8))
    where = int(rng.integers(0, 6))
    if where == 0:
        cam = [float(rng.uniform(-0.6, 0.6)), float(rng.uniform(-0.6, 0.6)), float(rng.uniform(-3.4, -1.2))]
    elif where == 1:
        cam = [float(rng.uniform(-0.9, 0.9)), float(rng.uniform(-0.9, 0.9)), float(rng.uniform(-0.9, 0.9))]
    elif where == 2:
        cam = [0.3 + float(rng.uniform(-0.1, 0.1)), 0.1 + float(rng.uniform(-0.1, 0.1)), -0.5 + float(rng.uniform(-0.1, 0.1))]
    elif where == 3:
        cam = [-0.4 + float(rng.uniform(-0.08, 0.08)), 0.8 + float(rng.uniform(-0.08, 0.08)), -0.5 + float(rng.uniform(-0.08, 0.08))]
    elif where == 4:
        d = rng.normal(size=3)
        d /= np.linalg.norm(d)
        cam = (np.array([0.3, 0.1, -0.5]) + d * (np.sqrt(0.075) + float(rng.uniform(1e-4, 0.05)))).tolist()
    else:
        cam = [float(rng.uniform(-3, 3)), float(rng.uniform(-3, 3)), float(rng.uniform(-6, -1.5))]
    lw = int(rng.integers(0, 5))
    if lw == 0:
        light = [float(rng.uniform(-0.9, 0.9)), float(rng.uniform(-0.9, 0.9)), float(rng.uniform(-0.9, 0.9))]
    elif lw == 1:
        light = [float(rng.uniform(-0.9, 0.9)), float(rng.uniform(-0.9, 0.9)), float(rng.uniform(-0.9, 0.9))]
        ax = int(rng.integers(0, 3))
        light[ax] = float(rng.choice([-1.0, 1.0])) + float(rng.uniform(-0.08, 0.08))
    elif lw == 2:
        light = [float(rng.uniform(-0.7, 0.7)), float(rng.uniform(0.0, 1.0)), float(rng.uniform(-0.7, 0.7))]
    elif lw == 3:
        light = [float(rng.uniform(-2.5, 2.5)), float(rng.uniform(-2.5, 2.5)), float(rng.uniform(-4.0, 2.5))]
    else:
        light = [0.3 + float(rng.uniform(-0.35, 0.35)), 0.1 + float(rng.uniform(-0.35, 0.35)), -0.5 + float(rng.uniform(-0.35, 0.35))]
    span = float(rng.choice([0.3, 0.8, np.pi / 2]))
    m = dict(W=W, H=H, aa=A, shadow_samples=S, max_bounces=B, focal=f, cam=cam, light=light,
             yaw=float(rng.uniform(-span, span)), pitch=float(rng.uniform(-span, span)))
    return m, u.Scene(v, nn, c), n_extra, where, (lw, coplanar)


def report(seed, verbose=True):
    m, scene, n_extra, where, lw = case(seed)
    strict = T._render(m, scene, True, split_pixels=False)
    fast = T._render(m, scene, False, split_pixels=False)
    d = channel_diff(fast, strict)
    m0 = dict(m, max_bounces=0)
    bounce_px = T._render(m0, scene, True, split_pixels=False) != strict
    off1, off16 = d > 1, d > 16
    line = (f"seed {seed:3d} {m['W']:3d}x{m['H']:2d} aa{m['aa']} S{m['shadow_samples']:2d} B{m['max_bounces']:2d} +{n_extra:2d} cam{where} light{lw[0]} coplanar={int(lw[1])}: "
            f"within1 {100 * (1 - off1.mean()):7.3f}%  >1: {int(off1.sum()):4d} (bounce px {int((off1 & bounce_px).sum()):4d}, direct {int((off1 & ~bounce_px).sum()):4d})  "
            f">16: {int(off16.sum()):4d} (bounce {int((off16 & bounce_px).sum()):4d})  bounce px share {100 * bounce_px.mean():5.1f}%")
    if verbose:
        print(line, flush=True)
    if os.environ.get("FUZZ_DUMP"):
        np.savez_compressed(os.path.join(ROOT, "gpurun_out", f"fuzz_{seed}_{os.environ['FUZZ_DUMP']}.npz"), fast=fast, strict=strict,
                            bounce_px=bounce_px, verts=scene.verts, normals=scene.normals, colors=scene.colors,
                            meta=np.array([m["W"], m["H"], m["aa"], m["shadow_samples"], m["max_bounces"], m["focal"], *m["cam"], *m["light"], m["yaw"], m["pitch"]], np.float64))
    return float(off1.mean()), line


if __name__ == "__main__":
    seeds = [int(a) for a in sys.argv[1:]]
    if seeds:
        for s in seeds:
            report(s)
    else:
        res = sorted((report(s, verbose=False) + (s,) for s in range(300)), reverse=True)
        for frac, line, s in res[:15]:
            print(line)
        print("seeds with every pixel within 1/255:", sum(1 for r in res if r[0] == 0.0), "of 300")
