"""GPU-box tool (TEST INFRASTRUCTURE): the reference's own OpenCL `draw` kernel on the B200 through NVIDIA's OpenCL
driver (oracle/ref_ocl.c), next to the oracle (parity, north-star tolerance) and the new CUDA path (timing).
Writes gpurun_out/ocl_ref.json.  Usage: python tests/tools/gpu_ocl_ref.py [head cfg1 cfg2 ...]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np

import uob_raytracer_b200 as u
from oracle import bind as ob

CASES = {"head": (1024, 1024, 2, 10, 10), "cfg1": (1024, 1024, 1, 1, 0), "cfg2": (1920, 1080, 2, 8, 10),
         "cfg3": (3840, 2160, 4, 10, 4)}


def stats(a, b):
    d = np.zeros(a.shape, np.int32)
    for sh in (16, 8, 0):
        d = np.maximum(d, np.abs(((a >> sh) & 255).astype(np.int32) - ((b >> sh) & 255).astype(np.int32)))
    return dict(neq=int((a != b).sum()), gt1=int((d > 1).sum()), maxdiff=int(d.max()), frac_le1=float((d <= 1).mean()))


def main():
    names = [a for a in sys.argv[1:] if a in CASES] or ["head", "cfg1", "cfg2"]
    counts = json.load(open(os.path.join(ROOT, "tests", "golden", "ray_counts.json")))
    scene = u.load_test_model()
    cam = u.Camera()
    rot, cam4, light4 = cam.rot(), cam.position.copy(), cam.light.copy()
    results = []
    for name in names:
        W, H, A, S, B = CASES[name]
        f = 1100.0 * A * H / 1024
        res = {"case": name, "W": W, "H": H, "aa": A, "shadow": S, "bounces": B,
               "source": "unmodified" if name == "head" else "parameter tokens substituted"}
        try:
            g, info = ob.ref_ocl_render(W, H, A, S, B, f, scene.verts, scene.normals, scene.colors, rot, cam4, light4, frames=5)
        except Exception as e:  # noqa: BLE001 - report and go on
            res["error"] = str(e)[:1500]
            results.append(res)
            print(json.dumps(res), flush=True)
            continue
        res["opencl"] = info
        rays = counts.get(name, {}).get("rays")
        if rays:
            res["opencl"]["Mrays_s"] = round(rays / info["kernel_ms"] / 1e3, 1)
        if W * H <= 1920 * 1080:
            o, _ = ob.oracle_render(W, H, A, S, B, f, scene.verts, scene.normals, scene.colors, rot, cam4, light4)
            res["opencl_vs_oracle"] = stats(g, o)
        for strict in (False, True):
            with u.Renderer(W, H, A, S, B, strict=strict) as r:
                r.upload_scene(scene)
                ours = r.render(rot, cam4, light4, f)
                ms = []
                for _ in range(5):
                    r.render_device(rot, cam4, light4, f)
                    ms.append(r.last_kernel_ms)
            key = "strict" if strict else "fast"
            res["cuda_" + key] = {"kernel_ms": round(min(ms), 4), "vs_opencl": stats(ours, g)}
        res["kernel_speedup_fast"] = round(info["kernel_ms"] / res["cuda_fast"]["kernel_ms"], 1)
        results.append(res)
        print(json.dumps(res), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(results, open(os.path.join(ROOT, "gpurun_out", "ocl_ref.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
