"""A/B timing of kernel-variant libraries (scripts/build_variants.sh) against the default build, in ONE gpurun call:
   python scripts/gpu_ab.py [reps]      -> for the default lib and every uob_raytracer_b200/variants/var_*.so, a child process
   renders HEAD / cfg2 / cfg3 (fast + strict), a 1/8 interleaved share of cfg2 and — if a mesh is asked for — cfg4, and prints
   the median and minimum kernel time over `reps` frames plus a sha of each frame (variants must agree bit for bit)."""
import glob, hashlib, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(reps: int, with_mesh: bool):
    import numpy as np
    import uob_raytracer_b200 as u
    scene = u.load_test_model()
    cam = u.Camera()
    rot, cam4, light4 = cam.rot(), cam.position.copy(), cam.light.copy()
    out = {}
    cases = [("head", 1024, 1024, 2, 10, 10, {}), ("cfg2", 1920, 1080, 2, 8, 10, {}), ("cfg3", 3840, 2160, 4, 10, 4, {}),
             ("cfg2_1of8", 1920, 1080, 2, 8, 10, dict(block_stride=8, block_phase=3)),
             ("cfg2_s4", 1920, 1080, 2, 4, 10, {})]  # four shadow samples: 51 KB of shared memory per block, four blocks fit an SM
    for name, W, H, A, S, B, kw in cases:
        f = 1100.0 * A * H / 1024
        for strict in (False, True):
            if strict and name == "cfg3":
                continue
            with u.Renderer(W, H, A, S, B, strict=strict, **kw) as r:
                r.upload_scene(scene)
                g = r.render(rot, cam4, light4, f)
                ms = []
                for _ in range(reps if name != "cfg3" else max(reps // 3, 5)):
                    r.render_device(rot, cam4, light4, f)
                    ms.append(r.last_kernel_ms)
                ms.sort()
                out[f"{name}_{'strict' if strict else 'fast'}"] = dict(med=round(ms[len(ms) // 2], 4), min=round(ms[0], 4),
                                                                      sha=hashlib.sha256(np.ascontiguousarray(g).tobytes()).hexdigest()[:12])
    if with_mesh:
        import tempfile
        path = os.path.join(tempfile.gettempdir(), f"ab_ico8_{os.getpid()}.obj")
        u.write_icosphere_obj(path, 8, 0.2, 0.05)
        mesh = u.load_obj(path)
        os.unlink(path)
        big = u.Scene(np.concatenate([scene.verts, mesh.verts]), np.concatenate([scene.normals, mesh.normals]), np.concatenate([scene.colors, mesh.colors]))
        for strict in (False, True):
            with u.Renderer(1920, 1080, 2, 8, 10, strict=strict) as r:
                r.upload_scene(big)
                g = r.render(rot, cam4, light4, 1100.0 * 2 * 1080 / 1024)
                ms = []
                for _ in range(max(reps // 3, 5)):
                    r.render_device(rot, cam4, light4, 1100.0 * 2 * 1080 / 1024)
                    ms.append(r.last_kernel_ms)
                ms.sort()
                out[f"cfg4_{'strict' if strict else 'fast'}"] = dict(med=round(ms[len(ms) // 2], 4), min=round(ms[0], 4),
                                                                    sha=hashlib.sha256(np.ascontiguousarray(g).tobytes()).hexdigest()[:12])
    print("AB " + json.dumps(out), flush=True)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        child(int(sys.argv[2]), sys.argv[3] == "1")
        return
    reps = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 30
    with_mesh = "mesh" in sys.argv
    libs = [("default", None)] + [(os.path.basename(f)[4:-3], f) for f in sorted(glob.glob(os.path.join(ROOT, "uob_raytracer_b200", "variants", "var_*.so")))]
    results = {}
    for rnd in range(2):  # two passes, interleaved: drift of the box shows up as a difference between the passes
        for name, lib in libs:
            env = dict(os.environ)
            if lib:
                env["UOB_RT_LIB"] = lib
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", str(reps), "1" if with_mesh else "0"], env=env,
                               stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
            line = [l for l in p.stdout.splitlines() if l.startswith("AB ")]
            if not line:
                print(f"== {name}: FAILED\n{p.stdout[-2000:]}")
                continue
            results.setdefault(name, []).append(json.loads(line[0][3:]))
    keys = list(next(iter(results.values()))[0].keys())
    print("%-14s" % "case" + "".join("%22s" % n for n in results))
    for k in keys:
        row = "%-14s" % k
        for n, passes in results.items():
            row += "%22s" % "/".join("%.4f" % q[k]["med"] for q in passes)
        print(row)
    base = results["default"][0]
    for n, passes in results.items():
        bad = [k for k in keys for q in passes if q[k]["sha"] != base[k]["sha"]]
        print(f"frames of {n}: {'identical to default' if not bad else 'DIFFER in ' + ','.join(sorted(set(bad)))}")
    json.dump(results, open(os.path.join(ROOT, "gpurun_out", "ab.json"), "w"), indent=1)


if __name__ == "__main__":
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    main()
