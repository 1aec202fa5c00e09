#!/usr/bin/env python3
"""Generate the golden fixtures of tests/golden/ from THE REFERENCE ITSELF.

Run in the build container (needs /root/reference):  python tests/golden/make_golden.py

Every fixture is produced by code compiled from the reference's own sources
where they lie (oracle/build_ref.py -> oracle/_ref/): frames by the verbatim
`draw` kernel of Source/kernels.cl, scenes by LoadTestModel / load_obj of
Source/TestModelH.h / Loader.cpp.  The ray/test counters are the only part that
comes from the C restatement (the reference has no counters); the restatement
is required to match the reference's frames bit for bit first.

Fixtures (all small):
  scene_cornell.npz      the 26-triangle Cornell Box as uploaded (skeleton.cpp:474-484)
  scene_ico2.npz         load_obj output for ico2.obj (320 faces)
  ico2.obj               the OBJ itself (written by the product's generator, frozen here)
  frames_small.npz       reference frames of every config variant at reduced size
  frame_hashes.json      sha256[:16] of the reference's full-size frames + frame metadata
  ray_counts.json        oracle ray/test counters per config at full size (+ per-row rays for cfg2)
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))

from oracle import bind as ob  # noqa: E402

CAM = [0.0, 0.0, -3.2]
LIGHT = [0.0, -0.5, -0.7]

# name: (W, H, aa, shadow samples, bounces, yaw, pitch, light_x)
SMALL = {
    "head_256": (256, 256, 2, 10, 10, 0.0, 0.0, 0.0),
    "head_256_rotated": (256, 256, 2, 10, 10, 0.25, -0.1, -0.3),
    "cfg1_256": (256, 256, 1, 1, 0, 0.0, 0.0, 0.0),
    "cfg2_480x270": (480, 270, 2, 8, 10, 0.0, 0.0, 0.0),
    "cfg3_240x135": (240, 135, 4, 10, 4, 0.0, 0.0, 0.0),
}
FULL = {
    "head": (1024, 1024, 2, 10, 10),
    "cfg1": (1024, 1024, 1, 1, 0),
    "cfg2": (1920, 1080, 2, 8, 10),
}
COUNTS_ONLY = {
    "cfg3": (3840, 2160, 4, 10, 4),
    "cfg5": (7680, 4320, 2, 10, 10),
}


def focal(aa, H):
    return 1100.0 * aa * H / 1024.0


def main():
    import subprocess
    subprocess.check_call([sys.executable, os.path.join(ROOT, "oracle", "build_ref.py")])
    v, n, c = ob.ref_load_test_model()
    np.savez_compressed(os.path.join(HERE, "scene_cornell.npz"), verts=v, normals=n, colors=c)

    obj = os.path.join(HERE, "ico2.obj")
    if not os.path.exists(obj):
        import uob_raytracer_b200 as u
        u.write_icosphere_obj(obj, 2, 0.2, 0.05)
    ov, on, oc = ob.ref_load_obj(obj)
    np.savez_compressed(os.path.join(HERE, "scene_ico2.npz"), verts=ov, normals=on, colors=oc)

    frames, meta = {}, {}
    for name, (W, H, A, S, B, yaw, pitch, lx) in SMALL.items():
        rot = ob.oracle_rot_matrix(yaw, pitch)
        light = [lx, LIGHT[1], LIGHT[2]]
        f = focal(A, H)
        frames[name] = ob.ref_render(W, H, A, S, B, f, v, n, c, rot, CAM, light)
        meta[name] = dict(W=W, H=H, aa=A, shadow_samples=S, max_bounces=B, yaw=yaw, pitch=pitch, light=light, cam=CAM,
                          focal=f, sha256_16=ob.frame_hash(frames[name]))
        o, _ = ob.oracle_render(W, H, A, S, B, f, v, n, c, rot, CAM, light)
        assert (o == frames[name]).all(), f"oracle != reference on {name}"
        print(name, meta[name]["sha256_16"])
    # a mesh inside the box (Loader.cpp call site skeleton.cpp:102-103): box + icosphere
    mv, mn, mc = np.concatenate([v, ov]), np.concatenate([n, on]), np.concatenate([c, oc])
    W, H, A, S, B = 192, 192, 2, 10, 10
    f = focal(A, H)
    rot = ob.oracle_rot_matrix(0.0, 0.0)
    frames["head_192_ico2"] = ob.ref_render(W, H, A, S, B, f, mv, mn, mc, rot, CAM, LIGHT)
    meta["head_192_ico2"] = dict(W=W, H=H, aa=A, shadow_samples=S, max_bounces=B, yaw=0.0, pitch=0.0, light=LIGHT, cam=CAM,
                                 focal=f, sha256_16=ob.frame_hash(frames["head_192_ico2"]), scene="cornell+ico2")
    np.savez_compressed(os.path.join(HERE, "frames_small.npz"), **frames)

    hashes = {"small": meta, "full": {}}
    counts = {}
    rot = ob.oracle_rot_matrix(0.0, 0.0)
    for name, (W, H, A, S, B) in FULL.items():
        f = focal(A, H)
        r = ob.ref_render(W, H, A, S, B, f, v, n, c, rot, CAM, LIGHT)
        o, ctr, rows = ob.oracle_render(W, H, A, S, B, f, v, n, c, rot, CAM, LIGHT, want_row_rays=True)
        assert (o == r).all(), f"oracle != reference on {name}"
        hashes["full"][name] = dict(W=W, H=H, aa=A, shadow_samples=S, max_bounces=B, focal=f, cam=CAM, light=LIGHT,
                                    sha256_16=ob.frame_hash(r))
        counts[name] = dict(W=W, H=H, aa=A, shadow_samples=S, max_bounces=B, focal=f, **ctr)
        if name == "cfg2":
            counts[name]["row_rays"] = [int(x) for x in rows]
        print(name, hashes["full"][name]["sha256_16"], ctr["rays"])
    for name, (W, H, A, S, B) in COUNTS_ONLY.items():
        f = focal(A, H)
        o, ctr = ob.oracle_render(W, H, A, S, B, f, v, n, c, rot, CAM, LIGHT)
        counts[name] = dict(W=W, H=H, aa=A, shadow_samples=S, max_bounces=B, focal=f, oracle_sha256_16=ob.frame_hash(o), **ctr)
        print(name, ctr["rays"])
    with open(os.path.join(HERE, "frame_hashes.json"), "w") as fh:
        json.dump(hashes, fh, indent=1)
    with open(os.path.join(HERE, "ray_counts.json"), "w") as fh:
        json.dump(counts, fh, indent=1)


if __name__ == "__main__":
    main()
