"""CPU tests of the oracle: the C restatement against the reference itself
(oracle/_ref, when built) and against the committed golden fixtures that were
generated from the reference (tests/golden/make_golden.py)."""
import numpy as np
import pytest


def _render_meta(ob, meta, scene, **kw):
    v, n, c = scene
    rot = ob.oracle_rot_matrix(meta["yaw"], meta["pitch"])
    return ob.oracle_render(meta["W"], meta["H"], meta["aa"], meta["shadow_samples"], meta["max_bounces"], meta["focal"],
                            v, n, c, rot, meta["cam"], meta["light"], **kw)


@pytest.mark.parametrize("name", ["head_256", "head_256_rotated", "cfg1_256", "cfg2_480x270", "cfg3_240x135"])
def test_oracle_matches_golden_frames(ob, golden_scene, golden_frames, golden_meta, name):
    meta = golden_meta["small"][name]
    frame, ctr = _render_meta(ob, meta, golden_scene)
    assert ob.frame_hash(golden_frames[name]) == meta["sha256_16"]  # fixture intact
    assert (frame == golden_frames[name]).all(), "oracle must reproduce the reference frame bit for bit"
    assert ctr["primary_rays"] == meta["W"] * meta["H"] * meta["aa"] ** 2
    assert ctr["pixels"] == meta["W"] * meta["H"]


def test_oracle_matches_golden_mesh_frame(ob, golden_scene, golden_ico2, golden_frames, golden_meta):
    meta = golden_meta["small"]["head_192_ico2"]
    scene = tuple(np.concatenate([a, b]) for a, b in zip(golden_scene, golden_ico2))
    frame, _ = _render_meta(ob, meta, scene)
    assert (frame == golden_frames["head_192_ico2"]).all()


def test_oracle_full_size_hash_cfg1(ob, golden_scene, golden_meta, ray_counts):
    """cfg1 (the reference's CPU-runnable config) at full size: frame hash + ray counters."""
    m = golden_meta["full"]["cfg1"]
    v, n, c = golden_scene
    frame, ctr = ob.oracle_render(m["W"], m["H"], m["aa"], m["shadow_samples"], m["max_bounces"], m["focal"], v, n, c,
                                  ob.oracle_rot_matrix(0, 0), m["cam"], m["light"])
    assert ob.frame_hash(frame) == m["sha256_16"]
    for k in ("rays", "primary_rays", "shadow_rays", "bounce_rays", "closest_tri_tests", "shadow_stage1_tests",
              "shadow_stage2_tests", "sphere_tests"):
        assert ctr[k] == ray_counts["cfg1"][k]


def test_oracle_thread_and_row_invariance(ob, golden_scene, golden_frames, golden_meta):
    """Pixels are independent: any thread count / row subset gives the same pixels."""
    meta = golden_meta["small"]["head_256"]
    one, _ = _render_meta(ob, meta, golden_scene, threads=1, y0=64, y1=96)
    many, _ = _render_meta(ob, meta, golden_scene, threads=5, y0=64, y1=96)
    assert (one == many).all()
    assert (one[64:96] == golden_frames["head_256"][64:96]).all()
    assert (one[:64] == 0).all() and (one[96:] == 0).all()
    strided, ctr, rows = _render_meta(ob, meta, golden_scene, row_step=7, want_row_rays=True)
    assert (strided[::7] == golden_frames["head_256"][::7]).all()
    assert int(rows.sum()) == ctr["rays"] and (rows[1::7] == 0).all()


@pytest.mark.parametrize("variant", [(2, 10, 10), (1, 1, 0), (2, 8, 10), (4, 10, 4)])
def test_oracle_vs_reference_kernel(ob, golden_scene, variant):
    """The restatement against the verbatim kernels.cl compiled by g++ (oracle/_ref), on a
    camera the fixtures do not contain."""
    aa, s, b = variant
    if not ob.ref_available(aa, s, b):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    v, n, c = golden_scene
    W, H = 160, 120
    rot = ob.oracle_rot_matrix(-0.4, 0.15)
    cam, light, f = [0.2, -0.1, -2.9], [0.3, -0.4, -0.6], 1100.0 * aa * H / 1024
    r = ob.ref_render(W, H, aa, s, b, f, v, n, c, rot, cam, light)
    o, _ = ob.oracle_render(W, H, aa, s, b, f, v, n, c, rot, cam, light)
    assert (r == o).all()


def test_reference_scene_matches_golden(ob, golden_scene, golden_ico2):
    if not ob.ref_available():
        pytest.skip("oracle/_ref not built")
    import os
    for got, want in zip(ob.ref_load_test_model(), golden_scene):
        assert got.tobytes() == want.tobytes()
    obj = os.path.join(os.path.dirname(__file__), "golden", "ico2.obj")
    for got, want in zip(ob.ref_load_obj(obj), golden_ico2):
        assert got.tobytes() == want.tobytes()


# ---- known-answer tests for the RNG (must be bit exact: north star item 3) -------

def _xorshift(s):
    s &= 0xFFFFFFFF
    s ^= (s << 13) & 0xFFFFFFFF
    s ^= s >> 17
    s ^= (s << 5) & 0xFFFFFFFF
    return s


def test_xorshift_kat(ob):
    assert list(ob.oracle_xorshift3([1, 2, 0])) == [_xorshift(1), _xorshift(2), 0]
    assert _xorshift(1) == 270369  # x=1: 1^(1<<13)=8193; ^>>17 = 8193; ^<<5 = 8193^262176 = 270369
    rng = np.random.default_rng(1)
    for v in rng.integers(0, 2 ** 32, size=(50, 3), dtype=np.uint64):
        assert list(ob.oracle_xorshift3(v.astype(np.uint32))) == [_xorshift(int(x)) for x in v]


def test_seed_and_float_pixel_id(ob):
    # pixel 0: all three seeds are 0 and stay 0 => jitter is -range/2 forever
    assert list(ob.oracle_seed(0)) == [0, 0, 0]
    assert np.allclose(ob.oracle_crush([0, 0, 0], 0.05), -0.025)
    # seeds: (uint)gid, (uint)(gid*91.0f), (uint)(gid*19.0f) in float arithmetic
    gid = 20_000_001
    f91 = int(np.float32(gid) * np.float32(91.0))
    f19 = int(np.float32(gid) * np.float32(19.0))
    assert list(ob.oracle_seed(gid)) == [_xorshift(gid), _xorshift(f91), _xorshift(f19)]
    # y*W+x is evaluated in float (kernels.cl:380): exact below 2^24, rounds above (8K frames)
    assert ob.oracle_global_id(5, 7, 1024) == 7 * 1024 + 5
    W = 7680
    y, x = 4000, 4001
    want = int(np.float32(np.float32(y) * np.float32(W)) + np.float32(x))
    assert ob.oracle_global_id(x, y, W) == want
    assert want != y * W + x  # the rounding is real at 8K and must be reproduced, not fixed


def test_crush_range(ob):
    hi = ob.oracle_crush([0xFFFFFFFF] * 3, 0.05)
    lo = ob.oracle_crush([0] * 3, 0.05)
    assert np.all(hi <= 0.025 + 1e-9) and np.all(lo == np.float32(-0.025))


def test_refract_tir_gives_nan_ray(ob):
    """TIR branch is dead code (kernels.cl:78): inside glass at a grazing angle the ray is NaN."""
    d = np.array([0.0, 0.8, 0.6], np.float32)
    n = np.array([0.0, 0.0, 1.0], np.float32)
    s, o, med = ob.oracle_bounce(1, d, n, [0, 0, 0], 1.52)
    assert np.isnan(o).all() and med == 1.0
    # and a NaN ray hits nothing
    z = np.load(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "scene_cornell.npz"))
    ids, *_ = ob.oracle_closest_hits([[0, 0, 0]], [o], z["verts"], z["normals"], z["colors"])
    assert ids[0] == -1


def test_reflect_keeps_bias_before_normalise(ob):
    d = np.array([0.6, 0.0, 0.8], np.float32)
    n = np.array([0.0, 0.0, -1.0], np.float32)
    s, o, med = ob.oracle_bounce(0, d, n, [1, 2, 3], 1.52)
    assert np.allclose(o, [0.6, 0.0, -0.8], atol=1e-6) and med == 1.0
    assert np.allclose(s, np.array([1, 2, 3]) + 1e-4 * np.array([0.6, 0.0, -0.8]), atol=1e-6)


def test_light_sequence(ob):
    xs = ob.oracle_light_sequence(400)
    assert xs[0] == np.float32(-0.025) and xs.min() >= -0.5 and xs.max() <= 0.5
    # it turns round: goes left first, later comes back right of its start
    assert xs[:50].min() < -0.4 and xs.max() > 0.4


def test_opencl_reference_library_carries_the_unmodified_kernel_text(ob):
    """oracle/_ref/libref_ocl.so (the reference kernel for NVIDIA's OpenCL driver on the GPU box) embeds kernels.cl
    verbatim; other configs only substitute the reference's own parameter tokens.  No OpenCL call is made here."""
    import os
    if not ob.ref_ocl_available():
        pytest.skip("oracle/_ref/libref_ocl.so not built (run oracle/build_ref.py where /root/reference exists)")
    text = ob.ref_ocl_source()
    assert "kernel void draw(" in text.replace("__kernel", "kernel") and "light_sources = 10" in text
    ref = "/root/reference/Source/kernels.cl"
    if os.path.exists(ref):
        with open(ref) as f:
            want = f.read().replace("\r", "")
        assert text.rstrip("\n") == want.rstrip("\n")
    cfg2 = ob.ref_ocl_source(2, 8, 10, 1920, 1080)
    changed = [(a, b) for a, b in zip(text.split("\n"), cfg2.split("\n")) if a != b]
    assert len(text.split("\n")) == len(cfg2.split("\n")) and len(changed) == 3  # width, height, shadow samples
    assert all(("SCREEN_" in a) or ("light_sources" in a) for a, _ in changed)
    cfg3 = ob.ref_ocl_source(4, 10, 4, 3840, 2160)
    assert "rays_x = 4;" in cfg3 and "#define aa_rays 16" in cfg3 and "const int bounces = 4;" in cfg3
    # with no OpenCL platform in this container the call reports an error instead of crashing
    import uob_raytracer_b200 as u
    sc, cam = u.load_test_model(), u.Camera()
    try:
        ob.ref_ocl_render(128, 4, 2, 10, 10, 100.0, sc.verts, sc.normals, sc.colors, cam.rot(), cam.position, cam.light, frames=1)
    except RuntimeError as e:
        assert "OpenCL" in str(e)


def test_speed_build_of_the_reference_is_a_different_rounding_not_a_different_image(ob, golden_scene):
    """The -O3/AVX2/FMA build that bench.py times as cpu_baseline.speed_build: same image within the north-star
    tolerance, not bit-identical (which is why parity always uses the strict build)."""
    if not ob.ref_speed_available(2, 10, 10):
        pytest.skip("speed build missing or this CPU lacks AVX2/FMA")
    from conftest import assert_within_tolerance
    v, n, c = golden_scene
    rot = ob.oracle_rot_matrix(0.0, 0.0)
    args = (160, 120, 2, 10, 10, 1100.0 * 2 * 120 / 1024, v, n, c, rot, [0, 0, -3.2], [0, -0.5, -0.7])
    strict = ob.ref_render(*args)
    fast = ob.ref_render(*args, speed=True)
    assert_within_tolerance(fast, strict, "speed build vs strict build")
