"""What bounds a 1/8 share of the 1080p frame?  Kernel time for variations of bounces / shadow samples / AA."""
import sys, os
sys.path.insert(0, os.getcwd())
import uob_raytracer_b200 as u
scene = u.load_test_model(); cam = u.Camera()
W, H = 1920, 1080
for stride in (8, 1):
    for (A, S, B) in ((2, 8, 10), (2, 8, 0), (2, 8, 1), (2, 8, 2), (2, 8, 4), (2, 1, 10), (1, 8, 10), (1, 1, 0)):
        row = []
        for split in (False, True):
            if split and A != 2:
                continue
            with u.Renderer(W, H, A, S, B, split_pixels=split, block_stride=stride, block_phase=0) as r:
                r.upload_scene(scene)
                ms = []
                for i in range(6):
                    r.render_device(cam.rot(), cam.position, cam.light, 1100.0 * A * H / 1024); ms.append(r.last_kernel_ms)
                row.append(f"{'split' if split else 'default'} {min(ms) * 1e3:.1f}")
        print(f"1/{stride} A={A} S={S} B={B}: " + "  ".join(row) + " us", flush=True)
