"""cfg4 diagnosis: kernel time of each 16-row band (one wave of blocks each -> ~ the slowest block of the band), plus a small PNG of the frame."""
import sys, os, tempfile, zlib, struct
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import uob_raytracer_b200 as u
cfg = u.CONFIGS["cfg4"]
path = os.path.join(tempfile.gettempdir(), "ico8.obj")
u.write_icosphere_obj(path, 8, 0.2, 0.05)
scene = u.load_test_model() + u.load_obj(path)
cam = u.Camera()
def png(path, rgb):
    h, w, _ = rgb.shape
    raw = b"".join(b"\0" + rgb[y].tobytes() for y in range(h))
    def chunk(t, d): return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xffffffff)
    open(path, "wb").write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(raw)) + chunk(b"IEND", b""))
with u.Renderer(cfg.width, cfg.height, cfg.aa, cfg.shadow_samples, cfg.max_bounces) as r:
    r.upload_scene(scene)
    img = np.asarray(r.render(cam.rot(), cam.position, cam.light, cfg.focal)).reshape(cfg.height, cfg.width)
    print("full frame kernel ms", r.last_kernel_ms)
    rgb = np.stack([(img >> 16) & 255, (img >> 8) & 255, img & 255], -1).astype(np.uint8)[::3, ::3]
    png("gpurun_out/cfg4_frame.png", np.ascontiguousarray(rgb))
band = int(sys.argv[1]) if len(sys.argv) > 1 else 16
out = []
for row0 in range(0, cfg.height, band):
    rows = min(band, cfg.height - row0)
    with u.Renderer(cfg.width, cfg.height, cfg.aa, cfg.shadow_samples, cfg.max_bounces, row0=row0, rows=rows) as r:
        r.upload_scene(scene)
        ms = []
        for _ in range(3):
            r.render_device(cam.rot(), cam.position, cam.light, cfg.focal); ms.append(r.last_kernel_ms)
        out.append((row0, min(ms)))
print(" ".join(f"{a}:{b:.3f}" for a, b in out))
print("sum of bands ms", sum(b for _, b in out))
