#!/usr/bin/env python3
"""bench.py — headline benchmark of the render path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step is one frame of the workload (default cfg2 of BASELINE.json: Cornell Box 1920x1080, 2x2 AA,
8-sample soft shadows, <=10 mirror/glass bounces) rendered through libuob_rt.so.  Rays per frame
are the ORACLE's deterministic counters for that config (tests/golden/ray_counts.json,
SURVEY.md §8d), never a GPU-side count.

  value        Mrays/s, device time (CUDA events on the launching stream, max over ranks); the
               scene is resident in HBM, the frame stays on the device.
  e2e          the same metric through the reference-facing call rt_render (offload_rendering
               semantics: per-frame arguments, kernel, BLOCKING read-back of the frame into pinned
               host memory), host wall clock.
  roofline     dominant kernel (draw_brute) vs the FP32 non-tensor pipe: algorithmic FLOPs per
               frame (37*T_c + 17*T_s1 + 22*T_s2 + 31*T_sph from the oracle's test
               counters) / measured kernel time, against an FFMA microbenchmark run just before.
  cpu_baseline the reference's own kernels.cl compiled for the host (oracle/_ref, kind
               "reference"; falls back to the C restatement, kind "port") on all host cores.

N > 1: the frame is split into N contiguous row tiles (one per rank, one process per GPU); every
step ends with an in-place NCCL all-gather of the tiles into the whole frame on every rank
(strong scaling: the frame is fixed).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "Mrays/sec (Cornell Box 1080p, AA+soft shadows)"
UNIT = "Mrays/s"


def load_counts(workload: str) -> dict:
    with open(os.path.join(ROOT, "tests", "golden", "ray_counts.json")) as f:
        return json.load(f)[workload]


def load_traffic(workload: str):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/traffic.json)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)[workload]
        return t["dram_bytes_read"] + t["dram_bytes_write"]
    except Exception:
        return None


def load_ncu(workload: str):
    """Pipe / issue utilisation of the dominant kernel from the committed ncu capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)[workload].get("ncu")
    except Exception:
        return None


def algorithmic_flops(c: dict) -> float:
    """minimal-operation form of the reference's brute-force algorithm, FMA = 2."""
    return 37.0 * c["closest_tri_tests"] + 17.0 * c["shadow_stage1_tests"] + 22.0 * c["shadow_stage2_tests"] + \
        31.0 * c["sphere_tests"]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []
        self.nvml, self.handle, self.samples, self.stop_flag = None, None, [], False

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        try:  # the CUDA device's own UUID: immune to CUDA_VISIBLE_DEVICES renumbering
            import torch
            uuid = str(torch.cuda.get_device_properties(self.index).uuid)
            h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
        return pynvml, h

    def _nvml_loop(self):
        nv, h = self.nvml, self.handle
        while not self.stop_flag:  # ~1-2 kHz; the short sleep keeps this thread off the GIL while the main thread launches frames
            try:
                self.samples.append((nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)))
            except Exception:
                pass
            time.sleep(0.0005)

    def start(self):
        # in-process NVML polling at ~1 kHz (a short timed region — 100 frames of 0.1 ms — ends before an nvidia-smi child
        # has printed its first line); nvidia-smi -lms as the fallback
        try:
            self.nvml, self.handle = self._nvml_handle()
            self.thread = threading.Thread(target=self._nvml_loop, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "20"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.nvml is not None:
            self.stop_flag = True
            self.thread.join(timeout=1.0)
            nv, h = self.nvml, self.handle
            source = "NVML, polled during the timed region"
            if not self.samples:  # a timed region of a few ms can end before the first poll returns
                try:
                    self.samples.append((nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)))
                    source = "NVML, one sample right at the end of the timed region (it was shorter than one poll)"
                except Exception:
                    pass
            bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
            sm = [x[0] for x in self.samples]
            reasons = sorted({k for x in self.samples for k, b in bits.items() if x[1] & b})
            try:
                mx, power = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)), round(nv.nvmlDeviceGetPowerUsage(h) / 1000.0, 1)
            except Exception:
                mx, power = None, None
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "power_w": power, "samples": len(sm),
                    "reasons": reasons, "source": source}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
                power.append(float(parts[2]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


_scene_cache = {}


def scene_and_camera(workload: str = "cfg2"):
    """Cornell box (+ for cfg4 the 1,310,720-face icosphere loaded through load_obj, appended as
    skeleton.cpp:102-103 would) and the reference's default camera / light."""
    import uob_raytracer_b200 as u
    mesh = workload == "cfg4"
    if mesh not in _scene_cache:
        scene = u.load_test_model()
        if mesh:
            import tempfile
            path = os.path.join(tempfile.gettempdir(), f"uob_ico8_{os.getpid()}.obj")
            u.write_icosphere_obj(path, 8, 0.2, 0.05)
            scene = scene + u.load_obj(path)
            os.unlink(path)
        _scene_cache[mesh] = scene
    cam = u.Camera()
    return _scene_cache[mesh], cam.rot(), cam.position.copy(), cam.light.copy()


def gpu_ray_counts(cfg, scene, rot, cam4, light4, device=0, row0=0, rows=0) -> dict:
    """Ray counts from the strict kernel's counters (RT_FLAG_COUNT_RAYS) — equal to the oracle's counters on every
    config the oracle can run (tests/test_gpu_parity.py::test_ray_counters_equal_the_oracle); used where the
    oracle cannot run (1.3 M triangles: ~0.1 core-second per pixel)."""
    import uob_raytracer_b200 as u
    with u.Renderer(cfg.width, cfg.height, cfg.aa, cfg.shadow_samples, cfg.max_bounces, device=device, strict=True,
                    count_rays=True, row0=row0, rows=rows) as r:
        r.upload_scene(scene)
        r.render(rot, cam4, light4, cfg.focal)
        return r.ray_counts()


def cpu_reference_frame(cfg, rows_step: int, threads: int = 0, speed: bool = False):
    """One (possibly row-subsampled) frame by the reference's own kernel on host threads.
    Returns (seconds, rays traced, kind)."""
    from oracle import bind as ob
    import uob_raytracer_b200 as u
    scene, rot, cam4, light4 = scene_and_camera(cfg.name)
    if cfg.name == "cfg4":
        # Brute force over 1.3 M triangles costs ~25 ms per ray per core, so the CPU sample is the same scene and camera
        # at 1/64 resolution (30x16 pixels, focal scaled; one row per host thread): ~17 k rays, ~30 s on 16 cores.
        from uob_raytracer_b200.configs import RenderConfig
        small = RenderConfig("cfg4-sample", cfg.width // 64, cfg.height // 64, cfg.aa, cfg.shadow_samples, cfg.max_bounces, "")
        rays = gpu_ray_counts(small, scene, rot, cam4, light4)["rays"]
        t0 = time.perf_counter()
        kind = "reference" if ob.ref_available(cfg.aa, cfg.shadow_samples, cfg.max_bounces) else "port"
        fn = ob.ref_render if kind == "reference" else ob.oracle_render
        fn(small.width, small.height, small.aa, small.shadow_samples, small.max_bounces, small.focal, scene.verts, scene.normals,
           scene.colors, rot, cam4, light4, threads=threads)
        return time.perf_counter() - t0, rays, kind
    counts = load_counts(cfg.name)
    if rows_step == 1:
        rays = counts["rays"]
    else:
        rr = counts.get("row_rays")
        rays = int(sum(rr[::rows_step])) if rr else counts["rays"] / rows_step
    t0 = time.perf_counter()
    if ob.ref_available(cfg.aa, cfg.shadow_samples, cfg.max_bounces):
        kind = "reference"
        ob.ref_render(cfg.width, cfg.height, cfg.aa, cfg.shadow_samples, cfg.max_bounces, cfg.focal, scene.verts,
                      scene.normals, scene.colors, rot, cam4, light4, row_step=rows_step, threads=threads, speed=speed)
    else:
        kind = "port"
        ob.oracle_render(cfg.width, cfg.height, cfg.aa, cfg.shadow_samples, cfg.max_bounces, cfg.focal, scene.verts,
                         scene.normals, scene.colors, rot, cam4, light4, row_step=rows_step, threads=threads)
    return time.perf_counter() - t0, rays, kind


def pick_row_step(cfg) -> int:
    """Bound the CPU sample: a full frame when it takes < ~3 s on this host, else every k-th row."""
    if cfg.name == "cfg4":
        return 540  # fixed two-row sample, see cpu_reference_frame
    t, _, _ = cpu_reference_frame(cfg, 16)
    full = t * 16
    step = 1
    while full / step > 3.0 and step < 64:
        step *= 2
    return step


def run_reference(args, cfg) -> int:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    step = pick_row_step(cfg)
    for _ in range(max(args.warmup, 1)):
        cpu_reference_frame(cfg, step)
    total_t, total_rays, kind = 0.0, 0, "reference"
    for _ in range(args.steps):
        t, rays, kind = cpu_reference_frame(cfg, step)
        total_t += t
        total_rays += rays
    value = total_rays / total_t / 1e6
    sample = (f"{args.steps} frames of {cfg.name}, " + ("all rows" if step == 1 else f"every {step}th row") +
              f" ({total_rays / args.steps:.0f} rays per step), verbatim kernels.cl via g++ shim, -O2 strict IEEE")
    line = {"impl": "reference", "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(total_t / args.steps * 1e3, 3),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg.name, "description": cfg.description, "width": cfg.width, "height": cfg.height,
                       "aa": cfg.aa, "shadow_samples": cfg.shadow_samples, "max_bounces": cfg.max_bounces},
            "cpu_baseline": {"value": round(value, 3), "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": round(value, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


# experiment switch (never set in a measured run): peers keep their pixels in their own frame, to separate the cost of
# the NVLink stores from the rest of the hand-over
_DEBUG_LOCAL_STORES = bool(os.environ.get("UOB_BENCH_DEBUG_LOCAL_STORES"))


def run_ours(args, cfg) -> int:
    import torch
    import uob_raytracer_b200 as u

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the render path has no CPU fallback")
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py: --gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # The north star quotes 1080p AND 4K at every GPU count: a default cfg2 run first takes a short measurement of
    # cfg3 (4K, 16 spp) with the same partition and gather and reports it as "also_4k" in the same JSON line.
    main_cfg = cfg
    also_4k = None
    passes = [(main_cfg, True)]
    if main_cfg.name == "cfg2" and not args.no_4k and not args.strict:
        passes.insert(0, (u.CONFIGS["cfg3"], False))
    for cfg, primary in passes:
        steps = args.steps if primary else max(5, min(args.steps, 20))
        W, H = cfg.width, cfg.height
        from uob_raytracer_b200 import tiles
        row0, rows = tiles.row_tile(H, world, rank)
        scene, rot, cam4, light4 = scene_and_camera(cfg.name)
        if cfg.name == "cfg4":
            counts = gpu_ray_counts(cfg, scene, rot, cam4, light4, device=local_rank)
            rays_source = "strict-kernel ray counters (RT_FLAG_COUNT_RAYS; equal to the oracle's on cfg1/cfg2 — the oracle cannot run 1.3 M triangles)"
        else:
            counts = load_counts(cfg.name)
            rays_source = "oracle counters (tests/golden/ray_counts.json)"

        # Multi-GPU gather of the frame on rank 0:
        #   nccl: contiguous row tiles, in-place NCCL all-gather into every rank's frame (the north-star path)
        #   p2p : 16x16 blocks interleaved over the ranks (near-perfect balance), every rank's draw kernel
        #         stores its pixels straight into rank 0's frame over NVLink (CUDA IPC mapping); the hand-over is
        #         a pair of flags in rank 0's memory (rt_peer_signal / rt_peer_wait: release store after a
        #         system fence, acquire spin) — no collective at all
        #   p2p-nccl: the same stores, ordered by a 4-byte NCCL all-reduce instead of the flags
        gather = args.gather if world > 1 else "none"
        if gather == "auto":
            gather = "p2p"
        stream = torch.cuda.Stream(device=local_rank)  # a real (non-default) stream: the C ABI launches on it
        torch.cuda.set_stream(stream)
        sptr = stream.cuda_stream
        assert sptr != 0
        frame = torch.zeros(H * W, dtype=torch.int32, device=f"cuda:{local_rank}")
        tile = frame[row0 * W:(row0 + rows) * W]
        host = torch.empty(H * W, dtype=torch.int32).pin_memory()
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local_rank}")  # > 126 MB L2
        token = torch.zeros(1, dtype=torch.int32, device=f"cuda:{local_rank}")
        peer_ptr = 0
        fno = [0]  # frame number, the value the hand-over flags count up to
        r = None
        if gather in ("p2p", "p2p-nccl"):
            r = u.Renderer(W, H, cfg.aa, cfg.shadow_samples, cfg.max_bounces, device=local_rank, block_stride=world,
                           block_phase=rank, strict=args.strict)
            handle = [r.ipc_export_frame() if rank == 0 else None]
            dist.broadcast_object_list(handle, src=0)
            ok = torch.ones(1, dtype=torch.int32, device=f"cuda:{local_rank}")
            try:
                target = r.device_frame_ptr if rank == 0 else r.ipc_open_frame(handle[0])
                peer_ptr = 0 if rank == 0 else target
            except u.RtError as exc:  # CUDA IPC not available between these processes
                print(f"bench.py: rank {rank}: {exc}; falling back to --gather nccl", file=sys.stderr)
                ok.zero_()
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()) == 0:
                if peer_ptr:
                    r.ipc_close_frame(peer_ptr)
                    peer_ptr = 0
                r.close()
                r = None
                gather = "nccl"
        if r is None:
            r = u.Renderer(W, H, cfg.aa, cfg.shadow_samples, cfg.max_bounces, device=local_rank, row0=row0, rows=rows,
                           strict=args.strict)
            target = frame.data_ptr()
        r.upload_scene(scene)
        r.set_stream(sptr)

        flags = target + 4 * W * H  # rank 0's hand-over flags: [r] = rank r finished frame f, [32] = rank 0 consumed frame f
        consumed_flag = flags + 4 * 32

        def render_only():
            fno[0] += 1
            if gather == "p2p" and rank != 0:
                r.gate_next_frame(consumed_flag, fno[0] - 1)  # rank 0 is done with the previous frame (checked inside the draw kernel)
            r.render_device(rot, cam4, light4, cfg.focal, dev_ptr=(0 if _DEBUG_LOCAL_STORES else target), stream=sptr)

        def gather_only(consume=True):
            if gather == "nccl":
                dist.all_gather_into_tensor(frame, tile)
            elif gather == "p2p-nccl":
                dist.all_reduce(token)  # orders rank 0's next use of its frame after every peer's stores
            elif gather == "p2p":
                if rank == 0:
                    r.peer_wait(flags + 4, world - 1, fno[0], sptr)
                    if consume:
                        r.peer_signal(consumed_flag, fno[0], sptr)
                else:
                    r.peer_signal(flags + 4 * rank, fno[0], sptr)

        def step_device():
            render_only()
            gather_only()

        fp32_peak = r.measure_fp32_peak() if (rank == 0 and primary) else 0.0

        # sanity (outside every timed region): the gathered frame on rank 0 has every pixel written
        step_device()
        torch.cuda.synchronize()
        if rank == 0:
            if gather in ("p2p", "p2p-nccl"):
                r.read_frame_host_ptr(host.data_ptr())
            else:
                host.copy_(frame)
            torch.cuda.synchronize()
            alpha = (host.numpy().view(np.uint32) >> 24)
            if not (alpha == 255).all() and not _DEBUG_LOCAL_STORES:
                raise SystemExit(f"bench.py: gathered frame incomplete ({int((alpha != 255).sum())} pixels unwritten)")

        # ---- device-timed throughput ------------------------------------------------
        for _ in range(max(args.warmup, 3)):
            step_device()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        launches0 = r.kernel_launches
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
              for _ in range(steps)]
        torch.cuda.synchronize()
        t_wall0 = time.perf_counter()
        for a, k, b in ev:
            flush.zero_()  # L2 flush between timed iterations; outside the event pair
            a.record(stream)
            render_only()
            k.record(stream)
            gather_only()
            b.record(stream)
        torch.cuda.synchronize()
        t_wall = time.perf_counter() - t_wall0
        clocks = sampler.stop() if rank == 0 else {}  # right at the end of the timed region
        launches = r.kernel_launches - launches0
        if dist is not None:
            dist.barrier()
        step_ms = [a.elapsed_time(b) for a, _, b in ev]
        kern_ms = [a.elapsed_time(k) for a, k, _ in ev]
        total_ms = torch.tensor([sum(step_ms), sum(kern_ms)], dtype=torch.float64, device=f"cuda:{local_rank}")
        if dist is not None:
            dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        total_step_ms, total_kern_ms = (float(x) for x in total_ms.cpu())
        per_rank = [round(sum(kern_ms) / steps, 4)]
        if dist is not None:  # every rank's own kernel time (device events): shows imbalance / hand-over waits
            mine = torch.tensor([sum(kern_ms) / steps, sum(step_ms) / steps, r.last_kernel_ms], dtype=torch.float64,
                                device=f"cuda:{local_rank}")
            allr = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(allr, mine)
            # [wait-for-consumed + draw kernel, whole step, draw kernel alone (last frame, the context's own events)]
            per_rank = [[round(float(t[0]), 4), round(float(t[1]), 4), round(float(t[2]), 4)] for t in allr]

        if not primary:
            if rank == 0:
                ms4 = total_step_ms / steps
                also_4k = {"workload": cfg.name, "description": cfg.description, "width": W, "height": H, "aa": cfg.aa,
                           "shadow_samples": cfg.shadow_samples, "max_bounces": cfg.max_bounces, "rays_per_frame": counts["rays"],
                           "steps": steps, "ms_per_step": round(ms4, 4), "kernel_ms_per_step": round(total_kern_ms / steps, 4),
                           "value": round(counts["rays"] / ms4 / 1e3, 1), "unit": UNIT, "gather": gather,
                           "achieved_tflops_algorithmic": round(algorithmic_flops(counts) / (ms4 * 1e-3) / 1e12, 1)}
            torch.cuda.synchronize()
            r.synchronize()
            if dist is not None:
                dist.barrier()
            if peer_ptr:
                r.ipc_close_frame(peer_ptr)
            if dist is not None:
                dist.barrier()
            r.close()
            del frame, host, flush
            torch.cuda.empty_cache()
            if dist is not None:
                dist.barrier()
            continue

        # ---- end to end through rt_render (host buffers) ------------------------------
        e2e_line = None
        if world == 1:
            hp = host.data_ptr()
            for _ in range(3):
                r.render_host_ptr(rot, cam4, light4, cfg.focal, hp)
            t0 = time.perf_counter()
            for _ in range(args.steps):
                r.render_host_ptr(rot, cam4, light4, cfg.focal, hp)
            t_e2e = (time.perf_counter() - t0) / args.steps
            e2e_line = {"value": round(counts["rays"] / t_e2e / 1e6, 1), "unit": UNIT, "ms_per_step": round(t_e2e * 1e3, 4),
                        "h2d_bytes_per_step": 84, "d2h_bytes_per_step": W * H * 4,
                        "api": "rt_render (blocking: per-frame args + kernel + read-back into pinned host memory)"}
            # the same through the pipelined pair rt_render_begin / rt_render_end (a render loop: the read-back of
            # frame k overlaps the kernel of frame k+1; every frame still lands in host memory)
            host2 = torch.empty(H * W, dtype=torch.int32).pin_memory()
            bufs = [hp, host2.data_ptr()]
            r.render_begin(rot, cam4, light4, cfg.focal, bufs[0])
            for i in range(1, 4):
                r.render_begin(rot, cam4, light4, cfg.focal, bufs[i & 1])
                r.render_end()
            t0 = time.perf_counter()
            for i in range(args.steps):
                r.render_begin(rot, cam4, light4, cfg.focal, bufs[i & 1])
                r.render_end()
            t_pipe = (time.perf_counter() - t0) / args.steps
            r.render_end()
            e2e_line["pipelined"] = {"value": round(counts["rays"] / t_pipe / 1e6, 1), "unit": UNIT, "ms_per_step": round(t_pipe * 1e3, 4),
                                     "api": "rt_render_begin / rt_render_end, two frames in flight"}
        else:
            # every rank: render tile -> all-gather -> rank 0 reads the whole frame back
            def step_e2e():
                render_only()
                gather_only(consume=False)
                if rank == 0:
                    if gather in ("p2p", "p2p-nccl"):
                        r.read_frame_host_ptr(host.data_ptr())  # D2H on the same stream, blocking
                        if gather == "p2p":
                            r.peer_signal(consumed_flag, fno[0], sptr)  # the peers may overwrite the frame now
                    else:
                        host.copy_(frame, non_blocking=True)
                torch.cuda.synchronize()
            for _ in range(3):
                step_e2e()
            dist.barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                step_e2e()
            dist.barrier()
            t_e2e = torch.tensor([(time.perf_counter() - t0) / args.steps], dtype=torch.float64, device=f"cuda:{local_rank}")
            dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
            t_e2e = float(t_e2e.cpu())
            e2e_line = {"value": round(counts["rays"] / t_e2e / 1e6, 1), "unit": UNIT, "ms_per_step": round(t_e2e * 1e3, 4),
                        "h2d_bytes_per_step": 84, "d2h_bytes_per_step": W * H * 4,
                        "api": "rt_render_device per rank + " + ("NCCL all-gather" if gather == "nccl" else "peer stores into rank 0's frame + " + ("flag hand-over" if gather == "p2p" else "4-byte all-reduce")) + " + read-back on rank 0"}

        if rank == 0:
            ms_per_step = total_step_ms / args.steps
            kern_ms_per_step = total_kern_ms / args.steps
            mesh = cfg.name == "cfg4"
            flops = 0.0 if mesh else algorithmic_flops(counts) / world  # per launch (one rank's tile; tiles are near-uniform)
            achieved = flops / (kern_ms_per_step * 1e-3) / 1e12
            nominal = 148 * 128 * 2 * 1.965e9 / 1e12
            line = {
                "metric": METRIC, "value": round(counts["rays"] / ms_per_step / 1e3, 1), "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms_per_step, 4),
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": cfg.name, "description": cfg.description, "width": W, "height": H, "aa": cfg.aa,
                           "shadow_samples": cfg.shadow_samples, "max_bounces": cfg.max_bounces, "focal": cfg.focal,
                           "rays_per_frame": counts["rays"], "rays_source": rays_source, "triangles": scene.n,
                           "partition": (f"{world} row tile(s) of {rows} rows" + (", in-place NCCL all-gather per frame" if world > 1 else ""))
                           if gather not in ("p2p", "p2p-nccl") else f"16x16-pixel blocks interleaved over {world} ranks, peer stores into rank 0's frame over NVLink + " + ("release/acquire flag hand-over" if gather == "p2p" else "4-byte all-reduce") + " per frame",
                           "gather": gather,
                           "l2": "flushed between timed iterations (256 MiB memset outside the timed event pairs); "
                                 "the scene is 3.4 KB and lives in shared memory",
                           "arithmetic": ("RT_FLAG_STRICT_IEEE: reference operation sequence, frames bit-identical to the reference's" if args.strict else
                                          "fast path (FMA, division-free shadow tests), within 1/255 on >= 99.9 % of pixels; "
                                          "bit_exact_mode = the same frame through RT_FLAG_STRICT_IEEE")},
                "also_4k": also_4k,
                "wall_ms_per_step_incl_flush": round(t_wall / args.steps * 1e3, 4),
                "kernel_ms_per_step": round(kern_ms_per_step, 4),
                "per_rank_kernel_step_draw_ms": per_rank,
                "e2e": e2e_line,
                "gpu_launches": int(launches),
                "clocks": clocks,
                "roofline": {"bound": "fp32", "kernel": "draw_fast_kernel<%d,true,false,false>" % (10 if cfg.shadow_samples % 10 == 0 else 8),
                             "achieved": round(achieved, 3), "peak": round(fp32_peak, 3), "unit": "TFLOP/s",
                             "frac": round(achieved / fp32_peak, 4) if fp32_peak else None,
                             "peak_source": "FFMA microbenchmark in this run (rt_measure_fp32_peak); MEASURED_PEAKS.json has no FP32 entry",
                             "peak_nominal": round(nominal, 1), "frac_of_nominal": round(achieved / nominal, 4),
                             "algorithmic_gflop_per_launch": round(flops / 1e9, 3),
                             "traffic": load_traffic(cfg.name), "ncu": load_ncu(cfg.name),
                             "note": "achieved = ALGORITHMIC FLOPs of the reference's brute-force algorithm (SURVEY.md 8d) / kernel time; the "
                                     "kernel executes fewer: conservative culls (tile binning, plane/edge/box culls) skip tests the "
                                     "reference must perform, so frac can exceed 1.  What bounds the kernel is instruction issue and "
                                     "latency (see ncu: issue slots and FMA pipe utilisation), not FMA throughput."},
            }
            if mesh:
                # BVH path: no FLOP bound.  Algorithmic bytes: a binary BVH with 4-triangle leaves needs
                # ceil(log2(N/4)) = 19 node visits x 64 B + 4 triangles x 48 B = 1.4 KB per ray.
                bytes_per_ray = 19 * 64 + 4 * 48
                peaks = {}
                try:
                    with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                        peaks = json.load(f)
                except Exception:
                    pass
                peak = float(peaks.get("hbm_gbs", 6650.0))
                ach = bytes_per_ray * counts["rays"] / world / (kern_ms_per_step * 1e-3) / 1e9
                line["roofline"] = {"bound": "hbm", "kernel": "draw_bvh_kernel<float,8>", "achieved": round(ach, 1), "peak": peak,
                                    "unit": "GB/s", "frac": round(ach / peak, 4),
                                    "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                                    "algorithmic_bytes_per_ray": bytes_per_ray, "traffic": load_traffic(cfg.name), "ncu": load_ncu(cfg.name),
                                    "note": "nodes 42 MB + triangles 63 MB stay in the 126 MB L2 (ncu: DRAM 0.1 %, L2 0.9 %, L1 22 % of peak): the "
                                            "traversal is latency / issue-bound inside the SM, the byte figure is only the SURVEY 8d yardstick"}
            if world == 1 and not args.strict:
                # the bit-exact path (frames identical to the reference's), same frame, device time
                with u.Renderer(W, H, cfg.aa, cfg.shadow_samples, cfg.max_bounces, device=local_rank, strict=True) as rs:
                    rs.upload_scene(scene)
                    ms = []
                    for _ in range(8):
                        rs.render_device(rot, cam4, light4, cfg.focal)
                        ms.append(rs.last_kernel_ms)
                    best = min(ms[3:])
                line["bit_exact_mode"] = {"flag": "RT_FLAG_STRICT_IEEE", "kernel_ms_per_step": round(best, 4),
                                          "value": round(counts["rays"] / best / 1e3, 2), "unit": UNIT,
                                          "note": "same culls, surviving tests and all shading in the reference's IEEE operation sequence"}
            # CPU baseline on this box's host cores (N = 1 only)
            if world == 1 and not args.no_cpu_baseline:
                step = pick_row_step(cfg)
                best, rays, kind = None, 0, "reference"
                t_budget = time.perf_counter()
                for i in range(1 if mesh else 6):
                    t, rays, kind = cpu_reference_frame(cfg, step)
                    best = t if best is None else min(best, t)
                    if time.perf_counter() - t_budget > 15.0:
                        break
                line["cpu_baseline"] = {
                    "value": float(f"{rays / best / 1e6:.4g}"), "unit": UNIT, "cores": os.cpu_count(), "kind": kind,
                    "sample": ("cfg4 scene at 30x16 pixels (1/64 resolution, same camera), all rows" if mesh else f"{cfg.name}, " + ("all rows" if step == 1 else f"every {step}th row")) +
                              f", best of {i + 1} frames; verbatim kernels.cl via g++ shim (-O2, strict IEEE), all host threads",
                    "ms_per_frame": None if mesh else round(best * step * 1e3, 1)}
                if not mesh:
                    try:  # the optimiser-let-loose build of the same text, reported separately (not bit-reproducible)
                        from oracle import bind as ob
                        if ob.ref_speed_available(cfg.aa, cfg.shadow_samples, cfg.max_bounces):
                            tb = min(cpu_reference_frame(cfg, step, speed=True)[0] for _ in range(3))
                            line["cpu_baseline"]["speed_build"] = {"value": round(rays / tb / 1e6, 2), "unit": UNIT,
                                                                   "flags": "-O3 -march=x86-64-v3 -ffp-contract=fast",
                                                                   "ms_per_frame": round(tb * step * 1e3, 1)}
                    except Exception as e:  # noqa: BLE001
                        line["cpu_baseline"]["speed_build"] = {"unavailable": str(e)[:200]}
                    # Same-box GPU yardstick, still the baseline leg: the reference's own OpenCL kernel on THIS B200 through
                    # NVIDIA's OpenCL driver (oracle/ref_ocl.c; parameter tokens substituted for cfg != HEAD).  Reported, never used.
                    try:
                        from oracle import bind as ob
                        if ob.ref_ocl_available():
                            scene_o, rot_o, cam_o, light_o = scene_and_camera(cfg.name)
                            _, info = ob.ref_ocl_render(cfg.width, cfg.height, cfg.aa, cfg.shadow_samples, cfg.max_bounces, cfg.focal,
                                                        scene_o.verts, scene_o.normals, scene_o.colors, rot_o, cam_o, light_o, frames=5)
                            full = load_counts(cfg.name)["rays"]
                            line["cpu_baseline"]["same_gpu_opencl"] = {
                                "what": "reference kernels.cl `draw`, -cl-fast-relaxed-math -cl-mad-enable, NDRange {W,H}/{128,4}",
                                "device": info["device"], "kernel_ms": round(info["kernel_ms"], 4),
                                "offload_rendering_ms": round(info["total_ms"], 4),
                                "value": round(full / info["kernel_ms"] / 1e3, 2), "unit": UNIT}
                    except Exception as e:  # noqa: BLE001 - optional yardstick
                        line["cpu_baseline"]["same_gpu_opencl"] = {"unavailable": str(e)[:200]}
            print(json.dumps(line), flush=True)
        torch.cuda.synchronize()
        r.synchronize()  # surfaces a timed-out rt_peer_wait
        if dist is not None:
            dist.barrier()
        if peer_ptr:
            r.ipc_close_frame(peer_ptr)
        if dist is not None:
            dist.barrier()
        r.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-4k", action="store_true", help="skip the short cfg3 (4K) measurement reported as also_4k")
    ap.add_argument("--strict", action="store_true", help="time the bit-exact path (RT_FLAG_STRICT_IEEE) instead of the fast one")
    ap.add_argument("--gather", choices=["auto", "nccl", "p2p", "p2p-nccl"], default="auto",
                    help="N>1: how the frame reaches rank 0 (auto = p2p)")
    args = ap.parse_args()
    import uob_raytracer_b200 as u
    cfg = u.CONFIGS[args.workload]
    if args.impl == "reference":
        return run_reference(args, cfg)
    return run_ours(args, cfg)


if __name__ == "__main__":
    sys.exit(main())
