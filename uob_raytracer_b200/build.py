"""Build the in-tree native library uob_raytracer_b200/libuob_rt.so with nvcc for sm_100a.

nvcc cross-compiles without a GPU; the .so travels to the GPU box with the repo
snapshot (it is git-ignored, not gpurun-ignored).
"""
from __future__ import annotations

import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.environ.get("UOB_RT_LIB") or os.path.join(PKG, "libuob_rt.so")
HOST_LIB = os.path.join(PKG, "libuob_host.so")
HOST_SOURCES = [os.path.join("host", "uob_host.cpp")]
HOST_FLAGS = ["-std=c++17", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-Wall", "-pthread"]

SOURCES = ["rt_api.cu", "rt_draw.cu", "rt_peak.cu", "rt_bvh.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math", "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    return "nvcc"


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    for root, _, files in os.walk(CSRC):
        for f in files:
            if os.path.getmtime(os.path.join(root, f)) > t:
                return True
    inc = os.path.join(PKG, "..", "include", "uob_rt.h")
    return os.path.getmtime(inc) > t


def build_host(force: bool = False) -> str:
    """libuob_host.so: scene sources, camera/light state, framebuffer dump (plain g++)."""
    srcs = [os.path.join(CSRC, s) for s in HOST_SOURCES]
    if not force and os.path.exists(HOST_LIB) and all(os.path.getmtime(s) < os.path.getmtime(HOST_LIB) for s in srcs):
        return HOST_LIB
    subprocess.check_call(["g++", *HOST_FLAGS, *srcs, "-o", HOST_LIB])
    return HOST_LIB


FAST_CHUNKS = [1, 2, 4, 5, 8, 10]


def build(force: bool = False, verbose: bool = False) -> str:
    build_host(force)
    if not force and not needs_build():
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    objdir = os.environ.get("UOB_BUILD_DIR") or os.path.join(PKG, "build")  # kernel-variant experiments keep their objects apart
    os.makedirs(objdir, exist_ok=True)
    extra = os.environ.get("UOB_NVCC_DEFS", "").split()
    jobs = [(src, [], src.replace(".cu", ".o")) for src in SOURCES]
    jobs += [("rt_draw_fast.cu", [f"-DRT_FAST_CH={ch}"], f"rt_draw_fast_ch{ch}.o") for ch in FAST_CHUNKS]

    def compile_one(job):
        src, defs, obj = job
        obj = os.path.join(objdir, obj)
        cmd = [_nvcc(), *NVCC_FLAGS, *extra, *defs, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if out.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src} {defs}:\n{out.stdout}")
        return out.stdout

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        logs = list(ex.map(compile_one, jobs))
    if verbose:
        sys.stderr.write("".join(logs))
    objs = [os.path.join(objdir, j[2]) for j in jobs]
    cmd = [_nvcc(), "-Wno-deprecated-gpu-targets", "-shared", "-o", LIB, *objs, "-lcudart"]
    subprocess.check_call(cmd)
    build_skeleton()
    return LIB


HOST_EXE = os.path.join(PKG, "skeleton_b200")


def build_skeleton() -> str:
    """The reference's host loop on the new path (csrc/host/skeleton_b200.cpp), linked against both libraries."""
    src = os.path.join(CSRC, "host", "skeleton_b200.cpp")
    if os.environ.get("UOB_RT_LIB"):
        return HOST_EXE  # kernel-variant experiment builds do not relink the host program
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", src, "-o", HOST_EXE, "-L", PKG, "-luob_rt", "-luob_host",
                           "-Wl,-rpath,$ORIGIN"])
    return HOST_EXE


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
