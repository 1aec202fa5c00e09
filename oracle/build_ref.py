#!/usr/bin/env python3
"""TEST INFRASTRUCTURE — NOT PRODUCT CODE.

Compile the reference's own hot-path sources, from where they lie under
/root/reference, into shared libraries under oracle/_ref/ (git-ignored, shipped
to the GPU box by gpurun like any other built .so):

  oracle/_ref/libref_a<A>_s<S>_b<B>.so   the verbatim `draw` kernel of
        Source/kernels.cl for one (AA edge, shadow samples, max bounces)
        triple — these three are compile-time constants in the reference
        (kernels.cl:12-14, :316, :343), W and H become run-time.
  oracle/_ref/libref_scene.so            LoadTestModel / load_obj
        (Source/TestModelH.h, Source/Loader.cpp, vendored GLM).

kernels.cl is OpenCL C; there is no OpenCL runtime in this image (no PoCL, no
ICD, no CL headers), so it is made g++-compilable by the six mechanical token
rewrites below plus oracle/cl_shim.h.  The rewritten text is written to a
temporary directory and deleted after compilation: no reference source is
copied into the repository.

Only run where /root/reference exists (the build container).  On the GPU box
the prebuilt libraries are used as they are.
"""
from __future__ import annotations

import os
import re
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REFERENCE = os.environ.get("UOB_REFERENCE", "/root/reference")

# (AA edge, shadow samples, max bounces): HEAD defaults / cfg5, cfg1, cfg2, cfg3
VARIANTS = [(2, 10, 10), (1, 1, 0), (2, 8, 10), (4, 10, 4)]

CXXFLAGS = ["-std=c++17", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-pthread", ]


def _split_top_level(s: str) -> list[str]:
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    out.append(cur)
    return out


def rewrite_kernel(src: str, aa: int, shadow: int, bounces: int) -> str:
    """The six mechanical rewrites."""
    # (1) drop the OpenCL pragma
    src = re.sub(r"^#pragma OPENCL.*$", "", src, flags=re.M)
    # (2) OpenCL vector literals -> constructor calls
    src = src.replace("((float3)UINT_MAX)", "(make_float3((float)UINT_MAX))")
    src = re.sub(r"\(\s*(float3|float4|uint3)\s*\)\s*\(", r"make_\1(", src)
    src = re.sub(r"\(\s*float3\s*\)\s*0\.0f", "make_float3(0.0f)", src)
    # (3) swizzle -> accessor
    src = re.sub(r"\.xyz\b", ".xyz()", src)
    # (4) &rays (pointer to array) -> rays
    src = src.replace("(&rays,", "(rays,")
    # (5) drop the excess third initialiser of the [SPHERES]=2 tables
    def trim(m: re.Match) -> str:
        items = _split_top_level(m.group(2))
        return m.group(1) + "{" + ",".join(items[:2]) + "}" + m.group(3)
    src = re.sub(r"^(constant\s+\w+\s+sphere_\w+\[SPHERES\]\s*=\s*)\{(.*)\}(\s*;)", trim, src, flags=re.M)
    # (6) parameter tokens
    n = 0
    def sub1(pat: str, rep: str) -> None:
        nonlocal src, n
        src, k = re.subn(pat, rep, src, flags=re.M)
        if k != 1:
            raise RuntimeError(f"rewrite {pat!r} matched {k} times")
        n += k
    sub1(r"^#define SCREEN_WIDTH .*$", "#define SCREEN_WIDTH (shim_screen_w)")
    sub1(r"^#define SCREEN_HEIGHT .*$", "#define SCREEN_HEIGHT (shim_screen_h)")
    sub1(r"^constant char rays_x = \d+;", f"constant char rays_x = {aa};")
    sub1(r"^constant char rays_y = \d+;", f"constant char rays_y = {aa};")
    sub1(r"^#define aa_rays \d+", f"#define aa_rays {aa * aa}")
    sub1(r"const short light_sources = \d+;", f"const short light_sources = {shadow};")
    sub1(r"const int bounces = \d+;", f"const int bounces = {bounces};")
    return src


# SURVEY 8d "speed build": the same text with the optimiser let loose (FMA contraction, AVX2) - reported separately
# by bench.py, never used for parity.  x86-64-v3 instead of -march=native: the GPU box's CPU is not this container's.
SPEED_FLAGS = ["-std=c++17", "-O3", "-march=x86-64-v3", "-ffp-contract=fast", "-fPIC", "-shared", "-pthread"]
SPEED_VARIANTS = [(2, 10, 10), (2, 8, 10)]


def build_variant(aa: int, shadow: int, bounces: int, tmp: str, speed: bool = False) -> str:
    with open(os.path.join(REFERENCE, "Source", "kernels.cl")) as f:
        src = rewrite_kernel(f.read(), aa, shadow, bounces)
    inc = os.path.join(tmp, f"k_a{aa}_s{shadow}_b{bounces}.inc")
    with open(inc, "w") as f:
        f.write(src)
    suffix = "_speed" if speed else ""
    out = os.path.join(OUT, f"libref_a{aa}_s{shadow}_b{bounces}{suffix}.so")
    flags = SPEED_FLAGS if speed else CXXFLAGS
    cmd = ["g++", *flags, f'-DREF_KERNEL_INC="{inc}"', f"-DREF_AA={aa}", f"-DREF_SHADOW={shadow}",
           f"-DREF_BOUNCES={bounces}", "-I", HERE, os.path.join(HERE, "ref_driver.cpp"), "-o", out]
    subprocess.check_call(cmd)
    return out


def build_scene() -> str:
    out = os.path.join(OUT, "libref_scene.so")
    cmd = ["g++", *CXXFLAGS, "-I", os.path.join(REFERENCE, "Source"), "-I", os.path.join(REFERENCE, "glm"),
           os.path.join(HERE, "ref_scene.cpp"), "-o", out]
    subprocess.check_call(cmd)
    return out


def build_ocl(tmp: str) -> str:
    """libref_ocl.so: oracle/ref_ocl.c + the UNMODIFIED kernels.cl text embedded as a C string (the GPU box
    has no /root/reference).  Runs the reference kernel through a real OpenCL driver when one is present."""
    with open(os.path.join(REFERENCE, "Source", "kernels.cl")) as f:
        text = f.read()
    inc = os.path.join(tmp, "kernels_cl_string.inc")
    with open(inc, "w") as f:
        for line in text.split("\n"):
            esc = line.replace("\\", "\\\\").replace('"', '\\"').replace("\t", "\\t").replace("\r", "")
            f.write('"' + esc + '\\n"\n')
    out = os.path.join(OUT, "libref_ocl.so")
    cmd = ["gcc", "-O2", "-fPIC", "-shared", f'-DREF_OCL_SOURCE_INC="{inc}"', os.path.join(HERE, "ref_ocl.c"), "-o", out, "-ldl"]
    subprocess.check_call(cmd)
    return out


def main(argv: list[str]) -> int:
    if not os.path.isfile(os.path.join(REFERENCE, "Source", "kernels.cl")):
        print(f"build_ref: {REFERENCE} not present — keeping prebuilt oracle/_ref/*.so", file=sys.stderr)
        return 0
    os.makedirs(OUT, exist_ok=True)
    variants = VARIANTS
    if len(argv) == 4:
        variants = [tuple(int(a) for a in argv[1:4])]
    with tempfile.TemporaryDirectory(prefix="uob_ref_") as tmp:
        for aa, s, b in variants:
            print("built", build_variant(aa, s, b, tmp))
        if len(argv) != 4:
            for aa, s, b in SPEED_VARIANTS:
                print("built", build_variant(aa, s, b, tmp, speed=True))
        print("built", build_ocl(tmp))
    print("built", build_scene())
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
