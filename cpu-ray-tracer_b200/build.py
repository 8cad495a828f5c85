"""Compile the CUDA library in-tree: cpu-ray-tracer_b200/librt_b200.so (sm_100a only).

Flags that matter for parity ("FP discipline"): -fmad=false (no FMA contraction: the
reference's hits are reproduced bit for bit), IEEE division / sqrt (nvcc defaults, stated explicitly),
no --use_fast_math; host code of the same files is compiled with -ffp-contract=off.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "librt_b200.so")
SOURCES = ["rt_scene.cu", "rt_render.cu", "rt_build.cu", "rt_construct.cu", "rt_multi.cu"]
HEADERS = ["rt_device.cuh", "rt_streams8.cuh", "rt_glibc_math.cuh", "rt_internal.h", os.path.join("..", "..", "include", "rt_b200.h")]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
              "-Xcompiler", "-fPIC,-ffp-contract=off,-O2", "-shared", "-cudart", "shared", "--threads", "0"]


def nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def up_to_date():
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return all(os.path.getmtime(d) <= t for d in deps)


HOST_LIB = os.path.join(HERE, "librt_host.so")
HOST_SRC = os.path.join(HERE, "host", "bvh_build.cpp")


def build_host(force=False, verbose=False):
    """host-side builders (reference SAH BVH / TLAS restated): plain g++, strict IEEE like the oracle build"""
    deps = [HOST_SRC, os.path.join(HERE, "..", "include", "rt_b200.h")]
    if not force and os.path.exists(HOST_LIB) and all(os.path.getmtime(d) <= os.path.getmtime(HOST_LIB) for d in deps):
        return HOST_LIB
    cmd = ["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", HOST_SRC, "-o", HOST_LIB]
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    return HOST_LIB


def build(force=False, verbose=False, extra=()):
    build_host(force, verbose)
    if not force and up_to_date():
        return LIB
    cmd = [nvcc(), *NVCC_FLAGS, *extra, *[os.path.join(CSRC, f) for f in SOURCES], "-o", LIB]
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    return LIB


# Build variant of the same library (same ABI, selected with RT_B200_LIB).  bench.py times it beside the default build in a child
# process ("libm.variants" of the bench line); nothing else loads it and the tests run against the default build only.
#   cudamath   CUDA's expf / atan2f / acosf: radiance within the tolerance of tests/test_gpu_parity.py instead of bit-identical
# (Round 1 also built the two halves separately and a float-float expf; round 2 moved the exact routines out of line, which
# removed their cost - profiles/r2_libm_ab.txt - and the experiments with them.)
VARIANTS = {"cudamath": ["-DRT_B200_GLIBC_MATH=0"]}


def variant_path(tag):
    return os.path.join(HERE, f"librt_b200_{tag}.so")


CUDA_MATH_LIB = variant_path("cudamath")


def build_variants(tags=None, force=False, verbose=False):
    """compiles the stale ones of the variants in parallel (one nvcc each)"""
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    procs = []
    for tag in (tags or list(VARIANTS)):
        out = variant_path(tag)
        if not force and os.path.exists(out) and all(os.path.getmtime(d) <= os.path.getmtime(out) for d in deps):
            continue
        cmd = [nvcc(), *NVCC_FLAGS, *VARIANTS[tag], *[os.path.join(CSRC, f) for f in SOURCES], "-o", out]
        if verbose:
            print(" ".join(cmd))
        procs.append((tag, subprocess.Popen(cmd)))
    for tag, pr in procs:
        if pr.wait() != 0:
            raise RuntimeError(f"nvcc failed for the {tag} variant")
    return [variant_path(t) for t in (tags or list(VARIANTS))]


def build_cuda_math(force=False, verbose=False):
    return build_variants(["cudamath"], force, verbose)[0]


if __name__ == "__main__":
    if "--variants" in sys.argv or "--ab-variants" in sys.argv:
        print(build_variants(force=True, verbose=True))
    elif "--cuda-math" in sys.argv:
        print(build_cuda_math(force=True, verbose=True))
    else:
        extra = ["-Xptxas", "-v"] if "-v" in sys.argv else []
        print(build(force=True, verbose=True, extra=extra))
        if "--no-variants" not in sys.argv:
            print(build_variants(verbose=True))  # a stale variant lacks newer exports and cannot be loaded
