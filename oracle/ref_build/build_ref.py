#!/usr/bin/env python3
"""Build the reference's own CPU implementation headless -> oracle/_ref/libref_<integrator>_<scene>.so

TEST INFRASTRUCTURE ONLY.  Nothing in the product path loads these libraries; they are the parity
checker (tests/) and the `--impl reference` / cpu_baseline arm of bench.py.

Recipe:
  1. copy the reference's hot-path sources from /root/reference into a scratch directory (never into
     this repo: the scratch directory is deleted afterwards, only the .so lands in oracle/_ref/);
  2. apply the mechanical MSVC->g++ patches below (each asserts that its pattern matched, so a
     changed reference fails loudly instead of silently building something else);
  3. compile ONE unity translation unit: shim precomp.h + the patched sources + our ref_api.cpp
     with strict IEEE flags (-O2 -ffp-contract=off, no -ffast-math).

Four libraries are produced, because integrator (Whitted | path tracer) and scene class
(FileScene | TLASFileScene) are compile-time choices in the reference
(`2. WhittedStyle/renderer.h:57`, `3. PathTracer/renderer.h:48`).
"""
import os
import re
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(REPO, "oracle", "_ref")
REF = os.environ.get("RT_REFERENCE_DIR", "/root/reference")

TEMPLATE_FILES = ["tmplmath.h", "tmplmath.cpp", "common.h", "ray.h", "camera.h", "primitives.h",
                  "texture.h", "material.h"]
INFRA_FILES = ["kdtree.cpp", "grid.cpp", "blas_grid.cpp", "tlas_grid.cpp", "blas_kdtree.cpp", "tlas_kdtree.cpp", "bvh.h", "bvh.cpp", "blas_bvh.h", "blas_bvh.cpp", "tlas_bvh.h", "tlas_bvh.cpp",
               "grid.h", "blas_grid.h", "tlas_grid.h", "kdtree.h", "blas_kdtree.h", "tlas_kdtree.h",
               "helper.h", "hit_info.h", "model.h", "model.cpp"]
SCENE_FILES = ["base_scene.h", "file_scene.h", "file_scene.cpp", "tlas_file_scene.h", "tlas_file_scene.cpp"]
RENDERER_DIRS = {"whitted": "2. WhittedStyle", "pt": "3. PathTracer"}
SCENE_TYPES = {"file": "FileScene", "tlas": "TLASFileScene"}


def sub(text, pattern, repl, name, count=0, flags=0, min_hits=1):
    new, n = re.subn(pattern, repl, text, count=count, flags=flags)
    if n < min_hits:
        raise RuntimeError(f"patch '{name}' did not match (reference changed?)")
    return new


def patch_sources(d, integrator, big_tlas, accel="bvh", tlas_accel="bvh"):
    def edit(fn, fun):
        p = os.path.join(d, fn)
        with open(p, encoding="utf-8-sig") as f:
            s = f.read()
        s = fun(s)
        with open(p, "w") as f:
            f.write(s)

    # MSVC __m128 member access (tmplmath.cpp:173-190, primitives.h:41-272)
    m128 = lambda s: sub(s, r"\.m128_f32\[", "[", "m128_f32")
    edit("tmplmath.cpp", m128)
    edit("primitives.h", m128)
    # __declspec(align(N)) (ray.h:6, tmplmath.h:632,643)
    edit("ray.h", lambda s: sub(s, r"__declspec\(align\(64\)\) class Ray", "class alignas(64) Ray", "ray align"))
    edit("tmplmath.h", lambda s: sub(s, r"__declspec\(align\((\d+)\)\)", r"alignas(\1)", "mat align", min_hits=2))
    # anonymous structs holding a float3 (non-trivial ctor) are rejected by g++ (ray.h:30-32)
    def ray_unions(s):
        for v, pad in (("O", "d0"), ("D", "d1"), ("rD", "d2")):
            s = sub(s, r"union \{ struct \{ float3 %s; float %s; \}; __m128 %s4; \};" % (v, pad, v),
                    "union { float3 %s; struct { float _%s[3]; float %s; }; __m128 %s4; };" % (v, v, pad, v),
                    "ray union " + v)
        return s
    edit("ray.h", ray_unions)
    # same for aabb (tmplmath.h:609-617): keep the layout, drop the nested anonymous structs
    def aabb_union(s):
        s = sub(s, r"union\s*\{\s*struct\s*\{\s*union \{ __m128 bmin4; float bmin\[4\]; struct \{ float3 bmin3; \}; \};\s*"
                   r"union \{ __m128 bmax4; float bmax\[4\]; struct \{ float3 bmax3; \}; \};\s*\};\s*"
                   r"__m128 bounds\[2\] = \{ _mm_setr_ps\( 1e34f, 1e34f, 1e34f, 0 \), _mm_setr_ps\( -1e34f, -1e34f, -1e34f, 0 \) \};\s*\};",
                "union { __m128 bmin4 = _mm_setr_ps( 1e34f, 1e34f, 1e34f, 0 ); float bmin[4]; float3 bmin3; };\n"
                "\tunion { __m128 bmax4 = _mm_setr_ps( -1e34f, -1e34f, -1e34f, 0 ); float bmax[4]; float3 bmax3; };",
                "aabb union")
        return s
    edit("tmplmath.h", aabb_union)
    # extra qualification inside class body (primitives.h:384,534)
    edit("primitives.h", lambda s: sub(s, r"Torus::(Torus\(|GetAlbedo\()", r"\1", "torus qualification", min_hits=2))
    # compile-time screen size becomes the shim's run-time globals (camera.h:4-5)
    edit("camera.h", lambda s: sub(s, r"#define SCRWIDTH\s+\d+\s*\n#define SCRHEIGHT\s+\d+", "", "scr size"))
    # README.md:45-51: choose the BVH in FileScene (ships as KD-tree, file_scene.h:10-12)
    # accel == "kdtree" keeps the shipped configuration
    if accel == "bvh":
        edit("file_scene.h", lambda s: sub(sub(s, r"//#define USE_BVH", "#define USE_BVH", "use bvh"),
                                           r"#define USE_KDTree", "//#define USE_KDTree", "no kdtree"))
    elif accel == "grid":
        edit("file_scene.h", lambda s: sub(sub(s, r"//#define USE_Grid", "#define USE_Grid", "use grid"),
                                           r"#define USE_KDTree", "//#define USE_KDTree", "no kdtree"))
        # resolution / cellSize / gridCells are private (grid.h:25-30); the flattener copies them
        edit("grid.h", lambda s: sub(s, r"private:", "public:", "grid private", min_hits=2))
    # TLASFileScene with BLAS grids / BLAS KD-trees under the same agglomerative TLAS (tlas_file_scene.h:12-14)
    if tlas_accel != "bvh":
        want = {"grid": "TLAS_USE_Grid", "kdtree": "TLAS_USE_KDTree"}[tlas_accel]
        edit("tlas_file_scene.h", lambda s: sub(sub(s, r"#define TLAS_USE_BVH", "//#define TLAS_USE_BVH", "no tlas bvh"),
                                                r"//#define " + want, "#define " + want, "tlas accel"))
        hdrs = {"grid": ("tlas_grid.h", "blas_grid.h"), "kdtree": ("tlas_kdtree.h", "blas_kdtree.h")}[tlas_accel]
        for h in hdrs:  # tlasNode / nodesUsed, BLASGrid::resolution / cellSize / triangles / gridCells are private
            edit(h, lambda s: sub(s, r"private:", "public:", "tlas alt private", min_hits=2))
    # TLAS internals are needed by the flattener (tlas_bvh.h:27 keeps tlasNode private)
    edit("tlas_bvh.h", lambda s: sub(s, r"private:", "public:", "tlas private"))
    # texel array is private (texture.h:98-101); the flattener copies it
    edit("texture.h", lambda s: sub(s, r"private:", "public:", "texture private"))
    if big_tlas:
        # lift the 256-instance cap for synthetic instanced scenes (tlas_bvh.cpp:21)
        edit("tlas_bvh.cpp", lambda s: sub(s, r"int nodeIdx\[256\]", "std::vector<int> nodeIdxV(blasCount + 1); int* nodeIdx = nodeIdxV.data(); int", "tlas cap")
             .replace("int nodeIndices = blasCount", "nodeIndices = blasCount"))
    # renderer.h: scene class and path are a default member initialiser; PrimitiveScene is out of scope
    def renderer_h(s):
        s = sub(s, r'#include "primitive_scene.h"\n', "", "no primitive scene")
        s = sub(s, r'(TLASFileScene|FileScene) scene = (TLASFileScene|FileScene)\("[^"]*"\);',
                "REF_SCENE_TYPE scene = REF_SCENE_TYPE(g_ref_scene_path);", "scene member")
        return s
    edit("renderer.h", renderer_h)
    if integrator == "pt":
        def pt_cpp(s):
            # renderer.cpp:125-126 binds a temporary Ray to Ray& (MSVC extension) and leaves the order of
            # the two jitter draws unspecified.  g++ and MSVC x64 both evaluate right-to-left, i.e. the
            # y jitter takes the first draw; make that explicit.
            s = sub(s, r"accumulator\[x \+ y \* SCRWIDTH\] \+=\s*float4\(Sample\(camera\.GetPrimaryRay\(\(float\)x \+ RandomFloat\(seed\),\s*"
                       r"\(float\)y \+ RandomFloat\(seed\)\), seed\), 0\);",
                    "{ const float jy = RandomFloat(seed); const float jx = RandomFloat(seed);\n"
                    "\t\t\t  Ray primary = camera.GetPrimaryRay((float)x + jx, (float)y + jy);\n"
                    "\t\t\t  accumulator[x + y * SCRWIDTH] += float4(Sample(primary, seed), 0); }", "pt jitter")
            # renderer.cpp:139 caps the job array at 4096 tiles (1080p needs 8040)
            s = sub(s, r"\} tileJob\[4096\];", "} tileJob[65536];", "tile cap")
            return s
        edit("renderer.cpp", pt_cpp)
    else:
        # renderer.cpp:172-178: console clear + prints every frame
        edit("renderer.cpp", lambda s: sub(sub(s, r'system\("cls"\);', "", "cls"), r'\n\s*printf\("(Total|Average|Peak)[^\n]*', "", "prints", min_hits=6))


def build_variant(integrator, scene_kind, cxx="g++", extra_flags=(), suffix="", big_tlas=False, verbose=False, gpuhost=False, accel="bvh", tlas_accel="bvh"):
    os.makedirs(OUT, exist_ok=True)
    d = tempfile.mkdtemp(prefix="ref_build_")
    try:
        for fn in TEMPLATE_FILES:
            shutil.copy(os.path.join(REF, "template", fn), d)
        for fn in INFRA_FILES:
            shutil.copy(os.path.join(REF, "infra", fn), d)
        for fn in SCENE_FILES:
            shutil.copy(os.path.join(REF, "infra", "scene", fn), d)
        for fn in ("renderer.h", "renderer.cpp"):
            shutil.copy(os.path.join(REF, RENDERER_DIRS[integrator], fn), d)
        for fn in os.listdir(d):
            os.chmod(os.path.join(d, fn), 0o644)
        patch_sources(d, integrator, big_tlas, accel, tlas_accel)
        unity = os.path.join(d, "unity.cpp")
        with open(unity, "w") as f:
            f.write('#include "precomp.h"\n')
            for src in ["tmplmath.cpp", "bvh.cpp", "blas_bvh.cpp", "tlas_bvh.cpp", "model.cpp"] + \
                       {"kdtree": ["kdtree.cpp"], "grid": ["grid.cpp"]}.get(accel, []) + \
                       {"kdtree": ["blas_kdtree.cpp", "tlas_kdtree.cpp"], "grid": ["blas_grid.cpp", "tlas_grid.cpp"]}.get(tlas_accel, []) + ["file_scene.cpp", "tlas_file_scene.cpp", "renderer.cpp"]:
                f.write(f'#include "{src}"\n')
            f.write(f'#include "{os.path.join(HERE, "ref_api.cpp")}"\n')
            if gpuhost:
                f.write(f'#include "{os.path.join(HERE, "gpuhost_api.cpp")}"\n')
        out = os.path.join(OUT, f"lib{'gpuhost' if gpuhost else 'ref'}_{integrator}_{scene_kind}{suffix}.so")
        if gpuhost:
            # the C++ drop-in adapters (cpu-ray-tracer_b200/host) + the CUDA library behind the C-ABI
            pkg = os.path.join(REPO, "cpu-ray-tracer_b200")
            extra_flags = (*extra_flags, "-I" + os.path.join(REPO, "include"), "-I" + os.path.join(pkg, "host"),
                           "-L" + pkg, "-lrt_b200", "-Wl,-rpath,$ORIGIN/../../cpu-ray-tracer_b200")
        cmd = [cxx, "-std=c++17", "-O2", "-fopenmp", "-msse4.1", "-ffp-contract=off", "-fpermissive", "-w", "-Wno-psabi",
               "-shared", "-fPIC", *extra_flags,
               f"-DREF_SCENE_TYPE={SCENE_TYPES[scene_kind]}",
               f"-DREF_INTEGRATOR_{integrator.upper()}=1", f"-DREF_SCENE_{scene_kind.upper()}=1",
               "-I" + os.path.join(HERE, "shim"), "-I" + d,
               "-I" + os.path.join(REF, "lib"), "-I" + os.path.join(REF, "lib", "rapidxml-1.13"),
               unity, "-o", out]
        if gpuhost:  # libraries after the object that needs them
            libs = [f for f in cmd if f.startswith(("-L", "-l", "-Wl,"))]
            cmd = [f for f in cmd if f not in libs] + libs
        if verbose:
            print(" ".join(cmd))
        subprocess.run(cmd, check=True)
        return out
    finally:
        shutil.rmtree(d, ignore_errors=True)


def build_all(verbose=False, gpuhost=True):
    if not os.path.isdir(REF):
        raise FileNotFoundError(f"{REF} not present: oracle/_ref can only be (re)built where the reference is mounted")
    jobs = []
    for integ in ("whitted", "pt"):
        for kind in ("file", "tlas"):
            jobs.append(dict(integrator=integ, scene_kind=kind))
    # "reference-like" flags for the CPU baseline: mirrors MSVC /O2 /arch:AVX2 /fp:fast
    # (whitted-style-bvh.vcxproj:93-102); not a parity oracle
    for kind in ("file", "tlas"):
        jobs.append(dict(integrator="pt", scene_kind=kind, extra_flags=("-O3", "-mavx2", "-mfma", "-ffast-math"), suffix="_fast"))
    # FileScene as shipped: KD-tree accelerator (file_scene.h:10-12), and its third option, the uniform grid
    for integ in ("whitted", "pt"):
        jobs.append(dict(integrator=integ, scene_kind="file", suffix="_kd", accel="kdtree"))
        jobs.append(dict(integrator=integ, scene_kind="file", suffix="_grid", accel="grid"))
        # TLASFileScene over per-object KD-trees / grids (TLAS_USE_KDTree / TLAS_USE_Grid)
        jobs.append(dict(integrator=integ, scene_kind="tlas", suffix="_kd", tlas_accel="kdtree"))
        jobs.append(dict(integrator=integ, scene_kind="tlas", suffix="_grid", tlas_accel="grid"))
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:  # g++ subprocesses, one temp dir each
        outs = list(ex.map(lambda kw: build_variant(verbose=verbose, **kw), jobs))
    if gpuhost:
        outs += build_gpuhost(verbose=verbose)
    return outs


def build_gpuhost(verbose=False):
    """libgpuhost_*: the reference's loaders/builders + our C++ adapters + librt_b200.so (drop-in check)"""
    if not os.path.exists(os.path.join(REPO, "cpu-ray-tracer_b200", "librt_b200.so")):
        raise FileNotFoundError("build cpu-ray-tracer_b200/librt_b200.so first (python __graft_entry__.py)")
    jobs = [dict(integrator=integ, scene_kind=kind) for integ in ("whitted", "pt") for kind in ("file", "tlas")]
    # FileScene with the KD-tree it ships with / the grid, through the same adapter header
    jobs += [dict(integrator="pt", scene_kind="file", suffix="_kd", accel="kdtree"),
             dict(integrator="whitted", scene_kind="file", suffix="_grid", accel="grid"),
             dict(integrator="whitted", scene_kind="tlas", suffix="_kd", tlas_accel="kdtree"),
             dict(integrator="pt", scene_kind="tlas", suffix="_grid", tlas_accel="grid")]
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        return list(ex.map(lambda kw: build_variant(verbose=verbose, gpuhost=True, **kw), jobs))


if __name__ == "__main__":
    for o in (build_gpuhost(verbose="-v" in sys.argv) if "gpuhost" in sys.argv else build_all(verbose="-v" in sys.argv)):
        print("built", o)
