#!/usr/bin/env python3
"""Flatten the named scenes with the reference's own loaders + builders -> oracle/_ref/scenes/*.rtscene.gz

TEST INFRASTRUCTURE ONLY (runs where /root/reference is mounted; the outputs are git-ignored and travel
to the GPU box with the snapshot).  Each scene is loaded by oracle/_ref/libref_* (XML via rapidxml, OBJ via
tiny_obj_loader, textures via stb_image, SAH BVH / TLAS built by the reference) and written out through
ref_flatten (oracle/ref_build/ref_api.cpp).  One subprocess per scene: the reference keeps global state.
"""
import gzip
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
OUT = os.path.join(HERE, "_ref", "scenes")

# name -> (scene kind, xml)
SCENES = {
    "bunny_flat": ("file", "bunny_scene.xml"),                 # BASELINE config 1
    "wok_teapot_flat": ("file", "wok_teapot_scene.xml"),       # BASELINE config 2
    "inside_tlas": ("tlas", "inside_scene.xml"),               # shipped scene (9 BLAS)
    "instanced_tlas": ("tlas", "instanced_scene.xml"),         # BASELINE config 3
    "inside_flat": ("file", "inside_scene.xml"),
    "wok_teapot_tlas": ("tlas", "wok_teapot_scene.xml"),
    # FileScene with its other two accelerators: KD-tree as shipped, uniform grid
    "wok_teapot_kd": ("file_kd", "wok_teapot_scene.xml"),
    "wok_teapot_grid": ("file_grid", "wok_teapot_scene.xml"),
    "bunny_kd": ("file_kd", "bunny_scene.xml"),
    "bunny_grid": ("file_grid", "bunny_scene.xml"),
    "inside_kd": ("file_kd", "inside_scene.xml"),
    "inside_grid": ("file_grid", "inside_scene.xml"),
    # TLASFileScene over per-object KD-trees / grids (TLAS_USE_KDTree / TLAS_USE_Grid)
    # (one scene each: a flattened median-split KD-tree per object is ~10x the size of the BVH scene)
    "inside_tlas_grid": ("tlas_grid", "inside_scene.xml"),
    "instanced_tlas_kd": ("tlas_kd", "instanced_scene.xml"),
}


def bake(name, force=False):
    kind, xml = SCENES[name]
    dst = os.path.join(OUT, name + ".rtscene.gz")
    if os.path.exists(dst) and not force:
        return dst
    os.makedirs(OUT, exist_ok=True)
    tmp = os.path.join(OUT, name + ".rtscene")
    subprocess.run([sys.executable, "-m", "oracle.refhost", "flatten", "pt", kind, xml, tmp], check=True, cwd=REPO)
    with open(tmp, "rb") as f, gzip.open(dst, "wb", compresslevel=6) as g:
        shutil.copyfileobj(f, g)
    os.remove(tmp)
    return dst


if __name__ == "__main__":
    for n in SCENES:
        print(bake(n, force="-f" in sys.argv))
