#!/usr/bin/env python3
"""Generate the committed golden vectors from the REFERENCE'S OWN CODE (oracle/_ref, built headless).

Run here (where /root/reference is mounted):   python tests/golden/make_golden.py
Outputs, per scene kind k in {file, tlas, kd, grid, tlas_kd, tlas_grid} (FileScene+USE_BVH / TLASFileScene+TLAS_USE_BVH /
FileScene+USE_KDTree as shipped / FileScene+USE_Grid / TLASFileScene+TLAS_USE_KDTree / +TLAS_USE_Grid);
`make_golden.py kd grid` regenerates only those:
  tests/golden/golden_<k>.rtscene.gz   the reference's flattened scene (inputs)
  tests/golden/golden_<k>.npz          what the reference computed on it:
      cam_*            two cameras (default, look-at) as 4x3 (camPos, topLeft, topRight, bottomLeft)
      prim<c>_*        FindNearest over every primary ray of camera c: t,u,v,obj,tri,traversed,tested
      shadow_rays / shadow_occluded    IsOccluded over shadow rays from camera-0 hit points
      info_N / info_uv / info_albedo   GetHitInfo + GetAlbedo / GetSkyColor for camera-0 primary hits
      whitted<c>       Whitted accumulator (Renderer::Tick once)
      pt<c>            path-tracer accumulator after FRAMES Ticks from spp = 1
`make_golden.py refit` writes tests/golden/golden_refit.npz: for the `file` and `tlas` scenes, the moved vertices handed to the
reference's own BVH::Refit / BLASBVH::Refit (bvh.cpp:26-43, through oracle/ref_build/ref_api.cpp ref_refit) and the node
array it left behind (file_verts / file_nodes; tlas_blas, tlas_verts / tlas_nodes): pins oracle orc_refit_bvh.
The scene (scenes/golden_scene.xml) uses tiny generated textures so the fixtures stay small.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

W, H, FRAMES = 128, 80, 3
LOOK_AT = ((1.6, 0.9, -1.4), (0.0, -0.4, 1.0))


def refit_golden():
    """the reference's Refit on moved vertices; one subprocess per scene class (the reference keeps global state)"""
    import subprocess
    import tempfile
    import cpu_ray_tracer_b200 as rtb
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import displaced, verts9
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        for kind in ("file", "tlas"):
            fs = rtb.FlatScene.load(os.path.join(HERE, f"golden_{kind}.rtscene.gz"))
            blas = int(np.argmax(fs.blas_table["tri_count"]))
            b = fs.blas_table[blas]
            to, tc = int(b["tri_offset"]), int(b["tri_count"])
            v = verts9(displaced(fs.tris[to:to + tc], 0.04, seed=11))
            np.save(os.path.join(tmp, "v.npy"), v)
            dst = os.path.join(tmp, f"{kind}.rtscene")
            subprocess.run([sys.executable, "-m", "oracle.refhost", "refit_flatten", "pt", kind, "golden_scene.xml",
                            os.path.join(tmp, "v.npy"), str(blas), dst], check=True, cwd=ROOT)
            ref = rtb.FlatScene.load(dst)
            rb = ref.blas_table[blas]
            no, nc = int(rb["node_offset"]), int(rb["node_count"])
            assert np.array_equal(ref.tri_indices, fs.tri_indices) and nc == int(b["node_count"])
            out[f"{kind}_blas"], out[f"{kind}_verts"], out[f"{kind}_nodes"] = np.int32(blas), v, ref.nodes[no:no + nc]
            print("refit", kind, "blas", blas, "nodes", nc)
    np.savez_compressed(os.path.join(HERE, "golden_refit.npz"), **out)


def main():
    if sys.argv[1:] == ["refit"]:
        return refit_golden()
    from oracle.refhost import RefRenderer
    import cpu_ray_tracer_b200 as rtb
    from cpu_ray_tracer_b200 import api

    libkind = {"file": "file", "tlas": "tlas", "kd": "file_kd", "grid": "file_grid", "tlas_kd": "tlas_kd", "tlas_grid": "tlas_grid"}
    for kind in (sys.argv[1:] or list(libkind)):
        out = {}
        wh = RefRenderer("whitted", libkind[kind], "golden_scene.xml", W, H)
        pt = RefRenderer("pt", libkind[kind], "golden_scene.xml", W, H)
        scene_path = os.path.join(HERE, f"golden_{kind}.rtscene")
        pt.flatten(scene_path)
        fs = rtb.FlatScene.load(scene_path)
        if kind == "tlas_kd":
            # 430 804 KD nodes for 3 852 triangles (4.4 MB gzipped): the fixture keeps triangles / transforms / TLAS and the
            # loader rebuilds the per-object trees (FlatScene.rebuild_blas_kdtrees); the vectors below still pin every hit
            # and every traversed / tested counter the reference produced on ITS trees
            fs.save_without_kdtrees(scene_path + ".gz")
        else:
            fs.save(scene_path + ".gz")
        os.remove(scene_path)
        for c in (0, 1):
            if c == 1:
                wh.set_camera(*LOOK_AT)
                pt.set_camera(*LOOK_AT)
            out[f"cam{c}"] = pt.get_camera()
            prim = pt.primary_hits()
            for k in ("t", "u", "v", "obj", "tri", "traversed", "tested"):
                out[f"prim{c}_{k}"] = prim[k]
            if c == 0:
                I = prim["O"] + prim["t"][:, None] * prim["D"]
                m = prim["obj"] >= 0
                L = fs.header["light_pos"][0][None, :] - I[m]
                dist = np.sqrt((L * L).sum(1)).astype(np.float32)
                L = (L / dist[:, None]).astype(np.float32)
                sr = api.make_rays(I[m] + L * np.float32(0.001), L, dist - np.float32(0.002))
                out["shadow_rays"] = sr
                out["shadow_occluded"] = pt.is_occluded(sr["O"].copy(), sr["D"].copy(), sr["tmax"].copy())
                N, uv, alb = pt.hit_info(prim["O"], prim["D"], prim)
                out["info_N"], out["info_uv"], out["info_albedo"] = N, uv, alb
            wh.tick(1)
            out[f"whitted{c}"] = wh.accumulator()
            pt.reset(1)
            pt.tick(FRAMES)
            out[f"pt{c}"] = pt.accumulator()
        out["meta"] = np.array([W, H, FRAMES], np.int32)
        out["look_at"] = np.array(LOOK_AT, np.float32)
        np.savez_compressed(os.path.join(HERE, f"golden_{kind}.npz"), **out)
        print(kind, "tris", fs.triangle_count, "hit fraction", float((out["prim0_obj"] >= 0).mean()))


if __name__ == "__main__":
    main()
