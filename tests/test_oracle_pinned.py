"""Pins the oracle (oracle/rt_oracle.c, the C restatement) to the reference:
  (1) against the committed golden vectors the reference's own code produced (tests/golden/make_golden.py);
  (2) where the headless reference build is present (oracle/_ref), against the reference run live.
Everything is bit-exact: same compiler family, -ffp-contract=off on both sides."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, biteq, shadow_rays_from

from cpu_ray_tracer_b200 import abi


def golden(kind):
    return np.load(os.path.join(GOLDEN, f"golden_{kind}.npz"))


def cam_from(po, g, c, W, H):
    cam = po.camera_default(W, H) if c == 0 else po.camera_look_at(g["look_at"][0], g["look_at"][1], W, H)
    got = np.array([list(cam.pos), list(cam.top_left), list(cam.top_right), list(cam.bottom_left)], np.float32)
    assert biteq(got, g[f"cam{c}"]), "camera restatement differs from Camera / SetCameraState"
    return cam


@pytest.mark.parametrize("kind", ["file", "tlas", "kd", "grid", "tlas_kd", "tlas_grid"])
@pytest.mark.parametrize("c", [0, 1])
def test_oracle_find_nearest_vs_golden(oracles, kind, c):
    g = golden(kind)
    W, H, _ = (int(x) for x in g["meta"])
    po = oracles(f"golden_{kind}")
    cam = cam_from(po, g, c, W, H)
    hits, _ = po.find_nearest(po.primary_rays(cam, W, H))
    for ours, theirs in (("t", "t"), ("u", "u"), ("v", "v"), ("obj_idx", "obj"), ("tri_idx", "tri"),
                         ("traversed", "traversed"), ("tested", "tested")):
        assert biteq(hits[ours], g[f"prim{c}_{theirs}"]), ours


@pytest.mark.parametrize("kind", ["file", "tlas", "kd", "grid", "tlas_kd", "tlas_grid"])
def test_oracle_occlusion_and_shading_vs_golden(oracles, kind):
    g = golden(kind)
    W, H, _ = (int(x) for x in g["meta"])
    po = oracles(f"golden_{kind}")
    occ, _ = po.is_occluded(g["shadow_rays"])
    assert np.array_equal(occ, g["shadow_occluded"])
    assert 0.05 < occ.mean() < 0.95
    cam = po.camera_default(W, H)
    rays = po.primary_rays(cam, W, H)
    hits, _ = po.find_nearest(rays)
    N, uv, albedo = po.hit_info(rays, hits)
    assert biteq(N, g["info_N"]) and biteq(uv, g["info_uv"]) and biteq(albedo, g["info_albedo"])


@pytest.mark.parametrize("kind", ["file", "tlas", "kd", "grid", "tlas_kd", "tlas_grid"])
@pytest.mark.parametrize("c", [0, 1])
def test_oracle_integrators_vs_golden(oracles, kind, c):
    from oracle import porthost
    g = golden(kind)
    W, H, frames = (int(x) for x in g["meta"])
    po = oracles(f"golden_{kind}")
    cam = cam_from(po, g, c, W, H)
    acc, st = po.render_whitted(cam, porthost.default_params(abi.RT_INTEGRATOR_WHITTED, W, H))
    assert biteq(acc, g[f"whitted{c}"])
    assert st["shadow_rays"] > 0 and st["extension_rays"] > W * H  # glass + mirror spawn secondary rays
    acc, st = po.render_pt(cam, porthost.default_params(abi.RT_INTEGRATOR_PATH, W, H), 1, frames, 1)
    assert biteq(acc, g[f"pt{c}"])
    assert st["paths"] == W * H * frames


def test_oracle_pt_frame_split_is_exact(oracles):
    """frames are independent given their spp counter: rendering spp 1..3 in one call or as 1, then 2..3
    accumulates identically (this is what sample-index sharding across GPUs relies on)"""
    from oracle import porthost
    po = oracles("golden_file")
    W, H = 64, 48
    cam = po.camera_default(W, H)
    p = porthost.default_params(abi.RT_INTEGRATOR_PATH, W, H)
    a, _ = po.render_pt(cam, p, 1, 3, 1)
    b, _ = po.render_pt(cam, p, 1, 1, 1)
    b, _ = po.render_pt(cam, p, 2, 2, 1, accumulator=b)
    assert biteq(a, b)


@pytest.mark.parametrize("kind", ["file", "tlas"])
def test_oracle_refit_vs_golden(flat_scenes, kind):
    """orc_refit_bvh == what the reference's own BVH::Refit / BLASBVH::Refit left behind on the same moved vertices
    (tests/golden/golden_refit.npz, written by make_golden.py refit through ref_api.cpp ref_refit), node 1 stale included"""
    from oracle import porthost
    g = np.load(os.path.join(GOLDEN, "golden_refit.npz"))
    fs = flat_scenes(f"golden_{kind}")
    blas = int(g[f"{kind}_blas"])
    b = fs.blas_table[blas]
    no, nc, to, tc = int(b["node_offset"]), int(b["node_count"]), int(b["tri_offset"]), int(b["tri_count"])
    tris = np.array(fs.tris[to:to + tc], copy=True)
    v = g[f"{kind}_verts"]
    tris["v0"], tris["v1"], tris["v2"] = v[:, 0:3], v[:, 3:6], v[:, 6:9]
    got = porthost.refit_bvh(fs.nodes[no:no + nc], tris, fs.tri_indices[to:to + tc])
    want = g[f"{kind}_nodes"]
    assert got.tobytes() == want.tobytes()
    assert got[1].tobytes() == fs.nodes[no + 1].tobytes() and got[2].tobytes() != fs.nodes[no + 2].tobytes()  # bvh.cpp:28 skips node 1
    fixed = porthost.refit_bvh(fs.nodes[no:no + nc], tris, fs.tri_indices[to:to + tc], all_nodes=True)
    assert fixed[1].tobytes() != got[1].tobytes() and fixed[3:].tobytes() == got[3:].tobytes()


# ---- live reference (only where oracle/_ref was built, i.e. where /root/reference is mounted) ----
def _ref_available():
    from oracle import refhost
    return all(refhost.available(i, k) for i in ("pt", "whitted") for k in ("file", "tlas", "file_kd", "file_grid", "tlas_kd", "tlas_grid"))


LIVE = [("pt", "file", "wok_teapot_scene.xml", "wok_teapot_flat"), ("whitted", "file", "bunny_scene.xml", "bunny_flat"),
        ("pt", "tlas", "inside_scene.xml", "inside_tlas"), ("whitted", "tlas", "instanced_scene.xml", "instanced_tlas"),
        # FileScene with the KD-tree it ships with (file_scene.h:10-12) and with the uniform grid
        ("pt", "file_kd", "wok_teapot_scene.xml", "wok_teapot_kd"), ("whitted", "file_kd", "bunny_scene.xml", "bunny_kd"),
        ("pt", "file_grid", "wok_teapot_scene.xml", "wok_teapot_grid"), ("whitted", "file_grid", "inside_scene.xml", "inside_grid"),
        # TLASFileScene over per-object KD-trees / grids
        ("pt", "tlas_kd", "instanced_scene.xml", "instanced_tlas_kd"), ("whitted", "tlas_kd", "instanced_scene.xml", "instanced_tlas_kd"),
        ("pt", "tlas_grid", "inside_scene.xml", "inside_tlas_grid"), ("whitted", "tlas_grid", "inside_scene.xml", "inside_tlas_grid")]


@pytest.mark.parametrize("integ,kind,xml,baked", LIVE)
def test_oracle_vs_live_reference(integ, kind, xml, baked, tmp_path):
    """runs the reference's own Renderer::Tick / FindNearest / IsOccluded in a subprocess (it keeps global
    state) and compares with the oracle on the scene the reference flattened"""
    if not _ref_available():
        pytest.skip("oracle/_ref not built here (needs /root/reference)")
    import subprocess
    import sys
    import cpu_ray_tracer_b200 as rtb
    from oracle import porthost
    from conftest import ROOT, scene_path
    if not os.path.exists(scene_path(baked)):
        pytest.skip("baked scene missing")
    W, H, frames = 192, 112, 2
    out = tmp_path / "ref.npz"
    subprocess.run([sys.executable, "-m", "oracle.refhost", "dump", integ, kind, xml, str(W), str(H), str(out), str(frames)],
                   check=True, cwd=ROOT)
    ref = np.load(out)
    fs = rtb.FlatScene.load(scene_path(baked))
    po = porthost.PortOracle(fs)
    cam = po.camera_default(W, H)
    rays = po.primary_rays(cam, W, H)
    assert biteq(rays["D"], ref["prim_D"])
    hits, _ = po.find_nearest(rays)
    for ours, theirs in (("t", "t"), ("u", "u"), ("v", "v"), ("obj_idx", "obj"), ("tri_idx", "tri"),
                         ("traversed", "traversed"), ("tested", "tested")):
        assert biteq(hits[ours], ref["prim_" + theirs]), ours
    if integ == "pt":
        acc, _ = po.render_pt(cam, porthost.default_params(abi.RT_INTEGRATOR_PATH, W, H), 1, frames, 1)
    else:
        acc, _ = po.render_whitted(cam, porthost.default_params(abi.RT_INTEGRATOR_WHITTED, W, H))
    assert biteq(acc, ref["accumulator"])


def test_oracle_passes_vs_live_reference(tmp_path):
    """Renderer::passes = 2 (two consecutive samples per pixel from the tile's stream, spp advancing by 2 per Tick)"""
    if not _ref_available():
        pytest.skip("oracle/_ref not built here (needs /root/reference)")
    import subprocess
    import sys
    import cpu_ray_tracer_b200 as rtb
    from oracle import porthost
    from conftest import ROOT, scene_path
    W, H, frames = 160, 96, 2
    out = tmp_path / "ref.npz"
    subprocess.run([sys.executable, "-m", "oracle.refhost", "dump", "pt", "file", "wok_teapot_scene.xml", str(W), str(H), str(out), str(frames)],
                   check=True, cwd=ROOT, env=dict(os.environ, RT_REF_PASSES="2"))
    ref = np.load(out)
    po = porthost.PortOracle(rtb.FlatScene.load(scene_path("wok_teapot_flat")))
    acc, st = po.render_pt(po.camera_default(W, H), porthost.default_params(abi.RT_INTEGRATOR_PATH, W, H, passes=2), 1, frames, 2)
    assert st["paths"] == W * (H // 16 * 16) * frames * 2
    assert biteq(acc, ref["accumulator"])
