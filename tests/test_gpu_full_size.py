"""Parity at the BASELINE.json sizes, and the multi-frame / deep-tree cases the small-image tests cannot show
(run on the B200 box: pytest -m gpu).

  * C2 (wok+teapot, FileScene BVH) and C3 (instanced TLAS scene, inside_scene): ONE FULL 1920x1080 frame of the path tracer,
    accumulator bit-identical to the oracle's (the oracle is pinned to the reference: tests/test_oracle_pinned.py);
  * a 64-frame job in ONE call == the reference's 64 Ticks, bit for bit (frame-ordered accumulation, k_sum_frames), also when
    the image budget forces the job into several launches;
  * C5: 2^20 rays through a 10 M-triangle mesh (SAH BVH built on the GPU, bit-identical to the reference builder), hits bit-exact;
  * trees deeper than the traversal stack are refused with RT_ERR_UNSUPPORTED (the reference overflows BVHNode* stack[64] silently,
    bvh.cpp:227); the deepest tree that fits is traced bit-exactly.
"""
import os

import numpy as np
import pytest

from conftest import baked_scenes, biteq, random_rays
from test_gpu_parity import assert_hits_equal, check_pt

from cpu_ray_tracer_b200 import abi

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["wok_teapot_flat", "instanced_tlas", "inside_tlas"])
def test_one_full_1080p_frame_is_bit_identical(name, oracles, gpu_scenes):
    from cpu_ray_tracer_b200 import api
    from oracle import porthost
    if name not in baked_scenes():
        pytest.skip(f"{name} is baked from the reference's assets where /root/reference is mounted (oracle/_ref/scenes)")
    po, sc = oracles(name), gpu_scenes(name, counters=False)
    W, H = 1920, 1080
    cam = po.camera_default(W, H)
    oacc, ost = po.render_pt(cam, porthost.default_params(abi.RT_INTEGRATOR_PATH, W, H), 1, 1, 1)
    r = api.GpuRenderer(sc, abi.RT_INTEGRATOR_PATH, W, H).Init()
    r.Tick(0)
    c = r.counters()
    assert c["extension_rays"] == ost["extension_rays"] and c["paths"] == ost["paths"] == 1920 * 1072
    check_pt(r.accumulator, oacc, 1, f"{name} 1080p")
    # primary hits of the same full frame
    rays = po.primary_rays(cam, W, H)
    ref, _ = po.find_nearest(rays)
    scc = gpu_scenes(name, counters=True)
    assert_hits_equal(scc.FindNearest(rays), ref, f"{name} 1080p primary rays")
    r.close()


@pytest.mark.parametrize("name,budget_mb", [("golden_file", None), ("golden_tlas", None), ("golden_file", 2), ("golden_kd", 3)])
def test_64_frame_job_equals_64_ticks_bit_for_bit(name, budget_mb, oracles, gpu_scenes, monkeypatch):
    from cpu_ray_tracer_b200 import api
    from oracle import porthost
    po, sc = oracles(name), gpu_scenes(name, counters=False)
    W, H, frames = 320, 192, 64
    if budget_mb:
        monkeypatch.setenv("RT_B200_IMAGE_BUDGET_MB", str(budget_mb))  # 320 x 192 x 16 B = 0.94 MB per image: 2-3 frames per launch
    cam = po.camera_default(W, H)
    oacc, ost = po.render_pt(cam, porthost.default_params(abi.RT_INTEGRATOR_PATH, W, H), 1, frames, 1)
    r = api.GpuRenderer(sc, abi.RT_INTEGRATOR_PATH, W, H).Init()
    r.render(frames, first_spp=1)
    assert r.counters()["extension_rays"] == ost["extension_rays"]
    one = r.accumulator
    assert biteq(one[..., :3], np.asarray(oacc[..., :3], np.float32)), f"{name}: render(64) differs from the reference's 64 Ticks"
    # and it is reproducible: the same call again, and frame by frame
    r.ClearAccumulator()
    r.render(frames, first_spp=1)
    assert biteq(r.accumulator, one)
    r.ClearAccumulator()
    for k in range(frames):
        r.render(1, first_spp=1 + k)
    assert biteq(r.accumulator, one)
    r.close()


def test_ten_million_triangle_mesh_hits_bit_exact():
    """BASELINE configs[4] mesh size: 10 M triangles, 2^20 rays (primary from a camera + random incoherent) against the oracle"""
    from cpu_ray_tracer_b200 import api, host_build
    from oracle import porthost
    tris = host_build.terrain_mesh(10_000_000, seed=1)
    fs = host_build.flat_scene_from_tris(tris, builder=api.build_bvh_gpu)
    assert len(fs.tris) >= 10_000_000
    po, sc = porthost.PortOracle(fs), api.open_scene(fs, counters=True)
    cam = po.camera_look_at((0.0, 6.0, -4.0), (0.0, -0.5, 6.0), 1024, 512)
    rays = po.primary_rays(cam, 1024, 512)
    ref, _ = po.find_nearest(rays)
    assert (ref["obj_idx"] >= 2).mean() > 0.3
    assert_hits_equal(sc.FindNearest(rays), ref, "10 M triangles, primary")
    rr = random_rays(fs, 1 << 19, seed=3)
    ref, _ = po.find_nearest(rr)
    assert_hits_equal(sc.FindNearest(rr), ref, "10 M triangles, random")
    occ, _ = po.is_occluded(rr)
    assert np.array_equal(sc.IsOccluded(rr), occ)
    sc.close()


def _caterpillar(levels):
    """a BVH in the reference's layout whose every interior node has a one-triangle leaf on the left and the rest of the
    chain on the right: `levels` levels deep (a degenerate SAH outcome / hand-made worst case)"""
    from cpu_ray_tracer_b200 import host_build
    n = levels  # triangles: one per leaf, the last interior node has two leaves
    x = np.arange(n, dtype=np.float32) * np.float32(0.05) - np.float32(0.025 * n)
    v0 = np.stack([x, np.full(n, -0.3, np.float32), np.full(n, 1.5, np.float32)], 1)
    v1 = v0 + np.array([0.04, 0.0, 0.0], np.float32)
    v2 = v0 + np.array([0.0, 0.6, 0.0], np.float32)
    tris = host_build.make_tris(v0, v1, v2)
    lo = np.minimum(np.minimum(v0, v1), v2)
    hi = np.maximum(np.maximum(v0, v1), v2)
    nodes = np.zeros(2 * n - 1, abi.NODE_DTYPE)
    # node 2k (k < n-1): interior over triangles k..n-1, children 2k+1 (leaf: triangle k) and 2k+2
    for k in range(n - 1):
        i = 2 * k
        nodes[i]["aabb_min"], nodes[i]["aabb_max"] = lo[k:].min(0), hi[k:].max(0)
        nodes[i]["left_first"], nodes[i]["tri_count"] = i + 1, 0
        nodes[i + 1]["aabb_min"], nodes[i + 1]["aabb_max"] = lo[k], hi[k]
        nodes[i + 1]["left_first"], nodes[i + 1]["tri_count"] = k, 1
    last = 2 * (n - 1)
    nodes[last]["aabb_min"], nodes[last]["aabb_max"] = lo[n - 1], hi[n - 1]
    nodes[last]["left_first"], nodes[last]["tri_count"] = n - 1, 1
    idx = np.arange(n, dtype=np.uint32)
    return host_build.flat_scene_from_tris(tris, builder=lambda t: (nodes, idx, 0.0))


def test_trees_deeper_than_the_traversal_stack_are_refused():
    from cpu_ray_tracer_b200 import api
    from oracle import porthost
    # 66 levels: up to 65 pending far children -> refused; 60 levels: fits, and must be traced like the reference does
    with pytest.raises(api.RtError) as e:
        api.open_scene(_caterpillar(66))
    assert e.value.status == abi.RT_ERR_UNSUPPORTED and "deep" in str(e.value)
    fs = _caterpillar(60)
    po, sc = porthost.PortOracle(fs), api.open_scene(fs, counters=True)
    W, H = 256, 128
    # from the far end of the chain, so that rays cross many boxes and the stack fills up
    for cam in (po.camera_default(W, H), po.camera_look_at((3.0, 0.0, 0.9), (-1.5, 0.0, 1.6), W, H)):
        rays = po.primary_rays(cam, W, H)
        ref, _ = po.find_nearest(rays)
        assert_hits_equal(sc.FindNearest(rays), ref, "60-level caterpillar")
    assert (ref["obj_idx"] >= 2).any()
    oacc, ost = po.render_pt(cam, porthost.default_params(abi.RT_INTEGRATOR_PATH, W, H), 1, 2, 1)
    r = api.GpuRenderer(sc, abi.RT_INTEGRATOR_PATH, W, H).Init()
    r.camera.SetCameraState((3.0, 0.0, 0.9), (-1.5, 0.0, 1.6))
    r.render(2)
    assert r.counters()["extension_rays"] == ost["extension_rays"]
    check_pt(r.accumulator, oacc, 2, "60-level caterpillar")
    r.close(), sc.close()


def test_unbalanced_tlas_plus_blas_depth_is_checked():
    """one stack serves both levels: a TLAS chain whose depth + the BLAS depth exceeds the stack is refused, although each level alone fits"""
    from cpu_ray_tracer_b200 import api, host_build
    mesh = host_build.terrain_mesh(1200, seed=9, size=0.8, height=0.5)
    fs = host_build.instanced_grid(mesh, 60)
    n = len(fs.blas_table)
    # replace the agglomerative TLAS by a chain: node 0 = root, interior node k has leaf k (left) and interior k+1 (right)
    leaves = [t for t in fs.tlas_nodes if int(t["left_right"]) == 0]
    assert len(leaves) == n
    t = np.zeros(2 * n - 1, abi.TLAS_NODE_DTYPE)
    lo = np.array([l["aabb_min"] for l in leaves]), np.array([l["aabb_max"] for l in leaves])
    for k in range(n):
        t[n - 1 + k] = leaves[k]            # leaves at n-1 .. 2n-2
    for k in range(n - 1):
        left, right = n - 1 + k, (k + 1 if k + 1 < n - 1 else 2 * n - 2)
        t[k]["aabb_min"], t[k]["aabb_max"] = lo[0][k:].min(0), lo[1][k:].max(0)
        t[k]["left_right"], t[k]["blas"] = left | (right << 16), 0
    fs.tlas_nodes = t
    # depth of the chain = 60 levels, BLAS ~12 levels: 59 + 1 + 11 > 64
    with pytest.raises(api.RtError) as e:
        api.open_scene(fs)
    assert e.value.status == abi.RT_ERR_UNSUPPORTED


def test_whitted_queue_overflow_is_recovered_not_sticky(oracles, gpu_scenes, monkeypatch):
    """a frame whose ray queues overflow is rendered again with larger queues (no dropped rays, no sticky error state)"""
    from cpu_ray_tracer_b200 import api
    from oracle import porthost
    po, sc = oracles("golden_tlas"), gpu_scenes("golden_tlas", counters=False)
    W, H = 192, 112
    cam = po.camera_default(W, H)
    ow, ost = po.render_whitted(cam, porthost.default_params(abi.RT_INTEGRATOR_WHITTED, W, H))
    monkeypatch.setenv("RT_B200_WHITTED_QUEUE_PER_PIXEL", "1.02")  # glass doubles rays: more than 1.02 rays per pixel are alive at depth 1
    r = api.GpuRenderer(sc, abi.RT_INTEGRATOR_WHITTED, W, H).Init()
    for _ in range(2):
        r.Tick(0)
        assert np.nan_to_num(np.abs(r.accumulator - ow)).max() <= 2e-5
    r.close()


def test_scene_closed_before_its_renderer(flat_scenes):
    """rt_scene_destroy while a renderer exists defers the release to the last rt_renderer_destroy"""
    from cpu_ray_tracer_b200 import api
    sc = api.open_scene(flat_scenes("golden_file"))
    r = api.GpuRenderer(sc, abi.RT_INTEGRATOR_PATH, 64, 48).Init()
    r.render(1)
    sc.close()
    r.render(1)          # still valid: the scene lives until the renderer goes
    a = r.accumulator
    assert np.isfinite(a).all() and a[..., :3].sum() > 0
    r.close()
