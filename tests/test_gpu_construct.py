"""Scene construction on the device (run on the B200 box: pytest -m gpu).

  * rt_scene_create with blas.nodes == NULL: the SAH build AND the traversal layout happen on the GPU, no host round trip.
    The layout read back (rt_scene_download_bvh) is the reference builder's arrays bit for bit; hits, work counters and
    path-traced frames equal the oracle's.
  * rt_build_tlas / rt_scene_create without TLAS nodes: TLASBVH::Build (agglomerative clustering, tlas_bvh.cpp:17-70) on
    the GPU - the same tree node for node as the host restatement, also above the reference's 32 767-instance cap.
  * rt_scene_refit: BVH::Refit (bvh.cpp:26-43) on the traversal layout, with the reference's skipped node 1 by default.
"""
import numpy as np
import pytest

from conftest import biteq, displaced, random_rays, shadow_rays_from
from test_gpu_parity import assert_hits_equal, check_pt

from cpu_ray_tracer_b200 import abi

pytestmark = pytest.mark.gpu


def nodes_equal(got, ref, what, skip_root_box=True):
    assert len(got) == len(ref), f"{what}: {len(got)} nodes instead of {len(ref)}"
    assert np.array_equal(got["left_first"], ref["left_first"]) and np.array_equal(got["tri_count"], ref["tri_count"]), f"{what}: topology differs"
    lo = 1 if skip_root_box else 0   # the device keeps child boxes in the parent: the root's own box is not stored
    assert biteq(got["aabb_min"][lo:], ref["aabb_min"][lo:]) and biteq(got["aabb_max"][lo:], ref["aabb_max"][lo:]), f"{what}: node boxes differ"


def device_built(flat, tlas=False):
    fs = flat.copy()
    fs.device_build = True
    fs.device_tlas = tlas
    return fs


@pytest.mark.parametrize("name", ["golden_file", "golden_tlas"])
def test_device_built_scene_equals_host_built(name, flat_scenes, oracles):
    """same .rtscene, BVHs (and TLAS) built + laid out on the device instead of taken from the file"""
    from cpu_ray_tracer_b200 import api
    from oracle import porthost
    flat, po = flat_scenes(name), oracles(name)
    sc = api.open_scene(device_built(flat, tlas=flat.kind == abi.RT_SCENE_TLAS), counters=True)
    sc.validate()
    info = sc.info()
    assert info["instances"] == len(flat.blas_table) and info["triangle_slots"] == len(flat.tris) and info["stack_entries"] <= 64
    for i, b in enumerate(flat.blas_table):
        no, nc, to, tc = int(b["node_offset"]), int(b["node_count"]), int(b["tri_offset"]), int(b["tri_count"])
        nodes, idx = sc.download_bvh(i)
        nodes_equal(nodes, flat.nodes[no:no + nc], f"{name} BLAS {i}")
        assert np.array_equal(idx, flat.tri_indices[to:to + tc]), f"{name} BLAS {i}: triangle order differs"
    W, H = 320, 192
    for cam in (po.camera_default(W, H), po.camera_look_at((2.0, 1.5, -2.5), (0.0, 0.0, 1.5), W, H)):
        rays = po.primary_rays(cam, W, H)
        ref, _ = po.find_nearest(rays)
        assert_hits_equal(sc.FindNearest(rays), ref, f"{name}, device-built")   # includes traversed / tested: the same trees
    sr = shadow_rays_from(flat, rays, ref)
    assert np.array_equal(sc.IsOccluded(sr), po.is_occluded(sr)[0])
    rr = random_rays(flat, 20000, seed=5)
    assert_hits_equal(sc.FindNearest(rr), po.find_nearest(rr)[0], f"{name}, device-built (random)")
    oacc, ost = po.render_pt(cam, porthost.default_params(abi.RT_INTEGRATOR_PATH, W, H), 1, 2, 1)
    r = api.GpuRenderer(sc, abi.RT_INTEGRATOR_PATH, W, H).Init()
    r.camera.SetCameraState((2.0, 1.5, -2.5), (0.0, 0.0, 1.5))
    r.render(2)
    assert r.counters()["extension_rays"] == ost["extension_rays"]
    check_pt(r.accumulator, oacc, 2, f"{name}, device-built")
    r.close(), sc.close()


def test_device_built_large_mesh_and_instancing():
    """120 k-triangle terrain built on the device inside rt_scene_create; 343 instances of one device-built mesh share it"""
    from cpu_ray_tracer_b200 import api, host_build
    from oracle import porthost
    flat = host_build.flat_scene_from_tris(host_build.terrain_mesh(120000, seed=5))
    sc = api.open_scene(device_built(flat), counters=True)
    nodes, idx = sc.download_bvh(0)
    nodes_equal(nodes, flat.nodes[:int(flat.blas_table[0]["node_count"])], "terrain")
    assert np.array_equal(idx, flat.tri_indices)
    po = porthost.PortOracle(flat)
    rays = po.primary_rays(po.camera_look_at((2.5, 2.0, -3.0), (0.0, 0.0, 2.5), 320, 192), 320, 192)
    assert_hits_equal(sc.FindNearest(rays), po.find_nearest(rays)[0], "terrain, device-built")
    sc.close()
    inst = host_build.instanced_grid(host_build.terrain_mesh(1200, seed=9, size=0.8, height=0.5), 343)
    sc = api.open_scene(device_built(inst, tlas=True), counters=True)
    sc.validate()
    assert sc.info()["meshes"] == 1 and sc.info()["instances"] == 343 and sc.info()["triangle_slots"] == len(inst.tris)
    po = porthost.PortOracle(inst)
    rays = po.primary_rays(po.camera_default(320, 192), 320, 192)
    ref, st = po.find_nearest(rays)
    assert st["blas_entries"] > len(rays) * 0.05
    assert_hits_equal(sc.FindNearest(rays), ref, "343 instances, mesh + TLAS built on the device")
    sc.close()


@pytest.mark.parametrize("mode", ["cluster", "single"])
@pytest.mark.parametrize("n", [1, 2, 3, 7, 63, 64, 65, 257, 1000, 5000])
def test_gpu_tlas_builder_equals_the_reference_clustering(n, mode, monkeypatch):
    """random boxes plus lattice-aligned ones (equal union areas: FindBestMatch keeps the FIRST best candidate)"""
    from cpu_ray_tracer_b200 import api, host_build
    # two kernels build the same tree: a thread-block cluster with the live boxes in distributed shared memory (n >= 64), and one CTA
    monkeypatch.setenv("RT_B200_TLAS_BUILD", mode)
    rng = np.random.default_rng(n)
    lo = rng.uniform(-20, 20, (n, 3)).astype(np.float32)
    lo[: n // 2] = np.round(lo[: n // 2])                      # ties
    hi = lo + rng.choice([0.5, 1.0, 2.0], (n, 3)).astype(np.float32)
    bounds = np.concatenate([lo, hi], 1)
    ref = host_build.build_tlas32(bounds)
    got = api.build_tlas_gpu(bounds)
    assert len(got) == len(ref) == 2 * n
    for f in ("left", "right"):
        assert np.array_equal(got[f], ref[f]), f
    assert biteq(got["aabb_min"], ref["aabb_min"]) and biteq(got["aabb_max"], ref["aabb_max"])
    if 2 * n <= 65535:
        r16 = host_build.build_tlas(bounds)                     # the reference's own node format: same tree
        inner = r16["left_right"] != 0
        assert np.array_equal(r16["left_right"][inner] & 0xffff, ref["left"][inner])
        assert np.array_equal(r16["left_right"][inner] >> 16, ref["right"][inner])
        assert np.array_equal(r16["blas"][~inner], ref["right"][~inner]) and (ref["left"][~inner] == 0).all()


def test_more_instances_than_the_reference_can_hold():
    """40 000 instances: above the 2 x 16-bit child indices of TLASBVHNode (tlas_bvh.h:10) and far above nodeIdx[256]
    (tlas_bvh.cpp:21).  TLAS built on the device == host restatement with 32-bit children; both traverse like the oracle."""
    from cpu_ray_tracer_b200 import api, host_build
    from oracle import porthost
    mesh = host_build.terrain_mesh(200, seed=3, size=0.5, height=0.3)
    n = 40000
    host = host_build.instanced_grid(mesh, n, tlas="host32")
    with pytest.raises(ValueError):
        host_build.build_tlas(host.instance_bounds)            # the reference's format cannot hold it
    import os
    for mode in ("cluster", "single"):
        os.environ["RT_B200_TLAS_BUILD"] = mode
        got, ms = api.build_tlas_gpu(host.instance_bounds, return_ms=True)
        assert np.array_equal(got["left"], host.tlas_nodes32["left"]) and np.array_equal(got["right"], host.tlas_nodes32["right"])
        assert biteq(got["aabb_min"], host.tlas_nodes32["aabb_min"]) and biteq(got["aabb_max"], host.tlas_nodes32["aabb_max"])
        print(f"rt_build_tlas [{mode}]: {n} instances in {ms:.1f} ms on the device")
    del os.environ["RT_B200_TLAS_BUILD"]
    po = porthost.PortOracle(host)
    W, H = 256, 128
    rays = po.primary_rays(po.camera_look_at((0.0, 9.0, 40.0), (0.0, 6.0, 20.0), W, H), W, H)   # from behind: the last instances are in front
    ref, st = po.find_nearest(rays)
    assert (ref["obj_idx"] >= 2).mean() > 0.3 and int(ref["obj_idx"].max()) > 32767 + 2
    dev = host.copy()
    dev.device_build, dev.device_tlas, dev.tlas_nodes32 = True, True, None
    for flat, what in ((host, "32-bit TLAS from the host"), (dev, "mesh + TLAS built on the device")):
        sc = api.open_scene(flat, counters=True)
        sc.validate()
        assert_hits_equal(sc.FindNearest(rays), ref, f"{n} instances, {what}")
        sc.close()


@pytest.mark.parametrize("name,all_nodes", [("golden_file", False), ("golden_file", True), ("golden_tlas", False)])
def test_refit_equals_the_reference_refit(name, all_nodes, flat_scenes):
    from cpu_ray_tracer_b200 import api
    from oracle import porthost
    flat = flat_scenes(name).copy()
    sc = api.open_scene(flat, counters=True)
    blas = int(np.argmax(flat.blas_table["tri_count"]))
    b = flat.blas_table[blas]
    no, nc, to, tc = int(b["node_offset"]), int(b["node_count"]), int(b["tri_offset"]), int(b["tri_count"])
    new = displaced(flat.tris[to:to + tc], 0.04, seed=11)
    sc.Refit(blas, new, all_nodes=all_nodes)
    sc.validate()
    want = porthost.refit_bvh(flat.nodes[no:no + nc], new, flat.tri_indices[to:to + tc], all_nodes=all_nodes)
    got, idx = sc.download_bvh(blas)
    nodes_equal(got, want, f"{name} refit")
    if not all_nodes and nc > 1:
        assert biteq(want["aabb_min"][1], flat.nodes["aabb_min"][no + 1]), "node 1 keeps its old box in the reference's Refit (bvh.cpp:28)"
    # the oracle traces the refitted arrays: hits, counters and frames must agree
    flat.tris[to:to + tc] = new
    flat.nodes[no:no + nc] = want
    po = porthost.PortOracle(flat)
    W, H = 320, 192
    rays = po.primary_rays(po.camera_default(W, H), W, H)
    ref, _ = po.find_nearest(rays)
    assert_hits_equal(sc.FindNearest(rays), ref, f"{name} after Refit")
    rr = random_rays(flat, 20000, seed=7)
    assert_hits_equal(sc.FindNearest(rr), po.find_nearest(rr)[0], f"{name} after Refit (random)")
    cam = po.camera_default(W, H)
    oacc, ost = po.render_pt(cam, porthost.default_params(abi.RT_INTEGRATOR_PATH, W, H), 1, 2, 1)
    r = api.GpuRenderer(sc, abi.RT_INTEGRATOR_PATH, W, H).Init()
    r.render(2)
    assert r.counters()["extension_rays"] == ost["extension_rays"]
    check_pt(r.accumulator, oacc, 2, f"{name} after Refit")   # normals / uvs of the moved triangles included
    r.close(), sc.close()


def test_refit_with_tlas_rebuild():
    """instances of one mesh; the mesh moves, Refit + SetTransform + TLASBVH::Build all on the device"""
    from cpu_ray_tracer_b200 import api, host_build
    from oracle import porthost
    mesh = host_build.terrain_mesh(1200, seed=9, size=0.8, height=0.5)
    flat = host_build.instanced_grid(mesh, 200)
    sc = api.open_scene(flat, counters=True)
    new = displaced(flat.tris, 0.15, seed=2)
    sc.Refit(0, new, rebuild_tlas=True)
    sc.validate()
    nodes = porthost.refit_bvh(flat.nodes, new, flat.tri_indices)
    bounds = np.stack([host_build.world_bounds(nodes[0]["aabb_min"], nodes[0]["aabb_max"], b["T"]) for b in flat.blas_table])
    flat.tris[:], flat.nodes[:] = new, nodes
    flat.tlas_nodes = host_build.build_tlas(bounds)
    po = porthost.PortOracle(flat)
    rays = po.primary_rays(po.camera_default(320, 192), 320, 192)
    ref, st = po.find_nearest(rays)
    assert st["blas_entries"] > len(rays) * 0.05
    assert_hits_equal(sc.FindNearest(rays), ref, "refit + TLAS rebuild")
    sc.close()


def test_refit_argument_errors(flat_scenes):
    from cpu_ray_tracer_b200 import api
    flat = flat_scenes("golden_file")
    sc = api.open_scene(flat)
    with pytest.raises(api.RtError) as e:
        sc.Refit(0, flat.tris[:10])
    assert e.value.status == abi.RT_ERR_INVALID
    with pytest.raises(api.RtError) as e:
        sc.Refit(3, flat.tris)
    assert e.value.status == abi.RT_ERR_INVALID
    with pytest.raises(api.RtError) as e:
        sc.Refit(0, flat.tris, rebuild_tlas=True)
    assert e.value.status == abi.RT_ERR_INVALID
    sc.close()
    kd = api.open_scene(flat_scenes("golden_kd"))
    with pytest.raises(api.RtError) as e:
        kd.Refit(0, flat_scenes("golden_kd").tris)
    assert e.value.status == abi.RT_ERR_UNSUPPORTED
    kd.close()


@pytest.mark.parametrize("n_tris", [1, 2, 3, 5, 17])
def test_device_build_of_tiny_meshes(n_tris):
    """edge cases of the device build + layout: a mesh of <= 2 triangles is a single leaf (no fat node at all), 3 triangles give one
    split; flat scene and the same mesh under a device-built TLAS of 1 and 3 instances; refit of the single leaf included"""
    from cpu_ray_tracer_b200 import api, host_build
    from oracle import porthost
    rng = np.random.default_rng(n_tris)
    c = rng.uniform(-0.6, 0.6, (n_tris, 3)).astype(np.float32) + np.array([0, 0.2, 1.5], np.float32)
    v0 = c
    v1 = (c + rng.uniform(0.2, 0.5, (n_tris, 3)).astype(np.float32) * np.array([1, 0, 0], np.float32)).astype(np.float32)
    v2 = (c + rng.uniform(0.2, 0.5, (n_tris, 3)).astype(np.float32) * np.array([0, 1, 0], np.float32)).astype(np.float32)
    tris = host_build.make_tris(v0, v1, v2)
    flat = host_build.flat_scene_from_tris(tris)
    po = porthost.PortOracle(flat)
    W, H = 192, 112
    rays = po.primary_rays(po.camera_default(W, H), W, H)
    ref, _ = po.find_nearest(rays)
    assert (ref["obj_idx"] >= 2).any()
    sc = api.open_scene(device_built(flat), counters=True)
    sc.validate()
    nodes, idx = sc.download_bvh(0)
    want = flat.nodes[:int(flat.blas_table[0]["node_count"])]
    nodes_equal(nodes, want, f"{n_tris} triangles")
    assert np.array_equal(idx, flat.tri_indices)
    assert_hits_equal(sc.FindNearest(rays), ref, f"{n_tris} triangles, device-built")
    # move the triangles: Refit on a mesh that may be a single leaf
    moved = displaced(tris, 0.05, seed=3)
    sc.Refit(0, moved, all_nodes=True)
    sc.validate()
    flat2 = flat.copy()
    flat2.tris[:] = moved
    flat2.nodes[:len(want)] = porthost.refit_bvh(want, moved, flat.tri_indices, all_nodes=True)
    po2 = porthost.PortOracle(flat2)
    assert_hits_equal(sc.FindNearest(rays), po2.find_nearest(rays)[0], f"{n_tris} triangles after Refit")
    sc.close()
    for n_inst in (1, 3):
        inst = host_build.instanced_grid(tris, n_inst, spacing=1.4)
        poi = porthost.PortOracle(inst)
        r = poi.primary_rays(poi.camera_default(W, H), W, H)
        sci = api.open_scene(device_built(inst, tlas=True), counters=True)
        sci.validate()
        assert_hits_equal(sci.FindNearest(r), poi.find_nearest(r)[0], f"{n_tris} triangles x {n_inst} instances, all built on the device")
        if n_inst == 3:
            sci.Refit(0, moved, rebuild_tlas=True)   # single-leaf meshes keep their root box on the host
            sci.validate()
            nodes2 = porthost.refit_bvh(inst.nodes, moved, inst.tri_indices)
            inst.tris[:], inst.nodes[:] = moved, nodes2
            bounds = np.stack([host_build.world_bounds(nodes2[0]["aabb_min"], nodes2[0]["aabb_max"], b["T"]) for b in inst.blas_table])
            inst.tlas_nodes = host_build.build_tlas(bounds)
            poj = porthost.PortOracle(inst)
            assert_hits_equal(sci.FindNearest(r), poj.find_nearest(r)[0], f"{n_tris} triangles x 3 instances after Refit + TLAS rebuild")
        sci.close()
