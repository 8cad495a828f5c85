"""Synthetic scaled scenes (BASELINE configs[3] / configs[4] shapes at sizes the oracle finishes in seconds):
host-built SAH BVH / TLAS (host_build.py, bit-identical to the reference's builders) -> C-ABI -> CUDA, checked
against the oracle.  Instanced scenes share ONE device copy of the mesh (true instancing)."""
import numpy as np
import pytest

from conftest import biteq, random_rays, shadow_rays_from
from test_gpu_parity import assert_hits_equal, check_pt

from cpu_ray_tracer_b200 import abi

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def synth():
    from cpu_ray_tracer_b200 import host_build
    terrain = host_build.terrain_mesh(120000, seed=5)
    flat = host_build.flat_scene_from_tris(terrain)
    inst = host_build.instanced_grid(host_build.terrain_mesh(1200, seed=9, size=0.8, height=0.5), 343)
    small = host_build.instanced_grid(host_build.terrain_mesh(1200, seed=9, size=0.8, height=0.5), 64)
    return {"terrain": flat, "instanced": inst,
            # the other accelerators on host-built scenes; the instanced ones share ONE tree / grid between all instances
            "terrain_kd": host_build.with_accelerator(host_build.flat_scene_from_tris(host_build.terrain_mesh(20000, seed=5)), "kdtree"),
            "terrain_grid": host_build.with_accelerator(flat, "grid"),
            "instanced_kd": host_build.with_accelerator(small, "kdtree"),
            "instanced_grid": host_build.with_accelerator(inst, "grid")}


ALL_SYNTH = ["terrain", "instanced", "terrain_kd", "terrain_grid", "instanced_kd", "instanced_grid"]


@pytest.mark.parametrize("which", ALL_SYNTH)
def test_synthetic_traversal_bit_exact(which, synth):
    from cpu_ray_tracer_b200 import api
    from oracle import porthost
    fs = synth[which]
    po, sc = porthost.PortOracle(fs), api.open_scene(fs, counters=True)
    W, H = 320, 192
    for cam in (po.camera_default(W, H), po.camera_look_at((2.5, 2.0, -3.0), (0.0, 0.0, 2.5), W, H)):
        rays = po.primary_rays(cam, W, H)
        ref, st = po.find_nearest(rays)
        assert (ref["obj_idx"] >= 2).mean() > 0.05
        assert_hits_equal(sc.FindNearest(rays), ref, which)
    sr = shadow_rays_from(fs, rays, ref)
    occ, _ = po.is_occluded(sr)
    assert np.array_equal(sc.IsOccluded(sr), occ)
    rr = random_rays(fs, 30000, seed=13)
    ref, _ = po.find_nearest(rr)
    assert_hits_equal(sc.FindNearest(rr), ref, which + " (random)")
    if which.startswith("instanced"):
        assert st["blas_entries"] > len(rays) * 0.05
    sc.close()


@pytest.mark.parametrize("which", ALL_SYNTH)
def test_synthetic_path_tracer_vs_oracle(which, synth):
    from cpu_ray_tracer_b200 import api
    from oracle import porthost
    fs = synth[which]
    po, sc = porthost.PortOracle(fs), api.open_scene(fs)
    W, H, frames = 160, 96, 2
    cam = po.camera_default(W, H)
    oacc, ost = po.render_pt(cam, porthost.default_params(abi.RT_INTEGRATOR_PATH, W, H), 1, frames, 1)
    r = api.GpuRenderer(sc, abi.RT_INTEGRATOR_PATH, W, H).Init()
    r.render(frames)
    assert r.counters()["extension_rays"] == ost["extension_rays"]
    check_pt(r.accumulator, oacc, frames, which)
    r.close()
    ow, wst = po.render_whitted(cam, porthost.default_params(abi.RT_INTEGRATOR_WHITTED, W, H))
    r = api.GpuRenderer(sc, abi.RT_INTEGRATOR_WHITTED, W, H).Init()
    r.Tick(0)
    c = r.counters()
    assert c["extension_rays"] == wst["extension_rays"] and c["shadow_rays"] == wst["shadow_rays"]
    assert np.nan_to_num(np.abs(r.accumulator - ow)).max() <= 2e-5
    r.close()
    sc.close()


def _soup(n, seed):
    """random triangle soup with clustered sizes, duplicate centroids and degenerate (zero-extent) axes"""
    from cpu_ray_tracer_b200 import host_build
    rng = np.random.default_rng(seed)
    c = rng.uniform(-3, 3, (n, 3)).astype(np.float32)
    c[: n // 8] = c[0]                      # many identical centroids: bins collapse, splits fail
    c[n // 8: n // 4, 1] = np.float32(0.5)  # a slab: one axis has zero centroid extent
    e = rng.uniform(0.01, 0.4, (n, 3, 3)).astype(np.float32)
    return host_build.make_tris(c + e[:, 0], c + e[:, 1], c + e[:, 2])


@pytest.mark.parametrize("case", ["golden", "terrain", "soup_small", "soup", "tiny"])
def test_gpu_bvh_builder_is_bit_identical_to_the_reference_builder(case, flat_scenes):
    """rt_build_bvh (csrc/rt_build.cu) against host/bvh_build.cpp, which tests/test_host_build.py pins to the
    reference's own builder: node boxes, node numbering and triangle order, bit for bit"""
    from cpu_ray_tracer_b200 import api, host_build
    if case == "golden":
        fs = flat_scenes("golden_file")
        sets = [fs.tris] + [flat_scenes("golden_tlas").tris[int(b["tri_offset"]):int(b["tri_offset"]) + int(b["tri_count"])]
                            for b in flat_scenes("golden_tlas").blas_table]
    elif case == "terrain":
        sets = [host_build.terrain_mesh(150000, seed=11)]
    elif case == "soup_small":
        sets = [_soup(n, n) for n in (3, 4, 7, 33, 257, 1000)]
    elif case == "soup":
        sets = [_soup(200000, 5)]
    else:
        sets = [_soup(1, 1), _soup(2, 2)]
    for tris in sets:
        ref_nodes, ref_idx, _ = host_build.build_bvh(tris)
        nodes, idx, ms = api.build_bvh_gpu(tris)
        assert len(nodes) == len(ref_nodes), (case, len(tris))
        assert np.array_equal(idx, ref_idx), f"{case}: triangle order differs ({(idx != ref_idx).sum()} of {len(idx)})"
        assert nodes.tobytes() == ref_nodes.tobytes(), f"{case}: node array differs"
        assert ms >= 0
