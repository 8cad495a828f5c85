"""Host-side builders (cpu-ray-tracer_b200/host/bvh_build.cpp) reproduce the reference's builders bit for bit:
fed the triangles of scenes the reference itself loaded and built (tests/golden, oracle/_ref/scenes), they
must return the reference's node arrays, triangle order, instance inverses, world bounds and TLAS."""
import numpy as np
import pytest

from conftest import all_scene_names, biteq

from cpu_ray_tracer_b200 import abi, host_build


def raw_equal(a, b):
    return a.shape == b.shape and a.tobytes() == b.tobytes()


def _is_alt(name):
    return name.endswith(("_kd", "_grid"))


@pytest.mark.parametrize("name", [n for n in all_scene_names() if n.endswith("_kd") and n != "golden_tlas_kd"])
def test_kdtree_builder_matches_reference(name, flat_scenes):
    """KDTree::Build / BLASKDTree::Build restated on the host: same nodes (boxes, split axis / distance, links), same
    leaf lists, per object for TLAS scenes (golden_tlas_kd is excluded: its trees are rebuilt by this very builder)"""
    fs = flat_scenes(name)
    if fs.blas_kd_table is None:
        nodes, idx, depth = host_build.build_kdtree(fs.tris)
        assert depth <= 20 and len(nodes) == len(fs.kd_nodes)
        assert raw_equal(nodes, fs.kd_nodes), "KD nodes differ from KDTree::Build"
        assert raw_equal(idx, fs.kd_tri_indices), "KD leaf lists differ"
        return
    for b, k in zip(fs.blas_table, fs.blas_kd_table):
        t0, tn = int(b["tri_offset"]), int(b["tri_count"])
        nodes, idx, _ = host_build.build_kdtree(fs.tris[t0:t0 + tn])
        n0, nn, i0, ni = (int(k[f]) for f in ("node_offset", "node_count", "idx_offset", "idx_count"))
        assert raw_equal(nodes, fs.kd_nodes[n0:n0 + nn]) and raw_equal(idx, fs.kd_tri_indices[i0:i0 + ni])


@pytest.mark.parametrize("name", [n for n in all_scene_names() if n.endswith("_grid")])
def test_grid_builder_matches_reference(name, flat_scenes):
    fs = flat_scenes(name)
    if fs.blas_grid_table is None:
        hdr, start, idx = host_build.build_grid(fs.tris)
        assert raw_equal(hdr, fs.grid_header), "resolution / cell size / bounds differ from Grid::Build"
        assert raw_equal(start, fs.grid_cell_start) and raw_equal(idx, fs.grid_tri_indices)
        return
    for b, g in zip(fs.blas_table, fs.blas_grid_table):
        t0, tn = int(b["tri_offset"]), int(b["tri_count"])
        hdr, start, idx = host_build.build_grid(fs.tris[t0:t0 + tn])
        for f in ("resolution", "cell_size", "bounds_min", "bounds_max"):
            assert raw_equal(np.ascontiguousarray(hdr[0][f]), np.ascontiguousarray(g[f])), f
        c0, cn, i0, ni = (int(g[f]) for f in ("cell_offset", "cell_count", "idx_offset", "idx_count"))
        assert raw_equal(start, fs.grid_cell_start[c0:c0 + cn + 1]) and raw_equal(idx, fs.grid_tri_indices[i0:i0 + ni])


def test_with_accelerator_rebuilds_the_same_scene(flat_scenes):
    kd = host_build.with_accelerator(flat_scenes("golden_file"), "kdtree")
    assert kd.kind == abi.RT_SCENE_FLAT_KDTREE and raw_equal(kd.kd_nodes, flat_scenes("golden_kd").kd_nodes)
    gr = host_build.with_accelerator(flat_scenes("golden_file"), "grid")
    assert gr.kind == abi.RT_SCENE_FLAT_GRID and raw_equal(gr.grid_tri_indices, flat_scenes("golden_grid").grid_tri_indices)


@pytest.mark.parametrize("name", [n for n in all_scene_names() if not _is_alt(n)])
def test_sah_builder_matches_reference(name, flat_scenes):
    fs = flat_scenes(name)
    for b in fs.blas_table:
        t0, tn = int(b["tri_offset"]), int(b["tri_count"])
        n0, nn = int(b["node_offset"]), int(b["node_count"])
        tris = fs.tris[t0:t0 + tn]
        nodes, idx, _ = host_build.build_bvh(tris)
        assert len(nodes) == nn
        assert raw_equal(idx, fs.tri_indices[t0:t0 + tn]), "triangle order differs from BVH::Build"
        ref = fs.nodes[n0:n0 + nn]
        for f in ("aabb_min", "aabb_max"):
            assert biteq(nodes[f], ref[f]), f
        assert np.array_equal(nodes["left_first"], ref["left_first"]) and np.array_equal(nodes["tri_count"], ref["tri_count"])


@pytest.mark.parametrize("name", [n for n in all_scene_names() if "tlas" in n])
def test_tlas_builder_and_transforms_match_reference(name, flat_scenes):
    """the agglomerative TLAS is the same for all three BLAS kinds; the leaf boxes are the BLAS' local bounds under T"""
    fs = flat_scenes(name)
    bounds = []
    for i, b in enumerate(fs.blas_table):
        assert biteq(host_build.invert_rigid(b["T"]), b["inv_T"]), "FastInvertedTransformNoScale"
        if fs.blas_kd_table is not None:
            root = fs.kd_nodes[int(fs.blas_kd_table[i]["node_offset"])]
            lo, hi = root["aabb_min"], root["aabb_max"]
        elif fs.blas_grid_table is not None:
            lo, hi = fs.blas_grid_table[i]["bounds_min"], fs.blas_grid_table[i]["bounds_max"]
        else:
            root = fs.nodes[int(b["node_offset"])]
            lo, hi = root["aabb_min"], root["aabb_max"]
        bounds.append(host_build.world_bounds(lo, hi, b["T"]))
    tlas = host_build.build_tlas(np.array(bounds))
    assert len(tlas) == len(fs.tlas_nodes)
    for f in ("aabb_min", "aabb_max"):
        assert biteq(tlas[f], fs.tlas_nodes[f]), f
    assert np.array_equal(tlas["left_right"], fs.tlas_nodes["left_right"])
    leaf = tlas["left_right"] == 0   # interior nodes never set BLAS: the reference leaves malloc garbage there
    assert np.array_equal(tlas["blas"][leaf], fs.tlas_nodes["blas"][leaf])


def test_synthetic_scenes_are_well_formed_and_oracle_traceable():
    from oracle import porthost
    tris = host_build.terrain_mesh(5000, seed=3)
    fs = host_build.flat_scene_from_tris(tris)
    assert fs.kind == abi.RT_SCENE_FLAT and len(fs.nodes) <= 2 * len(tris) - 1
    po = porthost.PortOracle(fs)
    W, H = 64, 40
    hits, st = po.find_nearest(po.primary_rays(po.camera_default(W, H), W, H))
    assert (hits["obj_idx"] >= 2).mean() > 0.1
    inst = host_build.instanced_grid(tris[:600], 27)
    assert inst.kind == abi.RT_SCENE_TLAS and len(inst.blas_table) == 27 and len(inst.tlas_nodes) == 2 * 27
    assert len(inst.tris) == 600                      # ONE copy of the mesh, 27 instances
    po = porthost.PortOracle(inst)
    hits, st = po.find_nearest(po.primary_rays(po.camera_default(W, H), W, H))
    assert (hits["obj_idx"] >= 2).any() and st["blas_entries"] > 0
    with pytest.raises(ValueError):
        host_build.build_tlas(np.zeros((40000, 6), np.float32))   # 2 x 16-bit child indices


def test_swap_partition_closed_form():
    """the closed form the GPU builder (csrc/rt_build.cu) uses for the reference's two-cursor swap partition
    (bvh.cpp:88-95), checked against the loop itself on random inputs"""
    import random

    def swap_loop(flags):
        x = list(range(len(flags)))
        i, j = 0, len(x) - 1
        while i <= j:
            if flags[x[i]]:
                i += 1
            else:
                x[i], x[j] = x[j], x[i]
                j -= 1
        return x

    def closed_form(flags):
        n, G = len(flags), sum(flags)
        left_bad = [p for p in range(G) if not flags[p]]
        right_good_desc = [p for p in range(n - 1, G - 1, -1) if flags[p]]
        M = len(left_bad)
        out = [None] * n
        for p in range(n):
            good, left = flags[p], p < G
            if good and left:
                dst = p
            elif not good and (left or p == G):
                m = left_bad.index(p) + 1 if left else M + 1
                dst = n - 1 if m == 1 else right_good_desc[m - 2] - 1
            elif not good:
                dst = p - 1
            else:
                dst = left_bad[right_good_desc.index(p)]
            assert out[dst] is None
            out[dst] = p
        return out

    rnd = random.Random(7)
    for _ in range(20000):
        n, pr = rnd.randint(1, 18), rnd.random()
        flags = [rnd.random() < pr for _ in range(n)]
        assert swap_loop(flags) == closed_form(flags), flags


@pytest.mark.parametrize("name", ["golden_kd", "golden_grid", "golden_tlas_kd", "golden_tlas_grid"])
def test_rtscene_roundtrip_of_kdtree_and_grid_chunks(name, flat_scenes, tmp_path):
    """.rtscene save -> load keeps every chunk of the KD-tree / grid scene kinds byte for byte, and the rt_scene_desc built
    from the reloaded file points at equal arrays (host data format, no GPU)"""
    import cpu_ray_tracer_b200 as rtb
    fs = flat_scenes(name)
    p = tmp_path / (name + ".rtscene.gz")
    fs.save(str(p))
    back = rtb.FlatScene.load(str(p))
    assert back.kind == fs.kind
    for chunk in ("header", "blas_table", "tris", "tlas_nodes", "obj_material", "materials", "tex_table", "tex_pixels",
                  "kd_nodes", "kd_tri_indices", "grid_header", "grid_cell_start", "grid_tri_indices", "blas_kd_table", "blas_grid_table"):
        a, b = getattr(fs, chunk), getattr(back, chunk)
        assert (a is None) == (b is None), chunk
        if a is not None:
            assert raw_equal(np.ascontiguousarray(a), np.ascontiguousarray(b)), chunk
    d = back.desc()
    assert d.kind == fs.kind and d.blas_count == len(fs.blas_table)
    if fs.kind in (abi.RT_SCENE_TLAS_KDTREE, abi.RT_SCENE_TLAS_GRID):
        assert bool(d.blas_accel) and d.tlas_node_count == len(fs.tlas_nodes)
    c = back.copy()
    assert c.kind == fs.kind and raw_equal(c.tris, fs.tris)


@pytest.mark.parametrize("accel,kind", [("kdtree", abi.RT_SCENE_TLAS_KDTREE), ("grid", abi.RT_SCENE_TLAS_GRID)])
def test_with_accelerator_on_tlas_scenes(accel, kind, flat_scenes):
    """a TLAS-over-BVH scene rebuilt over per-object KD-trees / grids equals what the reference built for that
    configuration (same XML, tests/golden), and instances that share a mesh share one tree / grid"""
    from oracle import porthost
    ref = flat_scenes("golden_tlas_kd" if accel == "kdtree" else "golden_tlas_grid")
    got = host_build.with_accelerator(flat_scenes("golden_tlas"), accel)
    assert got.kind == kind and raw_equal(got.tlas_nodes["left_right"], ref.tlas_nodes["left_right"])
    for f in ("aabb_min", "aabb_max"):
        assert biteq(got.tlas_nodes[f], ref.tlas_nodes[f])
    if accel == "kdtree":
        assert raw_equal(got.kd_nodes, ref.kd_nodes) and raw_equal(got.kd_tri_indices, ref.kd_tri_indices)
        assert raw_equal(got.blas_kd_table, ref.blas_kd_table)
    else:
        assert raw_equal(got.grid_tri_indices, ref.grid_tri_indices) and raw_equal(got.blas_grid_table, ref.blas_grid_table)
    tris = host_build.terrain_mesh(800, seed=5, size=1.0, height=0.5)
    inst = host_build.with_accelerator(host_build.instanced_grid(tris, 27), accel)
    table = inst.blas_kd_table if accel == "kdtree" else inst.blas_grid_table
    assert len(table) == 27 and len(set(table.tobytes()[i * table.itemsize:(i + 1) * table.itemsize] for i in range(27))) == 1
    po = porthost.PortOracle(inst)
    W, H = 64, 40
    hits, st = po.find_nearest(po.primary_rays(po.camera_default(W, H), W, H))
    assert (hits["obj_idx"] >= 2).any() and st["blas_entries"] > 0


# ---- ABI v5: 32-bit TLAS nodes, Refit restatement ---------------------------------------------------------------------
def test_tlas32_builder_is_the_same_clustering():
    """rtb_build_tlas32 (no 32 767-instance cap) builds the tree of rtb_build_tlas, and the oracle walks both alike"""
    from cpu_ray_tracer_b200 import host_build
    from oracle import porthost
    mesh = host_build.terrain_mesh(300, seed=4, size=0.8, height=0.5)
    a = host_build.instanced_grid(mesh, 150, tlas="host")
    b = host_build.instanced_grid(mesh, 150, tlas="host32")
    inner = a.tlas_nodes["left_right"] != 0
    assert np.array_equal(a.tlas_nodes["left_right"][inner] & 0xffff, b.tlas_nodes32["left"][inner])
    assert np.array_equal(a.tlas_nodes["left_right"][inner] >> 16, b.tlas_nodes32["right"][inner])
    assert np.array_equal(a.tlas_nodes["blas"][~inner], b.tlas_nodes32["right"][~inner])
    assert np.array_equal(a.tlas_nodes["aabb_min"].view(np.uint32), b.tlas_nodes32["aabb_min"].view(np.uint32))
    pa, pb = porthost.PortOracle(a), porthost.PortOracle(b)
    rays = pa.primary_rays(pa.camera_default(160, 96), 160, 96)
    ha, sa = pa.find_nearest(rays)
    hb, sb = pb.find_nearest(rays)
    assert ha.tobytes() == hb.tobytes() and sa == sb and sa["blas_entries"] > 0


def test_refit_restatement_properties():
    """orc_refit_bvh (BVH::Refit, bvh.cpp:26-43): refitting unmoved triangles reproduces the builder's boxes; after a move every
    leaf box is the bounds of its triangles, every interior box the union of its children - except node 1, which the reference's
    loop skips (`if (i != 1)`), unless the repaired variant is asked for"""
    from cpu_ray_tracer_b200 import host_build
    from oracle import porthost
    tris = host_build.terrain_mesh(5000, seed=2)
    nodes, idx, _ = host_build.build_bvh(tris)
    same = porthost.refit_bvh(nodes, tris, idx)
    assert same.tobytes() == nodes.tobytes()
    moved = np.array(tris, copy=True)
    for v in ("v0", "v1", "v2"):
        moved[v] = (moved[v] * np.float32(1.25) + np.float32(0.1)).astype(np.float32)
    for all_nodes in (False, True):
        out = porthost.refit_bvh(nodes, moved, idx, all_nodes=all_nodes)
        assert np.array_equal(out["left_first"], nodes["left_first"]) and np.array_equal(out["tri_count"], nodes["tri_count"])
        for i in range(len(out)):
            if i == 1 and not all_nodes:
                assert out[1].tobytes() == nodes[1].tobytes()       # stale: the reference never refits it
                continue
            n = out[i]
            if n["tri_count"] > 0:
                t = moved[idx[n["left_first"]:n["left_first"] + n["tri_count"]]]
                v = np.concatenate([t["v0"], t["v1"], t["v2"]])
                assert np.array_equal(n["aabb_min"], v.min(0)) and np.array_equal(n["aabb_max"], v.max(0))
            else:
                l, r = out[n["left_first"]], out[n["left_first"] + 1]
                assert np.array_equal(n["aabb_min"], np.minimum(l["aabb_min"], r["aabb_min"]))
                assert np.array_equal(n["aabb_max"], np.maximum(l["aabb_max"], r["aabb_max"]))


def test_scene_file_offsets_are_checked_before_they_become_pointers(flat_scenes):
    """a .rtscene whose tables point outside its arrays must not reach the C-ABI as wild pointers (FlatScene.desc)"""
    for name, field, value in (("golden_file", "tri_count", 10 ** 8), ("golden_tlas", "node_offset", 10 ** 8), ("golden_tlas", "tri_offset", 2 ** 31)):
        fs = flat_scenes(name).copy()
        fs.blas_table[len(fs.blas_table) - 1][field] = value
        with pytest.raises(ValueError):
            fs.desc()
    fs = flat_scenes("golden_file").copy()
    fs.tex_table[0]["pixel_offset"] = len(fs.tex_pixels)
    with pytest.raises(ValueError):
        fs.desc()
    # device-built scenes carry no nodes / indices: only the triangle range is checked
    fs = flat_scenes("golden_file").copy()
    fs.device_build = True
    d = fs.desc()
    assert not d.blas[0].nodes and not d.blas[0].tri_indices and d.blas[0].tris
