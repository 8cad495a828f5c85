import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
BAKED = os.path.join(ROOT, "oracle", "_ref", "scenes")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def biteq(a, b):
    """bitwise equality of float arrays (NaN == NaN when the payloads match)"""
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    return a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))


def baked_scenes():
    """scenes flattened by the reference's loaders (oracle/bake_scenes.py); absent on a fresh checkout"""
    return sorted(os.path.basename(p)[:-len(".rtscene.gz")] for p in glob.glob(os.path.join(BAKED, "*.rtscene.gz")))


def scene_path(name):
    if name.startswith("golden_"):
        return os.path.join(GOLDEN, name + ".rtscene.gz")
    return os.path.join(BAKED, name + ".rtscene.gz")


def all_scene_names():
    return ["golden_file", "golden_tlas", "golden_kd", "golden_grid", "golden_tlas_kd", "golden_tlas_grid"] + baked_scenes()


@pytest.fixture(scope="session")
def flat_scenes():
    import cpu_ray_tracer_b200 as rtb
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = rtb.FlatScene.load(scene_path(name))
        return cache[name]
    return get


@pytest.fixture(scope="session")
def oracles(flat_scenes):
    from oracle import porthost
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = porthost.PortOracle(flat_scenes(name))
        return cache[name]
    return get


@pytest.fixture(scope="session")
def gpu_scenes(flat_scenes):
    from cpu_ray_tracer_b200 import api
    cache = {}

    def get(name, counters=True):
        key = (name, counters)
        if key not in cache:
            cache[key] = api.open_scene(flat_scenes(name), counters=counters)
            if flat_scenes(name).kind in (0, 1):
                cache[key].validate()   # every reference in the device arrays in range, trees, stack depth (rt_scene_validate)
        return cache[key]
    yield get
    for s in cache.values():
        s.close()


def shadow_rays_from(flat, rays, hits):
    """shadow rays toward the scene light from every hit point, as DirectIllumination builds them
    (2. WhittedStyle/renderer.cpp:110-118); numpy fp32 arithmetic, only used as test input"""
    from cpu_ray_tracer_b200 import api
    m = hits["obj_idx"] >= 0
    I = rays["O"][m] + hits["t"][m, None] * rays["D"][m]
    L = flat.header["light_pos"][0][None, :] - I
    dist = np.sqrt((L * L).sum(1)).astype(np.float32)
    ok = dist > 1e-3
    L = (L[ok] / dist[ok, None]).astype(np.float32)
    return api.make_rays(I[ok] + L * np.float32(0.001), L, dist[ok] - np.float32(0.002))


def random_rays(flat, n, seed):
    """incoherent rays: origins in a box around the scene, uniform directions; some axis-aligned
    (zero direction components exercise the NaN-exact slab path) and some starting on box planes"""
    from cpu_ray_tracer_b200 import api
    rng = np.random.default_rng(seed)
    O = rng.uniform(-4, 4, (n, 3)).astype(np.float32)
    O[:, 1] = rng.uniform(-0.9, 4, n)
    D = rng.normal(size=(n, 3)).astype(np.float32)
    D /= np.linalg.norm(D, axis=1, keepdims=True).astype(np.float32)
    k = n // 16
    for axis in range(3):
        D[axis * k:(axis + 1) * k, axis] = 0.0
    D[3 * k:4 * k] = np.eye(3, dtype=np.float32)[rng.integers(0, 3, k)] * rng.choice([-1.0, 1.0], (k, 1)).astype(np.float32)
    # origins exactly on node-box planes: 0 * inf = NaN in the slab test
    if flat.kind in (2, 4): # KD-tree(s): node boxes (a split plane is the max / min of the two children); kind 4: object space
        planes = flat.kd_nodes["aabb_min"]
    elif flat.kind == 5:    # per-object grids: TLAS node boxes
        planes = flat.tlas_nodes["aabb_min"]
    elif flat.kind == 3:    # grid: cell corners
        g = flat.grid_header[0]
        ijk = np.stack([rng.integers(0, int(r) + 1, 4 * k) for r in g["resolution"]], 1)
        planes = (g["bounds_min"][None, :] + ijk.astype(np.float32) * g["cell_size"][None, :]).astype(np.float32)
    else:
        planes = flat.nodes["aabb_min"]
    pick = rng.integers(0, len(planes), k)
    O[3 * k:4 * k] = planes[pick]
    nrm = np.linalg.norm(D, axis=1, keepdims=True)
    D = (D / np.where(nrm > 0, nrm, 1)).astype(np.float32)
    return api.make_rays(O, D)


def displaced(tris, amp, seed):
    """the same mesh with every vertex moved by a smooth deterministic field (topology kept): what an animated model hands to
    Refit.  numpy fp32, only used as test INPUT (the golden refit vectors store the moved vertices themselves)."""
    rng = np.random.default_rng(seed)
    t = np.array(tris, copy=True)
    for v in ("v0", "v1", "v2"):
        t[v] = (t[v] + (amp * np.sin(7.0 * t[v][:, ::-1] + rng.uniform(0, 6.28, 3))).astype(np.float32)).astype(np.float32)
    t["centroid"] = ((t["v0"] + t["v1"] + t["v2"]) * np.float32(0.3333)).astype(np.float32)
    return t


def verts9(tris):
    return np.concatenate([tris["v0"], tris["v1"], tris["v2"]], 1).astype(np.float32)
