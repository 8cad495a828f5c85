"""Host-side scene construction for scenes that do not come from the reference's loaders: the reference's
SAH BVH and agglomerative TLAS builders restated in C++ (host/bvh_build.cpp, bit-identical output), plus
generators for the synthetic BASELINE configurations:

  instanced_grid(...)   configs[3]: one BLAS instanced N times on a jittered 3-D grid under a TLAS
                        (rigid transforms only: FastInvertedTransformNoScale needs it), sharing ONE copy of
                        the mesh: blas_table entries point at the same node / triangle ranges
  terrain_mesh(...)     configs[4]: a displaced grid mesh of ~n triangles in one flat SAH BVH

Everything here is host logic (numpy + g++); the GPU only ever sees the resulting FlatScene.
"""
import ctypes as C
import os

import numpy as np

from . import abi
from .scene_file import BLAS_TABLE_DTYPE, HEADER_DTYPE, TEX_TABLE_DTYPE, FlatScene

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "librt_host.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise OSError(f"{LIB_PATH} not built: run `python __graft_entry__.py build`")
        L = C.CDLL(LIB_PATH)
        vp = C.c_void_p
        L.rtb_build_bvh.argtypes = [vp, C.c_uint32, vp, vp, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.rtb_world_bounds.argtypes = [vp, vp, vp, vp]
        L.rtb_invert_rigid.argtypes = [vp, vp]
        L.rtb_build_tlas.argtypes = [vp, C.c_uint32, vp, C.POINTER(C.c_uint32)]
        L.rtb_build_tlas32.argtypes = [vp, C.c_uint32, vp, C.POINTER(C.c_uint32)]
        u32p = C.POINTER(C.c_uint32)
        L.rtb_kd_build.argtypes, L.rtb_kd_build.restype = [vp, C.c_uint32], vp
        L.rtb_kd_sizes.argtypes, L.rtb_kd_sizes.restype = [vp, u32p, u32p, u32p], None
        L.rtb_kd_copy.argtypes, L.rtb_kd_copy.restype = [vp, vp, vp], None
        L.rtb_kd_free.argtypes, L.rtb_kd_free.restype = [vp], None
        L.rtb_grid_build.argtypes, L.rtb_grid_build.restype = [vp, C.c_uint32], vp
        L.rtb_grid_sizes.argtypes, L.rtb_grid_sizes.restype = [vp, vp, u32p, u32p], None
        L.rtb_grid_copy.argtypes, L.rtb_grid_copy.restype = [vp, vp, vp], None
        L.rtb_grid_free.argtypes, L.rtb_grid_free.restype = [vp], None
        _lib = L
    return _lib


def build_bvh(tris):
    """BVH::Build on a TRI_DTYPE array -> (nodes[:nodesUsed], tri_indices, max_depth)"""
    tris = np.ascontiguousarray(tris, abi.TRI_DTYPE)
    n = len(tris)
    nodes = np.zeros(2 * n - 1, abi.NODE_DTYPE)
    idx = np.zeros(n, np.uint32)
    used, depth = C.c_uint32(), C.c_uint32()
    rc = lib().rtb_build_bvh(tris.ctypes.data, n, nodes.ctypes.data, idx.ctypes.data, C.byref(used), C.byref(depth))
    if rc != 0:
        raise ValueError(f"rtb_build_bvh failed: {rc}")
    return nodes[:used.value].copy(), idx, depth.value


def build_kdtree(tris):
    """KDTree::Build (kdtree.cpp:4-112) on a TRI_DTYPE array -> (kd_nodes, kd_tri_indices, max_depth), flattened as
    include/rt_b200.h rt_kd_node documents"""
    tris = np.ascontiguousarray(tris, abi.TRI_DTYPE)
    h = lib().rtb_kd_build(tris.ctypes.data, len(tris))
    if not h:
        raise ValueError("rtb_kd_build failed")
    nn, ni, depth = C.c_uint32(), C.c_uint32(), C.c_uint32()
    lib().rtb_kd_sizes(h, C.byref(nn), C.byref(ni), C.byref(depth))
    nodes = np.zeros(nn.value, abi.KD_NODE_DTYPE)
    idx = np.zeros(ni.value, np.uint32)
    lib().rtb_kd_copy(h, nodes.ctypes.data, idx.ctypes.data)
    lib().rtb_kd_free(h)
    return nodes, idx, depth.value


def build_grid(tris):
    """Grid::Build (grid.cpp:4-60) on a TRI_DTYPE array -> (grid_header[1], cell_start, tri_indices)"""
    tris = np.ascontiguousarray(tris, abi.TRI_DTYPE)
    h = lib().rtb_grid_build(tris.ctypes.data, len(tris))
    if not h:
        raise ValueError("rtb_grid_build failed")
    hdr = np.zeros(1, abi.GRID_HEADER_DTYPE)
    nc, ni = C.c_uint32(), C.c_uint32()
    lib().rtb_grid_sizes(h, hdr.ctypes.data, C.byref(nc), C.byref(ni))
    start = np.zeros(nc.value + 1, np.uint32)
    idx = np.zeros(ni.value, np.uint32)
    lib().rtb_grid_copy(h, start.ctypes.data, idx.ctypes.data)
    lib().rtb_grid_free(h)
    return hdr, start, idx


def _with_blas_accelerator(flat, accel):
    """TLASFileScene with per-object KD-trees / grids (tlas_file_scene.h:12-14) from one with per-object BVHs: same
    triangles, transforms and TLAS (the TLAS is the same agglomerative BVH for all three BLAS kinds; its leaf boxes
    are the BLAS' local bounds, which every builder derives from the same vertices).  BLAS that share a triangle
    range (true instancing) share one tree / grid: their table rows point at the same ranges."""
    chunks = {k: getattr(flat, k) for k in ("tris", "tlas_nodes", "obj_material", "materials", "tex_table", "tex_pixels")}
    chunks["header"] = flat.header.copy()
    chunks["blas_table"] = flat.blas_table.copy()
    chunks["blas_table"]["node_offset"], chunks["blas_table"]["node_count"] = 0, 0
    chunks["nodes"], chunks["tri_indices"] = np.zeros(0, abi.NODE_DTYPE), np.zeros(0, np.uint32)
    built = {}
    if accel == "kdtree":
        chunks["header"]["kind"] = abi.RT_SCENE_TLAS_KDTREE
        table, nodes, idx, no, io = np.zeros(len(flat.blas_table), abi.BLAS_KD_TABLE_DTYPE), [], [], 0, 0
        for i, b in enumerate(flat.blas_table):
            key = (int(b["tri_offset"]), int(b["tri_count"]))
            if key not in built:
                n, ix, _ = build_kdtree(flat.tris[key[0]:key[0] + key[1]])
                built[key] = (no, len(n), io, len(ix))
                nodes.append(n), idx.append(ix)
                no, io = no + len(n), io + len(ix)
            table[i] = built[key]
        chunks["blas_kd_table"], chunks["kd_nodes"], chunks["kd_tri_indices"] = table, np.concatenate(nodes), np.concatenate(idx)
    elif accel == "grid":
        chunks["header"]["kind"] = abi.RT_SCENE_TLAS_GRID
        table, starts, idx, co, io = np.zeros(len(flat.blas_table), abi.BLAS_GRID_TABLE_DTYPE), [], [], 0, 0
        for i, b in enumerate(flat.blas_table):
            key = (int(b["tri_offset"]), int(b["tri_count"]))
            if key not in built:
                hdr, st, ix = build_grid(flat.tris[key[0]:key[0] + key[1]])
                row = np.zeros(1, abi.BLAS_GRID_TABLE_DTYPE)
                for f in ("resolution", "cell_size", "bounds_min", "bounds_max"):
                    row[f] = hdr[f]
                row["cell_offset"], row["cell_count"], row["idx_offset"], row["idx_count"] = co, len(st) - 1, io, len(ix)
                built[key] = row[0]
                starts.append(st), idx.append(ix)
                co, io = co + len(st), io + len(ix)
            table[i] = built[key]
        chunks["blas_grid_table"], chunks["grid_cell_start"], chunks["grid_tri_indices"] = table, np.concatenate(starts), np.concatenate(idx)
    else:
        raise ValueError(accel)
    return FlatScene(chunks)


def with_accelerator(flat, accel):
    """The same FileScene with another of its accelerators (file_scene.h:10-12): accel = "kdtree" | "grid"; for a
    TLASFileScene, the same scene with per-object KD-trees / grids (tlas_file_scene.h:12-14).
    Triangles, materials and textures are shared with `flat`; the BVH arrays are dropped."""
    if flat.kind == abi.RT_SCENE_TLAS:
        return _with_blas_accelerator(flat, accel)
    chunks = {k: getattr(flat, k) for k in ("header", "blas_table", "tris", "obj_material", "materials", "tex_table", "tex_pixels")}
    chunks["header"] = flat.header.copy()
    chunks["blas_table"] = flat.blas_table[:1].copy()
    chunks["blas_table"]["node_offset"], chunks["blas_table"]["node_count"] = 0, 0
    chunks["nodes"], chunks["tri_indices"] = np.zeros(0, abi.NODE_DTYPE), np.zeros(0, np.uint32)
    chunks["tlas_nodes"] = np.zeros(0, abi.TLAS_NODE_DTYPE)
    if accel == "kdtree":
        chunks["header"]["kind"] = abi.RT_SCENE_FLAT_KDTREE
        chunks["kd_nodes"], chunks["kd_tri_indices"], _ = build_kdtree(flat.tris)
    elif accel == "grid":
        chunks["header"]["kind"] = abi.RT_SCENE_FLAT_GRID
        chunks["grid_header"], chunks["grid_cell_start"], chunks["grid_tri_indices"] = build_grid(flat.tris)
    else:
        raise ValueError(accel)
    return FlatScene(chunks)


def invert_rigid(T):
    T = np.ascontiguousarray(T, np.float32).reshape(16)
    out = np.zeros(16, np.float32)
    lib().rtb_invert_rigid(T.ctypes.data, out.ctypes.data)
    return out


def world_bounds(root_min, root_max, T):
    out = np.zeros(6, np.float32)
    a, b, t = (np.ascontiguousarray(x, np.float32) for x in (root_min, root_max, T))
    lib().rtb_world_bounds(a.ctypes.data, b.ctypes.data, t.ctypes.data, out.ctypes.data)
    return out


def build_tlas(bounds):
    """TLASBVH::Build on an (n, 6) array of world bounds -> tlas_nodes[:nodesUsed]"""
    bounds = np.ascontiguousarray(bounds, np.float32).reshape(-1, 6)
    n = len(bounds)
    out = np.zeros(2 * n, abi.TLAS_NODE_DTYPE)
    used = C.c_uint32()
    rc = lib().rtb_build_tlas(bounds.ctypes.data, n, out.ctypes.data, C.byref(used))
    if rc != 0:
        raise ValueError(f"rtb_build_tlas failed: {rc} (more than 32767 instances do not fit the reference's 2 x 16-bit child indices)")
    return out[:used.value].copy()


def build_tlas32(bounds):
    """the same clustering with 32-bit children (rt_tlas_node32): no 32 767-instance cap"""
    bounds = np.ascontiguousarray(bounds, np.float32).reshape(-1, 6)
    n = len(bounds)
    out = np.zeros(2 * n, abi.TLAS_NODE32_DTYPE)
    used = C.c_uint32()
    rc = lib().rtb_build_tlas32(bounds.ctypes.data, n, out.ctypes.data, C.byref(used))
    if rc != 0:
        raise ValueError(f"rtb_build_tlas32 failed: {rc}")
    return out[:used.value].copy()


def make_tris(v0, v1, v2, obj_idx=2, normals=None, uvs=None):
    """Tri records as the reference's loaders fill them (model.cpp:60-79): centroid = (v0 + v1 + v2) * 0.3333f"""
    n = len(v0)
    t = np.zeros(n, abi.TRI_DTYPE)
    t["v0"], t["v1"], t["v2"] = v0, v1, v2
    if normals is None:
        e1, e2 = (v1 - v0).astype(np.float32), (v2 - v0).astype(np.float32)
        nn = np.cross(e1, e2).astype(np.float32)
        ln = np.sqrt((nn * nn).sum(1, keepdims=True)).astype(np.float32)
        nn = (nn / np.where(ln > 0, ln, 1)).astype(np.float32)
        normals = (nn, nn, nn)
    t["n0"], t["n1"], t["n2"] = normals
    if uvs is not None:
        t["uv0"], t["uv1"], t["uv2"] = uvs
    t["centroid"] = ((t["v0"] + t["v1"]).astype(np.float32) + t["v2"]).astype(np.float32) * np.float32(0.3333)
    t["obj_idx"] = obj_idx
    return t


def _identity():
    return np.eye(4, dtype=np.float32).reshape(16)


def _header(kind, like=None, sky=-1, floor=-1):
    h = np.zeros(1, HEADER_DTYPE)
    if like is not None:
        h[0] = like.header[0]
    else:
        # FileScene defaults (file_scene.cpp:15-19): floor plane y = -1, light quad of size 1 at light_position
        lp = np.array([0.0, 3.0, 1.0], np.float32)
        T = _identity()
        T[3], T[7], T[11] = lp
        h["floor_n"], h["floor_d"], h["floor_invto"] = (0, 1, 0), 1.0, 1.0
        h["light_T"], h["light_inv_T"], h["light_size"] = T, invert_rigid(T), 0.5
        h["light_color"] = (24, 24, 22)
        h["light_pos"] = lp - np.array([0, 0.01, 0], np.float32)
    h["kind"], h["skydome_texture"], h["floor_texture"] = kind, sky, floor
    return h


def _gradient_sky(w=512, h=256):
    y = np.linspace(0, 1, h, dtype=np.float32)[:, None]
    x = np.linspace(0, 1, w, dtype=np.float32)[None, :]
    r = (255 * (0.35 + 0.4 * y + 0 * x)).astype(np.uint32)
    g = (255 * (0.55 + 0.3 * y + 0 * x)).astype(np.uint32)
    b = (255 * (0.95 - 0.25 * y + 0.05 * np.sin(6.28 * x))).astype(np.uint32)
    return ((r << 16) | (g << 8) | b).astype(np.uint32).reshape(-1), w, h


def _materials(kinds):
    m = np.zeros(len(kinds), abi.MATERIAL_DTYPE)
    for i, (refl, refr, albedo) in enumerate(kinds):
        m[i]["reflectivity"], m[i]["refractivity"], m[i]["albedo"], m[i]["texture"] = refl, refr, albedo, -1
        m[i]["absorption"] = (0.5, 0.1, 0.5) if refr > 0 else (0, 0, 0)
    return m


def flat_scene_from_tris(tris, materials=None, builder=None):
    """FileScene + USE_BVH equivalent: one SAH BVH over all triangles, generated sky, untextured floor.
    builder: callable(tris) -> (nodes, tri_indices, _); default = the host restatement, api.build_bvh_gpu = the
    GPU builder (same arrays bit for bit)"""
    nodes, idx, _ = (builder or build_bvh)(tris)
    sky, w, h = _gradient_sky()
    bt = np.zeros(1, BLAS_TABLE_DTYPE)
    bt[0]["node_count"], bt[0]["tri_count"] = len(nodes), len(tris)
    bt[0]["T"], bt[0]["inv_T"], bt[0]["obj_idx"], bt[0]["mat_idx"] = _identity(), _identity(), -1, -1
    objs = int(tris["obj_idx"].max()) - 1
    mats = _materials([(0.0, 0.0, (0.8, 0.8, 0.8))] * objs) if materials is None else materials
    tt = np.zeros(1, TEX_TABLE_DTYPE)
    tt[0]["width"], tt[0]["height"] = w, h
    return FlatScene({"header": _header(abi.RT_SCENE_FLAT, sky=0), "blas_table": bt, "nodes": nodes, "tris": tris, "tri_indices": idx,
                      "tlas_nodes": np.zeros(0, abi.TLAS_NODE_DTYPE), "obj_material": np.arange(objs, dtype=np.int32) % len(mats),
                      "materials": mats, "tex_table": tt, "tex_pixels": sky})


def terrain_mesh(n_tris, seed=1, size=None, height=0.8):
    """displaced grid: 2 * g * g triangles with g = ceil(sqrt(n_tris / 2)); deterministic fractal-ish height field.
    The grid cell is kept >= 0.03 units: the reference's Moeller-Trumbore rejects |det| < 1e-4 (bvh.cpp:207), a test
    that is not scale-invariant — triangles with edges much below 0.01 can never be hit by any ray."""
    g = int(np.ceil(np.sqrt(n_tris / 2)))
    if size is None:
        size = max(6.0, 0.03 * g)
    rng = np.random.default_rng(seed)
    xs = np.linspace(-size / 2, size / 2, g + 1, dtype=np.float32)
    X, Z = np.meshgrid(xs, xs + np.float32(2.0))
    Y = np.zeros_like(X)
    for octave in range(6):
        f = np.float32(2.0 ** octave)
        ph = rng.uniform(0, 6.28, 4).astype(np.float32)
        Y += (np.sin(f * X * 1.3 + ph[0]) * np.cos(f * Z * 1.1 + ph[1]) + 0.5 * np.sin(f * (X + Z) * 0.7 + ph[2])).astype(np.float32) / f
    Y = (Y * np.float32(height / 2) - np.float32(0.6)).astype(np.float32)
    P = np.stack([X, Y, Z], -1).astype(np.float32)
    a, b, c, d = P[:-1, :-1], P[:-1, 1:], P[1:, :-1], P[1:, 1:]
    v0 = np.concatenate([a.reshape(-1, 3), b.reshape(-1, 3)])
    v1 = np.concatenate([c.reshape(-1, 3), c.reshape(-1, 3)])
    v2 = np.concatenate([b.reshape(-1, 3), d.reshape(-1, 3)])
    return make_tris(v0, v1, v2, obj_idx=2)


def instanced_grid(mesh_tris, n_instances, seed=0x12345678, spacing=None, like=None, tlas="host"):
    """TLASFileScene equivalent with TRUE instancing: every instance references the same BLAS arrays.
    Rigid transforms: rotation about Y by a xorshift32 angle, translation on a jittered 3-D grid."""
    tris = np.ascontiguousarray(mesh_tris, abi.TRI_DTYPE).copy()
    nodes, idx, _ = build_bvh(tris)
    ext = nodes[0]["aabb_max"] - nodes[0]["aabb_min"]
    spacing = float(ext.max()) * 1.6 if spacing is None else spacing
    side = int(np.ceil(n_instances ** (1 / 3)))
    bt = np.zeros(n_instances, BLAS_TABLE_DTYPE)
    bounds = np.zeros((n_instances, 6), np.float32)
    s = np.uint32(seed)

    def rnd():
        nonlocal s
        s ^= np.uint32(s << np.uint32(13)); s ^= np.uint32(s >> np.uint32(17)); s ^= np.uint32(s << np.uint32(5))
        return np.float32(s) * np.float32(2.3283064365387e-10)
    with np.errstate(over="ignore"):
        for i in range(n_instances):
            gx, gy, gz = i % side, (i // side) % side, i // (side * side)
            ang = rnd() * np.float32(6.2831853)
            c, sn = np.float32(np.cos(ang)), np.float32(np.sin(ang))
            pos = np.array([(gx - side / 2 + rnd() * 0.3) * spacing, (gy + rnd() * 0.3) * spacing - 0.5,
                            (gz + rnd() * 0.3) * spacing + 2.0], np.float32)
            T = _identity()
            T[0], T[2], T[8], T[10] = c, sn, -sn, c   # mat4::RotateY (tmplmath.h:674) then Translate
            T[3], T[7], T[11] = pos
            bt[i]["node_count"], bt[i]["tri_count"] = len(nodes), len(tris)  # node_offset = tri_offset = 0: shared
            bt[i]["T"], bt[i]["inv_T"] = T, invert_rigid(T)
            bt[i]["obj_idx"], bt[i]["mat_idx"] = i + 2, i % 3
            bounds[i] = world_bounds(nodes[0]["aabb_min"], nodes[0]["aabb_max"], T)
    sky, w, h = _gradient_sky()
    tt = np.zeros(1, TEX_TABLE_DTYPE)
    tt[0]["width"], tt[0]["height"] = w, h
    mats = _materials([(0.0, 0.0, (0.8, 0.6, 0.4)), (0.9, 0.0, (0.9, 0.9, 0.9)), (0.1, 0.8, (0.9, 0.95, 1.0))])
    # tlas: "host" = the reference's node format (<= 32 767 instances), "host32" = 32-bit children, "none" = left to the device
    chunks = {"header": _header(abi.RT_SCENE_TLAS, like=like, sky=0), "blas_table": bt, "nodes": nodes, "tris": tris,
              "tri_indices": idx, "obj_material": (np.arange(n_instances) % 3).astype(np.int32),
              "materials": mats, "tex_table": tt, "tex_pixels": sky}
    if tlas == "host":
        chunks["tlas_nodes"] = build_tlas(bounds)
    elif tlas == "host32":
        chunks["tlas_nodes32"] = build_tlas32(bounds)
    fs = FlatScene(chunks)
    fs.instance_bounds = bounds
    if tlas == "none":
        fs.device_tlas = True
    return fs
