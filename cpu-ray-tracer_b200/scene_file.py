""".rtscene: the flattened scene the C-ABI consumes, as a chunk file (host-side data format).

A flattened scene is exactly what the reference's FileScene / TLASFileScene hold after their
constructors ran (file_scene.cpp:4-62, tlas_file_scene.cpp:4-93), in the reference's own POD layouts:

    header       scalars: kind, skydome/floor texture ids, floor plane, light quad, light colour/pos
    blas_table   per BLAS: node/tri ranges into the shared arrays, T, invT, objIdx, matIdx
    nodes        BVHNode[]   (32 B, blas_bvh.h:13-20)       all BLAS concatenated
    tris         Tri[]       (112 B, helper.h:6-26)
    tri_indices  uint[]      (per-BLAS local indices)
    tlas_nodes   TLASBVHNode[] (32 B, tlas_bvh.h:7-14)      TLAS scenes only
    tlas_nodes32 rt_tlas_node32[] (32 B)                    TLAS scenes with more than 32 767 instances: 32-bit children
    obj_material int[]       material index of object (objIdx - 2)
    materials    rt_material[]
    tex_table    per texture: offset into tex_pixels, width, height
    tex_pixels   uint[]      packed 0x00RRGGBB (texture.h:31-34)
    kd_nodes     rt_kd_node[] (48 B) + kd_tri_indices uint[]      kind 2: FileScene with its KD-tree (kdtree.cpp)
    grid_header  resolution / cellSize / localBounds + grid_cell_start uint[cells+1] + grid_tri_indices uint[]
                                                                  kind 3: FileScene with its uniform grid (grid.cpp)
    blas_kd_table / blas_grid_table   per BLAS: ranges into kd_nodes + kd_tri_indices, or grid header + ranges into
                 grid_cell_start (cells + 1 entries per BLAS) + grid_tri_indices       kinds 4 / 5: TLASFileScene over
                 per-object KD-trees / grids (tlas_kdtree.cpp, tlas_grid.cpp); indices are relative to the BLAS' own range

File layout: "RTSCN001", u32 chunk count, u32 pad, then per chunk: char name[24], u64 nbytes, payload
padded to 8 bytes.  A ".gz" suffix means the whole file is gzip-compressed.
Writers: oracle/ref_build/ref_api.cpp (from the reference's own loaders) and FlatScene.save (synthetic
scenes built by host_build.py).
"""
import ctypes as C
import gzip
import struct

import numpy as np

from . import abi

HEADER_DTYPE = np.dtype([("kind", "<i4"), ("skydome_texture", "<i4"), ("floor_texture", "<i4"), ("reserved", "<i4"),
                         ("floor_n", "<f4", 3), ("floor_d", "<f4"), ("floor_invto", "<f4"),
                         ("light_T", "<f4", 16), ("light_inv_T", "<f4", 16), ("light_size", "<f4"),
                         ("light_color", "<f4", 3), ("light_pos", "<f4", 3)])
BLAS_TABLE_DTYPE = np.dtype([("node_offset", "<u4"), ("node_count", "<u4"), ("tri_offset", "<u4"), ("tri_count", "<u4"),
                             ("T", "<f4", 16), ("inv_T", "<f4", 16), ("obj_idx", "<i4"), ("mat_idx", "<i4")])
TEX_TABLE_DTYPE = np.dtype([("pixel_offset", "<u8"), ("width", "<i4"), ("height", "<i4")])

_CHUNK_DTYPES = {
    "header": HEADER_DTYPE, "blas_table": BLAS_TABLE_DTYPE, "nodes": abi.NODE_DTYPE, "tris": abi.TRI_DTYPE,
    "tri_indices": np.dtype("<u4"), "tlas_nodes": abi.TLAS_NODE_DTYPE, "obj_material": np.dtype("<i4"),
    "materials": abi.MATERIAL_DTYPE, "tex_table": TEX_TABLE_DTYPE, "tex_pixels": np.dtype("<u4"),
    "kd_nodes": abi.KD_NODE_DTYPE, "kd_tri_indices": np.dtype("<u4"),
    "grid_header": abi.GRID_HEADER_DTYPE, "grid_cell_start": np.dtype("<u4"), "grid_tri_indices": np.dtype("<u4"),
    "blas_kd_table": abi.BLAS_KD_TABLE_DTYPE, "blas_grid_table": abi.BLAS_GRID_TABLE_DTYPE,
    "tlas_nodes32": abi.TLAS_NODE32_DTYPE,
}
_OPTIONAL = ("tlas_nodes32", "kd_nodes", "kd_tri_indices", "grid_header", "grid_cell_start", "grid_tri_indices", "blas_kd_table", "blas_grid_table")


class FlatScene:
    """Host-side container of the flattened arrays; builds the rt_scene_desc handed over the C-ABI."""

    def __init__(self, chunks):
        self.header = chunks["header"]
        self.blas_table = chunks["blas_table"]
        self.nodes = chunks["nodes"]
        self.tris = chunks["tris"]
        self.tri_indices = chunks["tri_indices"]
        self.tlas_nodes = chunks.get("tlas_nodes", np.zeros(0, abi.TLAS_NODE_DTYPE))
        self.obj_material = chunks["obj_material"]
        self.materials = chunks["materials"]
        self.tex_table = chunks["tex_table"]
        self.tex_pixels = chunks["tex_pixels"]
        for name in _OPTIONAL:
            setattr(self, name, chunks.get(name))
        self._keep = None
        # ABI v5 construction on the device (not stored in files): desc() then hands over triangles only
        self.device_build = False   # blas.nodes = blas.tri_indices = NULL: SAH build + layout on the GPU
        self.device_tlas = False    # tlas_nodes = tlas_nodes32 = NULL: TLASBVH::Build on the GPU

    def copy(self):
        """deep copy of the arrays (not of the ctypes tables a previous desc() call left behind)"""
        names = ["header", "blas_table", "nodes", "tris", "tri_indices", "tlas_nodes", "obj_material", "materials",
                 "tex_table", "tex_pixels"] + [n for n in _OPTIONAL if getattr(self, n) is not None]
        return FlatScene({n: np.array(getattr(self, n), copy=True) for n in names})

    @property
    def kind(self):
        return int(self.header["kind"][0])

    @property
    def triangle_count(self):
        return int(self.tris.shape[0])

    @staticmethod
    def load(path):
        opener = gzip.open if str(path).endswith(".gz") else open
        with opener(path, "rb") as f:
            data = f.read()
        if data[:8] != b"RTSCN001":
            raise ValueError(f"{path}: not an .rtscene file")
        (count,) = struct.unpack_from("<I", data, 8)
        off = 16
        chunks = {}
        for _ in range(count):
            name = data[off:off + 24].split(b"\0", 1)[0].decode()
            (nbytes,) = struct.unpack_from("<Q", data, off + 24)
            off += 32
            dt = _CHUNK_DTYPES[name]
            chunks[name] = np.frombuffer(data, dtype=dt, count=nbytes // dt.itemsize, offset=off).copy()
            off += (nbytes + 7) & ~7
        fs = FlatScene(chunks)
        if fs.kind == abi.RT_SCENE_TLAS_KDTREE and fs.blas_kd_table is None:
            fs.rebuild_blas_kdtrees()
        return fs

    def rebuild_blas_kdtrees(self):
        """A kind-4 file may omit the per-object KD-trees (the reference's median-split trees reach 20 MB for a 4 k-triangle
        scene): they are rebuilt here with the host restatement of BLASKDTree::Build (host/bvh_build.cpp, byte-identical to
        the reference's trees on every scene it flattened, tests/test_host_build.py)."""
        from . import host_build
        nodes, idx, table = [], [], np.zeros(len(self.blas_table), abi.BLAS_KD_TABLE_DTYPE)
        no, io = 0, 0
        for i, b in enumerate(self.blas_table):
            t0, tn = int(b["tri_offset"]), int(b["tri_count"])
            n, ix, _ = host_build.build_kdtree(self.tris[t0:t0 + tn])
            table[i] = (no, len(n), io, len(ix))
            nodes.append(n), idx.append(ix)
            no, io = no + len(n), io + len(ix)
        self.kd_nodes, self.kd_tri_indices, self.blas_kd_table = np.concatenate(nodes), np.concatenate(idx), table

    def save_without_kdtrees(self, path):
        keep = self.kd_nodes, self.kd_tri_indices, self.blas_kd_table
        self.kd_nodes = self.kd_tri_indices = self.blas_kd_table = None
        try:
            self.save(path)
        finally:
            self.kd_nodes, self.kd_tri_indices, self.blas_kd_table = keep

    def save(self, path):
        opener = gzip.open if str(path).endswith(".gz") else open
        names = ["header", "blas_table", "nodes", "tris", "tri_indices", "tlas_nodes", "obj_material",
                 "materials", "tex_table", "tex_pixels"] + [n for n in _OPTIONAL if getattr(self, n) is not None]
        with opener(path, "wb") as f:
            f.write(b"RTSCN001" + struct.pack("<II", len(names), 0))
            for name in names:
                raw = np.ascontiguousarray(getattr(self, name)).tobytes()
                f.write(name.encode().ljust(24, b"\0") + struct.pack("<Q", len(raw)) + raw)
                f.write(b"\0" * (((len(raw) + 7) & ~7) - len(raw)))

    def desc(self):
        """rt_scene_desc whose pointers reference this object's numpy arrays (kept alive by self)."""
        h = self.header[0]
        nb = len(self.blas_table)
        blas = (abi.rt_blas_desc * nb)()
        bvh_kind = int(h["kind"]) in (abi.RT_SCENE_FLAT, abi.RT_SCENE_TLAS)
        on_device = self.device_build and bvh_kind
        for i, b in enumerate(self.blas_table):
            no, nc, to, tc = int(b["node_offset"]), int(b["node_count"]), int(b["tri_offset"]), int(b["tri_count"])
            # offsets come from a file: never turn them into pointers unchecked
            if to + tc > len(self.tris) or (bvh_kind and not on_device and (no + nc > len(self.nodes) or to + tc > len(self.tri_indices))):
                raise ValueError(f"BLAS {i}: node / triangle range outside the scene's arrays")
            if not on_device:
                blas[i].nodes = self.nodes.ctypes.data + no * 32
                blas[i].node_count = nc
                blas[i].tri_indices = self.tri_indices.ctypes.data + to * 4
            blas[i].tris = self.tris.ctypes.data + to * 112
            blas[i].tri_count = tc
            blas[i].T = abi.f16(*b["T"].tolist())
            blas[i].inv_T = abi.f16(*b["inv_T"].tolist())
            blas[i].obj_idx = int(b["obj_idx"])
            blas[i].mat_idx = int(b["mat_idx"])
        nt = len(self.tex_table)
        tex = (abi.rt_texture * max(nt, 1))()
        for i, t in enumerate(self.tex_table):
            if int(t["pixel_offset"]) + int(t["width"]) * int(t["height"]) > len(self.tex_pixels) or int(t["width"]) < 0 or int(t["height"]) < 0:
                raise ValueError(f"texture {i}: texel range outside tex_pixels")
            tex[i].pixels = self.tex_pixels.ctypes.data + int(t["pixel_offset"]) * 4
            tex[i].width, tex[i].height = int(t["width"]), int(t["height"])
        d = abi.rt_scene_desc()
        d.kind = int(h["kind"])
        d.blas = C.cast(blas, C.POINTER(abi.rt_blas_desc))
        d.blas_count = nb
        if not self.device_tlas:
            d.tlas_nodes = self.tlas_nodes.ctypes.data if len(self.tlas_nodes) else None
            d.tlas_node_count = len(self.tlas_nodes)
            if self.tlas_nodes32 is not None and not len(self.tlas_nodes):
                d.tlas_nodes32, d.tlas_node32_count = self.tlas_nodes32.ctypes.data, len(self.tlas_nodes32)
        d.obj_material = self.obj_material.ctypes.data
        d.obj_count = len(self.obj_material)
        d.materials = self.materials.ctypes.data
        d.material_count = len(self.materials)
        d.textures = C.cast(tex, C.POINTER(abi.rt_texture))
        d.texture_count = nt
        d.skydome_texture, d.floor_texture = int(h["skydome_texture"]), int(h["floor_texture"])
        d.floor_n = abi.f3(*h["floor_n"].tolist())
        d.floor_d, d.floor_invto = float(h["floor_d"]), float(h["floor_invto"])
        d.light_T = abi.f16(*h["light_T"].tolist())
        d.light_inv_T = abi.f16(*h["light_inv_T"].tolist())
        d.light_size = float(h["light_size"])
        d.light_color = abi.f3(*h["light_color"].tolist())
        d.light_pos = abi.f3(*h["light_pos"].tolist())
        grid = None
        if self.blas_kd_table is not None or self.blas_grid_table is not None:
            # TLASFileScene over per-object KD-trees / grids
            accel = (abi.rt_blas_accel * nb)()
            grids = (abi.rt_grid_desc * nb)()
            for i in range(nb):
                if self.blas_kd_table is not None:
                    k = self.blas_kd_table[i]
                    accel[i].kd_nodes = self.kd_nodes.ctypes.data + int(k["node_offset"]) * 48
                    accel[i].kd_node_count = int(k["node_count"])
                    accel[i].kd_tri_indices = self.kd_tri_indices.ctypes.data + int(k["idx_offset"]) * 4
                    accel[i].kd_tri_index_count = int(k["idx_count"])
                else:
                    g = self.blas_grid_table[i]
                    grids[i].resolution = (C.c_int32 * 3)(*g["resolution"].tolist())
                    grids[i].cell_size = abi.f3(*g["cell_size"].tolist())
                    grids[i].bounds_min = abi.f3(*g["bounds_min"].tolist())
                    grids[i].bounds_max = abi.f3(*g["bounds_max"].tolist())
                    grids[i].cell_start = self.grid_cell_start.ctypes.data + int(g["cell_offset"]) * 4
                    grids[i].tri_indices = self.grid_tri_indices.ctypes.data + int(g["idx_offset"]) * 4
                    grids[i].index_count = int(g["idx_count"])
                    accel[i].grid = C.pointer(grids[i])
            d.blas_accel = C.cast(accel, C.POINTER(abi.rt_blas_accel))
            self._keep = (blas, tex, accel, grids)
            return d
        if self.kd_nodes is not None:
            d.kd_nodes, d.kd_node_count = self.kd_nodes.ctypes.data, len(self.kd_nodes)
            d.kd_tri_indices, d.kd_tri_index_count = self.kd_tri_indices.ctypes.data, len(self.kd_tri_indices)
        if self.grid_header is not None:
            g = self.grid_header[0]
            grid = abi.rt_grid_desc()
            grid.resolution = (C.c_int32 * 3)(*g["resolution"].tolist())
            grid.cell_size = abi.f3(*g["cell_size"].tolist())
            grid.bounds_min = abi.f3(*g["bounds_min"].tolist())
            grid.bounds_max = abi.f3(*g["bounds_max"].tolist())
            grid.cell_start = self.grid_cell_start.ctypes.data
            grid.tri_indices = self.grid_tri_indices.ctypes.data
            grid.index_count = len(self.grid_tri_indices)
            d.grid = C.pointer(grid)
        self._keep = (blas, tex, grid)
        return d
