"""ctypes mirror of include/rt_b200.h (the C-ABI).  Field order and types must match the header exactly;
tests/test_abi.py checks the struct sizes against the sizes the compiled library reports."""
import ctypes as C

import numpy as np

RT_OK, RT_ERR_INVALID, RT_ERR_NO_DEVICE, RT_ERR_CUDA, RT_ERR_UNSUPPORTED = 0, -1, -2, -3, -4
RT_SCENE_FLAT, RT_SCENE_TLAS, RT_SCENE_FLAT_KDTREE, RT_SCENE_FLAT_GRID, RT_SCENE_TLAS_KDTREE, RT_SCENE_TLAS_GRID = 0, 1, 2, 3, 4, 5
TLAS_KINDS = (RT_SCENE_TLAS, RT_SCENE_TLAS_KDTREE, RT_SCENE_TLAS_GRID)
RT_SCENE_FLAG_COUNTERS = 1
RT_INTEGRATOR_WHITTED, RT_INTEGRATOR_PATH = 0, 1
RT_SEED_REFERENCE_TILE, RT_SEED_PER_PIXEL = 0, 1
RT_SCHEDULE_AUTO, RT_SCHEDULE_WAVEFRONT, RT_SCHEDULE_STREAMS = 0, 1, 2
RT_MATH_EXPF, RT_MATH_ACOSF, RT_MATH_ATAN2F, RT_MATH_SKY_TEXEL = 0, 1, 2, 3
RT_IPC_HANDLE_BYTES = 64
RT_REFIT_ALL_NODES, RT_REFIT_REBUILD_TLAS = 1, 2
RT_RAYS_DEFAULT, RT_RAYS_INCOHERENT = 0, 1

f3 = C.c_float * 3
f16 = C.c_float * 16


class rt_blas_desc(C.Structure):
    _fields_ = [("nodes", C.c_void_p), ("node_count", C.c_uint32),
                ("tris", C.c_void_p), ("tri_indices", C.c_void_p), ("tri_count", C.c_uint32),
                ("T", f16), ("inv_T", f16), ("obj_idx", C.c_int32), ("mat_idx", C.c_int32)]


class rt_material(C.Structure):
    _fields_ = [("reflectivity", C.c_float), ("refractivity", C.c_float), ("absorption", f3),
                ("albedo", f3), ("is_light", C.c_int32), ("texture", C.c_int32)]


class rt_texture(C.Structure):
    _fields_ = [("pixels", C.c_void_p), ("width", C.c_int32), ("height", C.c_int32)]


class rt_grid_desc(C.Structure):
    _fields_ = [("resolution", C.c_int32 * 3), ("cell_size", f3), ("bounds_min", f3), ("bounds_max", f3),
                ("cell_start", C.c_void_p), ("tri_indices", C.c_void_p), ("index_count", C.c_uint32)]


class rt_blas_accel(C.Structure):
    _fields_ = [("kd_nodes", C.c_void_p), ("kd_node_count", C.c_uint32),
                ("kd_tri_indices", C.c_void_p), ("kd_tri_index_count", C.c_uint32),
                ("grid", C.POINTER(rt_grid_desc))]


class rt_scene_desc(C.Structure):
    _fields_ = [("kind", C.c_int32),
                ("blas", C.POINTER(rt_blas_desc)), ("blas_count", C.c_uint32),
                ("tlas_nodes", C.c_void_p), ("tlas_node_count", C.c_uint32),
                ("obj_material", C.c_void_p), ("obj_count", C.c_uint32),
                ("materials", C.c_void_p), ("material_count", C.c_uint32),
                ("textures", C.POINTER(rt_texture)), ("texture_count", C.c_uint32),
                ("skydome_texture", C.c_int32), ("floor_texture", C.c_int32),
                ("floor_n", f3), ("floor_d", C.c_float), ("floor_invto", C.c_float),
                ("light_T", f16), ("light_inv_T", f16), ("light_size", C.c_float),
                ("light_color", f3), ("light_pos", f3),
                ("kd_nodes", C.c_void_p), ("kd_node_count", C.c_uint32),
                ("kd_tri_indices", C.c_void_p), ("kd_tri_index_count", C.c_uint32),
                ("grid", C.POINTER(rt_grid_desc)),
                ("blas_accel", C.POINTER(rt_blas_accel)),
                ("tlas_nodes32", C.c_void_p), ("tlas_node32_count", C.c_uint32)]


class rt_camera(C.Structure):
    _fields_ = [("pos", f3), ("top_left", f3), ("top_right", f3), ("bottom_left", f3)]


class rt_render_params(C.Structure):
    _fields_ = [("integrator", C.c_int32), ("width", C.c_int32), ("height", C.c_int32),
                ("depth_limit", C.c_int32), ("epsilon", C.c_float), ("seed_mode", C.c_int32),
                ("tile_begin", C.c_int32), ("tile_end", C.c_int32), ("max_frames_in_flight", C.c_int32),
                ("schedule", C.c_int32), ("lookahead_frames", C.c_int32), ("passes", C.c_int32), ("tile_step", C.c_int32)]


class rt_counters(C.Structure):
    _fields_ = [("extension_rays", C.c_uint64), ("shadow_rays", C.c_uint64), ("paths", C.c_uint64),
                ("wavefront_iterations", C.c_uint64), ("kernel_launches", C.c_uint64)]


class rt_scene_info(C.Structure):
    _fields_ = [("fat_nodes", C.c_uint64), ("triangle_slots", C.c_uint64), ("instances", C.c_uint64), ("meshes", C.c_uint64),
                ("bytes_geometry", C.c_uint64), ("bytes_textures", C.c_uint64), ("stack_entries", C.c_int32), ("max_blas_depth", C.c_int32)]


class rt_stage_times(C.Structure):
    _fields_ = [("ms", C.c_double * 5), ("launches", C.c_uint64 * 5)]


STAGES = ("generate", "extend", "shade", "connect", "accumulate")

# numpy record layouts of the POD arrays (same bytes as the reference's structs)
NODE_DTYPE = np.dtype([("aabb_min", "<f4", 3), ("aabb_max", "<f4", 3), ("left_first", "<u4"), ("tri_count", "<u4")])
TRI_DTYPE = np.dtype([("v0", "<f4", 3), ("v1", "<f4", 3), ("v2", "<f4", 3),
                      ("n0", "<f4", 3), ("n1", "<f4", 3), ("n2", "<f4", 3),
                      ("uv0", "<f4", 2), ("uv1", "<f4", 2), ("uv2", "<f4", 2),
                      ("centroid", "<f4", 3), ("obj_idx", "<i4")])
TLAS_NODE_DTYPE = np.dtype([("aabb_min", "<f4", 3), ("left_right", "<u4"), ("aabb_max", "<f4", 3), ("blas", "<u4")])
TLAS_NODE32_DTYPE = np.dtype([("aabb_min", "<f4", 3), ("left", "<u4"), ("aabb_max", "<f4", 3), ("right", "<u4")])
KD_NODE_DTYPE = np.dtype([("aabb_min", "<f4", 3), ("left", "<i4"), ("aabb_max", "<f4", 3), ("right", "<i4"),
                          ("split_axis", "<i4"), ("split_distance", "<f4"), ("tri_start", "<u4"), ("tri_count", "<u4")])
GRID_HEADER_DTYPE = np.dtype([("resolution", "<i4", 3), ("cell_size", "<f4", 3), ("bounds_min", "<f4", 3), ("bounds_max", "<f4", 3)])
BLAS_KD_TABLE_DTYPE = np.dtype([("node_offset", "<u4"), ("node_count", "<u4"), ("idx_offset", "<u4"), ("idx_count", "<u4")])
BLAS_GRID_TABLE_DTYPE = np.dtype([("resolution", "<i4", 3), ("cell_size", "<f4", 3), ("bounds_min", "<f4", 3), ("bounds_max", "<f4", 3),
                                  ("cell_offset", "<u4"), ("cell_count", "<u4"), ("idx_offset", "<u4"), ("idx_count", "<u4")])
MATERIAL_DTYPE = np.dtype([("reflectivity", "<f4"), ("refractivity", "<f4"), ("absorption", "<f4", 3),
                           ("albedo", "<f4", 3), ("is_light", "<i4"), ("texture", "<i4")])
RAY_DTYPE = np.dtype([("O", "<f4", 3), ("tmax", "<f4"), ("D", "<f4", 3), ("inside", "<i4")])
HIT_DTYPE = np.dtype([("t", "<f4"), ("u", "<f4"), ("v", "<f4"), ("obj_idx", "<i4"), ("tri_idx", "<i4"),
                      ("traversed", "<i4"), ("tested", "<i4"), ("reserved", "<i4")])
assert NODE_DTYPE.itemsize == 32 and TRI_DTYPE.itemsize == 112 and TLAS_NODE_DTYPE.itemsize == 32
assert TLAS_NODE32_DTYPE.itemsize == 32
assert KD_NODE_DTYPE.itemsize == 48 and GRID_HEADER_DTYPE.itemsize == 48
assert MATERIAL_DTYPE.itemsize == 40 and RAY_DTYPE.itemsize == 32 and HIT_DTYPE.itemsize == 32
