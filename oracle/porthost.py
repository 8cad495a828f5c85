"""ctypes harness over oracle/liboracle.so, the C restatement in oracle/rt_oracle.c (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm may import this."""
import ctypes as C
import os
import subprocess

import numpy as np

import cpu_ray_tracer_b200 as rtb
from cpu_ray_tracer_b200 import abi

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")
SRC = os.path.join(HERE, "rt_oracle.c")

STATS_FIELDS = ["extension_rays", "shadow_rays", "paths", "interior_visits", "tlas_interior_visits",
                "leaf_visits", "tri_tests", "blas_entries"]


def build(force=False):
    """gcc, strict IEEE: no FMA contraction, no fast-math (matches oracle/ref_build flags)."""
    hdr = os.path.join(os.path.dirname(HERE), "include", "rt_b200.h")
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= max(os.path.getmtime(SRC), os.path.getmtime(hdr)):
        return LIB
    subprocess.run(["gcc", "-std=gnu11", "-O2", "-fopenmp", "-ffp-contract=off", "-fPIC", "-shared", "-w",
                    SRC, "-o", LIB, "-lm"], check=True)
    return LIB


class PortOracle:
    def __init__(self, scene: "rtb.FlatScene"):
        self.lib = C.CDLL(build())
        self.scene = scene
        self.desc = scene.desc()
        self._p = C.byref(self.desc)
        L = self.lib
        vp = C.c_void_p
        L.orc_find_nearest.argtypes = [vp, vp, vp, C.c_size_t, vp]
        L.orc_is_occluded.argtypes = [vp, vp, vp, C.c_size_t, vp]
        L.orc_primary_rays.argtypes = [vp, C.c_int, C.c_int, vp]
        L.orc_hit_info.argtypes = [vp, vp, vp, C.c_size_t, vp, vp, vp]
        L.orc_render_whitted.argtypes = [vp, vp, vp, vp, vp]
        L.orc_render_pt.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_int, vp, vp]
        L.orc_to_rgb8.argtypes = [vp, C.c_size_t, C.c_float, vp]
        L.orc_camera_default.argtypes = [vp, C.c_int, C.c_int]
        L.orc_camera_look_at.argtypes = [vp, vp, vp, C.c_int, C.c_int]
        L.orc_refit_bvh.argtypes = [vp, C.c_uint32, vp, vp, C.c_int]
        assert L.orc_sizeof_stats() == 8 * len(STATS_FIELDS)

    @staticmethod
    def _stats(arr):
        return dict(zip(STATS_FIELDS, (int(x) for x in arr)))

    def camera_default(self, w, h):
        cam = abi.rt_camera()
        self.lib.orc_camera_default(C.byref(cam), w, h)
        return cam

    def camera_look_at(self, pos, target, w, h):
        cam = abi.rt_camera()
        p = np.asarray(pos, np.float32)
        t = np.asarray(target, np.float32)
        self.lib.orc_camera_look_at(C.byref(cam), p.ctypes.data, t.ctypes.data, w, h)
        return cam

    def primary_rays(self, cam, w, h):
        rays = np.zeros(w * h, abi.RAY_DTYPE)
        self.lib.orc_primary_rays(C.byref(cam), w, h, rays.ctypes.data)
        return rays

    def find_nearest(self, rays):
        rays = np.ascontiguousarray(rays, abi.RAY_DTYPE)
        hits = np.zeros(len(rays), abi.HIT_DTYPE)
        st = np.zeros(len(STATS_FIELDS), np.uint64)
        self.lib.orc_find_nearest(self._p, rays.ctypes.data, hits.ctypes.data, len(rays), st.ctypes.data)
        return hits, self._stats(st)

    def is_occluded(self, rays):
        rays = np.ascontiguousarray(rays, abi.RAY_DTYPE)
        out = np.zeros(len(rays), np.uint8)
        st = np.zeros(len(STATS_FIELDS), np.uint64)
        self.lib.orc_is_occluded(self._p, rays.ctypes.data, out.ctypes.data, len(rays), st.ctypes.data)
        return out, self._stats(st)

    def hit_info(self, rays, hits):
        n = len(rays)
        N = np.zeros((n, 3), np.float32)
        uv = np.zeros((n, 2), np.float32)
        albedo = np.zeros((n, 3), np.float32)
        self.lib.orc_hit_info(self._p, np.ascontiguousarray(rays).ctypes.data, np.ascontiguousarray(hits).ctypes.data,
                              n, N.ctypes.data, uv.ctypes.data, albedo.ctypes.data)
        return N, uv, albedo

    def render_whitted(self, cam, params):
        acc = np.zeros((params.height, params.width, 4), np.float32)
        st = np.zeros(len(STATS_FIELDS), np.uint64)
        self.lib.orc_render_whitted(self._p, C.byref(cam), C.byref(params), acc.ctypes.data, st.ctypes.data)
        return acc, self._stats(st)

    def render_pt(self, cam, params, first_spp=1, count=1, stride=1, accumulator=None):
        acc = np.zeros((params.height, params.width, 4), np.float32) if accumulator is None else accumulator
        st = np.zeros(len(STATS_FIELDS), np.uint64)
        self.lib.orc_render_pt(self._p, C.byref(cam), C.byref(params), first_spp, count, stride, acc.ctypes.data, st.ctypes.data)
        return acc, self._stats(st)

    def to_rgb8(self, acc, scale):
        acc = np.ascontiguousarray(acc, np.float32)
        out = np.zeros(acc.shape[:-1], np.uint32)
        self.lib.orc_to_rgb8(acc.ctypes.data, out.size, scale, out.ctypes.data)
        return out


def refit_bvh(nodes, tris, tri_indices, all_nodes=False):
    """BVH::Refit (bvh.cpp:26-43) restated (orc_refit_bvh): returns the refitted copy of `nodes`"""
    L = C.CDLL(build())
    L.orc_refit_bvh.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_int]
    out = np.array(nodes, abi.NODE_DTYPE, copy=True)
    tris = np.ascontiguousarray(tris, abi.TRI_DTYPE)
    idx = np.ascontiguousarray(tri_indices, np.uint32)
    L.orc_refit_bvh(out.ctypes.data, len(out), tris.ctypes.data, idx.ctypes.data, 1 if all_nodes else 0)
    return out


def default_params(integrator, w, h, seed_mode=abi.RT_SEED_REFERENCE_TILE, depth_limit=5, passes=1):
    p = abi.rt_render_params()
    p.integrator, p.width, p.height = integrator, w, h
    p.depth_limit, p.epsilon, p.seed_mode = depth_limit, 0.001, seed_mode
    p.tile_begin, p.tile_end, p.max_frames_in_flight, p.schedule = 0, 0, 0, 0
    p.passes = passes
    return p
