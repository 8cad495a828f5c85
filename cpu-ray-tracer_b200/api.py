"""Host-side mirror of the reference's Scene / Renderer surface over the C-ABI (include/rt_b200.h).

Same names, argument meaning and error behaviour as the reference classes, batched:

    GpuFileScene / GpuTLASFileScene   <->  FileScene / TLASFileScene  (infra/scene/*.h):
        FindNearest, IsOccluded, GetLightPos, GetLightColor, GetTriangleCount
    Camera                            <->  Tmpl8::Camera (template/camera.h): SetCameraState
    GpuRenderer                       <->  Renderer : TheApp (2. WhittedStyle / 3. PathTracer renderer.h):
        Init, Tick, ClearAccumulator, accumulator, camera, spp, passes, depthLimit

There is no CPU fallback anywhere in this module: if librt_b200.so is missing or no CUDA device is
visible the calls raise (RtError / OSError).  The C++ twin of these adapters is host/rt_b200_adapters.h.
"""
import ctypes as C
import os

import numpy as np

from . import abi
from .scene_file import FlatScene

HERE = os.path.dirname(os.path.abspath(__file__))
# RT_B200_LIB: another build of the same library (development A/B runs, e.g. build.py --cuda-math); the default is in-tree
LIB_PATH = os.environ.get("RT_B200_LIB") or os.path.join(HERE, "librt_b200.so")

EXPORTS = [
    "rt_last_error", "rt_abi_version", "rt_device_count", "rt_scene_create", "rt_scene_destroy",
    "rt_find_nearest", "rt_is_occluded", "rt_find_nearest_device", "rt_is_occluded_device",
    "rt_camera_default", "rt_camera_look_at", "rt_render_params_default",
    "rt_renderer_create", "rt_renderer_destroy", "rt_renderer_set_stream", "rt_renderer_set_accumulator",
    "rt_renderer_set_camera", "rt_renderer_set_passes", "rt_renderer_clear", "rt_renderer_render", "rt_renderer_sync",
    "rt_renderer_read_accumulator", "rt_renderer_read_pixels", "rt_renderer_device_accumulator",
    "rt_renderer_get_counters", "rt_renderer_reset_counters",
    "rt_renderer_set_profiling", "rt_renderer_get_stage_times", "rt_renderer_get_launch_spans",
    "rt_renderer_get_queue_history", "rt_measure_gather_bandwidth", "rt_build_bvh", "rt_eval_shading_math",
    "rt_measure_l2_stream_bandwidth", "rt_renderer_export_accumulator", "rt_renderer_import_accumulator",
    "rt_multi_renderer_create", "rt_multi_renderer_destroy", "rt_multi_renderer_device_count", "rt_multi_renderer_set_camera",
    "rt_multi_renderer_set_passes", "rt_multi_renderer_clear", "rt_multi_renderer_render", "rt_multi_renderer_sync",
    "rt_multi_renderer_read_accumulator", "rt_multi_renderer_read_pixels", "rt_multi_renderer_get_counters",
    "rt_multi_renderer_reset_counters",
    "rt_build_tlas", "rt_scene_refit", "rt_scene_download_bvh", "rt_scene_get_info", "rt_scene_validate", "rt_find_nearest_device_ex",
]


class RtError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"rt_b200 status {status}: {msg}")
        self.status = status


_lib = None


def lib():
    """Loads the CUDA library; raises if it was not built (never substitutes a CPU path)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OSError(f"{LIB_PATH} not built: run `python __graft_entry__.py build` (nvcc, sm_100a). "
                      "There is no CPU fallback for the ray core.")
    L = C.CDLL(LIB_PATH)
    vp, i32, sz = C.c_void_p, C.c_int, C.c_size_t
    L.rt_last_error.restype = C.c_char_p
    L.rt_scene_create.argtypes = [C.POINTER(abi.rt_scene_desc), i32, C.c_uint32, C.POINTER(vp)]
    L.rt_scene_destroy.argtypes = [vp]
    L.rt_scene_destroy.restype = None
    L.rt_find_nearest.argtypes = [vp, vp, vp, sz]
    L.rt_is_occluded.argtypes = [vp, vp, vp, sz]
    L.rt_find_nearest_device.argtypes = [vp, vp, vp, sz, vp]
    L.rt_is_occluded_device.argtypes = [vp, vp, vp, sz, vp]
    L.rt_find_nearest_device_ex.argtypes = [vp, vp, vp, sz, vp, C.c_uint32]
    L.rt_camera_default.argtypes = [C.POINTER(abi.rt_camera), i32, i32]
    L.rt_camera_default.restype = None
    L.rt_camera_look_at.argtypes = [C.POINTER(abi.rt_camera), C.POINTER(C.c_float), C.POINTER(C.c_float), i32, i32]
    L.rt_camera_look_at.restype = None
    L.rt_render_params_default.argtypes = [C.POINTER(abi.rt_render_params), i32, i32, i32]
    L.rt_render_params_default.restype = None
    L.rt_renderer_create.argtypes = [vp, C.POINTER(abi.rt_render_params), C.POINTER(vp)]
    L.rt_renderer_destroy.argtypes = [vp]
    L.rt_renderer_destroy.restype = None
    L.rt_renderer_set_stream.argtypes = [vp, vp]
    L.rt_renderer_set_accumulator.argtypes = [vp, vp]
    L.rt_renderer_set_camera.argtypes = [vp, C.POINTER(abi.rt_camera)]
    L.rt_renderer_set_passes.argtypes = [vp, i32]
    L.rt_renderer_clear.argtypes = [vp]
    L.rt_renderer_render.argtypes = [vp, i32, i32, i32]
    L.rt_renderer_sync.argtypes = [vp]
    L.rt_renderer_read_accumulator.argtypes = [vp, vp]
    L.rt_renderer_read_pixels.argtypes = [vp, C.c_float, vp]
    L.rt_renderer_device_accumulator.argtypes = [vp]
    L.rt_renderer_device_accumulator.restype = vp
    L.rt_renderer_get_counters.argtypes = [vp, C.POINTER(abi.rt_counters)]
    L.rt_renderer_reset_counters.argtypes = [vp]
    L.rt_renderer_set_profiling.argtypes = [vp, i32]
    L.rt_renderer_get_stage_times.argtypes = [vp, C.POINTER(abi.rt_stage_times)]
    L.rt_renderer_get_launch_spans.argtypes = [vp, vp, vp, sz, C.POINTER(sz)]
    L.rt_renderer_get_queue_history.argtypes = [vp, vp, sz, C.POINTER(sz)]
    L.rt_measure_gather_bandwidth.argtypes = [i32, sz, i32, C.POINTER(C.c_double)]
    L.rt_build_bvh.argtypes = [i32, vp, C.c_uint32, vp, vp, C.POINTER(C.c_uint32), C.POINTER(C.c_double)]
    L.rt_eval_shading_math.argtypes = [i32, i32, vp, vp, vp, sz]
    L.rt_measure_l2_stream_bandwidth.argtypes = [i32, sz, C.POINTER(C.c_double)]
    L.rt_renderer_export_accumulator.argtypes = [vp, vp]
    L.rt_renderer_import_accumulator.argtypes = [vp, vp]
    L.rt_multi_renderer_create.argtypes = [C.POINTER(abi.rt_scene_desc), C.c_uint32, C.POINTER(i32), i32, C.POINTER(abi.rt_render_params), C.POINTER(vp)]
    L.rt_multi_renderer_destroy.argtypes = [vp]
    L.rt_multi_renderer_destroy.restype = None
    L.rt_multi_renderer_device_count.argtypes = [vp]
    L.rt_multi_renderer_set_camera.argtypes = [vp, C.POINTER(abi.rt_camera)]
    L.rt_multi_renderer_set_passes.argtypes = [vp, i32]
    L.rt_multi_renderer_clear.argtypes = [vp]
    L.rt_multi_renderer_render.argtypes = [vp, i32, i32, i32]
    L.rt_multi_renderer_sync.argtypes = [vp]
    L.rt_multi_renderer_read_accumulator.argtypes = [vp, vp]
    L.rt_multi_renderer_read_pixels.argtypes = [vp, C.c_float, vp]
    L.rt_multi_renderer_get_counters.argtypes = [vp, C.POINTER(abi.rt_counters)]
    L.rt_multi_renderer_reset_counters.argtypes = [vp]
    L.rt_build_tlas.argtypes = [i32, vp, C.c_uint32, vp, C.POINTER(C.c_uint32), C.POINTER(C.c_double)]
    L.rt_scene_refit.argtypes = [vp, C.c_uint32, vp, C.c_uint32, C.c_uint32]
    L.rt_scene_download_bvh.argtypes = [vp, C.c_uint32, vp, vp, C.POINTER(C.c_uint32)]
    L.rt_scene_get_info.argtypes = [vp, C.POINTER(abi.rt_scene_info)]
    L.rt_scene_validate.argtypes = [vp]
    _lib = L
    return L


def _check(status):
    if status != abi.RT_OK:
        raise RtError(status, lib().rt_last_error().decode(errors="replace"))


def device_count():
    return lib().rt_device_count()


def build_bvh_gpu(tris, device=0):
    """BVH::Build on the GPU (rt_build_bvh): TRI_DTYPE array -> (nodes[:nodesUsed], tri_indices, kernel milliseconds)"""
    tris = np.ascontiguousarray(tris, abi.TRI_DTYPE)
    n = len(tris)
    nodes = np.zeros(max(2 * n - 1, 1), abi.NODE_DTYPE)
    idx = np.zeros(n, np.uint32)
    used, ms = C.c_uint32(), C.c_double()
    _check(lib().rt_build_bvh(device, tris.ctypes.data, n, nodes.ctypes.data, idx.ctypes.data, C.byref(used), C.byref(ms)))
    return nodes[:used.value].copy(), idx, ms.value


def measure_gather_bandwidth(working_set_bytes, bypass_l1=True, device=0):
    """GB/s of random 64-byte gathers over a working set (L2 roofline denominator for L2-resident scenes)"""
    out = C.c_double()
    _check(lib().rt_measure_gather_bandwidth(device, working_set_bytes, 1 if bypass_l1 else 0, C.byref(out)))
    return out.value


def measure_l2_stream_bandwidth(working_set_bytes=48 << 20, device=0):
    """GB/s of coalesced streaming reads of an L2-resident buffer, L1 bypassed (the L2 peak of the roofline)"""
    out = C.c_double()
    _check(lib().rt_measure_l2_stream_bandwidth(device, working_set_bytes, C.byref(out)))
    return out.value


def eval_shading_math(fn, a, b=None, device=0):
    """expf / acosf / atan2f as the shading kernels compute them, evaluated on the device (rt_eval_shading_math)"""
    a = np.ascontiguousarray(a, np.float32)
    b = None if b is None else np.ascontiguousarray(b, np.float32)
    out = np.empty_like(a)
    _check(lib().rt_eval_shading_math(device, fn, a.ctypes.data, None if b is None else b.ctypes.data, out.ctypes.data, a.size))
    return out


def build_tlas_gpu(bounds, device=0, return_ms=False):
    """TLASBVH::Build (tlas_bvh.cpp:17-70) on the GPU over an (n, 6) array of world bounds: rt_build_tlas.
    Returns the 2n nodes in the reference's order with 32-bit children (abi.TLAS_NODE32_DTYPE)."""
    bounds = np.ascontiguousarray(bounds, np.float32).reshape(-1, 6)
    n = len(bounds)
    out = np.zeros(2 * n, abi.TLAS_NODE32_DTYPE)
    used, ms = C.c_uint32(), C.c_double()
    _check(lib().rt_build_tlas(device, bounds.ctypes.data, n, out.ctypes.data, C.byref(used), C.byref(ms)))
    return (out[:used.value], ms.value) if return_ms else out[:used.value]


class Camera:
    """Tmpl8::Camera (template/camera.h): default view (0,0,-2) -> +z, SetCameraState."""

    def __init__(self, width, height):
        self.width, self.height = width, height
        self.c = abi.rt_camera()
        lib().rt_camera_default(C.byref(self.c), width, height)

    def SetCameraState(self, position, target):
        p = (C.c_float * 3)(*[float(x) for x in position])
        t = (C.c_float * 3)(*[float(x) for x in target])
        lib().rt_camera_look_at(C.byref(self.c), p, t, self.width, self.height)

    @property
    def camPos(self):
        return np.array(list(self.c.pos), np.float32)


def make_rays(O, D, tmax=1e34, inside=0):
    O = np.asarray(O, np.float32)
    rays = np.zeros(O.shape[0], abi.RAY_DTYPE)
    rays["O"], rays["D"], rays["tmax"], rays["inside"] = O, np.asarray(D, np.float32), tmax, inside
    return rays


class GpuScene:
    """BaseScene surface (infra/scene/base_scene.h:16-32) served by the CUDA library."""

    KIND = None

    def __init__(self, flat: FlatScene, device=0, counters=False):
        if isinstance(flat, (str, os.PathLike)):
            flat = FlatScene.load(flat)
        if self.KIND is not None and flat.kind not in self.KIND:
            raise ValueError(f"{type(self).__name__} needs a scene of kind {self.KIND}, got {flat.kind}")
        self.flat = flat
        self.device = device
        self.handle = C.c_void_p()
        desc = flat.desc()
        _check(lib().rt_scene_create(C.byref(desc), device, abi.RT_SCENE_FLAG_COUNTERS if counters else 0, C.byref(self.handle)))

    def close(self):
        if getattr(self, "handle", None):
            lib().rt_scene_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- BaseScene ---------------------------------------------------------------------------
    def FindNearest(self, rays):
        """Batched FindNearest: rays = RAY_DTYPE array (host). Returns HIT_DTYPE array; miss <=> obj_idx == -1."""
        rays = np.ascontiguousarray(rays, abi.RAY_DTYPE)
        hits = np.empty(len(rays), abi.HIT_DTYPE)
        _check(lib().rt_find_nearest(self.handle, rays.ctypes.data, hits.ctypes.data, len(rays)))
        return hits

    def IsOccluded(self, rays):
        rays = np.ascontiguousarray(rays, abi.RAY_DTYPE)
        out = np.empty(len(rays), np.uint8)
        _check(lib().rt_is_occluded(self.handle, rays.ctypes.data, out.ctypes.data, len(rays)))
        return out

    def FindNearestDevice(self, d_rays_ptr, d_hits_ptr, n, stream=None, incoherent=False):
        """Buffers already in HBM (e.g. torch tensors' data_ptr()); asynchronous on `stream`.  incoherent=True: the batch is
        bounce rays, not camera rays (RT_RAYS_INCOHERENT: same hits, the traversal kernel that suits them)."""
        _check(lib().rt_find_nearest_device_ex(self.handle, d_rays_ptr, d_hits_ptr, n, stream, abi.RT_RAYS_INCOHERENT if incoherent else 0))

    def IsOccludedDevice(self, d_rays_ptr, d_out_ptr, n, stream=None):
        _check(lib().rt_is_occluded_device(self.handle, d_rays_ptr, d_out_ptr, n, stream))

    def GetLightPos(self):
        return self.flat.header["light_pos"][0].copy()

    def GetLightColor(self):
        return self.flat.header["light_color"][0].copy()

    def GetTriangleCount(self):
        return self.flat.triangle_count

    def info(self):
        """what the scene occupies on the device (rt_scene_get_info): meshes < instances = shared geometry"""
        i = abi.rt_scene_info()
        _check(lib().rt_scene_get_info(self.handle, C.byref(i)))
        return {k: int(getattr(i, k)) for k, _ in abi.rt_scene_info._fields_}

    def validate(self):
        """rt_scene_validate: structural check of the traversal data read back from device memory (raises RtError)"""
        _check(lib().rt_scene_validate(self.handle))

    # -- scene construction steps on the device -----------------------------------
    def Refit(self, blas_index, tris, all_nodes=False, rebuild_tlas=False):
        """BVH::Refit / BLASBVH::Refit (bvh.cpp:26-43) for the mesh of BLAS `blas_index` after its triangles moved:
        rt_scene_refit.  `tris`: the mesh's TRI_DTYPE array in its own order (same count)."""
        tris = np.ascontiguousarray(tris, abi.TRI_DTYPE)
        flags = (abi.RT_REFIT_ALL_NODES if all_nodes else 0) | (abi.RT_REFIT_REBUILD_TLAS if rebuild_tlas else 0)
        _check(lib().rt_scene_refit(self.handle, blas_index, tris.ctypes.data, len(tris), flags))

    def download_bvh(self, blas_index=0):
        """the device's traversal layout of one mesh read back as the reference's arrays: (nodes[:nodesUsed], tri_indices)"""
        n = int(self.flat.blas_table[blas_index]["tri_count"])
        nodes, idx, used = np.zeros(2 * n - 1, abi.NODE_DTYPE), np.zeros(n, np.uint32), C.c_uint32()
        _check(lib().rt_scene_download_bvh(self.handle, blas_index, nodes.ctypes.data, idx.ctypes.data, C.byref(used)))
        return nodes[:used.value], idx


class GpuFileScene(GpuScene):
    """FileScene (infra/scene/file_scene.h) with the accelerator its file_scene.h:10-12 switch selects: one flat
    SAH BVH over all triangles (USE_BVH), the KD-tree it ships with (USE_KDTree) or the uniform grid (USE_Grid)."""
    KIND = (abi.RT_SCENE_FLAT, abi.RT_SCENE_FLAT_KDTREE, abi.RT_SCENE_FLAT_GRID)


class GpuTLASFileScene(GpuScene):
    """TLASFileScene (infra/scene/tlas_file_scene.h): the agglomerative TLAS over per-object BLAS of the kind its
    tlas_file_scene.h:12-14 switch selects - BVH (TLAS_USE_BVH), KD-tree (TLAS_USE_KDTree) or grid (TLAS_USE_Grid)."""
    KIND = abi.TLAS_KINDS


def open_scene(path_or_flat, device=0, counters=False):
    flat = FlatScene.load(path_or_flat) if isinstance(path_or_flat, (str, os.PathLike)) else path_or_flat
    cls = GpuTLASFileScene if flat.kind in abi.TLAS_KINDS else GpuFileScene
    return cls(flat, device=device, counters=counters)


class GpuRenderer:
    """Renderer : TheApp (2. WhittedStyle/renderer.h:41-61, 3. PathTracer/renderer.h:29-53)."""

    def __init__(self, scene: GpuScene, integrator, width, height, depthLimit=5, seed_mode=abi.RT_SEED_REFERENCE_TILE,
                 tile_begin=0, tile_end=0, max_frames_in_flight=0, schedule=abi.RT_SCHEDULE_AUTO, lookahead_frames=0, tile_step=1):
        self.scene = scene
        self.integrator = integrator
        self.width, self.height = width, height
        self.depthLimit = depthLimit
        self.camera = Camera(width, height)
        self.spp, self.passes = 1, 1  # renderer.h:50
        self.params = abi.rt_render_params()
        lib().rt_render_params_default(C.byref(self.params), integrator, width, height)
        self.params.depth_limit = depthLimit
        self.params.seed_mode = seed_mode
        self.params.tile_begin, self.params.tile_end, self.params.tile_step = tile_begin, tile_end, tile_step
        self.params.max_frames_in_flight = max_frames_in_flight
        self.params.schedule = schedule
        self.params.lookahead_frames = lookahead_frames
        self.handle = C.c_void_p()
        self._initialised = False

    def Init(self):
        """Renderer::Init: allocate + clear the accumulator (renderer.cpp:8-13)."""
        if self.handle:
            lib().rt_renderer_destroy(self.handle)
            self.handle = C.c_void_p()
        _check(lib().rt_renderer_create(self.scene.handle, C.byref(self.params), C.byref(self.handle)))
        self._initialised = True
        return self

    def _need(self):
        if not self._initialised:
            self.Init()

    def close(self):
        if getattr(self, "handle", None):
            lib().rt_renderer_destroy(self.handle)
            self.handle = None
            self._initialised = False

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def ClearAccumulator(self):
        self._need()
        _check(lib().rt_renderer_clear(self.handle))

    def set_stream(self, cuda_stream_ptr):
        self._need()
        _check(lib().rt_renderer_set_stream(self.handle, cuda_stream_ptr))

    def set_accumulator(self, device_ptr):
        self._need()
        _check(lib().rt_renderer_set_accumulator(self.handle, device_ptr))

    def Tick(self, deltaTime=0.0):
        """One frame.  Path tracer: `passes` samples per pixel with the current `spp`, then spp += passes
        (renderer.cpp:144-168).  Whitted: overwrites the accumulator (renderer.cpp:131-157)."""
        self.render(1)

    def render(self, frames, first_spp=None, stride=None):
        """`frames` Ticks in one call (frames run concurrently on the GPU; result = running them in order).
        Every Tick takes `passes` samples per pixel and advances spp by `passes` (renderer.cpp:123,167), so the
        default stride between the frames' spp counters is `passes`."""
        self._need()
        _check(lib().rt_renderer_set_camera(self.handle, C.byref(self.camera.c)))
        if self.integrator == abi.RT_INTEGRATOR_PATH:
            _check(lib().rt_renderer_set_passes(self.handle, int(self.passes)))
        first = self.spp if first_spp is None else first_spp
        _check(lib().rt_renderer_render(self.handle, first, frames, self.passes if stride is None else stride))
        if self.integrator == abi.RT_INTEGRATOR_PATH and first_spp is None:
            self.spp += frames * self.passes

    def sync(self):
        self._need()
        _check(lib().rt_renderer_sync(self.handle))

    @property
    def accumulator(self):
        self._need()
        out = np.empty((self.height, self.width, 4), np.float32)
        _check(lib().rt_renderer_read_accumulator(self.handle, out.ctypes.data))
        return out

    def device_accumulator(self):
        self._need()
        return lib().rt_renderer_device_accumulator(self.handle)

    def read_accumulator_into(self, host_ptr):
        """rt_renderer_read_accumulator into a caller-owned host buffer of width * height float4 (e.g. pinned memory)"""
        self._need()
        _check(lib().rt_renderer_read_accumulator(self.handle, host_ptr))

    def export_accumulator(self):
        """CUDA IPC handle (bytes) of this renderer's accumulator, for the tile shards of other processes"""
        self._need()
        buf = C.create_string_buffer(abi.RT_IPC_HANDLE_BYTES)
        _check(lib().rt_renderer_export_accumulator(self.handle, buf))
        return buf.raw

    def import_accumulator(self, handle_bytes):
        """accumulate into another process' accumulator (peer-mapped over NVLink) from now on"""
        self._need()
        buf = C.create_string_buffer(bytes(handle_bytes), abi.RT_IPC_HANDLE_BYTES)
        _check(lib().rt_renderer_import_accumulator(self.handle, buf))

    def screen_pixels(self, scale=None):
        """screen->pixels as the reference displays them: accumulator * 1/(spp+passes) through RGBF32_to_RGB8
        (renderer.cpp:119,127-129); note that `spp` here is the value BEFORE the last Tick's increment."""
        self._need()
        if scale is None:
            scale = 1.0 if self.integrator == abi.RT_INTEGRATOR_WHITTED else 1.0 / float(self.spp)
        out = np.empty((self.height, self.width), np.uint32)
        _check(lib().rt_renderer_read_pixels(self.handle, scale, out.ctypes.data))
        return out

    def counters(self):
        self._need()
        c = abi.rt_counters()
        _check(lib().rt_renderer_get_counters(self.handle, C.byref(c)))
        return {k: int(getattr(c, k)) for k, _ in abi.rt_counters._fields_}

    def reset_counters(self):
        self._need()
        _check(lib().rt_renderer_reset_counters(self.handle))

    def set_profiling(self, on):
        self._need()
        _check(lib().rt_renderer_set_profiling(self.handle, 1 if on else 0))

    def launch_spans(self, capacity=1 << 16):
        """(stage ids, ms) of every launch since profiling was switched on / last read, in launch order"""
        self._need()
        stage = np.zeros(capacity, np.int32)
        ms = np.zeros(capacity, np.float32)
        n = C.c_size_t()
        _check(lib().rt_renderer_get_launch_spans(self.handle, stage.ctypes.data, ms.ctypes.data, capacity, C.byref(n)))
        return stage[:n.value], ms[:n.value]

    def queue_history(self, capacity=1 << 15):
        self._need()
        out = np.zeros(capacity, np.int32)
        n = C.c_size_t()
        _check(lib().rt_renderer_get_queue_history(self.handle, out.ctypes.data, capacity, C.byref(n)))
        return out[:n.value]

    def stage_times(self):
        """{stage: (device ms, launches)} since the last call (needs set_profiling(True))"""
        self._need()
        t = abi.rt_stage_times()
        _check(lib().rt_renderer_get_stage_times(self.handle, C.byref(t)))
        return {name: (float(t.ms[i]), int(t.launches[i])) for i, name in enumerate(abi.STAGES)}


class MultiGpuRenderer:
    """Renderer::Tick (3. PathTracer/renderer.cpp:144-168) on several GPUs of this process: rt_multi_renderer.  The scene is
    replicated, device k renders the interleaved tiles k, k + n, ... and every device accumulates into ONE image on devices[0]
    through peer-mapped memory; the accumulator is bit-identical to a one-GPU GpuRenderer's."""

    def __init__(self, flat, integrator, width, height, devices, depthLimit=5, seed_mode=abi.RT_SEED_REFERENCE_TILE,
                 schedule=abi.RT_SCHEDULE_AUTO, lookahead_frames=0, counters=False):
        if isinstance(flat, (str, os.PathLike)):
            flat = FlatScene.load(flat)
        self.flat, self.integrator, self.width, self.height = flat, integrator, width, height
        self.devices = [int(d) for d in devices]
        self.camera = Camera(width, height)
        self.spp, self.passes = 1, 1
        self.params = abi.rt_render_params()
        lib().rt_render_params_default(C.byref(self.params), integrator, width, height)
        self.params.depth_limit, self.params.seed_mode = depthLimit, seed_mode
        self.params.schedule, self.params.lookahead_frames = schedule, lookahead_frames
        self.handle = C.c_void_p()
        desc = flat.desc()
        dev = (C.c_int * len(self.devices))(*self.devices)
        _check(lib().rt_multi_renderer_create(C.byref(desc), abi.RT_SCENE_FLAG_COUNTERS if counters else 0, dev, len(self.devices),
                                              C.byref(self.params), C.byref(self.handle)))

    def close(self):
        if getattr(self, "handle", None):
            lib().rt_multi_renderer_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def ClearAccumulator(self):
        _check(lib().rt_multi_renderer_clear(self.handle))

    def Tick(self, deltaTime=0.0):
        self.render(1)

    def render(self, frames, first_spp=None, stride=None):
        _check(lib().rt_multi_renderer_set_camera(self.handle, C.byref(self.camera.c)))
        _check(lib().rt_multi_renderer_set_passes(self.handle, int(self.passes)))
        first = self.spp if first_spp is None else first_spp
        _check(lib().rt_multi_renderer_render(self.handle, first, frames, self.passes if stride is None else stride))
        if first_spp is None:
            self.spp += frames * self.passes

    def sync(self):
        _check(lib().rt_multi_renderer_sync(self.handle))

    @property
    def accumulator(self):
        out = np.empty((self.height, self.width, 4), np.float32)
        _check(lib().rt_multi_renderer_read_accumulator(self.handle, out.ctypes.data))
        return out

    def read_accumulator_into(self, host_ptr):
        _check(lib().rt_multi_renderer_read_accumulator(self.handle, host_ptr))

    def counters(self):
        c = abi.rt_counters()
        _check(lib().rt_multi_renderer_get_counters(self.handle, C.byref(c)))
        return {k: int(getattr(c, k)) for k, _ in abi.rt_counters._fields_}

    def reset_counters(self):
        _check(lib().rt_multi_renderer_reset_counters(self.handle))
