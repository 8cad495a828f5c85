"""Drop-in check of the C++ adapters (TEST INFRASTRUCTURE ONLY; run as a subprocess by tests/test_cpp_adapters.py).

    python -m oracle.gpuhost_check <whitted|pt> <file|tlas> <scene.xml> <W> <H> <frames> [devices, e.g. 0,1]

Loads oracle/_ref/libgpuhost_<integrator>_<kind>.so, which holds BOTH the reference's own Renderer
(ref_* entry points, CPU) and rtb200::GpuRenderer (gh_* entry points: the reference's loaders and
builders + cpu-ray-tracer_b200/host/rt_b200_adapters.h + librt_b200.so), constructs both from the same
scene XML and compares Tick output, FindNearest and IsOccluded.  Prints one JSON line.
"""
import ctypes as C
import json
import os
import sys

import numpy as np

from oracle import refhost

f32p, i32p, u8p = refhost.f32p, refhost.i32p, refhost.u8p


def lib_path(integrator, kind):
    return os.path.join(refhost.REF_DIR, f"libgpuhost_{integrator}_{kind}.so")


def available(integrator, kind):
    return os.path.exists(lib_path(integrator, kind)) and os.path.isdir(os.path.join(refhost.WORK, "assets"))


def main(integrator, kind, xml, W, H, frames, devices=""):
    W, H, frames = int(W), int(H), int(frames)
    devices = [int(d) for d in devices.split(",") if d != ""]
    L = C.CDLL(lib_path(integrator, kind))
    L.gh_last_error.restype = C.c_char_p
    L.gh_create.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int]
    L.gh_accumulator.restype = C.POINTER(C.c_float)
    L.gh_screen.restype = C.POINTER(C.c_uint32)
    L.gh_find_nearest.argtypes = [C.c_int, f32p, f32p, f32p, C.c_int, f32p, f32p, f32p, i32p, i32p]
    L.gh_is_occluded.argtypes = [C.c_int, f32p, f32p, f32p, u8p]
    L.gh_set_camera.argtypes = [f32p, f32p]
    L.ref_create.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int]
    L.ref_accumulator.restype = C.POINTER(C.c_float)
    L.ref_screen.restype = C.POINTER(C.c_uint32)
    L.ref_find_nearest.argtypes = [C.c_int, f32p, f32p, f32p, f32p, f32p, f32p, i32p, i32p, i32p, i32p]
    L.ref_is_occluded.argtypes = [C.c_int, f32p, f32p, f32p, u8p]
    L.ref_primary_hits.argtypes = [f32p, f32p, f32p, f32p, f32p, i32p, i32p, i32p, i32p]
    L.ref_set_camera.argtypes = [f32p, f32p]
    path = ("../assets/scenes/" + xml).encode()
    run = os.path.join(refhost.WORK, "run").encode()
    if L.ref_create(path, run, W, H) != 0:
        raise SystemExit("ref_create failed")
    if len(devices) > 1:
        # GpuRenderer over several GPUs of this process (tile jobs dealt to the devices, one peer-mapped accumulator)
        L.gh_create_multi.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int]
        rc = L.gh_create_multi(path, run, W, H, (C.c_int * len(devices))(*devices), len(devices))
    else:
        rc = L.gh_create(path, run, W, H, devices[0] if devices else 0)
    if rc != 0:
        raise SystemExit("gh_create failed: " + L.gh_last_error().decode())
    out = {"triangles": L.gh_triangle_count()}
    info = (C.c_ulonglong * 6)()
    if L.gh_scene_info(info) != 0:
        raise SystemExit(L.gh_last_error().decode())
    out["scene_info"] = dict(zip(("instances", "meshes", "fat_nodes", "triangle_slots", "bytes_geometry", "multi_devices"), (int(x) for x in info)))
    n = W * H
    res = {}
    for cam in (None, ((1.6, 0.9, -1.4), (0.0, -0.4, 1.0))):
        if cam is not None:
            p, t = np.asarray(cam[0], np.float32), np.asarray(cam[1], np.float32)
            L.ref_set_camera(p, t), L.gh_set_camera(p, t)
            L.ref_reset(1)
            if L.gh_reset(1) != 0:
                raise SystemExit(L.gh_last_error().decode())
            if integrator == "pt":  # second camera: two samples per pixel per Tick (the UI's "spp" slider, renderer.cpp:182)
                L.ref_set_passes(2), L.gh_set_passes(2)
        tag = "cam0" if cam is None else "cam1"
        # Renderer::Tick x frames on both sides
        L.ref_tick(frames)
        if L.gh_tick(frames) != 0:
            raise SystemExit("gh_tick failed: " + L.gh_last_error().decode())
        ra = np.ctypeslib.as_array(L.ref_accumulator(), (n * 4,)).reshape(H, W, 4).copy()
        ga = np.ctypeslib.as_array(L.gh_accumulator(), (n * 4,)).reshape(H, W, 4).copy()
        rs = np.ctypeslib.as_array(L.ref_screen(), (n,)).copy()
        gs = np.ctypeslib.as_array(L.gh_screen(), (n,)).copy()
        d = np.nan_to_num(np.abs(ga.astype(np.float64) - ra))
        chan = lambda p: np.stack([(p >> 16) & 255, (p >> 8) & 255, p & 255], -1).astype(np.int32)
        res[tag] = {"max_abs": float(d.max()), "rmse": float(np.sqrt((d ** 2).mean())),
                    "pixels_over_1e-4": int((d.max(-1) > 1e-4 * frames).sum()),
                    "bit_identical_fraction": float((ga.view(np.uint32) == ra.view(np.uint32)).mean()),
                    "screen_max_channel_diff": int(np.abs(chan(gs) - chan(rs)).max()),
                    "screen_pixels_differing": int((gs != rs).sum()),
                    "ref_spp": int(L.ref_spp()), "gpu_spp": int(L.gh_spp()), "mean": float(ra[..., :3].mean())}
    out["tick"] = res
    # BaseScene::FindNearest / IsOccluded: primary rays of the reference camera
    prim = dict(O=np.zeros((n, 3), np.float32), D=np.zeros((n, 3), np.float32), t=np.zeros(n, np.float32),
                u=np.zeros(n, np.float32), v=np.zeros(n, np.float32), obj=np.zeros(n, np.int32), tri=np.zeros(n, np.int32),
                traversed=np.zeros(n, np.int32), tested=np.zeros(n, np.int32))
    L.ref_primary_hits(prim["O"], prim["D"], prim["t"], prim["u"], prim["v"], prim["obj"], prim["tri"], prim["traversed"], prim["tested"])
    tmax = np.full(n, 1e34, np.float32)
    g = dict(t=np.zeros(n, np.float32), u=np.zeros(n, np.float32), v=np.zeros(n, np.float32), obj=np.zeros(n, np.int32), tri=np.zeros(n, np.int32))
    if L.gh_find_nearest(n, prim["O"], prim["D"], tmax, 0, g["t"], g["u"], g["v"], g["obj"], g["tri"]) != 0:
        raise SystemExit(L.gh_last_error().decode())
    out["find_nearest_mismatches"] = {k: int((g[k].view(np.uint32) != prim[k].view(np.uint32)).sum()) for k in g}
    out["hit_fraction"] = float((prim["obj"] >= 2).mean())
    # single-ray virtual calls (the UI pick path): the first 64 rays of a row that hits geometry
    k = 64
    row = int(np.argmax((prim["obj"] >= 2).reshape(H, W).sum(1))) * W
    sl = slice(row, row + k)
    s = dict(t=np.zeros(k, np.float32), u=np.zeros(k, np.float32), v=np.zeros(k, np.float32), obj=np.zeros(k, np.int32), tri=np.zeros(k, np.int32))
    L.gh_find_nearest(k, prim["O"][sl].copy(), prim["D"][sl].copy(), tmax[:k], 1, s["t"], s["u"], s["v"], s["obj"], s["tri"])
    out["single_ray_mismatches"] = int(sum((s[f].view(np.uint32) != prim[f][sl].view(np.uint32)).sum() for f in s))
    # shadow rays toward (0, 3, 1) from the hit points
    m = prim["obj"] >= 0
    I = prim["O"][m] + prim["t"][m, None] * prim["D"][m]
    Lv = np.array([0.0, 3.0, 1.0], np.float32)[None, :] - I
    dist = np.sqrt((Lv * Lv).sum(1)).astype(np.float32)
    Lv = (Lv / dist[:, None]).astype(np.float32)
    so = np.ascontiguousarray(I + Lv * np.float32(0.001), np.float32)
    st = np.ascontiguousarray(dist - np.float32(0.002), np.float32)
    ro, go = np.zeros(len(so), np.uint8), np.zeros(len(so), np.uint8)
    L.ref_is_occluded(len(so), so, Lv, st, ro)
    if L.gh_is_occluded(len(so), so, Lv, st, go) != 0:
        raise SystemExit(L.gh_last_error().decode())
    out["occlusion_mismatches"] = int((ro != go).sum())
    out["occluded_fraction"] = float(ro.mean())
    L.gh_destroy()
    print(json.dumps(out))


if __name__ == "__main__":
    main(*sys.argv[1:8])
