"""ctypes harness over oracle/_ref/libref_*.so — the reference's own CPU code (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's reference / cpu_baseline arm may import this.
One process can hold ONE reference library instance (the reference keeps its renderer in globals and
chdir()s into the asset tree), so callers that need several scenes run this module as a subprocess:

    python -m oracle.refhost dump  <integrator> <kind> <scene.xml> <W> <H> <out.npz> [frames] [cam...]
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
WORK = os.path.join(REF_DIR, "work")

f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")


def lib_path(integrator, kind, fast=False):
    """kind: file | tlas | file_kd (FileScene with the KD-tree it ships with)"""
    return os.path.join(REF_DIR, f"libref_{integrator}_{kind}{'_fast' if fast else ''}.so")


def available(integrator="pt", kind="file", fast=False):
    return os.path.exists(lib_path(integrator, kind, fast)) and os.path.isdir(os.path.join(WORK, "assets"))


class RefRenderer:
    """The reference's Renderer + scene, headless (see oracle/ref_build/ref_api.cpp)."""

    def __init__(self, integrator, kind, scene_xml, width, height, fast=False):
        self.lib = C.CDLL(lib_path(integrator, kind, fast))
        L = self.lib
        L.ref_create.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int]
        L.ref_set_camera.argtypes = [f32p, f32p]
        L.ref_get_camera.argtypes = [f32p]
        L.ref_find_nearest.argtypes = [C.c_int, f32p, f32p, f32p, f32p, f32p, f32p, i32p, i32p, i32p, i32p]
        L.ref_is_occluded.argtypes = [C.c_int, f32p, f32p, f32p, u8p]
        L.ref_primary_hits.argtypes = [f32p, f32p, f32p, f32p, f32p, i32p, i32p, i32p, i32p]
        L.ref_hit_info.argtypes = [C.c_int, f32p, f32p, f32p, f32p, f32p, i32p, i32p, f32p, f32p, f32p]
        L.ref_tick.argtypes = [C.c_int]
        L.ref_last_tick_seconds.restype = C.c_double
        L.ref_accumulator.restype = C.POINTER(C.c_float)
        L.ref_screen.restype = C.POINTER(C.c_uint32)
        L.ref_reset.argtypes = [C.c_int]
        L.ref_set_passes.argtypes = [C.c_int]
        L.ref_set_depth_limit.argtypes = [C.c_int]
        L.ref_energy.restype = C.c_float
        L.ref_flatten.argtypes = [C.c_char_p]
        L.ref_refit.argtypes = [C.c_int, C.c_int, f32p]
        self.width, self.height = width, height
        # scene files are addressed the way the reference does: "../assets/scenes/x.xml" from work/run
        if not scene_xml.startswith("../") and not os.path.isabs(scene_xml):
            scene_xml = "../assets/scenes/" + scene_xml
        rc = L.ref_create(scene_xml.encode(), os.path.join(WORK, "run").encode(), width, height)
        if rc != 0:
            raise RuntimeError(f"ref_create({scene_xml}) failed: {rc}")

    def threads(self):
        return self.lib.ref_threads()

    def set_camera(self, pos, target):
        self.lib.ref_set_camera(np.asarray(pos, np.float32), np.asarray(target, np.float32))

    def get_camera(self):
        out = np.zeros(12, np.float32)
        self.lib.ref_get_camera(out)
        return out.reshape(4, 3)

    def set_depth_limit(self, d):
        self.lib.ref_set_depth_limit(d)

    def find_nearest(self, O, D, tmax=None):
        O = np.ascontiguousarray(O, np.float32)
        D = np.ascontiguousarray(D, np.float32)
        n = O.shape[0]
        tmax = np.full(n, 1e34, np.float32) if tmax is None else np.ascontiguousarray(tmax, np.float32)
        out = dict(t=np.zeros(n, np.float32), u=np.zeros(n, np.float32), v=np.zeros(n, np.float32),
                   obj=np.zeros(n, np.int32), tri=np.zeros(n, np.int32),
                   traversed=np.zeros(n, np.int32), tested=np.zeros(n, np.int32))
        self.lib.ref_find_nearest(n, O, D, tmax, out["t"], out["u"], out["v"], out["obj"], out["tri"],
                                  out["traversed"], out["tested"])
        return out

    def is_occluded(self, O, D, tmax):
        O = np.ascontiguousarray(O, np.float32)
        D = np.ascontiguousarray(D, np.float32)
        tmax = np.ascontiguousarray(tmax, np.float32)
        out = np.zeros(O.shape[0], np.uint8)
        self.lib.ref_is_occluded(O.shape[0], O, D, tmax, out)
        return out

    def primary_hits(self):
        n = self.width * self.height
        out = dict(O=np.zeros((n, 3), np.float32), D=np.zeros((n, 3), np.float32),
                   t=np.zeros(n, np.float32), u=np.zeros(n, np.float32), v=np.zeros(n, np.float32),
                   obj=np.zeros(n, np.int32), tri=np.zeros(n, np.int32),
                   traversed=np.zeros(n, np.int32), tested=np.zeros(n, np.int32))
        self.lib.ref_primary_hits(out["O"], out["D"], out["t"], out["u"], out["v"], out["obj"], out["tri"],
                                  out["traversed"], out["tested"])
        return out

    def hit_info(self, O, D, hits):
        n = O.shape[0]
        N = np.zeros((n, 3), np.float32)
        uv = np.zeros((n, 2), np.float32)
        albedo = np.zeros((n, 3), np.float32)
        self.lib.ref_hit_info(n, np.ascontiguousarray(O, np.float32), np.ascontiguousarray(D, np.float32),
                              hits["t"], hits["u"], hits["v"], hits["obj"], hits["tri"], N, uv, albedo)
        return N, uv, albedo

    def tick(self, frames=1):
        self.lib.ref_tick(frames)
        return self.lib.ref_last_tick_seconds()

    def reset(self, spp=1):
        self.lib.ref_reset(spp)

    def set_passes(self, p):
        self.lib.ref_set_passes(p)

    def spp(self):
        return self.lib.ref_spp()

    def accumulator(self):
        n = self.width * self.height * 4
        return np.ctypeslib.as_array(self.lib.ref_accumulator(), (n,)).reshape(self.height, self.width, 4).copy()

    def screen(self):
        n = self.width * self.height
        return np.ctypeslib.as_array(self.lib.ref_screen(), (n,)).reshape(self.height, self.width).copy()

    def refit(self, blas, verts9):
        """overwrite the triangle vertices of one acceleration structure and call the reference's Refit() on it"""
        v = np.ascontiguousarray(verts9, np.float32).reshape(-1, 9)
        rc = self.lib.ref_refit(blas, len(v), v)
        if rc != 0:
            raise RuntimeError(f"ref_refit failed: {rc}")

    def flatten(self, path):
        rc = self.lib.ref_flatten(os.path.abspath(path).encode())
        if rc != 0:
            raise RuntimeError("ref_flatten failed")


def _main(argv):
    cmd = argv[0]
    if cmd == "dump":
        integrator, kind, scene, W, H, out = argv[1], argv[2], argv[3], int(argv[4]), int(argv[5]), argv[6]
        frames = int(argv[7]) if len(argv) > 7 else 1
        out = os.path.abspath(out)
        r = RefRenderer(integrator, kind, scene, W, H)
        if len(argv) > 8:
            cam = [float(x) for x in argv[8:14]]
            r.set_camera(cam[:3], cam[3:])
        prim = r.primary_hits()
        r.set_passes(int(os.environ.get("RT_REF_PASSES", "1")))  # Renderer::passes (path tracer only)
        secs = r.tick(frames)
        np.savez_compressed(out, accumulator=r.accumulator(), camera=r.get_camera(), seconds=secs,
                            threads=r.threads(), **{"prim_" + k: v for k, v in prim.items()})
    elif cmd == "bench":
        # one warm-up frame, then `frames` timed frames (Renderer::Tick), JSON on stdout
        import json
        integrator, kind, scene, W, H, frames = argv[1], argv[2], argv[3], int(argv[4]), int(argv[5]), int(argv[6])
        fast = len(argv) > 7 and argv[7] == "1"
        r = RefRenderer(integrator, kind, scene, W, H, fast=fast)
        r.tick(1)
        secs = r.tick(frames)
        print(json.dumps({"seconds": secs, "threads": r.threads(), "frames": frames, "fast": fast}))
    elif cmd == "bench_steps":
        # warmup x frames untimed, then steps x frames timed; JSON on stdout
        import json
        integrator, kind, scene, W, H, frames, warmup, steps = argv[1], argv[2], argv[3], int(argv[4]), int(argv[5]), int(argv[6]), int(argv[7]), int(argv[8])
        fast = len(argv) > 9 and argv[9] == "1"
        r = RefRenderer(integrator, kind, scene, W, H, fast=fast)
        for _ in range(warmup):
            r.tick(frames)
        secs = 0.0
        for _ in range(steps):
            secs += r.tick(frames)
        print(json.dumps({"seconds": secs, "threads": r.threads(), "frames": frames, "fast": fast,
                          "flags": "reference Renderer::Tick built headless with g++ " + ("-O3 -mavx2 -mfma -ffast-math" if fast else "-O2 -fopenmp -ffp-contract=off")}))
    elif cmd == "refit_flatten":
        # refit_flatten <integrator> <kind> <scene.xml> <verts.npy: (n, 9) float32> <blas> <out.rtscene>
        integrator, kind, scene, verts, blas, out = argv[1], argv[2], argv[3], argv[4], int(argv[5]), os.path.abspath(argv[6])
        v = np.load(os.path.abspath(verts))
        r = RefRenderer(integrator, kind, scene, 64, 64)
        r.refit(blas, v)
        r.flatten(out)
    elif cmd == "flatten":
        integrator, kind, scene, out = argv[1], argv[2], argv[3], os.path.abspath(argv[4])
        r = RefRenderer(integrator, kind, scene, 64, 64)
        r.flatten(out)
    else:
        raise SystemExit("unknown command " + cmd)


if __name__ == "__main__":
    _main(sys.argv[1:])
