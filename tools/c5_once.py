"""One ray set of the 10 M-triangle microbench without timing (for ncu captures): c5_once.py [scattered|primary|bounce|shadow] [launches]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
import ray_bench
from cpu_ray_tracer_b200 import abi, api, host_build
which = sys.argv[1] if len(sys.argv) > 1 else "scattered"
launches = int(sys.argv[2]) if len(sys.argv) > 2 else 3
tris = host_build.terrain_mesh(10_000_000, seed=1)
fs = host_build.flat_scene_from_tris(tris, builder=lambda t: (np.zeros(0, abi.NODE_DTYPE), np.zeros(0, np.uint32), 0.0))
fs.device_build = True
sc = api.open_scene(fs)
W = H = 4096
cam = api.Camera(W, H)
cam.SetCameraState((0.0, 6.0, -4.0), (0.0, -0.5, 6.0))
rays = ray_bench.primary(W, H, cam)
if which == "scattered":
    rng = np.random.default_rng(7)
    n = 1 << 24
    lo, hi = tris["v0"].min(0), tris["v0"].max(0)
    O = np.stack([rng.uniform(lo[0], hi[0], n), np.full(n, hi[1] + 0.5), rng.uniform(lo[2], hi[2], n)], 1).astype(np.float32)
    D = rng.normal(size=(n, 3)).astype(np.float32)
    D[:, 1] = -np.abs(D[:, 1]) - 1.0
    D /= np.linalg.norm(D, axis=1, keepdims=True).astype(np.float32)
    r = api.make_rays(O, D)
elif which == "primary":
    r = rays
else:
    hits = sc.FindNearest(rays)
    r = ray_bench.bounce_rays(fs, rays, hits) if which == "bounce" else ray_bench.shadow_rays(fs, rays, hits)
d_rays = torch.from_numpy(r.view(np.uint8).reshape(-1, 32)).cuda()
m = len(r)
stream = torch.cuda.Stream()
if which == "shadow":
    res = torch.empty(m, dtype=torch.uint8, device="cuda")
    fn = lambda: sc.IsOccludedDevice(d_rays.data_ptr(), res.data_ptr(), m, stream.cuda_stream)
else:
    res = torch.empty((m, 32), dtype=torch.uint8, device="cuda")
    fn = lambda: sc.FindNearestDevice(d_rays.data_ptr(), res.data_ptr(), m, stream.cuda_stream)
for _ in range(launches):
    fn()
    stream.synchronize()
print(which, m, "rays x", launches)
