"""Latency of single (tile x 64 frames) stream sets: how long is the critical path? (development tool)"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cpu_ray_tracer_b200 as rtb
from cpu_ray_tracer_b200 import abi, api
name = sys.argv[1] if len(sys.argv) > 1 else "wok_teapot_flat"
spp = 64
W, H = 1920, 1080
fs = rtb.FlatScene.load(os.path.join(ROOT, "oracle", "_ref", "scenes", name + ".rtscene.gz"))
sc = api.open_scene(fs)
tilesX = W // 16
res = []
for ty in range(2, 67, 4):
    for tx in range(4, 120, 8):
        t = ty * tilesX + tx
        r = api.GpuRenderer(sc, abi.RT_INTEGRATOR_PATH, W, H, tile_begin=t, tile_end=t + 1).Init()
        r.render(spp, first_spp=1); r.sync(); r.reset_counters()
        t0 = time.perf_counter(); r.render(spp, first_spp=1); r.sync(); dt = time.perf_counter() - t0
        c = r.counters()["extension_rays"]
        res.append((dt * 1e3, c / spp, tx, ty))
        r.close()
res.sort(reverse=True)
for ms, rays, tx, ty in res[:12]:
    print(f"tile ({tx:3d},{ty:2d}): {ms:7.2f} ms for 64 streams, {rays:7.1f} rays/stream, {ms*1e3/rays:6.2f} us per ray of the chain")
print("...")
for ms, rays, tx, ty in res[-3:]:
    print(f"tile ({tx:3d},{ty:2d}): {ms:7.2f} ms for 64 streams, {rays:7.1f} rays/stream, {ms*1e3/rays:6.2f} us per ray of the chain")
