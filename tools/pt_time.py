"""Times one path-traced job per schedule with CUDA-side sync (development tool)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cpu_ray_tracer_b200 as rtb
from cpu_ray_tracer_b200 import abi, api
name = sys.argv[1] if len(sys.argv) > 1 else "wok_teapot_flat"
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 64
W, H = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (1920, 1080)
fs = rtb.FlatScene.load(os.path.join(ROOT, "oracle", "_ref", "scenes", name + ".rtscene.gz"))
sc = api.open_scene(fs)
for sched, nm in ((abi.RT_SCHEDULE_STREAMS, "streams"), (abi.RT_SCHEDULE_WAVEFRONT, "wavefront")):
    r = api.GpuRenderer(sc, abi.RT_INTEGRATOR_PATH, W, H, schedule=sched).Init()
    r.render(spp, first_spp=1); r.sync()
    best = 1e9
    for _ in range(3):
        r.reset_counters()
        t0 = time.perf_counter(); r.render(spp, first_spp=1); r.sync(); dt = time.perf_counter() - t0
        best = min(best, dt)
    c = r.counters()
    print(f"{name} {W}x{H} {spp}spp {nm:10s}: {best*1e3:8.2f} ms  {c['extension_rays']/best/1e6:8.1f} Mrays/s  rays {c['extension_rays']}")
    r.close()
