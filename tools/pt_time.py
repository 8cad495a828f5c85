"""Times one path-traced job with CUDA-side sync (development tool).
usage: pt_time.py [scene[,scene...]] [spp] [W H]   env: RT_B200_STREAM_KERNEL, RT_B200_PT_SCHEDULE, RT_B200_STREAM_CTAS"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cpu_ray_tracer_b200 as rtb
from cpu_ray_tracer_b200 import abi, api
names = (sys.argv[1] if len(sys.argv) > 1 else "wok_teapot_flat").split(",")
spps = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "64").split(",")]
W, H = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (1920, 1080)
tag = f"kernel={os.environ.get('RT_B200_STREAM_KERNEL', 'default')} sched={os.environ.get('RT_B200_PT_SCHEDULE', 'streams')} ctas={os.environ.get('RT_B200_STREAM_CTAS', 'occ')}"
for name in names:
    fs = rtb.FlatScene.load(os.path.join(ROOT, "oracle", "_ref", "scenes", name + ".rtscene.gz"))
    sc = api.open_scene(fs)
    r = api.GpuRenderer(sc, abi.RT_INTEGRATOR_PATH, W, H, seed_mode=int(os.environ.get("RT_SEED_MODE", "0"))).Init()
    for spp in spps:
        r.render(spp, first_spp=1); r.sync()
        best = 1e9
        for _ in range(3):
            r.reset_counters()
            t0 = time.perf_counter(); r.render(spp, first_spp=1); r.sync(); dt = time.perf_counter() - t0
            best = min(best, dt)
        c = r.counters()
        print(f"{name:18s} {W}x{H} {spp:4d}spp [{tag}]: {best*1e3:8.2f} ms  {c['extension_rays']/best/1e6:8.1f} Mrays/s  rays {c['extension_rays']}", flush=True)
    r.close(); sc.close()
