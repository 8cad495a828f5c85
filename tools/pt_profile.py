"""Per-iteration profile of the path-tracer wavefront (development tool): queue width and stage time per iteration."""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cpu_ray_tracer_b200 as rtb
from cpu_ray_tracer_b200 import abi, api

name = sys.argv[1] if len(sys.argv) > 1 else "wok_teapot_flat"
W, H, spp = 1920, 1080, int(sys.argv[2]) if len(sys.argv) > 2 else 64
fs = rtb.FlatScene.load(os.path.join(ROOT, "oracle", "_ref", "scenes", name + ".rtscene.gz"))
sc = api.open_scene(fs)
r = api.GpuRenderer(sc, abi.RT_INTEGRATOR_PATH, W, H, schedule=abi.RT_SCHEDULE_WAVEFRONT).Init()
r.render(spp, first_spp=1); r.sync()
r.set_profiling(True)
r.reset_counters()
r.render(spp, first_spp=1)
stage, ms = r.launch_spans()
hist = r.queue_history()
c = r.counters()
ext = ms[stage == 1]; sh = ms[stage == 2]
n = min(len(ext), len(hist))
print("iterations", len(hist), "rays", c["extension_rays"], "extend ms", ext.sum(), "shade ms", sh.sum())
edges = [0, 16, 64, 128, 256, 384, 512, 768, 1024, 1280, 1536, 100000]
for a, b in zip(edges[:-1], edges[1:]):
    b = min(b, n)
    if a >= b: break
    h = hist[a:b].astype(np.float64)
    print(f"iter {a:5d}-{b:5d}: rays/iter {h.mean():10.0f}  extend us/iter {1000*ext[a:b].mean():8.1f}  shade us/iter {1000*sh[a:b].mean():8.1f}"
          f"  extend Mrays/s {h.sum()/ext[a:b].sum()/1e3:8.0f}  cum extend ms {ext[:b].sum():7.1f}")
np.savez(os.path.join(ROOT, "gpurun_out", f"pt_profile_{name}.npz"), hist=hist, ext=ext, sh=sh)
