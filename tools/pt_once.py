"""Path-traced jobs without timing (for ncu captures): pt_once.py scene spp [W H [repeats]]
(the second job of a view uses the measured tile order: capture it with ncu -s 1 -c 1)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cpu_ray_tracer_b200 as rtb
from cpu_ray_tracer_b200 import abi, api
name = sys.argv[1] if len(sys.argv) > 1 else "wok_teapot_flat"
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 64
W, H = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (1920, 1080)
sc = api.open_scene(rtb.FlatScene.load(os.path.join(ROOT, "oracle", "_ref", "scenes", name + ".rtscene.gz")))
r = api.GpuRenderer(sc, abi.RT_INTEGRATOR_PATH, W, H).Init()
for _ in range(int(sys.argv[5]) if len(sys.argv) > 5 else 1):
    r.render(spp, first_spp=1); r.sync()
print(name, spp, r.counters()["extension_rays"])
