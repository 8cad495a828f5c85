"""One Whitted frame without the CUDA graph (for ncu captures of the generate / extend / shade / connect stages):
whitted_once.py scene [W H]   (development tool)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["RT_B200_WHITTED_GRAPH"] = "0"
import cpu_ray_tracer_b200 as rtb
from cpu_ray_tracer_b200 import abi, api
name = sys.argv[1] if len(sys.argv) > 1 else "bunny_flat"
W, H = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (640, 360)
sc = api.open_scene(rtb.FlatScene.load(os.path.join(ROOT, "oracle", "_ref", "scenes", name + ".rtscene.gz")))
r = api.GpuRenderer(sc, abi.RT_INTEGRATOR_WHITTED, W, H).Init()
r.Tick(0); r.sync()
r.reset_counters()
r.Tick(0); r.sync()
print(name, W, H, r.counters())
