"""Renderer::Tick rate (one frame per call, the reference's own usage) with and without look-ahead frames.
usage: tick_time.py [scene] [ticks] [W H]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cpu_ray_tracer_b200 as rtb
from cpu_ray_tracer_b200 import abi, api
name = sys.argv[1] if len(sys.argv) > 1 else "wok_teapot_flat"
ticks = int(sys.argv[2]) if len(sys.argv) > 2 else 64
W, H = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (1920, 1080)
sc = api.open_scene(rtb.FlatScene.load(os.path.join(ROOT, "oracle", "_ref", "scenes", name + ".rtscene.gz")))
for L in (0, 16, 64):
    r = api.GpuRenderer(sc, abi.RT_INTEGRATOR_PATH, W, H, lookahead_frames=L).Init()
    r.Tick(0); r.sync()           # first call: pilot tile order (+ the first batch of frames)
    r.ClearAccumulator(); r.spp = 1
    per = []
    t0 = time.perf_counter()
    for _ in range(ticks):
        t = time.perf_counter(); r.Tick(0); r.sync(); per.append(time.perf_counter() - t)
    total = time.perf_counter() - t0
    px = r.screen_pixels()         # what the reference blits after a Tick
    print(json.dumps({"scene": name, "lookahead_frames": L, "ticks": ticks, "ms_per_tick_mean": round(1e3 * total / ticks, 3),
                      "ms_per_tick_max": round(1e3 * max(per), 2), "ms_per_tick_median": round(1e3 * sorted(per)[len(per) // 2], 3),
                      "ticks_per_s": round(ticks / total, 1)}), flush=True)
    r.close()
