"""How many accumulator floats of the CUDA integrators equal the oracle's bit for bit, per scene / schedule / call pattern
(development probe behind tests/test_glibc_math.py; run on the GPU box, writes gpurun_out/radiance_exact_probe.jsonl)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import cpu_ray_tracer_b200 as rtb
from cpu_ray_tracer_b200 import abi, api
from oracle import porthost
from conftest import all_scene_names, scene_path

def stats(g, o):
    g, o = np.ascontiguousarray(g[..., :3], np.float32), np.ascontiguousarray(o[..., :3], np.float32)
    diff = (g.view(np.uint32) != o.view(np.uint32)) & ~(np.isnan(g) & np.isnan(o))
    return {"floats_differ": int(diff.sum()), "pixels_differ": int(diff.any(-1).sum()), "pixels": int(diff.shape[0] * diff.shape[1]),
            "max_abs": float(np.nan_to_num(np.abs(g.astype(np.float64) - o)).max())}

out = open(os.path.join(ROOT, "gpurun_out", "radiance_exact_probe.jsonl"), "w")
W, H, frames = 320, 192, 3
for name in (sys.argv[1:] or all_scene_names()):
    fs = rtb.FlatScene.load(scene_path(name))
    po, sc = porthost.PortOracle(fs), api.open_scene(fs, counters=False)
    cam = po.camera_default(W, H)
    oacc, ost = po.render_pt(cam, porthost.default_params(abi.RT_INTEGRATOR_PATH, W, H), 1, frames, 1)
    for sched, sname in ((abi.RT_SCHEDULE_STREAMS, "streams"), (abi.RT_SCHEDULE_WAVEFRONT, "wavefront")):
        r = api.GpuRenderer(sc, abi.RT_INTEGRATOR_PATH, W, H, schedule=sched).Init()
        for _ in range(frames):
            r.Tick(0)
        rec = {"scene": name, "integrator": "pt", "schedule": sname, "calls": "tick", **stats(r.accumulator, oacc)}
        print(json.dumps(rec)); out.write(json.dumps(rec) + "\n")
        r.ClearAccumulator(); r.render(frames, first_spp=1)
        rec = {"scene": name, "integrator": "pt", "schedule": sname, "calls": "batched", **stats(r.accumulator, oacc)}
        print(json.dumps(rec)); out.write(json.dumps(rec) + "\n")
        r.close()
    ow, _ = po.render_whitted(cam, porthost.default_params(abi.RT_INTEGRATOR_WHITTED, W, H))
    r = api.GpuRenderer(sc, abi.RT_INTEGRATOR_WHITTED, W, H).Init()
    r.Tick(0)
    rec = {"scene": name, "integrator": "whitted", **stats(r.accumulator, ow)}
    print(json.dumps(rec)); out.write(json.dumps(rec) + "\n")
    r.close(); sc.close(); out.flush()
