"""Whitted frame time on the GPU next to the reference's CPU loop (BASELINE configs[0]: bunny 640x360, 1 spp).
usage: whitted_time.py [scene] [W H] [xml for the CPU reference]"""
import json, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import cpu_ray_tracer_b200 as rtb
from cpu_ray_tracer_b200 import abi, api
name = sys.argv[1] if len(sys.argv) > 1 else "bunny_flat"
W, H = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (640, 360)
xml = sys.argv[4] if len(sys.argv) > 4 else "bunny_scene.xml"
sc = api.open_scene(rtb.FlatScene.load(os.path.join(ROOT, "oracle", "_ref", "scenes", name + ".rtscene.gz")))
r = api.GpuRenderer(sc, abi.RT_INTEGRATOR_WHITTED, W, H).Init()
r.Tick(0); r.sync()
best = 1e9
for _ in range(20):
    r.reset_counters()
    t0 = time.perf_counter(); r.Tick(0); r.sync(); best = min(best, time.perf_counter() - t0)
c = r.counters()
rays = c["extension_rays"] + c["shadow_rays"]
t0 = time.perf_counter(); img = r.accumulator; e2e = best + (time.perf_counter() - t0)
out = {"config": f"Whitted {name} {W}x{H} 1 spp", "gpu_ms_per_frame": round(best * 1e3, 3), "rays": rays, "gpu_Mrays_per_s": round(rays / best / 1e6, 1),
       "gpu_Mpix_per_s": round(W * H / best / 1e6, 1), "gpu_ms_with_accumulator_readback": round(e2e * 1e3, 3)}
kind = "file" if name.endswith("flat") else "tlas"
try:
    from oracle import refhost
    if refhost.available("whitted", kind):
        p = subprocess.run([sys.executable, "-m", "oracle.refhost", "bench", "whitted", kind, xml, str(W), str(H), "20", "0"], cwd=ROOT, capture_output=True, text=True, timeout=600)
        ref = json.loads(p.stdout.strip().splitlines()[-1])
        out.update({"cpu_ms_per_frame": round(1e3 * ref["seconds"] / 20, 3), "cpu_threads": ref["threads"], "cpu_Mrays_per_s": round(rays * 20 / ref["seconds"] / 1e6, 1)})
except Exception as e:
    out["cpu_error"] = str(e)[:200]
print(json.dumps(out))
