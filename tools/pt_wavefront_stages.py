"""Per-stage device time of the wavefront path tracer (generate / extend / shade) with one slot per pixel
(RT_SEED_PER_PIXEL) or per tile (reference RNG): pt_wavefront_stages.py scene spp seed_mode   (development tool;
run under `ncu --metrics ... -k regex:k_pt_` for lanes per instruction and issue utilisation per stage)"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cpu_ray_tracer_b200 as rtb
from cpu_ray_tracer_b200 import abi, api
name = sys.argv[1] if len(sys.argv) > 1 else "wok_teapot_flat"
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 64
mode = int(sys.argv[3]) if len(sys.argv) > 3 else abi.RT_SEED_PER_PIXEL
W, H = 1920, 1080
sc = api.open_scene(rtb.FlatScene.load(os.path.join(ROOT, "oracle", "_ref", "scenes", name + ".rtscene.gz")))
r = api.GpuRenderer(sc, abi.RT_INTEGRATOR_PATH, W, H, schedule=abi.RT_SCHEDULE_WAVEFRONT, seed_mode=mode).Init()
r.render(spp, first_spp=1); r.sync()
r.set_profiling(True); r.reset_counters()
r.render(spp, first_spp=1)
st = r.stage_times()
c = r.counters()
total = sum(v[0] for v in st.values())
print(json.dumps({"scene": name, "spp": spp, "seed_mode": "per_pixel" if mode else "reference_tile", "rays": c["extension_rays"],
                  "stage_ms": {k: round(v[0], 2) for k, v in st.items()}, "launches": {k: v[1] for k, v in st.items()},
                  "sum_ms": round(total, 2), "Mrays_per_s_of_stage_sum": round(c["extension_rays"] / total / 1e3, 1)}))
