"""Quick GPU-vs-oracle check used during development (the real tests live in tests/)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import cpu_ray_tracer_b200 as rtb
from cpu_ray_tracer_b200 import abi, api
from oracle import porthost

def biteq(a, b):
    return np.array_equal(a.view(np.uint32), b.view(np.uint32))

def check_scene(name, W, H, frames=2):
    path = os.path.join(ROOT, "oracle", "_ref", "scenes", name + ".rtscene.gz")
    fs = rtb.FlatScene.load(path)
    po = porthost.PortOracle(fs)
    sc = api.open_scene(fs, counters=True)
    cam = po.camera_default(W, H)
    rays = po.primary_rays(cam, W, H)
    t0 = time.time(); oh, st = po.find_nearest(rays); t1 = time.time()
    gh = sc.FindNearest(rays); t2 = time.time()
    for f in ("t", "u", "v", "obj_idx", "tri_idx", "traversed", "tested"):
        print(name, "primary", f, "bit-equal:", biteq(oh[f], gh[f]), int((oh[f].view(np.uint32) != gh[f].view(np.uint32)).sum()))
    print("  oracle s", t1 - t0, "gpu e2e s", t2 - t1)
    # shadow rays
    I = rays["O"] + oh["t"][:, None] * rays["D"]
    m = oh["obj_idx"] >= 0
    L = fs.header["light_pos"][0][None, :] - I[m]
    dist = np.sqrt((L * L).sum(1)).astype(np.float32)
    L = (L / dist[:, None]).astype(np.float32)
    sr = api.make_rays(I[m] + L * np.float32(0.001), L, dist - np.float32(0.002))
    oo, _ = po.is_occluded(sr)
    go = sc.IsOccluded(sr)
    print(name, "occlusion equal:", np.array_equal(oo, go), oo.mean())
    for integ, nm in ((abi.RT_INTEGRATOR_WHITTED, "whitted"), (abi.RT_INTEGRATOR_PATH, "pt")):
        params = porthost.default_params(integ, W, H)
        if integ == abi.RT_INTEGRATOR_PATH:
            oacc, ost = po.render_pt(cam, params, 1, frames, 1)
        else:
            oacc, ost = po.render_whitted(cam, params)
        r = api.GpuRenderer(sc, integ, W, H).Init()
        t0 = time.time()
        r.render(frames if integ == abi.RT_INTEGRATOR_PATH else 1)
        gacc = r.accumulator
        t1 = time.time()
        c = r.counters()
        d = np.abs(gacc - oacc)
        bad = ~((gacc.view(np.uint32) == oacc.view(np.uint32)) | (np.isnan(gacc) & np.isnan(oacc)))
        print(name, nm, "max abs", np.nanmax(d), "mismatching floats", int(bad.sum()), "of", bad.size,
              "pixels >1e-4:", int((d.max(-1) > 1e-4).sum()), "gpu s", round(t1 - t0, 4))
        print("   oracle rays", ost["extension_rays"], ost["shadow_rays"], "gpu", c)
        r.close()
    sc.close()

if __name__ == "__main__":
    print("devices", api.device_count())
    for nm, W, H in (("bunny_flat", 640, 360), ("wok_teapot_flat", 480, 272), ("inside_tlas", 320, 192), ("instanced_tlas", 320, 192)):
        check_scene(nm, W, H)
