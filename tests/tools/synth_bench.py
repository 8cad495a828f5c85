"""BASELINE configs[3] and configs[4] at full size (measurement tool with an oracle parity sample, hence under tests/: only
tests/, smoke() and bench.py's CPU legs may execute oracle/; run on the GPU box).
  c5: synthetic ~10M-triangle mesh in one flat SAH BVH (host-built with the reference's algorithm):
      coherent primary vs incoherent bounce closest-hit vs shadow any-hit through the C-ABI device entry points
  c4: one ~5k-triangle mesh instanced ~20k times under a TLAS (~100M triangles, ONE device copy of the mesh):
      path tracer at 3840x2160, reduced spp (stated), with the size-independent properties checked
usage: synth_bench.py c5 [n_tris] | c4 [n_instances] [spp] [W H]"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
import cpu_ray_tracer_b200 as rtb
from cpu_ray_tracer_b200 import abi, api, host_build


def c5(n_tris):
    import ray_bench
    t0 = time.time()
    tris = host_build.terrain_mesh(n_tris, seed=1)
    t1 = time.time()
    fs = host_build.flat_scene_from_tris(tris, builder=api.build_bvh_gpu if os.environ.get("RT_B200_GPU_BUILD", "1") != "0" else None)
    t2 = time.time()
    sc = api.open_scene(fs)
    t3 = time.time()
    print(json.dumps({"config": "c5", "triangles": len(tris), "bvh_nodes": len(fs.nodes), "mesh_s": round(t1 - t0, 2),
                      "sah_build_s": round(t2 - t1, 2), "builder": "rt_build_bvh (GPU)" if os.environ.get("RT_B200_GPU_BUILD", "1") != "0" else "host", "upload_relayout_s": round(t3 - t2, 2),
                      "device_geometry_MB": round((len(fs.nodes) / 2 * 64 + len(tris) * (48 + 64)) / 1e6, 1)}), flush=True)
    s = torch.cuda.Stream(); torch.cuda.set_stream(s)
    W, H = 4096, 4096   # 2^24 primary rays
    rays = ray_bench.primary(W, H, api.Camera(W, H))
    hits = sc.FindNearest(rays)
    sets = {"primary (coherent)": (rays, False), "diffuse bounce (incoherent)": (ray_bench.bounce_rays(fs, rays, hits), False),
            "shadow (any-hit)": (ray_bench.shadow_rays(fs, rays, hits), True)}
    for label, (r, occl) in sets.items():
        d_rays = torch.from_numpy(r.view(np.uint8).reshape(-1, 32)).cuda()
        n = len(r)
        if occl:
            out = torch.empty(n, dtype=torch.uint8, device="cuda")
            ms = ray_bench.time_batch(lambda: sc.IsOccludedDevice(d_rays.data_ptr(), out.data_ptr(), n, s.cuda_stream), 5)
            frac = float(out.float().mean())
        else:
            out = torch.empty((n, 32), dtype=torch.uint8, device="cuda")
            ms = ray_bench.time_batch(lambda: sc.FindNearestDevice(d_rays.data_ptr(), out.data_ptr(), n, s.cuda_stream), 5)
            frac = float((out.cpu().numpy().reshape(-1).view(abi.HIT_DTYPE)["obj_idx"] >= 2).mean())
        print(json.dumps({"config": "c5", "rays": label, "n": n, "ms": round(ms, 3), "Mrays_per_s": round(n / ms / 1e3, 1),
                          "hit_or_occluded_fraction": round(frac, 3)}), flush=True)
    # parity on a sample: the oracle traces 2^16 of the same rays
    from oracle import porthost
    po = porthost.PortOracle(fs)
    sel = np.random.default_rng(0).choice(len(rays), 1 << 16, replace=False)
    for label in ("primary (coherent)", "diffuse bounce (incoherent)"):
        r = sets[label][0]
        sub = r[sel[sel < len(r)]]
        ref, st = po.find_nearest(sub)
        got = sc.FindNearest(sub)
        same = all(np.array_equal(ref[f].view(np.uint32), got[f].view(np.uint32)) for f in ("t", "u", "v", "obj_idx", "tri_idx"))
        rays_n = len(sub)
        I = (st["interior_visits"] + st["tlas_interior_visits"]) / rays_n
        T = st["tri_tests"] / rays_n
        print(json.dumps({"config": "c5", "parity_sample": label, "rays": rays_n, "bit_exact": bool(same),
                          "interior_visits_per_ray": round(I, 2), "tri_tests_per_ray": round(T, 2), "algorithmic_bytes_per_ray": round(64 * I + 52 * T + 48, 1)}), flush=True)
    sc.close()


def c4(n_inst, spp, W, H):
    p = os.path.join(ROOT, "oracle", "_ref", "scenes", "bunny_flat.rtscene.gz")
    if os.path.exists(p):
        b = rtb.FlatScene.load(p)
        mesh = b.tris.copy()
        c = (mesh["v0"].min(0) + mesh["v0"].max(0)) / 2
        for f in ("v0", "v1", "v2"):
            mesh[f] = (mesh[f] - c).astype(np.float32)
        mesh["centroid"] = ((mesh["v0"] + mesh["v1"]).astype(np.float32) + mesh["v2"]).astype(np.float32) * np.float32(0.3333)
        name = "bunny.obj (4968 triangles)"
    else:
        mesh, name = host_build.terrain_mesh(4968, seed=2, size=1.0, height=0.6), "terrain patch"
    t0 = time.time()
    fs = host_build.instanced_grid(mesh, n_inst)
    t1 = time.time()
    sc = api.open_scene(fs)
    t2 = time.time()
    print(json.dumps({"config": "c4", "mesh": name, "instances": n_inst, "triangles_total": int(n_inst) * len(mesh),
                      "tlas_nodes": len(fs.tlas_nodes), "host_build_s": round(t1 - t0, 2), "upload_s": round(t2 - t1, 2)}), flush=True)
    r = api.GpuRenderer(sc, abi.RT_INTEGRATOR_PATH, W, H).Init()
    side = int(np.ceil(n_inst ** (1 / 3)))
    r.camera.SetCameraState((0.0, side * 0.9, -side * 1.2), (0.0, side * 0.3, side * 0.8))
    r.render(1, first_spp=1); r.sync()
    best = 1e9
    for _ in range(2):
        r.ClearAccumulator(); r.reset_counters()
        t = time.perf_counter(); r.render(spp, first_spp=1); r.sync(); best = min(best, time.perf_counter() - t)
    cnt = r.counters()
    acc = r.accumulator
    print(json.dumps({"config": "c4", "width": W, "height": H, "spp": spp, "ms": round(best * 1e3, 2), "rays": cnt["extension_rays"],
                      "Mrays_per_s": round(cnt["extension_rays"] / best / 1e6, 1), "Msamples_per_s": round(cnt["paths"] / best / 1e6, 1),
                      "rays_per_path": round(cnt["extension_rays"] / cnt["paths"], 3), "nonzero_pixel_fraction": round(float((acc[..., :3].sum(-1) > 0).mean()), 3)}), flush=True)
    # size-independent property at full size: tile sharding in two halves reproduces the image
    tiles = (W // 16) * (H // 16)
    lo = api.GpuRenderer(sc, abi.RT_INTEGRATOR_PATH, W, H, tile_begin=0, tile_end=tiles // 2).Init()
    hi = api.GpuRenderer(sc, abi.RT_INTEGRATOR_PATH, W, H, tile_begin=tiles // 2, tile_end=tiles).Init()
    for q in (lo, hi):
        q.camera.c = r.camera.c
        q.render(spp, first_spp=1)
    d = np.abs(lo.accumulator + hi.accumulator - acc).max()
    print(json.dumps({"config": "c4", "tile_shard_sum_max_abs_diff": float(d),
                      "rays_equal": lo.counters()["extension_rays"] + hi.counters()["extension_rays"] == cnt["extension_rays"]}), flush=True)


if __name__ == "__main__":
    which = sys.argv[1]
    if which == "c5":
        c5(int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000)
    else:
        c4(int(sys.argv[2]) if len(sys.argv) > 2 else 20129, int(sys.argv[3]) if len(sys.argv) > 3 else 8,
           int(sys.argv[4]) if len(sys.argv) > 4 else 3840, int(sys.argv[5]) if len(sys.argv) > 5 else 2160)
