"""A/B of the L2 access-policy window over the fat-node array (RT_B200_L2_PERSIST_MB) on the 10 M-triangle mesh of BASELINE
configs[4]: the four ray sets of bench.py's C5 block, one child process per setting (the limit is a device-wide setting).
usage: python tools/c5_l2_window.py [mb,mb,...]      -> one JSON line per setting"""
import json, os, subprocess, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(n_tris):
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import ray_bench
    from cpu_ray_tracer_b200 import abi, api, host_build
    tris = host_build.terrain_mesh(n_tris, seed=1)
    fs = host_build.flat_scene_from_tris(tris, builder=lambda t: (np.zeros(0, abi.NODE_DTYPE), np.zeros(0, np.uint32), 0.0))
    fs.device_build = True
    t0 = time.time()
    sc = api.open_scene(fs)
    create_s = time.time() - t0
    info = sc.info()
    W = H = 4096
    cam = api.Camera(W, H)
    cam.SetCameraState((0.0, 6.0, -4.0), (0.0, -0.5, 6.0))
    rays = ray_bench.primary(W, H, cam)
    hits = sc.FindNearest(rays)
    rng = np.random.default_rng(7)
    n = 1 << 24
    lo, hi = tris["v0"].min(0), tris["v0"].max(0)
    O = np.stack([rng.uniform(lo[0], hi[0], n), np.full(n, hi[1] + 0.5), rng.uniform(lo[2], hi[2], n)], 1).astype(np.float32)
    D = rng.normal(size=(n, 3)).astype(np.float32)
    D[:, 1] = -np.abs(D[:, 1]) - 1.0
    D /= np.linalg.norm(D, axis=1, keepdims=True).astype(np.float32)
    sets = {"primary": (rays, False), "bounce": (ray_bench.bounce_rays(fs, rays, hits), False),
            "shadow": (ray_bench.shadow_rays(fs, rays, hits), True), "scattered": (api.make_rays(O, D), False)}
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    out = {"l2_persist_mb": int(os.environ.get("RT_B200_L2_PERSIST_MB", "0")), "scene_create_s": round(create_s, 3), "fat_node_mb": info["fat_nodes"] * 64 / 1e6}
    for label, (r, occl) in sets.items():
        d_rays = torch.from_numpy(r.view(np.uint8).reshape(-1, 32)).cuda()
        m = len(r)
        if occl:
            res = torch.empty(m, dtype=torch.uint8, device="cuda")
            ms = ray_bench.time_batch(lambda: sc.IsOccludedDevice(d_rays.data_ptr(), res.data_ptr(), m, stream.cuda_stream), 5)
        else:
            res = torch.empty((m, 32), dtype=torch.uint8, device="cuda")
            ms = ray_bench.time_batch(lambda: sc.FindNearestDevice(d_rays.data_ptr(), res.data_ptr(), m, stream.cuda_stream), 5)
            ms2 = ray_bench.time_batch(lambda: sc.FindNearestDevice(d_rays.data_ptr(), res.data_ptr(), m, stream.cuda_stream, incoherent=True), 5)
            out[label + "_hint_incoherent_Mrays"] = round(m / ms2 / 1e3, 1)
        out[label + "_Mrays"] = round(m / ms / 1e3, 1)
        del d_rays, res
    print(json.dumps(out))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child(int(sys.argv[2]))
    else:
        for mb in (sys.argv[1] if len(sys.argv) > 1 else "0,32,64,96").split(","):
            env = dict(os.environ, RT_B200_L2_PERSIST_MB=mb)
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "child", "10000000"], env=env, capture_output=True, text=True, timeout=600)
            print(p.stdout.strip().splitlines()[-1] if p.returncode == 0 and p.stdout.strip() else f"mb={mb} failed: {p.stderr[-400:]}")
