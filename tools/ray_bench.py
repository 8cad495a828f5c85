"""Ray-throughput microbench through the C-ABI device entry points (BASELINE configs[4] style):
coherent primary rays vs incoherent diffuse-bounce closest-hit rays vs shadow-ray any-hit.
usage: ray_bench.py scene[,scene..] [W H] [reps]"""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import cpu_ray_tracer_b200 as rtb
from cpu_ray_tracer_b200 import abi, api


def bounce_rays(flat, rays, hits, seed=1):
    """uniform-hemisphere directions around the geometric 'up-facing' normal proxy: random unit vector flipped
    toward the incoming side (test input only; incoherent by construction)"""
    rng = np.random.default_rng(seed)
    m = hits["obj_idx"] >= 0
    I = rays["O"][m] + hits["t"][m, None] * rays["D"][m]
    R = rng.normal(size=I.shape).astype(np.float32)
    R /= np.linalg.norm(R, axis=1, keepdims=True).astype(np.float32)
    flip = (R * rays["D"][m]).sum(1) > 0
    R[flip] *= -1
    return api.make_rays(I + R * np.float32(1e-3), R)


def shadow_rays(flat, rays, hits):
    m = hits["obj_idx"] >= 0
    I = rays["O"][m] + hits["t"][m, None] * rays["D"][m]
    L = flat.header["light_pos"][0][None, :] - I
    dist = np.sqrt((L * L).sum(1)).astype(np.float32)
    L = (L / dist[:, None]).astype(np.float32)
    return api.make_rays(I + L * np.float32(0.001), L, dist - np.float32(0.002))


def primary(W, H, cam):
    x, y = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32))
    u, v = x * np.float32(1.0 / W), y * np.float32(1.0 / H)
    tl, tr, bl, pos = (np.array(list(a), np.float32) for a in (cam.c.top_left, cam.c.top_right, cam.c.bottom_left, cam.c.pos))
    P = tl + u[..., None] * (tr - tl) + v[..., None] * (bl - tl)
    D = P - pos
    D /= np.linalg.norm(D, axis=-1, keepdims=True)
    return api.make_rays(np.broadcast_to(pos, (W * H, 3)).copy(), D.reshape(-1, 3).astype(np.float32))


def time_batch(fn, reps):
    stream = torch.cuda.current_stream()
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); fn(); b.record(stream); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def main():
    names = (sys.argv[1] if len(sys.argv) > 1 else "wok_teapot_flat").split(",")
    W, H = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1920, 1080)
    reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
    s = torch.cuda.Stream(); torch.cuda.set_stream(s)
    for name in names:
        p = os.path.join(ROOT, "oracle", "_ref", "scenes", name + ".rtscene.gz")
        flat = rtb.FlatScene.load(p if os.path.exists(p) else name)
        sc = api.open_scene(flat)
        rays = primary(W, H, api.Camera(W, H))
        hits = sc.FindNearest(rays)
        sets = {"primary (coherent)": (rays, False), "diffuse bounce (incoherent)": (bounce_rays(flat, rays, hits), False),
                "shadow (any-hit)": (shadow_rays(flat, rays, hits), True)}
        # tile the small sets up to ~4M rays so that one launch is long enough to time
        for label, (r, occl) in sets.items():
            if len(r) == 0:
                continue
            k = max(1, (4 << 20) // len(r))
            r = np.tile(r, k) if label.startswith("primary") else np.concatenate([r] * k)
            d_rays = torch.from_numpy(r.view(np.uint8).reshape(-1, 32)).cuda()
            n = len(r)
            if occl:
                out = torch.empty(n, dtype=torch.uint8, device="cuda")
                ms = time_batch(lambda: sc.IsOccludedDevice(d_rays.data_ptr(), out.data_ptr(), n, s.cuda_stream), reps)
                frac = float(out.float().mean())
            else:
                out = torch.empty((n, 32), dtype=torch.uint8, device="cuda")
                ms = time_batch(lambda: sc.FindNearestDevice(d_rays.data_ptr(), out.data_ptr(), n, s.cuda_stream), reps)
                frac = float((out.cpu().numpy().reshape(-1).view(abi.HIT_DTYPE)["obj_idx"] >= 2).mean())
            print(json.dumps({"scene": name, "rays": label, "n": n, "ms": round(ms, 3), "Mrays_per_s": round(n / ms / 1e3, 1),
                              "hit_or_occluded_fraction": round(frac, 3), "traversal": os.environ.get("RT_B200_TRAVERSAL", "persistent")}), flush=True)
        sc.close()


if __name__ == "__main__":
    main()
