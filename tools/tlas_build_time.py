"""TLASBVH::Build on the device: cluster + distributed shared memory vs one CTA vs the host restatement (rt_build_tlas / rtb_build_tlas32).
usage: python tools/tlas_build_time.py [n,n,...]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cpu_ray_tracer_b200 import api, host_build
for n in [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "1000,5000,20129,40000").split(",")]:
    rng = np.random.default_rng(n)
    side = int(np.ceil(n ** (1 / 3)))
    i = np.arange(n)
    lo = (np.stack([i % side, (i // side) % side, i // (side * side)], 1) * 1.6 + rng.uniform(0, 0.5, (n, 3))).astype(np.float32)
    b = np.concatenate([lo, lo + 1.0], 1).astype(np.float32)
    t0 = time.time(); ref = host_build.build_tlas32(b); host_ms = (time.time() - t0) * 1e3
    out = [f"n {n:6d}  host {host_ms:9.1f} ms"]
    for mode in ("cluster128", "cluster256", "cluster512", "cluster1024", "single"):
        os.environ["RT_B200_TLAS_BUILD"] = mode[:7] if mode.startswith("cluster") else mode
        if mode.startswith("cluster"):
            os.environ["RT_B200_TLAS_THREADS"] = mode[7:]
        api.build_tlas_gpu(b[:64])
        got, ms = api.build_tlas_gpu(b, return_ms=True)
        ok = np.array_equal(got["left"], ref["left"]) and np.array_equal(got["right"], ref["right"]) and got["aabb_min"].tobytes() == ref["aabb_min"].tobytes()
        out.append(f"{mode} {ms:9.1f} ms ({'same tree' if ok else 'DIFFERENT TREE'})")
    print("   ".join(out), flush=True)
