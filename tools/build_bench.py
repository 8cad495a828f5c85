"""BVH build: the reference's SAH builder restated on the host (single thread, like the reference) vs rt_build_bvh on the GPU.
usage: build_bench.py [n_tris ...]"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cpu_ray_tracer_b200 as rtb
from cpu_ray_tracer_b200 import api, host_build
for n in [int(a) for a in sys.argv[1:]] or [100000, 1000000, 10000000]:
    tris = host_build.terrain_mesh(n, seed=1)
    api.build_bvh_gpu(tris[:1000])  # context + module load
    t0 = time.perf_counter(); nodes, idx, ms = api.build_bvh_gpu(tris); wall = time.perf_counter() - t0
    t0 = time.perf_counter(); rn, ri, depth = host_build.build_bvh(tris); host = time.perf_counter() - t0
    same = bool(np.array_equal(idx, ri) and nodes.tobytes() == rn.tobytes())
    print(json.dumps({"triangles": len(tris), "nodes": len(nodes), "max_depth": depth, "gpu_kernels_ms": round(ms, 2),
                      "gpu_call_ms_with_copies": round(wall * 1e3, 1), "host_reference_algorithm_ms": round(host * 1e3, 1),
                      "speedup_kernels": round(host * 1e3 / ms, 1), "speedup_call": round(host / wall, 1), "bit_identical": same}), flush=True)
