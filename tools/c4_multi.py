"""BASELINE configs[3] across the GPUs of one box: the ~100M-triangle instanced TLAS scene at 3840x2160, sharded by TILE
(rank r renders tiles r, r + N, ... of the 240 x 135 tile grid, or a contiguous range with C4_TILES=contiguous; cpu-ray-tracer_b200/parallel.py).  Round 2: the mesh BVH
and the 20 129-instance TLAS are built on the device inside rt_scene_create (C4_BUILD=host for the round-1 path), and every rank writes its tiles
straight into rank 0's accumulator through peer-mapped memory (CUDA IPC; C4_COMBINE=reduce for the round-1 NCCL reduce of the float4 accumulators).  Launch:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1
--master-port P tools/c4_multi.py [spp] [n_instances]      (N = 1 works without torchrun)
Timing: CUDA events on the render stream, max over ranks; the scene build / upload is outside the timed region."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import cpu_ray_tracer_b200 as rtb
from cpu_ray_tracer_b200 import abi, api, host_build, parallel


def main():
    spp = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    n_inst = int(sys.argv[2]) if len(sys.argv) > 2 else 20129
    W, H = 3840, 2160
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    b = rtb.FlatScene.load(os.path.join(ROOT, "oracle", "_ref", "scenes", "bunny_flat.rtscene.gz"))
    mesh = b.tris.copy()
    c = (mesh["v0"].min(0) + mesh["v0"].max(0)) / 2
    for f in ("v0", "v1", "v2"):
        mesh[f] = (mesh[f] - c).astype(np.float32)
    mesh["centroid"] = ((mesh["v0"] + mesh["v1"]).astype(np.float32) + mesh["v2"]).astype(np.float32) * np.float32(0.3333)
    t0 = time.time()
    on_device = os.environ.get("C4_BUILD", "device") != "host"
    fs = host_build.instanced_grid(mesh, n_inst, tlas="none" if on_device else "host")
    fs.device_build = on_device
    host_s = time.time() - t0
    t0 = time.time()
    sc = api.GpuTLASFileScene(fs, device=local)
    create_s = time.time() - t0
    build_s = host_s + create_s
    interleaved = os.environ.get("C4_TILES", "interleaved") != "contiguous"
    shard = parallel.tile_shard(rank, world, W, H, interleaved=interleaved)
    r = api.GpuRenderer(sc, abi.RT_INTEGRATOR_PATH, W, H, tile_begin=shard.tile_begin, tile_end=shard.tile_end, tile_step=shard.tile_step).Init()
    side = int(np.ceil(n_inst ** (1 / 3)))
    r.camera.SetCameraState((0.0, side * 0.9, -side * 1.2), (0.0, side * 0.3, side * 0.8))
    stream = torch.cuda.Stream()
    r.set_stream(stream.cuda_stream)
    shared = os.environ.get("C4_COMBINE", "shared") != "reduce" and interleaved
    acc = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda") if not shared else None
    if shared:
        parallel.share_accumulator(r, rank, world)   # all ranks accumulate into rank 0's image over NVLink
    else:
        r.set_accumulator(acc.data_ptr())
    token = torch.zeros(1, device="cuda")
    with torch.cuda.stream(stream):
        r.render(2, first_spp=1)          # warm-up: pilot tile order, measured order on the second call
        r.render(2, first_spp=1)
        stream.synchronize()
        if world > 1:
            dist.barrier()
        r.ClearAccumulator() if shared else acc.zero_()
        stream.synchronize()
        r.reset_counters()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        r.render(spp, first_spp=1)
        if world > 1:
            dist.all_reduce(token) if shared else dist.reduce(acc, dst=0, op=dist.ReduceOp.SUM)  # shared: a 4-byte completion token
        e.record(stream)
        torch.cuda.synchronize()
    if shared and rank == 0:
        acc = torch.from_numpy(r.accumulator).cuda()
    ms = torch.tensor([a.elapsed_time(e)], device="cuda")
    cnt = r.counters()
    rays = torch.tensor([float(cnt["extension_rays"]), float(cnt["paths"])], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(rays, op=dist.ReduceOp.SUM)
    if rank == 0:
        t = float(ms.item()) * 1e-3
        print(json.dumps({"config": "c4", "n_gpus": world, "instances": n_inst, "triangles_total": int(n_inst) * len(mesh), "width": W, "height": H,
                          "spp": spp, "sharding": ("interleaved tiles" if interleaved else "contiguous tile ranges") + (", one accumulator on rank 0 written through peer-mapped memory" if shared else ", one NCCL reduce"), "ms": round(t * 1e3, 2), "rays": int(rays[0].item()),
                          "Mrays_per_s": round(rays[0].item() / t / 1e6, 1), "Msamples_per_s": round(rays[1].item() / t / 1e6, 1),
                          "scene_build_upload_s_per_rank": round(build_s, 2), "rt_scene_create_s": round(create_s, 3),
                          "scene_build": "mesh BVH + TLAS on the device (rt_scene_create)" if on_device else "host builders (librt_host.so), then upload",
                          "nonzero_pixel_fraction": round(float((acc[..., :3].sum(-1) > 0).float().mean().item()), 3),
                          "checksum": float(acc[..., :3].double().sum().item())}), flush=True)
    r.close(); sc.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
