"""N > 1 host logic on CPU: world_size-2 gloo processes shard a path-traced job by sample index and by tile
range (cpu-ray-tracer_b200/parallel.py), each renders its share with the ORACLE standing in for the GPU
renderer (same rt_render_params contract: first_spp / count / stride, tile_begin / tile_end), the
accumulators are reduced onto rank 0 and compared with the single-process render."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT, scene_path


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, mode, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import cpu_ray_tracer_b200 as rtb
    from cpu_ray_tracer_b200 import abi, parallel
    from oracle import porthost
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        W, H, frames, first = 96, 64, 5, 1
        po = porthost.PortOracle(rtb.FlatScene.load(scene_path("golden_tlas")))
        cam = po.camera_default(W, H)
        acc = np.zeros((H, W, 4), np.float32)
        t_acc = torch.from_numpy(acc)
        rays = [0]

        def render_frames(first_spp, count, stride):
            p = porthost.default_params(abi.RT_INTEGRATOR_PATH, W, H)
            _, st = po.render_pt(cam, p, first_spp, count, stride, accumulator=acc)
            rays[0] += st["extension_rays"]

        def render_tiles(tile_begin, tile_end, first_spp, count, tile_step=1):
            p = porthost.default_params(abi.RT_INTEGRATOR_PATH, W, H)
            p.tile_begin, p.tile_end, p.tile_step = tile_begin, tile_end, tile_step
            _, st = po.render_pt(cam, p, first_spp, count, 1, accumulator=acc)
            rays[0] += st["extension_rays"]

        parallel.render_sharded(render_frames, t_acc, first, frames, mode=mode, width=W, height=H, render_tiles=render_tiles)
        r = torch.tensor([rays[0]], dtype=torch.int64)
        dist.reduce(r, dst=0)
        if rank == 0:
            full, st = po.render_pt(cam, porthost.default_params(abi.RT_INTEGRATOR_PATH, W, H), first, frames, 1)
            np.savez(os.path.join(out_dir, f"{mode}.npz"), sharded=acc, full=full, rays=int(r[0]), rays_full=st["extension_rays"])
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["frames", "tiles", "tiles_interleaved"])
def test_two_rank_sharding_reduces_to_the_single_process_image(mode, tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, mode, str(tmp_path)), nprocs=2, join=True)
    z = np.load(tmp_path / f"{mode}.npz")
    assert int(z["rays"]) == int(z["rays_full"])          # the shards trace exactly the job's rays, once
    assert z["full"][..., :3].sum() > 0
    if mode.startswith("tiles"):
        # disjoint pixels: the sum has nothing to reassociate
        assert np.array_equal(z["sharded"].view(np.uint32), z["full"].view(np.uint32))
    else:
        assert np.abs(z["sharded"] - z["full"]).max() <= 1e-4


def test_shard_arithmetic():
    sys.path.insert(0, ROOT)
    from cpu_ray_tracer_b200 import parallel
    for world in (1, 2, 3, 4, 8):
        for frames in (0, 1, 5, 8, 64, 257):
            per_rank = parallel.covered_frames(world, 7, frames)
            flat = sorted(x for r in per_rank for x in r)
            assert flat == list(range(7, 7 + frames))
            assert max(len(r) for r in per_rank) - min(len(r) for r in per_rank) <= 1
        for (w, h) in ((1920, 1080), (3840, 2160), (144, 90), (16, 16)):
            tiles = (w // 16) * (h // 16)
            shards = [parallel.tile_shard(r, world, w, h) for r in range(world)]
            assert shards[0].tile_begin == 0 and shards[-1].tile_end == tiles
            assert all(a.tile_end == b.tile_begin for a, b in zip(shards, shards[1:]))
            sizes = [s.tile_end - s.tile_begin for s in shards]
            assert max(sizes) - min(sizes) <= 1
            inter = [list(parallel.tile_shard(r, world, w, h, interleaved=True).tiles()) for r in range(world)]
            assert sorted(t for r in inter for t in r) == list(range(tiles))
            assert max(len(r) for r in inter) - min(len(r) for r in inter) <= 1
    with pytest.raises(ValueError):
        parallel.frame_shard(2, 2, 1, 4)
