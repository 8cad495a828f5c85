"""Multi-GPU partitioning of a render job (one process per GPU, torch.distributed for the plumbing).

Pixels, tiles and samples are independent in the reference (every ProcessTile touches its own 256
accumulator entries and its own seed, 3. PathTracer/renderer.cpp:117-131), so the job shards with no
data-path exchange; the only collective is one reduce of the float4 accumulators at the end.

  sample-index sharding (default): rank r renders the frames whose reference `spp` counter is
      first_spp + r, first_spp + r + N, ...      (seeds depend on spp: InitSeed(tx + ty*W + spp*1799))
  tile sharding: rank r renders every N-th tile of the (W/16) x (H/16) row-major tile grid starting at tile r
      (interleaved, the default: cost varies across an image, 6.1x -> see profiles/r1_c4_multi_gpu_tile_sharding.jsonl)
      or a contiguous range of it

Both give, after the sum over ranks, the single-process image up to float reassociation.

  shared accumulator (round 2, what bench.py --gpus N uses): with interleaved tile shards the ranks touch disjoint
      pixels, so they can all accumulate into ONE image: rank 0 exports its accumulator as a CUDA IPC handle
      (share_accumulator), the others map it (peer memory over NVLink) and their frame-ordered sums write their tiles
      straight into it.  No reduce, no gather; after a barrier the image on rank 0 is bit-identical to a one-GPU render.

This module is host logic only (no CUDA): the caller supplies the per-rank render function.
"""
from dataclasses import dataclass


@dataclass(frozen=True)
class FrameShard:
    first_spp: int
    count: int
    stride: int


@dataclass(frozen=True)
class TileShard:
    tile_begin: int
    tile_end: int
    tile_step: int = 1

    def tiles(self):
        return range(self.tile_begin, self.tile_end, self.tile_step)


def frame_shard(rank, world, first_spp, frames):
    """frames first_spp .. first_spp+frames-1 dealt round-robin; ranks beyond `frames` get count 0"""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world {world}")
    count = (frames - rank + world - 1) // world if frames > rank else 0
    return FrameShard(first_spp + rank, count, world)


def tile_shard(rank, world, width, height, interleaved=False):
    """contiguous tile ranges whose sizes differ by at most one tile, or (interleaved) tiles rank, rank + world, ..."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world {world}")
    tiles = (width // 16) * (height // 16)  # integer division as renderer.cpp:151
    if interleaved:
        return TileShard(rank, tiles, world)
    base, extra = divmod(tiles, world)
    begin = rank * base + min(rank, extra)
    return TileShard(begin, begin + base + (1 if rank < extra else 0))


def covered_frames(world, first_spp, frames):
    """every spp counter of the job, per rank (used by tests: a partition, no overlap, nothing missing)"""
    out = []
    for r in range(world):
        s = frame_shard(r, world, first_spp, frames)
        out.append([s.first_spp + i * s.stride for i in range(s.count)])
    return out


def reduce_accumulator(acc, dst=0):
    """sum of the per-rank float4 accumulators onto `dst` (NCCL over NVLink on GPUs, gloo in CPU tests)"""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(acc, dst=dst, op=dist.ReduceOp.SUM)
    return acc


def render_sharded(render_frames, acc, first_spp, frames, mode="frames", width=None, height=None, render_tiles=None):
    """Runs this rank's share and reduces onto rank 0.

    render_frames(first_spp, count, stride) accumulates into `acc` (a torch tensor the renderer writes to);
    render_tiles(tile_begin, tile_end, first_spp, count[, tile_step]) likewise for mode == "tiles" (contiguous ranges)
    and mode == "tiles_interleaved".
    """
    import torch.distributed as dist
    init = dist.is_available() and dist.is_initialized()
    rank = dist.get_rank() if init else 0
    world = dist.get_world_size() if init else 1
    if mode == "frames":
        s = frame_shard(rank, world, first_spp, frames)
        if s.count:
            render_frames(s.first_spp, s.count, s.stride)
    elif mode == "tiles":
        t = tile_shard(rank, world, width, height)
        if t.tile_end > t.tile_begin:
            render_tiles(t.tile_begin, t.tile_end, first_spp, frames)
    elif mode == "tiles_interleaved":
        t = tile_shard(rank, world, width, height, interleaved=True)
        if t.tile_end > t.tile_begin:
            render_tiles(t.tile_begin, t.tile_end, first_spp, frames, t.tile_step)
    else:
        raise ValueError(mode)
    return reduce_accumulator(acc)


def share_accumulator(renderer, rank, world, src=0):
    """Rank `src` exports its renderer's accumulator (renderer.export_accumulator() -> bytes), every other rank maps it with
    renderer.import_accumulator(bytes).  The handle travels through torch.distributed's object broadcast (host side, once)."""
    import torch.distributed as dist
    if world <= 1:
        return None
    box = [renderer.export_accumulator() if rank == src else None]
    dist.broadcast_object_list(box, src=src)
    if rank != src:
        renderer.import_accumulator(box[0])
    return box[0]
