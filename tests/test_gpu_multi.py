"""Tile-sharded rendering into ONE accumulator (run on the B200 box: pytest -m gpu; uses as many GPUs as the box shows).

  * rt_multi_renderer: several GPUs of one process, peer-mapped accumulator on devices[0]: bit-identical to one GPU;
  * rt_renderer_export_accumulator / _import_accumulator: the same between PROCESSES through CUDA IPC (what bench.py --gpus N
    does under torchrun) - exercised here with a second process, on a second GPU when there is one, else on the same GPU.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import GOLDEN, biteq

from cpu_ray_tracer_b200 import abi

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("name,lookahead", [("golden_file", 0), ("golden_tlas", 0), ("golden_file", 4), ("golden_kd", 0)])
def test_multi_renderer_is_bit_identical_to_one_gpu(name, lookahead, flat_scenes):
    from cpu_ray_tracer_b200 import api
    flat = flat_scenes(name)
    W, H, frames = 320, 192, 5
    sc = api.open_scene(flat)
    one = api.GpuRenderer(sc, abi.RT_INTEGRATOR_PATH, W, H).Init()
    one.render(frames, first_spp=1)
    ref, rays = one.accumulator, one.counters()["extension_rays"]
    one.close(), sc.close()
    n = api.device_count()
    for devices in ([0], list(range(min(n, 2))), list(range(n))):
        m = api.MultiGpuRenderer(flat, abi.RT_INTEGRATOR_PATH, W, H, devices, lookahead_frames=lookahead)
        if lookahead:
            for _ in range(frames):
                m.Tick(0)      # one frame per call, served from frames rendered ahead on every device
        else:
            m.render(frames, first_spp=1)
        assert biteq(m.accumulator, ref), f"{len(devices)} device(s)"
        if not lookahead:
            assert m.counters()["extension_rays"] == rays
        m.ClearAccumulator()
        assert not m.accumulator.any()
        m.render(2, first_spp=1)
        m.render(3, first_spp=3)
        assert biteq(m.accumulator, ref)
        m.close()


def test_multi_renderer_argument_errors(flat_scenes):
    from cpu_ray_tracer_b200 import api
    flat = flat_scenes("golden_file")
    with pytest.raises(api.RtError) as e:
        api.MultiGpuRenderer(flat, abi.RT_INTEGRATOR_PATH, 64, 48, [0, 0])
    assert e.value.status == abi.RT_ERR_INVALID
    with pytest.raises(api.RtError) as e:
        api.MultiGpuRenderer(flat, abi.RT_INTEGRATOR_PATH, 64, 48, [99])
    assert e.value.status == abi.RT_ERR_NO_DEVICE
    with pytest.raises(api.RtError) as e:
        api.MultiGpuRenderer(flat, abi.RT_INTEGRATOR_WHITTED, 64, 48, [0])
    assert e.value.status == abi.RT_ERR_UNSUPPORTED


CHILD = r"""
import sys
sys.path.insert(0, sys.argv[1])
import cpu_ray_tracer_b200 as rtb
from cpu_ray_tracer_b200 import abi, api
scene, W, H, frames, device, handle = sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6]), bytes.fromhex(sys.argv[7])
sc = api.open_scene(rtb.FlatScene.load(scene), device=device)
tiles = (W // 16) * (H // 16)
r = api.GpuRenderer(sc, abi.RT_INTEGRATOR_PATH, W, H, tile_begin=1, tile_end=tiles, tile_step=2).Init()
r.import_accumulator(handle)
r.ClearAccumulator()          # its own tiles only
r.render(frames, first_spp=1)
r.sync()
print("child rays", r.counters()["extension_rays"])
r.close(); sc.close()
"""


def test_two_processes_share_one_accumulator_through_cuda_ipc(flat_scenes):
    from cpu_ray_tracer_b200 import api
    path = os.path.join(GOLDEN, "golden_file.rtscene.gz")
    flat = flat_scenes("golden_file")
    W, H, frames = 320, 192, 4
    sc = api.open_scene(flat)
    full = api.GpuRenderer(sc, abi.RT_INTEGRATOR_PATH, W, H).Init()
    full.render(frames, first_spp=1)
    ref, rays = full.accumulator, full.counters()["extension_rays"]
    full.close()
    tiles = (W // 16) * (H // 16)
    mine = api.GpuRenderer(sc, abi.RT_INTEGRATOR_PATH, W, H, tile_begin=0, tile_end=tiles, tile_step=2).Init()
    handle = mine.export_accumulator()
    mine.render(frames, first_spp=1)
    mine.sync()
    half = mine.accumulator
    assert half[..., :3].sum() > 0 and not biteq(half, ref)
    device = 1 if api.device_count() > 1 else 0
    out = subprocess.run([sys.executable, "-c", CHILD, ROOT, path, str(W), str(H), str(frames), str(device), handle.hex()],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr[-3000:]
    child_rays = int(out.stdout.strip().split()[-1])
    assert child_rays + mine.counters()["extension_rays"] == rays
    assert biteq(mine.accumulator, ref), "the two shards together are not the one-GPU image"
    mine.close(), sc.close()
