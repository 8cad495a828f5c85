"""The C++ drop-in adapters (cpu-ray-tracer_b200/host/rt_b200_adapters.h): the reference's own scene loaders and
SAH / TLAS builders feed rtb200::GpuScene / GpuRenderer, which call the C-ABI; results are compared in the
same process with the reference's own Renderer / Scene on the same XML scene.

The adapter libraries (oracle/_ref/libgpuhost_*.so) are built where /root/reference is mounted
(oracle/ref_build/build_ref.py) and travel to the GPU box; without them the tests skip.
"""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

WHITTED_TOL = 2e-5
CASES = [("pt", "file", "wok_teapot_scene.xml"), ("pt", "tlas", "inside_scene.xml"),
         ("whitted", "file", "bunny_scene.xml"), ("whitted", "tlas", "instanced_scene.xml"),
         # FileScene exactly as the reference ships it (KD-tree), and with its grid
         ("pt", "file_kd", "wok_teapot_scene.xml"), ("whitted", "file_grid", "bunny_scene.xml"),
         # TLASFileScene over per-object KD-trees / grids
         ("whitted", "tlas_kd", "instanced_scene.xml"), ("pt", "tlas_grid", "inside_scene.xml")]


def test_adapter_header_cites_and_covers_the_surface():
    """host logic, no GPU: the adapter implements every BaseScene virtual and the Renderer members"""
    src = open(os.path.join(ROOT, "cpu-ray-tracer_b200", "host", "rt_b200_adapters.h")).read()
    for virt in ("SetTime", "GetSkyColor", "GetLightPos", "GetLightColor", "FindNearest", "IsOccluded", "GetAlbedo",
                 "GetHitInfo", "GetTriangleCount"):
        assert virt + "(" in src, virt
    for member in ("accumulator", "camera", "spp", "passes", "depthLimit", "Init()", "Tick(", "ClearAccumulator()"):
        assert member in src, member
    assert "rt_scene_create" in src and "rt_renderer_render" in src


@pytest.mark.gpu
@pytest.mark.parametrize("integ,kind,xml", CASES)
def test_cpp_adapter_matches_reference(integ, kind, xml):
    from oracle import gpuhost_check
    if not gpuhost_check.available(integ, kind):
        pytest.skip("oracle/_ref/libgpuhost_* not built (needs /root/reference at build time)")
    W, H, frames = 192, 112, 2
    p = subprocess.run([sys.executable, "-m", "oracle.gpuhost_check", integ, kind, xml, str(W), str(H), str(frames)],
                       cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:] + p.stdout[-500:]
    r = json.loads(p.stdout.strip().splitlines()[-1])
    # FindNearest through the adapter: bit-exact against the reference's own traversal
    assert r["hit_fraction"] > 0.02
    assert all(v == 0 for v in r["find_nearest_mismatches"].values()), r["find_nearest_mismatches"]
    assert r["single_ray_mismatches"] == 0
    assert r["occlusion_mismatches"] == 0 and 0.0 < r["occluded_fraction"] < 1.0
    for tag, t in r["tick"].items():
        assert t["mean"] > 0
        if integ == "whitted":
            assert t["max_abs"] <= WHITTED_TOL, (tag, t)
        else:
            # cam0: passes = 1; cam1: passes = 2 (spp advances by `passes` per Tick, renderer.cpp:167)
            assert t["ref_spp"] == t["gpu_spp"] == 1 + frames * (2 if tag == "cam1" else 1)
            assert t["pixels_over_1e-4"] <= max(2, 2e-4 * W * H * frames * (2 if tag == "cam1" else 1)), (tag, t)
            assert t["rmse"] < 2e-3, (tag, t)
        # screen->pixels (RGBF32_to_RGB8 of accumulator * scale): at most one 8-bit step apart
        assert t["screen_max_channel_diff"] <= 1 or t["screen_pixels_differing"] <= 4, (tag, t)
    info = r["scene_info"]
    if xml == "instanced_scene.xml" and kind == "tlas":
        # true instancing from the XML: the reference builds one BLASBVH per <object>; the adapter hands identical ones over once
        # (2 x watch-tower and 4 x log_fence at equal scale -> 6 meshes for 10 instances)
        assert info["instances"] == 10 and info["meshes"] < info["instances"], info


@pytest.mark.gpu
def test_cpp_adapter_tick_on_all_gpus_of_the_box():
    """rtb200::GpuRenderer constructed with a device list: Renderer::Tick's tile jobs are dealt to the GPUs (rt_multi_renderer),
    the accumulator the reference's callers read is the same image"""
    from cpu_ray_tracer_b200 import api
    from oracle import gpuhost_check
    if not gpuhost_check.available("pt", "tlas"):
        pytest.skip("oracle/_ref/libgpuhost_* not built (needs /root/reference at build time)")
    n = api.device_count()
    if n < 2:
        pytest.skip("one GPU on this box: the device-list form is covered by tests/test_gpu_multi.py through the C-ABI")
    W, H, frames = 192, 112, 2
    outs = []
    for devices in ("0", ",".join(str(d) for d in range(n))):
        p = subprocess.run([sys.executable, "-m", "oracle.gpuhost_check", "pt", "tlas", "inside_scene.xml", str(W), str(H), str(frames), devices],
                           cwd=ROOT, capture_output=True, text=True, timeout=600)
        assert p.returncode == 0, p.stderr[-2000:] + p.stdout[-500:]
        outs.append(json.loads(p.stdout.strip().splitlines()[-1]))
    one, many = outs
    assert many["scene_info"]["multi_devices"] == n and one["scene_info"]["multi_devices"] == 0
    for tag in one["tick"]:
        assert one["tick"][tag] == many["tick"][tag], (tag, one["tick"][tag], many["tick"][tag])   # same statistics against the reference, digit for digit
