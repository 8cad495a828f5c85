"""The C-ABI library loads and exports every symbol include/rt_b200.h declares; the ctypes mirror has the
header's struct sizes; compute calls fail loudly (never fall back) without a CUDA device."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "rt_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rt_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_expected_surface():
    names = declared_functions()
    for required in ("rt_scene_create", "rt_find_nearest", "rt_is_occluded", "rt_renderer_create",
                     "rt_renderer_render", "rt_renderer_read_accumulator", "rt_camera_look_at"):
        assert required in names


def test_library_exports_every_declared_symbol():
    from cpu_ray_tracer_b200 import api
    L = api.lib()
    for name in declared_functions():
        assert hasattr(L, name), f"librt_b200.so does not export {name}"
    assert set(api.EXPORTS) == set(declared_functions())
    assert L.rt_abi_version() == 5


def test_struct_sizes_match_header(tmp_path):
    """compile a C probe against the header and compare sizeof() with the ctypes mirror"""
    from cpu_ray_tracer_b200 import abi
    structs = ["rt_bvh_node", "rt_tri", "rt_tlas_node", "rt_blas_desc", "rt_material", "rt_texture",
               "rt_scene_desc", "rt_ray", "rt_hit", "rt_camera", "rt_render_params", "rt_counters", "rt_kd_node", "rt_grid_desc"]
    probe = tmp_path / "probe.c"
    probe.write_text('#include <stdio.h>\n#include "rt_b200.h"\nint main(void){\n' +
                     "".join(f'printf("{s} %zu\\n", sizeof({s}));\n' for s in structs) + "return 0;}\n")
    exe = tmp_path / "probe"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(probe), "-o", str(exe)], check=True)
    sizes = dict(line.split() for line in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    mirror = {"rt_bvh_node": abi.NODE_DTYPE.itemsize, "rt_tri": abi.TRI_DTYPE.itemsize,
              "rt_tlas_node": abi.TLAS_NODE_DTYPE.itemsize, "rt_material": abi.MATERIAL_DTYPE.itemsize,
              "rt_ray": abi.RAY_DTYPE.itemsize, "rt_hit": abi.HIT_DTYPE.itemsize,
              "rt_blas_desc": C.sizeof(abi.rt_blas_desc), "rt_texture": C.sizeof(abi.rt_texture),
              "rt_scene_desc": C.sizeof(abi.rt_scene_desc), "rt_camera": C.sizeof(abi.rt_camera),
              "rt_render_params": C.sizeof(abi.rt_render_params), "rt_counters": C.sizeof(abi.rt_counters),
              "rt_kd_node": abi.KD_NODE_DTYPE.itemsize, "rt_grid_desc": C.sizeof(abi.rt_grid_desc),
              "rt_material_ct": C.sizeof(abi.rt_material)}
    for s in structs:
        assert int(sizes[s]) == mirror[s], s
    assert mirror["rt_material_ct"] == int(sizes["rt_material"])


def test_no_cpu_fallback(flat_scenes):
    """without a device the product path must refuse, not compute on the host"""
    from cpu_ray_tracer_b200 import abi, api
    if api.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(api.RtError) as e:
        api.open_scene(flat_scenes("golden_file"))
    assert e.value.status == abi.RT_ERR_NO_DEVICE


def test_eval_shading_math_argument_checks():
    """rt_eval_shading_math validates before it touches the device; without a device it refuses like every compute call"""
    import numpy as np
    from cpu_ray_tracer_b200 import abi, api
    L = api.lib()
    a = np.zeros(4, np.float32)
    out = np.zeros(4, np.float32)
    assert L.rt_eval_shading_math(0, 7, a.ctypes.data, None, out.ctypes.data, 4) == abi.RT_ERR_INVALID
    assert L.rt_eval_shading_math(0, abi.RT_MATH_ATAN2F, a.ctypes.data, None, out.ctypes.data, 4) == abi.RT_ERR_INVALID
    assert L.rt_eval_shading_math(0, abi.RT_MATH_EXPF, None, None, out.ctypes.data, 4) == abi.RT_ERR_INVALID
    assert L.rt_eval_shading_math(99, abi.RT_MATH_EXPF, a.ctypes.data, None, out.ctypes.data, 4) == abi.RT_ERR_NO_DEVICE
    if api.device_count() == 0:
        with pytest.raises(api.RtError) as e:
            api.eval_shading_math(abi.RT_MATH_EXPF, a)
        assert e.value.status == abi.RT_ERR_NO_DEVICE


def test_product_does_not_touch_the_oracle():
    """the oracle is test infrastructure: nothing under the package may import or link it"""
    pkg = os.path.join(ROOT, "cpu-ray-tracer_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, fn), errors="replace").read()
                assert "porthost" not in text and "refhost" not in text and "liboracle" not in text and "rt_oracle" not in text, fn
