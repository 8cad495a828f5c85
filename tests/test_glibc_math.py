"""expf / acosf / atan2f of the shading code = the host libm's, bit for bit.

The reference's integrators call exp() on floats (Beer's law, renderer.cpp:76-80) and atan2f / acosf (GetSkyColor,
file_scene.cpp:142-154) from the system libm (glibc 2.39 here and on the GPU box).  csrc/rt_glibc_math.cuh restates those
routines for the device; this file pins the restatement:
  * CPU: the header compiled for the host against libm, strided over the 2^32 arguments (every argument with
    RT_GLIBC_MATH_EXHAUSTIVE=1: 0 mismatches for expf, acosf, atanf; 6e8 atan2f pairs), including the sky lookup's inputs;
  * GPU: the device code (rt_eval_shading_math, i.e. exactly the functions the kernels inline) against libm;
  * GPU: with identical hits, identical RNG and identical libm bits the path tracer's accumulator is the oracle's, bit for bit.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, all_scene_names, biteq

from cpu_ray_tracer_b200 import abi

SRC = os.path.join(ROOT, "tests", "tools", "glibc_math_check.c")
INC = os.path.join(ROOT, "cpu-ray-tracer_b200", "csrc")
FLAGS = ["-O2", "-ffp-contract=off", "-fopenmp", "-mfma", "-I", INC]


@pytest.fixture(scope="module")
def checker(tmp_path_factory):
    d = tmp_path_factory.mktemp("glibc_math")
    exe, so = str(d / "glibc_math_check"), str(d / "libglibc_math_check.so")
    subprocess.run(["gcc", *FLAGS, SRC, "-o", exe, "-lm"], check=True)
    subprocess.run(["gcc", *FLAGS, "-fPIC", "-shared", SRC, "-o", so, "-lm"], check=True)
    L = C.CDLL(so)
    for f in (L.libm_eval, L.restated_eval):
        f.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        f.restype = None
    return exe, L


def host_eval(f, fn, a, b=None):
    a = np.ascontiguousarray(a, np.float32)
    b = None if b is None else np.ascontiguousarray(b, np.float32)
    out = np.empty_like(a)
    f(fn, a.ctypes.data, None if b is None else b.ctypes.data, out.ctypes.data, a.size)
    return out


def same_bits(x, y):
    return (x.view(np.uint32) == y.view(np.uint32)) | (np.isnan(x) & np.isnan(y))


def test_restatement_equals_host_libm(checker):
    exe, _ = checker
    exhaustive = os.environ.get("RT_GLIBC_MATH_EXHAUSTIVE") == "1"
    r = subprocess.run([exe, "1" if exhaustive else "1021", "200000000" if exhaustive else "1000000"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr[-2000:]
    assert r.stdout.count("mismatches 0") == 4


def sweep_arguments():
    bits = np.arange(0, 1 << 32, 257, dtype=np.uint64).astype(np.uint32)
    x = bits.view(np.float32)
    rng = np.random.default_rng(11)
    d = rng.normal(size=(4_000_000, 3)).astype(np.float32)
    d /= np.sqrt((d * d).sum(1, dtype=np.float32))[:, None]
    pairs = rng.integers(0, 1 << 32, (2, 2_000_000), dtype=np.uint64).astype(np.uint32).view(np.float32)
    y2 = np.concatenate([-d[:, 2], pairs[0], np.float32([0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, np.nan, 0.0])])
    x2 = np.concatenate([d[:, 0], pairs[1], np.float32([-1.0, -1.0, 0.0, 0.0, -np.inf, np.inf, 1.0, 0.0])])
    absorb = -(rng.uniform(0, 8, 4_000_000).astype(np.float32) * rng.uniform(0, 30, 4_000_000).astype(np.float32))
    return x, np.concatenate([x, absorb]), np.concatenate([x, -d[:, 1]]), y2, x2


def test_restated_arrays_equal_libm(checker):
    """the array entry points used by the GPU test below agree on the host (guards the harness itself)"""
    _, L = checker
    _, xe, xa, y2, x2 = sweep_arguments()
    for fn, a, b in ((0, xe[::16], None), (1, xa[::16], None), (2, y2[::4], x2[::4])):
        assert same_bits(host_eval(L.restated_eval, fn, a, b), host_eval(L.libm_eval, fn, a, b)).all()


def require_fma_libm(L):
    """glibc selects __expf_fma by ifunc on hosts with FMA + AVX2; the restatement follows that variant (any B200 host has both)"""
    x = np.linspace(-60, 60, 200001, dtype=np.float32)
    if not same_bits(host_eval(L.restated_eval, 0, x), host_eval(L.libm_eval, 0, x)).all():
        pytest.skip("host libm did not select the FMA variant of expf (CPU without FMA/AVX2): bit parity is defined on FMA hosts")


@pytest.mark.gpu
def test_device_math_equals_host_libm(checker):
    from cpu_ray_tracer_b200 import api
    _, L = checker
    require_fma_libm(L)
    _, xe, xa, y2, x2 = sweep_arguments()
    for what, fn, a, b in (("expf", abi.RT_MATH_EXPF, xe, None), ("acosf", abi.RT_MATH_ACOSF, xa, None), ("atan2f", abi.RT_MATH_ATAN2F, y2, x2)):
        got, ref = api.eval_shading_math(fn, a, b), host_eval(L.libm_eval, fn, a, b)
        bad = ~same_bits(got, ref)
        assert not bad.any(), f"{what}: {int(bad.sum())} of {a.size} device results differ from libm, first at {a[bad][:4]}"


@pytest.mark.gpu
def test_sky_texel_filter_never_accepts_a_different_texel():
    """the sky lookup accepts the texel found with CUDA's atan2f / acosf only away from texel borders (rt_device.cuh, sky_color):
    whenever it accepts, the texel is the one the restated glibc routines choose; the rest (a few %) goes to those routines"""
    from cpu_ray_tracer_b200 import api
    rng = np.random.default_rng(5)
    n = 16_000_000
    az = rng.uniform(-np.pi, np.pi, n).astype(np.float32)
    h = rng.uniform(-1, 1, n).astype(np.float32)
    # directions on texel borders of the 4096 x 2048 sky (phi = 2 pi i / 4096, theta = pi j / 2048) and just beside them
    i, j = rng.integers(0, 4096, n // 8), rng.integers(1, 2048, n // 8)
    jitter = rng.choice(np.float32([0, 1e-7, -1e-7, 1e-6, -1e-6, 3e-6, -3e-6, 1e-5]), n // 8)
    az[: n // 8] = (2 * np.pi * i / 4096 - np.pi + jitter).astype(np.float32)
    h[n // 8: n // 4] = (-np.cos(np.pi * j / 2048 + jitter)).astype(np.float32)
    out = api.eval_shading_math(abi.RT_MATH_SKY_TEXEL, az, h)
    assert not (out < 0).any(), f"{int((out < 0).sum())} lookups accepted a texel that differs from glibc's"
    handed = float((out[n // 4:] == 0).mean())
    assert 0.005 < handed < 0.06, f"fraction of random lookups handed to the exact routines: {handed}"


def render_pair(name, oracles, gpu_scenes, schedule, W=320, H=192, frames=3):
    from cpu_ray_tracer_b200 import api
    from oracle import porthost
    po, sc = oracles(name), gpu_scenes(name, counters=False)
    cam = po.camera_default(W, H)
    oacc, ost = po.render_pt(cam, porthost.default_params(abi.RT_INTEGRATOR_PATH, W, H), 1, frames, 1)
    r = api.GpuRenderer(sc, abi.RT_INTEGRATOR_PATH, W, H, schedule=schedule).Init()
    for _ in range(frames):
        r.Tick(0)
    gacc = r.accumulator.copy()
    rays = r.counters()["extension_rays"]
    r.close()
    return gacc, oacc, rays, ost["extension_rays"]


@pytest.mark.gpu
@pytest.mark.parametrize("schedule", [abi.RT_SCHEDULE_STREAMS, abi.RT_SCHEDULE_WAVEFRONT], ids=["streams", "wavefront"])
@pytest.mark.parametrize("name", all_scene_names())
def test_path_tracer_accumulator_bit_identical(name, schedule, oracles, gpu_scenes, checker):
    """reference RNG, one Tick per frame: every float of the accumulator equals the oracle's (which is pinned to the reference)"""
    require_fma_libm(checker[1])
    gacc, oacc, rays, orays = render_pair(name, oracles, gpu_scenes, schedule)
    assert rays == orays
    diff = ~same_bits(gacc[..., :3].astype(np.float32), oacc[..., :3].astype(np.float32))
    assert not diff.any(), f"{name}: {int(diff.any(-1).sum())} of {diff.shape[0] * diff.shape[1]} pixels differ in the last bits"
