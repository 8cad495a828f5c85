"""Checks of experimental code that no default path uses yet (named to run last: a failure here must not hide the parity suite).

rt_glibc_expf_ff (csrc/rt_glibc_math.cuh, build variant RT_B200_EXPF_FF): glibc's expf bits through float-float arithmetic.
The host comparison with libm over all 2^32 arguments is part of tests/tools/glibc_math_check.c; here the DEVICE evaluation is
compared with the host libm."""
import numpy as np
import pytest

from cpu_ray_tracer_b200 import abi
from test_glibc_math import checker, host_eval, require_fma_libm, same_bits, sweep_arguments  # noqa: F401


def test_float_float_expf_on_the_host(checker):
    _, L = checker
    if not hasattr(L, "restated_eval"):
        pytest.skip("checker library without array entry points")
    _, xe, _, _, _ = sweep_arguments()
    x = xe[::8]
    L.restated_eval.argtypes = L.libm_eval.argtypes
    assert same_bits(host_eval(L.restated_eval, 4, x), host_eval(L.libm_eval, 0, x)).all()


@pytest.mark.gpu
def test_device_float_float_expf_equals_host_libm(checker):
    from cpu_ray_tracer_b200 import api
    _, L = checker
    require_fma_libm(L)
    _, xe, _, _, _ = sweep_arguments()
    got, ref = api.eval_shading_math(abi.RT_MATH_EXPF_FF, xe), host_eval(L.libm_eval, 0, xe)
    bad = ~same_bits(got, ref)
    assert not bad.any(), f"{int(bad.sum())} of {xe.size} device results differ from libm, first at {xe[bad][:4]}"
