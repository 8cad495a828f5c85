"""Host-side pieces of bench.py that need no GPU: the committed-profile readers, the roofline's algorithmic work per ray from
the oracle's counters, the peak table, and the failure path of the child run that times the CUDA-libm build."""
import argparse
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_profile_readers_return_numbers():
    prof = bench.kernel_profile()
    assert prof is None or (os.path.exists(os.path.join(ROOT, prof["source"])) and prof.get("warp_instructions_per_launch", 1) > 0)
    peaks, kind = bench.measured_peaks()
    assert peaks["hbm_gbs"] > 1000 and isinstance(kind, str)


def test_algorithmic_bytes_per_ray_from_oracle_counters(flat_scenes):
    """64 I + 52 T + 64 B + 48 bytes per ray, I / T / B from the oracle's traversal counters"""
    w = bench.oracle_work_per_ray(flat_scenes("golden_file"), 64, 36, frames=1)
    assert w["interior_visits"] > 0 and w["tri_tests"] > 0 and w["blas_entries"] == 0
    assert abs(w["bytes_per_ray"] - (64 * w["interior_visits"] + 52 * w["tri_tests"] + 64 * w["blas_entries"] + 48)) < 1e-6
    w = bench.oracle_work_per_ray(flat_scenes("golden_tlas"), 64, 36, frames=1)
    assert w["blas_entries"] > 0


def test_both_arms_describe_the_job_with_the_same_config_keys():
    a = bench.config_block("w", 1920, 1080, 64, 1)
    b = bench.config_block("w", 1920, 1080, 64, 8)
    assert set(a) == set(b) and a["total_spp"] == 64 and b["total_spp"] == 512 and "tiles" in b["sharding"]


def test_alt_build_child_failure_is_reported_not_raised(monkeypatch, tmp_path):
    args = argparse.Namespace(steps=1, warmup=1, width=64, height=36, spp=1)
    monkeypatch.delenv("RT_B200_LIB", raising=False)
    res = bench.alt_build_line(args)
    # no GPU here: either the CUDA-libm build is absent (None) or the child run fails and says so
    assert res is None or res["value"] is not None or "failed" in res["note"]  # (a GPU box returns the timed line)
    monkeypatch.setenv("RT_B200_LIB", "/nonexistent.so")
    assert bench.alt_build_line(args) is None  # never recurses when a library override is already active


def test_reference_arm_prints_the_contract_line():
    """--impl reference on a tiny frame: the reference's own CPU code (oracle/_ref) or the oracle port, whichever is present"""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--width", "64", "--height", "36"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["unit"] == "Mrays/s"
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["cpu_baseline"]["kind"] in ("reference", "port")
    assert set(line["config"]) == set(bench.config_block("w", 64, 36, 64, 1))  # the GPU arm's keys
    if line["cpu_baseline"]["kind"] == "reference":
        assert "strict" in line["cpu_baseline"]["builds"]  # both builds of the reference are reported


def test_roofline_issue_figures_come_from_the_default_build_capture():
    """profiles/ holds ncu captures of several instances of the stream kernel (flat default build, TLAS): bench.py's roofline.issue must
    read the one of the timed configuration (flat scene, default build)"""
    import bench
    prof = bench.kernel_profile()
    assert prof is not None and "default" in prof["source"]
    assert prof["kernel"].startswith("void k_pt_streams8<0,")            # TLAS = 0
    assert 2e10 < prof["warp_instructions_per_launch"] < 6e10              # 64 spp of the bench scene: ~143 warp instructions per ray
    assert prof["dram_bytes_per_launch"] and prof["dram_bytes_per_launch"] < 1e10
