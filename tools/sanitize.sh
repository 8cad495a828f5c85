#!/bin/bash
# compute-sanitizer over the hot path (the reference's only sanitizer hook is the MSVC debug runtime):
#   gpurun --timeout 1500 -- 'bash tools/sanitize.sh'
# memcheck + racecheck + initcheck over __graft_entry__.smoke() (FindNearest / IsOccluded batches, 2-frame path-traced job,
# one Whitted frame on the TLAS golden scene), memcheck + racecheck over one TLAS parity test of each kind.
# Writes gpurun_out/r2_sanitizer_summary.txt (one line per run: tool, target, exit code, ERROR SUMMARY line).
set -u
OUT=gpurun_out
mkdir -p $OUT
SUM=$OUT/r2_sanitizer_summary.txt
: > $SUM
CS=/usr/local/cuda/bin/compute-sanitizer
run() {
    tool=$1; shift; tag=$1; shift
    log=$OUT/sanitize_${tool}_${tag}.log
    timeout 300 $CS --tool $tool --error-exitcode 9 --launch-timeout 0 "$@" > $log 2>&1
    rc=$?
    echo "tool=$tool target=$tag rc=$rc :: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $log | tail -1) :: $(grep -E 'passed|failed|smoke ok' $log | tail -1)" >> $SUM
}
T1='tests/test_gpu_parity.py::test_find_nearest_primary_bit_exact[golden_tlas]'
T2='tests/test_gpu_parity.py::test_secondary_rays_and_occlusion[golden_tlas]'
T3='tests/test_gpu_parity.py::test_path_tracer_vs_oracle_reference_rng[golden_tlas-streams]'
T4='tests/test_gpu_parity.py::test_path_tracer_vs_oracle_reference_rng[golden_tlas-wavefront]'
T5='tests/test_gpu_parity.py::test_whitted_vs_oracle[golden_tlas]'
for tool in memcheck racecheck initcheck; do
    run $tool smoke python __graft_entry__.py smoke
done
for tool in memcheck racecheck; do
    run $tool tlas_parity python -m pytest -q -x -m gpu "$T1" "$T2" "$T3" "$T4" "$T5"
done
run memcheck multi_gpu python -m pytest -q -x -m gpu tests/test_gpu_multi.py -k "bit_identical and golden_tlas"
cat $SUM
