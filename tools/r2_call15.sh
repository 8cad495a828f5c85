#!/bin/bash
# round 2, call 15: ncu capture of the current default build (flat + TLAS scene) and a knob sweep of the TLAS stream kernel
bash tools/ncu_stream_kernel.sh r2_v8b_default wok_teapot_flat 64 > gpurun_out/r2_call15_ncu_flat.log 2>&1
bash tools/ncu_stream_kernel.sh r2_v8b_tlas inside_tlas 64 > gpurun_out/r2_call15_ncu_tlas.log 2>&1
OUT=gpurun_out/r2_tlas_knobs.txt
: > $OUT
run() { echo "## $*" >> $OUT; env "$@" timeout 120 python tools/pt_time.py inside_tlas,instanced_tlas 64 >> $OUT 2>&1; }
run RT_B200_STREAM_KEEPSHIFT=1
run RT_B200_STREAM_KEEPSHIFT=2
run RT_B200_STREAM_MINB=8
run RT_B200_STREAM_MINB=8 RT_B200_STREAM_KEEPSHIFT=2
run RT_B200_STREAM_SMEM_SLOTS=0
run RT_B200_STREAM_CTAS=6
run RT_B200_STREAM_CTAS=5
run RT_B200_STREAM_KEEPSHIFT=1
cat $OUT
