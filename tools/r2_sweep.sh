#!/bin/bash
# round-2 A/B sweep of the stream kernels (development tool): gpurun -- 'bash tools/r2_sweep.sh'
OUT=gpurun_out/r2_sweep.log
: > $OUT
run() { echo "## $*" >> $OUT; env "$@" timeout 120 python tools/pt_time.py ${SCENES:-wok_teapot_flat} ${SPP:-64,256} >> $OUT 2>&1; }
run RT_B200_STREAM_KERNEL=5
run RT_B200_STREAM_KERNEL=8
run RT_B200_STREAM_KERNEL=8 RT_B200_STREAM_MINB=8
run RT_B200_STREAM_KERNEL=8 RT_B200_STREAM_SMEM_SLOTS=0
run RT_B200_STREAM_KERNEL=8 RT_B200_STREAM_SMEM_SLOTS=32
run RT_B200_STREAM_KERNEL=8 RT_B200_STREAM_SMEM_SLOTS=0 RT_B200_STREAM_MINB=8
run RT_B200_STREAM_KERNEL=8 RT_B200_ORDERED_FRAMES=0
run RT_B200_STREAM_KERNEL=8 RT_B200_STREAM_KEEPSHIFT=1
run RT_B200_STREAM_KERNEL=8 RT_B200_STREAM_KEEPSHIFT=3
SCENES=inside_tlas,instanced_tlas SPP=64 run RT_B200_STREAM_KERNEL=5
SCENES=inside_tlas,instanced_tlas SPP=64 run RT_B200_STREAM_KERNEL=8
SCENES=inside_tlas,instanced_tlas SPP=64 run RT_B200_STREAM_KERNEL=8 RT_B200_STREAM_MINB=8
cat $OUT
