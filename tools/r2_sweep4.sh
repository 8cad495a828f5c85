#!/bin/bash
OUT=gpurun_out/r2_sweep4.log
: > $OUT
run() { echo "## $*" >> $OUT; env "$@" timeout 120 python tools/pt_time.py ${SCENES:-wok_teapot_flat} ${SPP:-64,256} >> $OUT 2>&1; }
for rep in 1 2; do
run RT_B200_STREAM_FASTNODE=1
run RT_B200_STREAM_FASTNODE=0
done
run RT_B200_STREAM_FASTNODE=1 RT_B200_STREAM_KEEPSHIFT=8
run RT_B200_STREAM_FASTNODE=0 RT_B200_STREAM_KEEPSHIFT=8
SCENES=inside_tlas,instanced_tlas SPP=64 run RT_B200_STREAM_FASTNODE=1
SCENES=inside_tlas,instanced_tlas SPP=64 run RT_B200_STREAM_FASTNODE=0
cat $OUT
