#!/bin/bash
OUT=gpurun_out/r2_sweep5.log
: > $OUT
run() { echo "## $*" >> $OUT; env "$@" timeout 120 python tools/pt_time.py ${SCENES:-wok_teapot_flat} ${SPP:-64} >> $OUT 2>&1; }
run RT_B200_STREAM_CRIT_THETA=0
for th in 0.9 0.8 0.7; do for ln in 4 8 16; do
run RT_B200_STREAM_CRIT_THETA=$th RT_B200_STREAM_CRIT_LANES=$ln
done; done
run RT_B200_STREAM_CRIT_THETA=0
cat $OUT
