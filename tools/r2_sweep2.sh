#!/bin/bash
OUT=gpurun_out/r2_sweep2.log
: > $OUT
run() { echo "## $*" >> $OUT; env "$@" timeout 120 python tools/pt_time.py ${SCENES:-wok_teapot_flat} ${SPP:-64} >> $OUT 2>&1; }
for ctas in 4 5 6 7; do for ks in 1 2; do for sm in 0 24; do
run RT_B200_STREAM_CTAS=$ctas RT_B200_STREAM_KEEPSHIFT=$ks RT_B200_STREAM_SMEM_SLOTS=$sm
done; done; done
SPP=256 run RT_B200_STREAM_KEEPSHIFT=1 RT_B200_STREAM_SMEM_SLOTS=0
SPP=256 run RT_B200_STREAM_KEEPSHIFT=1 RT_B200_STREAM_SMEM_SLOTS=24
SPP=256 run RT_B200_STREAM_KEEPSHIFT=1 RT_B200_STREAM_SMEM_SLOTS=0 RT_B200_STREAM_MINB=8
RT_B200_DEBUG=1 run RT_B200_STREAM_KEEPSHIFT=1 RT_B200_STREAM_SMEM_SLOTS=0
cat $OUT
