#!/bin/bash
OUT=gpurun_out/r2_sweep3.log
: > $OUT
run() { echo "## $*" >> $OUT; env "$@" timeout 120 python tools/pt_time.py ${SCENES:-wok_teapot_flat} ${SPP:-64,256} >> $OUT 2>&1; }
run RT_B200_STREAM_KEEPSHIFT=1 RT_B200_STREAM_SMEM_SLOTS=0
run RT_B200_STREAM_KEEPSHIFT=2 RT_B200_STREAM_SMEM_SLOTS=0
run RT_B200_STREAM_KEEPSHIFT=1 RT_B200_STREAM_SMEM_SLOTS=24
run RT_B200_STREAM_KEEPSHIFT=1 RT_B200_STREAM_SMEM_SLOTS=0 RT_B200_STREAM_MINB=8
SCENES=inside_tlas,instanced_tlas SPP=64 run RT_B200_STREAM_KEEPSHIFT=1 RT_B200_STREAM_SMEM_SLOTS=0
SCENES=inside_tlas,instanced_tlas SPP=64 run RT_B200_STREAM_KEEPSHIFT=1 RT_B200_STREAM_SMEM_SLOTS=32
SCENES=inside_tlas,instanced_tlas SPP=64 run RT_B200_STREAM_KEEPSHIFT=2 RT_B200_STREAM_SMEM_SLOTS=0
cat $OUT
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
