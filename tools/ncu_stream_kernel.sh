#!/bin/bash
# one `ncu --set full` capture (source-level) of the path tracer's stream kernel on the bench job, default build:
#   gpurun --timeout 600 -- 'bash tools/ncu_stream_kernel.sh [tag] [scene] [spp]'
# writes gpurun_out/<tag>_summary.txt (tools/ncu_summary.py) and <tag>_source_hot.txt (tools/ncu_source_hot.py); the .ncu-rep is removed
set -u
TAG=${1:-r2_v8}
SCENE=${2:-wok_teapot_flat}
SPP=${3:-64}
OUT=gpurun_out
mkdir -p $OUT
timeout 60 python tools/pt_time.py $SCENE $SPP > $OUT/${TAG}_time.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_pt_streams -s 1 -c 1 -f -o $OUT/$TAG \
    python tools/pt_once.py $SCENE $SPP 1920 1080 2 > $OUT/${TAG}_ncu.log 2>&1
python tools/ncu_summary.py $OUT/$TAG.ncu-rep $OUT/${TAG}_summary.txt
python tools/ncu_source_hot.py $OUT/$TAG.ncu-rep 90 $OUT/${TAG}_lines.csv > $OUT/${TAG}_source_hot.txt 2>&1
rm -f $OUT/$TAG.ncu-rep
cat $OUT/${TAG}_time.log; head -30 $OUT/${TAG}_summary.txt
