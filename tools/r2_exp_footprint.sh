# A/B of stream-kernel builds (development tool): librt_b200_<tag>.so variants and the default build timed alternately on one box
#   usage: bash tools/r2_exp_footprint.sh <out name> "<tags, '' = default>" [scenes] [spps] [pytest: 1]
OUT=gpurun_out/${1:-r2_ab}.txt
TAGS=${2:-"_base "}
SCENES=${3:-wok_teapot_flat,inside_tlas}
SPPS=${4:-64,256}
if [ "${5:-0}" = 1 ]; then python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -x -q -m gpu -k "path_tracer or degenerate or 64_frame or full_1080p" 2>&1 | tail -3; fi
: > $OUT
for rep in 1 2; do
for L in $TAGS DEFAULT; do
  [ "$L" = DEFAULT ] && L=""
  echo "== lib$L" >> $OUT
  RT_B200_LIB=cpu-ray-tracer_b200/librt_b200$L.so python tools/pt_time.py $SCENES $SPPS >> $OUT 2>&1
done; done
cat $OUT
