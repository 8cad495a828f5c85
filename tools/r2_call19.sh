# round 2, call 19: the driver's round-end sequence on the final build + 8 CTAs/SM (64 registers) against the default 7
bash tools/r2_final.sh
OUT=gpurun_out/r2_minb8.txt
: > $OUT
for rep in 1 2; do
for M in 7 8; do
  echo "== RT_B200_STREAM_MINB=$M" >> $OUT
  RT_B200_STREAM_MINB=$M python tools/pt_time.py wok_teapot_flat,inside_tlas,instanced_tlas 64,256 >> $OUT 2>&1
done; done
cat $OUT
