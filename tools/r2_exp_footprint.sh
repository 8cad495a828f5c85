# A/B of the stream kernel's code footprint (development tool): parity of the default build, then base / A / default timed alternately
OUT=gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -x -q -m gpu -k "path_tracer or degenerate or 64_frame or full_1080p" 2>&1 | tail -3
: > $OUT/r2_park7.txt
for rep in 1 2; do
for L in _S ""; do
  echo "== lib$L" >> $OUT/r2_park7.txt
  RT_B200_LIB=cpu-ray-tracer_b200/librt_b200$L.so python tools/pt_time.py inside_tlas,instanced_tlas 64,256 >> $OUT/r2_park7.txt 2>&1
done; done
cat $OUT/r2_park7.txt
