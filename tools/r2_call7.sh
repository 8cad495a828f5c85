OUT=gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -4
for L in 32 24 16 8; do
  echo "== lanes per warp $L" >> $OUT/r2_force_lanes.txt
  RT_B200_DEBUG=1 RT_B200_STREAM_FORCE_LANES=$L python tools/pt_time.py wok_teapot_flat 64 >> $OUT/r2_force_lanes.txt 2>&1
done
cat $OUT/r2_force_lanes.txt
python bench.py > $OUT/r2_bench4.json 2> $OUT/r2_bench4.err; echo bench rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/r2_v8_launches.csv python bench.py --steps 2 --warmup 1 --no-extra > $OUT/r2_ncu_launches.log 2>&1
bash tools/ncu_stream_kernel.sh r2_v8_tlas inside_tlas 64 > $OUT/r2_ncu_tlas.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_find_nearest_persistent -s 2 -c 1 -f -o $OUT/r2_c5_scattered python tools/c5_once.py scattered 3 > $OUT/r2_c5_ncu.log 2>&1
python tools/ncu_summary.py $OUT/r2_c5_scattered.ncu-rep $OUT/r2_c5_scattered_summary.txt
python tools/ncu_source_hot.py $OUT/r2_c5_scattered.ncu-rep 40 > $OUT/r2_c5_scattered_source_hot.txt 2>&1
rm -f $OUT/r2_c5_scattered.ncu-rep
head -30 $OUT/r2_c5_scattered_summary.txt
