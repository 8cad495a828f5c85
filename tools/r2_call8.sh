OUT=gpurun_out
for N in 8 4 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$N bench.py --gpus $N --steps 5 --warmup 3 > $OUT/r2b_bench_${N}gpu.json 2> $OUT/r2b_bench_${N}gpu.err; echo bench$N rc=$?
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29633 tools/c4_multi.py 256 > $OUT/r2b_c4_8gpu.log 2>&1; tail -1 $OUT/r2b_c4_8gpu.log
python -m pytest tests/test_gpu_multi.py tests/test_gpu_full_size.py -x -q -m gpu 2>&1 | tail -3
for L in "" _fn10 _fn12; do
  echo "== lib$L" >> $OUT/r2_fn_minb.txt
  RT_B200_LIB=cpu-ray-tracer_b200/librt_b200$L.so python tools/c5_l2_window.py 0 >> $OUT/r2_fn_minb.txt 2>&1
done
cat $OUT/r2_fn_minb.txt
python - <<'PY'
import json
for n in (2,4,8):
    d=json.loads(open(f'gpurun_out/r2b_bench_{n}gpu.json').read().strip().splitlines()[-1])
    print(n, round(d['value'],1), round(d['ms_per_step'],2), round(d['e2e']['value'],1), round(d['strong_scaling']['ms'],1), round(d['strong_scaling']['value'],1))
PY
