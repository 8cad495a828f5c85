# round 2, call 17 (8 GPUs): bench --gpus 8 and --gpus 4 with the final build
nvidia-smi -L | wc -l
for N in 8 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N --steps 5 --warmup 3 --no-extra > gpurun_out/r2c_bench_${N}gpu.json 2> gpurun_out/r2c_bench_${N}gpu.err; echo bench$N rc=$?
head -c 300 gpurun_out/r2c_bench_${N}gpu.json; echo
done
