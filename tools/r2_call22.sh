# round 2, call 22 (2 GPUs): multi-GPU tests + a short bench with the final build
python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 --no-extra > gpurun_out/r2f_bench_2gpu.json 2> gpurun_out/r2f_bench_2gpu.err; echo bench2 rc=$?
head -c 260 gpurun_out/r2f_bench_2gpu.json
