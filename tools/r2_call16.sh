# round 2, call 16 (2 GPUs): multi-GPU tests + bench --gpus 2 with the final build, --no-extra to keep it short
nvidia-smi -L | wc -l
python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --no-extra > gpurun_out/r2c_bench_2gpu.json 2> gpurun_out/r2c_bench_2gpu.err; echo bench2 rc=$?
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2c_bench_2gpu_ref.json 2> gpurun_out/r2c_bench_2gpu_ref.err; echo ref2 rc=$?
head -c 600 gpurun_out/r2c_bench_2gpu.json; echo; tail -c 300 gpurun_out/r2c_bench_2gpu.err
