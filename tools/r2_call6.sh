nvidia-smi -L | wc -l
python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -3
for N in 8 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err; echo bench$N rc=$?
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tools/c4_multi.py 256 > gpurun_out/r2_c4_8gpu.log 2>&1; tail -2 gpurun_out/r2_c4_8gpu.log
C4_BUILD=host C4_COMBINE=reduce python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 tools/c4_multi.py 256 > gpurun_out/r2_c4_8gpu_r1path.log 2>&1; tail -2 gpurun_out/r2_c4_8gpu_r1path.log
python tools/c4_multi.py 32 > gpurun_out/r2_c4_1gpu_32spp.log 2>&1; tail -1 gpurun_out/r2_c4_1gpu_32spp.log
