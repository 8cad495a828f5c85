nvidia-smi -L
python -m pytest tests/test_gpu_multi.py tests/test_cpp_adapters.py -x -q -m gpu -k "multi or all_gpus or ipc or shard" 2>&1 | tail -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err; echo bench2 rc=$?
tail -c 600 gpurun_out/r2_bench_2gpu.err
python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_2gpu.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['strong_scaling'])
"
