python -m pytest tests/test_gpu_construct.py -x -q -m gpu -s 2>&1 | tail -8
python tools/tlas_build_time.py 1000,5000,20129,40000,56000 2>&1 | tee gpurun_out/r2_tlas_build_time.txt
