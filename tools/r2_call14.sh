python -m pytest tests/test_gpu_parity.py tests/test_gpu_construct.py -x -q -m gpu -k "device_pointer or tiny or tlas_builder" 2>&1 | tail -3
python tools/c5_l2_window.py 0 2>&1 | tee gpurun_out/r2_c5_hint.txt
