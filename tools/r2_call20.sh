# round 2, call 20: automatic 7 / 8 CTAs per SM - parity subset with the default choice and with 8 forced everywhere, then timing
T="tests/test_gpu_parity.py tests/test_gpu_full_size.py -x -q -m gpu -k"
python -m pytest $T "path_tracer or degenerate or 64_frame or full_1080p" 2>&1 | tail -2
RT_B200_STREAM_MINB=8 python -m pytest $T "path_tracer or 64_frame or full_1080p" 2>&1 | tail -2
python tools/pt_time.py wok_teapot_flat,inside_tlas,instanced_tlas 64,256 2>&1 | tee gpurun_out/r2_minb_auto.txt
