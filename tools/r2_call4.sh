python -m pytest tests -x -q -m gpu 2>&1 | tail -5
python bench.py > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err; echo bench rc=$?
python tools/c5_l2_window.py 0,32,64,96 > gpurun_out/r2_c5_l2_window.txt 2>&1; cat gpurun_out/r2_c5_l2_window.txt
