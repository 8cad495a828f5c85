python bench.py > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err; echo bench rc=$?
bash tools/ncu_stream_kernel.sh r2_v8b wok_teapot_flat 64 > gpurun_out/r2_ncu_v8b.log 2>&1
bash tools/sanitize.sh > gpurun_out/r2_sanitize.log 2>&1
cat gpurun_out/r2_sanitizer_summary.txt
