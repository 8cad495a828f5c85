# the driver's round-end sequence, run once by hand: GPU tests, smoke, both bench arms
OUT=gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py --impl reference > $OUT/r2_final_bench_reference.json 2> $OUT/r2_final_bench_reference.err; echo ref rc=$?
python bench.py > $OUT/r2_final_bench.json 2> $OUT/r2_final_bench.err; echo bench rc=$?
tail -c 400 $OUT/r2_final_bench.err
