#!/bin/bash
# A/B of the two builds of the library on the bench scene: the default build (glibc's expf / atan2f /
# acosf restated on the device) against librt_b200_cudamath.so (CUDA's routines).  Run on the GPU box from the repo root:
#   gpurun --timeout 600 -- 'bash tools/ncu_ab_libm.sh'
# 1. timing control without a profiler, both builds, same box;  2. one ncu --set full capture of the second k_pt_streams5 launch of
# each build (source-level, -lineinfo);  3. text summaries + per-source-line hot spots into gpurun_out/ (the .ncu-rep files are
# removed: two of them exceed what gpurun brings back).  Compare the two *_source_hot.txt files line by line.
# build() also compiles librt_b200_glibcexpf.so (glibc expf, CUDA sky routines), librt_b200_glibcsky.so (the reverse) and
# librt_b200_ffexpf.so (float-float expf); when present they are timed too, which splits the cost.
set -u
OUT=gpurun_out
mkdir -p $OUT
CM=$PWD/cpu-ray-tracer_b200/librt_b200_cudamath.so
SCENE=${1:-wok_teapot_flat}
for tag in glibc cuda glibc cuda; do
    if [ $tag = cuda ]; then export RT_B200_LIB=$CM; else unset RT_B200_LIB; fi
    timeout 60 python tools/pt_time.py $SCENE 64 >> $OUT/ab_libm_time_$tag.log 2>&1
done
cat $OUT/ab_libm_time_glibc.log $OUT/ab_libm_time_cuda.log
for tag in glibcexpf glibcsky ffexpf; do
    lib=$PWD/cpu-ray-tracer_b200/librt_b200_$tag.so
    [ -f $lib ] && RT_B200_LIB=$lib timeout 60 python tools/pt_time.py $SCENE 64 2>&1 | sed "s/^/$tag: /" | tee -a $OUT/ab_libm_time_variants.log
done
unset RT_B200_LIB
for tag in glibc cuda; do
    if [ $tag = cuda ]; then export RT_B200_LIB=$CM; else unset RT_B200_LIB; fi
    timeout 280 ncu --set full --clock-control none --import-source on -k regex:k_pt_streams5 -s 1 -c 1 -f -o $OUT/ab_libm_$tag \
        python tools/pt_once.py $SCENE 64 1920 1080 2 > $OUT/ab_libm_ncu_$tag.log 2>&1
    python tools/ncu_summary.py $OUT/ab_libm_$tag.ncu-rep $OUT/ab_libm_${tag}_summary.txt
    python tools/ncu_source_hot.py $OUT/ab_libm_$tag.ncu-rep 80 > $OUT/ab_libm_${tag}_source_hot.txt 2>&1
    rm -f $OUT/ab_libm_$tag.ncu-rep
done
grep -h "gpu__time_duration.sum\|smsp__inst_executed.sum\|thread_inst_executed_per_inst\|issue_active\|stalled_no_instruction\|stalled_math_pipe" $OUT/ab_libm_glibc_summary.txt $OUT/ab_libm_cuda_summary.txt
