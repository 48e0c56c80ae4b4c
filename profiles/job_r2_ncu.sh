#!/bin/bash
# Round 2 ncu evidence for the bench's dominant kernel (spmv_stream_kernel on config 5, N = 1):
#   1. the plain command (must exit 0), 2. the launch list with per-launch device times, 3. one --set full capture.
mkdir -p gpurun_out
CMD="python bench.py --gpus 1 --steps 2 --warmup 3 --no-extras"
$CMD > gpurun_out/r2_ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:spl:: -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv $CMD > gpurun_out/r2_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spmv_stream_kernel -s 3 -c 2 -f -o gpurun_out/r2_spmv_stream_c5 $CMD > gpurun_out/r2_ncu_full.log 2>&1
# the same kernel on config 1 (the 80 MB Laplacian) and the hot-column kernel on config 4
python profiles/prof_one.py c1 auto 8 > gpurun_out/r2_prof_c1_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:spmv_stream_kernel -s 2 -c 2 -f -o gpurun_out/r2_spmv_stream_c1 python profiles/prof_one.py c1 auto 8 > gpurun_out/r2_ncu_c1.log 2>&1
RMAT_SCALE=24 python profiles/prof_one.py c4 split 4 > gpurun_out/r2_prof_c4_plain.log 2>&1 &&
RMAT_SCALE=24 ncu --set full --clock-control none --import-source on -k regex:spmv_split_hot_kernel -c 1 -f -o gpurun_out/r2_spmv_split_hot_c4 python profiles/prof_one.py c4 split 4 > gpurun_out/r2_ncu_c4.log 2>&1
tail -2 gpurun_out/r2_ncu_plain.log | head -c 600; echo; tail -2 gpurun_out/r2_ncu_launches.log; tail -2 gpurun_out/r2_ncu_full.log; tail -2 gpurun_out/r2_ncu_c1.log; tail -2 gpurun_out/r2_ncu_c4.log; ls -la gpurun_out/*.ncu-rep
