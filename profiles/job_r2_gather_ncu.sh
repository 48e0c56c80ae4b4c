#!/bin/bash
# Round 2: ncu of the fused gather kernel's compute side (world of one) next to the stream kernel on the same matrix.
mkdir -p gpurun_out
timeout 200 python profiles/r2_gather_ncu.py > gpurun_out/r2_gather_ncu_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/r2_gather_ncu_plain.log; exit 1; }
timeout 500 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'spmv_gather_fused|spmv_stream_kernel' -s 2 -c 4 \
    -o gpurun_out/r2_gather_fused -f python profiles/r2_gather_ncu.py > gpurun_out/r2_gather_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r2_gather_ncu.log
