#!/bin/bash
# Round 2, second GPU job: GPU tests (full-size bit-exact C3/C4 added), SpMV sweeps of the stream kernel v2.
mkdir -p gpurun_out
free -g > gpurun_out/r2_host.txt 2>&1; nproc >> gpurun_out/r2_host.txt
timeout 1200 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/r2_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests.log
timeout 900 python profiles/r2_spmv_sweep.py c1 c2 c5 c4 > gpurun_out/r2_sweep.log 2>&1; echo "rc=$?" >> gpurun_out/r2_sweep.log
cat gpurun_out/r2_host.txt; tail -25 gpurun_out/r2_tests.log; grep -c . gpurun_out/r2_sweep.log
