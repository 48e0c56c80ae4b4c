#!/bin/bash
# Round 2, first GPU job: GPU tests (incl. the new stream kernel), DSMEM gather microbenchmark, SpMV sweeps.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_gpu.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests.log
timeout 120 ./profiles/micro/dsmem_gather > gpurun_out/r2_dsmem.txt 2>&1; echo "rc=$?" >> gpurun_out/r2_dsmem.txt
timeout 900 python profiles/r2_spmv_sweep.py c1 c2 c5 c3 c4 > gpurun_out/r2_sweep.log 2>&1; echo "rc=$?" >> gpurun_out/r2_sweep.log
tail -5 gpurun_out/r2_tests.log; cat gpurun_out/r2_dsmem.txt; tail -30 gpurun_out/r2_sweep.log
