#!/bin/bash
# Round 2, third GPU job: stream kernel v3 (pointers through the TMA ring) - parity tests of the SpMV kernels, sweeps.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_dist.py -m gpu -x -q -k "spmv or stream or tile_boundary or dist or peer or sharded or values_mut" > gpurun_out/r2_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests.log
timeout 900 python profiles/r2_spmv_sweep.py c1 c2 c5 c3 > gpurun_out/r2_sweep.log 2>&1; echo "rc=$?" >> gpurun_out/r2_sweep.log
tail -8 gpurun_out/r2_tests.log; grep -c . gpurun_out/r2_sweep.log
