#!/bin/bash
# launch lists: (1) the bench command, our kernels only; (2) conversions / assemblies on configs 1-3
mkdir -p gpurun_out
CMD="python bench.py --gpus 1 --steps 2 --warmup 3 --no-extras"
$CMD > gpurun_out/r2_ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:spl:: -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv $CMD > gpurun_out/r2_ncu_launches.log 2>&1
timeout 600 python profiles/r2_prof_conv.py > gpurun_out/r2_prof_conv_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:spl:: -c 3000 --csv --log-file gpurun_out/r2_launches_conv.csv python profiles/r2_prof_conv.py > gpurun_out/r2_prof_conv_ncu.log 2>&1
cat gpurun_out/r2_prof_conv_plain.log; tail -2 gpurun_out/r2_ncu_launches.log; tail -3 gpurun_out/r2_prof_conv_ncu.log
