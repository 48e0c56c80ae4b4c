#!/bin/bash
# launch list (per-kernel device times) of the conversions and assemblies on configs 1-3
mkdir -p gpurun_out
timeout 600 python profiles/r2_prof_conv.py > gpurun_out/r2_prof_conv_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2_launches_conv.csv python profiles/r2_prof_conv.py > gpurun_out/r2_prof_conv_ncu.log 2>&1
cat gpurun_out/r2_prof_conv_plain.log; tail -3 gpurun_out/r2_prof_conv_ncu.log
