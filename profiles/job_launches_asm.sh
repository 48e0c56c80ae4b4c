# launch list (ncu, gpu__time_duration) of one config-3 assembly; argument: nrows (default 1e7)
N=${1:-1e7}
python profiles/prof_asm.py $N 3 > gpurun_out/plain_asm.log 2>&1 && cat gpurun_out/plain_asm.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_asm.csv python profiles/prof_asm.py $N 1 > gpurun_out/ncu_asm.log 2>&1
python profiles/summarize_launches.py gpurun_out/launches_asm.csv | grep -v "native::\|at_cuda"
