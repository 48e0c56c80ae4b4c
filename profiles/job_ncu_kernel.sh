# one `ncu --set full` capture of the first launch of a kernel: args <kernel regex> <out name> <python script + args...>
K=$1; OUT=$2; shift 2
timeout 500 ncu --set full --clock-control none --import-source on -k regex:$K -c 1 -o gpurun_out/$OUT -f python "$@" > gpurun_out/ncu_$OUT.log 2>&1
tail -3 gpurun_out/ncu_$OUT.log
