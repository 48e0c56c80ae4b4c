#!/bin/bash
# Round 2, N GPUs: the sharded path against the oracle (pytest, 2 ranks) and the bench line at N.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo.txt 2>&1
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/r2_tests_multi.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests_multi.log
tail -15 gpurun_out/r2_tests_multi.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "bench rc=$?" >> gpurun_out/r2_bench_n$N.err
tail -12 gpurun_out/r2_bench_n$N.err; head -c 3000 gpurun_out/r2_bench_n$N.json
