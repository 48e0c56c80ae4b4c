# real multi-GPU check of the final build: NCCL tests + the sharded bench line; argument: N (default 2)
N=${1:-2}
timeout 300 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -4
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench rc=$?"; tail -c 2500 gpurun_out/bench_n$N.json; tail -3 gpurun_out/bench_n$N.err
