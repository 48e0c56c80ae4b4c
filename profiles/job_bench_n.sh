# bench line at N GPUs (one rank per GPU, torchrun as the driver launches it)
N=${1:-8}
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench rc=$?"; tail -c 1800 gpurun_out/bench_n$N.json
