#!/bin/bash
# Round 2: full GPU tests, smoke, the bench line at N = 1 and the reference arm, as the driver runs them.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r2_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/r2_smoke.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?" >> gpurun_out/r2_bench_n1.err
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo "ref rc=$?" >> gpurun_out/r2_bench_ref.err
tail -12 gpurun_out/r2_tests.log; cat gpurun_out/r2_smoke.log; tail -5 gpurun_out/r2_bench_n1.err; head -c 1500 gpurun_out/r2_bench_n1.json; echo; tail -3 gpurun_out/r2_bench_ref.err; head -c 600 gpurun_out/r2_bench_ref.json
