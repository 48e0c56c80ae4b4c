# final artefacts of the round on one GPU: full GPU test suite, smoke, bench line, ncu launch list of the bench,
# all five configs
timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 300 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>/dev/null; echo "ref rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 20 --warmup 3 --no-extras > gpurun_out/ncu_bench.log 2>&1
python profiles/summarize_launches.py gpurun_out/launches_r1.csv | head -8
timeout 600 python profiles/run_configs.py c1 c2 c3 c4 c5 > gpurun_out/configs.log 2>&1; echo "configs rc=$?"
