# pipelined host-vector product: its tests and the bench's e2e
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pipelined or spmv_random" 2>&1 | tail -3
timeout 200 python bench.py --no-extras > gpurun_out/bench_e2e.json 2> gpurun_out/bench_e2e.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_e2e.json").read().strip().splitlines()[-1])
print("value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), d["e2e"]["launches_per_step"])
PY
