# full GPU suite + the one-GPU bench line
timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 300 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n1.json").read().strip().splitlines()[-1])
print("value", round(d["value"],1), "frac", round(d["roofline"]["frac"],3), "e2e", d["e2e"], "cold", d.get("e2e_cold",{}).get("value"))
PY
tail -3 gpurun_out/bench_n1.err
