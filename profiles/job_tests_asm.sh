timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 200 python profiles/run_configs.py c1 c2 c3 > gpurun_out/c.log 2>&1
python - <<PY
import json
d=json.load(open("gpurun_out/configs.json"))
print("c1", {k: round(v["ms"],3) for k,v in d["c1"].items() if k.startswith("assembly")})
print("c2", round(d["c2"]["assembly_shuffled"]["ms"],3), "c3", round(d["c3"]["assembly"]["ms"],3))
PY
