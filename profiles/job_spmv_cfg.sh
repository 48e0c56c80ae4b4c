# SpMV tests + every kernel variant on the chosen configs (default c1 c2 c3 c5)
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "spmv" 2>&1 | tail -4
timeout 400 python profiles/run_configs.py ${@:-c1 c2 c3 c5} > gpurun_out/c.log 2>&1
python - <<PY
import json
d=json.load(open("gpurun_out/configs.json"))
for c,v in d.items():
    for key in ("spmv","spmv_l2_flushed","spmv_l2_resident"):
        if key in v:
            print(c, key, {k: round(x["ms"],4) for k,x in v[key].items() if isinstance(x,dict) and "ms" in x and k in ("auto","sliced","vector1","vector2","vector4","vector8")})
PY
