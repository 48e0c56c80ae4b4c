#!/bin/bash
# Compare the SpMV kernels on configs 2 and 1 (device-timed; prints value GB/s, ms/step, roofline frac)
[ "$1" = "--test" ] && timeout 300 python -m pytest tests -m gpu -x -q -k spmv 2>&1 | tail -15
run() { timeout 200 python bench.py --steps 100 --warmup 5 --no-extras "$@" > /tmp/o.log 2>&1; python - "$*" <<'PY'
import sys, json
lines = open('/tmp/o.log').read().strip().splitlines()
try:
    d = json.loads(lines[-1]); print(sys.argv[1], round(d['value']), round(d['ms_per_step'], 4), round(d['roofline']['frac'], 3), d['config']['spmv_kernel'])
except Exception as e:
    print(sys.argv[1], 'FAILED', lines[-3:])
PY
}
for k in ${KERNELS2:-stream1024 merge vector4 vector8 vector16 auto}; do run --kernel $k; done
for k in ${KERNELS1:-stream256 merge vector2 vector4 vector8 auto}; do run --workload laplace2d_1024_f64 --kernel $k; done
