#!/bin/bash
# Rebuilds the library on the GPU box with different radix tile shapes and times config 2
# (CSR->CSC, shuffled assembly) and config 3 (assembly, CSR->CSC).
# Usage: bash profiles/tune_radix.sh "IPT,MINBLOCKS[,THREADS] ..."
for v in $1; do
  IFS=, read ipt mb thr <<< "$v"; thr=${thr:-256}
  SPL_EXTRA_NVCC_FLAGS="-DRS_IPT_VALUE=$ipt -DRS_MIN_BLOCKS=$mb -DRS_THREADS_VALUE=$thr" python -m spalinalg_b200.build --force > /tmp/build.log 2>&1 || { echo "build failed $v"; tail -5 /tmp/build.log; continue; }
  python profiles/run_configs.py c2 c3 > /tmp/run.log 2>&1
  python - "$v" <<'PY'
import json, sys
d = json.load(open('gpurun_out/configs.json'))
print(sys.argv[1], 'c2 csr_to_csc ms', round(d['c2']['csr_to_csc']['ms'], 3), 'c2 assembly ms', round(d['c2']['assembly_shuffled']['ms'], 3),
      'c3 assembly ms', round(d['c3']['assembly']['ms'], 3), 'c3 csr_to_csc', round(d['c3']['csr_to_csc']['ms'], 3))
PY
done
