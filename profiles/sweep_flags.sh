#!/bin/bash
# Rebuilds the library on the GPU box with each set of extra nvcc flags and runs a command.
# Usage: bash profiles/sweep_flags.sh "<cmd>" "<flags A>" "<flags B>" ...
cmd=$1; shift
for f in "$@"; do
  SPL_EXTRA_NVCC_FLAGS="$f" python -m spalinalg_b200.build --force > /tmp/build.log 2>&1 || { echo "build failed: $f"; tail -5 /tmp/build.log; continue; }
  echo "=== flags: $f"
  eval "$cmd"
done
