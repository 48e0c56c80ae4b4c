"""Builds libspalinalg_b200.so (sm_100a only) in-tree with nvcc.  No torch involved: the library
is a plain C-ABI shared object (include/spl.h).  Run: python -m spalinalg_b200.build [--force]"""
from __future__ import annotations

import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libspalinalg_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

# --fmad=false: the reference (Rust) never contracts a*b+c; products and sums round separately.
# No --use_fast_math, no FTZ: subnormal sums must survive (hard part 1).
FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
         "--fmad=false", "--ftz=false", "--prec-div=true", "--prec-sqrt=true",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-Xptxas", "-v"]
FLAGS += os.environ.get("SPL_EXTRA_NVCC_FLAGS", "").split()      # tuning experiments (-DRS_IPT_VALUE=8 ...)


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers_mtime():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    hs.append(os.path.join(HERE, "..", "include", "spl.h"))
    return max(os.path.getmtime(h) for h in hs)


def _compile(src: str, force: bool) -> str:
    obj = os.path.join(OBJ, src[:-3] + ".o")
    s = os.path.join(CSRC, src)
    if (not force and os.path.exists(obj)
            and os.path.getmtime(obj) > max(os.path.getmtime(s), _headers_mtime())):
        return obj
    log = subprocess.run([NVCC, *FLAGS, "-c", s, "-o", obj], capture_output=True, text=True)
    with open(obj + ".log", "w") as f:      # ptxas -v: registers / spills / shared memory per kernel
        f.write(log.stdout + log.stderr)
    if log.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}:\n{log.stdout}\n{log.stderr}")
    return obj


def build(force: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    with cf.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(lambda s: _compile(s, force), _sources()))
    if force or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        # default visibility only for the extern "C" entry points (SPL_EXPORT in api.cu)
        cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
