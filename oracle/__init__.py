"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes front end of ``liboracle.so`` (built from ``spl_oracle.c`` /
``orc_impl.inc`` by ``make -C oracle``), the CPU restatement of the reference
(spalinalg v0.0.2) hot path.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this
module; nothing under ``spalinalg_b200/`` does.

All indices are ``uint64`` (Rust ``usize``), values ``float32`` / ``float64``;
the scalar type is taken from the value array's dtype.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build() -> str:
    """Compile liboracle.so (gcc, -ffp-contract=off) if missing or stale."""
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("spl_oracle.c", "orc_impl.inc")]
    if (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        for suf in ("f32", "f64"):
            getattr(_LIB, f"orc_compress_from_coo_{suf}").restype = C.c_size_t
            getattr(_LIB, f"orc_addsub_{suf}").restype = C.c_size_t
            getattr(_LIB, f"orc_csr_mul_{suf}").restype = C.c_size_t
            for name in ("orc_recompress", "orc_neg", "orc_csr_spmv", "orc_expand_to_coo"):
                getattr(_LIB, f"{name}_{suf}").restype = None
        _LIB.orc_validate_compressed.restype = C.c_int
    return _LIB


def _suf(dtype) -> str:
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return "f32"
    if dtype == np.float64:
        return "f64"
    raise TypeError(f"Scalar is implemented for f32 and f64 only (src/scalar.rs:55-57), got {dtype}")


def triplet_dtype(dtype) -> np.dtype:
    """AoS (usize, usize, T) with C layout: 24 bytes for both f32 and f64."""
    return np.dtype({"names": ["row", "col", "val"], "formats": ["<u8", "<u8", np.dtype(dtype)],
                     "offsets": [0, 8, 16], "itemsize": 24})


def make_triplets(rows, cols, vals) -> np.ndarray:
    vals = np.asarray(vals)
    out = np.zeros(len(vals), dtype=triplet_dtype(vals.dtype))
    out["row"] = rows
    out["col"] = cols
    out["val"] = vals
    return out


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _u64(a):
    return np.ascontiguousarray(a, dtype=np.uint64)


def _sz(x):
    return C.c_size_t(int(x))


def compress_from_coo(nrows, ncols, triplets, major="row", dedup=True, dropzero=True):
    """COO -> CSR (major='row') / CSC (major='col'); dedup=dropzero=False is DOK ->."""
    suf = _suf(triplets.dtype["val"])
    n = len(triplets)
    nmajor = nrows if major == "row" else ncols
    ptr = np.zeros(nmajor + 1, dtype=np.uint64)
    ind = np.zeros(max(n, 1), dtype=np.uint64)
    val = np.zeros(max(n, 1), dtype=triplets.dtype["val"])
    trip = np.ascontiguousarray(triplets)
    nnz = getattr(lib(), f"orc_compress_from_coo_{suf}")(
        C.c_int(1 if major == "row" else 0), _sz(nrows), _sz(ncols), _sz(n), _p(trip),
        C.c_int(int(dedup)), C.c_int(int(dropzero)), _p(ptr), _p(ind), _p(val))
    return ptr, ind[:nnz].copy(), val[:nnz].copy()


def recompress(nmajor, nminor, ptr, ind, val):
    """transpose() / CSR<->CSC: arrays of the same entries grouped by the other index."""
    val = np.ascontiguousarray(val)
    suf = _suf(val.dtype)
    ptr, ind = _u64(ptr), _u64(ind)
    nnz = int(ptr[nmajor])
    optr = np.zeros(nminor + 1, dtype=np.uint64)
    oind = np.zeros(max(nnz, 1), dtype=np.uint64)
    oval = np.zeros(max(nnz, 1), dtype=val.dtype)
    getattr(lib(), f"orc_recompress_{suf}")(_sz(nmajor), _sz(nminor), _p(ptr), _p(ind), _p(val),
                                            _p(optr), _p(oind), _p(oval))
    return optr, oind[:nnz].copy(), oval[:nnz].copy()


def addsub(subtract, nmajor, nminor, a, b):
    aptr, aind, aval = _u64(a[0]), _u64(a[1]), np.ascontiguousarray(a[2])
    bptr, bind, bval = _u64(b[0]), _u64(b[1]), np.ascontiguousarray(b[2], dtype=aval.dtype)
    suf = _suf(aval.dtype)
    cap = int(aptr[nmajor]) + int(bptr[nmajor])
    optr = np.zeros(nmajor + 1, dtype=np.uint64)
    oind = np.zeros(max(cap, 1), dtype=np.uint64)
    oval = np.zeros(max(cap, 1), dtype=aval.dtype)
    nnz = getattr(lib(), f"orc_addsub_{suf}")(
        C.c_int(int(subtract)), _sz(nmajor), _sz(nminor), _p(aptr), _p(aind), _p(aval),
        _p(bptr), _p(bind), _p(bval), _p(optr), _p(oind), _p(oval))
    return optr, oind[:nnz].copy(), oval[:nnz].copy()


def csr_mul(an, ak, bn, a, b, cap=None):
    """CSR(A*B); for CSC operands call csr_mul(bn, ak, an, b, a) (see orc_impl.inc).
    cap: known upper bound of nnz(C) -> one pass instead of size query + fill."""
    aptr, aind, aval = _u64(a[0]), _u64(a[1]), np.ascontiguousarray(a[2])
    bptr, bind, bval = _u64(b[0]), _u64(b[1]), np.ascontiguousarray(b[2], dtype=aval.dtype)
    suf = _suf(aval.dtype)
    fn = getattr(lib(), f"orc_csr_mul_{suf}")
    args = (_sz(an), _sz(ak), _sz(bn), _p(aptr), _p(aind), _p(aval), _p(bptr), _p(bind), _p(bval))
    nnz = cap if cap is not None else fn(*args, None, None, None)
    optr = np.zeros(an + 1, dtype=np.uint64)
    oind = np.zeros(max(nnz, 1), dtype=np.uint64)
    oval = np.zeros(max(nnz, 1), dtype=aval.dtype)
    nnz = fn(*args, _p(optr), _p(oind), _p(oval))
    return optr, oind[:nnz].copy(), oval[:nnz].copy()


def neg(val):
    val = np.ascontiguousarray(val)
    out = np.empty_like(val)
    getattr(lib(), f"orc_neg_{_suf(val.dtype)}")(_sz(len(val)), _p(val), _p(out))
    return out


def csr_spmv(nrows, ptr, ind, val, x):
    val = np.ascontiguousarray(val)
    x = np.ascontiguousarray(x, dtype=val.dtype)
    ptr, ind = _u64(ptr), _u64(ind)
    y = np.zeros(nrows, dtype=val.dtype)
    getattr(lib(), f"orc_csr_spmv_{_suf(val.dtype)}")(_sz(nrows), _p(ptr), _p(ind), _p(val), _p(x), _p(y))
    return y


def expand_to_coo(nmajor, ptr, ind, val, major="row"):
    val = np.ascontiguousarray(val)
    ptr, ind = _u64(ptr), _u64(ind)
    out = np.zeros(int(ptr[nmajor]), dtype=triplet_dtype(val.dtype))
    getattr(lib(), f"orc_expand_to_coo_{_suf(val.dtype)}")(
        C.c_int(1 if major == "row" else 0), _sz(nmajor), _p(ptr), _p(ind), _p(val), _p(out))
    return out


def validate_compressed(nrows, ncols, ptr, ind, nval, major="row") -> int:
    """0 if CsrMatrix::new / CscMatrix::new accepts, else ordinal of the failing assert."""
    ptr, ind = _u64(ptr), _u64(ind)
    if len(ptr) == 0:
        return 1 if nrows == 0 else (2 if ncols == 0 else 3)
    return int(lib().orc_validate_compressed(
        C.c_int(1 if major == "row" else 0), _sz(nrows), _sz(ncols), _sz(len(ptr)), _p(ptr),
        _sz(len(ind)), _p(ind), _sz(nval)))
