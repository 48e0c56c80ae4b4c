"""Host-side mirror of the reference's public API for the hot path (same names, argument
meaning and error behaviour), on top of the C ABI in include/spl.h.

Reference (Rust crate spalinalg v0.0.2, paths under /root/reference):
  CooMatrix  src/coo.rs:52-57     builder, host SoA triplets (API shell; input of assembly)
  DokMatrix  src/dok.rs:53-58     builder, host dict (API shell)
  CsrMatrix  src/csr.rs:65-72     device resident, immutable structure
  CscMatrix  src/csc.rs:65-72     device resident, immutable structure
Rust `From` impls become `from_*` constructors, `impl Add/Sub/Mul/Neg for &M` become the Python
operators on the matrix objects, `assert!` panics become `Panic` (an AssertionError).  CSR/CSC
data lives in HBM; `rowptr()/colind()/values()` download (and cache) host copies with `usize`
(uint64) indices, exactly sized.  There is no CPU fallback: every CSR/CSC operation goes through
libspalinalg_b200.so and raises if it is missing or no CUDA device is present.
"""
from __future__ import annotations

import ctypes as C
import threading
from typing import Iterable, Iterator, Optional, Tuple

import numpy as np

from . import _capi as capi


class Panic(AssertionError):
    """The reference panics (assert!/assert_eq!); the mirror raises this."""


class DeviceError(RuntimeError):
    """CUDA / resource failure reported by the library (no reference equivalent)."""


def _dtype_code(dtype) -> int:
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return capi.SPL_F32
    if dtype == np.float64:
        return capi.SPL_F64
    raise TypeError(f"Scalar is implemented for f32 and f64 only (src/scalar.rs:55-57), got {dtype}")


_NP = {capi.SPL_F32: np.float32, capi.SPL_F64: np.float64}


# --------------------------------------------------------------------------- context
class Context:
    """One spl_ctx: a device, a stream, the last error.  One per host thread."""

    def __init__(self, device: int = 0, stream: Optional[int] = None):
        self._lib = capi.load()
        h = C.c_void_p()
        st = self._lib.spl_ctx_create(int(device), C.c_void_p(stream) if stream else None, C.byref(h))
        if st != capi.SPL_OK:
            raise DeviceError(
                f"spl_ctx_create(device={device}) failed with {capi.STATUS_NAMES.get(st, st)}: "
                "a CUDA device is required, there is no CPU fallback")
        self._h = h
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.spl_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, status: int):
        if status == capi.SPL_OK:
            return
        msg = self._lib.spl_last_error(self._h).decode(errors="replace")
        name = capi.STATUS_NAMES.get(status, str(status))
        if status in (capi.SPL_ERR_SHAPE, capi.SPL_ERR_INVALID, capi.SPL_ERR_ARG):
            raise Panic(f"{name}: {msg}")
        raise DeviceError(f"{name}: {msg}")

    def invalid_reason(self) -> int:
        return int(self._lib.spl_invalid_reason(self._h))

    def sync(self):
        self.check(self._lib.spl_ctx_sync(self._h))

    def trim(self):
        """Synchronise and hand the freed device memory in the library's pool back to the driver."""
        self.check(self._lib.spl_ctx_trim(self._h))

    def launch_count(self) -> int:
        return int(self._lib.spl_launch_count(self._h))


_tls = threading.local()


def default_context() -> Context:
    ctx = getattr(_tls, "ctx", None)
    if ctx is None:
        ctx = Context(0)
        _tls.ctx = ctx
    return ctx


def set_default_context(ctx: Optional[Context]):
    _tls.ctx = ctx


class _PinnedBlock:
    """Owner of one spl_host_alloc block; freed when the last numpy view of it goes away."""

    def __init__(self, nbytes: int):
        self._lib = capi.load()
        p = C.c_void_p()
        status = self._lib.spl_host_alloc(max(int(nbytes), 1), C.byref(p))
        if status != capi.SPL_OK or not p.value:
            raise DeviceError("spl_host_alloc failed: no CUDA device, or out of pinnable memory")
        self.addr = p.value

    def __del__(self):
        if getattr(self, "addr", None):
            self._lib.spl_host_free(C.c_void_p(self.addr))
            self.addr = None


def pinned_empty(n: int, dtype=np.float64) -> np.ndarray:
    """A 1-D numpy array of n values in page-locked host memory (spl_host_alloc).  Passed to matvec (x, and
    out=) it lets spl_spmv_host pipeline the upload, the product and the download over row chunks; ordinary
    arrays give the same result through one staged copy each way."""
    dt = np.dtype(dtype)
    block = _PinnedBlock(int(n) * dt.itemsize)
    buf = (C.c_char * (int(n) * dt.itemsize)).from_address(block.addr)
    buf._owner = block                                        # numpy keeps buf (its base) alive, buf keeps the block
    return np.frombuffer(buf, dtype=dt, count=int(n))


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


# --------------------------------------------------------------------------- COO (host shell)
class CooMatrix:
    """Coordinate matrix: insertion-ordered triplets, duplicates allowed (src/coo.rs:52-57).
    Stored as SoA numpy arrays so the assembly call can hand them to the device as they are."""

    def __init__(self, nrows: int, ncols: int, dtype=np.float64):      # CooMatrix::new, coo.rs:104-112
        if not nrows > 0:
            raise Panic("assertion failed: nrows > 0")
        if not ncols > 0:
            raise Panic("assertion failed: ncols > 0")
        self._nrows, self._ncols = int(nrows), int(ncols)
        self._dtype = np.dtype(dtype)
        _dtype_code(self._dtype)
        self._row = np.empty(0, np.uint64)
        self._col = np.empty(0, np.uint64)
        self._val = np.empty(0, self._dtype)
        self._len = 0

    # constructors ------------------------------------------------------------
    @classmethod
    def new(cls, nrows, ncols, dtype=np.float64):
        return cls(nrows, ncols, dtype)

    @classmethod
    def eye(cls, size, dtype=np.float64):                               # coo.rs:127-139
        if not size > 0:
            raise Panic("assertion failed: size > 0")
        idx = np.arange(size, dtype=np.uint64)
        return cls.with_triplets(size, size, idx, idx, np.ones(size, dtype))

    @classmethod
    def with_capacity(cls, nrows, ncols, capacity, dtype=np.float64):   # coo.rs:162-170
        m = cls(nrows, ncols, dtype)
        m._reserve(int(capacity))
        return m

    @classmethod
    def with_entries(cls, nrows, ncols, entries: Iterable[Tuple[int, int, float]], dtype=np.float64):
        ents = list(entries)                                            # coo.rs:204-220
        r = np.array([e[0] for e in ents], dtype=np.uint64)
        c = np.array([e[1] for e in ents], dtype=np.uint64)
        v = np.array([e[2] for e in ents], dtype=dtype)
        return cls.with_triplets(nrows, ncols, r, c, v, dtype=dtype)

    @classmethod
    def with_triplets(cls, nrows, ncols, rowind, colind, values, dtype=None):   # coo.rs:254-288
        values = np.asarray(values) if dtype is None else np.asarray(values, dtype=dtype)
        if values.dtype not in (np.float32, np.float64):
            values = values.astype(np.float64)
        m = cls(nrows, ncols, values.dtype)
        rowind = np.ascontiguousarray(rowind, dtype=np.uint64)
        colind = np.ascontiguousarray(colind, dtype=np.uint64)
        if len(rowind) != len(values):
            raise Panic("assertion failed: rowind.len() == values.len()")
        if len(colind) != len(values):
            raise Panic("assertion failed: colind.len() == values.len()")
        if len(values) and not (rowind < nrows).all():
            raise Panic("assertion failed: *row < nrows")
        if len(values) and not (colind < ncols).all():
            raise Panic("assertion failed: *col < ncols")
        m._row, m._col, m._val = rowind.copy(), colind.copy(), np.ascontiguousarray(values).copy()
        m._len = len(values)
        return m

    # accessors -------------------------------------------------------------------
    def nrows(self): return self._nrows
    def ncols(self): return self._ncols
    def shape(self): return (self._nrows, self._ncols)
    def length(self): return self._len
    def capacity(self): return len(self._val)
    @property
    def dtype(self): return self._dtype

    def triplets(self):
        """SoA views of the stored entries (row, col, val), insertion order."""
        n = self._len
        return self._row[:n], self._col[:n], self._val[:n]

    def get(self, index: int):                                          # coo.rs:386-390
        if 0 <= index < self._len:
            return (int(self._row[index]), int(self._col[index]), self._val[index].item())
        return None

    def _reserve(self, cap: int):
        if cap > len(self._val):
            for name in ("_row", "_col", "_val"):
                old = getattr(self, name)
                new = np.empty(cap, old.dtype)
                new[:self._len] = old[:self._len]
                setattr(self, name, new)

    def push(self, row: int, col: int, value: float):                   # coo.rs:431-435
        if not row < self._nrows:
            raise Panic("assertion failed: row < self.nrows")
        if not col < self._ncols:
            raise Panic("assertion failed: col < self.ncols")
        if self._len == len(self._val):
            self._reserve(max(4, 2 * self._len))
        self._row[self._len], self._col[self._len], self._val[self._len] = row, col, value
        self._len += 1

    def pop(self):                                                      # coo.rs:450-452
        if self._len == 0:
            return None
        self._len -= 1
        i = self._len
        return (int(self._row[i]), int(self._col[i]), self._val[i].item())

    def clear(self):
        self._len = 0

    def iter(self) -> Iterator[Tuple[int, int, float]]:
        r, c, v = self.triplets()
        return iter(zip(r.tolist(), c.tolist(), v.tolist()))

    __iter__ = iter

    def extend(self, entries: Iterable[Tuple[int, int, float]]):        # coo.rs:566-573
        ents = list(entries)
        for (row, col, _) in ents:
            if not row < self._nrows:
                raise Panic("assertion failed: *row < self.nrows")
            if not col < self._ncols:
                raise Panic("assertion failed: *col < self.ncols")
        for e in ents:
            self.push(*e)

    def transpose(self) -> "CooMatrix":                                 # coo.rs:538-545
        r, c, v = self.triplets()
        return CooMatrix.with_triplets(self._ncols, self._nrows, c, r, v)

    # host-side ops (O(len) concatenations, coo.rs:751-804) ---------------------------------
    def _same_shape(self, rhs):
        if self._nrows != rhs._nrows or self._ncols != rhs._ncols:
            raise Panic("assertion `left == right` failed (shape)")

    def __add__(self, rhs: "CooMatrix"):
        self._same_shape(rhs)
        a, b = self.triplets(), rhs.triplets()
        return CooMatrix.with_triplets(self._nrows, self._ncols, np.concatenate([a[0], b[0]]),
                                       np.concatenate([a[1], b[1]]), np.concatenate([a[2], b[2]]))

    def __sub__(self, rhs: "CooMatrix"):
        self._same_shape(rhs)
        a, b = self.triplets(), rhs.triplets()
        return CooMatrix.with_triplets(self._nrows, self._ncols, np.concatenate([a[0], b[0]]),
                                       np.concatenate([a[1], b[1]]), np.concatenate([a[2], -b[2]]))

    def __neg__(self):
        r, c, v = self.triplets()
        return CooMatrix.with_triplets(self._nrows, self._ncols, r, c, -v)

    # From<&CsrMatrix>/<&CscMatrix>/<&DokMatrix> for CooMatrix (coo.rs:629-749)
    @classmethod
    def from_csr(cls, m: "CsrMatrix"): return m.to_coo()
    @classmethod
    def from_csc(cls, m: "CscMatrix"): return m.to_coo()

    @classmethod
    def from_dok(cls, dok: "DokMatrix"):
        items = list(dok.iter())
        return cls.with_entries(dok.nrows(), dok.ncols(), items, dtype=dok.dtype)


class PinnedCooMatrix(CooMatrix):
    """CooMatrix whose storage is an spl_coo (include/spl.h): pinned host SoA
    arrays, streamed to the device chunk by chunk while they are filled, so that
    CsrMatrix.from_coo / CscMatrix.from_coo find the triplets already in HBM.  Same surface as
    CooMatrix (push, extend, pop, clear, get, iter, length, capacity ...; src/coo.rs); needs a
    CUDA device — there is no CPU fallback for this storage."""

    def __init__(self, nrows: int, ncols: int, dtype=np.float64, capacity: int = 0, ctx=None):
        if not nrows > 0:
            raise Panic("assertion failed: nrows > 0")
        if not ncols > 0:
            raise Panic("assertion failed: ncols > 0")
        self._nrows, self._ncols = int(nrows), int(ncols)
        self._dtype = np.dtype(dtype)
        code = _dtype_code(self._dtype)
        self._ctx = ctx or default_context()
        self._lib = self._ctx._lib
        h = C.c_void_p()
        self._ctx.check(self._lib.spl_coo_create(self._ctx._h, code, self._nrows, self._ncols,
                                                 int(capacity), C.byref(h)))
        self._b = h

    def __del__(self):
        b = getattr(self, "_b", None)
        if b:
            self._lib.spl_coo_free(b)
            self._b = None

    def _check(self, status: int):
        if status == capi.SPL_OK:
            return
        msg = self._lib.spl_coo_last_error(self._b).decode(errors="replace")
        if status in (capi.SPL_ERR_ARG, capi.SPL_ERR_INVALID, capi.SPL_ERR_SHAPE):
            raise Panic(msg)
        raise DeviceError(f"{capi.STATUS_NAMES.get(status, status)}: {msg}")

    # constructors (src/coo.rs:104-112, 162-170, 204-220, 254-288) -----------------
    @classmethod
    def new(cls, nrows, ncols, dtype=np.float64, ctx=None):
        return cls(nrows, ncols, dtype, 0, ctx)

    @classmethod
    def with_capacity(cls, nrows, ncols, capacity, dtype=np.float64, ctx=None):
        return cls(nrows, ncols, dtype, capacity, ctx)

    @classmethod
    def with_triplets(cls, nrows, ncols, rowind, colind, values, dtype=None, ctx=None):
        values = np.asarray(values) if dtype is None else np.asarray(values, dtype=dtype)
        if values.dtype not in (np.float32, np.float64):
            values = values.astype(np.float64)
        if len(rowind) != len(values):
            raise Panic("assertion failed: rowind.len() == values.len()")
        if len(colind) != len(values):
            raise Panic("assertion failed: colind.len() == values.len()")
        m = cls(nrows, ncols, values.dtype, len(values), ctx)
        m.extend_triplets(rowind, colind, values)
        return m

    @classmethod
    def with_entries(cls, nrows, ncols, entries, dtype=np.float64, ctx=None):
        ents = list(entries)
        return cls.with_triplets(nrows, ncols, [e[0] for e in ents], [e[1] for e in ents],
                                 np.array([e[2] for e in ents], dtype=dtype), ctx=ctx)

    # accessors ---------------------------------------------------------------------
    @property
    def _len(self):
        return int(self._lib.spl_coo_len(self._b))

    def length(self): return self._len
    def capacity(self): return int(self._lib.spl_coo_capacity(self._b))
    def streamed(self): return int(self._lib.spl_coo_streamed(self._b))

    def triplets(self):
        """Views of the pinned arrays (valid until the next push / extend / reserve)."""
        n = self._len
        r, c, v = C.c_void_p(), C.c_void_p(), C.c_void_p()
        self._check(self._lib.spl_coo_host_ptrs(self._b, C.byref(r), C.byref(c), C.byref(v)))
        if n == 0:
            return np.empty(0, np.uint64), np.empty(0, np.uint64), np.empty(0, self._dtype)
        ct = C.c_float if self._dtype == np.float32 else C.c_double
        row = np.ctypeslib.as_array(C.cast(r, C.POINTER(C.c_uint64)), (n,))
        col = np.ctypeslib.as_array(C.cast(c, C.POINTER(C.c_uint64)), (n,))
        val = np.ctypeslib.as_array(C.cast(v, C.POINTER(ct)), (n,))
        return row, col, val

    def get(self, index: int):
        if 0 <= index < self._len:
            r, c, v = self.triplets()
            return (int(r[index]), int(c[index]), v[index].item())
        return None

    def _reserve(self, cap: int):
        self._check(self._lib.spl_coo_reserve(self._b, int(cap)))

    def push(self, row: int, col: int, value: float):                   # coo.rs:431-435
        if row < 0 or col < 0:
            raise Panic("index must be non-negative (usize)")
        v = np.array([value], dtype=self._dtype)
        self._check(self._lib.spl_coo_push(self._b, int(row), int(col), _ptr(v)))

    def extend_triplets(self, rowind, colind, values):
        """Bulk push of SoA triplets (with_triplets, coo.rs:254-288 / Extend, coo.rs:566-573)."""
        r = np.ascontiguousarray(rowind, dtype=np.uint64)
        c = np.ascontiguousarray(colind, dtype=np.uint64)
        v = np.ascontiguousarray(values, dtype=self._dtype)
        if not (len(r) == len(v) and len(c) == len(v)):
            raise Panic("assertion failed: rowind.len() == colind.len() == values.len()")
        self._check(self._lib.spl_coo_extend(self._b, len(v), _ptr(r), _ptr(c), _ptr(v)))

    def extend(self, entries):
        ents = list(entries)
        if ents:
            self.extend_triplets([e[0] for e in ents], [e[1] for e in ents], [e[2] for e in ents])

    def pop(self):                                                      # coo.rs:450-452
        n = self._len
        if n == 0:
            return None
        last = self.get(n - 1)
        self._check(self._lib.spl_coo_truncate(self._b, n - 1))
        return last

    def clear(self):                                                    # coo.rs:470-472
        self._check(self._lib.spl_coo_truncate(self._b, 0))


# --------------------------------------------------------------------------- DOK (host shell)
class DokMatrix:
    """Dictionary-of-keys matrix (src/dok.rs:53-58): unordered, keys unique."""

    def __init__(self, nrows: int, ncols: int, dtype=np.float64):
        if not nrows > 0:
            raise Panic("assertion failed: nrows > 0")
        if not ncols > 0:
            raise Panic("assertion failed: ncols > 0")
        self._nrows, self._ncols = int(nrows), int(ncols)
        self._dtype = np.dtype(dtype)
        _dtype_code(self._dtype)
        self._map = {}

    @classmethod
    def new(cls, nrows, ncols, dtype=np.float64): return cls(nrows, ncols, dtype)

    @classmethod
    def with_entries(cls, nrows, ncols, entries, dtype=np.float64):
        m = cls(nrows, ncols, dtype)
        for (r, c, v) in entries:
            m.insert(r, c, v)
        return m

    def nrows(self): return self._nrows
    def ncols(self): return self._ncols
    def shape(self): return (self._nrows, self._ncols)
    def length(self): return len(self._map)
    @property
    def dtype(self): return self._dtype

    def contains(self, row, col): return (row, col) in self._map
    def get(self, row, col): return self._map.get((row, col))

    def insert(self, row: int, col: int, value: float):                 # dok.rs:462-466
        if not row < self._nrows:
            raise Panic("assertion failed: row < self.nrows")
        if not col < self._ncols:
            raise Panic("assertion failed: col < self.ncols")
        old = self._map.get((row, col))
        self._map[(row, col)] = self._dtype.type(value)
        return old

    def clear(self): self._map.clear()

    def iter(self):
        return iter((r, c, v) for (r, c), v in self._map.items())

    __iter__ = iter

    def transpose(self):
        return DokMatrix.with_entries(self._ncols, self._nrows, ((c, r, v) for (r, c, v) in self.iter()),
                                      self._dtype)

    # conversions into DOK and DOK arithmetic: host-side shells (src/dok.rs:640-775)
    @classmethod
    def from_coo(cls, coo: "CooMatrix"):
        """From<&CooMatrix> (src/dok.rs:640-668): `*entry.or_default() += value`, i.e. every cell is
        summed from T::default() = 0.0 in insertion order (a lone -0.0 becomes +0.0); no zero drop."""
        m = cls(coo.nrows(), coo.ncols(), coo.dtype)
        zero = m._dtype.type(0)
        for (r, c, v) in coo.iter():
            m._map[(r, c)] = m._dtype.type(m._map.get((r, c), zero) + m._dtype.type(v))
        return m

    @classmethod
    def from_csr(cls, csr: "CsrMatrix"):
        """From<&CsrMatrix> (src/dok.rs:707-720): storage-order collect of the stored entries."""
        m = cls(csr.nrows(), csr.ncols(), csr.dtype)
        m._map = {(int(r), int(c)): m._dtype.type(v) for (r, c, v) in csr.iter()}
        return m

    @classmethod
    def from_csc(cls, csc: "CscMatrix"):
        """From<&CscMatrix> (src/dok.rs:676-693)."""
        return cls.from_csr(csc)

    def _merge(self, rhs: "DokMatrix", sign):
        out = DokMatrix(self._nrows, self._ncols, self._dtype)       # dims of self, like the reference
        out._map = dict(self._map)
        zero = self._dtype.type(0)
        for key, v in rhs._map.items():                               # or_default() then += / -=
            base = out._map.get(key, zero)
            out._map[key] = self._dtype.type(base + v) if sign > 0 else self._dtype.type(base - v)
        return out

    def __add__(self, rhs: "DokMatrix"): return self._merge(rhs, +1)       # src/dok.rs:722-736
    def __sub__(self, rhs: "DokMatrix"): return self._merge(rhs, -1)       # src/dok.rs:738-752

    def __neg__(self):                                                     # src/dok.rs:754-769
        out = DokMatrix(self._nrows, self._ncols, self._dtype)
        out._map = {k: self._dtype.type(-v) for k, v in self._map.items()}
        return out

    def triplets(self):
        n = len(self._map)
        r = np.fromiter((k[0] for k in self._map), np.uint64, n)
        c = np.fromiter((k[1] for k in self._map), np.uint64, n)
        v = np.fromiter(self._map.values(), self._dtype, n)
        return r, c, v


# --------------------------------------------------------------------------- CSR / CSC (device)
class _Compressed:
    _FORMAT = capi.SPL_CSR

    def __init__(self, nrows, ncols, ptr, ind, values, ctx: Optional[Context] = None):
        """CsrMatrix::new (src/csr.rs:137-164) / CscMatrix::new (src/csc.rs:137-164): validating."""
        ctx = ctx or default_context()
        values = np.ascontiguousarray(values)
        if values.dtype not in (np.float32, np.float64):
            values = values.astype(np.float64)
        ptr = np.ascontiguousarray(ptr, dtype=np.uint64)
        ind = np.ascontiguousarray(ind, dtype=np.uint64)
        h = C.c_void_p()
        ctx.check(ctx._lib.spl_mat_from_compressed(
            ctx._h, self._FORMAT, _dtype_code(values.dtype), int(nrows), int(ncols),
            len(ptr), _ptr(ptr), len(ind), _ptr(ind), len(values), _ptr(values), C.byref(h)))
        self._adopt(ctx, h)

    # -- plumbing
    def _adopt(self, ctx: Context, handle):
        self._ctx, self._h = ctx, handle
        fmt, dt = C.c_int(), C.c_int()
        nr, nc, nz = C.c_uint64(), C.c_uint64(), C.c_uint64()
        ctx._lib.spl_mat_info(handle, C.byref(fmt), C.byref(dt), C.byref(nr), C.byref(nc), C.byref(nz))
        assert fmt.value == self._FORMAT
        self._dtype = np.dtype(_NP[dt.value])
        self._nrows, self._ncols, self._nnz = nr.value, nc.value, nz.value
        self._host = None

    @classmethod
    def _wrap(cls, ctx: Context, handle):
        m = object.__new__(cls)
        m._adopt(ctx, handle)
        return m

    def __del__(self):
        h = getattr(self, "_h", None)
        ctx = getattr(self, "_ctx", None)
        if h and ctx is not None and getattr(ctx, "_h", None):
            try:
                ctx._lib.spl_mat_free(ctx._h, h)
            except Exception:
                pass
            self._h = None

    @classmethod
    def _other(cls):
        return CscMatrix if cls is CsrMatrix else CsrMatrix

    # -- constructors
    @classmethod
    def new(cls, nrows, ncols, ptr, ind, values, ctx=None):
        return cls(nrows, ncols, ptr, ind, values, ctx)

    @classmethod
    def eye(cls, size, dtype=np.float64, ctx=None):                    # src/csr.rs:179-188
        ctx = ctx or default_context()
        h = C.c_void_p()
        ctx.check(ctx._lib.spl_mat_eye(ctx._h, cls._FORMAT, _dtype_code(dtype), int(size), C.byref(h)))
        return cls._wrap(ctx, h)

    @classmethod
    def from_coo(cls, coo: CooMatrix, ctx=None):
        """From<&CooMatrix<T>> (src/csr/conv/coo.rs:3-116, src/csc/conv/coo.rs:3-116)."""
        ctx = ctx or default_context()
        h = C.c_void_p()
        if isinstance(coo, PinnedCooMatrix):         # triplets already streamed to the device
            ctx.check(ctx._lib.spl_mat_from_coo_builder(ctx._h, coo._b, cls._FORMAT, 1, 1, C.byref(h)))
            return cls._wrap(ctx, h)
        r, c, v = coo.triplets()
        ctx.check(ctx._lib.spl_mat_from_coo(ctx._h, cls._FORMAT, _dtype_code(v.dtype), coo.nrows(),
                                            coo.ncols(), len(v), _ptr(r), _ptr(c), _ptr(v), 1, 1,
                                            C.byref(h)))
        return cls._wrap(ctx, h)

    @classmethod
    def from_dok(cls, dok: DokMatrix, ctx=None):
        """From<&DokMatrix<T>> (src/csr/conv/dok.rs:3-76): no dedup needed, explicit zeros kept."""
        ctx = ctx or default_context()
        r, c, v = dok.triplets()
        h = C.c_void_p()
        ctx.check(ctx._lib.spl_mat_from_coo(ctx._h, cls._FORMAT, _dtype_code(v.dtype), dok.nrows(),
                                            dok.ncols(), len(v), _ptr(r), _ptr(c), _ptr(v), 0, 0,
                                            C.byref(h)))
        return cls._wrap(ctx, h)

    @classmethod
    def from_device_triplets(cls, nrows, ncols, length, row_dev: int, col_dev: int, val_dev: int, dtype,
                             dedup=True, dropzero=True, ctx=None):
        """Assembly from device-resident uint32 SoA triplets (raw device addresses)."""
        ctx = ctx or default_context()
        h = C.c_void_p()
        ctx.check(ctx._lib.spl_mat_from_coo_dev(ctx._h, cls._FORMAT, _dtype_code(dtype), nrows, ncols,
                                                length, C.c_void_p(row_dev), C.c_void_p(col_dev),
                                                C.c_void_p(val_dev), int(dedup), int(dropzero),
                                                C.byref(h)))
        return cls._wrap(ctx, h)

    @classmethod
    def from_device_arrays(cls, nrows, ncols, nnz, ptr_dev: int, ind_dev: int, val_dev: int, dtype,
                           validate=True, ctx=None):
        """CsrMatrix::new on device-resident uint32 ptr/ind and T values (copied)."""
        ctx = ctx or default_context()
        h = C.c_void_p()
        ctx.check(ctx._lib.spl_mat_from_compressed_dev(
            ctx._h, cls._FORMAT, _dtype_code(dtype), nrows, ncols, nnz, C.c_void_p(ptr_dev),
            C.c_void_p(ind_dev), C.c_void_p(val_dev), int(validate), C.byref(h)))
        return cls._wrap(ctx, h)

    @classmethod
    def from_device_arrays64(cls, nrows, ncols, nnz, ptr64_dev: int, ind_dev: int, val_dev: int, dtype,
                             validate=True, ctx=None):
        """CsrMatrix::new on device-resident arrays with a 64-bit pointer array (usize, src/csr.rs:66-72): the
        constructor for matrices with 2^32 - 65536 stored entries or more (spl_mat_from_compressed_dev64)."""
        ctx = ctx or default_context()
        h = C.c_void_p()
        ctx.check(ctx._lib.spl_mat_from_compressed_dev64(
            ctx._h, cls._FORMAT, _dtype_code(dtype), nrows, ncols, nnz, C.c_void_p(ptr64_dev),
            C.c_void_p(ind_dev), C.c_void_p(val_dev), int(validate), C.byref(h)))
        return cls._wrap(ctx, h)

    def device_ptr64(self):
        """Device address of the 64-bit pointer array of a wide matrix (None for the usual 32-bit ones)."""
        p = C.c_void_p()
        self._ctx._lib.spl_mat_device_ptr64(self._h, C.byref(p))
        return p.value

    def read_entries(self, start: int, count: int):
        """Stored entries [start, start + count) in storage order as (rows, cols, values) host arrays."""
        r = np.empty(count, np.uint64)
        c = np.empty(count, np.uint64)
        v = np.empty(count, self._dtype)
        self._ctx.check(self._ctx._lib.spl_mat_read_entries(self._ctx._h, self._h, int(start), int(count), _ptr(r), _ptr(c),
                                                            _ptr(v)))
        return r, c, v

    def _convert(self, target_cls):
        h = C.c_void_p()
        self._ctx.check(self._ctx._lib.spl_mat_convert(self._ctx._h, self._h, target_cls._FORMAT,
                                                       C.byref(h)))
        return target_cls._wrap(self._ctx, h)

    # -- accessors (src/csr.rs:200-289)
    def nrows(self): return self._nrows
    def ncols(self): return self._ncols
    def nnz(self): return self._nnz
    def shape(self): return (self._nrows, self._ncols)
    @property
    def dtype(self): return self._dtype

    def _download(self):
        if self._host is None:
            nmajor = self._nrows if self._FORMAT == capi.SPL_CSR else self._ncols
            ptr = np.empty(nmajor + 1, np.uint64)
            ind = np.empty(self._nnz, np.uint64)       # exactly sized (capacity == len in the reference)
            val = np.empty(self._nnz, self._dtype)
            self._ctx.check(self._ctx._lib.spl_mat_download(self._ctx._h, self._h, _ptr(ptr), _ptr(ind),
                                                            _ptr(val)))
            self._host = (ptr, ind, val)
        return self._host

    def values(self): return self._download()[2]

    def set_values(self, values):
        """Overwrite the stored values, structure kept (what a caller does through values_mut(),
        src/csr.rs:270-272 / src/csc.rs:270-272)."""
        v = np.ascontiguousarray(values, dtype=self._dtype)
        if len(v) != self._nnz:
            raise Panic("assertion `left == right` failed: values.len() == self.nnz()")
        self._ctx.check(self._ctx._lib.spl_mat_set_values(self._ctx._h, self._h, _ptr(v)))
        if self._host is not None:
            self._host = (self._host[0], self._host[1], v.copy())

    def values_mut(self):
        """`with m.values_mut() as v: v[...] = ...` — a writable host copy of the values that is
        written back to the device on exit (src/csr.rs:270-272)."""
        mat = self

        class _Mut:
            def __enter__(self_inner):
                self_inner.v = mat.values().copy()
                return self_inner.v

            def __exit__(self_inner, exc_type, exc, tb):
                if exc_type is None:
                    mat.set_values(self_inner.v)
                return False
        return _Mut()

    def device_ptrs(self):
        p, i, v = C.c_void_p(), C.c_void_p(), C.c_void_p()
        self._ctx._lib.spl_mat_device_ptrs(self._h, C.byref(p), C.byref(i), C.byref(v))
        return p.value, i.value, v.value

    def to_coo(self) -> CooMatrix:
        """From<&CsrMatrix>/<&CscMatrix> for CooMatrix (src/coo.rs:629-705): storage order."""
        r = np.empty(self._nnz, np.uint64)
        c = np.empty(self._nnz, np.uint64)
        v = np.empty(self._nnz, self._dtype)
        self._ctx.check(self._ctx._lib.spl_mat_to_coo(self._ctx._h, self._h, _ptr(r), _ptr(c), _ptr(v)))
        return CooMatrix.with_triplets(self._nrows, self._ncols, r, c, v)

    ITER_CHUNK = 1 << 16

    def iter(self):
        """iter() (src/csr.rs:303-316 / src/csc.rs:303-316): (row, col, value) in storage order, read
        from the device in chunks of ITER_CHUNK entries (spl_mat_read_entries): the host never holds
        more than one chunk, whatever nnz is."""
        n = self._nnz
        cap = min(self.ITER_CHUNK, max(n, 1))
        r = np.empty(cap, np.uint64)
        c = np.empty(cap, np.uint64)
        v = np.empty(cap, self._dtype)
        for start in range(0, n, cap):
            cnt = min(cap, n - start)
            self._ctx.check(self._ctx._lib.spl_mat_read_entries(self._ctx._h, self._h, start, cnt, _ptr(r), _ptr(c),
                                                                _ptr(v)))
            yield from zip(r[:cnt].tolist(), c[:cnt].tolist(), v[:cnt].tolist())

    __iter__ = iter

    def to_coo_device(self, row_dev: int, col_dev: int, val_dev: int):
        """From<&CsrMatrix>/<&CscMatrix> for CooMatrix with the triplets left on the device (raw device
        addresses of uint32 row/col and T value arrays with nnz slots): spl_mat_to_coo_dev."""
        self._ctx.check(self._ctx._lib.spl_mat_to_coo_dev(self._ctx._h, self._h, C.c_void_p(row_dev), C.c_void_p(col_dev),
                                                          C.c_void_p(val_dev)))

    # -- hot-path operators
    def transpose(self):
        """src/csr.rs:358-406 / src/csc.rs:358-406."""
        h = C.c_void_p()
        self._ctx.check(self._ctx._lib.spl_mat_transpose(self._ctx._h, self._h, C.byref(h)))
        return type(self)._wrap(self._ctx, h)

    def _binary(self, fn, rhs):
        if type(rhs) is not type(self):
            return NotImplemented
        h = C.c_void_p()
        self._ctx.check(fn(self._ctx._h, self._h, rhs._h, C.byref(h)))
        return type(self)._wrap(self._ctx, h)

    def __add__(self, rhs): return self._binary(self._ctx._lib.spl_mat_add, rhs)
    def __sub__(self, rhs): return self._binary(self._ctx._lib.spl_mat_sub, rhs)
    def __mul__(self, rhs): return self._binary(self._ctx._lib.spl_mat_mul, rhs)
    __matmul__ = __mul__

    def __neg__(self):
        h = C.c_void_p()
        self._ctx.check(self._ctx._lib.spl_mat_neg(self._ctx._h, self._h, C.byref(h)))
        return type(self)._wrap(self._ctx, h)


    def matvec(self, x: np.ndarray, out: Optional[np.ndarray] = None) -> np.ndarray:
        """y = A x with host vectors (extension; reference route is `&A * &X`, X n x 1,
        src/csr/ops/mul.rs:5-60, src/csc/ops/mul.rs:5-61).  Uploads x, runs the SpMV kernel, downloads y.
        On a CscMatrix the first product builds (and keeps) the CSR form on the device.  With x and out from
        pinned_empty() the three steps are pipelined over row chunks."""
        x = np.ascontiguousarray(x, dtype=self._dtype)
        if len(x) != self._ncols:
            raise Panic("assertion `left == right` failed: self.ncols() == rhs.nrows()")
        if out is not None:
            if out.dtype != self._dtype or out.shape != (self._nrows,) or not out.flags.c_contiguous:
                raise Panic("assertion `left == right` failed: self.nrows() == out.len()")
            y = out
        else:
            y = np.empty(self._nrows, self._dtype)
        self._ctx.check(self._ctx._lib.spl_spmv_host(self._ctx._h, self._h, _ptr(x), _ptr(y)))
        return y

    def spmv_device(self, x_dev: int, y_dev: int, kernel: int = capi.SPL_SPMV_AUTO, lanes: int = 0):
        """y = A x on raw device addresses (asynchronous on the context's stream)."""
        self._ctx.check(self._ctx._lib.spl_spmv_ex(self._ctx._h, self._h, C.c_void_p(x_dev),
                                                   C.c_void_p(y_dev), int(kernel) | (int(lanes) << 8)))

    def spmv_choice(self):
        k, l = C.c_int(), C.c_int()
        self._ctx.check(self._ctx._lib.spl_spmv_choice(self._ctx._h, self._h, C.byref(k), C.byref(l)))
        return k.value, l.value


class CsrMatrix(_Compressed):
    """Compressed sparse row matrix in HBM (reference: src/csr.rs:65-72)."""
    _FORMAT = capi.SPL_CSR

    def rowptr(self): return self._download()[0]
    def colind(self): return self._download()[1]

    @classmethod
    def from_csc(cls, csc: "CscMatrix"):
        """From<&CscMatrix<T>> for CsrMatrix<T> (src/csr/conv/csc.rs:3-53)."""
        return csc._convert(cls)

    def to_csc(self) -> "CscMatrix":
        return self._convert(CscMatrix)



class CscMatrix(_Compressed):
    """Compressed sparse column matrix in HBM (reference: src/csc.rs:65-72)."""
    _FORMAT = capi.SPL_CSC

    def colptr(self): return self._download()[0]
    def rowind(self): return self._download()[1]

    @classmethod
    def from_csr(cls, csr: CsrMatrix):
        """From<&CsrMatrix<T>> for CscMatrix<T> (src/csc/conv/csr.rs:3-53)."""
        return csr._convert(cls)

    def to_csr(self) -> CsrMatrix:
        return self._convert(CsrMatrix)
