"""BASELINE.json configs 3, 4, 5 on the GPU: bit-exact parity against the oracle at sizes the oracle
finishes in seconds, and size-independent properties at the full sizes (validity of the output,
idempotence of assembly, linearity against a direct COO evaluation, agreement of the two SpMV
kernels, (A+B)x = Ax + Bx)."""
import numpy as np
import pytest

import oracle as orc
import spalinalg_b200 as sp
from spalinalg_b200 import _capi as capi
from spalinalg_b200 import synthetic as syn

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(autouse=True, scope="module")
def one_stream():
    """torch (generators, checks) and the library share one explicit stream, so no cross-stream
    synchronisation is needed between a torch op and the API call that consumes its output."""
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    old = getattr(sp.matrix._tls, "ctx", None)
    sp.set_default_context(sp.Context(0, stream.cuda_stream))
    yield
    torch.cuda.synchronize()
    sp.set_default_context(old)
    torch.cuda.set_stream(torch.cuda.default_stream())


def gen():
    from spalinalg_b200 import synthetic_device as sd
    return sd


def dev_arrays(A):
    sd = gen()
    p, i, v = A.device_ptrs()
    nmajor = A.nrows() if isinstance(A, sp.CsrMatrix) else A.ncols()
    tdt = torch.float32 if A.dtype == np.float32 else torch.float64
    return (sd.device_view(torch, p, nmajor + 1, torch.int32), sd.device_view(torch, i, max(A.nnz(), 1), torch.int32)[:A.nnz()],
            sd.device_view(torch, v, max(A.nnz(), 1), tdt)[:A.nnz()])


def spmv(A, x, kernel=capi.SPL_SPMV_AUTO, lanes=0):
    y = torch.empty(A.nrows(), device="cuda", dtype=x.dtype)
    torch.cuda.synchronize()
    A.spmv_device(x.data_ptr(), y.data_ptr(), kernel, lanes)
    sp.default_context().sync()
    return y


def coo_matvec_f64(n, r, c, v, x):
    """Direct COO evaluation in f64 (order-free reference for the linearity property)."""
    y = torch.zeros(n, device="cuda", dtype=torch.float64)
    y.index_add_(0, r.long(), v.double() * x.double()[c.long()])
    return y


def same_host(A, want):
    assert np.array_equal(A.rowptr(), want[0]) and np.array_equal(A.colind(), want[1])
    assert A.values().tobytes() == want[2].tobytes()


def host_ram_gb():
    try:
        import psutil
        return psutil.virtual_memory().available / 2 ** 30
    except Exception:                                             # noqa: BLE001
        return 0.0


def aos_triplets_from_device(r, c, v):
    """The oracle's input — the reference's Vec<(usize, usize, T)> (src/coo.rs:52-57), 24 bytes per
    entry — built on the device and brought down in one copy (numpy would spend longer converting the
    three arrays than the oracle spends assembling them)."""
    n = r.numel()
    aos = torch.empty((n, 3), dtype=torch.int64, device="cuda")
    aos[:, 0] = r.long() & 0xFFFFFFFF                             # int32 tensors hold uint32 bit patterns
    aos[:, 1] = c.long() & 0xFFFFFFFF
    if v.dtype == torch.float32:
        aos[:, 2] = v.view(torch.int32).long() & 0xFFFFFFFF       # value in the low 4 bytes, padding above
    else:
        aos[:, 2] = v.view(torch.int64)
    host = aos.cpu().numpy()
    del aos
    return host.reshape(-1).view(orc.triplet_dtype(np.float32 if v.dtype == torch.float32 else np.float64))


def same_compressed_big(A, want):
    """Bit-exact comparison of a large device matrix with the oracle's arrays: pointers and indices on
    the host (uint64), values as raw bytes."""
    ptr, ind, val = (A.rowptr(), A.colind(), A.values()) if isinstance(A, sp.CsrMatrix) else (A.colptr(), A.rowind(), A.values())
    assert len(ind) == len(want[1]), (len(ind), len(want[1]))
    assert np.array_equal(ptr, want[0]), "pointers differ"
    assert np.array_equal(ind, want[1]), "indices differ"
    assert val.tobytes() == np.asarray(want[2]).tobytes(), "values differ (bitwise)"


# ------------------------------------------------------------------ config 3
def test_c3_full_size_bit_exact_vs_oracle():
    """Config 3 at its full size (168 M triplets, 5 % duplicates incl. exact cancellations and >= 3-fold
    cells) against the oracle's restatement of src/csr/conv/coo.rs:4-115, bit for bit; then the
    assembled matrix CSR -> CSC against src/csc/conv/csr.rs:4-52."""
    if host_ram_gb() < 24:
        pytest.skip(f"host RAM {host_ram_gb():.0f} GB < 24 GB: the oracle needs the 4 GB AoS list plus ~8 GB of work arrays")
    sd = gen()
    n = 10_000_000
    r, c, v = sd.random_uniform_coo_device(torch, n, 16, 8_000_000, torch.float32, seed=1)
    assert r.numel() == 168_000_000
    A = sp.CsrMatrix.from_device_triplets(n, n, r.numel(), r.data_ptr(), c.data_ptr(), v.data_ptr(), np.float32)
    trip = aos_triplets_from_device(r, c, v)
    del r, c, v
    torch.cuda.empty_cache()
    want = orc.compress_from_coo(n, n, trip, "row")
    del trip
    same_compressed_big(A, want)
    C = A.to_csc()
    wantc = orc.recompress(n, n, *want)
    del want
    same_compressed_big(C, wantc)


def test_c3_reduced_bit_exact_vs_oracle():
    sd = gen()
    n = 1_000_000
    r, c, v = sd.random_uniform_coo_device(torch, n, 16, 800_000, torch.float32, seed=1)
    A = sp.CsrMatrix.from_device_triplets(n, n, r.numel(), r.data_ptr(), c.data_ptr(), v.data_ptr(), np.float32)
    rh, ch, vh = r.cpu().numpy().astype(np.uint64), c.cpu().numpy().astype(np.uint64), v.cpu().numpy()
    want = orc.compress_from_coo(n, n, orc.make_triplets(rh, ch, vh), "row")
    same_host(A, want)
    C = sp.CscMatrix.from_device_triplets(n, n, r.numel(), r.data_ptr(), c.data_ptr(), v.data_ptr(), np.float32)
    wantc = orc.compress_from_coo(n, n, orc.make_triplets(rh, ch, vh), "col")
    assert np.array_equal(C.colptr(), wantc[0]) and np.array_equal(C.rowind(), wantc[1])
    assert C.values().tobytes() == wantc[2].tobytes()


def test_c3_full_size_properties():
    sd = gen()
    n, per_row, extra = 10_000_000, 16, 8_000_000
    r, c, v = sd.random_uniform_coo_device(torch, n, per_row, extra, torch.float32, seed=1)
    ln = r.numel()
    assert ln == 168_000_000
    A = sp.CsrMatrix.from_device_triplets(n, n, ln, r.data_ptr(), c.data_ptr(), v.data_ptr(), np.float32)
    ptr, ind, val = dev_arrays(A)
    # (1) the output is a valid CsrMatrix (sorted pointers, strictly increasing columns, in range)
    sp.CsrMatrix.from_device_arrays(n, n, A.nnz(), ptr.data_ptr(), ind.data_ptr(), val.data_ptr(), np.float32, validate=True)
    # (2) nnz = distinct cells minus cells whose in-order sum is exactly zero (at most the negations)
    key = r.long() * n + c.long()
    uniq = int(torch.unique(key).numel())
    del key
    assert uniq - extra // 100 - 1 <= A.nnz() <= uniq
    assert not bool((val == 0).any())                                       # zero sums were dropped
    # (3) linearity: A x equals the direct COO evaluation (f32 sums, order differs)
    x = torch.rand(n, device="cuda", dtype=torch.float32) - 0.5
    y = spmv(A, x).double()
    yref = coo_matvec_f64(n, r, c, v, x)
    scale = coo_matvec_f64(n, r, c, v.abs(), x.abs())
    assert bool(((y - yref).abs() <= 1e-5 * scale + 1e-30).all())
    # (4) idempotence: assembling the assembled matrix's own entries returns it bit for bit
    rows2 = torch.repeat_interleave(torch.arange(n, device="cuda", dtype=torch.int32),
                                    (ptr[1:] - ptr[:-1]).long())
    g = torch.Generator(device="cuda"); g.manual_seed(9)
    perm = torch.randperm(A.nnz(), device="cuda", generator=g)
    r2, c2, v2 = rows2[perm].contiguous(), ind[perm].contiguous(), val[perm].contiguous()
    B = sp.CsrMatrix.from_device_triplets(n, n, A.nnz(), r2.data_ptr(), c2.data_ptr(), v2.data_ptr(), np.float32)
    bp, bi, bv = dev_arrays(B)
    assert B.nnz() == A.nnz() and torch.equal(bp, ptr) and torch.equal(bi, ind)
    assert torch.equal(bv.view(torch.int32), val.view(torch.int32))


# ------------------------------------------------------------------ config 4
def test_c4_reduced_bit_exact_and_merge_path():
    sd = gen()
    scale = 18
    n = 1 << scale
    r, c, v = sd.rmat_coo_device(torch, scale, 32, torch.float32, seed=3)
    A = sp.CsrMatrix.from_device_triplets(n, n, r.numel(), r.data_ptr(), c.data_ptr(), v.data_ptr(), np.float32)
    rh, ch, vh = r.cpu().numpy().astype(np.uint64), c.cpu().numpy().astype(np.uint64), v.cpu().numpy()
    want = orc.compress_from_coo(n, n, orc.make_triplets(rh, ch, vh), "row")
    same_host(A, want)
    assert A.spmv_choice()[0] == capi.SPL_SPMV_SPLIT                       # skewed rows
    x = torch.rand(n, device="cuda", dtype=torch.float32) - 0.5
    xh = x.cpu().numpy()
    yw = orc.csr_spmv(n, *want, xh)
    sc = orc.csr_spmv(n, want[0], want[1], np.abs(want[2]), np.abs(xh))
    for kern, lanes in ((capi.SPL_SPMV_MERGE, 0), (capi.SPL_SPMV_SPLIT, 0), (capi.SPL_SPMV_VECTOR, 32), (capi.SPL_SPMV_AUTO, 0)):
        y = spmv(A, x, kern, lanes).cpu().numpy()
        assert np.all(np.abs(y - yw) <= 1e-5 * np.maximum(sc, 1e-30)), kern
    # transpose of a skewed matrix, bit-exact
    T = A.transpose()
    wt = orc.recompress(n, n, *want)
    same_host(T, wt)


def test_c4_full_size_bit_exact_vs_oracle():
    """Config 4 at its full size (R-MAT scale 24, 2^29 generated edges, heavy duplicate merging in the
    hot cells, rows of up to ~10^6 entries) against the oracle, bit for bit."""
    if host_ram_gb() < 56:
        pytest.skip(f"host RAM {host_ram_gb():.0f} GB < 56 GB: the oracle needs the 12.9 GB AoS list plus ~25 GB of work arrays")
    sd = gen()
    scale = 24
    n = 1 << scale
    r, c, v = sd.rmat_coo_device(torch, scale, 32, torch.float32, seed=3)
    A = sp.CsrMatrix.from_device_triplets(n, n, r.numel(), r.data_ptr(), c.data_ptr(), v.data_ptr(), np.float32)
    trip = aos_triplets_from_device(r, c, v)
    del r, c, v
    torch.cuda.empty_cache()
    want = orc.compress_from_coo(n, n, trip, "row")
    del trip
    same_compressed_big(A, want)


def test_c4_full_size_properties():
    sd = gen()
    scale = 24
    n = 1 << scale
    r, c, v = sd.rmat_coo_device(torch, scale, 32, torch.float32, seed=3)
    assert r.numel() == 1 << 29
    A = sp.CsrMatrix.from_device_triplets(n, n, r.numel(), r.data_ptr(), c.data_ptr(), v.data_ptr(), np.float32)
    ptr, ind, val = dev_arrays(A)
    sp.CsrMatrix.from_device_arrays(n, n, A.nnz(), ptr.data_ptr(), ind.data_ptr(), val.data_ptr(), np.float32, validate=True)
    assert A.nnz() < r.numel()                                               # duplicates merged
    x = torch.rand(n, device="cuda", dtype=torch.float32) - 0.5
    yref = coo_matvec_f64(n, r, c, v, x)
    sc = coo_matvec_f64(n, r, c, v.abs(), x.abs())
    del r, c, v
    ym = spmv(A, x, capi.SPL_SPMV_MERGE).double()
    yv = spmv(A, x, capi.SPL_SPMV_VECTOR, 32).double()
    # duplicates are summed in f32 during assembly as well: allow two roundings per addend
    assert bool(((ym - yref).abs() <= 4e-5 * sc + 1e-30).all())
    assert bool(((yv - ym).abs() <= 1e-5 * sc + 1e-30).all())
    ys = spmv(A, x, capi.SPL_SPMV_SPLIT).double()
    assert bool(((ys - ym).abs() <= 1e-5 * sc + 1e-30).all())
    assert torch.equal(ys, spmv(A, x, capi.SPL_SPMV_SPLIT).double())         # run-to-run identical


# ------------------------------------------------------------------ config 5
def _banded(n, offsets, r0=0, r1=None):
    sd = gen()
    r1 = n if r1 is None else r1
    ptr, col, val = sd.banded_device(torch, n, r0, r1, offsets, torch.float64)
    A = sp.CsrMatrix.from_device_arrays(r1 - r0, n, col.numel(), ptr.data_ptr(), col.data_ptr(), val.data_ptr(),
                                        np.float64, validate=True)
    return A, (ptr, col, val)


def test_c5_reduced_add_sub_bit_exact():
    n = 200_000
    A, (ap, ac, av) = _banded(n, range(-4, 5))
    B, (bp, bc, bv) = _banded(n, (-8, -2, 0, 2, 8))
    a = (ap.cpu().numpy().astype(np.uint64), ac.cpu().numpy().astype(np.uint64), av.cpu().numpy())
    b = (bp.cpu().numpy().astype(np.uint64), bc.cpu().numpy().astype(np.uint64), bv.cpu().numpy())
    same_host(A + B, orc.addsub(0, n, n, a, b))
    same_host(A - B, orc.addsub(1, n, n, a, b))


def test_c5_full_size_spmv_and_add():
    n = 100_000_000
    A, _ = _banded(n, range(-4, 5))
    assert A.nnz() == 899_999_980
    B, _ = _banded(n, (-8, -2, 0, 2, 8))
    x = torch.sin(torch.arange(n, device="cuda", dtype=torch.float64) * 1e-3)
    ya, yb = spmv(A, x), spmv(B, x)
    # analytic check of A x on a slice: y_i = sum_d (1/(1+|d|) + (i mod 7) 1e-3) x_{i+d}
    i = torch.arange(10, 1_000_010, device="cuda", dtype=torch.int64)
    want = torch.zeros(i.numel(), device="cuda", dtype=torch.float64)
    for d in range(-4, 5):
        want += (1.0 / (1 + abs(d)) + (i % 7).double() * 1e-3) * x[i + d]
    assert bool(((ya[i] - want).abs() <= 1e-12 * 10).all())
    C = A + B
    assert C.nnz() == 9 * n - 20 + 2 * (n - 8)                               # union adds offsets +-8
    yc = spmv(C, x)
    assert bool(((yc - (ya + yb)).abs() <= 1e-12 * 20).all())
    D = C - B                                                                # pattern of C, values of A (+0.0 on +-8)
    yd = spmv(D, x)
    assert bool(((yd - ya).abs() <= 1e-12 * 20).all())


# ------------------------------------------------------------------ beyond 2^32 stored entries
def test_wide_matrix_beyond_2_pow_32_entries():
    """A CsrMatrix with more than 2^32 stored entries (the reference indexes with usize, src/csr.rs:66-72;
    the device keeps a 64-bit pointer array for such matrices): CsrMatrix::new validation, SpMV against the
    analytic band, transpose against the analytic transposed band, transpose twice = the matrix bit for bit,
    entry chunks past position 2^32, and the strictly-increasing assertion on an entry beyond 2^32.
    f32, 65 diagonals, n = 2^26: 4.36 G entries, 35 GB per copy."""
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    sp.default_context().trim()                              # what earlier tests left in the library's pool
    free, _ = torch.cuda.mem_get_info()
    if free < 150 * 2 ** 30:
        pytest.skip(f"needs ~150 GB of free device memory, {free / 2 ** 30:.0f} GB available")
    sd = gen()
    n, half = 1 << 26, 32
    offs = torch.arange(-half, half + 1, device="cuda", dtype=torch.int64)
    nnz = sum(n - abs(d) for d in range(-half, half + 1))
    assert nnz > 2 ** 32
    ind = torch.empty(nnz, dtype=torch.int32, device="cuda")
    val = torch.empty(nnz, dtype=torch.float32, device="cuda")
    cnt = torch.empty(n, dtype=torch.int64, device="cuda")
    pos, chunk = 0, 1 << 21

    def band_values(i, d):                                   # A[i, i+d], f32
        return (1.0 / (1.0 + d.abs().double()) + (i % 7).double() * 1e-3).float()

    for s in range(0, n, chunk):
        i = torch.arange(s, min(n, s + chunk), device="cuda", dtype=torch.int64)
        cols = i[:, None] + offs[None, :]
        mask = (cols >= 0) & (cols < n)
        k = int(mask.sum().item())
        ind[pos:pos + k] = cols[mask].to(torch.int32)
        val[pos:pos + k] = band_values(i[:, None].expand_as(cols)[mask], offs[None, :].expand_as(cols)[mask])
        cnt[s:s + i.numel()] = mask.sum(1)
        pos += k
        del cols, mask, i
    assert pos == nnz
    ptr = torch.zeros(n + 1, dtype=torch.int64, device="cuda")
    ptr[1:] = torch.cumsum(cnt, 0)
    del cnt
    torch.cuda.synchronize()
    A = sp.CsrMatrix.from_device_arrays64(n, n, nnz, ptr.data_ptr(), ind.data_ptr(), val.data_ptr(), np.float32, validate=True)
    assert A.nnz() == nnz and A.spmv_choice()[0] == capi.SPL_SPMV_VECTOR
    # the strictly-increasing assertion (src/csr.rs:152-156) on an entry beyond position 2^32
    last = int(ptr[n - 5].item())
    assert last > 2 ** 32
    ind[last], ind[last + 1] = ind[last + 1].clone(), ind[last].clone()
    torch.cuda.synchronize()
    with pytest.raises(sp.Panic):
        sp.CsrMatrix.from_device_arrays64(n, n, nnz, ptr.data_ptr(), ind.data_ptr(), val.data_ptr(), np.float32, validate=True)
    assert sp.default_context().invalid_reason() == 9
    ind[last], ind[last + 1] = ind[last + 1].clone(), ind[last].clone()
    # entry chunks behind position 2^32: storage order, rows found from the 64-bit pointers
    r, c, v = A.read_entries(nnz - 70, 70)
    assert np.array_equal(c, ind[nnz - 70:].cpu().numpy().astype(np.uint64)) and v.tobytes() == val[nnz - 70:].cpu().numpy().tobytes()
    assert r[-1] == n - 1 and r[0] == n - 3 and np.all(np.diff(r.astype(np.int64)) >= 0)
    del ind, val
    torch.cuda.empty_cache()

    def check_product(M, transposed):
        x = torch.sin(torch.arange(n, device="cuda", dtype=torch.float64) * 1e-3).float()
        y = spmv(M, x)
        for lo in (0, n // 2 - 3, n - 200_000):               # first rows, the middle, rows whose entries lie beyond 2^32
            i = torch.arange(lo, min(n, lo + 200_000), device="cuda", dtype=torch.int64)
            want = torch.zeros(i.numel(), device="cuda", dtype=torch.float64)
            scale = torch.zeros_like(want)
            for d in range(-half, half + 1):
                j = i + d
                ok = (j >= 0) & (j < n)
                jj = j.clamp(0, n - 1)
                dd = torch.full_like(i, d)
                a = band_values(jj, -dd) if transposed else band_values(i, dd)      # A^T[i, j] = A[j, i] = band(j, i - j)
                term = torch.where(ok, a.double() * x[jj].double(), torch.zeros_like(want))
                want += term
                scale += term.abs()
            assert bool(((y[i].double() - want).abs() <= 1e-5 * scale + 1e-30).all()), (transposed, lo)

    check_product(A, False)
    T = A.transpose()                                          # histogram + 64-bit scan + scatter + repair (wide.cu)
    assert T.nnz() == nnz
    check_product(T, True)
    TT = T.transpose()
    del T
    torch.cuda.empty_cache()
    p64 = sd.device_view(torch, TT.device_ptr64(), n + 1, torch.int64)
    assert torch.equal(p64, ptr)
    _, ti, tv = TT.device_ptrs()
    _, ai, av = A.device_ptrs()
    step = 1 << 28                                             # compare in slices: no 35 GB temporaries
    for s in range(0, nnz, step):
        k = min(step, nnz - s)
        assert torch.equal(sd.device_view(torch, ti + 4 * s, k, torch.int32), sd.device_view(torch, ai + 4 * s, k, torch.int32))
        assert torch.equal(sd.device_view(torch, tv + 4 * s, k, torch.int32), sd.device_view(torch, av + 4 * s, k, torch.int32))
    # the 32-bit paths refuse a wide operand instead of truncating its positions
    with pytest.raises(sp.DeviceError):
        A + TT
