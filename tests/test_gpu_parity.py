"""GPU parity tests: the CUDA path, called through the C ABI (via the host mirror), against
(1) the reference's own golden vectors and (2) the CPU oracle on seeded random inputs.
Bar: bit-exact for structure, conversions, add/sub/mul/neg values; 1e-12 (f64) / 1e-5 (f32)
relative for SpMV (BASELINE.json north_star)."""
import numpy as np
import pytest

import oracle as orc
import spalinalg_b200 as sp
from spalinalg_b200 import synthetic as syn

pytestmark = pytest.mark.gpu

SPMV_RTOL = {np.dtype(np.float64): 1e-12, np.dtype(np.float32): 1e-5}


def arrays(m):
    if isinstance(m, sp.CsrMatrix):
        return m.rowptr(), m.colind(), m.values()
    return m.colptr(), m.rowind(), m.values()


def same(got, want, what=""):
    gp, gi, gv = got
    wp, wi, wv = want
    assert gp.dtype == np.uint64 and gi.dtype == np.uint64
    assert np.array_equal(gp, np.asarray(wp, np.uint64)), f"{what}: ptr differs"
    assert np.array_equal(gi, np.asarray(wi, np.uint64)), f"{what}: ind differs"
    wv = np.asarray(wv, gv.dtype)
    assert gv.tobytes() == wv.tobytes(), f"{what}: values differ (bitwise)"


def G(m, dtype=np.float64):
    return m[0], m[1], np.array(m[2], np.uint64), np.array(m[3], np.uint64), np.array(m[4], dtype)


# ------------------------------------------------------------------ the reference's own tests
def test_from_coo_golden(goldens):
    """src/csr/conv/coo.rs:129-145, src/csc/conv/coo.rs:129-145."""
    g = goldens["coo_pushes"]
    coo = sp.CooMatrix.new(g["nrows"], g["ncols"])
    for r, c, v in g["entries"]:
        coo.push(int(r), int(c), v)
    csr = sp.CsrMatrix.from_coo(coo)
    assert csr.rowptr().tolist() == g["csr"][2] and csr.colind().tolist() == g["csr"][3]
    assert csr.values().tolist() == g["csr"][4]
    csc = sp.CscMatrix.from_coo(coo)
    assert csc.colptr().tolist() == g["csc"][2] and csc.rowind().tolist() == g["csc"][3]
    assert csc.values().tolist() == g["csc"][4]


def test_from_dok_golden(goldens):
    g = goldens["dok_inserts"]
    dok = sp.DokMatrix.new(2, 3)
    for r, c, v in g["entries"]:
        dok.insert(int(r), int(c), v)
    same(arrays(sp.CsrMatrix.from_dok(dok)), G(g["csr"])[2:])
    same(arrays(sp.CscMatrix.from_dok(dok)), G(g["csc"])[2:])


def test_conversions_golden(goldens):
    g = goldens["csc_to_csr"]
    csc = sp.CscMatrix.new(*G(g["csc"]))
    same(arrays(sp.CsrMatrix.from_csc(csc)), G(g["csr"])[2:])
    g = goldens["csr_to_csc"]
    csr = sp.CsrMatrix.new(*G(g["csr"]))
    same(arrays(sp.CscMatrix.from_csr(csr)), G(g["csc"])[2:])


def test_transpose_golden(goldens):
    t = sp.CsrMatrix.new(*G(goldens["csr_transpose"]["in"])).transpose()
    same(arrays(t), G(goldens["csr_transpose"]["out"])[2:])
    t = sp.CscMatrix.new(*G(goldens["csc_transpose"]["in"])).transpose()
    same(arrays(t), G(goldens["csc_transpose"]["out"])[2:])


@pytest.mark.parametrize("name", ["csr_add", "csr_sub", "csc_add", "csc_sub"])
def test_add_sub_golden(goldens, name):
    g = goldens[name]
    cls = sp.CsrMatrix if name.startswith("csr") else sp.CscMatrix
    lhs, rhs = cls.new(*G(g["lhs"])), cls.new(*G(g["rhs"]))
    mat = lhs + rhs if name.endswith("add") else lhs - rhs
    assert (mat.nrows(), mat.ncols()) == (4, 4)
    same(arrays(mat), G(g["out"])[2:])
    assert len(arrays(mat)[1]) == mat.nnz() and len(arrays(mat)[2]) == mat.nnz()   # exactly sized


def test_mul_golden(goldens):
    """src/csc/ops/mul.rs:68-95; CSR Mul pinned through it."""
    g = goldens["csc_mul"]
    lhs, rhs = sp.CscMatrix.new(*G(g["lhs"])), sp.CscMatrix.new(*G(g["rhs"]))
    mat = lhs * rhs
    assert (mat.nrows(), mat.ncols()) == (5, 4)
    same(arrays(mat), G(g["out"])[2:])
    mat_csr = lhs.to_csr() * rhs.to_csr()
    same(arrays(mat_csr.to_csc()), G(g["out"])[2:])


def test_neg_golden(goldens):
    n = -sp.CsrMatrix.new(*G(goldens["csr_neg"]["in"]))
    same(arrays(n), G(goldens["csr_neg"]["out"])[2:])
    n = -sp.CscMatrix.new(*G(goldens["csc_neg"]["in"]))
    same(arrays(n), G(goldens["csc_neg"]["out"])[2:])


def test_new_panics(goldens):
    """src/csr.rs:470-510, src/csc.rs:470-510: the 7+7 #[should_panic] constructions."""
    for cls, key, major in ((sp.CsrMatrix, "csr_new_panics", "row"), (sp.CscMatrix, "csc_new_panics", "col")):
        for name, (nr, nc, ptr, ind, val) in goldens[key]["cases"].items():
            with pytest.raises(sp.Panic):
                cls.new(nr, nc, ptr, ind, val)
            want = orc.validate_compressed(nr, nc, ptr, ind, len(val), major)
            assert sp.default_context().invalid_reason() == want, name
    for m in goldens["valid_constructions"]["csr"]:
        sp.CsrMatrix.new(*m)
    for m in goldens["valid_constructions"]["csc"]:
        sp.CscMatrix.new(*m)


def test_eye_and_accessors():
    e = sp.CsrMatrix.eye(5)
    assert e.rowptr().tolist() == list(range(6)) and e.colind().tolist() == list(range(5))
    assert e.values().tolist() == [1.0] * 5 and e.nnz() == 5 and e.shape() == (5, 5)
    with pytest.raises(sp.Panic):
        sp.CscMatrix.eye(0)
    coo = e.to_coo()
    assert list(coo.iter()) == [(i, i, 1.0) for i in range(5)]


def test_shape_mismatch_panics():
    a, b = sp.CsrMatrix.eye(3), sp.CsrMatrix.eye(4)
    for op in (lambda: a + b, lambda: a - b, lambda: a * b):
        with pytest.raises(sp.Panic):
            op()


# ------------------------------------------------------------------ differential vs the oracle
def _oracle_assemble(nr, nc, r, c, v, major, dedup=True, dropzero=True):
    return orc.compress_from_coo(nr, nc, orc.make_triplets(r, c, v), major, dedup, dropzero)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("shape,n", [((1, 1), 5), ((7, 5), 0), ((7, 5), 200), ((300, 4000), 20000),
                                     ((70000, 3), 50000), ((1, 90000), 40000), ((100000, 100000), 300000)])
def test_assembly_random(dtype, shape, n):
    rng = np.random.default_rng(hash((shape, n)) % 2**32)
    nr, nc = shape
    r, c, v = syn.random_coo(rng, nr, nc, n, dtype, dup_frac=0.3, cancel_frac=0.05)
    # heavy duplication of a few cells: long in-order sums with cancellation
    if n:
        hot = rng.integers(0, len(v), 3)
        extra = 500
        r = np.concatenate([r, np.repeat(r[hot], extra)])
        c = np.concatenate([c, np.repeat(c[hot], extra)])
        v = np.concatenate([v, (rng.standard_normal(3 * extra) * 10.0 ** rng.integers(-8, 8, 3 * extra)).astype(dtype)])
    coo = sp.CooMatrix.with_triplets(nr, nc, r, c, v)
    same(arrays(sp.CsrMatrix.from_coo(coo)), _oracle_assemble(nr, nc, r, c, v, "row"), "csr")
    same(arrays(sp.CscMatrix.from_coo(coo)), _oracle_assemble(nr, nc, r, c, v, "col"), "csc")


def test_assembly_special_values():
    ent = [(0, 0, -0.0), (0, 1, float("nan")), (1, 0, 1e-320), (1, 1, 1.0), (1, 1, -1.0),
           (2, 2, 1e-310), (2, 2, 1e-310), (2, 0, 1e308), (2, 0, 1e308), (2, 1, 0.1), (2, 1, 0.2), (2, 1, -0.3)]
    r = np.array([e[0] for e in ent], np.uint64)
    c = np.array([e[1] for e in ent], np.uint64)
    v = np.array([e[2] for e in ent], np.float64)
    coo = sp.CooMatrix.with_triplets(3, 3, r, c, v)
    same(arrays(sp.CsrMatrix.from_coo(coo)), _oracle_assemble(3, 3, r, c, v, "row"))
    same(arrays(sp.CscMatrix.from_coo(coo)), _oracle_assemble(3, 3, r, c, v, "col"))


def test_assembly_out_of_bounds_is_rejected():
    import ctypes as C
    ctx = sp.default_context()
    r = np.array([0, 5], np.uint64); c = np.array([0, 0], np.uint64); v = np.array([1.0, 2.0])
    h = C.c_void_p()
    st = ctx._lib.spl_mat_from_coo(ctx._h, 0, 1, 3, 3, 2, r.ctypes.data, c.ctypes.data, v.ctypes.data, 1, 1, C.byref(h))
    assert st == 6 and not h.value


def _rand_csr(rng, n, m, density, dtype):
    nnz_target = int(n * m * density)
    r = rng.integers(0, n, nnz_target)
    c = rng.integers(0, m, nnz_target)
    key = np.unique(r.astype(np.int64) * m + c)
    r, c = key // m, key % m
    ptr = np.concatenate([[0], np.cumsum(np.bincount(r, minlength=n))]).astype(np.uint64)
    val = rng.standard_normal(len(key)).astype(dtype)
    return ptr, c.astype(np.uint64), val


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("n,m,density", [(1, 1, 1.0), (5, 9, 0.4), (1000, 700, 0.01), (3, 50000, 0.2),
                                         (50000, 3, 0.2), (20000, 20000, 0.0008)])
def test_transpose_convert_random(dtype, n, m, density):
    rng = np.random.default_rng(n * 31 + m)
    a = _rand_csr(rng, n, m, density, dtype)
    csr = sp.CsrMatrix.new(n, m, *a)
    want = orc.recompress(n, m, *a)
    t = csr.transpose()
    assert t.shape() == (m, n)
    same(arrays(t), want, "transpose")
    csc = csr.to_csc()
    assert csc.shape() == (n, m)
    same(arrays(csc), want, "csr->csc")
    same(arrays(csc.to_csr()), a, "csc->csr round trip")
    same(arrays(t.transpose()), a, "transpose twice")
    # CSC transpose: arrays regrouped by row
    same(arrays(csc.transpose()), orc.recompress(m, n, *want), "csc transpose")


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("n,m,density", [(4, 4, 0.5), (300, 300, 0.02), (200, 350, 0.03), (40000, 40000, 0.0002)])
def test_add_sub_random(dtype, n, m, density):
    rng = np.random.default_rng(n + m)
    a = _rand_csr(rng, n, m, density, dtype)
    b = _rand_csr(rng, n, m, density, dtype)
    # make some overlapping entries cancel exactly: explicit zeros must be kept
    A, B = sp.CsrMatrix.new(n, m, *a), sp.CsrMatrix.new(n, m, *b)
    same(arrays(A + B), orc.addsub(0, n, m, a, b), "add")
    same(arrays(A - B), orc.addsub(1, n, m, a, b), "sub")
    z = A - A
    assert z.nnz() == A.nnz() and not z.values().any()                   # a + (-a) stores 0.0
    Ac, Bc = A.to_csc(), B.to_csc()
    ac, bc = arrays(Ac), arrays(Bc)
    same(arrays(Ac + Bc), orc.addsub(0, m, n, ac, bc), "csc add")
    same(arrays(Ac - Bc), orc.addsub(1, m, n, ac, bc), "csc sub")


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("n,k,m,density", [(5, 3, 4, 0.6), (60, 80, 50, 0.1), (2000, 1500, 1800, 0.004),
                                           (400, 300, 600, 0.05),    # 128..1024 products per row: large hash table
                                           (300, 400, 500, 0.08)])   # > 1024 products per row: sort path
def test_mul_random(dtype, n, k, m, density):
    rng = np.random.default_rng(n * k + m)
    a = _rand_csr(rng, n, k, density, dtype)
    b = _rand_csr(rng, k, m, density, dtype)
    A, B = sp.CsrMatrix.new(n, k, *a), sp.CsrMatrix.new(k, m, *b)
    same(arrays(A * B), orc.csr_mul(n, k, m, a, b), "csr mul")
    Ac, Bc = A.to_csc(), B.to_csc()
    same(arrays(Ac * Bc), orc.csr_mul(m, k, n, arrays(Bc), arrays(Ac)), "csc mul")


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("fmt", ["csr", "csc"])
def test_mul_by_n_by_1_is_the_reference_spmv(dtype, fmt):
    """`&A * &X` with X n x 1 — the reference's only SpMV route (src/csr/ops/mul.rs:5-60): rows of A
    without a stored entry are absent from the result, values are the sequential ascending-k sums,
    bit for bit (the device skips the sort here: products are already in row order)."""
    rng = np.random.default_rng(77)
    n, k = 3000, 2500
    a = _rand_csr(rng, n, k, 0.004, dtype)
    xs = rng.standard_normal(k).astype(dtype)
    present = rng.random(k) < 0.9                                  # some rows of X are empty
    xptr = np.concatenate([[0], np.cumsum(present)]).astype(np.uint64)
    x = (xptr, np.zeros(int(present.sum()), np.uint64), xs[present])
    want = orc.csr_mul(n, k, 1, a, x)
    A, X = sp.CsrMatrix.new(n, k, *a), sp.CsrMatrix.new(k, 1, *x)
    if fmt == "csr":
        same(arrays(A * X), want, "csr A * X")
    else:
        got = A.to_csc() * X.to_csc()
        same(arrays(got.to_csr()), want, "csc A * X")


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("fmt", ["row", "col"])
def test_assembly_of_sorted_triplets_skips_the_sort(dtype, fmt):
    """Triplets already ordered by (major, minor) — duplicates adjacent, in insertion order — take
    the no-sort path; one element out of place takes the sort.  Both must equal the oracle bit for
    bit (in-order duplicate sums, cancellations dropped)."""
    rng = np.random.default_rng(12)
    n, m, length = 700, 900, 30000
    r = rng.integers(0, n, length).astype(np.uint64)
    c = rng.integers(0, m, length).astype(np.uint64)
    v = rng.standard_normal(length).astype(dtype)
    r = np.concatenate([r, r[:3000], r[:1000]]); c = np.concatenate([c, c[:3000], c[:1000]])
    v = np.concatenate([v, rng.standard_normal(3000).astype(dtype), -v[:1000]])
    order = np.lexsort((c, r), ) if fmt == "row" else np.lexsort((r, c))     # stable: duplicates keep their order
    r, c, v = r[order], c[order], v[order]
    cls = sp.CsrMatrix if fmt == "row" else sp.CscMatrix
    for variant in ("sorted", "one_swap"):
        if variant == "one_swap":
            r, c, v = r.copy(), c.copy(), v.copy()
            for a in (r, c, v):
                a[[10, 20000]] = a[[20000, 10]]
        got = cls.from_coo(sp.CooMatrix.with_triplets(n, m, r, c, v))
        want = orc.compress_from_coo(n, m, orc.make_triplets(r, c, v), fmt)
        same(arrays(got), want, f"{fmt} {variant}")


def test_values_mut_round_trip():
    """values_mut() (src/csr.rs:270-272): new values, same structure, seen by the next operation."""
    A = sp.CsrMatrix.new(2, 3, np.array([0, 2, 3], np.uint64), np.array([0, 2, 1], np.uint64), np.array([1.0, 2.0, 3.0]))
    with A.values_mut() as v:
        v *= 10.0
    assert A.values().tolist() == [10.0, 20.0, 30.0] and A.colind().tolist() == [0, 2, 1]
    assert A.matvec(np.array([1.0, 1.0, 1.0])).tolist() == [30.0, 30.0]
    with pytest.raises(sp.Panic):
        A.set_values(np.zeros(2))


def test_long_runs_of_empty_rows_and_columns():
    """A handful of entries in a 5e6 x 3e6 matrix: the pointer arrays are almost entirely runs of
    empty rows / columns (written by the queued gap filler, not by one thread)."""
    n, m = 5_000_000, 3_000_000
    r = np.array([7, 7, 2_500_000, n - 1, n - 1], np.uint64)
    c = np.array([5, 2_999_999, 0, 123, 123], np.uint64)
    v = np.array([1.0, 2.0, 3.0, 4.0, 0.5])
    coo = sp.CooMatrix.with_triplets(n, m, r, c, v)
    trip = orc.make_triplets(r, c, v)
    A = sp.CsrMatrix.from_coo(coo)
    want = orc.compress_from_coo(n, m, trip, "row")
    same(arrays(A), want, "csr")
    same(arrays(sp.CscMatrix.from_coo(coo)), orc.compress_from_coo(n, m, trip, "col"), "csc")
    same(arrays(A.transpose()), orc.recompress(n, m, *want), "transpose")
    same(arrays(A.to_csc()), orc.recompress(n, m, *want), "to_csc")
    same(arrays(A + A), orc.addsub(0, n, m, want, want), "add")


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("fmt,longest", [("row", 12), ("row", 40), ("col", 60), ("row", 200)])
def test_assembly_of_major_sorted_triplets(dtype, fmt, longest):
    """Triplets that come row by row (column by column for CSC) with the minor indices in arbitrary
    order and duplicates inside the segments: the per-segment sort route (segments up to 64) and, for
    the 200-entry case, the radix route must both equal the oracle bit for bit — duplicates summed
    in insertion order."""
    rng = np.random.default_rng(longest)
    n, m = 3000, 2500
    nmaj, nmin = (n, m) if fmt == "row" else (m, n)
    lens = rng.integers(0, longest + 1, nmaj)
    lens[5] = longest                                         # the longest segment decides the route
    maj = np.repeat(np.arange(nmaj, dtype=np.uint64), lens)
    mino = rng.integers(0, min(nmin, 3 * longest), len(maj)).astype(np.uint64)     # many duplicates per segment
    v = rng.standard_normal(len(maj)).astype(dtype)
    v[::7] = -v[1::7][: len(v[::7])] if len(v[1::7]) >= len(v[::7]) else v[::7]
    r, c = (maj, mino) if fmt == "row" else (mino, maj)
    cls = sp.CsrMatrix if fmt == "row" else sp.CscMatrix
    got = cls.from_coo(sp.CooMatrix.with_triplets(n, m, r, c, v))
    want = orc.compress_from_coo(n, m, orc.make_triplets(r, c, v), fmt)
    same(arrays(got), want, f"{fmt} longest {longest}")


def test_dok_round_trip_through_device():
    """From<&DokMatrix> for CsrMatrix / CscMatrix (src/csr/conv/dok.rs:3-76) and back
    (src/dok.rs:676-720): explicit zeros survive both ways, nothing is summed or dropped."""
    rng = np.random.default_rng(4)
    dok = sp.DokMatrix.new(40, 30)
    for _ in range(300):
        dok.insert(int(rng.integers(40)), int(rng.integers(30)), float(rng.standard_normal()))
    dok.insert(3, 3, 0.0)
    for cls, back in ((sp.CsrMatrix, sp.DokMatrix.from_csr), (sp.CscMatrix, sp.DokMatrix.from_csc)):
        m = cls.from_dok(dok)
        assert m.nnz() == dok.length()
        d2 = back(m)
        assert d2.shape() == dok.shape() and d2._map == dok._map


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_neg_random(dtype):
    rng = np.random.default_rng(11)
    a = _rand_csr(rng, 500, 400, 0.05, dtype)
    a[2][:3] = [0.0, -0.0, np.nan]
    N = -sp.CsrMatrix.new(500, 400, *a)
    got = arrays(N)
    assert np.array_equal(got[0], a[0]) and np.array_equal(got[1], a[1])
    assert got[2].tobytes() == orc.neg(a[2]).tobytes()


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("n,m,density", [(1, 1, 1.0), (64, 64, 0.3), (5000, 4000, 0.004), (3000, 3000, 0.05),
                                         (100000, 100000, 0.00005)])
def test_spmv_random(dtype, n, m, density):
    rng = np.random.default_rng(n ^ m)
    a = _rand_csr(rng, n, m, density, dtype)
    x = rng.standard_normal(m).astype(dtype)
    A = sp.CsrMatrix.new(n, m, *a)
    want = orc.csr_spmv(n, *a, x)
    # |y - y_ref| <= rtol * sum |a||x| (componentwise backward-error bound; rows with cancellation)
    scale = orc.csr_spmv(n, a[0], a[1], np.abs(a[2]), np.abs(x))
    rtol = SPMV_RTOL[np.dtype(dtype)]
    got = A.matvec(x)
    assert np.all(np.abs(got - want) <= rtol * np.maximum(scale, np.finfo(dtype).tiny)), "auto"
    assert np.all(got[np.diff(a[0].astype(np.int64)) == 0] == 0)          # empty rows give 0
    import torch
    xd = torch.from_numpy(x).cuda()
    for kernel, lanes in [(1, l) for l in (1, 2, 4, 8, 16, 32)] + [(2, 0), (3, 0), (5, 0)]:   # vector x6, merge, split, stream
        yd = torch.full((n,), 7.0, dtype=xd.dtype, device="cuda")
        torch.cuda.synchronize()
        A.spmv_device(xd.data_ptr(), yd.data_ptr(), kernel=kernel, lanes=lanes)
        sp.default_context().sync()
        got = yd.cpu().numpy()
        assert np.all(np.abs(got - want) <= rtol * np.maximum(scale, np.finfo(dtype).tiny)), (kernel, lanes)
        if kernel == 5 and A.spmv_choice()[1] == 1:       # stream kernel, one lane per row: ascending-column sum
            assert got.tobytes() == want.tobytes(), "stream (1 lane/row): not bit-identical to the sequential row sum"
    # sliced kernel (slices of 32 rows, lane per row): ascending-column sum, i.e. the reference's
    # `&A * &X` bit for bit; the copy is refused when the padding would exceed 4x
    lens = np.diff(a[0].astype(np.int64))
    padded = sum(32 * int(lens[i:i + 32].max()) for i in range(0, n, 32))
    yd = torch.full((n,), 7.0, dtype=xd.dtype, device="cuda")
    torch.cuda.synchronize()
    if padded <= 4 * len(a[1]) + 4096 and len(a[1]):
        A.spmv_device(xd.data_ptr(), yd.data_ptr(), kernel=4)
        sp.default_context().sync()
        assert yd.cpu().numpy().tobytes() == want.tobytes(), "sliced: not bit-identical to the sequential row sum"
    else:
        with pytest.raises(sp.DeviceError):
            A.spmv_device(xd.data_ptr(), yd.data_ptr(), kernel=4)


def test_spmv_skewed_rows():
    """Power-law-like rows (one row holding most entries) + many empty rows."""
    rng = np.random.default_rng(5)
    n = m = 20000
    lens = np.zeros(n, np.int64)
    lens[7] = 15000; lens[100:200] = 300; lens[1000::7] = 3
    ptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    ind = np.concatenate([np.sort(rng.choice(m, l, replace=False)) for l in lens if l]).astype(np.uint64)
    val = rng.standard_normal(len(ind))
    x = rng.standard_normal(m)
    A = sp.CsrMatrix.new(n, m, ptr, ind, val)
    want = orc.csr_spmv(n, ptr, ind, val, x)
    scale = orc.csr_spmv(n, ptr, ind, np.abs(val), np.abs(x))
    assert np.all(np.abs(A.matvec(x) - want) <= 1e-12 * np.maximum(scale, 1e-300))
    assert A.spmv_choice()[0] == 3                       # skewed rows select the balanced nnz-split kernel
    import torch
    xd = torch.from_numpy(x).cuda()
    for kernel, lanes in ((1, 32), (2, 0), (3, 0)):
        yd = torch.full((n,), 7.0, dtype=xd.dtype, device="cuda")
        torch.cuda.synchronize()
        A.spmv_device(xd.data_ptr(), yd.data_ptr(), kernel=kernel, lanes=lanes)
        sp.default_context().sync()
        assert np.all(np.abs(yd.cpu().numpy() - want) <= 1e-12 * np.maximum(scale, 1e-300)), kernel


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_spmv_split_long_runs_and_edges(dtype):
    """nnz-split kernel: a row spanning > 32 chunks (CTA fix-up), a row ending exactly on a chunk
    boundary, leading/trailing empty rows, rows of 33..200 entries (whole-warp sums), inf in x."""
    import torch
    rng = np.random.default_rng(9)
    n = m = 30000
    K = 128 if dtype == np.float64 else 256
    lens = np.zeros(n, np.int64)
    lens[5] = K - 3; lens[6] = 3; lens[7] = 40 * K + 17; lens[8] = 2 * K; lens[50:80] = rng.integers(33, 200, 30)
    lens[300:20000:3] = rng.integers(0, 9, len(lens[300:20000:3]))
    ptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    ind = np.concatenate([np.sort(rng.choice(m, l, replace=False)) for l in lens if l]).astype(np.uint64)
    val = rng.standard_normal(len(ind)).astype(dtype)
    x = rng.standard_normal(m).astype(dtype)
    A = sp.CsrMatrix.new(n, m, ptr, ind, val)
    want = orc.csr_spmv(n, ptr, ind, val, x)
    scale = orc.csr_spmv(n, ptr, ind, np.abs(val), np.abs(x))
    rtol = SPMV_RTOL[np.dtype(dtype)]
    xd = torch.from_numpy(x).cuda()
    yd = torch.full((n,), 7.0, dtype=xd.dtype, device="cuda")
    torch.cuda.synchronize()
    A.spmv_device(xd.data_ptr(), yd.data_ptr(), kernel=3)
    sp.default_context().sync()
    got = yd.cpu().numpy()
    assert np.all(np.abs(got - want) <= rtol * np.maximum(scale, np.finfo(dtype).tiny))
    assert np.all(got[lens == 0] == 0)
    # a column holding +inf must give inf (never 0 * inf = nan from a masked slot) in the rows using it
    col = int(ind[int(ptr[7]) + 3])
    x2 = x.copy(); x2[col] = np.inf
    for kernel, lanes in ((3, 0), (1, 4), (1, 32)):
        yd.fill_(7.0)
        xd2 = torch.from_numpy(x2).cuda()
        torch.cuda.synchronize()
        A.spmv_device(xd2.data_ptr(), yd.data_ptr(), kernel=kernel, lanes=lanes)
        sp.default_context().sync()
        got = yd.cpu().numpy()
        w2 = orc.csr_spmv(n, ptr, ind, val, x2)
        assert np.array_equal(np.isnan(got), np.isnan(w2)) and np.array_equal(np.isinf(got), np.isinf(w2)), kernel


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("n,m,density", [(1, 1, 1.0), (37, 91, 0.2), (4000, 2500, 0.01), (120000, 90000, 0.0001)])
def test_to_coo_random_matches_the_oracle(dtype, n, m, density):
    """From<&CsrMatrix> / From<&CscMatrix> for CooMatrix (src/coo.rs:629-705): storage-order expansion
    (row of every entry from rowptr, src/csr.rs:303-316), compared entry by entry with the oracle."""
    rng = np.random.default_rng(n * 7 + m)
    a = _rand_csr(rng, n, m, density, dtype)
    A = sp.CsrMatrix.new(n, m, *a)
    for M, major, nmaj in ((A, "row", n), (A.to_csc(), "col", m)):
        arr = arrays(M)
        want = orc.expand_to_coo(nmaj, arr[0], arr[1], arr[2], major)
        r, c, v = M.to_coo().triplets()
        assert r.dtype == np.uint64 and c.dtype == np.uint64 and len(v) == M.nnz()
        assert np.array_equal(r, want["row"]) and np.array_equal(c, want["col"]), major
        assert v.tobytes() == want["val"].tobytes(), major
        # iter(): the same entries, read chunk by chunk from the device
        M.ITER_CHUNK = 1000                                   # several chunks, a ragged last one
        got = list(M.iter())
        assert len(got) == M.nnz()
        if got:
            gr, gc, gv = (np.array(t) for t in zip(*got))
            assert np.array_equal(gr.astype(np.uint64), want["row"]) and np.array_equal(gc.astype(np.uint64), want["col"])
            assert gv.astype(dtype).tobytes() == want["val"].tobytes()
        # the device-side form: uint32 SoA triplets left in HBM, fed straight back into the assembly
        import torch
        nz = max(M.nnz(), 1)
        rd = torch.empty(nz, dtype=torch.int32, device="cuda")
        cd = torch.empty(nz, dtype=torch.int32, device="cuda")
        vd = torch.empty(nz, dtype=torch.float32 if dtype == np.float32 else torch.float64, device="cuda")
        torch.cuda.synchronize()
        M.to_coo_device(rd.data_ptr(), cd.data_ptr(), vd.data_ptr())
        sp.default_context().sync()
        k = M.nnz()
        assert np.array_equal(rd.cpu().numpy()[:k].astype(np.uint64), want["row"])
        assert np.array_equal(cd.cpu().numpy()[:k].astype(np.uint64), want["col"])
        assert vd.cpu().numpy()[:k].tobytes() == want["val"].tobytes()
        again = type(M).from_device_triplets(n, m, k, rd.data_ptr(), cd.data_ptr(), vd.data_ptr(), dtype)
        same(arrays(again), arr, f"device round trip through COO ({major})")
        # and back: CooMatrix -> the same compressed matrix (round trip through the builder format)
        back = type(M).from_coo(M.to_coo())
        same(arrays(back), arr, f"round trip through COO ({major})")       # values are non-zero: nothing is dropped


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("n,m,density", [(1, 1, 1.0), (700, 450, 0.02), (30000, 41000, 0.0004)])
def test_spmv_on_csc_matrix(dtype, n, m, density):
    """`&A * &X` with A a CscMatrix (src/csc/ops/mul.rs:5-61; X n x 1): y[i] accumulates over ascending k
    like the row-wise product.  spl_spmv on a CSC matrix runs on its cached CSR form; the scatter
    kernel is the copy-free variant (atomic adds, tolerance only); the CSC view of a CSR matrix B
    (same arrays, dims swapped) gives y = B^T x; values_mut drops the cached form."""
    import torch
    rng = np.random.default_rng(n + 3 * m)
    a = _rand_csr(rng, n, m, density, dtype)
    x = rng.standard_normal(m).astype(dtype)
    A = sp.CsrMatrix.new(n, m, *a)
    Cm = A.to_csc()
    want = orc.csr_spmv(n, *a, x)
    scale = orc.csr_spmv(n, a[0], a[1], np.abs(a[2]), np.abs(x))
    rtol = SPMV_RTOL[np.dtype(dtype)]
    tiny = np.finfo(dtype).tiny
    y = Cm.matvec(x)
    assert y.tobytes() == A.matvec(x).tobytes()                 # same kernel on the same CSR arrays
    assert np.all(np.abs(y - want) <= rtol * np.maximum(scale, tiny))
    assert Cm.spmv_choice() == A.spmv_choice()
    xd = torch.from_numpy(x).cuda()
    yd = torch.full((n,), 7.0, dtype=xd.dtype, device="cuda")
    torch.cuda.synchronize()
    Cm.spmv_device(xd.data_ptr(), yd.data_ptr(), kernel=6)      # SPL_SPMV_SCATTER
    sp.default_context().sync()
    assert np.all(np.abs(yd.cpu().numpy() - want) <= 4 * rtol * np.maximum(scale, tiny))
    with pytest.raises(sp.DeviceError):
        A.spmv_device(xd.data_ptr(), yd.data_ptr(), kernel=6)   # the column kernel needs a CSC matrix
    # y = B^T z through the CSC view of B's own arrays
    z = rng.standard_normal(n).astype(dtype)
    Bt = sp.CscMatrix.new(m, n, *a)                             # m x n matrix whose columns are B's rows
    t = orc.recompress(n, m, *a)                                # CSR arrays of B^T
    wt = orc.csr_spmv(m, *t, z)
    st = orc.csr_spmv(m, t[0], t[1], np.abs(t[2]), np.abs(z))
    assert np.all(np.abs(Bt.matvec(z) - wt) <= rtol * np.maximum(st, tiny))
    # values_mut: the cached CSR form follows the new values
    if len(a[2]):
        v2 = (2.0 * a[2]).astype(dtype)
        Bt.set_values(v2)
        assert np.all(np.abs(Bt.matvec(z) - 2.0 * wt) <= 2 * rtol * np.maximum(2 * st, tiny))


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("shape", ["laplace", "band9", "stencil27", "ragged"])
def test_spmv_stream_kernel(dtype, shape):
    """The persistent TMA-pipelined kernel (SPL_SPMV_STREAM): row counts that are not a multiple of the
    tile, tiles without entries, more tiles than CTAs (the ring wraps many times), and a chain of
    products y_t -> x_{t+1} launched back to back (programmatic dependent launch between them)
    against the oracle's sequential iteration.  One lane per row => bit-identical to `&A * &X`."""
    import torch
    rng = np.random.default_rng(11)
    if shape == "laplace":
        g = 301
        r, c, v = syn.laplacian_2d(g)
        n = g * g
        order = np.lexsort((c, r))
        a = (syn.csr_from_sorted_triplets(n, r, c, v)[0], c[order], v[order])
    elif shape == "band9":
        n = 700_001
        rows = np.repeat(np.arange(n), 9)
        cols = rows + np.tile(np.arange(-4, 5), n)
        ok = (cols >= 0) & (cols < n)
        rows, cols = rows[ok].astype(np.uint64), cols[ok].astype(np.uint64)
        a = (np.concatenate([[0], np.cumsum(np.bincount(rows.astype(np.int64), minlength=n))]).astype(np.uint64), cols,
             rng.standard_normal(len(cols)))
    elif shape == "stencil27":
        m_ = 40
        r, c, v = syn.stencil_27(m_)
        n = m_ ** 3
        a = syn.csr_from_sorted_triplets(n, r, c, v)
    else:                                                       # empty tiles, rows of 0..40 entries, a ragged end
        n = 70_003
        lens = rng.integers(0, 41, n)
        lens[1000:3000] = 0
        lens[-5:] = 0
        ptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
        ind = np.concatenate([np.sort(rng.choice(n, l, replace=False)) for l in lens if l]).astype(np.uint64)
        a = (ptr, ind, rng.standard_normal(len(ind)))
    a = (a[0], a[1], (np.asarray(a[2]) * 0.05).astype(dtype))   # scaled: the chain below neither blows up nor dies
    A = sp.CsrMatrix.new(n, n, *a)
    lanes = A.spmv_choice()[1]
    x0 = rng.standard_normal(n).astype(dtype)
    rtol = SPMV_RTOL[np.dtype(dtype)]
    xd = torch.from_numpy(x0).cuda()
    yd = torch.full((n,), 7.0, dtype=xd.dtype, device="cuda")
    torch.cuda.synchronize()
    A.spmv_device(xd.data_ptr(), yd.data_ptr(), kernel=5)
    sp.default_context().sync()
    want = orc.csr_spmv(n, *a, x0)
    scale = orc.csr_spmv(n, a[0], a[1], np.abs(a[2]), np.abs(x0))
    got = yd.cpu().numpy()
    assert np.all(np.abs(got - want) <= rtol * np.maximum(scale, np.finfo(dtype).tiny))
    if lanes == 1:
        assert got.tobytes() == want.tobytes()
    # chain: 12 products back to back on two buffers, no host synchronisation in between
    bufs = [xd.clone(), torch.empty_like(xd)]
    torch.cuda.synchronize()
    for it in range(12):
        A.spmv_device(bufs[it % 2].data_ptr(), bufs[(it + 1) % 2].data_ptr(), kernel=5)
    sp.default_context().sync()
    ref = x0
    for it in range(12):
        ref = orc.csr_spmv(n, *a, ref)
    got = bufs[0].cpu().numpy()
    if lanes == 1:
        assert got.tobytes() == ref.tobytes(), "chained stream products differ from the sequential iteration"
    else:
        assert np.allclose(got, ref, rtol=1e3 * rtol, atol=1e3 * rtol * float(np.abs(ref).max() + 1e-30))


# ------------------------------------------------------------------ BASELINE shapes, properties
def test_c1_laplacian_assembly_and_spmv():
    """Config 1 at full size: shuffled COO -> CSR equals the generator's canonical CSR bit for
    bit; A*1 equals the analytic row sums; CSR->CSC->CSR is the identity; A^T == A."""
    g = 1024
    r, c, v = syn.laplacian_2d(g)
    n = g * g
    want = syn.csr_from_sorted_triplets(n, r, c, v)          # emission order is not column-sorted:
    order = np.lexsort((c, r))
    want = (want[0], c[order], v[order])
    perm = np.random.default_rng(42).permutation(len(v))
    A = sp.CsrMatrix.from_coo(sp.CooMatrix.with_triplets(n, n, r[perm], c[perm], v[perm]))
    same(arrays(A), want, "C1 assembly")
    y = A.matvec(np.ones(n))
    deg = np.diff(want[0].astype(np.int64))
    assert np.array_equal(y, 4.0 - (deg - 1))               # exact in f64: small integers
    same(arrays(A.to_csc().to_csr()), want, "round trip")
    same(arrays(A.transpose()), want, "symmetric")


def test_c2_stencil_transpose_properties():
    """Config 2 structure at a reduced grid for the oracle (48^3) and full-size properties (128^3)."""
    m = 48
    r, c, v = syn.stencil_27(m)
    n = m ** 3
    a = syn.csr_from_sorted_triplets(n, r, c, v)
    A = sp.CsrMatrix.new(n, n, *a)
    same(arrays(A.to_csc()), orc.recompress(n, n, *a), "C2/48 csr->csc")
    x = 1.0 / (1.0 + (np.arange(n) % 97))
    want = orc.csr_spmv(n, *a, x)
    assert np.all(np.abs(A.matvec(x) - want) <= 1e-12 * orc.csr_spmv(n, a[0], a[1], np.abs(a[2]), np.abs(x)))


def test_c2_full_size_round_trip():
    m = 128
    r, c, v = syn.stencil_27(m)
    n = m ** 3
    assert len(v) == 382 ** 3
    a = syn.csr_from_sorted_triplets(n, r, c, v)
    del r
    A = sp.CsrMatrix.new(n, n, *a)
    C_ = A.to_csc()
    same(arrays(C_.to_csr()), a, "C2 csr->csc->csr")
    same(arrays(A.transpose()), a, "C2 symmetric pattern and values")
    y = A.matvec(np.ones(n))
    deg = np.diff(a[0].astype(np.int64))
    assert np.array_equal(y, 26.0 - (deg - 1))


# ------------------------------------------------------------------ boundary sizes
@pytest.mark.parametrize("length", [1, 2, 31, 32, 33, 255, 256, 257, 4095, 4096, 4097, 8191, 8193, 12288, 65535,
                                    65537, 4096 * 37 + 5])
def test_tile_boundary_lengths(length):
    """COO lengths straddling the warp, tile (4096) and multi-tile boundaries of the sort and tail
    kernels: assembly (both formats, f32), transpose, add and every SpMV kernel against the oracle."""
    import torch
    rng = np.random.default_rng(length)
    n, m = max(2, length // 7 + 3), max(2, length // 5 + 2)
    r = rng.integers(0, n, length).astype(np.uint64)
    c = rng.integers(0, m, length).astype(np.uint64)
    v = rng.standard_normal(length).astype(np.float32)
    if length > 4:
        r[-2:], c[-2:] = r[:2], c[:2]                     # duplicates across the whole list
        v[-1] = -v[1]
    trip = orc.make_triplets(r, c, v)
    coo = sp.CooMatrix.with_triplets(n, m, r, c, v)
    A = sp.CsrMatrix.from_coo(coo)
    want = orc.compress_from_coo(n, m, trip, "row")
    same(arrays(A), want, "csr")
    same(arrays(sp.CscMatrix.from_coo(coo)), orc.compress_from_coo(n, m, trip, "col"), "csc")
    same(arrays(A.transpose()), orc.recompress(n, m, *want), "transpose")
    same(arrays(A + A), orc.addsub(0, n, m, want, want), "add")
    x = rng.standard_normal(m).astype(np.float32)
    yw = orc.csr_spmv(n, *want, x)
    sc = orc.csr_spmv(n, want[0], want[1], np.abs(want[2]), np.abs(x))
    xd = torch.from_numpy(x).cuda()
    for kernel, lanes in ((1, 1), (1, 4), (1, 32), (2, 0), (3, 0), (5, 0)):
        yd = torch.full((n,), 7.0, dtype=torch.float32, device="cuda")
        torch.cuda.synchronize()
        A.spmv_device(xd.data_ptr(), yd.data_ptr(), kernel=kernel, lanes=lanes)
        sp.default_context().sync()
        assert np.all(np.abs(yd.cpu().numpy() - yw) <= 1e-5 * np.maximum(sc, 1e-30)), (kernel, lanes)


# ------------------------------------------------------------------ pinned, streaming CooMatrix (8f-4)
def test_pinned_coo_golden(goldens):
    """The reference's COO->CSR/CSC test (src/csr/conv/coo.rs:129-145) through the streamed storage."""
    g = goldens["coo_pushes"]
    coo = sp.PinnedCooMatrix.new(g["nrows"], g["ncols"])
    for r, c, v in g["entries"]:
        coo.push(int(r), int(c), v)
    assert coo.length() == len(g["entries"]) and coo.get(0) == (1, 2, 5.0) and coo.get(99) is None
    csr = sp.CsrMatrix.from_coo(coo)
    assert csr.rowptr().tolist() == g["csr"][2] and csr.colind().tolist() == g["csr"][3]
    assert csr.values().tolist() == g["csr"][4]
    csc = sp.CscMatrix.from_coo(coo)                       # the builder stays valid
    assert csc.colptr().tolist() == g["csc"][2] and csc.rowind().tolist() == g["csc"][3]
    assert csc.values().tolist() == g["csc"][4]
    with pytest.raises(sp.Panic):
        coo.push(2, 0, 1.0)                                # src/coo.rs:432
    with pytest.raises(sp.Panic):
        coo.push(0, 3, 1.0)                                # src/coo.rs:433
    with pytest.raises(sp.Panic):
        coo.extend([(0, 0, 1.0), (5, 0, 1.0)])             # Extend asserts before storing anything
    assert coo.length() == len(g["entries"])
    with pytest.raises(sp.Panic):
        sp.PinnedCooMatrix.new(0, 3)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("fmt", ["row", "col"])
def test_pinned_coo_streams_chunks_and_matches_the_oracle(dtype, fmt):
    """More than two transfer chunks (2^19 each), growth from a small capacity, duplicates,
    cancellations; pop / clear below the streamed watermark; conversion twice with pushes between."""
    rng = np.random.default_rng(11)
    n, m, length = 3000, 2500, (1 << 20) + 300_001
    r = rng.integers(0, n, length).astype(np.uint64)
    c = rng.integers(0, m, length).astype(np.uint64)
    v = rng.standard_normal(length).astype(dtype)
    v[1000:2000] = -v[0:1000]; r[1000:2000] = r[0:1000]; c[1000:2000] = c[0:1000]
    cls = sp.CsrMatrix if fmt == "row" else sp.CscMatrix
    coo = sp.PinnedCooMatrix.with_capacity(n, m, 1000, dtype)
    cut = 700_003
    coo.extend_triplets(r[:cut], c[:cut], v[:cut])
    assert coo.streamed() == 1 << 19 and coo.length() == cut and coo.capacity() >= cut
    first = cls.from_coo(coo)
    same(arrays(first), orc.compress_from_coo(n, m, orc.make_triplets(r[:cut], c[:cut], v[:cut]), fmt), "first")
    assert coo.streamed() == cut
    for i in range(cut, cut + 5):                                    # single pushes after a conversion
        coo.push(int(r[i]), int(c[i]), v[i])
    coo.extend_triplets(r[cut + 5:], c[cut + 5:], v[cut + 5:])
    tr = coo.triplets()
    assert np.array_equal(tr[0], r) and np.array_equal(tr[1], c) and tr[2].tobytes() == v.tobytes()
    same(arrays(cls.from_coo(coo)), orc.compress_from_coo(n, m, orc.make_triplets(r, c, v), fmt), "all")
    # pop below the watermark, refill with other entries: the positions are sent again
    assert coo.pop() == (int(r[-1]), int(c[-1]), v[-1].item())
    keep = 600_000
    while coo.length() > length - 3:
        coo.pop()
    coo._check(coo._lib.spl_coo_truncate(coo._b, keep))
    assert coo.length() == keep and coo.streamed() == keep
    r2, c2, v2 = r[::-1][:200_000].copy(), c[::-1][:200_000].copy(), (v[::-1][:200_000] * 2).astype(dtype)
    coo.extend_triplets(r2, c2, v2)
    want = orc.compress_from_coo(n, m, orc.make_triplets(np.concatenate([r[:keep], r2]),
                                                       np.concatenate([c[:keep], c2]),
                                                       np.concatenate([v[:keep], v2])), fmt)
    same(arrays(cls.from_coo(coo)), want, "after truncate")
    coo.clear()
    assert coo.length() == 0 and coo.pop() is None
    empty = cls.from_coo(coo)
    assert empty.nnz() == 0 and arrays(empty)[0].tolist() == [0] * ((n if fmt == "row" else m) + 1)


def test_pinned_coo_with_triplets_equals_plain():
    r, c, v = syn.laplacian_2d(96)
    rng = np.random.default_rng(5)
    p = rng.permutation(len(v))
    a = sp.CsrMatrix.from_coo(sp.PinnedCooMatrix.with_triplets(96 * 96, 96 * 96, r[p], c[p], v[p]))
    b = sp.CsrMatrix.from_coo(sp.CooMatrix.with_triplets(96 * 96, 96 * 96, r[p], c[p], v[p]))
    same(arrays(a), arrays(b))


# ------------------------------------------------------------------ hybrid route (>= 2^22 triplets): tail fused into the block kernel
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("fmt", ["row", "col"])
def test_hybrid_assembly_gaps_duplicates_zeros(dtype, fmt):
    """5 M triplets: a band of 100 000 empty majors (empty blocks), a last block that is not full,
    cells with 2..5 duplicates (order-sensitive sums), exact cancellations, explicit zeros and three
    majors with 1 000 records each (ranked by counting instead of the per-thread insertion sort)."""
    rng = np.random.default_rng(21)
    n, m, length = 400_001, 70_001, 5_000_000
    if fmt == "col":
        n, m = m, n
    nmaj = n if fmt == "row" else m
    maj = rng.integers(0, nmaj - 100_000, length)
    maj = np.where(maj >= 100_000, maj + 100_000, maj)
    mnr = rng.integers(0, m if fmt == "row" else n, length)
    for heavy in (5, 250_017, nmaj - 1):                       # rows too long for the per-thread insertion sort
        at = rng.choice(length, 1000, replace=False)
        maj[at], mnr[at] = heavy, rng.integers(0, 300, 1000)   # ~3 records per cell
    v = (rng.standard_normal(length) * 10.0 ** rng.integers(-6, 6, length)).astype(dtype)
    src = rng.integers(0, length // 2, 400_000)                # duplicates of earlier cells
    dst = length // 2 + np.arange(400_000)
    maj[dst], mnr[dst] = maj[src], mnr[src]
    v[dst[:50_000]] = -v[src[:50_000]]                          # many of these cancel exactly
    v[rng.integers(0, length, 1000)] = 0.0
    p = rng.permutation(length)
    maj, mnr, v = maj[p].astype(np.uint64), mnr[p].astype(np.uint64), v[p]
    r, c = (maj, mnr) if fmt == "row" else (mnr, maj)
    cls = sp.CsrMatrix if fmt == "row" else sp.CscMatrix
    got = cls.from_coo(sp.CooMatrix.with_triplets(n, m, r, c, v))
    same(arrays(got), orc.compress_from_coo(n, m, orc.make_triplets(r, c, v), fmt), "dedup")
    # DOK route: unique keys, explicit zeros kept, nothing summed
    key = maj * np.uint64(1 << 20) + mnr
    _, first = np.unique(key, return_index=True)
    first = rng.permutation(first)
    assert len(first) >= 1 << 22
    r1, c1, v1 = r[first], c[first], v[first]
    h = sp.default_context()
    import ctypes as C
    out = C.c_void_p()
    h.check(h._lib.spl_mat_from_coo(h._h, cls._FORMAT, sp.matrix._dtype_code(v1.dtype), n, m, len(v1),
                                    r1.ctypes.data_as(C.c_void_p), c1.ctypes.data_as(C.c_void_p),
                                    v1.ctypes.data_as(C.c_void_p), 0, 0, C.byref(out)))
    same(arrays(cls._wrap(h, out)),
         orc.compress_from_coo(n, m, orc.make_triplets(r1, c1, v1), fmt, dedup=False, dropzero=False), "dok")


@pytest.mark.parametrize("fmt", ["row", "col"])
def test_hybrid_assembly_big_blocks(fmt, monkeypatch):
    """The 8 192-record / 512-thread block kernel (f32, taken when it saves a global pass: 9 -> 8 high
    key bits here, 17 -> 16 on config 3), reached at an oracle-sized list by lowering the route's
    length threshold."""
    monkeypatch.setenv("SPL_HYBRID_MIN_LEN", "100000")
    rng = np.random.default_rng(33)
    nmaj, nmin, length = 65_536, 60_000, 1_310_720
    maj = rng.integers(0, nmaj, length)
    mnr = rng.integers(0, nmin, length)
    v = rng.standard_normal(length).astype(np.float32)
    src = rng.integers(0, length // 2, 100_000)
    dst = length // 2 + np.arange(100_000)
    maj[dst], mnr[dst] = maj[src], mnr[src]
    v[dst[:20_000]] = -v[src[:20_000]]
    p = rng.permutation(length)
    maj, mnr, v = maj[p].astype(np.uint64), mnr[p].astype(np.uint64), v[p]
    n, m = (nmaj, nmin) if fmt == "row" else (nmin, nmaj)
    r, c = (maj, mnr) if fmt == "row" else (mnr, maj)
    cls = sp.CsrMatrix if fmt == "row" else sp.CscMatrix
    got = cls.from_coo(sp.CooMatrix.with_triplets(n, m, r, c, v))
    same(arrays(got), orc.compress_from_coo(n, m, orc.make_triplets(r, c, v), fmt), "big blocks")
    # the same list in f64 takes the 4 096-record kernel
    got = cls.from_coo(sp.CooMatrix.with_triplets(n, m, r, c, v.astype(np.float64)))
    same(arrays(got), orc.compress_from_coo(n, m, orc.make_triplets(r, c, v.astype(np.float64)), fmt), "f64")


# ------------------------------------------------------------------ `&A * &x` with pinned host vectors: the pipelined path
@pytest.mark.parametrize("kind", ["laplace", "random", "tall"])
def test_spmv_host_pipelined_matches_device_product(kind):
    """spl_spmv_host with pinned x, y and >= 1 MB of rows runs in row chunks behind prefix uploads of
    x, downloads overlapping the next chunk: same kernel, so the bytes must equal the one-shot device
    product; and both must be within tolerance of the oracle."""
    import ctypes as C
    import torch
    rng = np.random.default_rng(17)
    if kind == "laplace":                                   # chunk c needs about a chunk of x more
        g = 600
        n = m = g * g
        r, c, v = syn.laplacian_2d(g)
    elif kind == "random":                                  # the first chunk already needs all of x
        n = m = 300_000
        r = np.repeat(np.arange(n, dtype=np.uint64), 6)
        c = rng.integers(0, m, len(r)).astype(np.uint64)
        v = rng.standard_normal(len(r))
    else:                                                   # more rows than columns, many empty rows
        n, m = 400_000, 1000
        r = np.sort(rng.integers(0, n, 900_000)).astype(np.uint64)
        c = rng.integers(0, m, len(r)).astype(np.uint64)
        v = rng.standard_normal(len(r))
    A = sp.CsrMatrix.from_coo(sp.CooMatrix.with_triplets(n, m, r, c, v))
    a = arrays(A)
    x = rng.standard_normal(m)
    hx = torch.from_numpy(x).pin_memory()
    hy = torch.full((n,), 7.0, dtype=torch.float64).pin_memory()
    ctx = sp.default_context()
    for _ in range(2):                                      # second call: plan cached, events reused
        hy.fill_(7.0)
        ctx.check(ctx._lib.spl_spmv_host(ctx._h, A._h, C.c_void_p(hx.data_ptr()), C.c_void_p(hy.data_ptr())))
        xd = hx.cuda()
        yd = torch.empty(n, dtype=torch.float64, device="cuda")
        torch.cuda.synchronize()
        A.spmv_device(xd.data_ptr(), yd.data_ptr())
        ctx.sync()
        assert torch.equal(hy, yd.cpu()), "pipelined host product differs from the device product"
    want = orc.csr_spmv(n, *a, x)
    scale = orc.csr_spmv(n, a[0], a[1], np.abs(a[2]), np.abs(x))
    assert np.all(np.abs(hy.numpy() - want) <= 1e-12 * np.maximum(scale, 1e-300))
    assert np.array_equal(A.matvec(x), hy.numpy())          # pageable vectors: the plain path, same result
    # the library's own page-locked vectors (spl_host_alloc) through the host mirror: the pipelined path again
    px, py = sp.pinned_empty(m, np.float64), sp.pinned_empty(n, np.float64)
    px[:] = x
    py[:] = 7.0
    assert A.matvec(px, out=py) is py and np.array_equal(py, hy.numpy())
    with pytest.raises(sp.Panic):
        A.matvec(px, out=sp.pinned_empty(n + 1, np.float64))
    del px, py


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_hybrid_assembly_wide_minor_index(dtype, monkeypatch):
    """2^27 columns: minor bits + in-block row bits exceed 32, so the block kernel keeps 64-bit keys in
    shared memory (its other instantiation); reached at an oracle-sized list through the threshold knob."""
    monkeypatch.setenv("SPL_HYBRID_MIN_LEN", "100000")
    rng = np.random.default_rng(44)
    n, m, length = 65_536, 1 << 27, 1_310_720
    r = rng.integers(0, n, length)
    c = rng.integers(0, m, length)
    v = rng.standard_normal(length).astype(dtype)
    src = rng.integers(0, length // 2, 100_000)
    dst = length // 2 + np.arange(100_000)
    r[dst], c[dst] = r[src], c[src]
    v[dst[:20_000]] = -v[src[:20_000]]
    p = rng.permutation(length)
    r, c, v = r[p].astype(np.uint64), c[p].astype(np.uint64), v[p]
    got = sp.CsrMatrix.from_coo(sp.CooMatrix.with_triplets(n, m, r, c, v))
    same(arrays(got), orc.compress_from_coo(n, m, orc.make_triplets(r, c, v), "row"), "wide minor")
