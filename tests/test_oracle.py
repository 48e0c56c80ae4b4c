"""The oracle (oracle/, CPU restatement) against every hot-path golden vector the
reference's own tests hold -- this is what pins parity."""
import numpy as np
import pytest

import oracle as orc


def M(m, dtype=np.float64):
    nrows, ncols, ptr, ind, val = m
    return nrows, ncols, np.array(ptr, np.uint64), np.array(ind, np.uint64), np.array(val, dtype)


def trip(entries, dtype=np.float64):
    e = np.array(entries, dtype=np.float64).reshape(-1, 3)
    return orc.make_triplets(e[:, 0].astype(np.uint64), e[:, 1].astype(np.uint64), e[:, 2].astype(dtype))


def same(got, want):
    _, _, ptr, ind, val = want
    assert got[0].tolist() == ptr.tolist()
    assert got[1].tolist() == ind.tolist()
    assert got[2].tobytes() == val.tobytes()      # bit-exact, like assert_eq! on f64


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_coo_to_csr_csc(goldens, dtype):
    g = goldens["coo_pushes"]
    t = trip(g["entries"], dtype)
    same(orc.compress_from_coo(g["nrows"], g["ncols"], t, "row"), M(g["csr"], dtype))
    same(orc.compress_from_coo(g["nrows"], g["ncols"], t, "col"), M(g["csc"], dtype))


def test_dok_to_csr_csc(goldens):
    g = goldens["dok_inserts"]
    for perm in ([0, 1, 2, 3], [3, 1, 0, 2], [2, 3, 1, 0]):   # HashMap order is arbitrary
        t = trip([g["entries"][i] for i in perm])
        same(orc.compress_from_coo(2, 3, t, "row", dedup=False, dropzero=False), M(g["csr"]))
        same(orc.compress_from_coo(2, 3, t, "col", dedup=False, dropzero=False), M(g["csc"]))


def test_dok_keeps_explicit_zero():
    # src/csr/conv/dok.rs has no zero-drop block
    t = trip([[0, 0, 0.0], [1, 1, 2.0]])
    ptr, ind, val = orc.compress_from_coo(2, 2, t, "row", dedup=False, dropzero=False)
    assert ptr.tolist() == [0, 1, 2] and ind.tolist() == [0, 1] and val.tolist() == [0.0, 2.0]


def test_conversions(goldens):
    g = goldens["csc_to_csr"]
    nr, nc, ptr, ind, val = M(g["csc"])
    same(orc.recompress(nc, nr, ptr, ind, val), M(g["csr"]))
    g = goldens["csr_to_csc"]
    nr, nc, ptr, ind, val = M(g["csr"])
    same(orc.recompress(nr, nc, ptr, ind, val), M(g["csc"]))


def test_transpose(goldens):
    nr, nc, ptr, ind, val = M(goldens["csr_transpose"]["in"])
    same(orc.recompress(nr, nc, ptr, ind, val), M(goldens["csr_transpose"]["out"]))
    nr, nc, ptr, ind, val = M(goldens["csc_transpose"]["in"])
    same(orc.recompress(nc, nr, ptr, ind, val), M(goldens["csc_transpose"]["out"]))


@pytest.mark.parametrize("name,sub,major", [("csr_add", 0, "row"), ("csr_sub", 1, "row"),
                                            ("csc_add", 0, "col"), ("csc_sub", 1, "col")])
def test_add_sub(goldens, name, sub, major):
    g = goldens[name]
    a, b = M(g["lhs"]), M(g["rhs"])
    nmajor, nminor = (a[0], a[1]) if major == "row" else (a[1], a[0])
    same(orc.addsub(sub, nmajor, nminor, a[2:], b[2:]), M(g["out"]))


def test_mul_csc_golden_and_csr_identity(goldens):
    g = goldens["csc_mul"]
    a, b, out = M(g["lhs"]), M(g["rhs"]), M(g["out"])
    # CSC(A*B) arrays == CSR(B^T * A^T) arrays: operands swapped, dims (bn, ak, an)
    same(orc.csr_mul(b[1], a[1], a[0], b[2:], a[2:]), out)
    # CSR route on the same matrices: convert CSC->CSR, multiply, convert back
    acsr = orc.recompress(a[1], a[0], *a[2:])
    bcsr = orc.recompress(b[1], b[0], *b[2:])
    ccsr = orc.csr_mul(a[0], a[1], b[1], acsr, bcsr)
    same(orc.recompress(a[0], b[1], *ccsr), out)


def test_neg(goldens):
    for name in ("csr_neg", "csc_neg"):
        g = goldens[name]
        assert orc.neg(M(g["in"])[4]).tobytes() == M(g["out"])[4].tobytes()
    assert np.signbit(orc.neg(np.array([0.0]))[0])          # 0.0 -> -0.0


def test_validation_panics(goldens):
    for major, key in (("row", "csr_new_panics"), ("col", "csc_new_panics")):
        for name, (nr, nc, ptr, ind, val) in goldens[key]["cases"].items():
            assert orc.validate_compressed(nr, nc, ptr, ind, len(val), major) != 0, name
    for m in goldens["valid_constructions"]["csr"]:
        assert orc.validate_compressed(m[0], m[1], m[2], m[3], len(m[4]), "row") == 0
    for m in goldens["valid_constructions"]["csc"]:
        assert orc.validate_compressed(m[0], m[1], m[2], m[3], len(m[4]), "col") == 0


def _rand_csr(rng, n, m, density, dtype):
    mask = rng.random((n, m)) < density
    ptr = np.concatenate([[0], np.cumsum(mask.sum(1))]).astype(np.uint64)
    ind = np.nonzero(mask)[1].astype(np.uint64)
    val = rng.standard_normal(len(ind)).astype(dtype)
    return ptr, ind, val


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_spmv_is_mul_by_nx1(dtype):
    """SpMV's pinned meaning: &A * &X with X n x 1 (src/csr/ops/mul.rs:8-59), bit for bit."""
    rng = np.random.default_rng(7)
    n, m = 57, 43
    a = _rand_csr(rng, n, m, 0.2, dtype)
    a[0][:] = a[0]                                   # keep
    x = rng.standard_normal(m).astype(dtype)
    xs = (np.arange(m + 1, dtype=np.uint64), np.zeros(m, np.uint64), x)   # m x 1, every row present
    cptr, cind, cval = orc.csr_mul(n, m, 1, a, xs)
    y = orc.csr_spmv(n, *a, x)
    present = np.diff(a[0].astype(np.int64)) > 0
    assert np.diff(cptr.astype(np.int64)).tolist() == present.astype(int).tolist()  # empty rows absent
    assert cval.tobytes() == y[present].tobytes()
    assert np.all(y[~present] == 0)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_assembly_equals_stable_sort_then_inorder_sum(dtype):
    """The property the device path is built on (hard part 1): the reference's double
    counting sort == stable sort by (row, col) + left-to-right segment sum + drop of == 0."""
    rng = np.random.default_rng(3)
    for case in range(30):
        nr, nc = int(rng.integers(1, 9)), int(rng.integers(1, 9))
        n = int(rng.integers(0, 120))
        r = rng.integers(0, nr, n).astype(np.uint64)
        c = rng.integers(0, nc, n).astype(np.uint64)
        v = rng.choice(np.array([1e-3, -1e-3, 1.0, -1.0, 1e8, -1e8, 0.0, 3.14159], dtype), n)
        got = orc.compress_from_coo(nr, nc, orc.make_triplets(r, c, v), "row")
        order = np.lexsort((np.arange(n), c, r))
        ptr = np.zeros(nr + 1, np.uint64)
        ind, val = [], []
        i = 0
        while i < n:
            j = i
            acc = v[order[i]]
            while j + 1 < n and r[order[j + 1]] == r[order[i]] and c[order[j + 1]] == c[order[i]]:
                j += 1
                acc = dtype(acc + v[order[j]])
            if acc != 0:
                ind.append(int(c[order[i]])); val.append(acc); ptr[int(r[order[i]]) + 1] += 1
            i = j + 1
        ptr = np.cumsum(ptr).astype(np.uint64)
        assert got[0].tolist() == ptr.tolist() and got[1].tolist() == ind
        assert got[2].tobytes() == np.array(val, dtype).tobytes()


def test_special_values():
    # -0.0 sums are dropped, NaN kept, subnormal kept (hard part 1)
    t = trip([[0, 0, -0.0], [0, 1, float("nan")], [1, 0, 1e-320], [1, 1, 1.0], [1, 1, -1.0]])
    ptr, ind, val = orc.compress_from_coo(2, 2, t, "row")
    assert ptr.tolist() == [0, 1, 2] and ind.tolist() == [1, 0]
    assert np.isnan(val[0]) and val[1] == 1e-320
    # empty COO => ptr all zero
    ptr, ind, val = orc.compress_from_coo(3, 2, trip([]), "row")
    assert ptr.tolist() == [0, 0, 0, 0] and len(ind) == 0 and len(val) == 0


def test_expand_round_trip():
    rng = np.random.default_rng(5)
    a = _rand_csr(rng, 9, 6, 0.3, np.float64)
    t = orc.expand_to_coo(9, *a, major="row")
    back = orc.compress_from_coo(9, 6, t, "row", dedup=False, dropzero=False)
    assert back[0].tolist() == a[0].tolist() and back[1].tolist() == a[1].tolist()
    assert back[2].tobytes() == a[2].tobytes()


# ------------------------------------------------------------------ second opinion: scipy.sparse
# The goldens pin the oracle on the reference's own small cases; here an independent implementation
# agrees with it on random inputs.  Values are small integers stored as floats, so every summation
# order gives the same bits and the comparison can be exact in structure AND values.
def _int_coo(rng, n, m, length, dtype):
    r = rng.integers(0, n, length).astype(np.uint64)
    c = rng.integers(0, m, length).astype(np.uint64)
    v = rng.integers(-3, 4, length).astype(dtype)            # zeros and cancellations happen
    return r, c, v


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("major", ["row", "col"])
def test_oracle_agrees_with_scipy_on_assembly_and_conversions(dtype, major):
    sps = pytest.importorskip("scipy.sparse")
    rng = np.random.default_rng(101)
    n, m = 300, 211
    r, c, v = _int_coo(rng, n, m, 6000, dtype)
    got = orc.compress_from_coo(n, m, orc.make_triplets(r, c, v), major)
    S = sps.coo_matrix((v, (r.astype(np.int64), c.astype(np.int64))), shape=(n, m))
    S = S.tocsr() if major == "row" else S.tocsc()
    S.sum_duplicates(); S.eliminate_zeros(); S.sort_indices()
    assert np.array_equal(got[0], S.indptr) and np.array_equal(got[1], S.indices)
    assert np.array_equal(got[2], S.data.astype(dtype))
    # the other format of the same matrix, and the transpose
    nmaj, nmin = (n, m) if major == "row" else (m, n)
    other = orc.recompress(nmaj, nmin, *got)
    T = S.tocsc() if major == "row" else S.tocsr()
    T.sort_indices()
    assert np.array_equal(other[0], T.indptr) and np.array_equal(other[1], T.indices)
    assert np.array_equal(other[2], T.data.astype(dtype))


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_oracle_agrees_with_scipy_on_add_sub_mul_spmv(dtype):
    sps = pytest.importorskip("scipy.sparse")
    rng = np.random.default_rng(202)
    n = 180
    a = orc.compress_from_coo(n, n, orc.make_triplets(*_int_coo(rng, n, n, 2500, dtype)), "row")
    b = orc.compress_from_coo(n, n, orc.make_triplets(*_int_coo(rng, n, n, 2500, dtype)), "row")
    A = sps.csr_matrix((a[2], a[1].astype(np.int64), a[0].astype(np.int64)), shape=(n, n))
    B = sps.csr_matrix((b[2], b[1].astype(np.int64), b[0].astype(np.int64)), shape=(n, n))
    for sub in (0, 1):
        got = orc.addsub(sub, n, n, a, b)
        # the reference keeps explicit zeros (pattern union): compare against the union pattern
        U = (abs(A) + abs(B)).tocsr(); U.sort_indices()
        want = (A - B if sub else A + B).tocsr()
        assert np.array_equal(got[0], U.indptr) and np.array_equal(got[1], U.indices)
        dense = np.asarray(want.todense())
        rows = np.repeat(np.arange(n), np.diff(U.indptr))
        assert np.array_equal(got[2], dense[rows, U.indices].astype(dtype))
    got = orc.csr_mul(n, n, n, a, b)
    P = (abs(A) @ abs(B)).tocsr(); P.sort_indices()             # structural product: no cancellation
    assert np.array_equal(got[0], P.indptr) and np.array_equal(got[1], P.indices)
    dense = np.asarray((A @ B).todense())
    rows = np.repeat(np.arange(n), np.diff(P.indptr))
    assert np.array_equal(got[2], dense[rows, P.indices].astype(dtype))
    x = rng.integers(-2, 3, n).astype(dtype)
    assert np.array_equal(orc.csr_spmv(n, *a, x), (A @ x).astype(dtype))
