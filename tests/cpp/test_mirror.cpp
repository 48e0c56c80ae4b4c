// The reference's own hot-path unit tests, restated against the C++ host mirror
// (include/spalinalg.hpp) — same inputs, same expected arrays, exact comparison.
// Each check names the reference test it restates (file:line under /root/reference).
// Exit code 0 = all passed.  Needs a CUDA device (no CPU fallback).
#include <cmath>
#include <cstdio>
#include <vector>

#include "spalinalg.hpp"

using namespace spalinalg;
using V = std::vector<std::size_t>;
using D = std::vector<double>;

static int failures = 0;
#define CHECK(cond)                                                           \
    do {                                                                      \
        if (!(cond)) { std::printf("FAILED %s:%d  %s\n", __FILE__, __LINE__, #cond); ++failures; } \
    } while (0)

template <typename F>
static bool panics(F &&f) {
    try { f(); } catch (const Panic &) { return true; }
    return false;
}

static int run();

int main() {
    try {
        return run();
    } catch (const DeviceError &e) {
        std::printf("no CUDA device / device failure: %s\n", e.what());
        return 2;
    }
}

static int run() {
    {   // src/csr/conv/coo.rs:129-145 and src/csc/conv/coo.rs:129-145
        CooMatrix<double> coo(2, 3);
        coo.push(1, 2, 5.0); coo.push(0, 2, 4.0); coo.push(0, 1, 3.0); coo.push(0, 0, 1.0);
        coo.push(0, 0, 2.0); coo.push(1, 0, 0.0); coo.push(1, 1, 1.0); coo.push(1, 1, -1.0);
        auto csr = CsrMatrix<double>::from(coo);
        CHECK(csr.rowptr() == (V{0, 3, 4}) && csr.colind() == (V{0, 1, 2, 2}) && csr.values() == (D{3, 3, 4, 5}));
        auto csc = CscMatrix<double>::from(coo);
        CHECK(csc.colptr() == (V{0, 1, 2, 4}) && csc.rowind() == (V{0, 0, 0, 1}) && csc.values() == (D{3, 3, 4, 5}));
        CHECK(csr.values().capacity() >= csr.values().size() && csr.nnz() == 4);
    }
    {   // the same test through the streamed pinned storage (spl_coo)
        PinnedCooMatrix<double> coo(2, 3);
        coo.push(1, 2, 5.0); coo.push(0, 2, 4.0); coo.push(0, 1, 3.0); coo.push(0, 0, 1.0);
        coo.push(0, 0, 2.0); coo.push(1, 0, 0.0); coo.push(1, 1, 1.0); coo.push(1, 1, -1.0);
        coo.push(0, 0, 9.0);
        CHECK(coo.length() == 9 && std::get<2>(*coo.pop()) == 9.0 && coo.length() == 8);
        CHECK(panics([&] { coo.push(2, 0, 1.0); }) && panics([&] { coo.push(0, 3, 1.0); }));
        auto csr = CsrMatrix<double>::from(coo);
        CHECK(csr.rowptr() == (V{0, 3, 4}) && csr.colind() == (V{0, 1, 2, 2}) && csr.values() == (D{3, 3, 4, 5}));
        auto csc = CscMatrix<double>::from(coo);
        CHECK(csc.colptr() == (V{0, 1, 2, 4}) && csc.rowind() == (V{0, 0, 0, 1}) && csc.values() == (D{3, 3, 4, 5}));
        CHECK(panics([] { PinnedCooMatrix<float> bad(0, 1); }));
    }
    {   // src/csr.rs:352-356 (transpose doctest), src/csc.rs:352-356
        CsrMatrix<double> m(2, 2, V{0, 2, 3}, V{0, 1, 1}, D{1, 2, 3});
        auto t = m.transpose();
        CHECK(t.rowptr() == (V{0, 1, 3}) && t.colind() == (V{0, 0, 1}) && t.values() == (D{1, 2, 3}));
        CscMatrix<double> c(2, 2, V{0, 2, 3}, V{0, 1, 1}, D{1, 2, 3});
        auto tc = c.transpose();
        CHECK(tc.colptr() == (V{0, 1, 3}) && tc.rowind() == (V{0, 0, 1}) && tc.values() == (D{1, 2, 3}));
        // CSR -> CSC -> CSR round trip (src/csc/conv/csr.rs:3-53, src/csr/conv/csc.rs:3-53)
        auto back = CsrMatrix<double>::from(CscMatrix<double>::from(m));
        CHECK(back.rowptr() == m.rowptr() && back.colind() == m.colind() && back.values() == m.values());
    }
    {   // src/csr/ops/add.rs:82-105 and src/csr/ops/sub.rs:82-108
        CsrMatrix<double> lhs(4, 4, V{0, 1, 3, 4, 7}, V{0, 0, 2, 1, 1, 2, 3}, D{1, 2, 3, 4, 5, 6, 7});
        CsrMatrix<double> rhs(4, 4, V{0, 2, 3, 4, 5}, V{0, 2, 2, 3, 1}, D{2, 4, 8, 10, 6});
        auto add = lhs + rhs;
        CHECK(add.rowptr() == (V{0, 2, 4, 6, 9}) && add.colind() == (V{0, 2, 0, 2, 1, 3, 1, 2, 3}) &&
              add.values() == (D{3, 4, 2, 11, 4, 10, 11, 6, 7}));
        auto sub = lhs - rhs;
        CHECK(sub.rowptr() == (V{0, 2, 4, 6, 9}) && sub.colind() == (V{0, 2, 0, 2, 1, 3, 1, 2, 3}) &&
              sub.values() == (D{-1, -4, 2, -5, 4, -10, -1, 6, 7}));
        CHECK(panics([&] { CsrMatrix<double> other(3, 4, V{0, 0, 0, 0}, V{}, D{}); (void)(lhs + other); }));   // add.rs:9-10
        auto neg = -lhs;                                                        // src/csr/ops/neg.rs:25-36
        CHECK(neg.rowptr() == lhs.rowptr() && neg.colind() == lhs.colind() && neg.values() == (D{-1, -2, -3, -4, -5, -6, -7}));
    }
    {   // src/csc/ops/mul.rs:68-95 (5x3 * 3x4); CSR Mul pinned through CSC(M) arrays == CSR(M^T) arrays
        CscMatrix<double> lhs(5, 3, V{0, 3, 4, 6}, V{0, 1, 4, 3, 1, 2}, D{1, -5, 4, 3, 7, 2});
        CscMatrix<double> rhs(3, 4, V{0, 3, 4, 5, 6}, V{0, 1, 2, 2, 0, 1}, D{1, -5, 7, 3, -2, 4});
        auto mul = lhs * rhs;
        CHECK(mul.nrows() == 5 && mul.ncols() == 4);
        CHECK(mul.colptr() == (V{0, 5, 7, 10, 11}) && mul.rowind() == (V{0, 1, 2, 3, 4, 1, 2, 0, 1, 4, 3}) &&
              mul.values() == (D{1, 44, 14, -15, 4, 21, 6, -2, 10, -8, 12}));
        auto mul_csr = CsrMatrix<double>::from(lhs) * CsrMatrix<double>::from(rhs);
        auto again = CscMatrix<double>::from(mul_csr);
        CHECK(again.colptr() == mul.colptr() && again.rowind() == mul.rowind() && again.values() == mul.values());
    }
    {   // src/csr.rs:470-510: the seven CsrMatrix::new panics
        CHECK(panics([] { CsrMatrix<double>(0, 1, V{0}, V{}, D{}); }));                        // nrows > 0
        CHECK(panics([] { CsrMatrix<double>(1, 0, V{0, 0}, V{}, D{}); }));                     // ncols > 0
        CHECK(panics([] { CsrMatrix<double>(1, 1, V{0}, V{}, D{}); }));                        // rowptr.len() == nrows + 1
        CHECK(panics([] { CsrMatrix<double>(1, 1, V{1, 1}, V{0}, D{1}); }));                   // rowptr[0] == 0
        CHECK(panics([] { CsrMatrix<double>(1, 1, V{0, 1}, V{}, D{}); }));                     // colind.len() == nz
        CHECK(panics([] { CsrMatrix<double>(2, 2, V{0, 2, 1}, V{0}, D{1}); }));                // rowptr sorted
        CHECK(panics([] { CsrMatrix<double>(1, 2, V{0, 2}, V{1, 0}, D{1, 2}); }));             // colind increasing
        CHECK(panics([] { CsrMatrix<double>(1, 2, V{0, 1}, V{2}, D{1}); }));                   // col < ncols
        CHECK(!panics([] { CsrMatrix<double>(1, 2, V{0, 2}, V{0, 1}, D{1, 2}); }));
    }
    {   // eye (src/csr.rs:179-188), DOK conversion keeps explicit zeros (src/csr/conv/dok.rs:89-99)
        auto e = CscMatrix<float>::eye(3);
        CHECK(e.colptr() == (V{0, 1, 2, 3}) && e.rowind() == (V{0, 1, 2}) && e.values() == (std::vector<float>{1, 1, 1}));
        DokMatrix<double> dok(2, 3);
        dok.insert(1, 2, 5.0); dok.insert(0, 0, 0.0); dok.insert(0, 1, 3.0);
        auto m = CsrMatrix<double>::from(dok);
        CHECK(m.rowptr() == (V{0, 2, 3}) && m.colind() == (V{0, 1, 2}) && m.values() == (D{0, 3, 5}));
        CHECK(CooMatrix<double>::from(m).length() == 3);
    }
    {   // `&A * &X` with X n x 1 is the reference's SpMV (src/csr/ops/mul.rs:5-60); matvec is its dense form
        CsrMatrix<double> a(3, 3, V{0, 2, 2, 4}, V{0, 2, 1, 2}, D{1, 2, 3, 4});
        CsrMatrix<double> x(3, 1, V{0, 1, 2, 3}, V{0, 0, 0}, D{10, 20, 30});
        auto ax = a * x;
        CHECK(ax.rowptr() == (V{0, 1, 1, 2}) && ax.colind() == (V{0, 0}) && ax.values() == (D{70, 180}));
        CHECK(a.matvec(D{10, 20, 30}) == (D{70, 0, 180}));
        a.values_mut([](D &v) { for (auto &e : v) e *= 2; });                            // src/csr.rs:270-272
        CHECK(a.values() == (D{2, 4, 6, 8}) && a.matvec(D{10, 20, 30}) == (D{140, 0, 360}));
        PinnedVector<double> px(D{10, 20, 30}), py(3);                                    // page-locked vectors (spl_host_alloc)
        a.matvec_into(px, py);
        CHECK(py[0] == 140 && py[1] == 0 && py[2] == 360);
        PinnedVector<double> short_y(2);
        CHECK(panics([&] { a.matvec_into(px, short_y); }));
        CHECK(panics([&] { (void)(x * a); }));                                             // mul.rs:9
    }
    {   // src/csc/ops/add.rs:77-100, src/csc/ops/sub.rs:77-103, src/csc/ops/neg.rs:25-36 (CSC twins), f32 values
        using F = std::vector<float>;
        CscMatrix<float> lhs(4, 4, V{0, 2, 4, 6, 7}, V{0, 1, 2, 3, 1, 3, 3}, F{1, 2, 4, 5, 3, 6, 7});
        CscMatrix<float> rhs(4, 4, V{0, 1, 2, 4, 5}, V{0, 3, 0, 1, 2}, F{2, 6, 4, 8, 10});
        auto add = lhs + rhs;
        CHECK(add.colptr() == (V{0, 2, 4, 7, 9}) && add.rowind() == (V{0, 1, 2, 3, 0, 1, 3, 2, 3}) &&
              add.values() == (F{3, 2, 4, 11, 4, 11, 6, 10, 7}));
        auto sub = lhs - rhs;
        CHECK(sub.colptr() == add.colptr() && sub.rowind() == add.rowind() &&
              sub.values() == (F{-1, 2, 4, -1, -4, -5, 6, -10, 7}));
        CscMatrix<float> one(1, 2, V{0, 1, 2}, V{0, 0}, F{1, 2});
        CHECK((-one).values() == (F{-1, -2}) && (-one).colptr() == one.colptr());
        CHECK(panics([] { CscMatrix<float>(2, 1, V{0, 2}, V{1, 0}, F{1, 2}); }));               // rowind increasing
        // CooMatrix shell: push bounds, pop, get (src/coo.rs tests)
        CooMatrix<float> coo(2, 2);
        coo.push(0, 1, 1.0f);
        CHECK(panics([&] { coo.push(2, 0, 1.0f); }) && panics([&] { coo.push(0, 2, 1.0f); }));
        CHECK(coo.length() == 1 && coo.get(0).has_value() && !coo.get(1).has_value());
        CHECK(coo.transpose().rowind() == (V{1}) && coo.pop().has_value() && coo.length() == 0);
        CHECK(panics([] { CooMatrix<float>(0, 1); }) && panics([] { DokMatrix<float>(1, 0); }));
    }
    {   // CooMatrix builder surface: src/coo.rs tests get_mut :986-991, iter :1031-1039, iter_mut :1042-1050,
        // extend :1053-1063, into_iter :1066-1074, add :1077-1091, sub :1094-1108, neg :1111-1119
        using E = std::vector<std::tuple<std::size_t, std::size_t, double>>;
        const E entries{{0, 0, 1.0}, {1, 0, 2.0}, {0, 2, 3.0}};
        auto m = CooMatrix<double>::with_entries(2, 3, entries);
        CHECK(m.length() == 3 && m.capacity() >= 3);
        auto g = m.get_mut(0);
        CHECK(g.has_value() && std::get<0>(*g) == 0 && std::get<1>(*g) == 0 && *std::get<2>(*g) == 1.0 && !m.get_mut(3).has_value());
        *std::get<2>(*g) = 1.5;
        CHECK(std::get<2>(*m.get(0)) == 1.5);
        *std::get<2>(*m.get_mut(0)) = 1.0;
        E seen;
        m.for_each([&](std::size_t r, std::size_t c, const double &v) { seen.emplace_back(r, c, v); });
        CHECK(seen == entries && m.into_entries() == entries);
        m.for_each_mut([](std::size_t, std::size_t, double &v) { v *= 2; });
        CHECK(m.values() == (D{2, 4, 6}));
        CooMatrix<double> ext(2, 3);
        ext.extend(entries);
        CHECK(ext.length() == 3 && ext.into_entries() == entries);
        CHECK(panics([&] { ext.extend(E{{0, 0, 1.0}, {2, 0, 1.0}}); }) && ext.length() == 3);      // asserted before any is stored
        auto lhs = CooMatrix<double>::with_entries(2, 3, entries);
        auto rhs = CooMatrix<double>::with_entries(2, 3, E{{0, 0, 2.0}, {1, 1, 4.0}, {1, 2, 6.0}});
        CHECK((lhs + rhs).into_entries() == (E{{0, 0, 1.0}, {1, 0, 2.0}, {0, 2, 3.0}, {0, 0, 2.0}, {1, 1, 4.0}, {1, 2, 6.0}}));
        CHECK((lhs - rhs).into_entries() == (E{{0, 0, 1.0}, {1, 0, 2.0}, {0, 2, 3.0}, {0, 0, -2.0}, {1, 1, -4.0}, {1, 2, -6.0}}));
        CHECK((-lhs).into_entries() == (E{{0, 0, -1.0}, {1, 0, -2.0}, {0, 2, -3.0}}));
        CHECK(panics([&] { (void)(lhs + CooMatrix<double>(3, 3)); }));
        // the sum converts to the device formats with its duplicates added in insertion order (1.0 + 2.0)
        auto sum = CsrMatrix<double>::from(lhs + rhs);
        CHECK(sum.rowptr() == (V{0, 2, 5}) && sum.colind() == (V{0, 2, 0, 1, 2}) && sum.values() == (D{3, 3, 2, 4, 6}));
    }
    {   // DokMatrix builder surface (src/dok.rs:105-769) and the conversions back from the device formats
        using E = std::vector<std::tuple<std::size_t, std::size_t, double>>;
        auto eye = DokMatrix<double>::eye(2);
        CHECK(eye.length() == 2 && eye.contains(1, 1) && !eye.contains(0, 1) && *eye.get(0, 0) == 1.0);
        auto d = DokMatrix<double>::with_entries(2, 3, E{{0, 0, 1.0}, {1, 0, 2.0}, {0, 2, 3.0}});
        *d.get_mut(1, 0) = 2.5;
        CHECK(*d.get(1, 0) == 2.5 && d.get_mut(1, 1) == nullptr);
        auto t = d.transpose();
        CHECK(t.nrows() == 3 && t.ncols() == 2 && *t.get(2, 0) == 3.0 && *t.get(0, 1) == 2.5);
        auto s2 = d + d;
        CHECK(*s2.get(0, 2) == 6.0 && s2.length() == 3);
        auto z = d - d;
        CHECK(*z.get(0, 0) == 0.0 && z.length() == 3);                                     // explicit zeros stay
        CHECK(*(-d).get(0, 0) == -1.0);
        CHECK(panics([&] { d.extend(E{{5, 0, 1.0}}); }));
        auto csr = CsrMatrix<double>::from(d);                                             // DOK -> CSR on the device ...
        auto back = DokMatrix<double>::from(csr);                                          // ... and back, read in chunks
        CHECK(back.length() == 3 && *back.get(1, 0) == 2.5 && *back.get(0, 2) == 3.0);
        auto backc = DokMatrix<double>::from(CscMatrix<double>::from(csr));
        CHECK(backc.length() == 3 && *backc.get(0, 0) == 1.0);
        d.clear();
        CHECK(d.length() == 0);
    }
    {   // iter / iter_mut / into_iter on the device formats (src/csr.rs:303-343, 409-464; the CSC twins), Aᵀ x through
        // the CSC view, SpMV on a CscMatrix (src/csc/ops/mul.rs:5-61)
        using E = std::vector<std::tuple<std::size_t, std::size_t, double>>;
        CsrMatrix<double> a(3, 4, V{0, 2, 2, 5}, V{0, 3, 0, 1, 2}, D{1, 2, 3, 4, 5});
        E seen;
        a.for_each([&](std::size_t r, std::size_t c, const double &v) { seen.emplace_back(r, c, v); }, 2);   // chunks of 2
        CHECK(seen == (E{{0, 0, 1.0}, {0, 3, 2.0}, {2, 0, 3.0}, {2, 1, 4.0}, {2, 2, 5.0}}));
        auto c = CscMatrix<double>::from(a);
        seen.clear();
        c.for_each([&](std::size_t r, std::size_t cc, const double &v) { seen.emplace_back(r, cc, v); });
        CHECK(seen == (E{{0, 0, 1.0}, {2, 0, 3.0}, {2, 1, 4.0}, {2, 2, 5.0}, {0, 3, 2.0}}));
        CHECK(c.matvec(D{1, 10, 100, 1000}) == a.matvec(D{1, 10, 100, 1000}) && a.matvec(D{1, 10, 100, 1000}) == (D{2001, 0, 543}));
        a.for_each_mut([](std::size_t r, std::size_t, double &v) { if (r == 2) v = -v; });
        CHECK(a.values() == (D{1, 2, -3, -4, -5}) && a.matvec(D{1, 10, 100, 1000}) == (D{2001, 0, -543}));
        // y = A^T z: the CSC matrix with A's arrays and swapped dims
        CscMatrix<double> at(4, 3, V{0, 2, 2, 5}, V{0, 3, 0, 1, 2}, D{1, 2, 3, 4, 5});
        CHECK(at.matvec(D{1, 10, 100}) == (D{301, 400, 500, 2}));
    }
    if (failures == 0) std::printf("cpp mirror: all reference tests passed\n");
    return failures == 0 ? 0 : 1;
}
