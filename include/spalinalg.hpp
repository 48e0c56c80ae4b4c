// spalinalg.hpp — C++ host-side mirror of the reference crate's public API for the hot path, on
// top of the C ABI in spl.h (libspalinalg_b200.so).
//
// The reference (Rust crate spalinalg v0.0.2) is compiled code and its toolchain is not part of
// this image, so the typed host side above the ABI is written in C++: same type names, method
// names, argument meaning and error behaviour as the crate (paths under /root/reference):
//   CooMatrix<T>  src/coo.rs:52-57      builder, insertion-ordered triplets (host, SoA)
//   DokMatrix<T>  src/dok.rs:53-58      builder, hash map (host)
//   PinnedCooMatrix<T>                  CooMatrix on pinned storage streamed to the device (spl_coo)
//   CsrMatrix<T>  src/csr.rs:65-72      device resident; rowptr()/colind()/values() are host
//   CscMatrix<T>  src/csc.rs:65-72      copies made on first use, exactly sized, usize indices
//   T in {float, double}                src/scalar.rs:55-57
// Rust `From` impls are the `from(...)` static functions, `impl Add/Sub/Mul/Neg for &M` are the
// C++ operators, `assert!` / `assert_eq!` panics are `spalinalg::Panic` exceptions.  There is no
// CPU fallback: every CSR/CSC operation is a call into the library.
#ifndef SPALINALG_HPP
#define SPALINALG_HPP

#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <tuple>
#include <unordered_map>
#include <utility>
#include <vector>

#include "spl.h"

namespace spalinalg {

// The reference panics (assert!, assert_eq!, index out of bounds); the mirror throws this.
struct Panic : std::logic_error {
    using std::logic_error::logic_error;
};
// CUDA / resource failure reported by the library (no reference equivalent).
struct DeviceError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

template <typename T> struct Scalar;                       // src/scalar.rs:55-57
template <> struct Scalar<float> { static constexpr int dtype = SPL_F32; };
template <> struct Scalar<double> { static constexpr int dtype = SPL_F64; };

// A fixed-length vector in page-locked host memory (spl_host_alloc), zero-initialised.  With PinnedVector
// arguments matvec_into pipelines upload, product and download over row chunks (spl_spmv_host); std::vector
// arguments give the same result through one staged copy each way.  An extension: the reference has no dense
// vectors.
template <typename T>
class PinnedVector {
public:
    explicit PinnedVector(std::size_t n) : n_(n) {
        void *p = nullptr;
        if (spl_host_alloc(static_cast<std::uint64_t>(n * sizeof(T)), &p) != SPL_OK || !p)
            throw DeviceError("spl_host_alloc failed: no CUDA device, or out of pinnable memory");
        data_ = static_cast<T *>(p);
        for (std::size_t i = 0; i < n_; ++i) data_[i] = T(0);
    }
    explicit PinnedVector(const std::vector<T> &x) : PinnedVector(x.size()) {
        for (std::size_t i = 0; i < n_; ++i) data_[i] = x[i];
    }
    ~PinnedVector() { if (data_) spl_host_free(data_); }
    PinnedVector(const PinnedVector &) = delete;
    PinnedVector &operator=(const PinnedVector &) = delete;
    PinnedVector(PinnedVector &&o) noexcept : data_(o.data_), n_(o.n_) { o.data_ = nullptr; o.n_ = 0; }
    std::size_t size() const { return n_; }
    T *data() { return data_; }
    const T *data() const { return data_; }
    T &operator[](std::size_t i) { return data_[i]; }
    const T &operator[](std::size_t i) const { return data_[i]; }
    const T *begin() const { return data_; }
    const T *end() const { return data_ + n_; }
private:
    T *data_ = nullptr;
    std::size_t n_ = 0;
};

// One spl_ctx per host thread (spl.h); created on first use on device 0.
class Context {
public:
    explicit Context(int device = 0, void *stream = nullptr) {
        if (spl_ctx_create(device, stream, &raw_) != SPL_OK)
            throw DeviceError("spl_ctx_create failed: a CUDA device is required, there is no CPU fallback");
    }
    ~Context() { if (raw_) spl_ctx_destroy(raw_); }
    Context(const Context &) = delete;
    Context &operator=(const Context &) = delete;
    spl_ctx *raw() const { return raw_; }
    void check(int status) const {
        if (status == SPL_OK) return;
        const std::string msg = spl_last_error(raw_);
        if (status == SPL_ERR_SHAPE || status == SPL_ERR_INVALID || status == SPL_ERR_ARG) throw Panic(msg);
        throw DeviceError(msg);
    }
    int invalid_reason() const { return spl_invalid_reason(raw_); }
    static Context &current() {
        thread_local Context ctx(0);
        return ctx;
    }
private:
    spl_ctx *raw_ = nullptr;
};

template <typename T> class CsrMatrix;
template <typename T> class CscMatrix;
template <typename T> class DokMatrix;
template <typename T> class PinnedCooMatrix;

// ---------------------------------------------------------------------------- CooMatrix
template <typename T>
class CooMatrix {
public:
    CooMatrix(std::size_t nrows, std::size_t ncols) : nrows_(nrows), ncols_(ncols) {      // coo.rs:104-112
        if (!(nrows > 0)) throw Panic("assertion failed: nrows > 0");
        if (!(ncols > 0)) throw Panic("assertion failed: ncols > 0");
    }
    static CooMatrix with_capacity(std::size_t nrows, std::size_t ncols, std::size_t capacity) {
        CooMatrix m(nrows, ncols);
        m.rows_.reserve(capacity); m.cols_.reserve(capacity); m.vals_.reserve(capacity);
        return m;
    }
    static CooMatrix with_entries(std::size_t nrows, std::size_t ncols,
                                  const std::vector<std::tuple<std::size_t, std::size_t, T>> &entries) {
        CooMatrix m(nrows, ncols);                                                        // coo.rs:204-220
        for (const auto &e : entries) m.push(std::get<0>(e), std::get<1>(e), std::get<2>(e));
        return m;
    }
    static CooMatrix with_triplets(std::size_t nrows, std::size_t ncols, std::vector<std::size_t> rowind,
                                   std::vector<std::size_t> colind, std::vector<T> values) {
        if (rowind.size() != colind.size() || colind.size() != values.size())             // coo.rs:254-288
            throw Panic("assertion `left == right` failed: triplet lengths differ");
        CooMatrix m(nrows, ncols);
        for (std::size_t i = 0; i < values.size(); ++i) m.push(rowind[i], colind[i], values[i]);
        return m;
    }
    static CooMatrix eye(std::size_t size) {                                              // coo.rs:127-139
        if (!(size > 0)) throw Panic("assertion failed: size > 0");
        CooMatrix m(size, size);
        for (std::size_t i = 0; i < size; ++i) m.push(i, i, T(1));
        return m;
    }
    std::size_t nrows() const { return nrows_; }
    std::size_t ncols() const { return ncols_; }
    std::size_t length() const { return vals_.size(); }
    void push(std::size_t row, std::size_t col, T value) {                                // coo.rs:431-435
        if (!(row < nrows_)) throw Panic("assertion failed: row < self.nrows");
        if (!(col < ncols_)) throw Panic("assertion failed: col < self.ncols");
        rows_.push_back(row); cols_.push_back(col); vals_.push_back(value);
    }
    std::optional<std::tuple<std::size_t, std::size_t, T>> pop() {                        // coo.rs:450-452
        if (vals_.empty()) return std::nullopt;
        auto e = std::make_tuple(rows_.back(), cols_.back(), vals_.back());
        rows_.pop_back(); cols_.pop_back(); vals_.pop_back();
        return e;
    }
    std::optional<std::tuple<std::size_t, std::size_t, T>> get(std::size_t index) const { // coo.rs:386-390
        if (index >= vals_.size()) return std::nullopt;
        return std::make_tuple(rows_[index], cols_[index], vals_[index]);
    }
    // get_mut (coo.rs:408-412): the indices stay const, the value is writable
    std::optional<std::tuple<std::size_t, std::size_t, T *>> get_mut(std::size_t index) {
        if (index >= vals_.size()) return std::nullopt;
        return std::make_tuple(rows_[index], cols_[index], &vals_[index]);
    }
    std::size_t capacity() const { return vals_.capacity(); }                             // coo.rs:366-368
    // iter / iter_mut / into_iter (coo.rs:491-518, 576-627): entries in insertion order
    template <typename F> void for_each(F &&f) const {
        for (std::size_t i = 0; i < vals_.size(); ++i) f(rows_[i], cols_[i], vals_[i]);
    }
    template <typename F> void for_each_mut(F &&f) {
        for (std::size_t i = 0; i < vals_.size(); ++i) f(rows_[i], cols_[i], vals_[i]);
    }
    std::vector<std::tuple<std::size_t, std::size_t, T>> into_entries() const {
        std::vector<std::tuple<std::size_t, std::size_t, T>> out;
        out.reserve(vals_.size());
        for_each([&](std::size_t r, std::size_t c, const T &v) { out.emplace_back(r, c, v); });
        return out;
    }
    // Extend (coo.rs:548-574): every entry asserted before any is stored
    void extend(const std::vector<std::tuple<std::size_t, std::size_t, T>> &entries) {
        for (const auto &e : entries) {
            if (!(std::get<0>(e) < nrows_)) throw Panic("assertion failed: *row < self.nrows");
            if (!(std::get<1>(e) < ncols_)) throw Panic("assertion failed: *col < self.ncols");
        }
        for (const auto &e : entries) push(std::get<0>(e), std::get<1>(e), std::get<2>(e));
    }
    // Add / Sub / Neg for &CooMatrix (coo.rs:751-804): concatenation, duplicates summed at conversion
    friend CooMatrix operator+(const CooMatrix &a, const CooMatrix &b) {
        if (a.nrows_ != b.nrows_ || a.ncols_ != b.ncols_) throw Panic("assertion `left == right` failed (shape)");
        CooMatrix m = a;
        m.rows_.insert(m.rows_.end(), b.rows_.begin(), b.rows_.end());
        m.cols_.insert(m.cols_.end(), b.cols_.begin(), b.cols_.end());
        m.vals_.insert(m.vals_.end(), b.vals_.begin(), b.vals_.end());
        return m;
    }
    friend CooMatrix operator-(const CooMatrix &a, const CooMatrix &b) { return a + (-b); }
    friend CooMatrix operator-(const CooMatrix &a) {
        CooMatrix m = a;
        for (T &v : m.vals_) v = -v;
        return m;
    }
    void clear() { rows_.clear(); cols_.clear(); vals_.clear(); }
    CooMatrix transpose() const {                                                         // coo.rs:538-545
        CooMatrix m(ncols_, nrows_);
        m.rows_ = cols_; m.cols_ = rows_; m.vals_ = vals_;
        return m;
    }
    const std::vector<std::size_t> &rowind() const { return rows_; }
    const std::vector<std::size_t> &colind() const { return cols_; }
    const std::vector<T> &values() const { return vals_; }
    static CooMatrix from(const CsrMatrix<T> &m) { return m.to_coo(); }                   // coo.rs:629-705
    static CooMatrix from(const CscMatrix<T> &m) { return m.to_coo(); }
    static CooMatrix from(const DokMatrix<T> &dok);
private:
    std::size_t nrows_, ncols_;
    std::vector<std::size_t> rows_, cols_;      // SoA: what spl_mat_from_coo takes as is
    std::vector<T> vals_;
};

// ---------------------------------------------------------------------------- PinnedCooMatrix
// CooMatrix whose storage is an spl_coo (spl.h): pinned host SoA arrays streamed to
// the device chunk by chunk while they are filled, so CsrMatrix::from / CscMatrix::from find the
// triplets already in HBM.  Same surface as CooMatrix (src/coo.rs); needs a CUDA device.
template <typename T>
class PinnedCooMatrix {
public:
    PinnedCooMatrix(std::size_t nrows, std::size_t ncols, std::size_t capacity = 0) : nrows_(nrows), ncols_(ncols) {
        if (!(nrows > 0)) throw Panic("assertion failed: nrows > 0");                     // coo.rs:105-106
        if (!(ncols > 0)) throw Panic("assertion failed: ncols > 0");
        Context &c = Context::current();
        c.check(spl_coo_create(c.raw(), Scalar<T>::dtype, nrows, ncols, capacity, &b_));
    }
    static PinnedCooMatrix with_capacity(std::size_t nrows, std::size_t ncols, std::size_t capacity) {
        return PinnedCooMatrix(nrows, ncols, capacity);                                   // coo.rs:162-170
    }
    ~PinnedCooMatrix() { if (b_) spl_coo_free(b_); }
    PinnedCooMatrix(PinnedCooMatrix &&o) noexcept : nrows_(o.nrows_), ncols_(o.ncols_), b_(o.b_) { o.b_ = nullptr; }
    PinnedCooMatrix(const PinnedCooMatrix &) = delete;
    PinnedCooMatrix &operator=(const PinnedCooMatrix &) = delete;
    std::size_t nrows() const { return nrows_; }
    std::size_t ncols() const { return ncols_; }
    std::size_t length() const { return spl_coo_len(b_); }                                // coo.rs:349-351
    std::size_t capacity() const { return spl_coo_capacity(b_); }
    void push(std::size_t row, std::size_t col, T value) { check(spl_coo_push(b_, row, col, &value)); }   // coo.rs:431-435
    void extend(const std::vector<std::size_t> &rowind, const std::vector<std::size_t> &colind,
                const std::vector<T> &values) {                                           // coo.rs:566-573
        if (rowind.size() != values.size() || colind.size() != values.size())
            throw Panic("assertion `left == right` failed: triplet lengths differ");
        check(spl_coo_extend(b_, values.size(), reinterpret_cast<const std::uint64_t *>(rowind.data()),
                             reinterpret_cast<const std::uint64_t *>(colind.data()), values.data()));
    }
    std::optional<std::tuple<std::size_t, std::size_t, T>> get(std::size_t index) const { // coo.rs:386-390
        if (index >= length()) return std::nullopt;
        const std::uint64_t *r = nullptr, *c = nullptr;
        const void *v = nullptr;
        spl_coo_host_ptrs(b_, &r, &c, &v);
        return std::make_tuple((std::size_t)r[index], (std::size_t)c[index], static_cast<const T *>(v)[index]);
    }
    std::optional<std::tuple<std::size_t, std::size_t, T>> pop() {                        // coo.rs:450-452
        const std::size_t n = length();
        if (n == 0) return std::nullopt;
        auto e = get(n - 1);
        check(spl_coo_truncate(b_, n - 1));
        return e;
    }
    void clear() { check(spl_coo_truncate(b_, 0)); }                                      // coo.rs:470-472
    spl_coo *raw() const { return b_; }
private:
    void check(int status) const {
        if (status == SPL_OK) return;
        const std::string msg = spl_coo_last_error(b_);
        if (status == SPL_ERR_SHAPE || status == SPL_ERR_INVALID || status == SPL_ERR_ARG) throw Panic(msg);
        throw DeviceError(msg);
    }
    std::size_t nrows_, ncols_;
    spl_coo *b_ = nullptr;
};

// ---------------------------------------------------------------------------- DokMatrix
template <typename T>
class DokMatrix {
    struct KeyHash {
        std::size_t operator()(const std::pair<std::size_t, std::size_t> &k) const {
            return std::hash<std::size_t>()(k.first * 0x9e3779b97f4a7c15ull ^ k.second);
        }
    };
public:
    using Map = std::unordered_map<std::pair<std::size_t, std::size_t>, T, KeyHash>;
    DokMatrix(std::size_t nrows, std::size_t ncols) : nrows_(nrows), ncols_(ncols) {      // dok.rs:105-113
        if (!(nrows > 0)) throw Panic("assertion failed: nrows > 0");
        if (!(ncols > 0)) throw Panic("assertion failed: ncols > 0");
    }
    std::size_t nrows() const { return nrows_; }
    std::size_t ncols() const { return ncols_; }
    std::size_t length() const { return map_.size(); }
    std::optional<T> insert(std::size_t row, std::size_t col, T value) {                  // dok.rs:462-466
        if (!(row < nrows_)) throw Panic("assertion failed: row < self.nrows");
        if (!(col < ncols_)) throw Panic("assertion failed: col < self.ncols");
        auto it = map_.find({row, col});
        std::optional<T> old;
        if (it != map_.end()) { old = it->second; it->second = value; }
        else map_.emplace(std::make_pair(row, col), value);
        return old;
    }
    bool contains(std::size_t row, std::size_t col) const { return map_.count({row, col}) != 0; }
    std::optional<T> get(std::size_t row, std::size_t col) const {
        auto it = map_.find({row, col});
        return it == map_.end() ? std::nullopt : std::optional<T>(it->second);
    }
    const Map &entries() const { return map_; }
    static DokMatrix with_capacity(std::size_t nrows, std::size_t ncols, std::size_t capacity) {   // dok.rs:163-171
        DokMatrix m(nrows, ncols);
        m.map_.reserve(capacity);
        return m;
    }
    static DokMatrix eye(std::size_t size) {                                              // dok.rs:128-135
        if (!(size > 0)) throw Panic("assertion failed: size > 0");
        DokMatrix m(size, size);
        for (std::size_t i = 0; i < size; ++i) m.map_[{i, i}] = T(1);
        return m;
    }
    static DokMatrix with_entries(std::size_t nrows, std::size_t ncols,
                                  const std::vector<std::tuple<std::size_t, std::size_t, T>> &entries) {   // dok.rs:205-221
        DokMatrix m(nrows, ncols);
        m.extend(entries);
        return m;
    }
    void extend(const std::vector<std::tuple<std::size_t, std::size_t, T>> &entries) {    // dok.rs:561-587
        for (const auto &e : entries) {
            if (!(std::get<0>(e) < nrows_)) throw Panic("assertion failed: *row < self.nrows");
            if (!(std::get<1>(e) < ncols_)) throw Panic("assertion failed: *col < self.ncols");
        }
        for (const auto &e : entries) map_[{std::get<0>(e), std::get<1>(e)}] = std::get<2>(e);
    }
    T *get_mut(std::size_t row, std::size_t col) {                                        // dok.rs:439-441
        auto it = map_.find({row, col});
        return it == map_.end() ? nullptr : &it->second;
    }
    void clear() { map_.clear(); }                                                        // dok.rs:484-486
    DokMatrix transpose() const {                                                         // dok.rs:547-558
        DokMatrix m(ncols_, nrows_);
        for (const auto &kv : map_) m.map_[{kv.first.second, kv.first.first}] = kv.second;
        return m;
    }
    template <typename F> void for_each(F &&f) const { for (const auto &kv : map_) f(kv.first.first, kv.first.second, kv.second); }
    template <typename F> void for_each_mut(F &&f) { for (auto &kv : map_) f(kv.first.first, kv.first.second, kv.second); }
    // Add / Sub / Neg for &DokMatrix (dok.rs:722-769): lhs entries, then or_default() += / -= rhs
    friend DokMatrix operator+(const DokMatrix &a, const DokMatrix &b) {
        DokMatrix m = a;
        for (const auto &kv : b.map_) m.map_[kv.first] += kv.second;
        return m;
    }
    friend DokMatrix operator-(const DokMatrix &a, const DokMatrix &b) {
        DokMatrix m = a;
        for (const auto &kv : b.map_) m.map_[kv.first] -= kv.second;
        return m;
    }
    friend DokMatrix operator-(const DokMatrix &a) {
        DokMatrix m = a;
        for (auto &kv : m.map_) kv.second = -kv.second;
        return m;
    }
    static DokMatrix from(const CsrMatrix<T> &m);                                         // dok.rs:702-720
    static DokMatrix from(const CscMatrix<T> &m);                                         // dok.rs:676-693
    static DokMatrix from(const CooMatrix<T> &coo) {                                      // dok.rs:640-668
        DokMatrix m(coo.nrows(), coo.ncols());
        for (std::size_t i = 0; i < coo.length(); ++i)
            m.map_[{coo.rowind()[i], coo.colind()[i]}] += coo.values()[i];                 // or_default() += v
        return m;
    }
private:
    std::size_t nrows_, ncols_;
    Map map_;
};

template <typename T>
CooMatrix<T> CooMatrix<T>::from(const DokMatrix<T> &dok) {
    CooMatrix m(dok.nrows(), dok.ncols());
    for (const auto &kv : dok.entries()) m.push(kv.first.first, kv.first.second, kv.second);
    return m;
}

// ---------------------------------------------------------------------------- CSR / CSC
namespace detail {

struct MatDeleter {
    void operator()(spl_mat *m) const { if (m) spl_mat_free(Context::current().raw(), m); }
};

// Shared implementation of the two compressed formats (device handle + lazy host mirror).
template <typename T, int FORMAT>
class Compressed {
public:
    std::size_t nrows() const { return nrows_; }
    std::size_t ncols() const { return ncols_; }
    std::size_t nnz() const { return nnz_; }                                              // csr.rs:287-289
    const std::vector<T> &values() const { download(); return host_->val; }
    spl_mat *raw() const { return dev_.get(); }
    // values_mut() (src/csr.rs:270-272): hand f a writable copy of the values, then store it back
    template <typename F>
    void values_mut(F &&f) {
        std::vector<T> v = values();
        f(v);
        if (v.size() != nnz_) throw Panic("values_mut: the number of values must not change");
        ctx().check(spl_mat_set_values(ctx().raw(), raw(), v.data()));
        if (host_) host_->val = v;
    }
    // iter() / into_iter() (csr.rs:303-316, 409-440): (row, col, value) in storage order, read from the
    // device in chunks (spl_mat_read_entries): the host never holds more than one chunk
    template <typename F>
    void for_each(F &&f, std::size_t chunk = std::size_t(1) << 16) const {
        std::vector<std::uint64_t> r(std::min(chunk, nnz_)), c(r.size());
        std::vector<T> v(r.size());
        for (std::size_t start = 0; start < nnz_; start += chunk) {
            const std::size_t cnt = std::min(chunk, nnz_ - start);
            ctx().check(spl_mat_read_entries(ctx().raw(), raw(), start, cnt, r.data(), c.data(), v.data()));
            for (std::size_t i = 0; i < cnt; ++i) f(std::size_t(r[i]), std::size_t(c[i]), v[i]);
        }
    }
    // iter_mut (csr.rs:330-343): f may change the values; they are stored back afterwards
    template <typename F>
    void for_each_mut(F &&f) {
        download();
        const std::size_t nmajor = FORMAT == SPL_CSR ? nrows_ : ncols_;
        std::vector<T> v = host_->val;
        for (std::size_t m = 0; m < nmajor; ++m)
            for (std::size_t p = host_->ptr[m]; p < host_->ptr[m + 1]; ++p)
                FORMAT == SPL_CSR ? f(m, host_->ind[p], v[p]) : f(host_->ind[p], m, v[p]);
        ctx().check(spl_mat_set_values(ctx().raw(), raw(), v.data()));
        host_->val = v;
    }
    // y = A x with host vectors: the dense form of `&A * &X`, X n x 1 (src/csr/ops/mul.rs:5-60,
    // src/csc/ops/mul.rs:5-61); a CSC matrix multiplies through its cached CSR form on the device
    std::vector<T> matvec(const std::vector<T> &x) const {
        if (x.size() != ncols_) throw Panic("assertion `left == right` failed: self.ncols() == rhs.nrows()");
        std::vector<T> y(nrows_);
        ctx().check(spl_spmv_host(ctx().raw(), raw(), x.data(), y.data()));
        return y;
    }
    // the same into caller-owned page-locked vectors: pipelined over row chunks
    void matvec_into(const PinnedVector<T> &x, PinnedVector<T> &y) const {
        if (x.size() != ncols_) throw Panic("assertion `left == right` failed: self.ncols() == rhs.nrows()");
        if (y.size() != nrows_) throw Panic("assertion `left == right` failed: self.nrows() == y.len()");
        ctx().check(spl_spmv_host(ctx().raw(), raw(), x.data(), y.data()));
    }
    CooMatrix<T> to_coo() const {                                                         // coo.rs:629-705
        std::vector<std::uint64_t> r(nnz_), c(nnz_);
        std::vector<T> v(nnz_);
        ctx().check(spl_mat_to_coo(ctx().raw(), raw(), r.data(), c.data(), v.data()));
        return CooMatrix<T>::with_triplets(nrows_, ncols_, std::vector<std::size_t>(r.begin(), r.end()),
                                           std::vector<std::size_t>(c.begin(), c.end()), std::move(v));
    }
protected:
    struct Host {
        std::vector<std::size_t> ptr, ind;
        std::vector<T> val;
    };
    static Context &ctx() { return Context::current(); }
    void adopt(spl_mat *m) {
        dev_ = std::shared_ptr<spl_mat>(m, MatDeleter{});
        std::uint64_t nr = 0, nc = 0, nz = 0;
        spl_mat_info(m, nullptr, nullptr, &nr, &nc, &nz);
        nrows_ = nr; ncols_ = nc; nnz_ = nz;
        host_.reset();
    }
    void construct(std::size_t nrows, std::size_t ncols, const std::vector<std::size_t> &ptr,
                   const std::vector<std::size_t> &ind, const std::vector<T> &val) {
        static_assert(sizeof(std::size_t) == sizeof(std::uint64_t), "usize is 64 bit");
        spl_mat *m = nullptr;
        ctx().check(spl_mat_from_compressed(ctx().raw(), FORMAT, Scalar<T>::dtype, nrows, ncols, ptr.size(),
                                            reinterpret_cast<const std::uint64_t *>(ptr.data()), ind.size(),
                                            reinterpret_cast<const std::uint64_t *>(ind.data()), val.size(),
                                            val.data(), &m));
        adopt(m);
    }
    void from_triplets(const CooMatrix<T> &coo, int dedup, int dropzero) {
        spl_mat *m = nullptr;
        ctx().check(spl_mat_from_coo(ctx().raw(), FORMAT, Scalar<T>::dtype, coo.nrows(), coo.ncols(), coo.length(),
                                     reinterpret_cast<const std::uint64_t *>(coo.rowind().data()),
                                     reinterpret_cast<const std::uint64_t *>(coo.colind().data()),
                                     coo.values().data(), dedup, dropzero, &m));
        adopt(m);
    }
    void download() const {
        if (host_) return;
        auto h = std::make_shared<Host>();
        const std::size_t nmajor = FORMAT == SPL_CSR ? nrows_ : ncols_;
        h->ptr.resize(nmajor + 1); h->ind.resize(nnz_); h->val.resize(nnz_);   // exactly sized
        ctx().check(spl_mat_download(ctx().raw(), raw(), reinterpret_cast<std::uint64_t *>(h->ptr.data()),
                                     reinterpret_cast<std::uint64_t *>(h->ind.data()), h->val.data()));
        host_ = h;
    }
    std::shared_ptr<spl_mat> dev_;
    std::size_t nrows_ = 0, ncols_ = 0, nnz_ = 0;
    mutable std::shared_ptr<Host> host_;
};

}  // namespace detail

template <typename T>
class CsrMatrix : public detail::Compressed<T, SPL_CSR> {
    using Base = detail::Compressed<T, SPL_CSR>;
public:
    // CsrMatrix::new (src/csr.rs:137-164): validating constructor, panics on the failing assertion.
    CsrMatrix(std::size_t nrows, std::size_t ncols, const std::vector<std::size_t> &rowptr,
              const std::vector<std::size_t> &colind, const std::vector<T> &values) {
        this->construct(nrows, ncols, rowptr, colind, values);
    }
    static CsrMatrix eye(std::size_t size) {                                              // csr.rs:179-188
        spl_mat *m = nullptr;
        Base::ctx().check(spl_mat_eye(Base::ctx().raw(), SPL_CSR, Scalar<T>::dtype, size, &m));
        return CsrMatrix(m);
    }
    static CsrMatrix from(const CooMatrix<T> &coo) { CsrMatrix r; r.from_triplets(coo, 1, 1); return r; }   // csr/conv/coo.rs:3-116
    static CsrMatrix from(const PinnedCooMatrix<T> &coo) {                                // same, triplets already streamed
        spl_mat *m = nullptr;
        Base::ctx().check(spl_mat_from_coo_builder(Base::ctx().raw(), coo.raw(), SPL_CSR, 1, 1, &m));
        return CsrMatrix(m);
    }
    static CsrMatrix from(const DokMatrix<T> &dok) {                                      // csr/conv/dok.rs:3-76
        CsrMatrix r; r.from_triplets(CooMatrix<T>::from(dok), 0, 0); return r;
    }
    static CsrMatrix from(const CscMatrix<T> &csc);                                       // csr/conv/csc.rs:3-53
    const std::vector<std::size_t> &rowptr() const { this->download(); return this->host_->ptr; }
    const std::vector<std::size_t> &colind() const { this->download(); return this->host_->ind; }
    CsrMatrix transpose() const {                                                         // csr.rs:358-406
        spl_mat *m = nullptr;
        Base::ctx().check(spl_mat_transpose(Base::ctx().raw(), this->raw(), &m));
        return CsrMatrix(m);
    }
    friend CsrMatrix operator+(const CsrMatrix &a, const CsrMatrix &b) { return binary(spl_mat_add, a, b); }   // csr/ops/add.rs:5-75
    friend CsrMatrix operator-(const CsrMatrix &a, const CsrMatrix &b) { return binary(spl_mat_sub, a, b); }   // csr/ops/sub.rs:5-75
    friend CsrMatrix operator*(const CsrMatrix &a, const CsrMatrix &b) { return binary(spl_mat_mul, a, b); }   // csr/ops/mul.rs:5-60
    friend CsrMatrix operator-(const CsrMatrix &a) {                                      // csr/ops/neg.rs:5-18
        spl_mat *m = nullptr;
        Base::ctx().check(spl_mat_neg(Base::ctx().raw(), a.raw(), &m));
        return CsrMatrix(m);
    }
private:
    friend class CscMatrix<T>;
    CsrMatrix() = default;
    explicit CsrMatrix(spl_mat *m) { this->adopt(m); }
    static CsrMatrix binary(int (*fn)(spl_ctx *, const spl_mat *, const spl_mat *, spl_mat **),
                            const CsrMatrix &a, const CsrMatrix &b) {
        spl_mat *m = nullptr;
        Base::ctx().check(fn(Base::ctx().raw(), a.raw(), b.raw(), &m));
        return CsrMatrix(m);
    }
};

template <typename T>
class CscMatrix : public detail::Compressed<T, SPL_CSC> {
    using Base = detail::Compressed<T, SPL_CSC>;
public:
    // CscMatrix::new (src/csc.rs:137-164)
    CscMatrix(std::size_t nrows, std::size_t ncols, const std::vector<std::size_t> &colptr,
              const std::vector<std::size_t> &rowind, const std::vector<T> &values) {
        this->construct(nrows, ncols, colptr, rowind, values);
    }
    static CscMatrix eye(std::size_t size) {                                              // csc.rs:179-188
        spl_mat *m = nullptr;
        Base::ctx().check(spl_mat_eye(Base::ctx().raw(), SPL_CSC, Scalar<T>::dtype, size, &m));
        return CscMatrix(m);
    }
    static CscMatrix from(const CooMatrix<T> &coo) { CscMatrix r; r.from_triplets(coo, 1, 1); return r; }   // csc/conv/coo.rs:3-116
    static CscMatrix from(const PinnedCooMatrix<T> &coo) {
        spl_mat *m = nullptr;
        Base::ctx().check(spl_mat_from_coo_builder(Base::ctx().raw(), coo.raw(), SPL_CSC, 1, 1, &m));
        return CscMatrix(m);
    }
    static CscMatrix from(const DokMatrix<T> &dok) {                                      // csc/conv/dok.rs:3-76
        CscMatrix r; r.from_triplets(CooMatrix<T>::from(dok), 0, 0); return r;
    }
    static CscMatrix from(const CsrMatrix<T> &csr) {                                      // csc/conv/csr.rs:3-53
        spl_mat *m = nullptr;
        Base::ctx().check(spl_mat_convert(Base::ctx().raw(), csr.raw(), SPL_CSC, &m));
        return CscMatrix(m);
    }
    const std::vector<std::size_t> &colptr() const { this->download(); return this->host_->ptr; }
    const std::vector<std::size_t> &rowind() const { this->download(); return this->host_->ind; }
    CscMatrix transpose() const {                                                         // csc.rs:358-406
        spl_mat *m = nullptr;
        Base::ctx().check(spl_mat_transpose(Base::ctx().raw(), this->raw(), &m));
        return CscMatrix(m);
    }
    friend CscMatrix operator+(const CscMatrix &a, const CscMatrix &b) { return binary(spl_mat_add, a, b); }   // csc/ops/add.rs:5-70
    friend CscMatrix operator-(const CscMatrix &a, const CscMatrix &b) { return binary(spl_mat_sub, a, b); }   // csc/ops/sub.rs:5-70
    friend CscMatrix operator*(const CscMatrix &a, const CscMatrix &b) { return binary(spl_mat_mul, a, b); }   // csc/ops/mul.rs:5-61
    friend CscMatrix operator-(const CscMatrix &a) {                                      // csc/ops/neg.rs:5-18
        spl_mat *m = nullptr;
        Base::ctx().check(spl_mat_neg(Base::ctx().raw(), a.raw(), &m));
        return CscMatrix(m);
    }
private:
    friend class CsrMatrix<T>;
    CscMatrix() = default;
    explicit CscMatrix(spl_mat *m) { this->adopt(m); }
    static CscMatrix binary(int (*fn)(spl_ctx *, const spl_mat *, const spl_mat *, spl_mat **),
                            const CscMatrix &a, const CscMatrix &b) {
        spl_mat *m = nullptr;
        Base::ctx().check(fn(Base::ctx().raw(), a.raw(), b.raw(), &m));
        return CscMatrix(m);
    }
};

template <typename T>
CsrMatrix<T> CsrMatrix<T>::from(const CscMatrix<T> &csc) {
    spl_mat *m = nullptr;
    Base::ctx().check(spl_mat_convert(Base::ctx().raw(), csc.raw(), SPL_CSR, &m));
    return CsrMatrix(m);
}

template <typename T>
DokMatrix<T> DokMatrix<T>::from(const CsrMatrix<T> &m) {
    DokMatrix d(m.nrows(), m.ncols());
    m.for_each([&](std::size_t r, std::size_t c, const T &v) { d.map_[{r, c}] = v; });
    return d;
}
template <typename T>
DokMatrix<T> DokMatrix<T>::from(const CscMatrix<T> &m) {
    DokMatrix d(m.nrows(), m.ncols());
    m.for_each([&](std::size_t r, std::size_t c, const T &v) { d.map_[{r, c}] = v; });
    return d;
}

}  // namespace spalinalg

#endif  // SPALINALG_HPP
