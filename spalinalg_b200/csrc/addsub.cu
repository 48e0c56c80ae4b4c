// addsub.cu — C = A + B / A - B on compressed arrays of the same format.
//
// Reference: impl Add/Sub for &CsrMatrix (src/csr/ops/add.rs:5-75, sub.rs:5-75) and &CscMatrix
// (src/csc/ops/add.rs:5-70, sub.rs:5-70).  There: three transposes around a stamp-array merge.
// Result: per segment the sorted union of the two index sets, explicit zeros kept; value a+b
// (a-b) where both are stored, a where only lhs, b (-b, sub.rs:47) where only rhs.  One IEEE
// operation per output, so the device result is bit-exact.
// Device formulation: two passes over the segments, symbolic (count the union) then numeric
// (merge); the count's exclusive scan is the output pointer array.  No transposes needed:
// inputs are already index-sorted inside each segment (CsrMatrix::new asserts it).
#include "kernels.cuh"
#include "scan.cuh"

namespace spl {

namespace {

__global__ void __launch_bounds__(256)
union_count_kernel(uint32_t nmajor, const uint32_t *__restrict__ aptr,
                   const uint32_t *__restrict__ aind, const uint32_t *__restrict__ bptr,
                   const uint32_t *__restrict__ bind, uint32_t *__restrict__ cnt) {
    const uint64_t m = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= nmajor) return;
    uint32_t pa = aptr[m], ea = aptr[m + 1], pb = bptr[m], eb = bptr[m + 1], c = 0;
    while (pa < ea && pb < eb) {
        const uint32_t ia = __ldg(aind + pa), ib = __ldg(bind + pb);
        pa += ia <= ib;
        pb += ib <= ia;
        ++c;
    }
    cnt[m] = c + (ea - pa) + (eb - pb);
}

template <typename T, bool SUB>
__global__ void __launch_bounds__(256)
union_fill_kernel(uint32_t nmajor, const uint32_t *__restrict__ aptr,
                  const uint32_t *__restrict__ aind, const T *__restrict__ aval,
                  const uint32_t *__restrict__ bptr, const uint32_t *__restrict__ bind,
                  const T *__restrict__ bval, const uint32_t *__restrict__ cptr,
                  uint32_t *__restrict__ cind, T *__restrict__ cval) {
    const uint64_t m = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= nmajor) return;
    uint32_t pa = aptr[m], ea = aptr[m + 1], pb = bptr[m], eb = bptr[m + 1], pc = cptr[m];
    while (pa < ea && pb < eb) {
        const uint32_t ia = __ldg(aind + pa), ib = __ldg(bind + pb);
        if (ia < ib) {
            cind[pc] = ia; cval[pc] = aval[pa]; ++pa;
        } else if (ib < ia) {
            cind[pc] = ib; cval[pc] = SUB ? flip_sign(bval[pb]) : bval[pb]; ++pb;
        } else {
            cind[pc] = ia; cval[pc] = SUB ? aval[pa] - bval[pb] : aval[pa] + bval[pb]; ++pa; ++pb;
        }
        ++pc;
    }
    for (; pa < ea; ++pa, ++pc) { cind[pc] = aind[pa]; cval[pc] = aval[pa]; }
    for (; pb < eb; ++pb, ++pc) { cind[pc] = bind[pb]; cval[pc] = SUB ? flip_sign(bval[pb]) : bval[pb]; }
}

template <typename T>
void fill(spl_ctx *ctx, const spl_mat *a, const spl_mat *b, spl_mat *c, int subtract) {
    const unsigned grid = div_up(a->nmajor(), 256);
    if (subtract)
        union_fill_kernel<T, true><<<grid, 256, 0, ctx->stream>>>(
            a->nmajor(), a->ptr, a->ind, (const T *)a->val, b->ptr, b->ind, (const T *)b->val, c->ptr,
            c->ind, (T *)c->val);
    else
        union_fill_kernel<T, false><<<grid, 256, 0, ctx->stream>>>(
            a->nmajor(), a->ptr, a->ind, (const T *)a->val, b->ptr, b->ind, (const T *)b->val, c->ptr,
            c->ind, (T *)c->val);
    check_launch(ctx, "union_fill");
}

}  // namespace

spl_mat *addsub(spl_ctx *ctx, const spl_mat *a, const spl_mat *b, int subtract) {
    SPL_REQUIRE(a->nrows == b->nrows && a->ncols == b->ncols, SPL_ERR_SHAPE,
                "add/sub: shapes differ (assert_eq!, src/csr/ops/add.rs:9-10)");
    SPL_REQUIRE(a->format == b->format && a->dtype == b->dtype, SPL_ERR_ARG,
                "add/sub: operands must share format and scalar type");
    SPL_REQUIRE((uint64_t)a->nnz + b->nnz < (1ull << 32), SPL_ERR_UNSUPPORTED,
                "add/sub: nnz(A)+nnz(B) must stay below 2^32");
    const uint32_t nmajor = a->nmajor();
    Tmp<uint32_t> cnt(ctx, nmajor);
    Tmp<uint32_t> cptr(ctx, (size_t)nmajor + 1);
    union_count_kernel<<<div_up(nmajor, 256), 256, 0, ctx->stream>>>(nmajor, a->ptr, a->ind, b->ptr,
                                                                    b->ind, cnt);
    check_launch(ctx, "union_count");
    exclusive_scan_u32(ctx, cnt, nmajor, cptr);
    uint32_t nnz = 0;
    read_back(ctx, cptr.p + nmajor, &nnz, 1);   // exact-size output (add.rs:103-105)
    spl_mat *c = new_mat(ctx, a->format, a->dtype, a->nrows, a->ncols, nnz);
    dfree(ctx, c->ptr);
    c->ptr = cptr.release();
    try {
        if (a->dtype == SPL_F32) fill<float>(ctx, a, b, c, subtract);
        else fill<double>(ctx, a, b, c, subtract);
    } catch (...) {
        free_mat(ctx, c);
        throw;
    }
    return c;
}

}  // namespace spl
