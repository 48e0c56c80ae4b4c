// spgemm.cu — C = A * B for compressed operands.
//
// Reference: impl Mul for &CsrMatrix (src/csr/ops/mul.rs:5-60) and &CscMatrix
// (src/csc/ops/mul.rs:5-61): Gustavson on the transposes with a dense accumulator; each
// C[i,j] = sum over ascending k of round(A[i,k]*B[k,j]), the first term copied, the pattern the
// structural union (explicit zeros kept).  Device formulation (expand - sort - compress):
// every A entry emits its products against B's row k in storage order, which is ascending k
// for a fixed (i,j); a stable sort by (i,j) keeps that order and the in-order segmented sum
// of assembly (dedup=1, dropzero=0) reproduces the accumulation bit for bit.
// CSC operands: CSC(A*B) arrays == CSR(B^T * A^T) arrays, same products and k order, so the
// caller passes them swapped.
#include "kernels.cuh"
#include "radix_sort.cuh"
#include "scan.cuh"

namespace spl {

namespace {

__global__ void __launch_bounds__(256)
product_count_kernel(uint32_t annz, const uint32_t *__restrict__ aind,
                     const uint32_t *__restrict__ bptr, uint32_t *__restrict__ cnt,
                     unsigned long long *__restrict__ total64) {
    const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t c = 0;
    if (p < annz) {
        const uint32_t k = aind[p];
        c = bptr[k + 1] - bptr[k];
        cnt[p] = c;
    }
    unsigned long long s = c;   // 64-bit total: the 32-bit scan below would wrap silently
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane_id() == 0 && s) atomicAdd(total64, s);
}

template <typename K, typename T>
__global__ void __launch_bounds__(256)
expand_kernel(uint32_t an, uint32_t annz, const uint32_t *__restrict__ aptr,
              const uint32_t *__restrict__ aind, const T *__restrict__ aval,
              const uint32_t *__restrict__ bptr, const uint32_t *__restrict__ bind,
              const T *__restrict__ bval, const uint32_t *__restrict__ off, int minor_bits,
              K *__restrict__ keys, T *__restrict__ vals) {
    const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= annz) return;
    const uint32_t i = upper_bound_u32(aptr, 0u, an + 1u, (uint32_t)p) - 1u;
    const uint32_t k = aind[p];
    const T a = aval[p];
    uint32_t o = off[p];
    for (uint32_t q = bptr[k]; q < bptr[k + 1]; ++q, ++o) {
        keys[o] = (K)(((uint64_t)i << minor_bits) | bind[q]);
        vals[o] = a * bval[q];   // rounded product; the add happens in seg_reduce (no FMA)
    }
}

template <typename K, typename T, typename VB>
spl_mat *spgemm_impl(spl_ctx *ctx, int format, int dtype, uint32_t out_rows, uint32_t out_cols,
                     uint32_t an, uint32_t bn, const spl_mat *a, const spl_mat *b) {
    // a: an segments over indices < ak; b: ak segments over indices < bn (both read as CSR)
    const uint32_t annz = a->nnz;
    Tmp<uint32_t> cnt(ctx, annz), off(ctx, (size_t)annz + 1);
    uint32_t total = 0;
    if (annz) {
        unsigned long long *total64 = reinterpret_cast<unsigned long long *>(ctx->d_scratch + 2);
        SPL_CUDA(cudaMemsetAsync(total64, 0, sizeof(unsigned long long), ctx->stream));
        product_count_kernel<<<div_up(annz, 256), 256, 0, ctx->stream>>>(annz, a->ind, b->ptr, cnt,
                                                                        total64);
        check_launch(ctx, "product_count");
        uint32_t t64[2];
        read_back(ctx, ctx->d_scratch + 2, t64, 2);
        SPL_REQUIRE(t64[1] == 0 && t64[0] < kMaxEntries, SPL_ERR_UNSUPPORTED,
                    "mul: more than 2^32 intermediate products (expand-sort-compress limit)");
        total = t64[0];
        exclusive_scan_u32(ctx, cnt, annz, off);
    }
    const int minor_bits = bits_for(bn);
    const int bits = bits_for(an) + minor_bits;
    const size_t sort_n = minor_bits == 0 ? 0 : total;      // no sort buffers on the n x 1 fast path
    Tmp<K> k0(ctx, sort_n), k1(ctx, sort_n), k2(ctx, total);
    Tmp<VB> v0(ctx, sort_n), v1(ctx, sort_n), v2(ctx, total);
    if (total) {
        expand_kernel<K, T><<<div_up(annz, 256), 256, 0, ctx->stream>>>(
            an, annz, a->ptr, a->ind, (const T *)a->val, b->ptr, b->ind, (const T *)b->val, off,
            minor_bits, k2, reinterpret_cast<T *>(v2.p));
        check_launch(ctx, "expand");
    }
    if (minor_bits == 0)     // B is n x 1 (the reference's SpMV route): the key is the row of A and the
        return finish_from_sorted(ctx, format, dtype, out_rows, out_cols, total, sizeof(K) == 8, k2.p,
                                  v2.p, minor_bits, /*dedup=*/1, /*dropzero=*/0);   // products are in row order
    K *kb[2] = {k0, k1};
    VB *vb[2] = {v0, v1};
    NoPayload *nb[2] = {nullptr, nullptr};
    const int r = radix_sort<K, VB, NoPayload>(ctx, total, bits, LoadPlain<K>{k2}, LoadPlain<VB>{v2},
                                               LoadNone{}, kb, vb, nb);
    return finish_from_sorted(ctx, format, dtype, out_rows, out_cols, total, sizeof(K) == 8, kb[r],
                              vb[r], minor_bits, /*dedup=*/1, /*dropzero=*/0);
}

}  // namespace

spl_mat *spgemm(spl_ctx *ctx, const spl_mat *a, const spl_mat *b) {
    SPL_REQUIRE(a->ncols == b->nrows, SPL_ERR_SHAPE,
                "mul: lhs.ncols != rhs.nrows (assert_eq!, src/csr/ops/mul.rs:9)");
    SPL_REQUIRE(a->format == b->format && a->dtype == b->dtype, SPL_ERR_ARG,
                "mul: operands must share format and scalar type");
    const uint32_t out_rows = a->nrows, out_cols = b->ncols;
    // read both as CSR: for CSC operands swap (see header comment)
    const spl_mat *l = a->format == SPL_CSR ? a : b;
    const spl_mat *r = a->format == SPL_CSR ? b : a;
    const uint32_t an = l->nmajor(), bn = r->nminor();
    const bool k64 = bits_for(an) + bits_for(bn) > 32;
    if (a->dtype == SPL_F32) {
        return k64 ? spgemm_impl<uint64_t, float, uint32_t>(ctx, a->format, a->dtype, out_rows,
                                                            out_cols, an, bn, l, r)
                   : spgemm_impl<uint32_t, float, uint32_t>(ctx, a->format, a->dtype, out_rows,
                                                            out_cols, an, bn, l, r);
    }
    return k64 ? spgemm_impl<uint64_t, double, uint64_t>(ctx, a->format, a->dtype, out_rows, out_cols,
                                                         an, bn, l, r)
               : spgemm_impl<uint32_t, double, uint64_t>(ctx, a->format, a->dtype, out_rows, out_cols,
                                                         an, bn, l, r);
}

}  // namespace spl
