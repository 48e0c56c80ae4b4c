// recompress.cu — the counting-sort re-compression behind CsrMatrix::transpose
// (src/csr.rs:358-406), CscMatrix::transpose (src/csc.rs:358-406) and the CSR<->CSC
// conversions (src/csc/conv/csr.rs:3-53, src/csr/conv/csc.rs:3-53): histogram of the minor
// index, exclusive scan, stable scatter (major ascending inside each output segment).
// On the device the three steps run once per 8-bit digit of the minor index (stable LSD radix
// sort carrying (major, value)); only ceil(log2(nminor)) bits are sorted.  The output pointer
// array is read off the sorted minor keys.
#include "kernels.cuh"
#include "radix_sort.cuh"

namespace spl {

namespace {

template <typename VB>
void recompress_impl(spl_ctx *ctx, uint32_t nmajor, uint32_t nminor, uint32_t nnz,
                     const uint32_t *ptr, const uint32_t *ind, const VB *val, uint32_t *out_ptr,
                     uint32_t *out_ind, VB *out_val) {
    if (nnz == 0) {
        SPL_CUDA(cudaMemsetAsync(out_ptr, 0, sizeof(uint32_t) * ((size_t)nminor + 1), ctx->stream));
        return;
    }
    const int bits = bits_for(nminor);
    const int passes = rs_num_passes(bits);
    Tmp<uint32_t> k0(ctx, nnz), k1(ctx, nnz), m_tmp(ctx, passes > 1 ? nnz : 1);
    Tmp<VB> v_tmp(ctx, passes > 1 ? nnz : 1);
    // the last pass must land in the caller's arrays: pass p writes set p&1
    const int last = (passes - 1) & 1;
    uint32_t *kb[2] = {k0, k1};
    uint32_t *mb[2];
    VB *vb[2];
    mb[last] = out_ind;
    vb[last] = out_val;
    mb[last ^ 1] = m_tmp;
    vb[last ^ 1] = v_tmp;
    // major index per entry: one coalesced expansion pass, then every radix pass reads plain arrays
    Tmp<uint32_t> major(ctx, nnz);
    expand_major(ctx, nmajor, nnz, ptr, major);
    LoadPlain<uint32_t> lk{ind};
    LoadPlain<uint32_t> lm{major};
    LoadPlain<VB> lv{val};
    const int r = radix_sort<uint32_t, uint32_t, VB>(ctx, nnz, bits, lk, lm, lv, kb, mb, vb);
    fill_ptr(ctx, kb[r], nnz, nminor, out_ptr);
}

}  // namespace

void recompress(spl_ctx *ctx, int dtype, uint32_t nmajor, uint32_t nminor, uint32_t nnz,
                const uint32_t *ptr, const uint32_t *ind, const void *val, uint32_t *out_ptr,
                uint32_t *out_ind, void *out_val) {
    if (dtype == SPL_F32)
        recompress_impl<uint32_t>(ctx, nmajor, nminor, nnz, ptr, ind, (const uint32_t *)val, out_ptr,
                                  out_ind, (uint32_t *)out_val);
    else
        recompress_impl<uint64_t>(ctx, nmajor, nminor, nnz, ptr, ind, (const uint64_t *)val, out_ptr,
                                  out_ind, (uint64_t *)out_val);
}

}  // namespace spl
