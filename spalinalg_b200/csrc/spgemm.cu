// spgemm.cu — C = A * B for compressed operands.
//
// Reference: impl Mul for &CsrMatrix (src/csr/ops/mul.rs:5-60) and &CscMatrix
// (src/csc/ops/mul.rs:5-61): Gustavson on the transposes with a dense accumulator; each
// C[i,j] = sum over ascending k of round(A[i,k]*B[k,j]), the first term copied, the pattern the
// structural union (explicit zeros kept).  Two device formulations: rows whose products fit a warp's
// shared-memory hash table accumulate row-wise in ascending k (spgemm_hash_kernel, no sort); anything
// else goes through expand - sort - compress:
// every A entry emits its products against B's row k in storage order, which is ascending k
// for a fixed (i,j); a stable sort by (i,j) keeps that order and the in-order segmented sum
// of assembly (dedup=1, dropzero=0) reproduces the accumulation bit for bit.
// CSC operands: CSC(A*B) arrays == CSR(B^T * A^T) arrays, same products and k order, so the
// caller passes them swapped.
#include "kernels.cuh"
#include "radix_sort.cuh"
#include "scan.cuh"

#include <algorithm>

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

// ---- row-wise hash accumulation (rows with at most 1 024 products) ---------------------------
// One warp owns one row i of C and a private hash table in shared memory (column j -> running
// value).  It walks A's row in storage order — ascending k, one entry per step — and in each step
// its lanes take the entries of B's row k: their columns are distinct, so within a step no two lanes
// touch the same cell, and across steps the warp is sequential.  Every C[i,j] therefore accumulates
// round(a*b) in ascending k with the first product copied, exactly the reference's order
// (src/csr/ops/mul.rs:25-40) — no sort of the products, which never leave the SM.  A symbolic
// pass (same walk, keys only) counts the distinct columns per row, the scan gives rowptr, the
// numeric pass accumulates, then ranks the row's columns (ascending, as the reference's final
// transpose leaves them) and writes them out.
// Two table sizes: 256 slots (rows of up to 128 products, 8 warps per CTA) and 2 048 slots (up to
// 1 024 products; 6 warps per CTA for f64, 8 for f32: one CTA fills an SM's shared memory).
constexpr uint32_t HS_EMPTY = 0xffffffffu;
constexpr uint32_t HS_SMALL_PRODUCTS = 128, HS_LARGE_PRODUCTS = 1024;

template <int SLOT_BITS>
__device__ __forceinline__ uint32_t hs_hash(uint32_t j) { return (j * 2654435761u) >> (32 - SLOT_BITS); }

// slot of column j in the warp's table, inserting it if absent; *fresh tells whether it was inserted
template <int SLOT_BITS>
__device__ __forceinline__ uint32_t hs_find_or_insert(uint32_t *keys, uint32_t j, bool *fresh) {
    uint32_t s = hs_hash<SLOT_BITS>(j);
    for (;;) {
        const uint32_t seen = atomicCAS(keys + s, HS_EMPTY, j);
        if (seen == HS_EMPTY) { *fresh = true; return s; }
        if (seen == j) { *fresh = false; return s; }
        s = (s + 1) & ((1u << SLOT_BITS) - 1u);
    }
}

__global__ void row_products_kernel(const uint32_t *__restrict__ aptr, const uint32_t *__restrict__ off,
                                    uint32_t an, uint32_t *__restrict__ max_out) {
    uint32_t m = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < an; i += (uint64_t)gridDim.x * blockDim.x)
        m = max(m, off[aptr[i + 1]] - off[aptr[i]]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane_id() == 0 && m) atomicMax(max_out, m);
}

template <typename T, bool NUMERIC, int SLOT_BITS, int WARPS>
constexpr size_t hs_smem_bytes() {
    constexpr size_t slots = (size_t)1 << SLOT_BITS, maxp = slots / 2;
    return WARPS * (slots * 4 + (NUMERIC ? slots * sizeof(T) + maxp * 4 + maxp * sizeof(T) : 0));
}

// Memory is touched in two round trips per 32 entries of A's row: the lanes fetch (k, a, B's row
// range) of one entry each, then every product of those entries is fetched at once — lane t takes
// product t, found by a search over the scanned row lengths — and only then do the steps run, in
// ascending k, out of registers.
template <typename T, bool NUMERIC, int SLOT_BITS, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
spgemm_hash_kernel(uint32_t an, const uint32_t *__restrict__ aptr, const uint32_t *__restrict__ aind,
                   const T *__restrict__ aval, const uint32_t *__restrict__ bptr,
                   const uint32_t *__restrict__ bind, const T *__restrict__ bval,
                   uint32_t *__restrict__ cnt, const uint32_t *__restrict__ cptr, uint32_t *__restrict__ cind,
                   T *__restrict__ cval) {
    constexpr uint32_t SLOTS = 1u << SLOT_BITS, MAXP = SLOTS / 2;
    extern __shared__ __align__(16) unsigned char hs_raw[];
    // layout: values first (8-byte aligned), then the 4-byte arrays
    T *vals_all = reinterpret_cast<T *>(hs_raw);
    T *dv_all = vals_all + (NUMERIC ? (size_t)WARPS * SLOTS : 0);
    uint32_t *keys_all = reinterpret_cast<uint32_t *>(dv_all + (NUMERIC ? (size_t)WARPS * MAXP : 0));
    uint32_t *dk_all = keys_all + (size_t)WARPS * SLOTS;
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
    uint32_t *keys = keys_all + (size_t)warp * SLOTS;
    T *vals = vals_all + (NUMERIC ? (size_t)warp * SLOTS : 0);
    for (uint64_t i = (uint64_t)blockIdx.x * WARPS + warp; i < an; i += (uint64_t)gridDim.x * WARPS) {
        for (uint32_t s = lane; s < SLOTS; s += 32) keys[s] = HS_EMPTY;
        __syncwarp();
        uint32_t mine = 0;                                  // columns this lane inserted
        const uint32_t pa = __ldg(aptr + i), ea = __ldg(aptr + i + 1);
        for (uint32_t p0 = pa; p0 < ea; p0 += 32) {         // 32 entries of A's row per trip, ascending k
            const uint32_t p = p0 + lane;
            const bool valid = p < ea;
            const uint32_t k = valid ? __ldg(aind + p) : 0u;
            T a = (T)0;
            if (NUMERIC && valid) a = __ldg(aval + p);
            const uint32_t qb = valid ? __ldg(bptr + k) : 0u;
            const uint32_t len = valid ? __ldg(bptr + k + 1) - qb : 0u;
            const uint32_t incl = warp_inclusive_scan(len);
            const uint32_t excl = incl - len;
            const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
            for (uint32_t t0 = 0; t0 < total; t0 += 32) {   // 32 products per batch, in emission order
                const uint32_t t = t0 + lane;
                const bool has = t < total;
                // step of product t = first lane whose inclusive count exceeds t
                uint32_t lo = 0, hi = 31;
#pragma unroll
                for (int it = 0; it < 5; ++it) {
                    const uint32_t mid = (lo + hi) >> 1;
                    const uint32_t im = __shfl_sync(0xffffffffu, incl, mid);
                    if (im <= t) lo = mid + 1; else hi = mid;
                }
                const uint32_t step = has ? lo : 31u;
                const uint32_t q = __shfl_sync(0xffffffffu, qb, step) + (t - __shfl_sync(0xffffffffu, excl, step));
                const T av = __shfl_sync(0xffffffffu, a, step);
                const uint32_t j = has ? __ldg(bind + q) : 0u;
                T prod = (T)0;
                if (NUMERIC && has) prod = av * __ldg(bval + q);     // rounded product (no FMA), then the add
                const uint32_t first = __shfl_sync(0xffffffffu, step, 0);
                const unsigned live = __ballot_sync(0xffffffffu, has);
                const uint32_t last = __shfl_sync(0xffffffffu, step, 31 - __clz(live));
                for (uint32_t sgo = first; sgo <= last; ++sgo) {      // one k at a time: columns distinct inside a step
                    if (has && step == sgo) {
                        bool fresh;
                        const uint32_t s = hs_find_or_insert<SLOT_BITS>(keys, j, &fresh);
                        mine += fresh;
                        if (NUMERIC) vals[s] = fresh ? prod : vals[s] + prod;
                    }
                    __syncwarp();                                     // the next k sees this step's cells
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
        if (!NUMERIC) {
            if (lane == 0) cnt[i] = mine;
            __syncwarp();
            continue;
        }
        // dense list of the stored columns, then rank = number of stored columns below: ascending order out
        uint32_t *dk = dk_all + (size_t)warp * MAXP;
        T *dv = dv_all + (size_t)warp * MAXP;
        uint32_t filled = 0;
        for (uint32_t c0 = 0; c0 < SLOTS; c0 += 32) {
            const uint32_t key = keys[c0 + lane];
            const bool occ = key != HS_EMPTY;
            const unsigned bal = __ballot_sync(0xffffffffu, occ);
            if (occ) {
                const uint32_t pos = filled + __popc(bal & lanemask_lt());
                dk[pos] = key;
                dv[pos] = vals[c0 + lane];
            }
            filled += __popc(bal);
        }
        __syncwarp();
        const uint32_t base = __ldg(cptr + i);
        for (uint32_t e = lane; e < mine; e += 32) {
            const uint32_t j = dk[e];
            uint32_t rank = 0;
            for (uint32_t t = 0; t < mine; ++t) rank += dk[t] < j;
            cind[base + rank] = j;
            cval[base + rank] = dv[e];
        }
        __syncwarp();
    }
}

template <typename T, int SLOT_BITS, int WARPS>
spl_mat *spgemm_hash(spl_ctx *ctx, int format, int dtype, uint32_t out_rows, uint32_t out_cols, uint32_t an,
                     const spl_mat *a, const spl_mat *b) {
    Tmp<uint32_t> cnt(ctx, an);
    Tmp<uint32_t> cptr(ctx, (size_t)an + 1 + 4);            // becomes the result's pointer array: same slack as new_mat
    auto ksym = spgemm_hash_kernel<T, false, SLOT_BITS, WARPS>;
    auto knum = spgemm_hash_kernel<T, true, SLOT_BITS, WARPS>;
    constexpr size_t sm_sym = hs_smem_bytes<T, false, SLOT_BITS, WARPS>();
    constexpr size_t sm_num = hs_smem_bytes<T, true, SLOT_BITS, WARPS>();
    SPL_CUDA(cudaFuncSetAttribute(ksym, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_sym));
    SPL_CUDA(cudaFuncSetAttribute(knum, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_num));
    SPL_CUDA(cudaFuncSetAttribute(knum, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    const unsigned grid = std::min<unsigned>(div_up(an, WARPS), (unsigned)ctx->num_sms * 8u);
    ksym<<<grid, WARPS * 32, sm_sym, ctx->stream>>>(an, a->ptr, a->ind, nullptr, b->ptr, b->ind, nullptr, cnt,
                                                    nullptr, nullptr, nullptr);
    check_launch(ctx, "spgemm_hash_symbolic");
    exclusive_scan_u32(ctx, cnt, an, cptr);
    uint32_t nnz = 0;
    read_back(ctx, cptr.p + an, &nnz, 1);      // exact-size output
    spl_mat *c = new_mat(ctx, format, dtype, out_rows, out_cols, nnz);
    dfree(ctx, c->ptr);
    c->ptr = cptr.release();
    try {
        knum<<<grid, WARPS * 32, sm_num, ctx->stream>>>(an, a->ptr, a->ind, (const T *)a->val, b->ptr, b->ind,
                                                        (const T *)b->val, nullptr, c->ptr, c->ind, (T *)c->val);
        check_launch(ctx, "spgemm_hash_numeric");
    } catch (...) {
        free_mat(ctx, c);
        throw;
    }
    return c;
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
        // rows whose products fit the warp's hash table take the row-wise path (no sort)
        SPL_CUDA(cudaMemsetAsync(ctx->d_scratch, 0, sizeof(uint32_t), ctx->stream));
        row_products_kernel<<<std::min<unsigned>(div_up(an, 256), (unsigned)ctx->num_sms * 8u), 256, 0, ctx->stream>>>(
            a->ptr, off, an, ctx->d_scratch);
        check_launch(ctx, "row_products");
        uint32_t max_products = 0;
        read_back(ctx, ctx->d_scratch, &max_products, 1);
#ifndef SPL_NO_HASH_SPGEMM
        if (max_products <= HS_SMALL_PRODUCTS && bn > 1)
            return spgemm_hash<T, 8, 8>(ctx, format, dtype, out_rows, out_cols, an, a, b);
        if (max_products <= HS_LARGE_PRODUCTS && bn > 1)
            return spgemm_hash<T, 11, (sizeof(T) == 8 ? 6 : 8)>(ctx, format, dtype, out_rows, out_cols, an, a, b);
#endif
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
