// assemble.cu — COO -> CSR / CSC assembly on the device.
//
// Reference semantics (src/csr/conv/coo.rs:3-116, src/csc/conv/coo.rs:3-116): entries ordered
// by (major, minor); duplicates of one cell added left to right in insertion order, the first
// value copied (:43-52); cells whose sum == 0 removed (:60-73); exactly sized outputs.
// Device formulation: stable LSD radix sort of key = major << minor_bits | minor carrying the
// value, then one sequential in-order sum per run of equal keys (a tree reduction would change
// the rounding and, through the zero drop, the structure), flag, compact, build the pointers.
#include "kernels.cuh"
#include "radix_sort.cuh"
#include "scan.cuh"

namespace spl {

namespace {

constexpr int CP_THREADS = 256;
constexpr int CP_IPT = 16;
constexpr int CP_TILE = CP_THREADS * CP_IPT;

// CooMatrix::push bounds (src/coo.rs:432-433) re-checked for raw ABI callers.
__global__ void coo_bounds_kernel(const uint32_t *__restrict__ row, const uint32_t *__restrict__ col,
                                  uint32_t n, uint32_t nrows, uint32_t ncols, uint32_t *flag) {
    uint32_t bad = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (uint64_t)gridDim.x * blockDim.x)
        bad |= (row[i] >= nrows) | (col[i] >= ncols);
    if (__any_sync(0xffffffffu, bad) && lane_id() == 0) atomicOr(flag, 1u);
}

// One thread per record; the thread on a run head walks its run and adds in order.
// flags[i] = 1 iff record i is a head whose (summed) value survives.  vals updated in place
// at heads only (non-heads are never written, so concurrent readers see the inputs).
template <typename K, typename T>
__global__ void __launch_bounds__(256)
seg_reduce_kernel(const K *__restrict__ keys, T *vals, uint32_t n, int dedup, int dropzero,
                  uint8_t *__restrict__ flags) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const K k = keys[i];
    bool head = true;
    if (dedup && i > 0) head = keys[i - 1] != k;
    uint8_t keep = 0;
    if (head) {
        T s = vals[i];
        if (dedup) {
            uint64_t j = i + 1;
            bool grew = false;
            while (j < n && keys[j] == k) {
                s = s + vals[j];          // one rounding per addend, insertion order
                ++j;
                grew = true;
            }
            if (grew) vals[i] = s;
        }
        keep = dropzero ? (s != (T)0) : 1;   // -0.0 and +0.0 dropped, NaN kept
    }
    flags[i] = keep;
}

__global__ void __launch_bounds__(CP_THREADS)
flag_count_kernel(const uint8_t *__restrict__ flags, uint32_t n, uint32_t *__restrict__ tile_sums) {
    __shared__ uint32_t ws[CP_THREADS / 32 + 1];
    const uint64_t base = (uint64_t)blockIdx.x * CP_TILE + (uint64_t)threadIdx.x * CP_IPT;
    uint32_t c = 0;
    if (base + CP_IPT <= n) {
        uint4 v = __ldg(reinterpret_cast<const uint4 *>(flags + base));
        c = __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w);
    } else {
        for (int i = 0; i < CP_IPT; ++i)
            if (base + i < n) c += flags[base + i];
    }
    uint32_t total;
    block_exclusive_scan(c, ws, &total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// Survivors to their final slots: minor index, value, and (temporarily) the major index.
template <typename K, typename T>
__global__ void __launch_bounds__(CP_THREADS)
compact_kernel(const uint8_t *__restrict__ flags, const K *__restrict__ keys,
               const T *__restrict__ vals, uint32_t n, const uint32_t *__restrict__ tile_offsets,
               int minor_bits, uint32_t *__restrict__ out_ind, T *__restrict__ out_val,
               uint32_t *__restrict__ out_major) {
    __shared__ uint32_t ws[CP_THREADS / 32 + 1];
    const uint64_t base = (uint64_t)blockIdx.x * CP_TILE + (uint64_t)threadIdx.x * CP_IPT;
    uint8_t f[CP_IPT];
    uint32_t c = 0;
#pragma unroll
    for (int i = 0; i < CP_IPT; ++i) {
        f[i] = (base + i < n) ? flags[base + i] : 0;
        c += f[i];
    }
    uint32_t pos = block_exclusive_scan(c, ws, nullptr) + tile_offsets[blockIdx.x];
    const K minor_mask = (minor_bits >= (int)(8 * sizeof(K))) ? ~(K)0 : (((K)1 << minor_bits) - 1);
#pragma unroll
    for (int i = 0; i < CP_IPT; ++i) {
        if (f[i]) {
            const K k = keys[base + i];
            out_ind[pos] = (uint32_t)(k & minor_mask);
            out_major[pos] = minor_bits >= (int)(8 * sizeof(K)) ? 0u : (uint32_t)(k >> minor_bits);
            out_val[pos] = vals[base + i];
            ++pos;
        }
    }
}

__global__ void fill_ptr_kernel(const uint32_t *__restrict__ sorted_major, uint32_t nnz,
                                uint32_t nmajor, uint32_t *__restrict__ ptr) {
    const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p > nnz) return;
    // entries p-1 and p bracket the majors whose segment starts at p
    const int64_t lo = p == 0 ? 0 : (int64_t)sorted_major[p - 1] + 1;
    const int64_t hi = p == nnz ? (int64_t)nmajor : (int64_t)sorted_major[p];
    for (int64_t q = lo; q <= hi; ++q) ptr[q] = (uint32_t)p;
}

template <typename K, typename T>
spl_mat *finish_impl(spl_ctx *ctx, int format, int dtype, uint32_t nrows, uint32_t ncols, uint32_t n,
                     const K *keys, T *vals, int minor_bits, int dedup, int dropzero) {
    const uint32_t nmajor = format == SPL_CSR ? nrows : ncols;
    if (n == 0) {   // empty COO => all-zero pointer array (coo.rs:16-22 with len == 0)
        spl_mat *e = new_mat(ctx, format, dtype, nrows, ncols, 0);
        cudaMemsetAsync(e->ptr, 0, sizeof(uint32_t) * ((size_t)nmajor + 1), ctx->stream);
        return e;
    }
    Tmp<uint8_t> flags(ctx, (size_t)n + CP_IPT);
    seg_reduce_kernel<K, T><<<div_up(n, 256), 256, 0, ctx->stream>>>(keys, vals, n, dedup, dropzero,
                                                                     flags);
    check_launch(ctx, "seg_reduce");
    const unsigned tiles = div_up(n, CP_TILE);
    Tmp<uint32_t> sums(ctx, tiles + 1);
    flag_count_kernel<<<tiles, CP_THREADS, 0, ctx->stream>>>(flags, n, sums);
    check_launch(ctx, "flag_count");
    scan_spine_kernel<<<1, 1024, 0, ctx->stream>>>(sums, tiles);
    check_launch(ctx, "scan_spine");
    uint32_t nnz = 0;
    read_back(ctx, sums.p + tiles, &nnz, 1);   // the one host sync of assembly: exact-size output

    spl_mat *m = new_mat(ctx, format, dtype, nrows, ncols, nnz);
    try {
        Tmp<uint32_t> major(ctx, nnz);
        compact_kernel<K, T><<<tiles, CP_THREADS, 0, ctx->stream>>>(
            flags, keys, vals, n, sums, minor_bits, m->ind, static_cast<T *>(m->val), major);
        check_launch(ctx, "compact");
        fill_ptr(ctx, major, nnz, nmajor, m->ptr);
    } catch (...) {
        free_mat(ctx, m);
        throw;
    }
    return m;
}

template <typename K, typename VB>
spl_mat *assemble_impl(spl_ctx *ctx, int format, int dtype, uint32_t nrows, uint32_t ncols,
                       uint32_t len, const uint32_t *major_idx, const uint32_t *minor_idx,
                       const VB *val, int minor_bits, int bits, int dedup, int dropzero) {
    Tmp<K> k0(ctx, len), k1(ctx, len);
    Tmp<VB> v0(ctx, len), v1(ctx, len);
    K *kb[2] = {k0, k1};
    VB *vb[2] = {v0, v1};
    NoPayload *nb[2] = {nullptr, nullptr};
    LoadPack<K> lk{major_idx, minor_idx, minor_bits};
    LoadPlain<VB> lv{val};
    const int r = radix_sort<K, VB, NoPayload>(ctx, len, bits, lk, lv, LoadNone{}, kb, vb, nb);
    if (dtype == SPL_F32)
        return finish_impl<K, float>(ctx, format, dtype, nrows, ncols, len, kb[r],
                                     reinterpret_cast<float *>(vb[r]), minor_bits, dedup, dropzero);
    return finish_impl<K, double>(ctx, format, dtype, nrows, ncols, len, kb[r],
                                  reinterpret_cast<double *>(vb[r]), minor_bits, dedup, dropzero);
}

}  // namespace

void fill_ptr(spl_ctx *ctx, const uint32_t *sorted_major, uint32_t nnz, uint32_t nmajor,
              uint32_t *ptr) {
    fill_ptr_kernel<<<div_up((uint64_t)nnz + 1, 256), 256, 0, ctx->stream>>>(sorted_major, nnz,
                                                                            nmajor, ptr);
    check_launch(ctx, "fill_ptr");
}

spl_mat *finish_from_sorted(spl_ctx *ctx, int format, int dtype, uint32_t nrows, uint32_t ncols,
                            uint32_t n, bool key64, const void *keys, void *vals, int minor_bits,
                            int dedup, int dropzero) {
    if (key64) {
        if (dtype == SPL_F32)
            return finish_impl<uint64_t, float>(ctx, format, dtype, nrows, ncols, n,
                                                (const uint64_t *)keys, (float *)vals, minor_bits,
                                                dedup, dropzero);
        return finish_impl<uint64_t, double>(ctx, format, dtype, nrows, ncols, n,
                                             (const uint64_t *)keys, (double *)vals, minor_bits,
                                             dedup, dropzero);
    }
    if (dtype == SPL_F32)
        return finish_impl<uint32_t, float>(ctx, format, dtype, nrows, ncols, n,
                                            (const uint32_t *)keys, (float *)vals, minor_bits, dedup,
                                            dropzero);
    return finish_impl<uint32_t, double>(ctx, format, dtype, nrows, ncols, n, (const uint32_t *)keys,
                                         (double *)vals, minor_bits, dedup, dropzero);
}

spl_mat *assemble_from_coo_dev(spl_ctx *ctx, int format, int dtype, uint32_t nrows, uint32_t ncols,
                               uint32_t len, const uint32_t *row, const uint32_t *col,
                               const void *val, int dedup, int dropzero) {
    if (len > 0) {
        SPL_CUDA(cudaMemsetAsync(ctx->d_scratch, 0, sizeof(uint32_t), ctx->stream));
        unsigned grid = div_up(len, 256 * 8);
        coo_bounds_kernel<<<grid, 256, 0, ctx->stream>>>(row, col, len, nrows, ncols, ctx->d_scratch);
        check_launch(ctx, "coo_bounds");
        uint32_t bad = 0;
        read_back(ctx, ctx->d_scratch, &bad, 1);
        SPL_REQUIRE(bad == 0, SPL_ERR_ARG,
                    "COO entry out of bounds (CooMatrix::push asserts row < nrows, col < ncols)");
    }
    const uint32_t *major_idx = format == SPL_CSR ? row : col;
    const uint32_t *minor_idx = format == SPL_CSR ? col : row;
    const int major_bits = bits_for(format == SPL_CSR ? nrows : ncols);
    const int minor_bits = bits_for(format == SPL_CSR ? ncols : nrows);
    const int bits = major_bits + minor_bits;
    if (bits <= 32) {
        if (dtype == SPL_F32)
            return assemble_impl<uint32_t, uint32_t>(ctx, format, dtype, nrows, ncols, len, major_idx,
                                                     minor_idx, (const uint32_t *)val, minor_bits,
                                                     bits, dedup, dropzero);
        return assemble_impl<uint32_t, uint64_t>(ctx, format, dtype, nrows, ncols, len, major_idx,
                                                 minor_idx, (const uint64_t *)val, minor_bits, bits,
                                                 dedup, dropzero);
    }
    if (dtype == SPL_F32)
        return assemble_impl<uint64_t, uint32_t>(ctx, format, dtype, nrows, ncols, len, major_idx,
                                                 minor_idx, (const uint32_t *)val, minor_bits, bits,
                                                 dedup, dropzero);
    return assemble_impl<uint64_t, uint64_t>(ctx, format, dtype, nrows, ncols, len, major_idx,
                                             minor_idx, (const uint64_t *)val, minor_bits, bits, dedup,
                                             dropzero);
}

}  // namespace spl
