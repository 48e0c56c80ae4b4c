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

template <typename K, typename VB, typename LoadK>
spl_mat *assemble_impl(spl_ctx *ctx, int format, int dtype, uint32_t nrows, uint32_t ncols,
                       uint32_t len, LoadK lk, const VB *val, int minor_bits, int bits, int dedup,
                       int dropzero) {
    Tmp<K> k0(ctx, len), k1(ctx, len);
    Tmp<VB> v0(ctx, len), v1(ctx, len);
    K *kb[2] = {k0, k1};
    VB *vb[2] = {v0, v1};
    NoPayload *nb[2] = {nullptr, nullptr};
    LoadPlain<VB> lv{val};
    const int r = radix_sort<K, VB, NoPayload>(ctx, len, bits, lk, lv, LoadNone{}, kb, vb, nb);
    if (dtype == SPL_F32)
        return finish_impl<K, float>(ctx, format, dtype, nrows, ncols, len, kb[r],
                                     reinterpret_cast<float *>(vb[r]), minor_bits, dedup, dropzero);
    return finish_impl<K, double>(ctx, format, dtype, nrows, ncols, len, kb[r],
                                  reinterpret_cast<double *>(vb[r]), minor_bits, dedup, dropzero);
}

// ---- row-sharded assembly: routing of triplets to their owners ----------------
// owner of a major index: the block [starts[g], starts[g+1]) that contains it (world <= 8)
struct Owners {
    uint32_t start[SPL_MAX_PEERS + 1];
    int world;
    __device__ __forceinline__ uint32_t of(uint32_t major) const {
        uint32_t g = 0;
#pragma unroll
        for (int q = 1; q < SPL_MAX_PEERS; ++q) g += (q < world && major >= start[q]) ? 1u : 0u;
        return g;
    }
};
struct LoadOwner {
    const uint32_t *major;
    Owners own;
    __device__ __forceinline__ uint32_t operator()(uint32_t i, uint32_t &) const {
        return own.of(major[i]);
    }
};
// key the owner will sort: (major - owner's first major) << minor_bits | minor
struct LoadPackLocal {
    const uint32_t *major;
    const uint32_t *minor;
    Owners own;
    int minor_bits;
    __device__ __forceinline__ uint64_t operator()(uint32_t i, uint32_t &) const {
        const uint32_t m = major[i];
        return ((uint64_t)(m - own.start[own.of(m)]) << minor_bits) | (uint64_t)minor[i];
    }
};
template <typename K>
struct LoadNarrow {   // packed 64-bit keys that fit K
    const uint64_t *p;
    __device__ __forceinline__ K operator()(uint32_t i, uint32_t &) const { return (K)p[i]; }
};

__global__ void owner_count_kernel(const uint32_t *__restrict__ major, uint32_t n, Owners own,
                                   uint32_t *__restrict__ counts) {
    __shared__ uint32_t s[SPL_MAX_PEERS];
    if (threadIdx.x < SPL_MAX_PEERS) s[threadIdx.x] = 0;
    __syncthreads();
    uint32_t c[SPL_MAX_PEERS] = {};
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t g = own.of(major[i]);
#pragma unroll
        for (int q = 0; q < SPL_MAX_PEERS; ++q) c[q] += g == (uint32_t)q;
    }
#pragma unroll
    for (int q = 0; q < SPL_MAX_PEERS; ++q) {
        uint32_t v = c[q];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane_id() == 0 && v) atomicAdd(&s[q], v);
    }
    __syncthreads();
    if (threadIdx.x < SPL_MAX_PEERS && s[threadIdx.x]) atomicAdd(counts + threadIdx.x, s[threadIdx.x]);
}

__global__ void packed_bounds_kernel(const uint64_t *__restrict__ keys, uint32_t n, int minor_bits,
                                     uint32_t nmajor, uint32_t nminor, uint32_t *flag) {
    uint32_t bad = 0;
    const uint64_t mask = minor_bits >= 64 ? ~0ull : ((1ull << minor_bits) - 1ull);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t k = keys[i];
        bad |= ((k >> minor_bits) >= nmajor) | ((k & mask) >= nminor);
    }
    if (__any_sync(0xffffffffu, bad) && lane_id() == 0) atomicOr(flag, 1u);
}

void check_coo_bounds(spl_ctx *ctx, uint32_t len, const uint32_t *row, const uint32_t *col,
                      uint32_t nrows, uint32_t ncols) {
    if (len == 0) return;
    SPL_CUDA(cudaMemsetAsync(ctx->d_scratch, 0, sizeof(uint32_t), ctx->stream));
    unsigned grid = div_up(len, 256 * 8);
    coo_bounds_kernel<<<grid, 256, 0, ctx->stream>>>(row, col, len, nrows, ncols, ctx->d_scratch);
    check_launch(ctx, "coo_bounds");
    uint32_t bad = 0;
    read_back(ctx, ctx->d_scratch, &bad, 1);
    SPL_REQUIRE(bad == 0, SPL_ERR_ARG,
                "COO entry out of bounds (CooMatrix::push asserts row < nrows, col < ncols)");
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
    check_coo_bounds(ctx, len, row, col, nrows, ncols);
    const uint32_t *major_idx = format == SPL_CSR ? row : col;
    const uint32_t *minor_idx = format == SPL_CSR ? col : row;
    const int major_bits = bits_for(format == SPL_CSR ? nrows : ncols);
    const int minor_bits = bits_for(format == SPL_CSR ? ncols : nrows);
    const int bits = major_bits + minor_bits;
    if (bits <= 32) {
        LoadPack<uint32_t> lk{major_idx, minor_idx, minor_bits};
        if (dtype == SPL_F32)
            return assemble_impl<uint32_t, uint32_t>(ctx, format, dtype, nrows, ncols, len, lk,
                                                     (const uint32_t *)val, minor_bits, bits, dedup,
                                                     dropzero);
        return assemble_impl<uint32_t, uint64_t>(ctx, format, dtype, nrows, ncols, len, lk,
                                                 (const uint64_t *)val, minor_bits, bits, dedup,
                                                 dropzero);
    }
    LoadPack<uint64_t> lk{major_idx, minor_idx, minor_bits};
    if (dtype == SPL_F32)
        return assemble_impl<uint64_t, uint32_t>(ctx, format, dtype, nrows, ncols, len, lk,
                                                 (const uint32_t *)val, minor_bits, bits, dedup,
                                                 dropzero);
    return assemble_impl<uint64_t, uint64_t>(ctx, format, dtype, nrows, ncols, len, lk,
                                             (const uint64_t *)val, minor_bits, bits, dedup, dropzero);
}

// Packed keys as produced by route_coo_dev on the sending ranks (and concatenated in source-rank
// order by the all-to-all, which keeps the global insertion order among duplicates).
spl_mat *assemble_from_packed_dev(spl_ctx *ctx, int format, int dtype, uint32_t nrows, uint32_t ncols,
                                  uint32_t len, const uint64_t *keys, const void *val, int dedup,
                                  int dropzero) {
    const uint32_t nmajor = format == SPL_CSR ? nrows : ncols;
    const uint32_t nminor = format == SPL_CSR ? ncols : nrows;
    const int minor_bits = bits_for(nminor);
    const int bits = bits_for(nmajor) + minor_bits;
    if (len) {
        SPL_CUDA(cudaMemsetAsync(ctx->d_scratch, 0, sizeof(uint32_t), ctx->stream));
        packed_bounds_kernel<<<div_up(len, 256 * 8), 256, 0, ctx->stream>>>(keys, len, minor_bits, nmajor,
                                                                          nminor, ctx->d_scratch);
        check_launch(ctx, "packed_bounds");
        uint32_t bad = 0;
        read_back(ctx, ctx->d_scratch, &bad, 1);
        SPL_REQUIRE(bad == 0, SPL_ERR_ARG, "packed COO key out of bounds for this shard");
    }
    if (bits <= 32) {
        LoadNarrow<uint32_t> lk{keys};
        if (dtype == SPL_F32)
            return assemble_impl<uint32_t, uint32_t>(ctx, format, dtype, nrows, ncols, len, lk,
                                                     (const uint32_t *)val, minor_bits, bits, dedup,
                                                     dropzero);
        return assemble_impl<uint32_t, uint64_t>(ctx, format, dtype, nrows, ncols, len, lk,
                                                 (const uint64_t *)val, minor_bits, bits, dedup,
                                                 dropzero);
    }
    LoadPlain<uint64_t> lk{keys};
    if (dtype == SPL_F32)
        return assemble_impl<uint64_t, uint32_t>(ctx, format, dtype, nrows, ncols, len, lk,
                                                 (const uint32_t *)val, minor_bits, bits, dedup,
                                                 dropzero);
    return assemble_impl<uint64_t, uint64_t>(ctx, format, dtype, nrows, ncols, len, lk,
                                             (const uint64_t *)val, minor_bits, bits, dedup, dropzero);
}

// Stable partition of this rank's triplets by the rank that owns their major index: one radix
// pass over the owner id carrying (packed local key, value).  Stability keeps every owner's
// share in insertion order.  counts[g] = entries routed to rank g.
template <typename VB>
static void route_impl(spl_ctx *ctx, uint32_t len, const uint32_t *major, const uint32_t *minor,
                       const VB *val, const Owners &own, int minor_bits, uint64_t *out_keys,
                       VB *out_val) {
    Tmp<uint32_t> okeys(ctx, len);
    uint32_t *kb[2] = {okeys, nullptr};
    uint64_t *ab[2] = {out_keys, nullptr};
    VB *vb[2] = {out_val, nullptr};
    LoadOwner lo{major, own};
    LoadPackLocal lp{major, minor, own, minor_bits};
    LoadPlain<VB> lv{val};
    radix_sort<uint32_t, uint64_t, VB>(ctx, len, bits_for((uint64_t)own.world), lo, lp, lv, kb, ab, vb);
}

void route_coo_dev(spl_ctx *ctx, int format, int dtype, uint32_t nrows, uint32_t ncols, uint32_t len,
                   const uint32_t *row, const uint32_t *col, const void *val, int world,
                   const uint64_t *major_starts, uint64_t *out_keys, void *out_val,
                   uint64_t *counts_host) {
    SPL_REQUIRE(world >= 1 && world <= SPL_MAX_PEERS, SPL_ERR_ARG, "world must be 1..8");
    const uint32_t nmajor = format == SPL_CSR ? nrows : ncols;
    SPL_REQUIRE(major_starts[0] == 0 && major_starts[world] == nmajor, SPL_ERR_ARG,
                "major_starts must run from 0 to the number of rows (CSR) / columns (CSC)");
    Owners own;
    own.world = world;
    for (int g = 0; g <= SPL_MAX_PEERS; ++g) {
        const uint64_t v = major_starts[g < world ? g : world];
        SPL_REQUIRE(g == 0 || g > world || v >= major_starts[g - 1], SPL_ERR_ARG,
                    "major_starts must be non-decreasing");
        own.start[g] = (uint32_t)v;
    }
    for (int g = 0; g < world; ++g) counts_host[g] = 0;
    if (len == 0) return;
    check_coo_bounds(ctx, len, row, col, nrows, ncols);
    const uint32_t *major = format == SPL_CSR ? row : col;
    const uint32_t *minor = format == SPL_CSR ? col : row;
    const int minor_bits = bits_for(format == SPL_CSR ? ncols : nrows);
    uint32_t *cnt = ctx->d_scratch + 8;
    SPL_CUDA(cudaMemsetAsync(cnt, 0, SPL_MAX_PEERS * sizeof(uint32_t), ctx->stream));
    unsigned grid = div_up(len, 256 * 8);
    if (grid > (unsigned)ctx->num_sms * 8u) grid = (unsigned)ctx->num_sms * 8u;
    owner_count_kernel<<<grid, 256, 0, ctx->stream>>>(major, len, own, cnt);
    check_launch(ctx, "owner_count");
    if (dtype == SPL_F32)
        route_impl<uint32_t>(ctx, len, major, minor, (const uint32_t *)val, own, minor_bits, out_keys,
                             (uint32_t *)out_val);
    else
        route_impl<uint64_t>(ctx, len, major, minor, (const uint64_t *)val, own, minor_bits, out_keys,
                             (uint64_t *)out_val);
    uint32_t c[SPL_MAX_PEERS];
    read_back(ctx, cnt, c, SPL_MAX_PEERS);
    for (int g = 0; g < world; ++g) counts_host[g] = c[g];
}

}  // namespace spl
