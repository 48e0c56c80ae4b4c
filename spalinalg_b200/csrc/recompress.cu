// recompress.cu — the counting-sort re-compression behind CsrMatrix::transpose
// (src/csr.rs:358-406), CscMatrix::transpose (src/csc.rs:358-406) and the CSR<->CSC
// conversions (src/csc/conv/csr.rs:3-53, src/csr/conv/csc.rs:3-53): histogram of the minor
// index, exclusive scan, stable scatter (major ascending inside each output segment).
// Two device routes.  Near-diagonal matrices with short output segments: the three steps as they
// are, the scatter taking its slot from an atomic counter, followed by a per-segment repair that
// restores the major-ascending order (recompress_by_scatter).  Everything else: the three steps
// once per 8-bit digit of the minor index (stable LSD radix sort carrying (major, value)); only
// ceil(log2(nminor)) bits are sorted and the output pointer array is read off the sorted keys.
#include "kernels.cuh"
#include "radix_sort.cuh"
#include "scan.cuh"

#include <algorithm>

namespace spl {

namespace {

// ---- histogram + exclusive scan + scatter, made deterministic by a per-segment repair ---------
// The reference's own three steps (src/csr.rs:367-396).  On the device the scatter takes its slot
// from an atomic cursor, so entries reach their output segment (a column of the transposed /
// converted matrix) in no particular order; the repair pass then ranks every segment by major index,
// which restores exactly the order the reference's row-major sweep produces (majors ascend inside a
// segment; they are distinct because a CSR row holds a column at most once).  Used when the entries
// stay near the diagonal (banded, stencil: the scattered stores then merge in L2) and no segment is
// longer than 64 entries; random columns and heavy tails keep the stable radix passes.
constexpr uint32_t kRepairMax = 64;

__global__ void minor_count_kernel(const uint32_t *__restrict__ ind, uint32_t nnz, uint32_t *__restrict__ counts) {
    for (uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < nnz;
         p += (uint64_t)gridDim.x * blockDim.x)
        atomicAdd(counts + ind[p], 1u);
}

// How far the entries stray from the (scaled) diagonal: max over segments of |minor - major*ratio| at
// the segment's first and last entry (indices ascend inside a segment).  It bounds the window of
// output segments that rows processed at about the same time write to; if that window stays in L2 the
// scattered 4/8/12-byte stores merge there, otherwise (random columns) every one of them is a DRAM
// read-modify-write of a sector and the radix passes are several times faster (measured: config 3
// 6.7 ms by radix passes, 24.8 ms by scatter).
__global__ void bandwidth_kernel(const uint32_t *__restrict__ ptr, const uint32_t *__restrict__ ind, uint32_t nmajor,
                                 double ratio, uint32_t *__restrict__ out) {
    uint32_t b = 0;
    for (uint64_t m = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; m < nmajor; m += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t lo = ptr[m], hi = ptr[m + 1];
        if (hi > lo) {
            const double centre = (double)m * ratio;
            const double d0 = fabs((double)ind[lo] - centre), d1 = fabs((double)ind[hi - 1] - centre);
            b = max(b, (uint32_t)fmin(fmax(d0, d1), 4.0e9));
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) b = max(b, __shfl_xor_sync(0xffffffffu, b, o));
    if (lane_id() == 0 && b) atomicMax(out, b);
}

__global__ void max_u32_kernel(const uint32_t *__restrict__ a, uint32_t n, uint32_t *__restrict__ out) {
    uint32_t m = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        m = max(m, a[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane_id() == 0 && m) atomicMax(out, m);
}

// LPR lanes walk one input segment (row); every entry claims the next free slot of its output
// segment by counting the segment's remaining-entries counter down.
template <typename VB, int LPR>
__global__ void __launch_bounds__(256)
minor_scatter_kernel(const uint32_t *__restrict__ ptr, const uint32_t *__restrict__ ind,
                     const VB *__restrict__ val, uint32_t nmajor, const uint32_t *__restrict__ out_ptr,
                     uint32_t *__restrict__ remaining, uint32_t *__restrict__ out_ind, VB *__restrict__ out_val) {
    const uint64_t gtid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t m = gtid / LPR;
    if (m >= nmajor) return;
    const uint32_t e = __ldg(ptr + m + 1);
    // four entries in flight per lane: the chain index -> output pointer -> atomic slot -> stores is three
    // dependent round trips, and one entry at a time left the kernel latency-bound (config 2: 725 us)
    constexpr int U = 4;
    for (uint32_t p = __ldg(ptr + m) + (uint32_t)(gtid % LPR); p < e; p += U * LPR) {
        uint32_t c[U], slot[U];
        VB v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t q = p + u * LPR;
            const bool ok = q < e;
            c[u] = ok ? __ldg(ind + q) : 0xffffffffu;
            if (ok) v[u] = __ldg(val + q);
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (c[u] != 0xffffffffu) slot[u] = __ldg(out_ptr + c[u]);
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (c[u] != 0xffffffffu) slot[u] += atomicSub(remaining + c[u], 1u) - 1u;
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (c[u] != 0xffffffffu) {
                out_ind[slot[u]] = (uint32_t)m;
                out_val[slot[u]] = v[u];
            }
    }
}

// LPS lanes own one output segment of at most 2*LPS entries and put it in ascending major order.
template <typename VB, int LPS>
__global__ void __launch_bounds__(256)
segment_repair_kernel(const uint32_t *__restrict__ out_ptr, uint32_t nseg, uint32_t *__restrict__ out_ind,
                      VB *__restrict__ out_val) {
    const uint64_t gtid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t seg = gtid / LPS;
    const unsigned sub = (unsigned)(gtid % LPS);
    const unsigned lane = lane_id();
    const unsigned group_base = lane - sub;                 // first lane of this segment's group
    uint32_t lo = 0, len = 0;
    if (seg < nseg) {
        lo = __ldg(out_ptr + seg);
        len = __ldg(out_ptr + seg + 1) - lo;
    }
    uint32_t r[2] = {0xffffffffu, 0xffffffffu};
    VB v[2] = {};
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const uint32_t e = sub + u * LPS;
        if (e < len) { r[u] = out_ind[lo + e]; v[u] = out_val[lo + e]; }
    }
    uint32_t rank[2] = {0, 0};
#pragma unroll
    for (int u2 = 0; u2 < 2; ++u2) {
#pragma unroll
        for (int t = 0; t < LPS; ++t) {                     // every entry of the group, by shuffle
            const uint32_t other = __shfl_sync(0xffffffffu, r[u2], group_base + t);
            rank[0] += other < r[0];
            rank[1] += other < r[1];
        }
    }
    __syncwarp();
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const uint32_t e = sub + u * LPS;
        if (e < len) { out_ind[lo + rank[u]] = r[u]; out_val[lo + rank[u]] = v[u]; }
    }
}

template <typename VB, int LPR>
void launch_minor_scatter(spl_ctx *ctx, uint32_t nmajor, const uint32_t *ptr, const uint32_t *ind, const VB *val,
                          const uint32_t *out_ptr, uint32_t *remaining, uint32_t *out_ind, VB *out_val) {
    minor_scatter_kernel<VB, LPR><<<div_up((uint64_t)nmajor * LPR, 256), 256, 0, ctx->stream>>>(
        ptr, ind, val, nmajor, out_ptr, remaining, out_ind, out_val);
    check_launch(ctx, "minor_scatter");
}

// returns false (nothing written) when a segment is too long for the repair pass
template <typename VB>
bool recompress_by_scatter(spl_ctx *ctx, uint32_t nmajor, uint32_t nminor, uint32_t nnz, const uint32_t *ptr,
                           const uint32_t *ind, const VB *val, uint32_t *out_ptr, uint32_t *out_ind,
                           VB *out_val) {
    const unsigned sgrid = (unsigned)ctx->num_sms * 16u;
    {   // locality first: is the window of output segments written at one time small enough for L2?
        SPL_CUDA(cudaMemsetAsync(ctx->d_scratch, 0, sizeof(uint32_t), ctx->stream));
        bandwidth_kernel<<<std::min<unsigned>(div_up(nmajor, 256), sgrid), 256, 0, ctx->stream>>>(
            ptr, ind, nmajor, (double)nminor / (double)nmajor, ctx->d_scratch);
        check_launch(ctx, "bandwidth");
        uint32_t band = 0;
        read_back(ctx, ctx->d_scratch, &band, 1);
        const double mean_in = (double)nnz / nmajor;
        const double lanes = mean_in <= 2.0 ? 1.0 : mean_in <= 12.0 ? 4.0 : mean_in <= 64.0 ? 8.0 : 32.0;
        const double rows_in_flight = (double)ctx->num_sms * 2048.0 / lanes;           // resident threads / lanes per row
        const double in_flight_minors = 2.0 * band + rows_in_flight * ((double)nminor / (double)nmajor);
        const double window_bytes = in_flight_minors * ((double)nnz / nminor) * (4.0 + sizeof(VB));
        if (window_bytes > 48.0e6) return false;
    }
    Tmp<uint32_t> counts(ctx, nminor);
    SPL_CUDA(cudaMemsetAsync(counts, 0, sizeof(uint32_t) * (size_t)nminor, ctx->stream));
    SPL_CUDA(cudaMemsetAsync(ctx->d_scratch, 0, sizeof(uint32_t), ctx->stream));
    minor_count_kernel<<<sgrid, 256, 0, ctx->stream>>>(ind, nnz, counts);
    check_launch(ctx, "minor_count");
    max_u32_kernel<<<std::min<unsigned>(div_up(nminor, 256), sgrid), 256, 0, ctx->stream>>>(counts, nminor,
                                                                                         ctx->d_scratch);
    check_launch(ctx, "max_u32");
    uint32_t longest = 0;
    read_back(ctx, ctx->d_scratch, &longest, 1);
    if (longest > kRepairMax) return false;
    exclusive_scan_u32(ctx, counts, nminor, out_ptr);
    const double mean = (double)nnz / nmajor;
    if (mean <= 2.0) launch_minor_scatter<VB, 1>(ctx, nmajor, ptr, ind, val, out_ptr, counts, out_ind, out_val);
    else if (mean <= 12.0) launch_minor_scatter<VB, 4>(ctx, nmajor, ptr, ind, val, out_ptr, counts, out_ind, out_val);
    else if (mean <= 64.0) launch_minor_scatter<VB, 8>(ctx, nmajor, ptr, ind, val, out_ptr, counts, out_ind, out_val);
    else launch_minor_scatter<VB, 32>(ctx, nmajor, ptr, ind, val, out_ptr, counts, out_ind, out_val);
    if (longest > 1) {
        if (longest <= 16)
            segment_repair_kernel<VB, 8><<<div_up((uint64_t)nminor * 8, 256), 256, 0, ctx->stream>>>(
                out_ptr, nminor, out_ind, out_val);
        else if (longest <= 32)
            segment_repair_kernel<VB, 16><<<div_up((uint64_t)nminor * 16, 256), 256, 0, ctx->stream>>>(
                out_ptr, nminor, out_ind, out_val);
        else
            segment_repair_kernel<VB, 32><<<div_up((uint64_t)nminor * 32, 256), 256, 0, ctx->stream>>>(
                out_ptr, nminor, out_ind, out_val);
        check_launch(ctx, "segment_repair");
    }
    return true;
}

template <typename VB>
void recompress_impl(spl_ctx *ctx, uint32_t nmajor, uint32_t nminor, uint32_t nnz,
                     const uint32_t *ptr, const uint32_t *ind, const VB *val, uint32_t *out_ptr,
                     uint32_t *out_ind, VB *out_val) {
    if (nnz == 0) {
        SPL_CUDA(cudaMemsetAsync(out_ptr, 0, sizeof(uint32_t) * ((size_t)nminor + 1), ctx->stream));
        return;
    }
#ifndef SPL_NO_SCATTER_RECOMPRESS
    if (recompress_by_scatter<VB>(ctx, nmajor, nminor, nnz, ptr, ind, val, out_ptr, out_ind, out_val)) return;
#endif
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
