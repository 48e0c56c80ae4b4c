// assemble.cu — COO -> CSR / CSC assembly on the device.
//
// Reference semantics (src/csr/conv/coo.rs:3-116, src/csc/conv/coo.rs:3-116): entries ordered
// by (major, minor); duplicates of one cell added left to right in insertion order, the first
// value copied (:43-52); cells whose sum == 0 removed (:60-73); exactly sized outputs.
// Device formulation: one pass over the indices checks bounds and detects already sorted input (no
// sort then).  Otherwise the key major << minor_bits | minor is sorted stably, carrying the value:
// large lists by the hybrid route (global radix passes on the high key bits, packing the key on the
// fly; then one CTA per block of whole rows finishes the sort in shared memory AND runs the tail
// there), the rest by full LSD radix passes and the streaming tail.  The tail is one sequential
// in-order sum per run of equal keys (a tree reduction would change the rounding and, through the
// zero drop, the structure), flag, compact, build the pointers.
#include <cstdlib>

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

// Pointer entries are written by the record that starts a major: ptr[lo..hi] = position.  A run of
// empty rows/columns can be arbitrarily long (one stored entry in a 10^8-row matrix), and one thread
// must not write millions of entries: ranges of 64 or more go to a queue that fill_gaps_kernel
// empties with a CTA per range.
struct GapQueue {
    uint32_t *count;     // number of queued ranges (may exceed cap: the excess was written inline)
    uint4 *items;        // (lo, hi, value, -)
    uint32_t cap;
};
constexpr uint32_t kGapCap = 16384;

__device__ __noinline__ void emit_ptr_long(uint32_t *__restrict__ ptr, uint64_t lo, uint64_t hi, uint32_t value,
                                           const GapQueue &gq) {
    if (gq.items) {
        const uint32_t slot = atomicAdd(gq.count, 1u);
        if (slot < gq.cap) {
            gq.items[slot] = make_uint4((uint32_t)lo, (uint32_t)hi, value, 0u);
            return;
        }
    }
    for (uint64_t q = lo; q <= hi; ++q) ptr[q] = value;
}
// ptr[lo..hi] = value (hi inclusive; nothing if hi < lo).  The common cases — no entry, one entry, a
// few — stay inline and 32-bit; only long ranges take the queue.
__device__ __forceinline__ void emit_ptr(uint32_t *__restrict__ ptr, uint64_t lo, uint64_t hi, uint32_t value,
                                         const GapQueue &gq) {
    if (hi < lo) return;
    if (hi - lo >= 64) { emit_ptr_long(ptr, lo, hi, value, gq); return; }
    const uint32_t l = (uint32_t)lo, n = (uint32_t)(hi - lo) + 1u;     // lo + n - 1 <= nmajor < 2^32
    for (uint32_t q = 0; q < n; ++q) ptr[l + q] = value;
}

__global__ void __launch_bounds__(256) fill_gaps_kernel(uint32_t *__restrict__ ptr, GapQueue gq) {
    const uint32_t n = min(*gq.count, gq.cap);
    for (uint32_t g = blockIdx.x; g < n; g += gridDim.x) {
        const uint4 it = gq.items[g];
        for (uint64_t q = (uint64_t)it.x + threadIdx.x; q <= (uint64_t)it.y; q += blockDim.x) ptr[q] = it.z;
    }
}

// Owns the queue's temporary storage for one pointer build.
struct GapQueueOwner {
    Tmp<uint32_t> mem;
    GapQueue q;
    GapQueueOwner(spl_ctx *ctx) : mem(ctx, 4 + 4 * (size_t)kGapCap) {
        q.count = mem.p;
#ifdef SPL_NO_GAPQ
        q.items = nullptr;
#else
        q.items = reinterpret_cast<uint4 *>(mem.p + 4);
#endif
        q.cap = kGapCap;
        cudaMemsetAsync(mem.p, 0, sizeof(uint32_t), ctx->stream);
    }
    void drain(spl_ctx *ctx, uint32_t *ptr) {
        fill_gaps_kernel<<<(unsigned)ctx->num_sms * 2u, 256, 0, ctx->stream>>>(ptr, q);
        check_launch(ctx, "fill_gaps");
    }
};

// Tile = CP_THREADS x CP_IPT consecutive records, warp-striped: in step i thread t looks at record
// tile_base + i*CP_THREADS + t, so every load and store below is coalesced, and all of a thread's
// loads are issued before the first dependent instruction (in-order issue would otherwise expose
// one memory latency per record).

// Pass 1 of the tail.  A record whose key differs from its predecessor's heads a run of equal keys
// (one matrix cell); the head adds the run left to right, one rounding per addend, in insertion
// order (src/csr/conv/coo.rs:43-52) and stores the sum in place.  flags[i] = 1 iff record i is a
// head whose sum survives the zero test (:60-73: -0.0 and +0.0 dropped, NaN kept).  tile_sums[b] =
// survivors of tile b.
template <typename K, typename T>
__global__ void __launch_bounds__(CP_THREADS)
seg_reduce_kernel(const K *__restrict__ keys, T *vals, uint32_t n, int dedup, int dropzero,
                  uint8_t *__restrict__ flags, uint32_t *__restrict__ tile_sums) {
    __shared__ uint32_t ws[CP_THREADS / 32 + 1];
    const uint64_t tile_base = (uint64_t)blockIdx.x * CP_TILE;
    const unsigned lane = lane_id();
    K k[CP_IPT], kn[CP_IPT], kp[CP_IPT];
    T s[CP_IPT];
#pragma unroll
    for (int i = 0; i < CP_IPT; ++i) {
        const uint64_t idx = tile_base + (uint64_t)i * CP_THREADS + threadIdx.x;
        const bool ok = idx < n;
        k[i] = ok ? keys[idx] : (K)0;
        s[i] = ok ? vals[idx] : (T)0;
        // neighbours that live in another warp's (or tile's) registers come from memory
        kp[i] = (lane == 0 && ok && idx > 0) ? keys[idx - 1] : (K)0;
        kn[i] = (lane == 31 && idx + 1 < n) ? keys[idx + 1] : (K)0;
    }
    uint32_t kept = 0;
#pragma unroll
    for (int i = 0; i < CP_IPT; ++i) {
        const uint64_t idx = tile_base + (uint64_t)i * CP_THREADS + threadIdx.x;
        const bool ok = idx < n;
        const K up = __shfl_up_sync(0xffffffffu, k[i], 1), dn = __shfl_down_sync(0xffffffffu, k[i], 1);
        const K prev = lane == 0 ? kp[i] : up, next = lane == 31 ? kn[i] : dn;
        const bool head = ok && (!dedup || idx == 0 || prev != k[i]);
        if (head && dedup && idx + 1 < n && next == k[i]) {          // duplicates: rare, walk the run
            T acc = s[i];
            for (uint64_t j = idx + 1; j < n && keys[j] == k[i]; ++j) acc = acc + vals[j];
            s[i] = acc;
            vals[idx] = acc;        // only heads are written; non-heads keep their input value
        }
        const bool keep = head && (!dropzero || s[i] != (T)0);
        if (ok) flags[idx] = keep ? 1 : 0;
        kept += keep;
    }
    uint32_t total;
    block_exclusive_scan(kept, ws, &total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// Pass 2 of the tail: survivors to their final slots (minor index, value) and the pointer array.
// Position of record idx = tile offset + survivors before it in the tile, from 16 ballots per warp
// and one scan of the (step, warp) counts.  ptr[q] = first final position whose major is >= q: the
// first record of every major (survivor or not) writes the entries for the majors it skips.
template <typename K, typename T>
__global__ void __launch_bounds__(CP_THREADS)
compact_kernel(const uint8_t *__restrict__ flags, const K *__restrict__ keys,
               const T *__restrict__ vals, uint32_t n, const uint32_t *__restrict__ tile_offsets,
               int minor_bits, uint32_t nmajor, uint32_t *__restrict__ out_ind, T *__restrict__ out_val,
               uint32_t *__restrict__ ptr, GapQueue gq) {
    constexpr int W = CP_THREADS / 32;
    __shared__ uint32_t s_cnt[CP_IPT * W + 1];
    const uint64_t tile_base = (uint64_t)blockIdx.x * CP_TILE;
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
    const bool wide = minor_bits >= (int)(8 * sizeof(K));
    const K minor_mask = wide ? ~(K)0 : (((K)1 << minor_bits) - 1);
    K k[CP_IPT], kp[CP_IPT];
    T v[CP_IPT];
    uint32_t below[CP_IPT];
    uint32_t fbits = 0;
#pragma unroll
    for (int i = 0; i < CP_IPT; ++i) {
        const uint64_t idx = tile_base + (uint64_t)i * CP_THREADS + threadIdx.x;
        const bool ok = idx < n;
        const bool f = ok && flags[idx];
        fbits |= (uint32_t)f << i;
        k[i] = ok ? keys[idx] : (K)0;
        v[i] = f ? vals[idx] : (T)0;
        kp[i] = (lane == 0 && ok && idx > 0) ? keys[idx - 1] : (K)0;
    }
#pragma unroll
    for (int i = 0; i < CP_IPT; ++i) {
        const unsigned bal = __ballot_sync(0xffffffffu, (fbits >> i) & 1u);
        below[i] = __popc(bal & lanemask_lt());
        if (lane == 0) s_cnt[i * W + warp] = __popc(bal);
    }
    __syncthreads();
    if (warp == 0) {                       // exclusive scan of the CP_IPT*W (= 128) counts: 4 per lane
        constexpr int PER = CP_IPT * W / 32;
        uint32_t c[PER], sum = 0;
#pragma unroll
        for (int q = 0; q < PER; ++q) { c[q] = s_cnt[lane * PER + q]; sum += c[q]; }
        uint32_t run = warp_inclusive_scan(sum) - sum;
#pragma unroll
        for (int q = 0; q < PER; ++q) { s_cnt[lane * PER + q] = run; run += c[q]; }
    }
    __syncthreads();
    const uint32_t off = tile_offsets[blockIdx.x];
#pragma unroll
    for (int i = 0; i < CP_IPT; ++i) {
        const uint64_t idx = tile_base + (uint64_t)i * CP_THREADS + threadIdx.x;
        const K up = __shfl_up_sync(0xffffffffu, k[i], 1);
        if (idx >= n) continue;
        const uint32_t pos = off + s_cnt[i * W + warp] + below[i];      // survivors before record idx
        if ((fbits >> i) & 1u) {
            out_ind[pos] = (uint32_t)(k[i] & minor_mask);
            out_val[pos] = v[i];
        }
        const K prev = lane == 0 ? kp[i] : up;
        const uint32_t mj = wide ? 0u : (uint32_t)(k[i] >> minor_bits);
        const uint32_t mp = wide ? 0u : (uint32_t)(prev >> minor_bits);
        // majors (mp, mj] start at pos; record 0 also writes major 0 .. mj = 0
        if (idx == 0) emit_ptr(ptr, 0, mj, 0u, gq);
        else emit_ptr(ptr, (uint64_t)mp + 1, mj, pos, gq);
        if (idx == (uint64_t)n - 1) emit_ptr(ptr, (uint64_t)mj + 1, nmajor, pos + ((fbits >> i) & 1u), gq);
    }
}

__global__ void fill_ptr_kernel(const uint32_t *__restrict__ sorted_major, uint32_t nnz,
                                uint32_t nmajor, uint32_t *__restrict__ ptr, GapQueue gq) {
    const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p > nnz) return;
    // entries p-1 and p bracket the majors whose segment starts at p
    const uint64_t lo = p == 0 ? 0 : (uint64_t)sorted_major[p - 1] + 1;
    const uint64_t hi = p == nnz ? (uint64_t)nmajor : (uint64_t)sorted_major[p];
    emit_ptr(ptr, lo, hi, (uint32_t)p, gq);
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
    const unsigned tiles = div_up(n, CP_TILE);
    Tmp<uint32_t> sums(ctx, tiles + 1);
    seg_reduce_kernel<K, T><<<tiles, CP_THREADS, 0, ctx->stream>>>(keys, vals, n, dedup, dropzero, flags,
                                                                   sums);
    check_launch(ctx, "seg_reduce");
    scan_spine_kernel<<<1, 1024, 0, ctx->stream>>>(sums, tiles);
    check_launch(ctx, "scan_spine");
    uint32_t nnz = 0;
    read_back(ctx, sums.p + tiles, &nnz, 1);   // the one host sync of assembly: exact-size output

    spl_mat *m = new_mat(ctx, format, dtype, nrows, ncols, nnz);
    try {
        GapQueueOwner gaps(ctx);
        compact_kernel<K, T><<<tiles, CP_THREADS, 0, ctx->stream>>>(
            flags, keys, vals, n, sums, minor_bits, nmajor, m->ind, static_cast<T *>(m->val), m->ptr, gaps.q);
        check_launch(ctx, "compact");
        gaps.drain(ctx, m->ptr);
    } catch (...) {
        free_mat(ctx, m);
        throw;
    }
    return m;
}

// First pass over the caller's triplets, reading the indices only: the bounds of CooMatrix::push
// (src/coo.rs:432-433), whether the packed keys major << minor_bits | minor are already
// non-decreasing and whether at least the majors are.  A sorted list — triplets emitted row by row,
// column by column, as element loops and stencil generators do — needs no sort at all: a stable
// sort would leave it where it is.  The general route never materialises the unsorted keys: its
// first radix pass packs them on the fly from the caller's arrays (LoadPack) and reads the values
// in place.
template <typename K>
__global__ void __launch_bounds__(256)
coo_scan_kernel(const uint32_t *__restrict__ major, const uint32_t *__restrict__ minor, uint32_t n,
                uint32_t nmajor, uint32_t nminor, int minor_bits, uint32_t *__restrict__ flags) {
    uint32_t bad = 0, unsorted = 0, major_unsorted = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t mj = major[i], mn = minor[i];
        bad |= (mj >= nmajor) | (mn >= nminor);
        if (i + 1 < n) {
            const uint32_t mjn = major[i + 1];
            const K k = (K)(((uint64_t)mj << minor_bits) | (uint64_t)mn);
            const K kn = (K)(((uint64_t)mjn << minor_bits) | (uint64_t)minor[i + 1]);
            unsorted |= kn < k;
            major_unsorted |= mjn < mj;
        }
    }
    if (__any_sync(0xffffffffu, bad) && lane_id() == 0) atomicOr(flags, 1u);
    if (__any_sync(0xffffffffu, unsorted) && lane_id() == 0) atomicOr(flags + 1, 1u);
    if (__any_sync(0xffffffffu, major_unsorted) && lane_id() == 0) atomicOr(flags + 2, 1u);
}

// Packed keys and a private copy of the values for the routes that skip the global sort (the tail
// sums in place and the caller's arrays must stay untouched).
template <typename K, typename VB>
__global__ void __launch_bounds__(256)
pack_kernel(const uint32_t *__restrict__ major, const uint32_t *__restrict__ minor, const VB *__restrict__ val,
            uint32_t n, int minor_bits, K *__restrict__ keys, VB *__restrict__ vals) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (uint64_t)gridDim.x * blockDim.x) {
        keys[i] = (K)(((uint64_t)major[i] << minor_bits) | (uint64_t)minor[i]);
        vals[i] = val[i];
    }
}

// Triplets that already come major by major (row by row) only need each segment put in (minor,
// position) order: LPS lanes own one segment of at most 2*LPS records and rank it by shuffles.  The
// position is the tie-break, so duplicates of a cell keep their insertion order (what the stable sort
// guarantees on the general path).  Out of place: (keys, vals) -> (out_k, out_v).
template <typename K, typename VB, int LPS>
__global__ void __launch_bounds__(256)
segment_sort_kernel(const uint32_t *__restrict__ segptr, uint32_t nseg, const K *__restrict__ keys,
                    const VB *__restrict__ vals, K *__restrict__ out_k, VB *__restrict__ out_v) {
    const uint64_t gtid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t seg = gtid / LPS;
    const unsigned sub = (unsigned)(gtid % LPS);
    const unsigned group_base = lane_id() - sub;
    uint32_t lo = 0, len = 0;
    if (seg < nseg) {
        lo = __ldg(segptr + seg);
        len = __ldg(segptr + seg + 1) - lo;
    }
    K k[2] = {~(K)0, ~(K)0};
    VB v[2] = {};
    bool have[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const uint32_t e = sub + u * LPS;
        have[u] = e < len;
        if (have[u]) { k[u] = keys[lo + e]; v[u] = vals[lo + e]; }
    }
    uint32_t rank[2] = {0, 0};
#pragma unroll
    for (int u2 = 0; u2 < 2; ++u2) {
#pragma unroll
        for (int t = 0; t < LPS; ++t) {
            const K other = __shfl_sync(0xffffffffu, k[u2], group_base + t);
            const bool there = __shfl_sync(0xffffffffu, (int)have[u2], group_base + t) != 0;
            const uint32_t opos = t + u2 * LPS;
#pragma unroll
            for (int u = 0; u < 2; ++u)
                rank[u] += there && (other < k[u] || (other == k[u] && opos < sub + u * LPS));
        }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u)
        if (have[u]) { out_k[lo + rank[u]] = k[u]; out_v[lo + rank[u]] = v[u]; }
}

__global__ void max_seglen_counts_kernel(const uint32_t *__restrict__ counts, uint32_t n, uint32_t *__restrict__ out) {
    uint32_t m = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        m = max(m, counts[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane_id() == 0 && m) atomicMax(out, m);
}

__global__ void max_seglen_kernel(const uint32_t *__restrict__ segptr, uint32_t nseg, uint32_t *__restrict__ out) {
    uint32_t m = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nseg; i += (uint64_t)gridDim.x * blockDim.x)
        m = max(m, segptr[i + 1] - segptr[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane_id() == 0 && m) atomicMax(out, m);
}

// ---- hybrid sort: global passes on the high key bits, the rest inside shared memory -----------
// The radix passes are bound by instructions and latency, not by bytes, so the way to get faster is
// to run fewer of them.  Only the top H key bits (high bits of the major index) are sorted globally:
// that cuts the list into blocks of a few thousand records — whole rows, still in insertion order
// because the passes are stable.  One CTA then finishes a block in shared memory: count per row
// (atomics), scan, drop every record into its row's segment (atomic slot: arrival order does not
// matter), and rank each record inside its row by (minor, position) — the position breaks ties, so
// duplicates of a cell leave in insertion order, exactly what the stable sort of the full key gives.
// Two global passes instead of five or six on the benchmark shapes.  Blocks are sized for ~2 700
// records; if any block would exceed the 4 096 the kernel holds (skewed rows), the full radix sort
// runs instead — decided from an exact histogram of the block ids, after a cheap sampled one.
constexpr uint32_t BL_CAP = 4096;             // records a block kernel holds (256 threads)
constexpr uint32_t BL_CAP_BIG = 8192;         // ... with 512 threads: f32 values and 32-bit in-block keys only
constexpr uint32_t BL_TARGET = 2700;
constexpr int BL_MAX_ROW_BITS = 10;            // at most 1 024 rows per block

template <typename K, typename LoadK>
__global__ void block_hist_kernel(LoadK lk, uint32_t n, int S, uint32_t step, uint32_t *__restrict__ bcount) {
    uint32_t state = 0;
    for (uint64_t i = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * step; i < n;
         i += (uint64_t)gridDim.x * blockDim.x * step)
        atomicAdd(bcount + (uint32_t)((K)lk((uint32_t)i, state) >> S), 1u);
}

// LK: type of the key inside the block.  All records of a block share the key bits above S, so
// only the low S bits (local row, minor) are kept — 32 bits whenever S <= 32, which halves the
// shared memory per record and raises the number of resident CTAs.
//
// The kernel is the whole tail of the assembly for its block, not only the sort: a block holds
// whole rows, so every run of equal keys (one matrix cell) is inside it.  After the ranking the
// sorted order sits in shared memory; the head of each run adds the run left to right in insertion
// order (src/csr/conv/coo.rs:43-52, one rounding per addend), the zero test runs on the rounded sum
// (:60-73), and the survivors leave compacted to the front of the block's own range of the
// temporary arrays, as (minor index, value) — 4 + V bytes instead of the sorted 8 + V.  What is
// left for the global level is the survivor count per block, the pointer entries relative to the
// block, and one gather pass (block_gather_kernel) once the exact nnz is known.
// resident CTAs per SM: what the shared memory allows (227 KB, ~10.5 KB static + reserved per CTA), at most 1 024 threads
constexpr int bl_min_ctas(int cap, int rec_bytes) {
    const int by_smem = (227 * 1024) / (cap * rec_bytes + 10752);
    const int by_threads = 16384 / cap;
    return by_smem < 1 ? 1 : (by_smem < by_threads ? by_smem : by_threads);
}

template <typename K, typename T, typename LK, int CAP>
__global__ void __launch_bounds__(CAP / 16, bl_min_ctas(CAP, (int)(sizeof(LK) + sizeof(T) + 2)))
block_finish_kernel(const K *__restrict__ keys, const T *__restrict__ vals, const uint32_t *__restrict__ bptr,
                    int minor_bits, int row_bits, int dedup, int dropzero, uint32_t nmajor,
                    uint32_t *__restrict__ tmp_ind, T *__restrict__ tmp_val, uint32_t *__restrict__ block_kept,
                    uint32_t *__restrict__ local_ptr) {
    constexpr int IPT = 16;                                   // sorted positions per thread
    constexpr int THREADS = CAP / IPT;
    constexpr int W = THREADS / 32;
    extern __shared__ __align__(16) unsigned char bl_raw[];   // CAP * (sizeof(LK) + sizeof(T) + 2) bytes
    constexpr bool kValFirst = sizeof(T) > sizeof(LK);        // the wider array first: alignment
    // values stay in ARRIVAL order (coalesced in); keys go to slot order (grouped by row), then
    // to sorted order in place
    T *s_val = reinterpret_cast<T *>(bl_raw + (kValFirst ? 0 : sizeof(LK) * CAP));
    LK *s_key = reinterpret_cast<LK *>(bl_raw + (kValFirst ? sizeof(T) * CAP : 0));
    // arrival index of the record in a slot (the tie-break); later: of the record at a sorted position
    uint16_t *s_arr = reinterpret_cast<uint16_t *>(bl_raw + (sizeof(LK) + sizeof(T)) * CAP);
    __shared__ uint32_t s_off[(1 << BL_MAX_ROW_BITS) + 1];
    __shared__ uint32_t s_cur[1 << BL_MAX_ROW_BITS];
    __shared__ uint32_t ws[W + 1];
    __shared__ uint32_t s_cnt[IPT * W + 1];                   // survivors before each (step, warp)
    __shared__ uint32_t s_bal[IPT * W];                       // survivor lanes of each (step, warp)
    const uint32_t R = 1u << row_bits;
    const uint32_t lo = bptr[blockIdx.x], cnt = bptr[blockIdx.x + 1] - lo;
    const uint64_t first_row = (uint64_t)blockIdx.x << row_bits;
    if (cnt == 0) {
        for (uint32_t r = threadIdx.x; r < R; r += THREADS)
            if (first_row + r < nmajor) local_ptr[first_row + r] = 0u;
        if (threadIdx.x == 0) block_kept[blockIdx.x] = 0u;
        return;
    }
    const int S = minor_bits + row_bits;
    const K low_mask = S >= (int)(8 * sizeof(K)) ? ~(K)0 : (((K)1 << S) - 1);
    const LK minor_mask = minor_bits >= (int)(8 * sizeof(LK)) ? ~(LK)0 : (LK)(((LK)1 << minor_bits) - 1);
    auto row_of = [&](LK k) -> uint32_t { return row_bits ? (uint32_t)(k >> minor_bits) : 0u; };
    for (uint32_t r = threadIdx.x; r < R; r += THREADS) s_cur[r] = 0;
    // all of a thread's loads are issued before the first dependent instruction
    LK kreg[IPT];
    {
        T vreg[IPT];
#pragma unroll
        for (int i = 0; i < IPT; ++i) {
            const uint32_t e = threadIdx.x + (uint32_t)i * THREADS;
            kreg[i] = e < cnt ? (LK)(keys[lo + e] & low_mask) : (LK)0;
            vreg[i] = e < cnt ? vals[lo + e] : (T)0;
        }
        __syncthreads();                                          // counters are zero
#pragma unroll
        for (int i = 0; i < IPT; ++i) {
            const uint32_t e = threadIdx.x + (uint32_t)i * THREADS;
            if (e < cnt) {
                s_val[e] = vreg[i];
                atomicAdd(&s_cur[row_of(kreg[i])], 1u);           // records per row
            }
        }
    }
    __syncthreads();
    {   // exclusive scan of the row counts (R <= 1024), counters reset for the placement
        constexpr int PER = (1 << BL_MAX_ROW_BITS) / THREADS;
        uint32_t c[PER], sum = 0;
#pragma unroll
        for (int q = 0; q < PER; ++q) {
            const uint32_t r = threadIdx.x * PER + q;
            c[q] = r < R ? s_cur[r] : 0u;
            sum += c[q];
        }
        uint32_t run = block_exclusive_scan(sum, ws, nullptr);
#pragma unroll
        for (int q = 0; q < PER; ++q) {
            const uint32_t r = threadIdx.x * PER + q;
            if (r < R) { s_off[r] = run; s_cur[r] = 0; }
            run += c[q];
        }
        if (threadIdx.x == THREADS - 1) s_off[R] = run;
    }
    __syncthreads();
    // every key (and its arrival index) into a slot of its row's segment; the order inside a
    // segment is whatever the atomics give
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
        const uint32_t e = threadIdx.x + (uint32_t)i * THREADS;
        if (e < cnt) {
            const uint32_t r = row_of(kreg[i]);
            const uint32_t slot = s_off[r] + atomicAdd(&s_cur[r], 1u);
            s_key[slot] = kreg[i];
            s_arr[slot] = (uint16_t)e;
        }
    }
    __syncthreads();
    // rank inside the row by (minor, arrival): one shared-memory word per comparison, no branch;
    // arrival indices are looked at only when the row really holds the key more than once (ncu: the
    // kernel is issue-bound on this loop, and with 5 % duplicates nearly every warp has one lane
    // that needs the tie-break).  Key,
    // sorted position and arrival index stay in registers until every thread is done reading.
    uint32_t packed[IPT];
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
        const uint32_t slot = threadIdx.x + (uint32_t)i * THREADS;
        packed[i] = 0xffffffffu;
        if (slot < cnt) {
            const LK k = s_key[slot];
            const uint32_t e = s_arr[slot];
            const uint32_t r = row_of(k);
            const uint32_t a = s_off[r], b = s_off[r + 1];
            uint32_t lt = 0, eq = 0, xs = 0;
            for (uint32_t t = a; t < b; ++t) {
                const LK kt = s_key[t];
                const bool same = kt == k;
                lt += kt < k;
                eq += same;
                xs ^= same ? t : 0u;                      // slots holding this key, xor-ed (own slot included)
            }
            if (eq == 2)                                  // the usual duplicate: one partner, found without a second pass
                lt += s_arr[xs ^ slot] < e;
            else if (eq > 2)
                for (uint32_t t = a; t < b; ++t) lt += (s_key[t] == k) & (s_arr[t] < e);
            kreg[i] = k;
            packed[i] = ((a + lt) << 16) | e;
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < IPT; ++i)
        if (packed[i] != 0xffffffffu) {                           // keys and arrival indices in sorted order
            s_key[packed[i] >> 16] = kreg[i];
            s_arr[packed[i] >> 16] = (uint16_t)(packed[i] & 0xffffu);
        }
    __syncthreads();
    // heads of runs of equal keys: in-order sum (stored in the head's own slot), zero test
    uint32_t keepbits = 0;
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
        const uint32_t p = threadIdx.x + (uint32_t)i * THREADS;
        if (p < cnt) {
            const LK k = s_key[p];
            const bool head = !dedup || p == 0 || s_key[p - 1] != k;
            if (head) {
                const uint32_t e = s_arr[p];
                T acc = s_val[e];
                if (dedup) {
                    bool more = false;
                    for (uint32_t j = p + 1; j < cnt && s_key[j] == k; ++j) {
                        acc = acc + s_val[s_arr[j]];
                        more = true;
                    }
                    if (more) s_val[e] = acc;      // nobody else reads a head's value
                }
                keepbits |= (uint32_t)(!dropzero || acc != (T)0) << i;
            }
        }
    }
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
        const unsigned bal = __ballot_sync(0xffffffffu, (keepbits >> i) & 1u);
        if (lane == 0) { s_cnt[i * W + warp] = __popc(bal); s_bal[i * W + warp] = bal; }
    }
    __syncthreads();
    if (warp == 0) {                       // exclusive scan of the IPT*W counts: PER per lane
        constexpr int PER = IPT * W / 32;
        uint32_t c[PER], sum = 0;
#pragma unroll
        for (int q = 0; q < PER; ++q) { c[q] = s_cnt[lane * PER + q]; sum += c[q]; }
        const uint32_t incl = warp_inclusive_scan(sum);
        uint32_t run = incl - sum;
#pragma unroll
        for (int q = 0; q < PER; ++q) { s_cnt[lane * PER + q] = run; run += c[q]; }
        if (lane == 31) s_cnt[IPT * W] = incl;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
        if ((keepbits >> i) & 1u) {
            const uint32_t p = threadIdx.x + (uint32_t)i * THREADS;
            const uint32_t lp = s_cnt[i * W + warp] + __popc(s_bal[i * W + warp] & lanemask_lt());
            tmp_ind[lo + lp] = (uint32_t)(s_key[p] & minor_mask);
            tmp_val[lo + lp] = s_val[s_arr[p]];
        }
    }
    // pointer entries relative to the block: survivors before the first sorted position of the row
    const uint32_t total = s_cnt[IPT * W];
    for (uint32_t r = threadIdx.x; r < R; r += THREADS) {
        if (first_row + r >= nmajor) break;
        const uint32_t p = s_off[r];
        uint32_t before = total;
        if (p < cnt) {
            const uint32_t cell = (p / THREADS) * W + ((p % THREADS) >> 5);
            before = s_cnt[cell] + __popc(s_bal[cell] & ((1u << (p & 31u)) - 1u));
        }
        local_ptr[first_row + r] = before;
    }
    if (threadIdx.x == 0) block_kept[blockIdx.x] = total;
}

// Survivors of every block to their final place (the exact nnz is known by now) and the pointer
// array: base of the block + the entry relative to it.  One CTA per block; coalesced both ways.
template <typename T>
__global__ void __launch_bounds__(256)
block_gather_kernel(const uint32_t *__restrict__ tmp_ind, const T *__restrict__ tmp_val,
                    const uint32_t *__restrict__ bptr, const uint32_t *__restrict__ block_base,
                    const uint32_t *__restrict__ local_ptr, int row_bits, uint32_t nmajor,
                    uint32_t *__restrict__ out_ind, T *__restrict__ out_val, uint32_t *__restrict__ ptr) {
    const uint32_t lo = bptr[blockIdx.x], base = block_base[blockIdx.x];
    const uint32_t kept = block_base[blockIdx.x + 1] - base;
#pragma unroll 4
    for (uint32_t i = threadIdx.x; i < kept; i += 256) {
        out_ind[base + i] = tmp_ind[lo + i];
        out_val[base + i] = tmp_val[lo + i];
    }
    const uint32_t R = 1u << row_bits;
    const uint64_t first_row = (uint64_t)blockIdx.x << row_bits;
    for (uint32_t r = threadIdx.x; r < R; r += 256) {
        if (first_row + r >= nmajor) break;
        ptr[first_row + r] = base + local_ptr[first_row + r];
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) ptr[nmajor] = block_base[gridDim.x];
}

// bptr[b] = first position whose block id is >= b, read off the block-sorted keys (same bracket
// logic as fill_ptr_kernel; block ids ascend)
template <typename K>
__global__ void __launch_bounds__(256)
block_bounds_kernel(const K *__restrict__ keys, uint32_t n, int S, uint32_t nblocks,
                    uint32_t *__restrict__ bptr, GapQueue gq) {
    constexpr int U = 4;                                   // positions per thread, loads issued together
    const uint64_t base = ((uint64_t)blockIdx.x * blockDim.x) * U + threadIdx.x;
    uint32_t cur[U], prev[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const uint64_t p = base + (uint64_t)u * blockDim.x;
        cur[u] = p < n ? (uint32_t)(keys[p] >> S) : nblocks;             // position n closes the last block
        prev[u] = (p > 0 && p <= n) ? (uint32_t)(keys[p - 1] >> S) : 0u;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const uint64_t p = base + (uint64_t)u * blockDim.x;
        if (p > n) continue;
        const uint64_t lo = p == 0 ? 0 : (uint64_t)prev[u] + 1;
        if (lo <= (uint64_t)cur[u]) emit_ptr(bptr, lo, cur[u], (uint32_t)p, gq);   // most positions: same block, nothing to do
    }
}

// Assembles by the hybrid route.  Input: the loaders (lk, lv) over the caller's triplets; (k0, v0)
// and (k1, v1) are scratch pairs.  On success *result is the finished matrix (the tail runs inside
// the block kernel).  On failure *result stays NULL: either nothing was touched (*sorted_k == NULL,
// the caller sorts from the loaders), or a block would not fit (skewed rows) and the records are
// left, in a stable order, in *sorted_k / *sorted_v for the caller's full radix sort.
template <typename K, typename VB, typename LoadK, typename LoadV>
void hybrid_sort(spl_ctx *ctx, int format, int dtype, uint32_t nrows, uint32_t ncols, int dedup, int dropzero,
                 uint32_t len, uint32_t nmajor, int major_bits, int minor_bits, LoadK lk, LoadV lv, K *k0, VB *v0,
                 K *k1, VB *v1, K **sorted_k, VB **sorted_v, spl_mat **result) {
    *sorted_k = nullptr;
    *sorted_v = nullptr;
    *result = nullptr;
    // small lists: the extra launches cost more (SPL_HYBRID_MIN_LEN lowers the bar: tests of this route at oracle sizes)
    const char *knob = std::getenv("SPL_HYBRID_MIN_LEN");
    const uint32_t min_len = knob ? (uint32_t)std::strtoul(knob, nullptr, 10) : (1u << 22);
    if (len < min_len || major_bits < 1) return;
    // rows per block: the largest power of two that keeps the average block at BL_TARGET records
    const double rows_per_block = (double)BL_TARGET * (double)nmajor / (double)len;
    if (rows_per_block < 1.0) return;                              // rows longer than a block on average
    int row_bits = 0;
    while (row_bits < BL_MAX_ROW_BITS && (double)(2u << row_bits) <= rows_per_block) ++row_bits;
    if (row_bits > major_bits) row_bits = major_bits;
    // Blocks twice as large (and twice the threads) when that saves a whole global pass and the block
    // still fits two to an SM: 4-byte values, 32-bit in-block keys (config 3: 17 -> 16 key bits).
    uint32_t cap = BL_CAP;
    if (sizeof(VB) == 4 && row_bits + 1 <= BL_MAX_ROW_BITS && row_bits + 1 <= major_bits &&
        minor_bits + row_bits + 1 <= 32 && major_bits - row_bits - 1 >= 1 &&
        rs_num_passes(major_bits - row_bits - 1) < rs_num_passes(major_bits - row_bits)) {
        ++row_bits;
        cap = BL_CAP_BIG;
    }
    const int H = major_bits - row_bits;                           // key bits sorted globally
    if (H < 1 || H > 24) return;
    const int S = minor_bits + row_bits;                           // block id = key >> S
    const uint32_t nblocks = (uint32_t)(((uint64_t)(nmajor - 1) >> row_bits) + 1);
    Tmp<uint32_t> bptr(ctx, (size_t)nblocks + 1);
    const unsigned sgrid = (unsigned)ctx->num_sms * 16u;
    {   // cheap early look: histogram of every 64th record's block id
        Tmp<uint32_t> bcount(ctx, nblocks);
        SPL_CUDA(cudaMemsetAsync(bcount, 0, sizeof(uint32_t) * (size_t)nblocks, ctx->stream));
        SPL_CUDA(cudaMemsetAsync(ctx->d_scratch, 0, sizeof(uint32_t), ctx->stream));
        block_hist_kernel<K, LoadK><<<sgrid, 256, 0, ctx->stream>>>(lk, len, S, 64u, bcount);
        check_launch(ctx, "block_hist");
        unsigned g = div_up(nblocks, 256);
        max_seglen_counts_kernel<<<g < sgrid ? g : sgrid, 256, 0, ctx->stream>>>(bcount, nblocks, ctx->d_scratch);
        check_launch(ctx, "max_block");
        uint32_t m = 0;
        read_back(ctx, ctx->d_scratch, &m, 1);
        if ((uint64_t)m * 64 > (uint64_t)cap + cap / 2) return;                // clearly skewed
    }
    K *kb[2] = {k1, k0};
    VB *vb[2] = {v1, v0};
    NoPayload *nb[2] = {nullptr, nullptr};
    const int r = radix_sort<K, VB, NoPayload>(ctx, len, H, lk, lv, LoadNone{}, kb, vb, nb, S);
    K *ik = kb[r], *ok = kb[r ^ 1];
    VB *iv = vb[r], *ov = vb[r ^ 1];
    *sorted_k = ik;
    *sorted_v = iv;
    {   // exact block boundaries and the longest block, from the block-sorted keys
        GapQueueOwner gaps(ctx);
        block_bounds_kernel<K><<<div_up((uint64_t)len + 1, 256 * 4), 256, 0, ctx->stream>>>(ik, len, S, nblocks, bptr,
                                                                                         gaps.q);
        check_launch(ctx, "block_bounds");
        gaps.drain(ctx, bptr);
        SPL_CUDA(cudaMemsetAsync(ctx->d_scratch, 0, sizeof(uint32_t), ctx->stream));
        unsigned g = div_up(nblocks, 256);
        max_seglen_kernel<<<g < sgrid ? g : sgrid, 256, 0, ctx->stream>>>(bptr, nblocks, ctx->d_scratch);
        check_launch(ctx, "max_block");
        uint32_t longest = 0;
        read_back(ctx, ctx->d_scratch, &longest, 1);
        if (longest > cap) return;                 // the caller sorts (ik, iv) fully; the passes so far were stable
    }
    // one CTA per block: sort, in-order sum, zero drop, compaction inside the block's range of (ok, ov)
    Tmp<uint32_t> block_kept(ctx, nblocks), block_base(ctx, (size_t)nblocks + 1), local_ptr(ctx, nmajor);
    uint32_t *tmp_ind = reinterpret_cast<uint32_t *>(ok);
    auto finish = [&](auto lk_tag, auto t_tag, auto cap_tag) -> spl_mat * {
        using LK = decltype(lk_tag);
        using T = decltype(t_tag);
        constexpr int CAP = decltype(cap_tag)::value;
        static_assert(sizeof(T) == sizeof(VB), "value container and scalar must have one size");
        constexpr size_t kSmem = (size_t)CAP * (sizeof(LK) + sizeof(T) + 2);
        auto kern = block_finish_kernel<K, T, LK, CAP>;
        SPL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem));
        SPL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                      cudaSharedmemCarveoutMaxShared));
        kern<<<nblocks, CAP / 16, kSmem, ctx->stream>>>(ik, reinterpret_cast<const T *>(iv), bptr, minor_bits,
                                                         row_bits, dedup, dropzero, nmajor, tmp_ind,
                                                         reinterpret_cast<T *>(ov), block_kept, local_ptr);
        check_launch(ctx, "block_finish");
        exclusive_scan_u32(ctx, block_kept, nblocks, block_base);
        uint32_t nnz = 0;
        read_back(ctx, block_base.p + nblocks, &nnz, 1);      // the one host sync of the tail: exact-size output
        spl_mat *m = new_mat(ctx, format, dtype, nrows, ncols, nnz);
        block_gather_kernel<T><<<nblocks, 256, 0, ctx->stream>>>(tmp_ind, reinterpret_cast<const T *>(ov), bptr,
                                                               block_base, local_ptr, row_bits, nmajor, m->ind,
                                                               static_cast<T *>(m->val), m->ptr);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) {
            free_mat(ctx, m);
            throw Error{SPL_ERR_CUDA, std::string("block_gather: ") + cudaGetErrorString(e)};
        }
        count_launch(ctx);
        return m;
    };
    using Small = std::integral_constant<int, (int)BL_CAP>;
    using Big = std::integral_constant<int, (int)BL_CAP_BIG>;
    if (dtype == SPL_F32) {
        if constexpr (sizeof(VB) == 4)
            *result = cap == BL_CAP_BIG ? finish(uint32_t{}, float{}, Big{})
                      : S <= 32         ? finish(uint32_t{}, float{}, Small{})
                                        : finish(K{}, float{}, Small{});
    } else {
        if constexpr (sizeof(VB) == 8)
            *result = S <= 32 ? finish(uint32_t{}, double{}, Small{}) : finish(K{}, double{}, Small{});
    }
}

// General route, shared by the COO front (keys packed on the fly from the caller's index arrays)
// and the receive side of the sharded assembly (packed keys as routed): hybrid route first, full
// LSD radix passes and the streaming tail when it does not apply.
template <typename K, typename VB, typename LoadK>
spl_mat *assemble_impl(spl_ctx *ctx, int format, int dtype, uint32_t nrows, uint32_t ncols,
                       uint32_t len, LoadK lk, const VB *val, int minor_bits, int bits, int dedup,
                       int dropzero) {
    const uint32_t nmajor = format == SPL_CSR ? nrows : ncols;
    Tmp<K> k0(ctx, len), k1(ctx, len);
    Tmp<VB> v0(ctx, len), v1(ctx, len);
    const LoadPlain<VB> lv{val};
    auto finish = [&](K *keys, VB *vals) {
        if (dtype == SPL_F32)
            return finish_impl<K, float>(ctx, format, dtype, nrows, ncols, len, keys,
                                         reinterpret_cast<float *>(vals), minor_bits, dedup, dropzero);
        return finish_impl<K, double>(ctx, format, dtype, nrows, ncols, len, keys,
                                      reinterpret_cast<double *>(vals), minor_bits, dedup, dropzero);
    };
    K *src_k = nullptr;
    VB *src_v = nullptr;
#ifndef SPL_NO_HYBRID_SORT
    {
        spl_mat *fused = nullptr;      // the hybrid route ends in the finished matrix (tail fused into its last kernel)
        hybrid_sort<K, VB>(ctx, format, dtype, nrows, ncols, dedup, dropzero, len, nmajor, bits - minor_bits,
                           minor_bits, lk, lv, k0.p, v0.p, k1.p, v1.p, &src_k, &src_v, &fused);
        if (fused) return fused;
    }
#endif
    NoPayload *nb[2] = {nullptr, nullptr};
    if (src_k) {          // stable partial order left by the hybrid route: pass 0 writes the other pair
        K *kb[2] = {src_k == k0.p ? k1.p : k0.p, src_k};
        VB *vb[2] = {src_v == v0.p ? v1.p : v0.p, src_v};
        const int r = radix_sort<K, VB, NoPayload>(ctx, len, bits, LoadPlain<K>{src_k}, LoadPlain<VB>{src_v},
                                                   LoadNone{}, kb, vb, nb);
        return finish(kb[r], vb[r]);
    }
    K *kb[2] = {k0.p, k1.p};
    VB *vb[2] = {v0.p, v1.p};
    const int r = radix_sort<K, VB, NoPayload>(ctx, len, bits, lk, lv, LoadNone{}, kb, vb, nb);
    return finish(kb[r], vb[r]);
}

template <typename K, typename VB>
spl_mat *assemble_coo(spl_ctx *ctx, int format, int dtype, uint32_t nrows, uint32_t ncols, uint32_t len,
                      const uint32_t *major_idx, const uint32_t *minor_idx, const VB *val, int minor_bits,
                      int bits, int dedup, int dropzero) {
    const uint32_t nmajor = format == SPL_CSR ? nrows : ncols, nminor = format == SPL_CSR ? ncols : nrows;
    uint32_t flags[3] = {0, 0, 0};
    const unsigned sgrid = (unsigned)ctx->num_sms * 16u;
    if (len) {
        SPL_CUDA(cudaMemsetAsync(ctx->d_scratch, 0, 4 * sizeof(uint32_t), ctx->stream));
        const unsigned grid = div_up(len, 256 * 4);
        coo_scan_kernel<K><<<grid < sgrid ? grid : sgrid, 256, 0, ctx->stream>>>(major_idx, minor_idx, len, nmajor,
                                                                               nminor, minor_bits, ctx->d_scratch);
        check_launch(ctx, "coo_scan");
        read_back(ctx, ctx->d_scratch, flags, 3);
        SPL_REQUIRE(flags[0] == 0, SPL_ERR_ARG,
                    "COO entry out of bounds (CooMatrix::push asserts row < nrows, col < ncols)");
    }
    auto finish = [&](K *keys, VB *vals) {
        if (dtype == SPL_F32)
            return finish_impl<K, float>(ctx, format, dtype, nrows, ncols, len, keys,
                                         reinterpret_cast<float *>(vals), minor_bits, dedup, dropzero);
        return finish_impl<K, double>(ctx, format, dtype, nrows, ncols, len, keys,
                                      reinterpret_cast<double *>(vals), minor_bits, dedup, dropzero);
    };
    auto pack = [&](K *keys, VB *vals) {
        if (!len) return;
        const unsigned grid = div_up(len, 256 * 4);
        pack_kernel<K, VB><<<grid < sgrid ? grid : sgrid, 256, 0, ctx->stream>>>(major_idx, minor_idx, val, len,
                                                                               minor_bits, keys, vals);
        check_launch(ctx, "pack");
    };
    if (!flags[1]) {                      // already sorted: no sort at all
        Tmp<K> k0(ctx, len);
        Tmp<VB> v0(ctx, len);
        pack(k0, v0);
        return finish(k0, v0);
    }
    if (!flags[2]) {                      // majors already in order: sort inside the segments only
        Tmp<uint32_t> segptr(ctx, (size_t)nmajor + 1);
        fill_ptr(ctx, major_idx, len, nmajor, segptr);
        SPL_CUDA(cudaMemsetAsync(ctx->d_scratch, 0, sizeof(uint32_t), ctx->stream));
        unsigned g = div_up(nmajor, 256);
        max_seglen_kernel<<<g < sgrid ? g : sgrid, 256, 0, ctx->stream>>>(segptr, nmajor, ctx->d_scratch);
        check_launch(ctx, "max_seglen");
        uint32_t longest = 0;
        read_back(ctx, ctx->d_scratch, &longest, 1);
        if (longest <= 64) {
            Tmp<K> k0(ctx, len), k1(ctx, len);
            Tmp<VB> v0(ctx, len), v1(ctx, len);
            pack(k0, v0);
            if (longest <= 16)
                segment_sort_kernel<K, VB, 8><<<div_up((uint64_t)nmajor * 8, 256), 256, 0, ctx->stream>>>(
                    segptr, nmajor, k0, v0, k1, v1);
            else if (longest <= 32)
                segment_sort_kernel<K, VB, 16><<<div_up((uint64_t)nmajor * 16, 256), 256, 0, ctx->stream>>>(
                    segptr, nmajor, k0, v0, k1, v1);
            else
                segment_sort_kernel<K, VB, 32><<<div_up((uint64_t)nmajor * 32, 256), 256, 0, ctx->stream>>>(
                    segptr, nmajor, k0, v0, k1, v1);
            check_launch(ctx, "segment_sort");
            return finish(k1, v1);
        }
    }
    // general route: keys packed on the fly by the first pass, values read in place
    return assemble_impl<K, VB>(ctx, format, dtype, nrows, ncols, len, LoadPack<K>{major_idx, minor_idx, minor_bits},
                                val, minor_bits, bits, dedup, dropzero);
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
    GapQueueOwner gaps(ctx);
    fill_ptr_kernel<<<div_up((uint64_t)nnz + 1, 256), 256, 0, ctx->stream>>>(sorted_major, nnz,
                                                                            nmajor, ptr, gaps.q);
    check_launch(ctx, "fill_ptr");
    gaps.drain(ctx, ptr);
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
    const uint32_t *major_idx = format == SPL_CSR ? row : col;
    const uint32_t *minor_idx = format == SPL_CSR ? col : row;
    const int major_bits = bits_for(format == SPL_CSR ? nrows : ncols);
    const int minor_bits = bits_for(format == SPL_CSR ? ncols : nrows);
    const int bits = major_bits + minor_bits;
    if (bits <= 32) {
        if (dtype == SPL_F32)
            return assemble_coo<uint32_t, uint32_t>(ctx, format, dtype, nrows, ncols, len, major_idx, minor_idx,
                                                    (const uint32_t *)val, minor_bits, bits, dedup, dropzero);
        return assemble_coo<uint32_t, uint64_t>(ctx, format, dtype, nrows, ncols, len, major_idx, minor_idx,
                                                (const uint64_t *)val, minor_bits, bits, dedup, dropzero);
    }
    if (dtype == SPL_F32)
        return assemble_coo<uint64_t, uint32_t>(ctx, format, dtype, nrows, ncols, len, major_idx, minor_idx,
                                                (const uint32_t *)val, minor_bits, bits, dedup, dropzero);
    return assemble_coo<uint64_t, uint64_t>(ctx, format, dtype, nrows, ncols, len, major_idx, minor_idx,
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

// ---- fused routing + exchange over peer memory -----------------------------------------------
// route_count_dev: bounds and how many of this rank's triplets each rank owns (host needs them to
// lay out the receive buffers).  route_coo_peers_dev: the same stable partition as route_coo_dev,
// but the pass writes every record straight into its owner's receive buffer (CUDA IPC mapping, NVLink)
// at the slot the caller computed from everybody's counts: no staging buffer, no all-to-all.
static Owners make_owners(int format, uint32_t nrows, uint32_t ncols, int world, const uint64_t *major_starts) {
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
    return own;
}

void route_count_dev(spl_ctx *ctx, int format, uint32_t nrows, uint32_t ncols, uint32_t len,
                     const uint32_t *row, const uint32_t *col, int world, const uint64_t *major_starts,
                     uint64_t *counts_host) {
    const Owners own = make_owners(format, nrows, ncols, world, major_starts);
    for (int g = 0; g < world; ++g) counts_host[g] = 0;
    if (len == 0) return;
    uint32_t *cnt = ctx->d_scratch + 8;
    SPL_CUDA(cudaMemsetAsync(ctx->d_scratch, 0, sizeof(uint32_t), ctx->stream));
    SPL_CUDA(cudaMemsetAsync(cnt, 0, SPL_MAX_PEERS * sizeof(uint32_t), ctx->stream));
    unsigned grid = div_up(len, 256 * 8);
    if (grid > (unsigned)ctx->num_sms * 8u) grid = (unsigned)ctx->num_sms * 8u;
    coo_bounds_kernel<<<grid, 256, 0, ctx->stream>>>(row, col, len, nrows, ncols, ctx->d_scratch);
    check_launch(ctx, "coo_bounds");
    owner_count_kernel<<<grid, 256, 0, ctx->stream>>>(format == SPL_CSR ? row : col, len, own, cnt);
    check_launch(ctx, "owner_count");
    uint32_t w[16];
    read_back(ctx, ctx->d_scratch, w, 16);
    SPL_REQUIRE(w[0] == 0, SPL_ERR_ARG,
                "COO entry out of bounds (CooMatrix::push asserts row < nrows, col < ncols)");
    for (int g = 0; g < world; ++g) counts_host[g] = w[8 + g];
}

void route_coo_peers_dev(spl_ctx *ctx, int format, int dtype, uint32_t nrows, uint32_t ncols, uint32_t len,
                         const uint32_t *row, const uint32_t *col, const void *val, int world,
                         const uint64_t *major_starts, void *const *key_bufs, void *const *val_bufs,
                         const uint64_t *dst_offsets) {
    const Owners own = make_owners(format, nrows, ncols, world, major_starts);
    if (len == 0) return;
    const uint32_t *major = format == SPL_CSR ? row : col;
    const uint32_t *minor = format == SPL_CSR ? col : row;
    const int minor_bits = bits_for(format == SPL_CSR ? ncols : nrows);
    LoadOwner lo{major, own};
    LoadPackLocal lp{major, minor, own, minor_bits};
    auto run = [&](auto tag) {
        using VB = decltype(tag);
        DestPeers<uint64_t, VB> dest{};
        for (int g = 0; g < world; ++g) {
            SPL_REQUIRE(key_bufs[g] && val_bufs[g], SPL_ERR_ARG, "NULL receive buffer");
            SPL_REQUIRE(dst_offsets[g] < kMaxEntries, SPL_ERR_UNSUPPORTED, "receive offset beyond 2^32");
            dest.a[g] = static_cast<uint64_t *>(key_bufs[g]);
            dest.b[g] = static_cast<VB *>(val_bufs[g]);
            dest.off[g] = (uint32_t)dst_offsets[g];
        }
        radix_pass_to_peers<uint64_t, VB>(ctx, len, bits_for((uint64_t)world), lo, lp,
                                          LoadPlain<VB>{static_cast<const VB *>(val)}, dest);
    };
    if (dtype == SPL_F32) run(uint32_t{});
    else run(uint64_t{});
}

}  // namespace spl
