// radix_sort.cuh — stable LSD radix sort of (key, payload A[, payload B]) for sm_100a.
//
// Stability is what makes COO assembly bit-exact (duplicates must be summed in insertion
// order, src/csr/conv/coo.rs:36-57) and what makes the CSR<->CSC scatter deterministic
// (inner indices ascending, src/csr.rs:385-396).  Per digit pass:
//   rs_hist     each block histograms its contiguous chunk of tiles        (reads keys)
//   rs_scan     one block per digit: exclusive prefix over the blocks + digit totals   (tiny)
//   rs_scatter  each block re-walks its chunk tile by tile: warp-striped coalesced loads,
//               match.any ranking per warp, digit scan in shared memory, exchange through
//               shared memory so runs of equal digits leave as contiguous coalesced stores.
// Only the significant key bits are sorted (ceil(bits/8) passes, bits split evenly).
// The grid is a multiple of the SM count; every block owns whole tiles.
#pragma once

#include "common.cuh"

namespace spl {

#ifndef RS_THREADS_VALUE
#define RS_THREADS_VALUE 256
#endif
constexpr int RS_THREADS = RS_THREADS_VALUE;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_BINS = 256;
#ifndef RS_IPT_VALUE
#define RS_IPT_VALUE 16
#endif
constexpr int RS_IPT = RS_IPT_VALUE;             // items per thread
constexpr int RS_TILE = RS_THREADS * RS_IPT;     // 3072 items per tile
#ifndef RS_MIN_BLOCKS
#define RS_MIN_BLOCKS 2                          // resident downsweep CTAs per SM the compiler must allow
#endif

// ---- loaders: what pass 0 reads (later passes read the ping-pong buffers) ------------------
// Every loader takes a per-thread `state` word (0 at kernel start) it may use as a cursor.
template <typename V>
struct LoadPlain {
    const V *p;
    __device__ __forceinline__ V operator()(uint32_t i, uint32_t &) const { return p[i]; }
};
// key = hi[i] << lobits | lo[i]   ((row, col) for CSR assembly, (col, row) for CSC)
template <typename K>
struct LoadPack {
    const uint32_t *hi;
    const uint32_t *lo;
    int lobits;
    __device__ __forceinline__ K operator()(uint32_t i, uint32_t &) const {
        return (K)(((uint64_t)hi[i] << lobits) | (uint64_t)lo[i]);
    }
};
// major index of compressed entry i: the segment of ptr[] that contains i.  A thread asks for
// ascending i, so the search gallops forward from the previous answer (state) instead of
// bisecting the whole pointer array: 1-3 probes for neighbouring rows instead of log2(nmajor).
struct LoadMajor {
    const uint32_t *ptr;
    uint32_t nmajor;
    __device__ __forceinline__ uint32_t operator()(uint32_t i, uint32_t &state) const {
        uint32_t lo = state;                       // ptr[lo] <= i holds (ptr[0] == 0)
        if (__ldg(ptr + lo) > i) lo = 0;           // not ascending after all: restart
        uint32_t step = 1, hi = lo + 1;
        while (hi <= nmajor && __ldg(ptr + hi) <= i) {   // gallop: ptr[hi] <= i
            lo = hi;
            step <<= 1;
            hi = lo + step;
        }
        if (hi > nmajor + 1u) hi = nmajor + 1u;
        // answer is the last index in [lo, hi) with ptr[idx] <= i
        const uint32_t r = upper_bound_u32(ptr, lo + 1u, hi, i) - 1u;
        state = r;
        return r;
    }
};
struct LoadNone {
    __device__ __forceinline__ NoPayload operator()(uint32_t, uint32_t &) const { return NoPayload{}; }
};

inline int rs_num_passes(int bits) { return bits <= 0 ? 1 : (bits + 7) / 8; }

// ---- upsweep ---------------------------------------------------------------------------------
// Same block <-> chunk mapping as the downsweep, but 512 threads and 8 keys in flight per thread:
// the pass only reads keys, so it wants bytes in flight, not registers.
constexpr int RH_THREADS = 512;
constexpr int RH_WARPS = RH_THREADS / 32;

template <typename K, typename LoadK>
__global__ void __launch_bounds__(RH_THREADS)
rs_hist_kernel(LoadK lk, uint32_t n, uint32_t tiles_per_block, int shift, uint32_t mask,
               uint32_t *__restrict__ counts) {
    __shared__ uint32_t hist[RH_WARPS][RS_BINS];
    for (int j = threadIdx.x; j < RH_WARPS * RS_BINS; j += RH_THREADS) (&hist[0][0])[j] = 0;
    __syncthreads();
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
    const uint64_t begin = (uint64_t)blockIdx.x * tiles_per_block * RS_TILE;
    uint64_t end = begin + (uint64_t)tiles_per_block * RS_TILE;
    if (end > n) end = n;
    constexpr int U = 8;
    uint32_t state = 0;
    for (uint64_t blk = begin; blk < end; blk += (uint64_t)RH_THREADS * U) {
        uint32_t d[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            uint64_t i = blk + (uint64_t)u * RH_THREADS + threadIdx.x;
            d[u] = i < end ? ((uint32_t)(lk((uint32_t)i, state) >> shift) & mask) : 0xffffffffu;
        }
        // shared-memory atomics on the warp's private histogram: measured 1 437 Gkeys/s on B200
        // against 149 Gkeys/s for match.any aggregation (profiles/micro/match_bench.cu)
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (d[u] != 0xffffffffu) atomicAdd(&hist[warp][d[u]], 1u);
    }
    __syncthreads();
    if (threadIdx.x < RS_BINS) {
        uint32_t c = 0;
#pragma unroll
        for (int w = 0; w < RH_WARPS; ++w) c += hist[w][threadIdx.x];
        counts[(size_t)threadIdx.x * gridDim.x + blockIdx.x] = c;
    }
}

// ---- spine: one block per digit turns counts[digit][0..grid) into exclusive prefixes over the
// blocks and records the digit total; the downsweep scans the 256 totals itself ------------------
static __global__ void __launch_bounds__(512) rs_scan_kernel(uint32_t *counts, uint32_t grid,
                                                             uint32_t *__restrict__ totals) {
    __shared__ uint32_t ws[17];
    uint32_t *row = counts + (size_t)blockIdx.x * grid;
    const uint32_t per = (grid + blockDim.x - 1) / blockDim.x;
    const uint32_t lo = min(threadIdx.x * per, grid), hi = min(lo + per, grid);
    uint32_t s = 0;
    for (uint32_t i = lo; i < hi; ++i) s += row[i];
    uint32_t total;
    uint32_t run = block_exclusive_scan(s, ws, &total);
    for (uint32_t i = lo; i < hi; ++i) {
        uint32_t v = row[i];
        row[i] = run;
        run += v;
    }
    if (threadIdx.x == 0) totals[blockIdx.x] = total;
}

// ---- where a pass writes ---------------------------------------------------------------------
// DestSingle: the usual ping-pong buffers, digits laid out one after the other.
// DestPeers:  the digit is the rank that owns the record (row-sharded assembly): digit d
//             goes to rank d's receive buffers — peer memory mapped over NVLink — starting at off[d],
//             the slot this rank's share begins at there.  The keys (owner ids) are not written.
template <typename K, typename A, typename B>
struct DestSingle {
    K *k;
    A *a;
    B *b;
    static constexpr bool kSingle = true;
    __device__ __forceinline__ uint32_t base(uint32_t, uint32_t below) const { return below; }
};
template <typename A, typename B>
struct DestPeers {
    A *a[SPL_MAX_PEERS];
    B *b[SPL_MAX_PEERS];
    uint32_t off[SPL_MAX_PEERS];
    static constexpr bool kSingle = false;
    __device__ __forceinline__ uint32_t base(uint32_t d, uint32_t) const {
        return d < (uint32_t)SPL_MAX_PEERS ? off[d] : 0u;
    }
};

// ---- downsweep -------------------------------------------------------------------------------
template <typename K, typename A, typename B, typename LoadK, typename LoadA, typename LoadB,
          typename Dest = DestSingle<K, A, B>>
__global__ void __launch_bounds__(RS_THREADS, RS_MIN_BLOCKS)
rs_scatter_kernel(LoadK lk, LoadA la, LoadB lb, uint32_t n, uint32_t tiles_per_block, int shift,
                  uint32_t mask, const uint32_t *__restrict__ offsets,
                  const uint32_t *__restrict__ totals, const Dest dest) {
    constexpr bool kHasA = !std::is_same<A, NoPayload>::value;
    constexpr bool kHasB = !std::is_same<B, NoPayload>::value;
    extern __shared__ __align__(16) unsigned char exch_raw[];   // RS_TILE * 8 bytes (dynamic: > 48 KB total)
    __shared__ uint32_t whist[RS_WARPS][RS_BINS];
    __shared__ uint32_t wmask[RS_WARPS][RS_BINS];   // lanes per digit of the item being ranked (zero between items)
    __shared__ uint32_t glob[RS_BINS];      // global position of local position 0 of each digit
    __shared__ uint32_t running[RS_BINS];   // next free global slot per digit for this block
    __shared__ uint32_t ws[RS_WARPS + 1];

    const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
    {   // global start of this block's share of each digit = digits below + earlier blocks
        const uint32_t tot = threadIdx.x < RS_BINS ? totals[threadIdx.x] : 0u;
        const uint32_t below = block_exclusive_scan(tot, ws, nullptr);
        if (threadIdx.x < RS_BINS)
            running[threadIdx.x] = dest.base(threadIdx.x, below) + offsets[(size_t)threadIdx.x * gridDim.x + blockIdx.x];
    }
    for (int j = threadIdx.x; j < RS_WARPS * RS_BINS; j += RS_THREADS) (&wmask[0][0])[j] = 0;
    uint32_t st_k = 0, st_a = 0, st_b = 0;

    const uint64_t begin = (uint64_t)blockIdx.x * tiles_per_block * RS_TILE;
    for (uint32_t t = 0; t < tiles_per_block; ++t) {
        const uint64_t tile_base = begin + (uint64_t)t * RS_TILE;
        if (tile_base >= n) break;
        const uint32_t tile_count = (uint32_t)(((uint64_t)n - tile_base) < RS_TILE
                                                   ? ((uint64_t)n - tile_base) : RS_TILE);
        const uint32_t my_base = warp * 32 * RS_IPT + lane;   // + i*32 : warp-striped

        for (int j = threadIdx.x; j < RS_WARPS * RS_BINS; j += RS_THREADS) (&whist[0][0])[j] = 0;

        K keys[RS_IPT];
#pragma unroll
        for (int i = 0; i < RS_IPT; ++i) {
            uint32_t loc = my_base + i * 32;
            keys[i] = loc < tile_count ? lk((uint32_t)(tile_base + loc), st_k) : (K)0;
        }
        // payload A is requested now so that its DRAM latency hides behind the ranking
        A pa[kHasA ? RS_IPT : 1];
        if constexpr (kHasA) {
#pragma unroll
            for (int i = 0; i < RS_IPT; ++i) {
                const uint32_t loc = my_base + i * 32;
                if (loc < tile_count) pa[i] = la((uint32_t)(tile_base + loc), st_a);
            }
        }
        __syncthreads();

        // rank inside the warp, digit by digit occurrence order (stable)
        uint32_t rank[RS_IPT];
#pragma unroll
        for (int i = 0; i < RS_IPT; ++i) {
            const bool valid = my_base + i * 32 < tile_count;
            const uint32_t d = valid ? ((uint32_t)(keys[i] >> shift) & mask) : 0u;
            // lanes holding the same digit: every lane ORs its bit into the digit's word of a per-warp
            // shared-memory table, then reads the word back (measured on B200, profiles/micro/
            // match_bench.cu: 507 Gkeys/s, against 288 for eight ballots over the digit bits and 149 for
            // the match.any instruction; OR is commutative, so the mask — and with it the stable rank —
            // does not depend on the order the hardware applies the atomics in).  A/B on one box: CSR->CSC
            // -5.5 % (config 2) / -7 % (config 3); 64-bit-key assembly +2 %, so those keep the ballots.
            unsigned peers;
            if constexpr (sizeof(K) == 4) {
                if (valid) atomicOr(&wmask[warp][d], 1u << lane);
                __syncwarp();
                peers = valid ? wmask[warp][d] : 0u;
                __syncwarp();
            } else {        // 64-bit keys sit at the register cap: the ballot form measured 2 % faster there
                peers = __ballot_sync(0xffffffffu, valid);
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    const bool bit = (d >> b) & 1u;
                    const unsigned bal = __ballot_sync(0xffffffffu, bit);
                    peers &= bit ? bal : ~bal;
                }
            }
            const int leader = valid ? __ffs(peers) - 1 : (int)lane;
            uint32_t old = 0;
            if (valid && (int)lane == leader) {
                if constexpr (sizeof(K) == 4) wmask[warp][d] = 0;     // ready for the next item
                old = whist[warp][d];
                whist[warp][d] = old + __popc(peers);
            }
            old = __shfl_sync(0xffffffffu, old, leader);
            rank[i] = old + __popc(peers & lanemask_lt());
            __syncwarp();
        }
        __syncthreads();

        // per digit: exclusive prefix over warps, then over digits
        uint32_t cnt = 0;
        if (threadIdx.x < RS_BINS) {
#pragma unroll
            for (int w = 0; w < RS_WARPS; ++w) {
                uint32_t c = whist[w][threadIdx.x];
                whist[w][threadIdx.x] = cnt;
                cnt += c;
            }
        }
        const uint32_t dbase = block_exclusive_scan(cnt, ws, nullptr);
        if (threadIdx.x < RS_BINS) {
#pragma unroll
            for (int w = 0; w < RS_WARPS; ++w) whist[w][threadIdx.x] += dbase;
            glob[threadIdx.x] = running[threadIdx.x] - dbase;
            running[threadIdx.x] += cnt;
        }
        __syncthreads();

        // keys: to shared memory at their tile-local sorted position, then out coalesced
        K *exk = reinterpret_cast<K *>(exch_raw);
#pragma unroll
        for (int i = 0; i < RS_IPT; ++i) {
            if (my_base + i * 32 < tile_count) {
                const uint32_t d = (uint32_t)(keys[i] >> shift) & mask;
                rank[i] += whist[warp][d];
                exk[rank[i]] = keys[i];
            }
        }
        // payload B is requested before the keys leave: its latency hides behind the two exchanges
        B pb[kHasB ? RS_IPT : 1];
        if constexpr (kHasB) {
#pragma unroll
            for (int i = 0; i < RS_IPT; ++i) {
                const uint32_t loc = my_base + i * 32;
                if (loc < tile_count) pb[i] = lb((uint32_t)(tile_base + loc), st_b);
            }
        }
        __syncthreads();
        uint32_t gpos[RS_IPT];
        uint32_t dgt[Dest::kSingle ? 1 : RS_IPT];      // DestPeers: which rank's buffers slot j goes to
#pragma unroll
        for (int j = 0; j < RS_IPT; ++j) {
            const uint32_t p = j * RS_THREADS + threadIdx.x;
            if (p < tile_count) {
                const K k = exk[p];
                const uint32_t dg = (uint32_t)(k >> shift) & mask;
                gpos[j] = glob[dg] + p;
                if constexpr (Dest::kSingle) dest.k[gpos[j]] = k;
                else dgt[j] = dg;
            }
        }
        if constexpr (kHasA) {
            __syncthreads();
            A *exa = reinterpret_cast<A *>(exch_raw);
#pragma unroll
            for (int i = 0; i < RS_IPT; ++i) {
                const uint32_t loc = my_base + i * 32;
                if (loc < tile_count) exa[rank[i]] = pa[i];
            }
            __syncthreads();
#pragma unroll
            for (int j = 0; j < RS_IPT; ++j) {
                const uint32_t p = j * RS_THREADS + threadIdx.x;
                if (p < tile_count) {
                    if constexpr (Dest::kSingle) dest.a[gpos[j]] = exa[p];
                    else dest.a[dgt[j]][gpos[j]] = exa[p];
                }
            }
        }
        if constexpr (kHasB) {
            __syncthreads();
            B *exb = reinterpret_cast<B *>(exch_raw);
#pragma unroll
            for (int i = 0; i < RS_IPT; ++i) {
                const uint32_t loc = my_base + i * 32;
                if (loc < tile_count) exb[rank[i]] = pb[i];
            }
            __syncthreads();
#pragma unroll
            for (int j = 0; j < RS_IPT; ++j) {
                const uint32_t p = j * RS_THREADS + threadIdx.x;
                if (p < tile_count) {
                    if constexpr (Dest::kSingle) dest.b[gpos[j]] = exb[p];
                    else dest.b[dgt[j]][gpos[j]] = exb[p];
                }
            }
        }
        __syncthreads();
    }
}

// ---- driver ----------------------------------------------------------------------------------
// Sorts n records by `bits` bits of the key starting at bit `first_bit` (default: the low bits).  Pass 0 reads through the loaders and
// writes buffer set 0; pass p writes set p&1.  Returns the index (0/1) of the set that holds
// the sorted result, i.e. (passes-1)&1.  A/B = NoPayload (with LoadNone, null buffers) drops
// that payload.
template <typename K, typename A, typename B, typename LoadK, typename LoadA, typename LoadB>
int radix_sort(spl_ctx *ctx, uint32_t n, int bits, LoadK lk0, LoadA la0, LoadB lb0, K *k_buf[2],
               A *a_buf[2], B *b_buf[2], int first_bit = 0) {
    const int passes = rs_num_passes(bits);
    if (n == 0) return (passes - 1) & 1;
    if (bits <= 0) bits = 1;
    const uint32_t tiles = div_up(n, RS_TILE);
    // one wave: grid = SMs x resident downsweep CTAs per SM (register-limited; asked, not guessed)
    int occ = 0;
    SPL_CUDA(cudaFuncSetAttribute(rs_scatter_kernel<K, A, B, LoadK, LoadA, LoadB>,
                                  cudaFuncAttributePreferredSharedMemoryCarveout,
                                  cudaSharedmemCarveoutMaxShared));
    constexpr size_t kExch = (size_t)RS_TILE * 8;
    SPL_CUDA(cudaFuncSetAttribute(rs_scatter_kernel<K, A, B, LoadK, LoadA, LoadB>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kExch));
    SPL_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
        &occ, rs_scatter_kernel<K, A, B, LoadK, LoadA, LoadB>, RS_THREADS, kExch));
    if (occ < 1) occ = 1;
    uint32_t grid = (uint32_t)ctx->num_sms * (uint32_t)occ;
    if (grid > tiles) grid = tiles;
    const uint32_t tiles_per_block = div_up(tiles, grid);
    grid = div_up(tiles, tiles_per_block);
    Tmp<uint32_t> counts(ctx, (size_t)RS_BINS * grid);
    Tmp<uint32_t> totals(ctx, RS_BINS);

    int shift = first_bit, done = 0;       // sorts key bits [first_bit, first_bit + bits)
    for (int p = 0; p < passes; ++p) {
        const int nb = (bits - done + (passes - p) - 1) / (passes - p);   // even split
        const uint32_t mask = (1u << nb) - 1u;
        K *ok = k_buf[p & 1];
        A *oa = a_buf[p & 1];
        B *ob = b_buf[p & 1];
        if (p == 0) {
            rs_hist_kernel<K, LoadK><<<grid, RH_THREADS, 0, ctx->stream>>>(
                lk0, n, tiles_per_block, shift, mask, counts);
            check_launch(ctx, "rs_hist");
            rs_scan_kernel<<<RS_BINS, 512, 0, ctx->stream>>>(counts, grid, totals);
            check_launch(ctx, "rs_scan");
            rs_scatter_kernel<K, A, B, LoadK, LoadA, LoadB><<<grid, RS_THREADS, kExch, ctx->stream>>>(
                lk0, la0, lb0, n, tiles_per_block, shift, mask, counts, totals, DestSingle<K, A, B>{ok, oa, ob});
            check_launch(ctx, "rs_scatter");
        } else {
            using PA = typename std::conditional<std::is_same<A, NoPayload>::value, LoadNone,
                                                 LoadPlain<A>>::type;
            using PB = typename std::conditional<std::is_same<B, NoPayload>::value, LoadNone,
                                                 LoadPlain<B>>::type;
            LoadPlain<K> lk{k_buf[(p - 1) & 1]};
            PA la;
            PB lb;
            if constexpr (!std::is_same<A, NoPayload>::value) la.p = a_buf[(p - 1) & 1];
            if constexpr (!std::is_same<B, NoPayload>::value) lb.p = b_buf[(p - 1) & 1];
            rs_hist_kernel<K, LoadPlain<K>><<<grid, RH_THREADS, 0, ctx->stream>>>(
                lk, n, tiles_per_block, shift, mask, counts);
            check_launch(ctx, "rs_hist");
            rs_scan_kernel<<<RS_BINS, 512, 0, ctx->stream>>>(counts, grid, totals);
            check_launch(ctx, "rs_scan");
            SPL_CUDA(cudaFuncSetAttribute(rs_scatter_kernel<K, A, B, LoadPlain<K>, PA, PB>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kExch));
            SPL_CUDA(cudaFuncSetAttribute(rs_scatter_kernel<K, A, B, LoadPlain<K>, PA, PB>,
                                          cudaFuncAttributePreferredSharedMemoryCarveout,
                                          cudaSharedmemCarveoutMaxShared));
            rs_scatter_kernel<K, A, B, LoadPlain<K>, PA, PB><<<grid, RS_THREADS, kExch, ctx->stream>>>(
                lk, la, lb, n, tiles_per_block, shift, mask, counts, totals, DestSingle<K, A, B>{ok, oa, ob});
            check_launch(ctx, "rs_scatter");
        }
        shift += nb;
        done += nb;
    }
    return (passes - 1) & 1;
}

// One stable pass over a digit of at most 3 bits (the owning rank) whose output goes straight to the
// owners' receive buffers.  `bits` = bits of the owner id.
template <typename A, typename B, typename LoadK, typename LoadA, typename LoadB>
void radix_pass_to_peers(spl_ctx *ctx, uint32_t n, int bits, LoadK lk, LoadA la, LoadB lb,
                         const DestPeers<A, B> &dest) {
    if (n == 0) return;
    if (bits <= 0) bits = 1;
    using Kern = DestPeers<A, B>;
    auto kern = rs_scatter_kernel<uint32_t, A, B, LoadK, LoadA, LoadB, Kern>;
    constexpr size_t kExch = (size_t)RS_TILE * 8;
    const uint32_t tiles = div_up(n, RS_TILE);
    int occ = 0;
    SPL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    SPL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kExch));
    SPL_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, RS_THREADS, kExch));
    if (occ < 1) occ = 1;
    uint32_t grid = (uint32_t)ctx->num_sms * (uint32_t)occ;
    if (grid > tiles) grid = tiles;
    const uint32_t tiles_per_block = div_up(tiles, grid);
    grid = div_up(tiles, tiles_per_block);
    Tmp<uint32_t> counts(ctx, (size_t)RS_BINS * grid);
    Tmp<uint32_t> totals(ctx, RS_BINS);
    const uint32_t mask = (1u << bits) - 1u;
    rs_hist_kernel<uint32_t, LoadK><<<grid, RH_THREADS, 0, ctx->stream>>>(lk, n, tiles_per_block, 0, mask, counts);
    check_launch(ctx, "rs_hist");
    rs_scan_kernel<<<RS_BINS, 512, 0, ctx->stream>>>(counts, grid, totals);
    check_launch(ctx, "rs_scan");
    kern<<<grid, RS_THREADS, kExch, ctx->stream>>>(lk, la, lb, n, tiles_per_block, 0, mask, counts, totals, dest);
    check_launch(ctx, "rs_scatter_to_peers");
}

}  // namespace spl
