// gather.cu — row-sharded y = A x for GENERAL shards (random / unstructured columns), the
// all-gather of x fused into the product ("Collective (general)").
//
// A banded or stencil shard copies its few halo columns inside the barrier kernel (peer.cu).  A
// shard with random columns cannot: every remote gather would be a 4-byte NVLink transaction.  It
// needs the whole of x locally, and done as "all-gather, then SpMV" the exchange (35 MB per rank at
// 8 GPUs on config 3) costs as much as the product and nothing overlaps.  Here ONE persistent
// kernel does both:
//   * the shard is stored blocked by column owner: block 0 holds, row by row, the entries whose
//     column lives on this rank, the following blocks the peers' columns in ring order (rank+1,
//     rank+2, ...), one block per peer or per GROUP of consecutive peers (block_first);
//   * every CTA carries one COPY warp beside its 8 compute warps.  Its elected thread moves the
//     CTA's share of the peers' slices with TMA bulk copies: peer memory -> a small shared-memory
//     ring -> the local full-length x (cp.async.bulk + mbarrier, bulk groups for the stores), three
//     2 KB loads in flight per CTA = several MB in flight on the chip, which is what a ~3 us NVLink
//     round trip needs (profiles/r2_nvlink_copy.txt: 128-bit loads from 64 CTAs reach 400 GB/s
//     alone and 220 GB/s beside the gathers of the product, which saturate the same L1/LSU path;
//     TMA bypasses it and reaches the copy engines' rate with one thread per CTA).  When a CTA's
//     chunks of a slice have been written it bumps that slice's counter (release);
//   * the COMPUTE warps own a fixed range of rows per CTA, keep the row sums in shared memory, and
//     walk the blocks in order: block 0 needs only the rank's own slice and runs while the first
//     slices are in flight; before block b they wait (acquire) for the counters of its slices.
//     Transfer and math overlap block by block, and y is written once.
// All CTAs are resident at once (cooperative launch: the grid is what fits), so the waits cannot
// deadlock.  Row sums add the blocks in ring order, not in ascending column order: tolerance parity
// (1e-12 / 1e-5), like every multi-lane kernel.
#include <algorithm>
#include <cstdlib>

#include "kernels.cuh"

namespace spl {

namespace {

struct GatherPeers {
    const void *slice[SPL_MAX_PEERS];
    uint32_t start[SPL_MAX_PEERS + 1];
    int world, rank;
};

// The device-side barrier of spl_peer_barrier, folded into the kernel: flags[g] = rank g's flag block
// (peer memory), flags_mine = this rank's own block (NULL: the caller ran the barrier itself)
struct GatherBarrier {
    uint32_t *flags[SPL_MAX_PEERS];
    const uint32_t *flags_mine;
    uint32_t epoch;
    unsigned long long timeout_ns;
    uint32_t *failed;
};

__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

constexpr int GF_THREADS = 256;  // consumer threads of a CTA; one warp more feeds them tiles, another copies x
constexpr int GF_CTA = GF_THREADS + 64;
constexpr uint32_t GF_CHUNK = 2048;      // bytes per bulk copy of x
constexpr int GF_STAGES = 6;             // x ring slots per CTA
constexpr int GF_AHEAD = 4;              // loads run this many chunks in front of the stores (STAGES >= AHEAD + 2)
constexpr int GF_LAG = 3;                // stores whose completion is not waited for before the next chunk moves
constexpr uint32_t GF_RING = GF_CHUNK * GF_STAGES;

// block b of the shard = the columns owned by ring offsets [first[b], first[b+1]); first[0] = 0, first[1] = 1
struct GatherBlocks {
    uint32_t first[SPL_MAX_PEERS + 1];
    int n;
};

__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t gf_smem(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void gf_mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(gf_smem(bar)), "r"(count));
}
__device__ __forceinline__ void gf_mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred q;\nGFW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 q, [%0], %1;\n@q bra GFD_%=;\nbra GFW_%=;\nGFD_%=:\n}\n" ::"r"(
            gf_smem(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void gf_mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(gf_smem(bar)) : "memory");
}
__device__ __forceinline__ void gf_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(gf_smem(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void gf_g2s(void *dst_smem, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(gf_smem(dst_smem)),
                 "l"(src), "r"(bytes), "r"(gf_smem(bar))
                 : "memory");
}
__device__ __forceinline__ void gf_g2s_stream(void *dst_smem, const void *src, uint32_t bytes, uint64_t *bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            gf_smem(dst_smem)),
        "l"(src), "r"(bytes), "r"(gf_smem(bar)), "l"(policy)
        : "memory");
}

// ---- the copy warp of one CTA ----
template <typename T>
__device__ void gather_copy_warp(const GatherPeers &gp, T *x_full, uint32_t *ready, const GatherBarrier &gb,
                                 unsigned char *ring, uint64_t *full, unsigned long long *timeline) {
    const int G = gp.world;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned long long cta = blockIdx.x, ncta = gridDim.x;
    if (gb.flags_mine) {
        if (cta == 0 && (int)lane < G && (int)lane != gp.rank) {
            // fused barrier, arrival: this rank's published slice is final (stream order put its writers before us)
            __threadfence_system();
            asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(gb.flags[lane] + gp.rank), "r"(gb.epoch) : "memory");
        }
        if ((int)lane < G && (int)lane != gp.rank) {       // wait: every peer's slice is final once its epoch shows up here
            const unsigned long long t0 = global_timer_ns();
            for (;;) {
                uint32_t v;
                asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(gb.flags_mine + lane) : "memory");
                if ((int32_t)(v - gb.epoch) >= 0) break;
                if (global_timer_ns() - t0 > gb.timeout_ns) { atomicExch(gb.failed, 1u); break; }
                __nanosleep(100);
            }
        }
        __syncwarp();
    }
    if (timeline && cta == 0 && lane == 0) timeline[0] = global_timer_ns();
    // chunks of the bulk path per slice (ring offset k = 1..G-1): whole 16-byte units of slices whose two ends are
    // 16-byte aligned (IPC blocks are; slice starts of an f32 vector need not be); the rest goes element by element
    unsigned long long cum[SPL_MAX_PEERS + 1];             // cum[k] = chunks in ring offsets 1..k
    cum[0] = 0;
    for (int k = 1; k < G; ++k) {
        const int g = (gp.rank + k) % G;
        const unsigned char *src = static_cast<const unsigned char *>(gp.slice[g]);
        unsigned char *dst = reinterpret_cast<unsigned char *>(x_full + gp.start[g]);
        const unsigned long long n = gp.start[g + 1] - gp.start[g], bytes = n * sizeof(T);
        const bool wide = ((((unsigned long long)(uintptr_t)src) | ((unsigned long long)(uintptr_t)dst)) & 15ull) == 0;
        const unsigned long long b16 = wide ? bytes & ~15ull : 0ull;
        cum[k] = cum[k - 1] + (b16 + GF_CHUNK - 1) / GF_CHUNK;
        // elements outside the bulk path: this CTA's share, all lanes
        const unsigned long long e0 = b16 / sizeof(T), rest = n - e0;
        if (rest) {
            const T *se = reinterpret_cast<const T *>(src);
            T *de = reinterpret_cast<T *>(dst);
            const unsigned long long per = (rest + ncta - 1) / ncta, lo = e0 + cta * per,
                                     hi = lo + per < n ? lo + per : n;
            for (unsigned long long e = lo + lane; e < hi; e += 32) de[e] = __ldcg(se + e);
        }
    }
    __syncwarp();
    if (lane == 0) {
        for (int s = 0; s < GF_STAGES; ++s)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(gf_smem(full + s)), "r"(1u));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const unsigned long long total = cum[G - 1];
        const unsigned long long mine = total > cta ? (total - cta + ncta - 1) / ncta : 0;
        int kl = 1, ks = 1, ksig = 1;
        auto slice_of = [&](int k, const unsigned char *&src, unsigned char *&dst, unsigned long long &b16) {
            const int g = (gp.rank + k) % G;
            src = static_cast<const unsigned char *>(gp.slice[g]);
            dst = reinterpret_cast<unsigned char *>(x_full + gp.start[g]);
            b16 = ((unsigned long long)(gp.start[g + 1] - gp.start[g]) * sizeof(T)) & ~15ull;
        };
        for (unsigned long long i = 0; i < mine + GF_AHEAD; ++i) {
            if (i < mine) {
                const unsigned long long c = cta + i * ncta;
                while (c >= cum[kl]) ++kl;
                const unsigned char *src; unsigned char *dst; unsigned long long b16;
                slice_of(kl, src, dst, b16);
                const unsigned long long off = (c - cum[kl - 1]) * GF_CHUNK;
                const uint32_t len = (uint32_t)(b16 - off < GF_CHUNK ? b16 - off : GF_CHUNK);
                const int s = (int)(i % GF_STAGES);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(gf_smem(full + s)), "r"(len) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 gf_smem(ring + (size_t)s * GF_CHUNK)),
                             "l"(src + off), "r"(len), "r"(gf_smem(full + s))
                             : "memory");
            }
            if (i >= (unsigned long long)GF_AHEAD) {
                const unsigned long long j = i - GF_AHEAD, c = cta + j * ncta;
                while (c >= cum[ks]) ++ks;
                const unsigned char *src; unsigned char *dst; unsigned long long b16;
                slice_of(ks, src, dst, b16);
                const unsigned long long off = (c - cum[ks - 1]) * GF_CHUNK;
                const uint32_t len = (uint32_t)(b16 - off < GF_CHUNK ? b16 - off : GF_CHUNK);
                const int s = (int)(j % GF_STAGES);
                const uint32_t parity = (uint32_t)((j / GF_STAGES) & 1);
                asm volatile(
                    "{\n.reg .pred q;\nGFW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 q, [%0], %1;\n@q bra GFD_%=;\nbra GFW_%=;\nGFD_%=:\n}\n" ::"r"(
                        gf_smem(full + s)),
                    "r"(parity)
                    : "memory");
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + off),
                             "r"(gf_smem(ring + (size_t)s * GF_CHUNK)), "r"(len)
                             : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // every store but the newest has left its slot
                asm volatile("cp.async.bulk.wait_group %0;" ::"n"(GF_LAG) : "memory");   // all but the newest LAG are written
                // slices that lie wholly below this CTA's oldest unwritten chunk hold nothing of it that is still in flight
                if (j >= (unsigned long long)GF_LAG) {
                    const unsigned long long cdone = cta + (j - GF_LAG + 1) * ncta;
                    while (ksig < G && cum[ksig] <= cdone) {
                        __threadfence();
                        atomicAdd(ready + ksig, 1u);
                        ++ksig;
                    }
                }
            }
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        __threadfence();
        for (; ksig < G; ++ksig) atomicAdd(ready + ksig, 1u);
    }
}

// Shared memory of a CTA: [tile stages: indices | values | row pointers] [row sums] [x ring] [tile bounds] [barriers]
struct GatherShape {
    uint32_t rows_per_cta;    // multiple of 32
    uint32_t ntiles;          // tiles of R rows per block and CTA
    uint32_t cap;             // entries a tile stage holds (16-byte aligned superset of the largest tile)
    uint32_t stages;          // tile stages
    uint32_t off_val, off_ptr, off_sum, off_ring, off_tlo, off_bar;      // byte offsets
};

template <typename T, int TR>
__global__ void __launch_bounds__(GF_CTA, 3)
spmv_gather_fused_kernel(uint32_t nloc, const uint32_t *__restrict__ bptr, uint32_t pstride,
                         const uint32_t *__restrict__ bind, const T *__restrict__ bval, GatherPeers gp, GatherBlocks gk,
                         T *x_full, T *__restrict__ y, uint32_t *ready, uint32_t target, GatherShape sh, GatherBarrier gb,
                         unsigned long long *timeline) {
    constexpr uint32_t R = TR;                    // rows per tile (chosen so that a tile holds ~2048 entries)
    constexpr int U = sizeof(T) == 8 ? 4 : 8;     // gathers a lane issues per round (two rounds are in flight: registers)
    constexpr uint32_t PTRS = R + 4;              // pointer slots per stage: R + 1 needed, whole 16-byte units
    extern __shared__ __align__(128) unsigned char gf_raw[];
    uint32_t *s_ind = reinterpret_cast<uint32_t *>(gf_raw);
    T *s_val = reinterpret_cast<T *>(gf_raw + sh.off_val);
    uint32_t *s_ptr = reinterpret_cast<uint32_t *>(gf_raw + sh.off_ptr);
    T *s_sum = reinterpret_cast<T *>(gf_raw + sh.off_sum);
    uint32_t *s_tlo = reinterpret_cast<uint32_t *>(gf_raw + sh.off_tlo);
    uint64_t *full = reinterpret_cast<uint64_t *>(gf_raw + sh.off_bar);
    uint64_t *empty = full + sh.stages;
    uint64_t *xfull = empty + sh.stages;
    const uint32_t c = blockIdx.x;
    const uint64_t rs64 = (uint64_t)c * sh.rows_per_cta;
    const uint32_t rs = rs64 < nloc ? (uint32_t)rs64 : nloc;
    const uint32_t re = rs + sh.rows_per_cta < nloc ? rs + sh.rows_per_cta : nloc;
    const uint32_t ntiles = (re - rs + R - 1) / R;          // 0 for a CTA without rows (it still copies x)
    const uint32_t nb = (uint32_t)gk.n;

    if (threadIdx.x >= GF_THREADS + 32) {
        // ---- warp 9: this CTA's share of the peers' slices of x ----
        if (gp.world > 1) gather_copy_warp<T>(gp, x_full, ready, gb, gf_raw + sh.off_ring, xfull, timeline);
        return;
    }
    if (threadIdx.x >= GF_THREADS) {
        // ---- warp 8: feeds the consumers.  All lanes fetch the tile bounds of every block (one round trip), then one
        // thread issues the bulk copies of each tile's row pointers, indices and values, `stages` tiles ahead; the
        // matrix does not depend on x, so the tiles of block b are in shared memory before its slices have landed ----
        const uint32_t lane = threadIdx.x & 31u;
        if (ntiles == 0) return;
        for (uint32_t i = lane; i < nb * (ntiles + 1); i += 32) {
            const uint32_t k = i / (ntiles + 1), t = i % (ntiles + 1);
            const uint32_t r = rs + t * R < re ? rs + t * R : re;
            s_tlo[i] = __ldg(bptr + (size_t)k * pstride + r);
        }
        if (lane == 0) {
            for (uint32_t s = 0; s < sh.stages; ++s) {
                gf_mbar_init(full + s, 1);
                gf_mbar_init(empty + s, GF_THREADS / 32);
            }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("bar.arrive 2, %0;" ::"n"(GF_THREADS + 32) : "memory");      // barriers are live: consumers may wait on them
        if (lane != 0) return;
        uint64_t pol;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
        uint32_t s = 0, phase = 0, i = 0;
        for (uint32_t k = 0; k < nb; ++k)
            for (uint32_t t = 0; t < ntiles; ++t, ++i) {
                const uint32_t lo = s_tlo[k * (ntiles + 1) + t], hi = s_tlo[k * (ntiles + 1) + t + 1];
                if (i >= sh.stages) gf_mbar_wait(empty + s, phase ^ 1u);
                const uint32_t za = lo & ~3u, zb = (hi + 3u) & ~3u;           // 16-byte aligned superset; the arrays carry slack
                const uint32_t cnt = zb - za;
                const uint32_t r0 = rs + t * R;
                const uint32_t left = (pstride - r0) & ~3u;                   // pointer entries from r0 to the end of the block's array
                const uint32_t np = left < PTRS ? left : PTRS;
                gf_expect_tx(full + s, cnt * (uint32_t)(sizeof(uint32_t) + sizeof(T)) + np * (uint32_t)sizeof(uint32_t));
                gf_g2s(s_ptr + (size_t)s * PTRS, bptr + (size_t)k * pstride + r0, np * (uint32_t)sizeof(uint32_t), full + s);
                if (cnt) {
                    gf_g2s_stream(s_ind + (size_t)s * sh.cap, bind + za, cnt * (uint32_t)sizeof(uint32_t), full + s, pol);
                    gf_g2s_stream(s_val + (size_t)s * sh.cap, bval + za, cnt * (uint32_t)sizeof(T), full + s, pol);
                }
                if (++s == sh.stages) { s = 0; phase ^= 1u; }
            }
        return;
    }

    // ---- consumers.  A tile's rows are dealt to the 8 warps in contiguous runs (so are its entries), and every warp
    // works through its run on its own, in two steps (pointers, indices and values are in shared memory):
    //   1. the run's ENTRIES go to the lanes, 32 apart, eight per lane and round: gather x, multiply, put the product
    //      back in place of the value.  Every lane keeps eight gathers in flight whatever the rows look like;
    //   2. the run's ROWS go to the lanes: a row's products are summed in column order (one lane, sequentially) and
    //      added to the row's running sum, which the same lane owns in every block — no atomics, one fixed order.
    // The two steps are software-pipelined: the gathers of the NEXT tile's run are issued before the sums of the
    // current one, so the gather latency of a tile hides behind the shared-memory work of its predecessor.  No CTA-wide
    // barrier: the warps drift apart, which keeps the gathers of an SM flowing instead of arriving in bursts.
    // Nothing but x comes from global memory ----
    if (ntiles == 0) return;
    asm volatile("bar.sync 2, %0;" ::"n"(GF_THREADS + 32) : "memory");
    const T *own = static_cast<const T *>(gp.slice[gp.rank]) - gp.start[gp.rank];
    constexpr uint32_t WR = R / (GF_THREADS / 32);          // rows of a tile per warp
    const uint32_t lane = threadIdx.x & 31u, w0 = (threadIdx.x >> 5) * WR;
    const bool stamp = timeline && c == 0 && threadIdx.x == 0;

    struct Run { uint32_t s, t, k, r0, r1, j0, j1, za; };    // warp-uniform: stage, tile, block, rows and entries of the run
    // wait for the stage of tile (k, t), then issue the first round of its gathers
    auto issue = [&](uint32_t k, uint32_t t, uint32_t s, uint32_t phase, Run &run, T (&xv)[U]) {
        gf_mbar_wait(full + s, phase);
        const uint32_t *cp = s_ptr + (size_t)s * PTRS;
        const uint32_t *ci = s_ind + (size_t)s * sh.cap;
        const uint32_t rows = re - (rs + t * R) < R ? re - (rs + t * R) : R;
        run.s = s; run.t = t; run.k = k;
        run.r0 = w0 < rows ? w0 : rows;
        run.r1 = w0 + WR < rows ? w0 + WR : rows;
        run.za = cp[0] & ~3u;
        run.j0 = cp[run.r0] - run.za;
        run.j1 = cp[run.r1] - run.za;
        // block 0 reads the published slice (constant for the whole kernel); x_full is written by this very kernel,
        // behind the acquire that let this block start: L2 loads
        if (k == 0) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const uint32_t j = run.j0 + u * 32 + lane;
                xv[u] = j < run.j1 ? __ldg(own + ci[j]) : (T)0;
            }
        } else {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const uint32_t j = run.j0 + u * 32 + lane;
                xv[u] = j < run.j1 ? __ldcg(x_full + ci[j]) : (T)0;
            }
        }
    };
    // products of the first round in place, further rounds of a long run, then the row sums; releases the stage
    auto finish = [&](const Run &run, T (&xv)[U]) {
        const uint32_t *cp = s_ptr + (size_t)run.s * PTRS;
        const uint32_t *ci = s_ind + (size_t)run.s * sh.cap;
        T *cv = s_val + (size_t)run.s * sh.cap;
        const T *xb = run.k == 0 ? own : x_full;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t j = run.j0 + u * 32 + lane;
            if (j < run.j1) cv[j] *= xv[u];
        }
        for (uint32_t base = run.j0 + 32 * U; base < run.j1; base += 32 * U) {       // warp-uniform trip count
            T more[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const uint32_t j = base + u * 32 + lane;
                more[u] = j < run.j1 ? __ldcg(xb + ci[j]) : (T)0;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const uint32_t j = base + u * 32 + lane;
                if (j < run.j1) cv[j] *= more[u];
            }
        }
        __syncwarp();                                                // products of the run are in place
        for (uint32_t i = run.r0 + lane; i < run.r1; i += 32) {
            T acc = (T)0;
            for (uint32_t j = cp[i] - run.za, e = cp[i + 1] - run.za; j < e; ++j) acc += cv[j];
            T *slot = s_sum + (size_t)run.t * R + i;
            *slot = run.k == 0 ? acc : *slot + acc;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");            // our writes to the stage, before the next bulk copy into it
        __syncwarp();
        if (lane == 0) gf_mbar_arrive(empty + run.s);
    };
    // have the slices of block k landed?  (acquire; lane 0 looks, the warp learns)
    auto landed = [&](uint32_t k) {
        int ok = 1;
        if (lane == 0)
            for (uint32_t o = gk.first[k]; o < gk.first[k + 1]; ++o) ok &= (int32_t)(ld_acquire_gpu(ready + o) - target) >= 0;
        return __shfl_sync(0xffffffffu, ok, 0) != 0;
    };

    Run cur{}, nxt{};
    T xa[U], xn[U];
    uint32_t s = 0, phase = 0;
    if (stamp) timeline[1] = timeline[2] = global_timer_ns();
    issue(0, 0, s, phase, cur, xa);
    if (++s == sh.stages) { s = 0; phase ^= 1u; }
    for (uint32_t k = 0; k < nb; ++k)
        for (uint32_t t = 0; t < ntiles; ++t) {
            const bool last = t + 1 == ntiles;
            const uint32_t k2 = last ? k + 1 : k, t2 = last ? 0 : t + 1;
            bool have = false;
            if (k2 < nb) {
                if (last && stamp) timeline[1 + 3 * k2] = global_timer_ns();                      // block k2: wait begins
                if (!last || landed(k2)) {
                    if (last && stamp) timeline[2 + 3 * k2] = global_timer_ns();                  // its slices have landed
                    issue(k2, t2, s, phase, nxt, xn);
                    have = true;
                }
            }
            finish(cur, xa);
            if (last && stamp) timeline[3 + 3 * k] = global_timer_ns();                           // block k done (this warp)
            if (k2 < nb && !have) {
                while (!landed(k2)) __nanosleep(64);
                if (stamp) timeline[2 + 3 * k2] = global_timer_ns();
                issue(k2, t2, s, phase, nxt, xn);
            }
            if (k2 < nb) {
                if (++s == sh.stages) { s = 0; phase ^= 1u; }
                cur = nxt;
#pragma unroll
                for (int u = 0; u < U; ++u) xa[u] = xn[u];
            }
        }
    // a fused barrier that gave up: NaN instead of sums over a half-written x (the status call reports it)
    const bool poisoned = gb.failed && *reinterpret_cast<const volatile uint32_t *>(gb.failed) != 0u;
    for (uint32_t t = 0; t < ntiles; ++t) {                              // every lane writes the rows it summed
        const uint32_t rows = re - (rs + t * R) < R ? re - (rs + t * R) : R;
        const uint32_t r1 = w0 + WR < rows ? w0 + WR : rows;
        for (uint32_t i = w0 + lane; i < r1; i += 32) y[rs + t * R + i] = poisoned ? (T)NAN : s_sum[(size_t)t * R + i];
    }
    if (timeline && threadIdx.x == 0) atomicMax(timeline + 1 + 3 * nb, global_timer_ns());           // the last CTA to finish
}

}  // namespace

// Launch shape: as many CTAs as stay resident (cooperative launch); the rows of a CTA (its shared-memory sums) shrink
// as the CTAs per SM grow, so the two are found together.  tile_caps = the most entries any 64 / 128 / 256 / 512 / 1024
// consecutive rows (starting at a multiple of 32) hold in one block.
template <typename T, int TR>
void spmv_gather_fused_t(spl_ctx *ctx, uint32_t nloc, const uint32_t *bptr, uint32_t pstride, const uint32_t *bind,
                         const T *bval, const uint32_t *tile_caps, const GatherPeers &gp, const GatherBlocks &gk, T *x_full,
                         T *y, uint32_t *ready, uint32_t epoch, const GatherBarrier &gb, unsigned long long *timeline) {
    auto k = spmv_gather_fused_kernel<T, TR>;
    constexpr uint32_t R = TR;
    static_assert(R == 64 || R == 128 || R == 256 || R == 512 || R == 1024, "tile rows");
    const uint32_t cap = (tile_caps[R == 64 ? 0 : R == 128 ? 1 : R == 256 ? 2 : R == 512 ? 3 : 4] + 6u + 3u) & ~3u;
    const char *pc = std::getenv("SPL_GATHER_CTAS_PER_SM");          // measurement knobs
    const char *ps = std::getenv("SPL_GATHER_STAGES");
    const int most = pc ? std::max(1, std::atoi(pc)) : 3;
    uint32_t ncta = 0;
    size_t smem = 0;
    GatherShape sh{};
    auto shape = [&](uint32_t ctas, uint32_t stages) {
        sh.rows_per_cta = (uint32_t)(((uint64_t)nloc + ctas - 1) / ctas);
        sh.rows_per_cta = (sh.rows_per_cta + 31u) & ~31u;
        sh.ntiles = (sh.rows_per_cta + R - 1) / R;
        sh.cap = cap;
        sh.stages = stages;
        size_t o = (size_t)stages * cap * sizeof(uint32_t);
        sh.off_val = (uint32_t)o;
        o += (size_t)stages * cap * sizeof(T);
        sh.off_ptr = (uint32_t)o;
        o += (size_t)stages * (R + 4) * sizeof(uint32_t);
        o = (o + 15) & ~(size_t)15;
        sh.off_sum = (uint32_t)o;
        o += (size_t)sh.ntiles * R * sizeof(T);
        o = (o + 127) & ~(size_t)127;
        sh.off_ring = (uint32_t)o;
        o += GF_RING;
        sh.off_tlo = (uint32_t)o;
        o += (size_t)gk.n * (sh.ntiles + 1) * sizeof(uint32_t);
        o = (o + 7) & ~(size_t)7;
        sh.off_bar = (uint32_t)o;
        o += (size_t)(2 * stages + GF_STAGES) * sizeof(uint64_t);
        return o;
    };
    for (int want = most; want >= 1 && ncta == 0; --want)
        for (uint32_t stages = ps ? (uint32_t)std::max(2, std::atoi(ps)) : 3u; stages >= 2 && ncta == 0; --stages) {
            const uint32_t ctas = (uint32_t)ctx->num_sms * (uint32_t)want;
            smem = shape(ctas, stages);
            if (smem > 226 * 1024 / (size_t)want) continue;
            if (smem > 48 * 1024) SPL_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            int resident = 0;
            SPL_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, k, GF_CTA, smem));
            if (resident >= want) ncta = ctas;
        }
    SPL_REQUIRE(ncta > 0, SPL_ERR_UNSUPPORTED,
                "fused gather SpMV: a tile of the shard (or its row sums) does not fit in shared memory");
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(ncta);
    cfg.blockDim = dim3(GF_CTA);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative;           // all CTAs resident or the launch fails: the waits cannot hang
    at[0].val.cooperative = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    SPL_CUDA(cudaLaunchKernelEx(&cfg, k, nloc, bptr, pstride, bind, bval, gp, gk, x_full, y, ready, epoch * ncta, sh, gb,
                                timeline));
    check_launch(ctx, "spmv_gather_fused");
}

void spmv_gather_fused(spl_ctx *ctx, int dtype, uint32_t nloc, int world, int rank, const uint64_t *col_starts,
                       const void *const *x_slices, int nblocks, const uint32_t *block_first, const uint32_t *bptr,
                       uint32_t pstride, const uint32_t *tile_caps, const uint32_t *bind, const void *bval, void *x_full,
                       void *y, uint32_t *ready, uint32_t epoch, double entries_per_row, void *const *flag_ptrs,
                       uint32_t barrier_epoch, uint32_t timeout_ms, unsigned long long *timeline) {
    GatherBarrier gb{};
    if (flag_ptrs && world > 1) {
        for (int g = 0; g < world; ++g) gb.flags[g] = static_cast<uint32_t *>(flag_ptrs[g]);
        gb.flags_mine = gb.flags[rank];
        gb.epoch = barrier_epoch;
        gb.timeout_ns = (unsigned long long)(timeout_ms ? timeout_ms : 2000) * 1000000ull;
    }
    gb.failed = ctx->d_scratch + 32;
    GatherPeers gp{};
    gp.world = world;
    gp.rank = rank;
    for (int g = 0; g <= SPL_MAX_PEERS; ++g) gp.start[g] = (uint32_t)col_starts[g < world ? g : world];
    for (int g = 0; g < world; ++g) gp.slice[g] = x_slices[g];
    GatherBlocks gk{};
    uint32_t widest = 1;
    if (block_first) {
        SPL_REQUIRE(nblocks >= 1 && nblocks <= world && block_first[0] == 0 && block_first[nblocks] == (uint32_t)world &&
                        (world == 1 || block_first[1] == 1),
                    SPL_ERR_ARG, "block_first must run 0, 1, ..., world (block 0 = the own slice alone)");
        gk.n = nblocks;
        for (int b = 0; b <= nblocks; ++b) gk.first[b] = block_first[b];
        for (int b = 0; b < nblocks; ++b) {
            SPL_REQUIRE(block_first[b] < block_first[b + 1], SPL_ERR_ARG, "block_first must increase");
            widest = std::max(widest, block_first[b + 1] - block_first[b]);
        }
    } else {
        gk.n = world;
        for (int b = 0; b <= world; ++b) gk.first[b] = (uint32_t)b;
    }
    // entries a row holds in the WIDEST block
    const double e = entries_per_row * (double)widest / (double)world;
    auto go = [&](auto tag, auto rows) {
        using T = decltype(tag);
        spmv_gather_fused_t<T, decltype(rows)::value>(ctx, nloc, bptr, pstride, bind, (const T *)bval, tile_caps, gp, gk,
                                                      (T *)x_full, (T *)y, ready, epoch, gb, timeline);
    };
    // rows per tile so that a tile holds about 2048 entries: one round of eight gathers per consumer
    const char *pl = std::getenv("SPL_GATHER_TILE_ROWS");            // measurement knob
    const int rows = pl ? std::atoi(pl) : (e <= 3.0 ? 1024 : e <= 6.0 ? 512 : e <= 12.0 ? 256 : e <= 24.0 ? 128 : 64);
    auto pick = [&](auto tag) {
        if (rows >= 1024) go(tag, std::integral_constant<int, 1024>{});
        else if (rows >= 512) go(tag, std::integral_constant<int, 512>{});
        else if (rows >= 256) go(tag, std::integral_constant<int, 256>{});
        else if (rows >= 128) go(tag, std::integral_constant<int, 128>{});
        else go(tag, std::integral_constant<int, 64>{});
    };
    if (dtype == SPL_F32) pick(float{});
    else pick(double{});
}

}  // namespace spl
