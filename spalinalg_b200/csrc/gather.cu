// gather.cu — row-sharded y = A x for GENERAL shards (random / unstructured columns), the
// all-gather of x fused into the product ("Collective (general)").
//
// A banded or stencil shard gathers its few halo columns straight from the owners' memory inside
// the SpMV kernel (spmv.cu, XPeer).  A shard with random columns cannot: every remote gather would
// be a 4-byte NVLink transaction.  It needs the whole of x locally, and done as "all-gather, then
// SpMV" the exchange (35 MB per rank at 8 GPUs on config 3) costs more than the product and nothing
// overlaps.  Here ONE persistent kernel does both:
//   * the shard is stored blocked by column owner, own block first, then the peers in ring order
//     (rank+1, rank+2, ...): block k holds, row by row, the entries whose column lives on that rank;
//   * a few COPY CTAs pull the peers' slices of x over NVLink (128-bit loads from the CUDA-IPC
//     mappings, 8 in flight per thread) into the local full-length x, slice by slice in the same
//     ring order, and bump a per-slice counter (release) when their share of a slice has landed;
//   * the COMPUTE CTAs own a fixed range of rows each, keep the row sums in shared memory, and walk
//     the blocks in order: block 0 needs only the rank's own slice and runs while the first slices
//     are in flight; before block k they wait (acquire) for slice k's counter.  Transfer and math
//     overlap block by block, and y is written once.
// All CTAs are resident at once (grid = what fits), so the waits cannot deadlock.  Row sums add the
// blocks in ring order, not in ascending column order: tolerance parity (1e-12 / 1e-5), like every
// multi-lane kernel.
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

constexpr int GF_THREADS = 256;
constexpr int GF_INFLIGHT = 8;    // entries in flight per thread = rows in flight x entries per row and trip: a block
                                  // holds only nnz / world entries of a row, so short rows need many rows in flight

__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

template <typename T, int LPR, int U>
__global__ void __launch_bounds__(GF_THREADS)
spmv_gather_fused_kernel(uint32_t nloc, const uint32_t *__restrict__ bptr, const uint32_t *__restrict__ bind,
                         const T *__restrict__ bval, GatherPeers gp, T *x_full, T *__restrict__ y, uint32_t ncopy,
                         uint32_t *ready, uint32_t target, uint32_t rows_per_cta, GatherBarrier gb,
                         unsigned long long *timeline) {
    extern __shared__ __align__(16) unsigned char gf_raw[];
    const int G = gp.world;
    if (blockIdx.x < ncopy) {
        // ---- copy role: every WARP owns a contiguous 1/(8 ncopy) share of each peer slice, slices in ring order;
        // warps run independently (no CTA barrier), 8 x 16 bytes in flight per lane ----
        const unsigned lane = lane_id();
        const unsigned long long w = (unsigned long long)blockIdx.x * (GF_THREADS / 32) + (threadIdx.x >> 5);
        const unsigned long long nwarps = (unsigned long long)ncopy * (GF_THREADS / 32);
        if (gb.flags_mine && w == 0 && (int)lane < G && (int)lane != gp.rank) {
            // fused barrier, arrival: this rank's published slice is final (stream order put its writers before us)
            __threadfence_system();
            asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(gb.flags[lane] + gp.rank), "r"(gb.epoch) : "memory");
        }
        if (timeline && w == 0 && lane == 0) timeline[0] = global_timer_ns();
        for (int k = 1; k < G; ++k) {
            const int g = (gp.rank + k) % G;
            if (gb.flags_mine) {         // fused barrier, wait: rank g's slice is final once its epoch shows up here
                if (lane == 0) {
                    const uint64_t t0 = global_timer_ns();
                    for (;;) {
                        uint32_t v;
                        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(gb.flags_mine + g) : "memory");
                        if ((int32_t)(v - gb.epoch) >= 0) break;
                        if (global_timer_ns() - t0 > gb.timeout_ns) { atomicExch(gb.failed, 1u); break; }
                        __nanosleep(100);
                    }
                }
                __syncwarp();
            }
            const unsigned char *src = static_cast<const unsigned char *>(gp.slice[g]);
            unsigned char *dst = reinterpret_cast<unsigned char *>(x_full + gp.start[g]);
            const unsigned long long bytes = (unsigned long long)(gp.start[g + 1] - gp.start[g]) * sizeof(T);
            // whole 16-byte units when both ends are aligned (IPC blocks are; slice starts of f32 vectors may not be)
            const bool wide = ((((unsigned long long)(uintptr_t)src) | ((unsigned long long)(uintptr_t)dst)) & 15ull) == 0;
            const unsigned long long n16 = wide ? bytes / 16 : 0;
            const unsigned long long per = (n16 + nwarps - 1) / nwarps;
            const unsigned long long lo = w * per;
            const unsigned long long hi = lo + per < n16 ? lo + per : n16;
            const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
            uint4 *d4 = reinterpret_cast<uint4 *>(dst);
            unsigned long long i = lo + lane;
            for (; i + 7ull * 32 < hi; i += 8ull * 32) {
                uint4 v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = __ldcg(s4 + i + (unsigned long long)u * 32);
#pragma unroll
                for (int u = 0; u < 8; ++u) d4[i + (unsigned long long)u * 32] = v[u];
            }
            for (; i < hi; i += 32) d4[i] = __ldcg(s4 + i);
            // the unaligned / trailing bytes: element by element, by the first copy warp
            if (w == 0) {
                const T *se = reinterpret_cast<const T *>(src);
                T *de = reinterpret_cast<T *>(dst);
                const unsigned long long n = bytes / sizeof(T);
                for (unsigned long long e = n16 * (16 / sizeof(T)) + lane; e < n; e += 32) de[e] = __ldcg(se + e);
            }
            __syncwarp();
            if (lane == 0) {
                __threadfence();                       // this warp's stores of the slice, before the count
                atomicAdd(ready + k, 1u);
            }
        }
        return;
    }

    // ---- compute role: rows [rs, re), sums in shared memory, blocks in ring order ----
    T *acc = reinterpret_cast<T *>(gf_raw);
    const uint32_t c = blockIdx.x - ncopy;
    const uint64_t rs64 = (uint64_t)c * rows_per_cta;
    if (rs64 >= nloc) return;
    const uint32_t rs = (uint32_t)rs64;
    const uint32_t re = rs + rows_per_cta < nloc ? rs + rows_per_cta : nloc;
    for (uint32_t i = threadIdx.x; i < re - rs; i += GF_THREADS) acc[i] = (T)0;
    __syncthreads();
    const T *own = static_cast<const T *>(gp.slice[gp.rank]) - gp.start[gp.rank];
    for (int k = 0; k < G; ++k) {
        const T *xb = k == 0 ? own : x_full;
        if (timeline && c == 0 && threadIdx.x == 0) timeline[1 + 3 * k] = global_timer_ns();          // block k: wait begins
        if (k > 0) {
            if (threadIdx.x == 0)
                while ((int32_t)(ld_acquire_gpu(ready + k) - target) < 0) __nanosleep(64);
            __syncthreads();
        }
        if (timeline && c == 0 && threadIdx.x == 0) timeline[2 + 3 * k] = global_timer_ns();          // slice k has landed
        const uint32_t *p = bptr + (size_t)k * (nloc + 1);
        constexpr uint32_t RL = GF_THREADS / LPR;          // rows a CTA covers per step of one q
        constexpr int ROWS = GF_INFLIGHT / U;
        const uint32_t sub = threadIdx.x % LPR;
        for (uint32_t base0 = rs; base0 < re; base0 += RL * ROWS) {       // uniform trip count: shuffles below
            const uint32_t base = base0 + threadIdx.x / LPR;
            uint32_t a[ROWS], b[ROWS];
            T s[ROWS];
#pragma unroll
            for (int q = 0; q < ROWS; ++q) {
                const uint32_t r = base + q * RL;
                a[q] = b[q] = 0;
                s[q] = (T)0;
                if (r < re) { a[q] = __ldg(p + r) + sub; b[q] = __ldg(p + r + 1); }
            }
            for (;;) {
                uint32_t col[ROWS][U];
                T v[ROWS][U], xv[ROWS][U];
                bool any = false;
#pragma unroll
                for (int q = 0; q < ROWS; ++q)
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const uint32_t j = a[q] + u * LPR;
                        const bool ok = j < b[q];
                        col[q][u] = ok ? ld_stream(bind + j) : 0xffffffffu;
                        v[q][u] = ok ? ld_stream(bval + j) : (T)0;
                        any |= ok;
                    }
                if (!any) break;
#pragma unroll
                for (int q = 0; q < ROWS; ++q)
#pragma unroll
                    for (int u = 0; u < U; ++u)
                        xv[q][u] = col[q][u] != 0xffffffffu ? __ldcg(xb + col[q][u]) : (T)0;     // L2: x_full changes under L1
#pragma unroll
                for (int q = 0; q < ROWS; ++q) {
#pragma unroll
                    for (int u = 0; u < U; ++u)
                        if (col[q][u] != 0xffffffffu) s[q] += v[q][u] * xv[q][u];
                    a[q] += U * LPR;
                }
            }
#pragma unroll
            for (int q = 0; q < ROWS; ++q) {
#pragma unroll
                for (int o = LPR / 2; o > 0; o >>= 1) s[q] += __shfl_xor_sync(0xffffffffu, s[q], o);
                const uint32_t r = base + q * RL;
                if (sub == 0 && r < re) acc[r - rs] += s[q];      // one owner per row for the whole kernel: no race
            }
        }
        if (timeline && c == 0 && threadIdx.x == 0) timeline[3 + 3 * k] = global_timer_ns();          // block k done (this CTA)
    }
    __syncthreads();
    // a fused barrier that gave up: NaN instead of sums over a half-written x (the status call reports it)
    const bool poisoned = gb.failed && *reinterpret_cast<const volatile uint32_t *>(gb.failed) != 0u;
    for (uint32_t i = threadIdx.x; i < re - rs; i += GF_THREADS) y[rs + i] = poisoned ? (T)NAN : acc[i];
}

}  // namespace

// Launch shape: `ncopy` copy CTAs plus as many compute CTAs as stay resident beside them; the row range
// of a compute CTA (its shared-memory sums) shrinks as the CTAs per SM grow, so the two are found together.
template <typename T, int LPR, int U>
void spmv_gather_fused_t(spl_ctx *ctx, uint32_t nloc, const uint32_t *bptr, const uint32_t *bind, const T *bval,
                         const GatherPeers &gp, T *x_full, T *y, uint32_t *ready, uint32_t epoch, const GatherBarrier &gb,
                         unsigned long long *timeline) {
    auto k = spmv_gather_fused_kernel<T, LPR, U>;
    const char *nc = std::getenv("SPL_GATHER_COPY_CTAS");            // measurement knob
    const uint32_t ncopy = gp.world > 1 ? (nc ? (uint32_t)std::atoi(nc) : std::min<uint32_t>(64u, (uint32_t)ctx->num_sms / 2u)) : 0u;
    uint32_t ncompute = 0, rows_per_cta = 0;
    size_t smem = 0;
    for (int want = 8; want >= 1; --want) {
        const uint32_t total = (uint32_t)ctx->num_sms * (uint32_t)want;
        if (total <= ncopy) continue;
        ncompute = total - ncopy;
        rows_per_cta = (uint32_t)(((uint64_t)nloc + ncompute - 1) / ncompute);
        rows_per_cta = (rows_per_cta + 31u) & ~31u;
        smem = (size_t)rows_per_cta * sizeof(T);
        if (smem > 200 * 1024) continue;
        if (smem > 48 * 1024) SPL_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int resident = 0;
        SPL_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, k, GF_THREADS, smem));
        if (resident >= want) break;
        ncompute = 0;
    }
    SPL_REQUIRE(ncompute > 0, SPL_ERR_UNSUPPORTED,
                "fused gather SpMV: the shard's rows do not fit in the shared-memory sums of one resident grid");
    k<<<ncopy + ncompute, GF_THREADS, smem, ctx->stream>>>(nloc, bptr, bind, bval, gp, x_full, y, ncopy, ready,
                                                          epoch * ncopy * (GF_THREADS / 32), rows_per_cta, gb, timeline);
    check_launch(ctx, "spmv_gather_fused");
}

void spmv_gather_fused(spl_ctx *ctx, int dtype, uint32_t nloc, int world, int rank, const uint64_t *col_starts,
                       const void *const *x_slices, const uint32_t *bptr, const uint32_t *bind, const void *bval,
                       void *x_full, void *y, uint32_t *ready, uint32_t epoch, double entries_per_row_block,
                       void *const *flag_ptrs, uint32_t barrier_epoch, uint32_t timeout_ms, unsigned long long *timeline) {
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
    // lanes per row x entries per lane and trip from the entries a row holds in ONE block (nnz / rows / world):
    // one trip for a typical row
    const double e = entries_per_row_block;
    auto go = [&](auto tag, auto lpr, auto u) {
        using T = decltype(tag);
        spmv_gather_fused_t<T, decltype(lpr)::value, decltype(u)::value>(ctx, nloc, bptr, bind, (const T *)bval, gp, (T *)x_full,
                                                                        (T *)y, ready, epoch, gb, timeline);
    };
    using I1 = std::integral_constant<int, 1>;
    using I2 = std::integral_constant<int, 2>;
    using I4 = std::integral_constant<int, 4>;
    auto pick = [&](auto tag) {
        if (e <= 1.5) go(tag, I1{}, I1{});
        else if (e <= 3.0) go(tag, I1{}, I2{});
        else if (e <= 6.0) go(tag, I2{}, I2{});
        else if (e <= 12.0) go(tag, I2{}, I4{});
        else go(tag, I4{}, I4{});
    };
    if (dtype == SPL_F32) pick(float{});
    else pick(double{});
}

}  // namespace spl
