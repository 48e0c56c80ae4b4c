// scan.cuh — device-wide exclusive scan of uint32 (reduce / spine / downsweep) and the
// flag-compaction offsets built on it.  Hand-written with warp shuffles; no CUB.
#pragma once

#include "common.cuh"

namespace spl {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_IPT = 16;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_IPT;   // 4096 elements per block

// Per-tile sums.  Thread t owns elements [t*IPT, (t+1)*IPT) of the tile (uint4 loads).
static __global__ void __launch_bounds__(SCAN_THREADS)
scan_reduce_kernel(const uint32_t *__restrict__ in, uint32_t n, uint32_t *__restrict__ tile_sums) {
    __shared__ uint32_t ws[SCAN_THREADS / 32 + 1];
    const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_IPT;
    uint32_t s = 0;
    if (base + SCAN_IPT <= n) {
        const uint4 *p = reinterpret_cast<const uint4 *>(in + base);
#pragma unroll
        for (int i = 0; i < SCAN_IPT / 4; ++i) {
            uint4 v = __ldg(p + i);
            s += v.x + v.y + v.z + v.w;
        }
    } else {
        for (int i = 0; i < SCAN_IPT; ++i)
            if (base + i < n) s += in[base + i];
    }
    uint32_t total;
    block_exclusive_scan(s, ws, &total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// Single-block in-place exclusive scan of `m` values (the spine); data[m] receives the total.
static __global__ void __launch_bounds__(1024) scan_spine_kernel(uint32_t *data, uint32_t m) {
    __shared__ uint32_t ws[33];
    const uint32_t per = (m + blockDim.x - 1) / blockDim.x;
    const uint64_t lo = (uint64_t)threadIdx.x * per;
    const uint64_t hi = lo + per < m ? lo + per : m;
    uint32_t s = 0;
    for (uint64_t i = lo; i < hi; ++i) s += data[i];
    uint32_t total;
    uint32_t run = block_exclusive_scan(s, ws, &total);
    for (uint64_t i = lo; i < hi; ++i) {
        uint32_t v = data[i];
        data[i] = run;
        run += v;
    }
    if (threadIdx.x == 0) data[m] = total;
}

// out[i] = tile_offset + exclusive prefix inside the tile, for i < n; out[n] = grand total.
static __global__ void __launch_bounds__(SCAN_THREADS)
scan_downsweep_kernel(const uint32_t *__restrict__ in, uint32_t n,
                      const uint32_t *__restrict__ tile_offsets, uint32_t *__restrict__ out) {
    __shared__ uint32_t ws[SCAN_THREADS / 32 + 1];
    const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_IPT;
    uint32_t v[SCAN_IPT];
    uint32_t s = 0;
    // whole 64-byte group of 16-byte aligned arrays: 128-bit loads and stores
    const bool full = base + SCAN_IPT <= n && ((((uintptr_t)in) | ((uintptr_t)out)) & 15u) == 0;
    if (full) {
        const uint4 *p = reinterpret_cast<const uint4 *>(in + base);
#pragma unroll
        for (int i = 0; i < SCAN_IPT / 4; ++i) {
            const uint4 q = __ldg(p + i);
            v[4 * i] = q.x; v[4 * i + 1] = q.y; v[4 * i + 2] = q.z; v[4 * i + 3] = q.w;
        }
#pragma unroll
        for (int i = 0; i < SCAN_IPT; ++i) s += v[i];
    } else {
#pragma unroll
        for (int i = 0; i < SCAN_IPT; ++i) {
            v[i] = (base + i < n) ? in[base + i] : 0u;
            s += v[i];
        }
    }
    uint32_t run = block_exclusive_scan(s, ws, nullptr) + tile_offsets[blockIdx.x];
    if (full) {
        uint4 *o = reinterpret_cast<uint4 *>(out + base);
#pragma unroll
        for (int i = 0; i < SCAN_IPT / 4; ++i) {
            uint4 q;
            q.x = run; run += v[4 * i];
            q.y = run; run += v[4 * i + 1];
            q.z = run; run += v[4 * i + 2];
            q.w = run; run += v[4 * i + 3];
            o[i] = q;
        }
    } else {
#pragma unroll
        for (int i = 0; i < SCAN_IPT; ++i) {
            if (base + i < n) out[base + i] = run;
            run += v[i];
        }
    }
    // the thread owning element n-1 also writes the grand total
    if (n > 0 && base <= (uint64_t)n - 1 && (uint64_t)n - 1 < base + SCAN_IPT) out[n] = run;
}

// Exclusive scan: out[0..n] (n+1 slots) from in[0..n).  in and out may alias only if identical
// is NOT allowed (out is one longer); pass distinct buffers.
inline void exclusive_scan_u32(spl_ctx *ctx, const uint32_t *in, uint32_t n, uint32_t *out) {
    if (n == 0) {
        SPL_CUDA(cudaMemsetAsync(out, 0, sizeof(uint32_t), ctx->stream));
        return;
    }
    const unsigned tiles = div_up(n, SCAN_TILE);
    Tmp<uint32_t> sums(ctx, tiles + 1);
    scan_reduce_kernel<<<tiles, SCAN_THREADS, 0, ctx->stream>>>(in, n, sums);
    check_launch(ctx, "scan_reduce");
    scan_spine_kernel<<<1, 1024, 0, ctx->stream>>>(sums, tiles);
    check_launch(ctx, "scan_spine");
    scan_downsweep_kernel<<<tiles, SCAN_THREADS, 0, ctx->stream>>>(in, n, sums, out);
    check_launch(ctx, "scan_downsweep");
}

}  // namespace spl
