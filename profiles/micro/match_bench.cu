// Microbenchmark: cost of the per-warp "who shares my digit" step of a radix pass on sm_100a.
//   v0 match.any instruction    v1 eight ballots over the digit bits    v2 shared-memory atomics
//   v3 lane bitmask per digit built with shared-memory atomicOr (same `peers` mask as v0/v1, stable)
//   v4 as v3, four keys per round on four mask tables (one syncwarp per phase and round)
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o match_bench match_bench.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned lanemask_lt() { unsigned m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }

template <int V>
__global__ void __launch_bounds__(512) k(const uint32_t *keys, size_t n, uint32_t *out) {
    __shared__ uint32_t hist[16][256];
    extern __shared__ uint32_t mask_raw[];
    uint32_t (*mask)[16][256] = reinterpret_cast<uint32_t (*)[16][256]>(mask_raw);
    for (int j = threadIdx.x; j < 16 * 256; j += 512) { (&hist[0][0])[j] = 0; (&mask[0][0][0])[j] = 0; if (V == 4) { (&mask[1][0][0])[j] = 0; (&mask[2][0][0])[j] = 0; (&mask[3][0][0])[j] = 0; } }
    __syncthreads();
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t acc = 0;
    for (size_t i = (size_t)blockIdx.x * 512 * 8 + threadIdx.x; i < n; i += (size_t)gridDim.x * 512 * 8) {
        uint32_t d[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) d[u] = (i + u * 512 < n ? keys[i + u * 512] : 0u) & 255u;
        if (V == 4) {
#pragma unroll
            for (int r = 0; r < 8; r += 4) {
                unsigned peers[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) atomicOr(&mask[q][warp][d[r + q]], 1u << lane);
                __syncwarp();
#pragma unroll
                for (int q = 0; q < 4; ++q) peers[q] = mask[q][warp][d[r + q]];
                __syncwarp();
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (lane == (unsigned)(__ffs(peers[q]) - 1)) {
                        mask[q][warp][d[r + q]] = 0;
                        hist[warp][d[r + q]] += __popc(peers[q]);
                    }
                    acc += __popc(peers[q] & lanemask_lt());
                }
                __syncwarp();
            }
            continue;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (V == 0) {
                unsigned peers = __match_any_sync(0xffffffffu, d[u]);
                if (lane == (unsigned)(__ffs(peers) - 1)) hist[warp][d[u]] += __popc(peers);
                acc += __popc(peers & lanemask_lt());
                __syncwarp();
            } else if (V == 1) {
                unsigned peers = 0xffffffffu;
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    const unsigned bal = __ballot_sync(0xffffffffu, (d[u] >> b) & 1u);
                    peers &= ((d[u] >> b) & 1u) ? bal : ~bal;
                }
                if (lane == (unsigned)(__ffs(peers) - 1)) hist[warp][d[u]] += __popc(peers);
                acc += __popc(peers & lanemask_lt());
                __syncwarp();
            } else if (V == 2) {
                acc += atomicAdd(&hist[warp][d[u]], 1u);
            } else {
                atomicOr(&mask[0][warp][d[u]], 1u << lane);
                __syncwarp();
                const unsigned peers = mask[0][warp][d[u]];
                __syncwarp();
                if (lane == (unsigned)(__ffs(peers) - 1)) {
                    mask[0][warp][d[u]] = 0;
                    hist[warp][d[u]] += __popc(peers);
                }
                acc += __popc(peers & lanemask_lt());
                __syncwarp();
            }
        }
    }
    __syncthreads();
    if (threadIdx.x < 256) {
        uint32_t c = 0;
        for (int w = 0; w < 16; ++w) c += hist[w][threadIdx.x];
        atomicAdd(out + threadIdx.x, c + (acc & 1));
    }
}

int main() {
    const size_t n = 1ull << 28;
    uint32_t *keys, *out;
    cudaMalloc(&keys, n * 4);
    cudaMalloc(&out, 1024);
    uint32_t *h = (uint32_t *)malloc(n * 4);
    uint32_t s = 12345;
    for (size_t i = 0; i < n; ++i) { s = s * 1664525u + 1013904223u; h[i] = s >> 11; }
    cudaMemcpy(keys, h, n * 4, cudaMemcpyHostToDevice);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int v = 0; v < 5; ++v) {
        for (int rep = 0; rep < 3; ++rep) {
            cudaMemset(out, 0, 1024);
            cudaEventRecord(a);
            if (v == 0) k<0><<<148 * 4, 512, 16 * 1024>>>(keys, n, out);
            if (v == 1) k<1><<<148 * 4, 512, 16 * 1024>>>(keys, n, out);
            if (v == 2) k<2><<<148 * 4, 512, 16 * 1024>>>(keys, n, out);
            if (v == 3) k<3><<<148 * 4, 512, 16 * 1024>>>(keys, n, out);
            if (v == 4) { cudaFuncSetAttribute(k<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024); k<4><<<148 * 4, 512, 64 * 1024>>>(keys, n, out); }
            cudaEventRecord(b);
            cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            if (rep == 2) printf("variant %d: %.3f ms  %.1f Gkeys/s  %.0f GB/s\n", v, ms, n / ms / 1e6, n * 4 / ms / 1e6);
        }
    }
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
