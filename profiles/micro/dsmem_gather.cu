// dsmem_gather.cu — can a thread-block cluster serve the hot-column table of the power-law SpMV?
//
// Question: the 128 KB shared-memory table of the nnz-split kernel covers 35 % of
// the R-MAT entries; a cluster of 2-8 CTAs could hold 2-8x the columns in distributed shared memory.
// That only pays if scattered 4-byte loads from a PEER CTA's shared memory (ld.shared::cluster) run
// at a useful rate.  This microbenchmark measures, per SM and clock, random 4-byte gathers from
//   (a) the CTA's own shared memory,
//   (b) a table spread uniformly over the cluster (fraction (CS-1)/CS of the gathers are remote),
//   (c) a 64 MB global array (L2-resident), the path the cold gathers take today,
// with 8 independent gathers in flight per thread and 1 024 threads per SM.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o dsmem_gather dsmem_gather.cu
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

namespace cg = cooperative_groups;

#define CK(x)                                                                          \
    do {                                                                               \
        cudaError_t e_ = (x);                                                          \
        if (e_ != cudaSuccess) {                                                       \
            std::fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_));              \
            std::exit(1);                                                              \
        }                                                                              \
    } while (0)

__device__ __forceinline__ uint32_t mix(uint32_t s) {      // xorshift step: independent streams per thread
    s ^= s << 13; s ^= s >> 17; s ^= s << 5;
    return s;
}

constexpr int THREADS = 1024;
constexpr int U = 8;

// table of `per_cta` floats per CTA; indices uniform over the whole cluster's table
__global__ void __launch_bounds__(THREADS, 1)
cluster_gather(uint32_t per_cta, int iters, float *out) {
    extern __shared__ float table[];
    cg::cluster_group cluster = cg::this_cluster();
    const uint32_t cs = cluster.num_blocks();
    for (uint32_t j = threadIdx.x; j < per_cta; j += THREADS) table[j] = (float)(j & 1023) * 1e-3f;
    cluster.sync();
    uint32_t s = (blockIdx.x * THREADS + threadIdx.x) * 2654435761u + 12345u;
    const uint32_t total = per_cta * cs;
    float acc = 0.f;
    for (int it = 0; it < iters; ++it) {
        float v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            s = mix(s);
            const uint32_t idx = (uint32_t)(((uint64_t)s * total) >> 32);
            const uint32_t r = idx / per_cta, o = idx - r * per_cta;
            v[u] = *cluster.map_shared_rank(table + o, r);      // mapa + ld.shared::cluster
        }
#pragma unroll
        for (int u = 0; u < U; ++u) acc += v[u];
    }
    cluster.sync();                      // nobody leaves while a peer may still read its table
    if (acc == 123.456f) out[0] = acc;
}

__global__ void __launch_bounds__(THREADS, 1)
global_gather(const float *__restrict__ x, uint32_t n, int iters, float *out) {
    uint32_t s = (blockIdx.x * THREADS + threadIdx.x) * 2654435761u + 12345u;
    float acc = 0.f;
    for (int it = 0; it < iters; ++it) {
        float v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            s = mix(s);
            v[u] = __ldg(x + (uint32_t)(((uint64_t)s * n) >> 32));
        }
#pragma unroll
        for (int u = 0; u < U; ++u) acc += v[u];
    }
    if (acc == 123.456f) out[0] = acc;
}

int main() {
    int dev = 0, sms = 0, khz = 0;
    CK(cudaSetDevice(dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
    float *out;
    CK(cudaMalloc(&out, 4));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const int iters = 2000;
    const uint32_t per_cta = 128 * 1024 / 4;
    const size_t smem = per_cta * 4;
    CK(cudaFuncSetAttribute(cluster_gather, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(cluster_gather, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    std::printf("# %d SMs, max clock %.0f MHz, %d threads/SM, %d gathers in flight per thread, table 128 KB per CTA\n", sms,
                khz / 1e3, THREADS, U);
    std::printf("# what, cluster, remote_fraction, ms, Ggather/s, gathers/clk/SM (at max clock)\n");
    for (int cs : {1, 2, 4, 8}) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(sms / cs * cs);
        cfg.blockDim = dim3(THREADS);
        cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = cs;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        for (int rep = 0; rep < 3; ++rep) {
            CK(cudaEventRecord(e0));
            CK(cudaLaunchKernelEx(&cfg, cluster_gather, per_cta, iters, out));
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep == 2) {
                const double g = (double)cfg.gridDim.x * THREADS * U * iters;
                std::printf("shared, %d, %.3f, %.3f, %.1f, %.3f\n", cs, (cs - 1.0) / cs, ms, g / ms / 1e6,
                            g / (ms * 1e-3) / (khz * 1e3) / cfg.gridDim.x);
            }
        }
    }
    for (uint32_t mb : {4u, 64u, 512u}) {
        const uint32_t n = mb * 1024 * 1024 / 4;
        float *x;
        CK(cudaMalloc(&x, (size_t)n * 4));
        CK(cudaMemset(x, 0, (size_t)n * 4));
        for (int rep = 0; rep < 3; ++rep) {
            CK(cudaEventRecord(e0));
            global_gather<<<sms, THREADS>>>(x, n, iters / 4, out);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep == 2) {
                const double g = (double)sms * THREADS * U * (iters / 4);
                std::printf("global %u MB, 1, -, %.3f, %.1f, %.3f\n", mb, ms, g / ms / 1e6,
                            g / (ms * 1e-3) / (khz * 1e3) / sms);
            }
        }
        CK(cudaFree(x));
    }
    return 0;
}
