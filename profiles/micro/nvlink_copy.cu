// nvlink_copy.cu — how fast can ONE kernel move a peer's slice of x over NVLink, and with how many SMs?
//
// Question behind it (profiles/r2_spmv_notes.md, "General shards"): in the fused all-gather + SpMV the
// slices arrived at ~220 GB/s per rank with 64 copy CTAs of 128-bit loads, far below the link.  Variants,
// every device of the box running the same thing at once (one process, peer access enabled; the
// library's CUDA-IPC mappings take the same path):
//   ldst  pull / push : warps own contiguous shares, 8 x 16 bytes in flight per lane (the round-2 copy role)
//   tma   pull / push : one thread per CTA, cp.async.bulk global -> shared ring -> global (bulk groups)
//   ce                : cudaMemcpyPeerAsync per peer (the copy engines, no SM)
// pull = remote loads + local stores; push = local loads + remote (posted) stores.
//
// nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o nvlink_copy nvlink_copy.cu
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x)                                                                                     \
    do {                                                                                          \
        cudaError_t e_ = (x);                                                                     \
        if (e_ != cudaSuccess) {                                                                  \
            std::fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
            std::exit(1);                                                                         \
        }                                                                                         \
    } while (0)

constexpr int MAXD = 8;
struct Pairs {
    const unsigned char *src[MAXD];
    unsigned char *dst[MAXD];
    int n;
    unsigned long long bytes;
};

__global__ void __launch_bounds__(256) copy_ldst_kernel(Pairs p, int concurrent) {
    const unsigned lane = threadIdx.x & 31u;
    unsigned long long w = (unsigned long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    unsigned long long nw = (unsigned long long)gridDim.x * 8;
    int k0 = 0, k1 = p.n;
    if (concurrent) {           // the CTAs are dealt round-robin to the pairs, all pairs move at once
        const int k = blockIdx.x % p.n;
        k0 = k; k1 = k + 1;
        w = (unsigned long long)(blockIdx.x / p.n) * 8 + (threadIdx.x >> 5);
        nw = (unsigned long long)((gridDim.x - k + p.n - 1) / p.n) * 8;
    }
    for (int k = k0; k < k1; ++k) {
        const uint4 *s4 = reinterpret_cast<const uint4 *>(p.src[k]);
        uint4 *d4 = reinterpret_cast<uint4 *>(p.dst[k]);
        const unsigned long long n16 = p.bytes / 16, per = (n16 + nw - 1) / nw, lo = w * per,
                                 hi = lo + per < n16 ? lo + per : n16;
        unsigned long long i = lo + lane;
        for (; i + 7ull * 32 < hi; i += 8ull * 32) {
            uint4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = __ldcg(s4 + i + (unsigned long long)u * 32);
#pragma unroll
            for (int u = 0; u < 8; ++u) d4[i + (unsigned long long)u * 32] = v[u];
        }
        for (; i < hi; i += 32) d4[i] = __ldcg(s4 + i);
    }
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// one thread per CTA drives a ring of `stages` chunks: loads run `ahead` chunks in front of the stores
__global__ void __launch_bounds__(32) copy_tma_kernel(Pairs p, uint32_t chunk, int stages, int ahead, int concurrent) {
    extern __shared__ __align__(128) unsigned char ring[];
    __shared__ __align__(8) uint64_t full[16];
    if (threadIdx.x != 0) return;
    for (int s = 0; s < stages; ++s)
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(full + s)), "r"(1u));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const unsigned long long per_pair = (p.bytes + chunk - 1) / chunk;
    const unsigned long long total = per_pair * p.n;
    // chunk id c -> (pair, offset): sequential = pair-major (pairs one after the other), concurrent = pair-minor
    const unsigned long long mine = (total - blockIdx.x + gridDim.x - 1) / gridDim.x;
    auto locate = [&](unsigned long long i, int &k, unsigned long long &off, uint32_t &len) {
        const unsigned long long c = blockIdx.x + i * gridDim.x;
        unsigned long long j;
        if (concurrent) { k = (int)(c % p.n); j = c / p.n; }
        else { k = (int)(c / per_pair); j = c % per_pair; }
        off = j * chunk;
        len = (uint32_t)(p.bytes - off < chunk ? p.bytes - off : chunk);
    };
    const int pend = stages - 1 - ahead;              // bulk store groups that may still be reading shared memory
    for (unsigned long long i = 0; i < mine + ahead; ++i) {
        if (i < mine) {
            const int s = (int)(i % stages);
            if (i >= (unsigned long long)stages) {
                if (pend == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                else if (pend == 1) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                else if (pend == 2) asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
                else asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
            }
            int k; unsigned long long off; uint32_t len;
            locate(i, k, off, len);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(full + s)), "r"(len) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             smem_u32(ring + (size_t)s * chunk)),
                         "l"(p.src[k] + off), "r"(len), "r"(smem_u32(full + s))
                         : "memory");
        }
        if (i >= (unsigned long long)ahead) {
            const unsigned long long j = i - ahead;
            const int s = (int)(j % stages);
            const uint32_t parity = (uint32_t)((j / stages) & 1);
            asm volatile(
                "{\n.reg .pred q;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 q, [%0], %1;\n@q bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(
                    smem_u32(full + s)),
                "r"(parity)
                : "memory");
            int k; unsigned long long off; uint32_t len;
            locate(j, k, off, len);
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(p.dst[k] + off),
                         "r"(smem_u32(ring + (size_t)s * chunk)), "r"(len)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main(int argc, char **argv) {
    int nd = 0;
    CK(cudaGetDeviceCount(&nd));
    if (argc > 1) nd = std::min(nd, std::atoi(argv[1]));
    nd = std::min(nd, MAXD);
    if (nd < 2) { std::printf("needs >= 2 GPUs (found %d)\n", nd); return 0; }
    std::vector<unsigned long long> sizes = {5000000ull / 16 * 16, 20000000ull};
    const int reps = 20;
    std::vector<cudaStream_t> st(nd);
    std::vector<cudaEvent_t> e0(nd), e1(nd);
    for (int d = 0; d < nd; ++d) {
        CK(cudaSetDevice(d));
        for (int g = 0; g < nd; ++g)
            if (g != d) {
                int can = 0;
                CK(cudaDeviceCanAccessPeer(&can, d, g));
                if (!can) { std::printf("no peer access %d -> %d\n", d, g); return 0; }
                cudaError_t e = cudaDeviceEnablePeerAccess(g, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CK(e);
                (void)cudaGetLastError();
            }
        CK(cudaStreamCreate(&st[d]));
        CK(cudaEventCreate(&e0[d]));
        CK(cudaEventCreate(&e1[d]));
        CK(cudaFuncSetAttribute(copy_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    }
    std::printf("devices %d, reps %d; GB/s = bytes RECEIVED (pull) or SENT (push) per device and second, min over devices\n", nd, reps);
    for (unsigned long long B : sizes) {
        std::vector<unsigned char *> slice(nd), fullv(nd);
        std::vector<unsigned char> host(B * nd), pat(B);
        for (int d = 0; d < nd; ++d) {
            CK(cudaSetDevice(d));
            CK(cudaMalloc(&slice[d], B));
            CK(cudaMalloc(&fullv[d], B * nd));
            for (unsigned long long i = 0; i < B; ++i) pat[i] = (unsigned char)((i * 131u + d * 17u + (i >> 12)) & 0xff);
            CK(cudaMemcpy(slice[d], pat.data(), B, cudaMemcpyHostToDevice));
        }
        auto pairs_of = [&](int d, bool push) {
            Pairs p{};
            p.bytes = B;
            for (int k = 1; k < nd; ++k) {
                const int g = (d + k) % nd;
                p.src[p.n] = push ? slice[d] : slice[g];
                p.dst[p.n] = push ? fullv[g] + (size_t)d * B : fullv[d] + (size_t)g * B;
                ++p.n;
            }
            return p;
        };
        auto verify = [&](const char *what) {
            for (int d = 0; d < nd; ++d) { CK(cudaSetDevice(d)); CK(cudaDeviceSynchronize()); }
            CK(cudaSetDevice(0));
            CK(cudaMemcpy(host.data(), fullv[0], B * nd, cudaMemcpyDeviceToHost));
            for (int g = 1; g < nd; ++g)
                for (unsigned long long i = 0; i < B; ++i) {
                    const unsigned char want = (unsigned char)((i * 131u + g * 17u + (i >> 12)) & 0xff);
                    if (host[(size_t)g * B + i] != want) {
                        std::printf("MISMATCH %s slice %d byte %llu\n", what, g, i);
                        std::exit(2);
                    }
                }
        };
        auto clear = [&]() {
            for (int d = 0; d < nd; ++d) { CK(cudaSetDevice(d)); CK(cudaMemset(fullv[d], 0, B * nd)); CK(cudaDeviceSynchronize()); }
        };
        auto run = [&](const char *name, auto launch) {
            clear();
            for (int d = 0; d < nd; ++d) { CK(cudaSetDevice(d)); launch(d); }
            verify(name);
            for (int w = 0; w < 2; ++w)
                for (int d = 0; d < nd; ++d) { CK(cudaSetDevice(d)); launch(d); }
            for (int d = 0; d < nd; ++d) { CK(cudaSetDevice(d)); CK(cudaDeviceSynchronize()); }
            for (int d = 0; d < nd; ++d) { CK(cudaSetDevice(d)); CK(cudaEventRecord(e0[d], st[d])); }
            for (int r = 0; r < reps; ++r)
                for (int d = 0; d < nd; ++d) { CK(cudaSetDevice(d)); launch(d); }
            for (int d = 0; d < nd; ++d) { CK(cudaSetDevice(d)); CK(cudaEventRecord(e1[d], st[d])); }
            float worst = 0;
            for (int d = 0; d < nd; ++d) {
                CK(cudaSetDevice(d));
                CK(cudaEventSynchronize(e1[d]));
                float ms = 0;
                CK(cudaEventElapsedTime(&ms, e0[d], e1[d]));
                worst = std::max(worst, ms);
            }
            const double us = worst * 1e3 / reps;
            std::printf("  %-44s %8.1f us  %7.1f GB/s\n", name, us, (double)B * (nd - 1) / (us * 1e-6) / 1e9);
            std::fflush(stdout);
        };
        std::printf("slice %.1f MB, %d peers -> %.1f MB per device\n", B / 1e6, nd - 1, B * (nd - 1) / 1e6);
        char name[128];
        for (int push = 0; push < 2; ++push) {
            for (int conc = 0; conc < 2; ++conc)
                for (int ctas : {16, 32, 64, 148, 296}) {
                    std::snprintf(name, sizeof name, "ldst %s %s ctas=%d", push ? "push" : "pull", conc ? "allpeers" : "ring", ctas);
                    run(name, [&](int d) {
                        copy_ldst_kernel<<<ctas, 256, 0, st[d]>>>(pairs_of(d, push), conc);
                        CK(cudaGetLastError());
                    });
                }
            for (int conc = 0; conc < 2; ++conc)
                for (int ctas : {8, 16, 32, 64})
                    for (int cfg = 0; cfg < 4; ++cfg) {
                        const uint32_t chunk = cfg == 0 ? 8192u : cfg == 1 ? 16384u : cfg == 2 ? 32768u : 16384u;
                        const int stages = cfg == 3 ? 12 : 6, ahead = cfg == 3 ? 9 : 4;
                        std::snprintf(name, sizeof name, "tma  %s %s ctas=%d chunk=%uK stages=%d", push ? "push" : "pull",
                                      conc ? "allpeers" : "ring", ctas, chunk / 1024, stages);
                        run(name, [&](int d) {
                            copy_tma_kernel<<<ctas, 32, (size_t)chunk * stages, st[d]>>>(pairs_of(d, push), chunk, stages, ahead, conc);
                            CK(cudaGetLastError());
                        });
                    }
        }
        run("ce   cudaMemcpyPeerAsync per peer", [&](int d) {
            const Pairs p = pairs_of(d, false);
            for (int k = 0; k < p.n; ++k) {
                const int g = (d + 1 + k) % nd;
                CK(cudaMemcpyPeerAsync(p.dst[k], d, p.src[k], g, B, st[d]));
            }
        });
        for (int d = 0; d < nd; ++d) { CK(cudaSetDevice(d)); CK(cudaFree(slice[d])); CK(cudaFree(fullv[d])); }
    }
    return 0;
}
