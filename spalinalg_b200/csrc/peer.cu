// peer.cu — peer-visible device memory and a flag barrier over it.
//
// One process per GPU.  A rank allocates its slice of x (and a small flag block) with
// cudaMalloc, exports it with CUDA IPC, and maps its peers' blocks; on an NVSwitch box every
// mapping is a full-bandwidth NVLink path.  spl_spmv_peer then gathers x straight from the
// owning slice, so the "exchange" of the sharded SpMV is part of the kernel's loads.  The only
// thing left between iterations is ordering: a rank's writes to its slice must be visible before
// a peer's next kernel reads them.  peer_barrier_kernel does that on the device: release-store of
// the epoch into every peer's flag block, acquire-spin on the own block, bounded by a timeout so
// a missing peer can never hang the GPU.
#include "kernels.cuh"

namespace spl {

namespace {

struct FlagBlocks {
    uint32_t *block[SPL_MAX_PEERS];
};

__device__ __forceinline__ uint64_t global_timer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__global__ void peer_barrier_kernel(FlagBlocks f, int world, int rank, uint32_t epoch,
                                    uint64_t timeout_ns, uint32_t *timed_out) {
    const int g = threadIdx.x;
    if (g >= world || g == rank) return;
    __threadfence_system();                      // earlier kernels' writes to this rank's slice
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f.block[g] + rank), "r"(epoch) : "memory");
    const uint64_t t0 = global_timer_ns();
    const uint32_t *mine = f.block[rank] + g;
    for (;;) {
        uint32_t v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
        if ((int32_t)(v - epoch) >= 0) break;
        if (global_timer_ns() - t0 > timeout_ns) {
            atomicExch(timed_out, 1u);
            break;
        }
        __nanosleep(200);
    }
}

// Barrier + halo in one launch.  Threads 0..world-1 run the flag protocol above; once every peer has
// arrived, the whole CTA copies this rank's halo — the halo_left columns before and the halo_right
// columns after its own slice — element by element from the owning ranks' slices (peer memory) into
// the padding that surrounds the own slice.  The product that follows then gathers from ONE local
// array (own slice + halo) with the plain local policy: for a banded or stencil shard this moves the
// same few values over NVLink as the in-kernel peer gather, without an owner test on every gather.
struct HaloSlices {
    unsigned char *slice[SPL_MAX_PEERS];
    unsigned long long start[SPL_MAX_PEERS + 1];
    unsigned int halo_left, halo_right, vsize;
};

__global__ void __launch_bounds__(256)
peer_barrier_halo_kernel(FlagBlocks f, int world, int rank, uint32_t epoch, uint64_t timeout_ns, uint32_t *timed_out,
                         HaloSlices h) {
    __shared__ int s_failed;
    if (threadIdx.x == 0) s_failed = 0;
    __syncthreads();
    const int g = threadIdx.x;
    if (g < world && g != rank) {
        __threadfence_system();
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f.block[g] + rank), "r"(epoch) : "memory");
        const uint64_t t0 = global_timer_ns();
        const uint32_t *mine = f.block[rank] + g;
        for (;;) {
            uint32_t v;
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
            if ((int32_t)(v - epoch) >= 0) break;
            if (global_timer_ns() - t0 > timeout_ns) {
                atomicExch(timed_out, 1u);
                s_failed = 1;
                break;
            }
            __nanosleep(200);
        }
    }
    __syncthreads();
    if (s_failed) return;                        // the product behind this writes NaN (XPeer / status call)
    const unsigned long long my0 = h.start[rank], my1 = h.start[rank + 1], n = h.start[world];
    const unsigned int total = h.halo_left + h.halo_right;
    for (unsigned int i = threadIdx.x; i < total; i += blockDim.x) {
        long long col = i < h.halo_left ? (long long)my0 - h.halo_left + i : (long long)my1 + (i - h.halo_left);
        if (col < 0 || (unsigned long long)col >= n) continue;
        int owner = 0;
        for (int q = 0; q < world; ++q)
            if ((unsigned long long)col >= h.start[q] && (unsigned long long)col < h.start[q + 1]) owner = q;
        const unsigned char *src = h.slice[owner] + ((unsigned long long)col - h.start[owner]) * h.vsize;
        unsigned char *dst = h.slice[rank] + (col - (long long)my0) * (long long)h.vsize;     // before / after the own slice
        if (h.vsize == 8) *reinterpret_cast<unsigned long long *>(dst) = __ldcg(reinterpret_cast<const unsigned long long *>(src));
        else *reinterpret_cast<unsigned int *>(dst) = __ldcg(reinterpret_cast<const unsigned int *>(src));
    }
}

// All-gather of x by pulling: every rank copies its peers' slices (peer memory, NVLink reads) into
// its own full-length vector with 128-bit coalesced loads.  One launch moves all slices; the grid
// is split among the source ranks in proportion to their slice lengths.
struct PullSlices {
    const unsigned char *src[SPL_MAX_PEERS];
    unsigned long long dst_off[SPL_MAX_PEERS];     // byte offset of slice g in the full vector
    unsigned long long bytes[SPL_MAX_PEERS];
    int world, rank;
};

__global__ void __launch_bounds__(256) peer_pull_kernel(PullSlices ps, unsigned char *__restrict__ dst) {
    // blockIdx.y = source rank; blocks of a row stride over that slice
    const int g = blockIdx.y;
    if (g >= ps.world || g == ps.rank) return;
    const unsigned char *src = ps.src[g];
    unsigned char *out = dst + ps.dst_off[g];
    const unsigned long long n = ps.bytes[g];
    const unsigned long long n16 = (((unsigned long long)(uintptr_t)src | (unsigned long long)(uintptr_t)out) & 15ull) ? 0ull : n / 16;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
    uint4 *d4 = reinterpret_cast<uint4 *>(out);
    for (; i + 3 * stride < n16; i += 4 * stride) {          // four 16-byte loads in flight per thread
        const uint4 a = __ldcg(s4 + i), b = __ldcg(s4 + i + stride), c = __ldcg(s4 + i + 2 * stride),
                    d = __ldcg(s4 + i + 3 * stride);
        d4[i] = a; d4[i + stride] = b; d4[i + 2 * stride] = c; d4[i + 3 * stride] = d;
    }
    for (; i < n16; i += stride) d4[i] = __ldcg(s4 + i);
    for (unsigned long long b = n16 * 16 + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; b < n; b += stride)
        out[b] = src[b];
}

}  // namespace

void peer_pull(spl_ctx *ctx, int world, int rank, size_t vsize, const uint64_t *starts,
               const void *const *slices, void *x_full) {
    SPL_REQUIRE(world >= 1 && world <= SPL_MAX_PEERS && rank >= 0 && rank < world, SPL_ERR_ARG,
                "world must be 1..8 and rank inside it");
    if (world == 1) return;
    PullSlices ps{};
    ps.world = world;
    ps.rank = rank;
    for (int g = 0; g < world; ++g) {
        SPL_REQUIRE(starts[g] <= starts[g + 1], SPL_ERR_ARG, "starts must be non-decreasing");
        SPL_REQUIRE(slices[g] || starts[g] == starts[g + 1], SPL_ERR_ARG, "NULL slice");
        ps.src[g] = static_cast<const unsigned char *>(slices[g]);
        ps.dst_off[g] = starts[g] * vsize;
        ps.bytes[g] = (starts[g + 1] - starts[g]) * vsize;
    }
    dim3 grid((unsigned)ctx->num_sms * 2u, (unsigned)world);
    peer_pull_kernel<<<grid, 256, 0, ctx->stream>>>(ps, static_cast<unsigned char *>(x_full));
    check_launch(ctx, "peer_pull");
}

void peer_barrier_halo(spl_ctx *ctx, int world, int rank, void *const *flag_ptrs, uint32_t epoch, uint32_t timeout_ms,
                       size_t vsize, const uint64_t *starts, void *const *slices, uint32_t halo_left, uint32_t halo_right) {
    SPL_REQUIRE(world >= 1 && world <= SPL_MAX_PEERS && rank >= 0 && rank < world, SPL_ERR_ARG,
                "world must be 1..8 and rank inside it");
    FlagBlocks f{};
    HaloSlices h{};
    for (int g = 0; g < world; ++g) {
        SPL_REQUIRE(flag_ptrs[g], SPL_ERR_ARG, "NULL flag block");
        SPL_REQUIRE(slices[g] || starts[g] == starts[g + 1], SPL_ERR_ARG, "NULL slice");
        f.block[g] = static_cast<uint32_t *>(flag_ptrs[g]);
        h.slice[g] = static_cast<unsigned char *>(slices[g]);
    }
    for (int g = 0; g <= SPL_MAX_PEERS; ++g) h.start[g] = starts[g < world ? g : world];
    h.halo_left = halo_left;
    h.halo_right = halo_right;
    h.vsize = (unsigned int)vsize;
    peer_barrier_halo_kernel<<<1, 256, 0, ctx->stream>>>(f, world, rank, epoch,
                                                        (uint64_t)(timeout_ms ? timeout_ms : 2000) * 1000000ull,
                                                        ctx->d_scratch + 32, h);
    check_launch(ctx, "peer_barrier_halo");
}

void peer_barrier(spl_ctx *ctx, int world, int rank, void *const *flag_ptrs, uint32_t epoch,
                  uint32_t timeout_ms) {
    SPL_REQUIRE(world >= 1 && world <= SPL_MAX_PEERS && rank >= 0 && rank < world, SPL_ERR_ARG,
                "world must be 1..8 and rank inside it");
    if (world == 1) return;
    FlagBlocks f{};
    for (int g = 0; g < world; ++g) {
        SPL_REQUIRE(flag_ptrs[g], SPL_ERR_ARG, "NULL flag block");
        f.block[g] = static_cast<uint32_t *>(flag_ptrs[g]);
    }
    peer_barrier_kernel<<<1, 32, 0, ctx->stream>>>(f, world, rank, epoch,
                                                   (uint64_t)(timeout_ms ? timeout_ms : 2000) * 1000000ull,
                                                   ctx->d_scratch + 32);
    check_launch(ctx, "peer_barrier");
}

}  // namespace spl
