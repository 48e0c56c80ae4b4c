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

}  // namespace

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
