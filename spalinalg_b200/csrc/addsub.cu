// addsub.cu — C = A + B / A - B on compressed arrays of the same format.
//
// Reference: impl Add/Sub for &CsrMatrix (src/csr/ops/add.rs:5-75, sub.rs:5-75) and &CscMatrix
// (src/csc/ops/add.rs:5-70, sub.rs:5-70).  There: three transposes around a stamp-array merge.
// Result: per segment the sorted union of the two index sets, explicit zeros kept; value a+b
// (a-b) where both are stored, a where only lhs, b (-b, sub.rs:47) where only rhs.  One IEEE
// operation per output, so the device result is bit-exact.
// Device formulation: two passes over the segments, symbolic (count the union) then numeric
// (merge); the count's exclusive scan is the output pointer array.  No transposes needed:
// inputs are already index-sorted inside each segment (CsrMatrix::new asserts it).
// Algorithmic bytes: (nnzA + nnzB + nnzC)*(4+V) + 3(n+1)*4; the symbolic pass re-reads the
// indices (overhead, not credit).
#include "kernels.cuh"
#include "scan.cuh"

namespace spl {

namespace {

// One thread merges one segment (row of CSR / column of CSC).  A CTA owns blockDim.x consecutive
// segments, whose entries are contiguous in A, B and C: it stages A's and B's slices in shared
// memory with coalesced loads (all of a thread's loads in flight at once), merges out of shared
// memory (a merge step costs a shared-memory latency, not an HBM one), parks the result in shared
// memory and streams it out coalesced.  Only indices are staged: the merge records, per output,
// the slots its value(s) come from, and the values go global -> register -> global in the
// coalesced stream-out.  Shared-memory positions are skewed by one word per 32 so
// that segments of 8, 16, 32 ... entries do not all start in the same bank.  A CTA whose slices do
// not fit (skewed rows) merges straight from global memory instead.
__device__ __forceinline__ uint32_t skew(uint32_t j) { return j + (j >> 5); }

// Global -> shared copies that never pass through registers (cp.async, SASS LDGSTS): a thread
// issues all of its copies back to back and waits once, so a CTA has its whole slice in flight.
// (A plain load/store loop stalls on every store until its load has landed: in-order issue.)
template <int BYTES>
__device__ __forceinline__ void cp_async(void *smem_dst, const void *gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(smem_dst)),
                 "l"(gmem_src), "n"(BYTES)
                 : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

struct Caps {
    uint32_t a, b;   // staged entries of A and B per CTA (C's stage holds a + b)
};

template <typename T, bool SUB, bool NUMERIC>
__global__ void __launch_bounds__(128)
union_block_kernel(uint32_t nmajor, const uint32_t *__restrict__ aptr,
                   const uint32_t *__restrict__ aind, const T *__restrict__ aval,
                   const uint32_t *__restrict__ bptr, const uint32_t *__restrict__ bind,
                   const T *__restrict__ bval, const uint32_t *__restrict__ cptr,
                   uint32_t *__restrict__ cnt, uint32_t *__restrict__ cind, T *__restrict__ cval,
                   Caps caps) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t sa_n = skew(caps.a) + 1, sb_n = skew(caps.b) + 1, sc_n = skew(caps.a + caps.b) + 1;
    uint32_t *sa_ind = reinterpret_cast<uint32_t *>(smem_raw);
    uint32_t *sb_ind = sa_ind + sa_n;
    uint32_t *sc_ind = sb_ind + sb_n;                       // numeric pass only
    uint16_t *sc_sa = reinterpret_cast<uint16_t *>(sc_ind + sc_n);   // source slot in A's slice
    uint16_t *sc_sb = sc_sa + sc_n;                                  // source slot in B's slice
    constexpr uint16_t kNone = 0xffffu;

    const uint32_t m0 = blockIdx.x * blockDim.x;
    const uint32_t m1 = min(m0 + blockDim.x, nmajor);
    const uint32_t m = m0 + threadIdx.x;
    const uint32_t a0 = __ldg(aptr + m0), a1 = __ldg(aptr + m1);
    const uint32_t b0 = __ldg(bptr + m0), b1 = __ldg(bptr + m1);
    const bool staged = a1 - a0 <= caps.a && b1 - b0 <= caps.b;     // uniform over the CTA
    uint32_t pa = 0, ea = 0, pb = 0, eb = 0;
    if (m < m1) {
        pa = __ldg(aptr + m); ea = __ldg(aptr + m + 1);
        pb = __ldg(bptr + m); eb = __ldg(bptr + m + 1);
    }
    if (staged) {
        // indices only: the values never pass through shared memory
        for (uint32_t j = threadIdx.x; j < a1 - a0; j += blockDim.x) cp_async<4>(sa_ind + skew(j), aind + a0 + j);
        for (uint32_t j = threadIdx.x; j < b1 - b0; j += blockDim.x) cp_async<4>(sb_ind + skew(j), bind + b0 + j);
        cp_async_wait_all();
        __syncthreads();
        pa -= a0; ea -= a0; pb -= b0; eb -= b0;
        if (!NUMERIC) {
            uint32_t c = 0;
            while (pa < ea && pb < eb) {
                const uint32_t ia = sa_ind[skew(pa)], ib = sb_ind[skew(pb)];
                pa += ia <= ib;
                pb += ib <= ia;
                ++c;
            }
            if (m < m1) cnt[m] = c + (ea - pa) + (eb - pb);
            return;
        }
        // merge: output index and where its value(s) come from
        const uint32_t c0 = __ldg(cptr + m0), c1 = __ldg(cptr + m1);
        uint32_t pc = m < m1 ? __ldg(cptr + m) - c0 : 0u;
        while (pa < ea && pb < eb) {
            const uint32_t ia = sa_ind[skew(pa)], ib = sb_ind[skew(pb)];
            const uint32_t q = skew(pc);
            sc_ind[q] = ia < ib ? ia : ib;
            sc_sa[q] = ia <= ib ? (uint16_t)pa : kNone;
            sc_sb[q] = ib <= ia ? (uint16_t)pb : kNone;
            pa += ia <= ib;
            pb += ib <= ia;
            ++pc;
        }
        for (; pa < ea; ++pa, ++pc) { const uint32_t q = skew(pc); sc_ind[q] = sa_ind[skew(pa)]; sc_sa[q] = (uint16_t)pa; sc_sb[q] = kNone; }
        for (; pb < eb; ++pb, ++pc) { const uint32_t q = skew(pc); sc_ind[q] = sb_ind[skew(pb)]; sc_sa[q] = kNone; sc_sb[q] = (uint16_t)pb; }
        __syncthreads();
        // stream out: consecutive threads, consecutive outputs; source slots ascend with the
        // output position, so the value loads are (nearly) coalesced too.  4 outputs in flight.
        const T *av = aval + a0, *bv = bval + b0;
        constexpr int U = 4;
        for (uint32_t j0 = threadIdx.x; j0 < c1 - c0; j0 += U * blockDim.x) {
            uint16_t qa[U], qb[U];
            uint32_t ix[U];
            T x[U], z[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const uint32_t j = j0 + u * blockDim.x;
                const uint32_t q = skew(j < c1 - c0 ? j : j0);
                qa[u] = sc_sa[q]; qb[u] = sc_sb[q]; ix[u] = sc_ind[q];
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                x[u] = qa[u] != kNone ? __ldg(av + qa[u]) : (T)0;
                z[u] = qb[u] != kNone ? __ldg(bv + qb[u]) : (T)0;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const uint32_t j = j0 + u * blockDim.x;
                if (j >= c1 - c0) continue;
                T r;
                if (qb[u] == kNone) r = x[u];                                   // only lhs: a
                else if (qa[u] == kNone) r = SUB ? flip_sign(z[u]) : z[u];      // only rhs: b / -b
                else r = SUB ? x[u] - z[u] : x[u] + z[u];                       // both: one IEEE op
                cind[c0 + j] = ix[u];
                cval[c0 + j] = r;
            }
        }
        return;
    }
    // slices too long for the stage: merge from global memory
    if (m >= m1) return;
    if (!NUMERIC) {
        uint32_t c = 0;
        while (pa < ea && pb < eb) {
            const uint32_t ia = __ldg(aind + pa), ib = __ldg(bind + pb);
            pa += ia <= ib;
            pb += ib <= ia;
            ++c;
        }
        cnt[m] = c + (ea - pa) + (eb - pb);
        return;
    }
    uint32_t pc = __ldg(cptr + m);
    while (pa < ea && pb < eb) {
        const uint32_t ia = __ldg(aind + pa), ib = __ldg(bind + pb);
        if (ia < ib) {
            cind[pc] = ia; cval[pc] = aval[pa]; ++pa;
        } else if (ib < ia) {
            cind[pc] = ib; cval[pc] = SUB ? flip_sign(bval[pb]) : bval[pb]; ++pb;
        } else {
            cind[pc] = ia; cval[pc] = SUB ? aval[pa] - bval[pb] : aval[pa] + bval[pb]; ++pa; ++pb;
        }
        ++pc;
    }
    for (; pa < ea; ++pa, ++pc) { cind[pc] = aind[pa]; cval[pc] = aval[pa]; }
    for (; pb < eb; ++pb, ++pc) { cind[pc] = bind[pb]; cval[pc] = SUB ? flip_sign(bval[pb]) : bval[pb]; }
}

// Rows per CTA and stage capacities from the mean segment lengths: 25 % headroom over the mean,
// at most ~96 KB of shared memory per CTA so that at least two CTAs stay resident.
struct BlockPlan {
    unsigned rows;
    Caps caps;
    size_t smem_numeric, smem_symbolic;
};

inline BlockPlan plan_blocks(const spl_mat *a, const spl_mat *b) {
    const double ma = (double)a->nnz / a->nmajor(), mb = (double)b->nnz / b->nmajor();
    BlockPlan best{};
    for (unsigned rows : {128u, 32u}) {
        Caps c{(uint32_t)(ma * rows * 1.25) + 64u, (uint32_t)(mb * rows * 1.25) + 64u};
        const size_t in_words = (size_t)(c.a + c.a / 32 + 1) + (c.b + c.b / 32 + 1);
        const size_t out_words = (size_t)(c.a + c.b) + (c.a + c.b) / 32 + 1;
        // indices of A, B (4 B each); per output: index (4 B) + two 16-bit source slots
        best = BlockPlan{rows, c, in_words * 4 + out_words * 8 + 16, in_words * 4 + 16};
        if (best.smem_numeric <= 64 * 1024 && c.a + c.b < 0xffffu) return best;
    }
    // very long segments: tiny stage, every CTA takes the global-memory path
    best.caps = Caps{0, 0};
    best.smem_numeric = 64;
    best.smem_symbolic = 64;
    return best;
}

template <typename T>
void fill(spl_ctx *ctx, const spl_mat *a, const spl_mat *b, spl_mat *c, int subtract, const BlockPlan &bp) {
    const unsigned grid = div_up(a->nmajor(), bp.rows);
    auto kadd = union_block_kernel<T, false, true>;
    auto ksub = union_block_kernel<T, true, true>;
    auto k = subtract ? ksub : kadd;
    SPL_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bp.smem_numeric));
    SPL_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    k<<<grid, bp.rows, bp.smem_numeric, ctx->stream>>>(
        a->nmajor(), a->ptr, a->ind, (const T *)a->val, b->ptr, b->ind, (const T *)b->val, c->ptr, nullptr,
        c->ind, (T *)c->val, bp.caps);
    check_launch(ctx, "union_fill");
}

}  // namespace

spl_mat *addsub(spl_ctx *ctx, const spl_mat *a, const spl_mat *b, int subtract) {
    SPL_REQUIRE(a->nrows == b->nrows && a->ncols == b->ncols, SPL_ERR_SHAPE,
                "add/sub: shapes differ (assert_eq!, src/csr/ops/add.rs:9-10)");
    SPL_REQUIRE(a->format == b->format && a->dtype == b->dtype, SPL_ERR_ARG,
                "add/sub: operands must share format and scalar type");
    SPL_REQUIRE((uint64_t)a->nnz + b->nnz < kMaxEntries, SPL_ERR_UNSUPPORTED,
                "add/sub: nnz(A)+nnz(B) must stay below 2^32 - 65536");
    const uint32_t nmajor = a->nmajor();
    Tmp<uint32_t> cnt(ctx, nmajor);
    Tmp<uint32_t> cptr(ctx, (size_t)nmajor + 1 + 4);      // becomes the result's pointer array: same slack as new_mat
    const BlockPlan bp = plan_blocks(a, b);
    {
        auto k = union_block_kernel<float, false, false>;       // symbolic: indices only
        SPL_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bp.smem_symbolic));
        SPL_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        k<<<div_up(nmajor, bp.rows), bp.rows, bp.smem_symbolic, ctx->stream>>>(
            nmajor, a->ptr, a->ind, nullptr, b->ptr, b->ind, nullptr, nullptr, cnt, nullptr, nullptr, bp.caps);
        check_launch(ctx, "union_count");
    }
    exclusive_scan_u32(ctx, cnt, nmajor, cptr);
    uint32_t nnz = 0;
    read_back(ctx, cptr.p + nmajor, &nnz, 1);   // exact-size output (add.rs:103-105)
    spl_mat *c = new_mat(ctx, a->format, a->dtype, a->nrows, a->ncols, nnz);
    dfree(ctx, c->ptr);
    c->ptr = cptr.release();
    try {
        if (a->dtype == SPL_F32) fill<float>(ctx, a, b, c, subtract, bp);
        else fill<double>(ctx, a, b, c, subtract, bp);
    } catch (...) {
        free_mat(ctx, c);
        throw;
    }
    return c;
}

}  // namespace spl
