// spmv.cu — y = A x for CSR A and dense x, y.
//
// The reference has no SpMV; its pinned meaning is `&A * &X` with X an n x 1 matrix
// (src/csr/ops/mul.rs:5-60): per row the column-ascending sum of round(a*x).  Parity bar
// (BASELINE.json north_star): 1e-12 (f64) / 1e-5 (f32) relative, so the in-row reduction
// order is free.  Two hand-written kernels:
//   vector  LPR lanes per row (2..32), coalesced col/val loads inside the sub-warp, shuffle
//           reduction; best when rows are long and regular.
//   merge   merge-path tiles over (row ends, nnz): every CTA takes the same number of
//           (row + nnz) items, streams its nnz with coalesced 128-bit loads into shared memory
//           as products, reduces rows from shared memory and hands partial rows to a fix-up.
//           Balanced for skewed (power-law) rows and bandwidth-efficient for very short rows.
// Algorithmic bytes per launch: nnz*(4+V) + (ncols+nrows)*V  (SURVEY.md 8d); HBM-bound.
#include "kernels.cuh"

namespace spl {

namespace {

// ------------------------------------------------------------------ vector kernel
template <typename T, int LPR>
__global__ void __launch_bounds__(256)
spmv_vector_kernel(uint32_t nrows, const uint32_t *__restrict__ ptr,
                   const uint32_t *__restrict__ ind, const T *__restrict__ val,
                   const T *__restrict__ x, T *__restrict__ y) {
    const uint64_t gtid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t row = gtid / LPR;
    const uint32_t sub = (uint32_t)(gtid % LPR);
    T acc0 = (T)0, acc1 = (T)0;
    if (row < nrows) {
        uint32_t p = __ldg(ptr + row) + sub;
        const uint32_t e = __ldg(ptr + row + 1);
        for (; p + LPR < e; p += 2 * LPR) {
            const uint32_t c0 = __ldg(ind + p), c1 = __ldg(ind + p + LPR);
            const T v0 = __ldg(val + p), v1 = __ldg(val + p + LPR);
            acc0 += v0 * __ldg(x + c0);
            acc1 += v1 * __ldg(x + c1);
        }
        if (p < e) acc0 += __ldg(val + p) * __ldg(x + __ldg(ind + p));
    }
    T acc = acc0 + acc1;
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (sub == 0 && row < nrows) y[row] = acc;
}

template <typename T, int LPR>
void launch_vector(spl_ctx *ctx, const spl_mat *a, const T *x, T *y) {
    const uint64_t threads = (uint64_t)a->nrows * LPR;
    const unsigned grid = div_up(threads, 256);
    spmv_vector_kernel<T, LPR><<<grid, 256, 0, ctx->stream>>>(a->nrows, a->ptr, a->ind,
                                                              static_cast<const T *>(a->val), x, y);
    check_launch(ctx, "spmv_vector");
}

template <typename T>
void spmv_vector(spl_ctx *ctx, const spl_mat *a, const T *x, T *y, int lanes) {
    switch (lanes) {
        case 1: launch_vector<T, 1>(ctx, a, x, y); break;
        case 2: launch_vector<T, 2>(ctx, a, x, y); break;
        case 4: launch_vector<T, 4>(ctx, a, x, y); break;
        case 8: launch_vector<T, 8>(ctx, a, x, y); break;
        case 16: launch_vector<T, 16>(ctx, a, x, y); break;
        case 32: launch_vector<T, 32>(ctx, a, x, y); break;
        default: throw Error{SPL_ERR_ARG, "lanes per row must be 1, 2, 4, 8, 16 or 32"};
    }
}

// ------------------------------------------------------------------ row statistics (plan)
__global__ void max_row_len_kernel(const uint32_t *__restrict__ ptr, uint32_t nrows, uint32_t *out) {
    uint32_t m = 0;
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < nrows;
         r += (uint64_t)gridDim.x * blockDim.x)
        m = max(m, ptr[r + 1] - ptr[r]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane_id() == 0 && m) atomicMax(out, m);
}

}  // namespace

void spmv_plan(spl_ctx *ctx, spl_mat *a) {
    if (a->plan_ready) return;
    uint32_t mx = 0;
    if (a->nnz) {
        SPL_CUDA(cudaMemsetAsync(ctx->d_scratch, 0, sizeof(uint32_t), ctx->stream));
        unsigned grid = min(div_up(a->nrows, 256), (unsigned)ctx->num_sms * 8u);
        max_row_len_kernel<<<grid, 256, 0, ctx->stream>>>(a->ptr, a->nrows, ctx->d_scratch);
        check_launch(ctx, "max_row_len");
        read_back(ctx, ctx->d_scratch, &mx, 1);
    }
    a->max_row_len = mx;
    const double mean = a->nrows ? (double)a->nnz / a->nrows : 0.0;
    int lanes = 2;
    while (lanes < 32 && lanes * 2 <= mean) lanes *= 2;   // largest power of two <= mean, >= 2
    a->plan_lanes = lanes;
    a->plan_kernel = SPL_SPMV_VECTOR;
    a->plan_ready = 1;
}

void spmv(spl_ctx *ctx, const spl_mat *a, const void *x, void *y, int kernel, int lanes) {
    SPL_REQUIRE(a->format == SPL_CSR, SPL_ERR_UNSUPPORTED, "spl_spmv needs a CSR matrix");
    if (kernel == SPL_SPMV_AUTO || lanes == 0) {
        spmv_plan(ctx, const_cast<spl_mat *>(a));
        if (kernel == SPL_SPMV_AUTO) kernel = a->plan_kernel;
        if (lanes == 0) lanes = a->plan_lanes;
    }
    if (kernel == SPL_SPMV_VECTOR) {
        if (a->dtype == SPL_F32) spmv_vector<float>(ctx, a, (const float *)x, (float *)y, lanes);
        else spmv_vector<double>(ctx, a, (const double *)x, (double *)y, lanes);
        return;
    }
    throw Error{SPL_ERR_UNSUPPORTED, "unknown SpMV kernel"};
}

}  // namespace spl
