// Diagonal preconditioners of the reference (src/linsolve/precond/diagonal.jl, jacobi.jl, spai.jl) behind the same
// handle type as ILU(0), so that jb_krylov_create / jb_krylov_solve accept them unchanged:
//   Jacobi   D_i = w * inv(A_ii)                                  (jacobi.jl:15-18, default w = 2/3)
//   SPAI(0)  D_i = inv(sum_k |A_ik|_F^2) * A_ii  over the row     (spai.jl:42-62, StaticSparsityMatrixCSR method)
//   apply!   x_i = D_i * y_i per block                            (diagonal.jl:31-49)
// One thread per block row; both kernels are single streaming passes (HBM-bound).
#include "jb_internal.cuh"
#include "jb_krylov_scalars.cuh"

template <int BS>
__global__ void __launch_bounds__(256) diag_factor_kernel(i64 n, int kind, double w, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ diag,
                                                          const double* __restrict__ val, double* __restrict__ dinv, int32_t* status) {
    constexpr int B2 = BS * BS;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        double Aii[B2], D[B2];
        const size_t dk = (size_t)__ldg(diag + i);
#pragma unroll
        for (int q = 0; q < B2; q++) Aii[q] = __ldg(val + dk * B2 + q);
        if (kind == 1) {
            blk_inv<BS>(Aii, D);
#pragma unroll
            for (int q = 0; q < B2; q++) D[q] = w * D[q];
        } else {
            double norm_sum = 0.0;
            for (int32_t k = __ldg(rowptr + i); k < __ldg(rowptr + i + 1); k++)
#pragma unroll
                for (int q = 0; q < B2; q++) { const double v = __ldg(val + (size_t)k * B2 + q); norm_sum += v * v; }
            const double inv_ns = 1.0 / norm_sum;
#pragma unroll
            for (int q = 0; q < B2; q++) D[q] = inv_ns * Aii[q];
        }
        bool bad = false;
#pragma unroll
        for (int q = 0; q < B2; q++) { dinv[(size_t)i * B2 + q] = D[q]; bad |= !isfinite(D[q]); }
        if (bad) *status = JB_BAD_PIVOT;
    }
}

template <int BS>
__global__ void __launch_bounds__(256) diag_apply_kernel(i64 n, const double* __restrict__ dinv, const double* b, double* x, const double* sc) {
    constexpr int B2 = BS * BS;
    if (sc && sc[KS_DONE] != 0.0) return;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        double d[B2], v[BS], out[BS];
#pragma unroll
        for (int q = 0; q < B2; q++) d[q] = __ldg(dinv + (size_t)i * B2 + q);
#pragma unroll
        for (int e = 0; e < BS; e++) v[e] = b[(size_t)i * BS + e];
        blk_mulvec<BS>(d, v, out);
#pragma unroll
        for (int e = 0; e < BS; e++) x[(size_t)i * BS + e] = out[e];
    }
}

static int dgrid(jb_ctx* ctx, i64 n) { return (int)std::max<i64>(1, std::min<i64>((n + 255) / 256, (i64)ctx->sm_count * 8)); }

int jb_launch_diag_factor(jb_ilu* F) {
    jb_csr* A = F->csr;
    jb_ctx* ctx = A->ctx;
    ProfScope _ps(ctx, JB_PROF_ILU_FACTOR);
    JB_CUDA(ctx, cudaMemsetAsync(F->d_status.p, 0, sizeof(int32_t), ctx->stream));
    const int g = dgrid(ctx, F->n);
#define JB_DF(BS) diag_factor_kernel<BS><<<g, 256, 0, ctx->stream>>>(F->n, F->diag_kind, F->diag_w, A->d_rowptr.p, A->d_diag.p, A->d_val.p, F->d_dinv.p, F->d_status.p)
    switch (F->bs) {
        case 1: JB_DF(1); break;
        case 2: JB_DF(2); break;
        case 3: JB_DF(3); break;
        case 4: JB_DF(4); break;
        default: return JB_ERR_UNSUPPORTED;
    }
#undef JB_DF
    JB_CHECK_LAUNCH(ctx);
    return JB_OK;
}
int jb_launch_diag_apply(jb_ilu* F, const double* d_b, double* d_x, const double* d_sc) {
    jb_ctx* ctx = F->csr->ctx;
    ProfScope _ps(ctx, JB_PROF_ILU_APPLY);
    const int g = dgrid(ctx, F->n);
#define JB_DA(BS) diag_apply_kernel<BS><<<g, 256, 0, ctx->stream>>>(F->n, F->d_dinv.p, d_b, d_x, d_sc)
    switch (F->bs) {
        case 1: JB_DA(1); break;
        case 2: JB_DA(2); break;
        case 3: JB_DA(3); break;
        case 4: JB_DA(4); break;
        default: return JB_ERR_UNSUPPORTED;
    }
#undef JB_DA
    JB_CHECK_LAUNCH(ctx);
    return JB_OK;
}

extern "C" int32_t jb_diag_precond_create(jb_csr* A, int32_t kind, double w, jb_ilu** out) {
    if (!A || !out || (kind != 1 && kind != 2)) return JB_ERR_ARG;
    jb_ctx* ctx = A->ctx;
    for (i64 r = 0; r < A->n; r++)
        if (A->h_diag[r] < 0) JB_FAIL(ctx, JB_ERR_ARG, "jb_diag_precond_create: diagonal must be present in the sparsity pattern");
    jb_ilu* F = new jb_ilu();
    F->csr = A; F->n = A->n; F->bs = A->bs; F->nL = F->nU = 0; F->nlevF = F->nlevB = 0;
    F->diag_kind = kind; F->diag_w = w;
    if (F->d_dinv.alloc((size_t)std::max<i64>(A->n, 1) * A->bs * A->bs) != cudaSuccess || F->d_status.alloc(1) != cudaSuccess) {
        delete F; JB_FAIL(ctx, JB_ERR_ALLOC, "jb_diag_precond_create: allocation failed");
    }
    A->n_ident_chunks = -1;
    *out = F;
    return JB_OK;
}
