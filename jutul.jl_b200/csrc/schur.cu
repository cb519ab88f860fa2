// MultiLinearizedSystem with reduction = :schur_apply on the device (src/linsolve/multimodel.jl:17-160; the block system is
// set up by setup_linearized_system!(::MultiModel), src/multimodel/model.jl:534-601):
//
//      [ B  C ] [x]   [a]        B : the reservoir Jacobian (jb_csr, bs x bs blocks, already resident)
//      [ D  E ] [y] = [b]        C_i, D_i, E_i : the coupling / eliminated blocks of groups i = 1..ng (wells, facility)
//
// The Krylov solver sees S = B - sum_i C_i E_i^{-1} D_i (linear_operator(sys), multimodel.jl:70-91 -> schur_mul!, :139-160),
// the right-hand side is reduced first (prepare_linear_solve!, :17-33: a -= C (E \ b)) and the eliminated unknowns are
// recovered afterwards (update_dx_from_vector! / schur_dx_update!, :97-137: y_i = E_i \ (D_i dx - b_i)).
//
// B200 design: all groups are concatenated into ONE eliminated index space of M = sum m_i dofs, so the correction is three
// small kernels whatever the number of wells: t = D x (CSR, warp per row), u = blockdiag(E_i^{-1}) t (dense rows, warp per
// row), res -= alpha C u (CSR over the rows of B that C touches, warp per row: no atomics, deterministic). E_i is factorised
// on the host with partial pivoting (the reference calls lu!/ldiv! on the host as well) and uploaded as the explicit inverse:
// the groups are small (tens to hundreds of dofs), the inverse turns the two triangular solves of every operator
// application into one bandwidth-bound dense matrix-vector product. Everything the Krylov iteration touches stays in HBM.
#include <algorithm>
#include <numeric>

#include "jb_internal.cuh"
#include "jb_reduce.cuh"

#define JB_SCHUR_MAX_GROUP 2048

struct jb_schur {
    jb_csr* B = nullptr;
    jb_ctx* ctx = nullptr;
    int ng = 0;
    i64 M = 0, nrows = 0;
    std::vector<i64> goff;                       // ng+1: first eliminated dof of each group
    // D: CSR M x nrows; C: CSR over its non-empty rows; E: dense per group
    std::vector<int32_t> h_Dptr, h_Dcol, h_Dmap; // map: COO entry -> CSR slot
    std::vector<int32_t> h_Crow, h_Cptr, h_Ccol, h_Cmap;
    std::vector<i64> h_Eptr;                     // ng+1 offsets into the COO arrays of E
    std::vector<int32_t> h_Er, h_Ec;             // local (row, col) of every COO entry of E, 0-based
    std::vector<i64> h_Einv_off;                 // ng+1 offsets into the dense inverse storage
    std::vector<double> h_Dval, h_Cval, h_Einv;
    DBuf<int32_t> d_Dptr, d_Dcol, d_Crow, d_Cptr, d_Ccol;
    DBuf<double> d_Dval, d_Cval, d_Einv;
    DBuf<i64> d_rowoff;                          // per eliminated dof: offset of its dense row in d_Einv
    DBuf<int32_t> d_rowg0, d_roww;               // per eliminated dof: first dof and width of its group
    DBuf<double> d_b, d_t, d_u;
    bool have_values = false;
};

// t[r] = sgn * sum_k D[r,k] x[col_k]  (- b[r] when sub_b)            mul!(buf, D_i, x)
__global__ void __launch_bounds__(256) schur_Dx_kernel(i64 M, const int32_t* __restrict__ ptr, const int32_t* __restrict__ col,
                                                       const double* __restrict__ val, const double* __restrict__ x, double sgn,
                                                       const double* __restrict__ b, double* __restrict__ t) {
    const i64 w = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= M) return;
    double s = 0.0;
    for (int32_t k = __ldg(ptr + w) + lane; k < __ldg(ptr + w + 1); k += 32) s = fma(__ldg(val + k), __ldg(x + __ldg(col + k)), s);
    s = warp_sum(s);
    if (lane == 0) t[w] = b ? sgn * s - __ldg(b + w) : sgn * s;
}
// u[r] = sum_j Einv[r, j] t[g0 + j]                                   ldiv!(buf, E_i, t)
__global__ void __launch_bounds__(256) schur_Einv_kernel(i64 M, const i64* __restrict__ rowoff, const int32_t* __restrict__ g0,
                                                         const int32_t* __restrict__ width, const double* __restrict__ Einv,
                                                         const double* __restrict__ t, double* __restrict__ u) {
    const i64 w = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= M) return;
    const double* row = Einv + __ldg(rowoff + w);
    const double* tv = t + __ldg(g0 + w);
    const int n = __ldg(width + w);
    double s = 0.0;
    for (int j = lane; j < n; j += 32) s = fma(__ldg(row + j), tv[j], s);
    s = warp_sum(s);
    if (lane == 0) u[w] = s;
}
// res[row] -= alpha * sum_k C[row,k] u[col_k]                         mul!(res, C_i, buf, -alpha, true)
__global__ void __launch_bounds__(256) schur_Cu_kernel(i64 nr, const int32_t* __restrict__ rows, const int32_t* __restrict__ ptr,
                                                       const int32_t* __restrict__ col, const double* __restrict__ val,
                                                       const double* __restrict__ u, double alpha, double* __restrict__ res) {
    const i64 w = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= nr) return;
    double s = 0.0;
    for (int32_t k = __ldg(ptr + w) + lane; k < __ldg(ptr + w + 1); k += 32) s = fma(__ldg(val + k), __ldg(u + __ldg(col + k)), s);
    s = warp_sum(s);
    if (lane == 0) { const int32_t r = __ldg(rows + w); res[r] = fma(-alpha, s, res[r]); }
}

static inline int wgrid(i64 warps) { return (int)std::max<i64>(1, (warps * 32 + 255) / 256); }

// res -= alpha * C (E \ (D x)), enqueue only
int jb_schur_correct_launch(jb_schur* S, const double* d_x, double* d_res, double alpha) {
    jb_ctx* ctx = S->ctx;
    if (S->M == 0) return JB_OK;
    if (!S->have_values) JB_FAIL(ctx, JB_ERR_ARG, "schur: jb_schur_update has not been called");
    ProfScope _ps(ctx, JB_PROF_SPMV);
    cudaStream_t st = ctx->stream;
    schur_Dx_kernel<<<wgrid(S->M), 256, 0, st>>>(S->M, S->d_Dptr.p, S->d_Dcol.p, S->d_Dval.p, d_x, 1.0, nullptr, S->d_t.p);
    JB_CHECK_LAUNCH(ctx);
    schur_Einv_kernel<<<wgrid(S->M), 256, 0, st>>>(S->M, S->d_rowoff.p, S->d_rowg0.p, S->d_roww.p, S->d_Einv.p, S->d_t.p, S->d_u.p);
    JB_CHECK_LAUNCH(ctx);
    const i64 nr = (i64)S->h_Crow.size();
    if (nr > 0) {
        schur_Cu_kernel<<<wgrid(nr), 256, 0, st>>>(nr, S->d_Crow.p, S->d_Cptr.p, S->d_Ccol.p, S->d_Cval.p, S->d_u.p, alpha, d_res);
        JB_CHECK_LAUNCH(ctx);
    }
    return JB_OK;
}
i64 jb_schur_nrows(jb_schur* S) { return S->nrows; }

// COO (1-based I, J over [0, nr) x [0, nc)) -> CSR with duplicates merged (sparse(I, J, V) sums them); map[k] = slot of entry k
static int coo_to_csr(jb_ctx* ctx, i64 nnz, const int64_t* I, const int64_t* J, i64 nr, i64 nc,
                      std::vector<int32_t>& ptr, std::vector<int32_t>& col, std::vector<int32_t>& map) {
    std::vector<i64> order(nnz);
    std::iota(order.begin(), order.end(), (i64)0);
    for (i64 k = 0; k < nnz; k++)
        if (I[k] < 1 || I[k] > nr || J[k] < 1 || J[k] > nc) JB_FAIL(ctx, JB_ERR_ARG, "jb_schur_create: index out of range");
    std::sort(order.begin(), order.end(), [&](i64 a, i64 b) { return I[a] != I[b] ? I[a] < I[b] : (J[a] != J[b] ? J[a] < J[b] : a < b); });
    ptr.assign((size_t)nr + 1, 0);
    col.clear();
    map.assign((size_t)nnz, 0);
    i64 prev_r = -1, prev_c = -1;
    for (i64 q = 0; q < nnz; q++) {
        const i64 k = order[q], r = I[k] - 1, c = J[k] - 1;
        if (r != prev_r || c != prev_c) { col.push_back((int32_t)c); ptr[(size_t)r + 1]++; prev_r = r; prev_c = c; }
        map[(size_t)k] = (int32_t)col.size() - 1;
    }
    for (i64 r = 0; r < nr; r++) ptr[(size_t)r + 1] += ptr[(size_t)r];
    return JB_OK;
}

// dense inverse by LU with partial pivoting (row-major n x n in A, destroyed); returns false on a zero / non-finite pivot
static bool dense_inverse(int n, std::vector<double>& A, double* inv) {
    std::vector<int> piv(n);
    for (int c = 0; c < n; c++) {
        int p = c;
        double best = std::fabs(A[(size_t)c * n + c]);
        for (int r = c + 1; r < n; r++) { const double v = std::fabs(A[(size_t)r * n + c]); if (v > best) { best = v; p = r; } }
        if (!(best > 0.0) || !std::isfinite(best)) return false;
        piv[c] = p;
        if (p != c) for (int k = 0; k < n; k++) std::swap(A[(size_t)c * n + k], A[(size_t)p * n + k]);
        const double ip = 1.0 / A[(size_t)c * n + c];
        for (int r = c + 1; r < n; r++) {
            const double f = A[(size_t)r * n + c] * ip;
            A[(size_t)r * n + c] = f;
            if (f != 0.0) for (int k = c + 1; k < n; k++) A[(size_t)r * n + k] -= f * A[(size_t)c * n + k];
        }
    }
    // solve A X = I column by column: P A = L U
    std::vector<double> y(n);
    for (int j = 0; j < n; j++) {
        for (int i = 0; i < n; i++) y[i] = (i == j) ? 1.0 : 0.0;
        for (int c = 0; c < n; c++) if (piv[c] != c) std::swap(y[c], y[piv[c]]);
        for (int i = 0; i < n; i++) { double s = y[i]; for (int k = 0; k < i; k++) s -= A[(size_t)i * n + k] * y[k]; y[i] = s; }
        for (int i = n - 1; i >= 0; i--) {
            double s = y[i];
            for (int k = i + 1; k < n; k++) s -= A[(size_t)i * n + k] * y[k];
            y[i] = s / A[(size_t)i * n + i];
        }
        for (int i = 0; i < n; i++) inv[(size_t)i * n + j] = y[i];
    }
    return true;
}

extern "C" {

int32_t jb_schur_create(jb_csr* B, int32_t ngroups, const int64_t* msize, const int64_t* C_ptr, const int64_t* C_I, const int64_t* C_J,
                        const int64_t* D_ptr, const int64_t* D_I, const int64_t* D_J, const int64_t* E_ptr, const int64_t* E_I,
                        const int64_t* E_J, jb_schur** out) {
    if (!B || !out || ngroups < 0 || (ngroups > 0 && (!msize || !C_ptr || !D_ptr || !E_ptr))) return JB_ERR_ARG;
    jb_ctx* ctx = B->ctx;
    jb_schur* S = new jb_schur();
    S->B = B; S->ctx = ctx; S->ng = ngroups; S->nrows = B->n * B->bs;
    S->goff.assign((size_t)ngroups + 1, 0);
    S->h_Einv_off.assign((size_t)ngroups + 1, 0);
    for (int g = 0; g < ngroups; g++) {
        if (msize[g] < 1 || msize[g] > JB_SCHUR_MAX_GROUP) {
            delete S;
            JB_FAIL(ctx, JB_ERR_UNSUPPORTED, "jb_schur_create: eliminated group size must be in 1..2048 (dense inverse of E_i)");
        }
        S->goff[g + 1] = S->goff[g] + msize[g];
        S->h_Einv_off[g + 1] = S->h_Einv_off[g] + msize[g] * msize[g];
    }
    S->M = S->goff[ngroups];
    const i64 M = S->M;
    if (ngroups > 0 && (C_ptr[0] != 1 || D_ptr[0] != 1 || E_ptr[0] != 1)) { delete S; JB_FAIL(ctx, JB_ERR_ARG, "jb_schur_create: *_ptr must start at 1"); }
    for (int g = 0; g < ngroups; g++)
        if (C_ptr[g + 1] < C_ptr[g] || D_ptr[g + 1] < D_ptr[g] || E_ptr[g + 1] < E_ptr[g]) {
            delete S;
            JB_FAIL(ctx, JB_ERR_ARG, "jb_schur_create: *_ptr must be non-decreasing");
        }
    // concatenate the groups: C columns and D rows shifted into the common eliminated index space
    const i64 nC = ngroups ? C_ptr[ngroups] - 1 : 0, nD = ngroups ? D_ptr[ngroups] - 1 : 0, nE = ngroups ? E_ptr[ngroups] - 1 : 0;
    std::vector<int64_t> cI(nC), cJ(nC), dI(nD), dJ(nD);
    S->h_Er.resize(nE); S->h_Ec.resize(nE);
    S->h_Eptr.assign(E_ptr, E_ptr + ngroups + 1);
    for (auto& v : S->h_Eptr) v -= 1;
    for (int g = 0; g < ngroups; g++) {
        const i64 mg = msize[g];
        for (i64 k = C_ptr[g] - 1; k < C_ptr[g + 1] - 1; k++) {
            if (C_J[k] < 1 || C_J[k] > mg) { delete S; JB_FAIL(ctx, JB_ERR_ARG, "jb_schur_create: C column out of range"); }
            cI[k] = C_I[k]; cJ[k] = C_J[k] + S->goff[g];
        }
        for (i64 k = D_ptr[g] - 1; k < D_ptr[g + 1] - 1; k++) {
            if (D_I[k] < 1 || D_I[k] > mg) { delete S; JB_FAIL(ctx, JB_ERR_ARG, "jb_schur_create: D row out of range"); }
            dI[k] = D_I[k] + S->goff[g]; dJ[k] = D_J[k];
        }
        for (i64 k = E_ptr[g] - 1; k < E_ptr[g + 1] - 1; k++) {
            if (E_I[k] < 1 || E_I[k] > mg || E_J[k] < 1 || E_J[k] > mg) { delete S; JB_FAIL(ctx, JB_ERR_ARG, "jb_schur_create: E index out of range"); }
            S->h_Er[k] = (int32_t)(E_I[k] - 1); S->h_Ec[k] = (int32_t)(E_J[k] - 1);
        }
    }
    int rc = coo_to_csr(ctx, nD, dI.data(), dJ.data(), M, S->nrows, S->h_Dptr, S->h_Dcol, S->h_Dmap);
    if (rc != JB_OK) { delete S; return rc; }
    std::vector<int32_t> fullptr;
    rc = coo_to_csr(ctx, nC, cI.data(), cJ.data(), S->nrows, M, fullptr, S->h_Ccol, S->h_Cmap);
    if (rc != JB_OK) { delete S; return rc; }
    S->h_Cptr.assign(1, 0);
    for (i64 r = 0; r < S->nrows; r++)
        if (fullptr[(size_t)r + 1] > fullptr[(size_t)r]) { S->h_Crow.push_back((int32_t)r); S->h_Cptr.push_back(fullptr[(size_t)r + 1]); }
    S->h_Dval.assign(S->h_Dcol.size(), 0.0);
    S->h_Cval.assign(S->h_Ccol.size(), 0.0);
    S->h_Einv.assign((size_t)S->h_Einv_off[ngroups], 0.0);
    std::vector<i64> rowoff((size_t)M);
    std::vector<int32_t> rowg0((size_t)M), roww((size_t)M);
    for (int g = 0; g < ngroups; g++)
        for (i64 r = 0; r < msize[g]; r++) {
            const size_t q = (size_t)(S->goff[g] + r);
            rowoff[q] = S->h_Einv_off[g] + r * msize[g]; rowg0[q] = (int32_t)S->goff[g]; roww[q] = (int32_t)msize[g];
        }
    cudaStream_t st = ctx->stream;
    bool ok = S->d_Dptr.upload(S->h_Dptr, st) == cudaSuccess && S->d_Dcol.upload(S->h_Dcol, st) == cudaSuccess &&
              S->d_Crow.upload(S->h_Crow, st) == cudaSuccess && S->d_Cptr.upload(S->h_Cptr, st) == cudaSuccess &&
              S->d_Ccol.upload(S->h_Ccol, st) == cudaSuccess && S->d_rowoff.upload(rowoff, st) == cudaSuccess &&
              S->d_rowg0.upload(rowg0, st) == cudaSuccess && S->d_roww.upload(roww, st) == cudaSuccess &&
              S->d_Dval.alloc(S->h_Dval.size()) == cudaSuccess && S->d_Cval.alloc(S->h_Cval.size()) == cudaSuccess &&
              S->d_Einv.alloc(S->h_Einv.size()) == cudaSuccess && S->d_b.alloc((size_t)M) == cudaSuccess &&
              S->d_t.alloc((size_t)M) == cudaSuccess && S->d_u.alloc((size_t)M) == cudaSuccess;
    if (!ok) { delete S; JB_FAIL(ctx, JB_ERR_ALLOC, "jb_schur_create: allocation failed"); }
    if (M > 0) cudaMemsetAsync(S->d_b.p, 0, (size_t)M * sizeof(double), st);
    cudaStreamSynchronize(st);
    *out = S;
    return JB_OK;
}

int32_t jb_schur_destroy(jb_schur* S) { delete S; return JB_OK; }
int64_t jb_schur_size(jb_schur* S) { return S ? S->M : 0; }

// new values of the coupling blocks (COO order of jb_schur_create, duplicates summed) and factorisation of every E_i:
// get_schur_blocks!(sys, update = true) -> update_preconditioner!(F, lu, lu!, E) (multimodel.jl:36-54)
int32_t jb_schur_update(jb_schur* S, const double* C_V, const double* D_V, const double* E_V) {
    if (!S) return JB_ERR_ARG;
    jb_ctx* ctx = S->ctx;
    if (S->M == 0) { S->have_values = true; return JB_OK; }
    if (!C_V || !D_V || !E_V) return JB_ERR_ARG;
    std::fill(S->h_Cval.begin(), S->h_Cval.end(), 0.0);
    std::fill(S->h_Dval.begin(), S->h_Dval.end(), 0.0);
    for (size_t k = 0; k < S->h_Cmap.size(); k++) S->h_Cval[(size_t)S->h_Cmap[k]] += C_V[k];
    for (size_t k = 0; k < S->h_Dmap.size(); k++) S->h_Dval[(size_t)S->h_Dmap[k]] += D_V[k];
    std::vector<double> dense;
    for (int g = 0; g < S->ng; g++) {
        const int mg = (int)(S->goff[g + 1] - S->goff[g]);
        dense.assign((size_t)mg * mg, 0.0);
        for (i64 k = S->h_Eptr[g]; k < S->h_Eptr[g + 1]; k++) dense[(size_t)S->h_Er[k] * mg + S->h_Ec[k]] += E_V[k];
        if (!dense_inverse(mg, dense, S->h_Einv.data() + S->h_Einv_off[g])) {
            S->have_values = false;
            JB_FAIL(ctx, JB_BAD_PIVOT, "jb_schur_update: singular eliminated block E_" + std::to_string(g + 1));
        }
    }
    cudaStream_t st = ctx->stream;
    if (!S->h_Cval.empty()) JB_CUDA(ctx, cudaMemcpyAsync(S->d_Cval.p, S->h_Cval.data(), S->h_Cval.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    if (!S->h_Dval.empty()) JB_CUDA(ctx, cudaMemcpyAsync(S->d_Dval.p, S->h_Dval.data(), S->h_Dval.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    JB_CUDA(ctx, cudaMemcpyAsync(S->d_Einv.p, S->h_Einv.data(), S->h_Einv.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    JB_CUDA(ctx, cudaStreamSynchronize(st));
    S->have_values = true;
    return JB_OK;
}

// prepare_linear_solve! (multimodel.jl:17-33): a -= C (E \ b); b (host, M) stays resident for jb_schur_dx_update
int32_t jb_schur_prepare(jb_schur* S, double* d_a, const double* b) {
    if (!S || !d_a) return JB_ERR_ARG;
    jb_ctx* ctx = S->ctx;
    if (S->M == 0) return JB_OK;
    if (!b) return JB_ERR_ARG;
    if (!S->have_values) JB_FAIL(ctx, JB_ERR_ARG, "jb_schur_prepare: jb_schur_update has not been called");
    cudaStream_t st = ctx->stream;
    JB_CUDA(ctx, cudaMemcpyAsync(S->d_b.p, b, (size_t)S->M * sizeof(double), cudaMemcpyHostToDevice, st));
    {
        ProfScope _ps(ctx, JB_PROF_OTHER);
        schur_Einv_kernel<<<wgrid(S->M), 256, 0, st>>>(S->M, S->d_rowoff.p, S->d_rowg0.p, S->d_roww.p, S->d_Einv.p, S->d_b.p, S->d_u.p);
        JB_CHECK_LAUNCH(ctx);
        const i64 nr = (i64)S->h_Crow.size();
        if (nr > 0) {
            schur_Cu_kernel<<<wgrid(nr), 256, 0, st>>>(nr, S->d_Crow.p, S->d_Cptr.p, S->d_Ccol.p, S->d_Cval.p, S->d_u.p, 1.0, d_a);
            JB_CHECK_LAUNCH(ctx);
        }
    }
    JB_CUDA(ctx, cudaStreamSynchronize(st));
    return JB_OK;
}

// schur_mul! (multimodel.jl:150-160): res <- beta res + alpha (B x - C (E \ (D x)))
int32_t jb_schur_mul(jb_schur* S, double alpha, const double* d_x, double beta, double* d_res) {
    if (!S || !d_x || !d_res) return JB_ERR_ARG;
    int rc = jb_launch_spmv(S->B, alpha, d_x, beta, d_res);
    if (rc != JB_OK) return rc;
    rc = jb_schur_correct_launch(S, d_x, d_res, alpha);
    if (rc != JB_OK) return rc;
    JB_CUDA(S->ctx, cudaStreamSynchronize(S->ctx->stream));
    return JB_OK;
}

// update_dx_from_vector! / schur_dx_update! (multimodel.jl:97-137). d_dx is what jb_krylov_solve left: dx = -x, x the
// solver's solution (= dx_from_solver); y_i = E_i \ (D_i x - b_i) goes to the host array y (M).
int32_t jb_schur_dx_update(jb_schur* S, const double* d_dx, double* y) {
    if (!S || !d_dx) return JB_ERR_ARG;
    jb_ctx* ctx = S->ctx;
    if (S->M == 0) return JB_OK;
    if (!y) return JB_ERR_ARG;
    if (!S->have_values) JB_FAIL(ctx, JB_ERR_ARG, "jb_schur_dx_update: jb_schur_update has not been called");
    cudaStream_t st = ctx->stream;
    {
        ProfScope _ps(ctx, JB_PROF_OTHER);
        schur_Dx_kernel<<<wgrid(S->M), 256, 0, st>>>(S->M, S->d_Dptr.p, S->d_Dcol.p, S->d_Dval.p, d_dx, -1.0, S->d_b.p, S->d_t.p);
        JB_CHECK_LAUNCH(ctx);
        schur_Einv_kernel<<<wgrid(S->M), 256, 0, st>>>(S->M, S->d_rowoff.p, S->d_rowg0.p, S->d_roww.p, S->d_Einv.p, S->d_t.p, S->d_u.p);
        JB_CHECK_LAUNCH(ctx);
    }
    JB_CUDA(ctx, cudaMemcpyAsync(y, S->d_u.p, (size_t)S->M * sizeof(double), cudaMemcpyDeviceToHost, st));
    JB_CUDA(ctx, cudaStreamSynchronize(st));
    return JB_OK;
}

}  // extern "C"
