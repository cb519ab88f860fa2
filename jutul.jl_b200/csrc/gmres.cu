// GMRES for sm_100a in the operation order of Krylov.jl 0.9 `gmres!` (the GenericKrylov() default solver,
// src/linsolve/krylov.jl:43,212-238): x0 = 0, Arnoldi with modified Gram-Schmidt, Givens rotations, memory 20 with
// restart = false (Jutul's serial call — the basis grows past `memory`), right (N) / left (M) / no preconditioning.
//
// Device-resident like the BiCGStab: the Hessenberg column, the Givens recurrences, the residual estimate and the
// termination flags live in device memory and are updated by the last CTA of the reducing kernel. One modified
// Gram-Schmidt step is ONE pass over q: the kernel subtracts h_i v_i (h_i produced by the previous pass) and, in the same
// sweep, accumulates the next inner product <v_{i+1}, q> (or <q, q> on the last step), so Arnoldi step j costs j + 1
// passes instead of 2j + 1.
#include "jb_internal.cuh"
#include "jb_krylov_scalars.cuh"
#include "jb_reduce.cuh"

int jb_launch_ilu_apply_sc(jb_ilu* F, const double* d_b, double* d_x, const double* d_sc);

enum { GM_H = 20, GM_HBIS = 21, GM_BETA = 22, GM_INNER = 23, GM_BASE = 24 };   // extra slots of the scalar block
// GM_INNER: Arnoldi steps of the current pass, GM_BASE: iterations of the earlier passes (restart = true); KS_ITER = GM_BASE + GM_INNER

__device__ __forceinline__ double gm_sign(double x) { return (double)((x > 0.0) - (x < 0.0)); }
__device__ __forceinline__ void gm_sym_givens(double a, double b, double& c, double& s, double& rho) {
    if (b == 0.0) { c = (a == 0.0) ? 1.0 : gm_sign(a); s = 0.0; rho = fabs(a); }
    else if (a == 0.0) { c = 0.0; s = gm_sign(b); rho = fabs(b); }
    else if (fabs(b) > fabs(a)) { const double t = a / b; s = gm_sign(b) / sqrt(1.0 + t * t); c = s * t; rho = b / s; }
    else { const double t = b / a; c = gm_sign(a) / sqrt(1.0 + t * t); s = c * t; rho = a / c; }
}

// r0 -> |r0|; V1 is formed by gm_scale_kernel
__global__ void __launch_bounds__(256) gm_init_kernel(i64 m, const double* __restrict__ r0, double* sc, double* hist, double* z, double* partials,
                                                      unsigned int* counter) {
    double d[1] = {0.0};
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (i64)gridDim.x * blockDim.x) d[0] = fma(__ldg(r0 + i), __ldg(r0 + i), d[0]);
    grid_reduce<1, OpSum>(d, partials, counter, [=](double(&t)[1]) {
        const double beta = sqrt(t[0]);
        sc[GM_BETA] = beta; sc[KS_RNORM] = beta; sc[KS_R0] = beta;
        sc[KS_EPS] = sc[KS_ATOL] + sc[KS_RTOL] * beta;
        sc[KS_ITER] = 0.0; sc[GM_INNER] = 0.0; sc[GM_BASE] = 0.0;
        hist[0] = beta; z[0] = beta;
        double done = 0.0, status = 1.0;
        if (beta == 0.0 || beta <= sc[KS_EPS]) { done = 1.0; status = 0.0; }
        else if (sc[KS_ITMAX] <= 0.0) { done = 1.0; status = 1.0; }
        sc[KS_DONE] = done; sc[KS_STATUS] = status;
    });
}
// v = src / sc[slot]
__global__ void __launch_bounds__(256) gm_scale_kernel(i64 m, const double* sc, int slot, const double* __restrict__ src, double* __restrict__ v) {
    if (sc[KS_DONE] != 0.0) return;
    const double a = sc[slot];
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (i64)gridDim.x * blockDim.x) v[i] = __ldg(src + i) / a;
}
// first inner product of an Arnoldi step: h_1 = <v_1, q>
__global__ void __launch_bounds__(256) gm_dot_kernel(i64 m, double* sc, const double* __restrict__ v, const double* __restrict__ q, double* R, int pos,
                                                     double* partials, unsigned int* counter) {
    if (sc[KS_DONE] != 0.0) return;
    double d[1] = {0.0};
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (i64)gridDim.x * blockDim.x) d[0] = fma(__ldg(v + i), __ldg(q + i), d[0]);
    grid_reduce<1, OpSum>(d, partials, counter, [=](double(&t)[1]) { sc[GM_H] = t[0]; R[pos] = t[0]; });
}
// q -= h_i v_i ; then <v_next, q> (i < j) or <q, q> + Givens update (i == j)
__global__ void __launch_bounds__(256) gm_mgs_kernel(i64 m, double* sc, const double* __restrict__ vi, const double* __restrict__ vnext, double* __restrict__ q,
                                                     double* R, double* gc, double* gs, double* z, double* hist, int hist_cap, int nr, int i, int j,
                                                     int base, double* partials, unsigned int* counter) {
    if (sc[KS_DONE] != 0.0) return;
    const double h = sc[GM_H];
    const bool last = (i == j);
    double d[1] = {0.0};
    for (i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x; k < m; k += (i64)gridDim.x * blockDim.x) {
        const double qk = fma(-h, __ldg(vi + k), q[k]);
        q[k] = qk;
        d[0] = fma(last ? qk : __ldg(vnext + k), qk, d[0]);
    }
    grid_reduce<1, OpSum>(d, partials, counter, [=](double(&t)[1]) {
        if (!last) { sc[GM_H] = t[0]; R[nr + i] = t[0]; return; }     // h_{i+1,j}
        // end of Arnoldi step j: h_{j+1,j}, Givens rotations, residual estimate, termination (Krylov.jl gmres!)
        const double Hbis = sqrt(t[0]);
        for (int k = 1; k <= j - 1; k++) {
            const double Rtmp = gc[k - 1] * R[nr + k - 1] + gs[k - 1] * R[nr + k];
            R[nr + k] = gs[k - 1] * R[nr + k - 1] - gc[k - 1] * R[nr + k];
            R[nr + k - 1] = Rtmp;
        }
        double c, s, rho;
        gm_sym_givens(R[nr + j - 1], Hbis, c, s, rho);
        gc[j - 1] = c; gs[j - 1] = s; R[nr + j - 1] = rho;
        const double zeta = s * z[j - 1];
        z[j - 1] = c * z[j - 1];
        const double rnorm = fabs(zeta);
        sc[KS_RNORM] = rnorm; sc[KS_ITER] = (double)(base + j); sc[GM_INNER] = (double)j; sc[GM_HBIS] = Hbis;
        if (base + j < hist_cap) hist[base + j] = rnorm;
        const bool mach = (rnorm + 1.0 <= 1.0);
        const bool solved = (rnorm <= sc[KS_EPS]) || mach;
        const bool breakdown = Hbis <= 1.8189894035458565e-12;       // eps(Float64)^(3/4)
        const bool tired = (double)(base + j) >= sc[KS_ITMAX];
        if (solved) { sc[KS_DONE] = 1.0; sc[KS_STATUS] = 0.0; }
        else if (breakdown) { sc[KS_DONE] = 1.0; sc[KS_STATUS] = 2.0; }
        else if (tired) { sc[KS_DONE] = 1.0; sc[KS_STATUS] = 1.0; }
        else z[j] = zeta;
    });
}
// R y = z by back substitution (packed columns), one thread
__global__ void gm_backsolve_kernel(const double* sc, const double* R, const double* z, double* y) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int inner = (int)sc[GM_INNER];
    const int nr = inner * (inner + 1) / 2;
    for (int i = 0; i < inner; i++) y[i] = z[i];
    for (int i = inner; i >= 1; i--) {
        int pos = nr + i - inner - 1;
        for (int j = inner; j >= i + 1; j--) { y[i - 1] -= R[pos] * y[j - 1]; pos = pos - j + 1; }
        y[i - 1] = (fabs(R[pos]) <= 1.8189894035458565e-12) ? 0.0 : y[i - 1] / R[pos];
    }
}
// acc = sum_i y_i v_i
__global__ void __launch_bounds__(256) gm_combine_kernel(i64 m, const double* sc, const double* __restrict__ y, const double* const* __restrict__ V,
                                                         double* __restrict__ acc) {
    const int inner = (int)sc[GM_INNER];
    for (i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x; k < m; k += (i64)gridDim.x * blockDim.x) {
        double a = 0.0;
        for (int i = 0; i < inner; i++) a = fma(__ldg(y + i), __ldg(V[i] + k), a);
        acc[k] = a;
    }
}
__global__ void __launch_bounds__(256) gm_negate_kernel(i64 m, const double* __restrict__ x, double* __restrict__ dx) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (i64)gridDim.x * blockDim.x) dx[i] = -__ldg(x + i);
}

// restart: r0 = b - A x (q holds A x); then |r0| -> beta, z[0]; the pass counters move on
__global__ void __launch_bounds__(256) gm_residual_kernel(i64 m, const double* __restrict__ b, const double* __restrict__ q, double* __restrict__ r) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (i64)gridDim.x * blockDim.x) r[i] = __ldg(b + i) - __ldg(q + i);
}
__global__ void __launch_bounds__(256) gm_reinit_kernel(i64 m, const double* __restrict__ r0, double* sc, double* z, double* partials,
                                                        unsigned int* counter) {
    double d[1] = {0.0};
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (i64)gridDim.x * blockDim.x) d[0] = fma(__ldg(r0 + i), __ldg(r0 + i), d[0]);
    grid_reduce<1, OpSum>(d, partials, counter, [=](double(&t)[1]) {
        const double beta = sqrt(t[0]);
        sc[GM_BETA] = beta; sc[KS_RNORM] = beta;
        sc[GM_BASE] = sc[KS_ITER]; sc[GM_INNER] = 0.0;
        z[0] = beta;
        if (beta == 0.0) { sc[KS_DONE] = 1.0; sc[KS_STATUS] = 0.0; }
    });
}
// x = pass (first pass) or x += pass
__global__ void __launch_bounds__(256) gm_accumulate_kernel(i64 m, const double* __restrict__ pass, double* __restrict__ x, int first) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (i64)gridDim.x * blockDim.x) x[i] = first ? __ldg(pass + i) : x[i] + __ldg(pass + i);
}

static int vgrid(jb_ctx* ctx, i64 m) {
    return (int)std::max<i64>(1, std::min<i64>((m + 255) / 256, std::min<i64>((i64)ctx->sm_count * 8, JB_MAX_PARTIALS)));
}

static int gm_ensure_basis(jb_krylov* K, int count) {
    jb_ctx* ctx = K->csr->ctx;
    bool grew = false;
    while ((int)K->gm_V.size() < count) {
        double* p = nullptr;
        if (cudaMalloc((void**)&p, (size_t)K->m * sizeof(double)) != cudaSuccess) JB_FAIL(ctx, JB_ERR_ALLOC, "gmres: basis allocation failed");
        K->gm_V.push_back(p);
        grew = true;
    }
    if (grew) {
        if (K->gm_Vptr.alloc(K->gm_V.size() + 64) != cudaSuccess) JB_FAIL(ctx, JB_ERR_ALLOC, "gmres: pointer table allocation failed");
        JB_CUDA(ctx, cudaMemcpyAsync(K->gm_Vptr.p, K->gm_V.data(), K->gm_V.size() * sizeof(double*), cudaMemcpyHostToDevice, ctx->stream));
        JB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return JB_OK;
}

static int gm_ensure_z(jb_krylov* K, int count) {
    jb_ctx* ctx = K->csr->ctx;
    bool grew = false;
    while ((int)K->gm_Z.size() < count) {
        double* p = nullptr;
        if (cudaMalloc((void**)&p, (size_t)K->m * sizeof(double)) != cudaSuccess) JB_FAIL(ctx, JB_ERR_ALLOC, "fgmres: basis allocation failed");
        K->gm_Z.push_back(p);
        grew = true;
    }
    if (grew) {
        if (K->gm_Zptr.alloc(K->gm_Z.size() + 64) != cudaSuccess) JB_FAIL(ctx, JB_ERR_ALLOC, "fgmres: pointer table allocation failed");
        JB_CUDA(ctx, cudaMemcpyAsync(K->gm_Zptr.p, K->gm_Z.data(), K->gm_Z.size() * sizeof(double*), cudaMemcpyHostToDevice, ctx->stream));
        JB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return JB_OK;
}

int jb_gmres_solve_impl(jb_krylov* K, const double* d_b, double* d_dx, double rtol, double atol, int itmax, int side, int* iters, double* hist,
                        int hist_cap, int* status_out) {
    jb_csr* A = K->csr;
    jb_ctx* ctx = A->ctx;
    cudaStream_t st = ctx->stream;
    if (K->dist) JB_FAIL(ctx, JB_ERR_UNSUPPORTED, "gmres: the distributed solve uses bicgstab in this build");
    const i64 m = K->m;
    const int g = vgrid(ctx, m);
    jb_ilu* F = K->ilu;
    const bool right = (side == 0 && F), left = (side == 1 && F);
    const int mem = std::max(1, K->gm_memory);
    const bool restart = K->gm_restart;
    const bool flex = K->gm_flexible && right;      // fgmres!: x = x0 + Z y with Z_j = N^{-1} v_j kept (Krylov.jl fgmres!)
    if (itmax > K->hist_cap - 2) itmax = K->hist_cap - 2;
    double* sc = K->d_sc.p;
    const int cols = restart ? std::min(mem, std::max(itmax, 1)) : itmax;     // Arnoldi steps a pass can take
    const size_t rsize = (size_t)(cols + 2) * (cols + 3) / 2;
    if (K->gm_R.n < rsize && K->gm_R.alloc(rsize) != cudaSuccess) JB_FAIL(ctx, JB_ERR_ALLOC, "gmres: allocation failed");
    if (K->gm_cs.n < (size_t)4 * (cols + 2) && K->gm_cs.alloc((size_t)4 * (cols + 2)) != cudaSuccess) JB_FAIL(ctx, JB_ERR_ALLOC, "gmres: allocation failed");
    double* gc = K->gm_cs.p; double* gs = gc + (cols + 2); double* z = gs + (cols + 2); double* y = z + (cols + 2);
    int rc = gm_ensure_basis(K, std::min(cols, mem) + 1);     // memory = 20; grows on demand when restart = false
    if (rc != JB_OK) return rc;

    double h_sc[KS_SIZE];
    memset(h_sc, 0, sizeof(h_sc));
    h_sc[KS_ATOL] = atol; h_sc[KS_RTOL] = rtol; h_sc[KS_ITMAX] = (double)itmax;
    memcpy(K->h_flags + 2 * KS_SIZE, h_sc, sizeof(h_sc));
    JB_CUDA(ctx, cudaMemcpyAsync(sc, K->h_flags + 2 * KS_SIZE, sizeof(h_sc), cudaMemcpyHostToDevice, st));
    JB_CUDA(ctx, cudaMemsetAsync(K->gm_cs.p, 0, K->gm_cs.n * sizeof(double), st));

#define GM_BEGIN if (ctx->prof_on) jb_prof_begin(ctx, JB_PROF_VECTOR);
#define GM_END if (ctx->prof_on) jb_prof_end(ctx);
    const double* r0 = d_b;
    if (left) { rc = jb_launch_ilu_apply_sc(F, d_b, K->r.p, nullptr); if (rc != JB_OK) return rc; r0 = K->r.p; }
    GM_BEGIN
    gm_init_kernel<<<g, 256, 0, st>>>(m, r0, sc, K->d_hist.p, z, ctx->d_partials, ctx->d_counters); JB_CHECK_LAUNCH(ctx);
    gm_scale_kernel<<<g, 256, 0, st>>>(m, sc, GM_BETA, r0, K->gm_V[0]); JB_CHECK_LAUNCH(ctx);
    GM_END
    double* q = K->q.p;
    int base = 0;          // iterations of the completed passes
    bool first_pass = true;
    for (;;) {
        const int inner_max = restart ? std::min(mem, itmax - base) : itmax - base;
        JB_CUDA(ctx, cudaMemcpyAsync(K->h_flags, sc, KS_SIZE * sizeof(double), cudaMemcpyDeviceToHost, st));
        JB_CUDA(ctx, cudaEventRecord(K->ev[0], st));
        for (int j = 1; j <= inner_max; j++) {
            if (j >= 2) {   // one Arnoldi step in flight ahead of the host
                const int slot_prev = j & 1;
                JB_CUDA(ctx, cudaEventSynchronize(K->ev[slot_prev]));
                if (K->h_flags[slot_prev * KS_SIZE + KS_DONE] != 0.0) break;
            }
            if ((rc = gm_ensure_basis(K, j + 1)) != JB_OK) return rc;
            if (flex && (rc = gm_ensure_z(K, j)) != JB_OK) return rc;
            const int nr = j * (j - 1) / 2;
            const double* pv = K->gm_V[j - 1];
            if (right) {
                double* dst = flex ? K->gm_Z[j - 1] : K->p.p;
                rc = jb_launch_ilu_apply_sc(F, K->gm_V[j - 1], dst, sc); if (rc != JB_OK) return rc;
                pv = dst;
            }
            if (!left) {
                rc = jb_launch_spmv_dots(A, pv, q, JB_DOT_NONE, nullptr, nullptr); if (rc != JB_OK) return rc;
                if (K->schur && (rc = jb_schur_correct_launch(K->schur, pv, q, 1.0)) != JB_OK) return rc;   // q = S pv (multimodel.jl:139-160)
            } else {
                rc = jb_launch_spmv_dots(A, pv, K->t.p, JB_DOT_NONE, nullptr, nullptr); if (rc != JB_OK) return rc;
                if (K->schur && (rc = jb_schur_correct_launch(K->schur, pv, K->t.p, 1.0)) != JB_OK) return rc;
                rc = jb_launch_ilu_apply_sc(F, K->t.p, q, sc); if (rc != JB_OK) return rc;
            }
            GM_BEGIN
            gm_dot_kernel<<<g, 256, 0, st>>>(m, sc, K->gm_V[0], q, K->gm_R.p, nr, ctx->d_partials, ctx->d_counters); JB_CHECK_LAUNCH(ctx);
            for (int i = 1; i <= j; i++) {
                gm_mgs_kernel<<<g, 256, 0, st>>>(m, sc, K->gm_V[i - 1], i < j ? K->gm_V[i] : nullptr, q, K->gm_R.p, gc, gs, z, K->d_hist.p, K->hist_cap, nr, i, j,
                                                 base, ctx->d_partials, ctx->d_counters);
                JB_CHECK_LAUNCH(ctx);
            }
            if (j < inner_max) {   // v_{j+1} = q / h_{j+1,j} (a no-op once done); the last step of a pass needs no next basis vector
                gm_scale_kernel<<<g, 256, 0, st>>>(m, sc, GM_HBIS, q, K->gm_V[j]); JB_CHECK_LAUNCH(ctx);
            }
            GM_END
            const int slot = j & 1;
            JB_CUDA(ctx, cudaMemcpyAsync(K->h_flags + slot * KS_SIZE, sc, KS_SIZE * sizeof(double), cudaMemcpyDeviceToHost, st));
            JB_CUDA(ctx, cudaEventRecord(K->ev[slot], st));
        }
        // end of the pass: x (+)= N^{-1} V y  (fgmres: Z y)
        GM_BEGIN
        gm_backsolve_kernel<<<1, 32, 0, st>>>(sc, K->gm_R.p, z, y); JB_CHECK_LAUNCH(ctx);
        gm_combine_kernel<<<g, 256, 0, st>>>(m, sc, y, (const double* const*)(flex ? K->gm_Zptr.p : K->gm_Vptr.p), K->s.p); JB_CHECK_LAUNCH(ctx);
        GM_END
        const double* pass = K->s.p;
        if (right && !flex) { rc = jb_launch_ilu_apply_sc(F, K->s.p, K->t.p, nullptr); if (rc != JB_OK) return rc; pass = K->t.p; }
        GM_BEGIN
        gm_accumulate_kernel<<<g, 256, 0, st>>>(m, pass, K->x.p, first_pass ? 1 : 0); JB_CHECK_LAUNCH(ctx);
        GM_END
        first_pass = false;
        JB_CUDA(ctx, cudaMemcpyAsync(K->h_flags + 3 * KS_SIZE, sc, KS_SIZE * sizeof(double), cudaMemcpyDeviceToHost, st));
        JB_CUDA(ctx, cudaStreamSynchronize(st));
        const double* f = K->h_flags + 3 * KS_SIZE;
        base = (int)f[KS_ITER];
        if (!restart || f[KS_DONE] != 0.0 || base >= itmax) break;
        // restart: r0 = M^{-1}(b - A x), beta = |r0|, v_1 = r0 / beta
        rc = jb_launch_spmv_dots(A, K->x.p, q, JB_DOT_NONE, nullptr, nullptr); if (rc != JB_OK) return rc;
        if (K->schur && (rc = jb_schur_correct_launch(K->schur, K->x.p, q, 1.0)) != JB_OK) return rc;
        GM_BEGIN
        gm_residual_kernel<<<g, 256, 0, st>>>(m, d_b, q, K->r.p); JB_CHECK_LAUNCH(ctx);
        GM_END
        const double* rr = K->r.p;
        if (left) { rc = jb_launch_ilu_apply_sc(F, K->r.p, K->s.p, nullptr); if (rc != JB_OK) return rc; rr = K->s.p; }
        GM_BEGIN
        gm_reinit_kernel<<<g, 256, 0, st>>>(m, rr, sc, z, ctx->d_partials, ctx->d_counters); JB_CHECK_LAUNCH(ctx);
        gm_scale_kernel<<<g, 256, 0, st>>>(m, sc, GM_BETA, rr, K->gm_V[0]); JB_CHECK_LAUNCH(ctx);
        GM_END
    }
    GM_BEGIN
    gm_negate_kernel<<<g, 256, 0, st>>>(m, K->x.p, d_dx); JB_CHECK_LAUNCH(ctx);   // dx = -x (update_dx_from_vector!)
    GM_END
    JB_CUDA(ctx, cudaMemcpyAsync(K->h_flags + 3 * KS_SIZE, sc, KS_SIZE * sizeof(double), cudaMemcpyDeviceToHost, st));
    JB_CUDA(ctx, cudaStreamSynchronize(st));
    const double* f = K->h_flags + 3 * KS_SIZE;
    const int niter = (int)f[KS_ITER];
    if (iters) *iters = niter;
    if (hist && hist_cap > 0) JB_CUDA(ctx, cudaMemcpy(hist, K->d_hist.p, std::min(hist_cap, niter + 1) * sizeof(double), cudaMemcpyDeviceToHost));
    int status = (int)f[KS_STATUS];
    if (f[KS_DONE] == 0.0) status = JB_NOT_CONVERGED;
    *status_out = status;
    return JB_OK;
}

extern "C" int32_t jb_krylov_set_gmres(jb_krylov* K, int32_t memory, int32_t restart, int32_t flexible) {
    if (!K || memory < 1) return JB_ERR_ARG;
    K->gm_memory = memory; K->gm_restart = restart != 0; K->gm_flexible = flexible != 0;
    return JB_OK;
}
