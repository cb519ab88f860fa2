// ILU(0)-preconditioned BiCGStab for sm_100a, in the operation order of Krylov.jl
// 0.9 `bicgstab!` as Jutul calls it (src/linsolve/krylov.jl:71-182,212-238):
// x0 = 0, c = b, right preconditioning by default (src/linsolve/utils.jl:25),
// left for the distributed path (ext/JutulPartitionedArraysExt/krylov.jl:60).
//
// B200 design: the whole iteration lives on the device. The 17 vector streams
// of an iteration are fused into three update kernels; the five inner products
// are fused into the kernels that produce their operands (<c,v> and <t,s>,<t,t>
// into the SpMV, <c,r>,<r,r> into the residual update) and reduced
// deterministically; alpha, omega, beta, |r| and the termination flags are
// device-resident scalars written by the last CTA of the reducing kernel, so no
// host round trip sits between kernels. The host enqueues one iteration ahead
// and polls a pinned copy of the flags; once `done` is set all later kernels of
// the solve return immediately.
//
// Distributed form (jb_krylov_set_dist; ext/JutulPartitionedArraysExt/krylov.jl:1-146,
// linalg.jl:37-93): vectors hold [owned | ghost] cells; every vector update and inner
// product runs over the owned rows; the raw local sums are all-reduced in place with NCCL
// on the same stream and a one-thread kernel applies the scalar recurrence, so every rank
// holds bitwise-identical scalars and takes identical decisions; the operand of each SpMV
// gets its ghost section from a halo exchange (consistent!(X), linalg.jl:46).
#include <cstdlib>
#include "jb_internal.cuh"
#include "jb_krylov_scalars.cuh"
#include "jb_reduce.cuh"
#include "jb_stream2.cuh"

int jb_launch_ilu_apply_sc(jb_ilu* F, const double* d_b, double* d_x, const double* d_sc);
int jb_dist_krylov_vectors(jb_dist* D, i64 n_local, double** y, double** z);
bool jb_krylov_persistent_ok(jb_krylov* K, int side, i64 n_own);
bool jb_krylov_persistent_aligned(const double* d_b, const double* d_dx);
int jb_krylov_solve_persistent(jb_krylov* K, const double* d_b, double* d_dx, double rtol, double atol, int itmax, int min_it, int* iters,
                               double* hist, int hist_cap, int* status_out);

__global__ void __launch_bounds__(256) bicg_init_kernel(i64 m, const double* __restrict__ b, const double* rin, double* r, double* __restrict__ p,
                                                        double* __restrict__ x, double* __restrict__ v, double* __restrict__ s, double* sc,
                                                        double* hist, double* partials, unsigned int* counter) {
    double d[2] = {0.0, 0.0};
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (i64)gridDim.x * blockDim.x) {
        const double ri = rin[i];
        r[i] = ri; p[i] = ri; x[i] = 0.0; v[i] = 0.0; s[i] = 0.0;
        d[0] = fma(ri, ri, d[0]);
        d[1] = fma(__ldg(b + i), ri, d[1]);
    }
    grid_reduce<2, OpSum>(d, partials, counter, [=](double(&t)[2]) {
        sc[KS_SUM0] = t[0]; sc[KS_SUM1] = t[1];
        if (sc[KS_DIST] == 0.0) ks_fin_init(sc, hist);
    });
}

// s = r - alpha v        (x += alpha y is deferred to update2: nothing reads x in between and the termination flag can only
//                         change in update2, so the two axpys on x run back to back in registers — same rounding, one pass)
__global__ void __launch_bounds__(256) bicg_update1_kernel(i64 m, const double* sc, const double* __restrict__ r, const double* __restrict__ v,
                                                           double* __restrict__ s) {
    if (sc[KS_DONE] != 0.0) return;
    const double alpha = sc[KS_ALPHA];
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (i64)gridDim.x * blockDim.x)
        s[i] = fma(-alpha, __ldg(v + i), __ldg(r + i));
}

// x += alpha y ; x += omega z ; r = s - omega t ; rho' = <c,r>, |r| ; beta ; termination
__global__ void __launch_bounds__(256) bicg_update2_kernel(i64 m, double* sc, const double* __restrict__ s, const double* __restrict__ t,
                                                           const double* __restrict__ y, const double* __restrict__ z,
                                                           const double* __restrict__ c, double* __restrict__ x,
                                                           double* __restrict__ r, double* hist, int hist_cap, double* partials,
                                                           unsigned int* counter) {
    if (sc[KS_DONE] != 0.0) return;
    const double omega = sc[KS_OMEGA], alpha = sc[KS_ALPHA];
    double d[2] = {0.0, 0.0};
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (i64)gridDim.x * blockDim.x) {
        x[i] = fma(omega, __ldg(z + i), fma(alpha, __ldg(y + i), x[i]));
        const double ri = fma(-omega, __ldg(t + i), __ldg(s + i));
        r[i] = ri;
        d[0] = fma(__ldg(c + i), ri, d[0]);
        d[1] = fma(ri, ri, d[1]);
    }
    grid_reduce<2, OpSum>(d, partials, counter, [=](double(&tt)[2]) {
        sc[KS_SUM0] = tt[0]; sc[KS_SUM1] = tt[1];
        if (sc[KS_DIST] == 0.0) ks_fin_update2(sc, hist, hist_cap);
    });
}

// p = r + beta (p - omega v)
__global__ void __launch_bounds__(256) bicg_update3_kernel(i64 m, const double* sc, const double* __restrict__ r, const double* __restrict__ v,
                                                           double* __restrict__ p) {
    if (sc[KS_DONE] != 0.0) return;
    const double beta = sc[KS_BETA], omega = sc[KS_OMEGA];
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (i64)gridDim.x * blockDim.x) {
        const double pa = fma(-omega, __ldg(v + i), p[i]);
        p[i] = fma(beta, pa, __ldg(r + i));
    }
}

// stand-alone dots for the left-preconditioned path
template <int MODE>
__global__ void __launch_bounds__(256) bicg_dot_kernel(i64 m, double* sc, const double* __restrict__ a, const double* __restrict__ b,
                                                       double* partials, unsigned int* counter) {
    if (sc[KS_DONE] != 0.0) return;
    double d[2] = {0.0, 0.0};
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (i64)gridDim.x * blockDim.x) {
        const double ai = __ldg(a + i);
        d[0] = fma(ai, __ldg(b + i), d[0]);
        if (MODE == JB_DOT_TS_TT) d[1] = fma(ai, ai, d[1]);
    }
    grid_reduce<2, OpSum>(d, partials, counter, [=](double(&t)[2]) {
        sc[KS_SUM0] = t[0]; sc[KS_SUM1] = t[1];
        if (sc[KS_DIST] == 0.0) { if (MODE == JB_DOT_CV) ks_fin_alpha(sc); else ks_fin_omega(sc); }
    });
}

// distributed: scalar recurrence after the all-reduce of sc[KS_SUM0..1]. which: 0 init, 1 alpha, 2 omega, 3 update2
__global__ void krylov_finalize_kernel(int which, double* sc, double* hist, int hist_cap) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (which == 0) { ks_fin_init(sc, hist); return; }
    if (sc[KS_DONE] != 0.0) return;
    if (which == 1) ks_fin_alpha(sc);
    else if (which == 2) ks_fin_omega(sc);
    else ks_fin_update2(sc, hist, hist_cap);
}

__global__ void __launch_bounds__(256) negate_kernel(i64 m, const double* __restrict__ x, double* __restrict__ dx) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (i64)gridDim.x * blockDim.x) dx[i] = -__ldg(x + i);
}

static int vec_grid(jb_ctx* ctx, i64 m) {
    i64 want = (m + 255) / 256;
    i64 cap = std::min<i64>((i64)ctx->sm_count * 8, JB_MAX_PARTIALS);
    return (int)std::max<i64>(1, std::min(want, cap));
}

extern "C" {

int32_t jb_krylov_create(jb_csr* A, jb_ilu* ilu, int32_t kind, jb_krylov** out) {
    if (!A || !out) return JB_ERR_ARG;
    jb_ctx* ctx = A->ctx;
    if (kind != 0 && kind != 1) JB_FAIL(ctx, JB_ERR_UNSUPPORTED, "jb_krylov_create: kind must be 0 (bicgstab) or 1 (gmres)");
    jb_krylov* K = new jb_krylov();
    K->csr = A; K->ilu = ilu; K->kind = kind; K->m = A->n * A->bs;
    const size_t m = (size_t)K->m;
    bool ok = K->r.alloc(m) == cudaSuccess && K->p.alloc(m) == cudaSuccess && K->v.alloc(m) == cudaSuccess && K->s.alloc(m) == cudaSuccess &&
              K->y.alloc(m) == cudaSuccess && K->z.alloc(m) == cudaSuccess && K->t.alloc(m) == cudaSuccess && K->x.alloc(m) == cudaSuccess &&
              K->q.alloc(m) == cudaSuccess && K->d_sc.alloc(KS_SIZE) == cudaSuccess;
    K->hist_cap = 1026;
    ok = ok && K->d_hist.alloc(K->hist_cap) == cudaSuccess;
    ok = ok && cudaMallocHost((void**)&K->h_flags, 4 * KS_SIZE * sizeof(double)) == cudaSuccess;
    if (ok && cudaHostAlloc((void**)&K->h_prog, 64, cudaHostAllocMapped) == cudaSuccess) {
        if (cudaHostGetDevicePointer((void**)&K->d_prog, K->h_prog, 0) != cudaSuccess) { cudaFreeHost(K->h_prog); K->h_prog = nullptr; K->d_prog = nullptr; }
        else K->h_prog[0] = 0ULL;
    } else { K->h_prog = nullptr; cudaGetLastError(); }
    ok = ok && cudaEventCreateWithFlags(&K->ev[0], cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&K->ev[1], cudaEventDisableTiming) == cudaSuccess;
    K->yv = K->y.p; K->zv = K->z.p;
    if (ok) {   // ghost sections of SpMV operands must never hold NaN garbage before the first exchange
        cudaMemsetAsync(K->yv, 0, m * sizeof(double), ctx->stream); cudaMemsetAsync(K->zv, 0, m * sizeof(double), ctx->stream);
        cudaMemsetAsync(K->p.p, 0, m * sizeof(double), ctx->stream); cudaMemsetAsync(K->s.p, 0, m * sizeof(double), ctx->stream);
        cudaStreamSynchronize(ctx->stream);
    }
    if (!ok) { delete K; JB_FAIL(ctx, JB_ERR_ALLOC, "jb_krylov_create: allocation failed"); }
    *out = K;
    return JB_OK;
}
int32_t jb_krylov_destroy(jb_krylov* K) {
    if (K) {
        if (K->h_flags) cudaFreeHost(K->h_flags);
        if (K->h_prog) cudaFreeHost(K->h_prog);
        if (K->ev[0]) cudaEventDestroy(K->ev[0]);
        if (K->ev[1]) cudaEventDestroy(K->ev[1]);
        for (double* p : K->gm_V) cudaFree(p);
        for (double* p : K->gm_Z) cudaFree(p);
    }
    delete K;
    return JB_OK;
}
// MultiModel: the operator of the solve becomes S = B - C E^-1 D (linear_operator(::MultiLinearizedSystem), multimodel.jl:70-91);
// the preconditioner keeps acting on B = jacobian(sys) (multimodel.jl:162-169). NULL restores the plain Jacobian.
int32_t jb_krylov_set_schur(jb_krylov* K, jb_schur* S) {
    if (!K) return JB_ERR_ARG;
    if (S && jb_schur_nrows(S) != K->m) return JB_ERR_ARG;
    K->schur = S;
    return JB_OK;
}
int32_t jb_krylov_set_dist(jb_krylov* K, jb_dist* D) {
    if (!K) return JB_ERR_ARG;
    if (D && jb_dist_n_owned(D) > K->csr->n) return JB_ERR_ARG;
    K->dist = D;
    K->pk.built = false;
    // peer-memory runs keep the SpMV operands inside the rank's symmetric buffer: the persistent solve has the neighbours
    // store their boundary values straight into the ghost sections (krylov_persistent.cu)
    K->yv = K->y.p; K->zv = K->z.p;
    if (D && jb_dist_is_p2p(D)) {
        double *sy = nullptr, *sz = nullptr;
        if (jb_dist_krylov_vectors(D, K->csr->n, &sy, &sz) == JB_OK && sy && sz) { K->yv = sy; K->zv = sz; }
    }
    // interior/boundary SpMV overlap with the halo exchange: opt-in (JB_OVERLAP=1); it pays only when the per-rank problem is
    // large enough that the second (boundary) launch costs less than the exposed halo wait
    const char* ov = getenv("JB_OVERLAP");
    K->overlap = ov && ov[0] == '1';
    if (D && K->overlap && !K->csr->has_split) jb_csr_split_owned(K->csr, jb_dist_n_owned(D));
    return JB_OK;
}

}  // extern "C"

// Rows whose product with A is known without touching A when x = N^{-1} w comes from a two-colour ILU(0): a row i of the
// first colour (no L entries) whose off-diagonal couplings are all kept in U has D_i = A_ii, U_ij = A_ij and
// x_i = D_i^{-1}(w_i - sum_j U_ij x_j), hence (A x)_i = w_i. A stream chunk made of such rows only is flagged; the SpMV
// copies w there (48 B per row instead of ~270 B). Mathematically the same operator; rounding differs at the level of
// cond(A_ii) * eps, like any other reassociation. JB_RB_IDENTITY=0 switches it off.
static int krylov_prepare_ident(jb_krylov* K, i64 n_own) {
    jb_csr* A = K->csr;
    jb_ilu* F = K->ilu;
    if (A->n_ident_chunks >= 0 && A->ident_for == (const void*)F) return JB_OK;
    A->n_ident_chunks = 0;
    A->ident_for = (const void*)F;
    const char* e = getenv("JB_RB_IDENTITY");
    if (e && e[0] == '0') return JB_OK;
    if (!F || !F->two_colour || !F->stream_ok || A->h_chunks.empty()) return JB_OK;
    const int nch = (int)A->h_chunks.size() - 1;
    std::vector<unsigned char> flag(nch, 0);
    int cnt = 0;
    for (int c = 0; c < nch; c++) {
        bool ok = true;
        for (int32_t i = A->h_chunks[c]; ok && i < A->h_chunks[c + 1]; i++) {
            const int32_t rowlen = A->h_rowptr[i + 1] - A->h_rowptr[i];
            ok = i < n_own && F->h_Lend[i] == F->h_Lstart[i] && F->h_Uend[i] - F->h_Ustart[i] == rowlen - 1;
        }
        flag[c] = ok ? 1 : 0;
        cnt += ok;
    }
    // the same analysis on the chunk table of the TMA-staged SpMV (64-row chunks: more of them qualify)
    int cnt2 = 0;
    std::vector<S2Chunk> tab2;
    if (A->s2_ok) {
        // qualifying chunks that follow each other are merged into identity entries of up to JB_S2_IDENT_ROWS rows: they are
        // streamed as plain vectors, and an entry should carry about as many bytes as a matrix chunk
        for (const S2Chunk& ch0 : A->h_s2) {
            S2Chunk ch = ch0;
            bool ok = true;
            for (int32_t i = ch.t0; ok && i < ch.t0 + ch.nr; i++) {
                const int32_t rowlen = A->h_rowptr[i + 1] - A->h_rowptr[i];
                ok = i < n_own && F->h_Lend[i] == F->h_Lstart[i] && F->h_Uend[i] - F->h_Ustart[i] == rowlen - 1;
            }
            if (ok) {
                cnt2++;
                ch.flags |= 1;
                if (!tab2.empty() && (tab2.back().flags & 1) && tab2.back().t0 + tab2.back().nr == ch.t0 &&
                    tab2.back().nr + ch.nr <= JB_S2_IDENT_ROWS) {
                    tab2.back().nr += ch.nr; tab2.back().cnt += ch.cnt;
                    continue;
                }
            }
            tab2.push_back(ch);
        }
    }
    if (cnt == 0 && cnt2 == 0) return JB_OK;
    if (A->d_ident.upload(flag, A->ctx->stream) != cudaSuccess) return JB_ERR_ALLOC;
    if (cnt2 > 0 && A->d_s2_ident.upload(tab2, A->ctx->stream) != cudaSuccess) return JB_ERR_ALLOC;
    A->n_s2_ident = cnt2 > 0 ? (int)tab2.size() : 0;
    const char* sv = getenv("JB_STREAM_VARIANT");
    const bool use2 = cnt2 > 0 && !(sv && sv[0] == '1');      // the s2 kernel is the one that runs when its table exists (spmv.cu)
    A->n_ident_chunks = use2 ? cnt2 : cnt;
    A->n_ident_rows = A->n_ident_blocks = 0;
    if (use2) {
        for (const S2Chunk& ch : tab2)
            if (ch.flags & 1) { A->n_ident_rows += ch.nr; A->n_ident_blocks += ch.cnt; }
    } else {
        for (int c = 0; c < nch; c++)
            if (flag[c]) {
                A->n_ident_rows += A->h_chunks[c + 1] - A->h_chunks[c];
                A->n_ident_blocks += A->h_rowptr[A->h_chunks[c + 1]] - A->h_rowptr[A->h_chunks[c]];
            }
    }
    return JB_OK;
}
extern "C" int32_t jb_krylov_info(jb_krylov* K, int64_t* info) {
    if (!K || !info) return JB_ERR_ARG;
    jb_csr* A = K->csr;
    for (int q = 3; q < 8; q++) info[q] = 0;
    if (K->pk.built && K->pk.ok) {   // fused iteration kernel: identity rows are the prefix [0, n_id)
        info[0] = K->pk.n_id > 0 ? 1 : 0; info[1] = K->pk.n_id; info[2] = K->pk.n_id_blocks;
        info[3] = 1; info[4] = K->pk.b1 - K->pk.b0; info[5] = K->pk.n_own; info[6] = K->pk.grid; info[7] = K->pk.n_iso;
        return JB_OK;
    }
    const bool on = A->n_ident_chunks > 0 && A->ident_for == (const void*)K->ilu;
    info[0] = on ? A->n_ident_chunks : 0;
    info[1] = on ? A->n_ident_rows : 0;
    info[2] = on ? A->n_ident_blocks : 0;
    return JB_OK;
}

// Enqueue-only solve; the caller synchronises. Returns the status through *status_out after a
// final sync inside (the flags have to be read anyway).
int jb_krylov_solve_impl(jb_krylov* K, const double* d_b, double* d_dx, double rtol, double atol, int itmax, int min_it, int side,
                         int* iters, double* hist, int hist_cap, int* status_out) {
    jb_csr* A = K->csr;
    jb_ctx* ctx = A->ctx;
    cudaStream_t st = ctx->stream;
    jb_dist* D = K->dist;
    jb_schur* SC = K->schur;
    if (SC && D) JB_FAIL(ctx, JB_ERR_UNSUPPORTED, "bicgstab: the Schur-reduced operator is not available in the distributed solve");
    const int bs = A->bs;
    const i64 n_own = D ? jb_dist_n_owned(D) : A->n;     // rows that enter updates and inner products
    const i64 m = n_own * bs;
    const int g = vec_grid(ctx, m);
    jb_ilu* F = K->ilu;
    const bool right = (side == 0 && F), left = (side == 1 && F);
    if (itmax > K->hist_cap - 2) itmax = K->hist_cap - 2;
    double* sc = K->d_sc.p;
    if (right) { const int rci = krylov_prepare_ident(K, n_own); if (rci != JB_OK) return rci; }
    // the identity rows are exact only for factors built from the Jacobian's CURRENT values: a lagged preconditioner
    // (update_preconditioner = false, or values written after jb_ilu0_update) falls back to the full SpMV
    const bool ident = right && A->n_ident_chunks > 0 && A->ident_for == (const void*)F && F->factored_gen == A->val_gen;

    double h_sc[KS_SIZE];
    memset(h_sc, 0, sizeof(h_sc));
    const bool manual = min_it > 1;
    h_sc[KS_MANUAL] = manual ? 1.0 : 0.0;
    h_sc[KS_ABS_TOL] = atol; h_sc[KS_REL_TOL] = rtol; h_sc[KS_MIN_IT] = (double)min_it;
    h_sc[KS_ATOL] = manual ? 1e-20 : atol;
    h_sc[KS_RTOL] = manual ? 1e-20 : rtol;
    h_sc[KS_ITMAX] = (double)itmax;
    h_sc[KS_DIST] = D ? 1.0 : 0.0;
    memcpy(K->h_flags + 2 * KS_SIZE, h_sc, sizeof(h_sc));
    JB_CUDA(ctx, cudaMemcpyAsync(sc, K->h_flags + 2 * KS_SIZE, sizeof(h_sc), cudaMemcpyHostToDevice, st));

    int rc;
    // all-reduce of the two raw sums + scalar recurrence (distributed only)
    auto reduce_fin = [&](int which) -> int {
        if (!D) return JB_OK;
        if (jb_dist_is_p2p(D))   // all-reduce over peer memory fused with the recurrence: one kernel
            return jb_dist_allreduce_fin_launch(D, sc + KS_SUM0, 2, 0, which, sc, K->d_hist.p, K->hist_cap);
        int r2 = jb_dist_allreduce_launch(D, sc + KS_SUM0, 2, 0);
        if (r2 != JB_OK) return r2;
        krylov_finalize_kernel<<<1, 32, 0, st>>>(which, sc, K->d_hist.p, K->hist_cap);
        JB_CHECK_LAUNCH(ctx);
        return JB_OK;
    };
#define JB_VEC_BEGIN if (ctx->prof_on) jb_prof_begin(ctx, JB_PROF_VECTOR);
#define JB_VEC_END if (ctx->prof_on) jb_prof_end(ctx);

    const double* rin = d_b;
    if (left) {
        rc = jb_launch_ilu_apply_sc(F, d_b, K->r.p, nullptr);
        if (rc != JB_OK) return rc;
        rin = K->r.p;
    }
    JB_VEC_BEGIN
    bicg_init_kernel<<<g, 256, 0, st>>>(m, d_b, rin, K->r.p, K->p.p, K->x.p, K->v.p, K->s.p, sc, K->d_hist.p, ctx->d_partials, ctx->d_counters);
    JB_CHECK_LAUNCH(ctx);
    JB_VEC_END
    if ((rc = reduce_fin(0)) != JB_OK) return rc;
    JB_CUDA(ctx, cudaMemcpyAsync(K->h_flags, sc, KS_SIZE * sizeof(double), cudaMemcpyDeviceToHost, st));
    JB_CUDA(ctx, cudaEventRecord(K->ev[0], st));

    for (int it = 1; it <= itmax; it++) {
        // keep one iteration in flight ahead of the host: before enqueuing iteration `it`, make sure the
        // flags written after iteration it-2 (same pinned slot this iteration will reuse) are not `done`.
        if (it >= 2) {
            const int slot_prev = it & 1;
            JB_CUDA(ctx, cudaEventSynchronize(K->ev[slot_prev]));
            if (K->h_flags[slot_prev * KS_SIZE + KS_DONE] != 0.0) break;
        }
        double* yv = K->p.p;
        if (right) { rc = jb_launch_ilu_apply_sc(F, K->p.p, K->yv, sc); if (rc != JB_OK) return rc; yv = K->yv; }
        const bool overlap = D && !left && jb_dist_is_p2p(D) && A->has_split && K->overlap;
        if (overlap) {
            // consistent!(y) overlapped with the interior rows: push -> interior SpMV -> pull -> boundary SpMV (+ reduction)
            if ((rc = jb_dist_halo_push_launch(D, yv, bs)) != JB_OK) return rc;
            A->ident_src = ident ? K->p.p : nullptr;
            A->split_phase = 1; rc = jb_launch_spmv_dots(A, yv, K->v.p, JB_DOT_CV, d_b, sc, n_own); A->split_phase = 0;
            if (rc != JB_OK) { A->ident_src = nullptr; return rc; }
            if ((rc = jb_dist_halo_pull_launch(D, yv, bs)) != JB_OK) { A->ident_src = nullptr; return rc; }
            A->split_phase = 2; rc = jb_launch_spmv_dots(A, yv, K->v.p, JB_DOT_CV, d_b, sc, n_own); A->split_phase = 0;
            A->ident_src = nullptr;
            if (rc != JB_OK) return rc;
        } else if (D && (rc = jb_dist_halo_launch(D, yv, bs)) != JB_OK) return rc;    // consistent!(y)
        if (overlap) {
        } else if (!left && SC) {
            // v = S y = B y - C E^-1 D y: the inner product cannot ride on the SpMV, it follows the correction
            rc = jb_launch_spmv_dots(A, yv, K->v.p, JB_DOT_NONE, nullptr, nullptr); if (rc != JB_OK) return rc;
            rc = jb_schur_correct_launch(SC, yv, K->v.p, 1.0); if (rc != JB_OK) return rc;
            JB_VEC_BEGIN
            bicg_dot_kernel<JB_DOT_CV><<<g, 256, 0, st>>>(m, sc, K->v.p, d_b, ctx->d_partials, ctx->d_counters);
            JB_CHECK_LAUNCH(ctx);
            JB_VEC_END
        } else if (!left) {
            A->ident_src = ident ? K->p.p : nullptr;
            rc = jb_launch_spmv_dots(A, yv, K->v.p, JB_DOT_CV, d_b, sc, n_own);       // v = A y, alpha = rho/<c,v>
            A->ident_src = nullptr;
            if (rc != JB_OK) return rc;
        } else {
            rc = jb_launch_spmv_dots(A, yv, K->q.p, JB_DOT_NONE, nullptr, nullptr); if (rc != JB_OK) return rc;
            if (SC && (rc = jb_schur_correct_launch(SC, yv, K->q.p, 1.0)) != JB_OK) return rc;
            rc = jb_launch_ilu_apply_sc(F, K->q.p, K->v.p, sc); if (rc != JB_OK) return rc;
            JB_VEC_BEGIN
            bicg_dot_kernel<JB_DOT_CV><<<g, 256, 0, st>>>(m, sc, K->v.p, d_b, ctx->d_partials, ctx->d_counters);
            JB_CHECK_LAUNCH(ctx);
            JB_VEC_END
        }
        if ((rc = reduce_fin(1)) != JB_OK) return rc;
        JB_VEC_BEGIN
        bicg_update1_kernel<<<g, 256, 0, st>>>(m, sc, K->r.p, K->v.p, K->s.p);
        JB_CHECK_LAUNCH(ctx);
        JB_VEC_END
        double* zv = K->s.p;
        if (right) { rc = jb_launch_ilu_apply_sc(F, K->s.p, K->zv, sc); if (rc != JB_OK) return rc; zv = K->zv; }
        if (overlap) {
            if ((rc = jb_dist_halo_push_launch(D, zv, bs)) != JB_OK) return rc;
            A->ident_src = ident ? K->s.p : nullptr;
            A->split_phase = 1; rc = jb_launch_spmv_dots(A, zv, K->t.p, JB_DOT_TS_TT, K->s.p, sc, n_own); A->split_phase = 0;
            if (rc != JB_OK) { A->ident_src = nullptr; return rc; }
            if ((rc = jb_dist_halo_pull_launch(D, zv, bs)) != JB_OK) { A->ident_src = nullptr; return rc; }
            A->split_phase = 2; rc = jb_launch_spmv_dots(A, zv, K->t.p, JB_DOT_TS_TT, K->s.p, sc, n_own); A->split_phase = 0;
            A->ident_src = nullptr;
            if (rc != JB_OK) return rc;
        } else if (D && (rc = jb_dist_halo_launch(D, zv, bs)) != JB_OK) return rc;    // consistent!(z)
        if (overlap) {
        } else if (!left && SC) {
            rc = jb_launch_spmv_dots(A, zv, K->t.p, JB_DOT_NONE, nullptr, nullptr); if (rc != JB_OK) return rc;
            rc = jb_schur_correct_launch(SC, zv, K->t.p, 1.0); if (rc != JB_OK) return rc;
            JB_VEC_BEGIN
            bicg_dot_kernel<JB_DOT_TS_TT><<<g, 256, 0, st>>>(m, sc, K->t.p, K->s.p, ctx->d_partials, ctx->d_counters);
            JB_CHECK_LAUNCH(ctx);
            JB_VEC_END
        } else if (!left) {
            A->ident_src = ident ? K->s.p : nullptr;
            rc = jb_launch_spmv_dots(A, zv, K->t.p, JB_DOT_TS_TT, K->s.p, sc, n_own);  // t = A z, omega = <t,s>/<t,t>
            A->ident_src = nullptr;
            if (rc != JB_OK) return rc;
        } else {
            rc = jb_launch_spmv_dots(A, zv, K->q.p, JB_DOT_NONE, nullptr, nullptr); if (rc != JB_OK) return rc;
            if (SC && (rc = jb_schur_correct_launch(SC, zv, K->q.p, 1.0)) != JB_OK) return rc;
            rc = jb_launch_ilu_apply_sc(F, K->q.p, K->t.p, sc); if (rc != JB_OK) return rc;
            JB_VEC_BEGIN
            bicg_dot_kernel<JB_DOT_TS_TT><<<g, 256, 0, st>>>(m, sc, K->t.p, K->s.p, ctx->d_partials, ctx->d_counters);
            JB_CHECK_LAUNCH(ctx);
            JB_VEC_END
        }
        if ((rc = reduce_fin(2)) != JB_OK) return rc;
        JB_VEC_BEGIN
        bicg_update2_kernel<<<g, 256, 0, st>>>(m, sc, K->s.p, K->t.p, yv, zv, d_b, K->x.p, K->r.p, K->d_hist.p, K->hist_cap, ctx->d_partials,
                                               ctx->d_counters);
        JB_CHECK_LAUNCH(ctx);
        JB_VEC_END
        if ((rc = reduce_fin(3)) != JB_OK) return rc;
        JB_VEC_BEGIN
        bicg_update3_kernel<<<g, 256, 0, st>>>(m, sc, K->r.p, K->v.p, K->p.p);
        JB_CHECK_LAUNCH(ctx);
        JB_VEC_END
        const int slot = it & 1;
        JB_CUDA(ctx, cudaMemcpyAsync(K->h_flags + slot * KS_SIZE, sc, KS_SIZE * sizeof(double), cudaMemcpyDeviceToHost, st));
        JB_CUDA(ctx, cudaEventRecord(K->ev[slot], st));
    }
    JB_VEC_BEGIN
    negate_kernel<<<g, 256, 0, st>>>(m, K->x.p, d_dx);   // dx = -x (update_dx_from_vector!)
    JB_CHECK_LAUNCH(ctx);
    JB_VEC_END
    JB_CUDA(ctx, cudaMemcpyAsync(K->h_flags + 3 * KS_SIZE, sc, KS_SIZE * sizeof(double), cudaMemcpyDeviceToHost, st));
    JB_CUDA(ctx, cudaStreamSynchronize(st));
    const double* f = K->h_flags + 3 * KS_SIZE;
    const int niter = (int)f[KS_ITER];
    if (iters) *iters = niter;
    if (hist && hist_cap > 0) {
        const int nh = std::min(hist_cap, niter + 1);
        JB_CUDA(ctx, cudaMemcpy(hist, K->d_hist.p, nh * sizeof(double), cudaMemcpyDeviceToHost));
    }
    int status = (int)f[KS_STATUS];
    if (f[KS_DONE] == 0.0) status = JB_NOT_CONVERGED;
    *status_out = status;
    return JB_OK;
}

int jb_gmres_solve_impl(jb_krylov* K, const double* d_b, double* d_dx, double rtol, double atol, int itmax, int side, int* iters, double* hist,
                        int hist_cap, int* status_out);
// kind 0: bicgstab, kind 1: gmres (min_iterations is honoured by bicgstab only)
int jb_krylov_dispatch(jb_krylov* K, const double* d_b, double* d_dx, double rtol, double atol, int itmax, int min_it, int side, int* iters,
                       double* hist, int hist_cap, int* status_out) {
    if (K->kind == 1) return jb_gmres_solve_impl(K, d_b, d_dx, rtol, atol, itmax, side, iters, hist, hist_cap, status_out);
    // two-colour ILU(0), 2x2 blocks, right preconditioning: the whole solve is one persistent kernel (krylov_persistent.cu)
    const i64 n_own = K->dist ? jb_dist_n_owned(K->dist) : K->csr->n;
    if (jb_krylov_persistent_aligned(d_b, d_dx) && jb_krylov_persistent_ok(K, side, n_own))
        return jb_krylov_solve_persistent(K, d_b, d_dx, rtol, atol, itmax, min_it, iters, hist, hist_cap, status_out);
    return jb_krylov_solve_impl(K, d_b, d_dx, rtol, atol, itmax, min_it, side, iters, hist, hist_cap, status_out);
}

extern "C" int32_t jb_krylov_solve(jb_krylov* K, const double* d_r, double* d_dx, double rtol, double atol, int32_t itmax, int32_t min_it,
                                   int32_t side, int32_t* iters, double* hist, int32_t hist_cap) {
    if (!K || !d_r || !d_dx || itmax < 0) return JB_ERR_ARG;
    int status = 0;
    int rc = jb_krylov_dispatch(K, d_r, d_dx, rtol, atol, itmax, min_it, side, iters, hist, hist_cap, &status);
    if (rc != JB_OK) return rc;
    return status;
}
