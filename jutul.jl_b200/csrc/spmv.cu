// Block-CSR SpMV for sm_100a: y = alpha*A*x + beta*y  (csr_mul_add!,
// src/StaticCSR/mat.jl:41-61; block_mul! src/linsolve/block_cpu.jl:1-9).
//
// HBM-bound: per row the kernel streams the row's blocks (8*bs^2 B each) and
// column indices (4 B) once, gathers x blocks (L2-resident for local numberings)
// and writes y. LPR lanes cooperate on a row so that a warp's loads of colidx /
// values cover contiguous memory (4 rows x ~7 blocks x 32 B for bs = 2); the
// partial sums are combined with LPR-wide shuffles. Optionally the Krylov dot
// products that consume y are fused in (<c,y>, or <y,u> and <y,y>), reduced
// deterministically (jb_reduce.cuh) and turned into alpha / omega by the last CTA.
#include "jb_internal.cuh"
#include "jb_krylov_scalars.cuh"
#include "jb_reduce.cuh"
#include "jb_stream.cuh"
#include "jb_stream2.cuh"

template <int BS> struct BlockLoad;
template <> struct BlockLoad<1> {
    __device__ static void mat(const double* __restrict__ v, size_t k, double (&a)[1]) { a[0] = __ldg(v + k); }
    __device__ static void vec(const double* __restrict__ x, int c, double (&o)[1]) { o[0] = __ldg(x + c); }
};
template <> struct BlockLoad<2> {
    __device__ static void mat(const double* __restrict__ v, size_t k, double (&a)[4]) {
        const double2* p = reinterpret_cast<const double2*>(v + 4 * k);
        double2 a0 = __ldg(p), a1 = __ldg(p + 1);
        a[0] = a0.x; a[1] = a0.y; a[2] = a1.x; a[3] = a1.y;
    }
    __device__ static void vec(const double* __restrict__ x, int c, double (&o)[2]) {
        double2 t = __ldg(reinterpret_cast<const double2*>(x) + c);
        o[0] = t.x; o[1] = t.y;
    }
};
template <int BS> struct BlockLoadN {
    __device__ static void mat(const double* __restrict__ v, size_t k, double (&a)[BS * BS]) {
#pragma unroll
        for (int i = 0; i < BS * BS; i++) a[i] = __ldg(v + k * BS * BS + i);
    }
    __device__ static void vec(const double* __restrict__ x, int c, double (&o)[BS]) {
#pragma unroll
        for (int i = 0; i < BS; i++) o[i] = __ldg(x + (size_t)c * BS + i);
    }
};
template <> struct BlockLoad<3> : BlockLoadN<3> {};
template <> struct BlockLoad<4> : BlockLoadN<4> {};

template <int BS, int LPR, int MODE>
__global__ void __launch_bounds__(256) spmv_kernel(i64 n, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                                                   const double* __restrict__ val, const double* __restrict__ x,
                                                   double* __restrict__ y, double alpha, double beta,
                                                   const double* __restrict__ u, double* sc, double* partials,
                                                   unsigned int* counter) {
    if (MODE != JB_DOT_NONE) {
        if (sc[KS_DONE] != 0.0) return;
    }
    const int lane = threadIdx.x % LPR;
    constexpr int ROWS_PER_WARP = 32 / LPR;
    const i64 warps_per_cta = blockDim.x >> 5;
    const i64 stride = (i64)gridDim.x * warps_per_cta * ROWS_PER_WARP;
    double d0 = 0.0, d1 = 0.0;
    // warp-uniform loop bound: every lane of a warp takes part in the shuffles
    for (i64 base = ((i64)blockIdx.x * warps_per_cta + (threadIdx.x >> 5)) * ROWS_PER_WARP; base < n; base += stride) {
        const i64 row = base + (threadIdx.x & 31) / LPR;
        const bool live = row < n;
        double acc[BS];
#pragma unroll
        for (int e = 0; e < BS; e++) acc[e] = 0.0;
        if (live) {
            const int32_t k0 = __ldg(rowptr + row), k1 = __ldg(rowptr + row + 1);
            for (int32_t k = k0 + lane; k < k1; k += LPR) {
                const int32_t c = __ldg(colidx + k);
                double a[BS * BS], xv[BS];
                BlockLoad<BS>::mat(val, (size_t)k, a);
                BlockLoad<BS>::vec(x, c, xv);
#pragma unroll
                for (int p = 0; p < BS; p++)
#pragma unroll
                    for (int e = 0; e < BS; e++) acc[e] = fma(a[p * BS + e], xv[p], acc[e]);
            }
        }
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1)
#pragma unroll
            for (int e = 0; e < BS; e++) acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], o);
        if (live && lane == 0) {
#pragma unroll
            for (int e = 0; e < BS; e++) {
                double v = alpha * acc[e];
                if (beta != 0.0) v += beta * y[row * BS + e];
                y[row * BS + e] = v;
                if (MODE == JB_DOT_CV) d0 = fma(__ldg(u + row * BS + e), v, d0);
                if (MODE == JB_DOT_TS_TT) { d0 = fma(v, __ldg(u + row * BS + e), d0); d1 = fma(v, v, d1); }
            }
        }
    }
    if (MODE == JB_DOT_CV) {
        double r[1] = {d0};
        grid_reduce<1, OpSum>(r, partials, counter, [=](double(&t)[1]) {
            sc[KS_SUM0] = t[0];
            if (sc[KS_DIST] == 0.0) ks_fin_alpha(sc);   // alpha_k = rho_k / <c, v_k>
        });
    } else if (MODE == JB_DOT_TS_TT) {
        double r[2] = {d0, d1};
        grid_reduce<2, OpSum>(r, partials, counter, [=](double(&t)[2]) {
            sc[KS_SUM0] = t[0]; sc[KS_SUM1] = t[1];
            if (sc[KS_DIST] == 0.0) ks_fin_omega(sc);   // omega_k = <t,s>/<t,t>
        });
    }
}

// ---- row-chunk stream form (default): see jb_stream.cuh -------------------------------------------------
template <int BS, int MODE>
__global__ void __launch_bounds__(256) spmv_stream_kernel(int nchunks, const int32_t* __restrict__ chunk_ptr, const int32_t* __restrict__ rowptr,
                                                          const int32_t* __restrict__ colidx, const double* __restrict__ val,
                                                          const double* __restrict__ x, double* __restrict__ y, double alpha, double beta,
                                                          const double* __restrict__ u, double* sc, double* partials, unsigned int* counter,
                                                          i64 n_dot, const int32_t* __restrict__ chunk_list, int slot_offset, int total_slots,
                                                          int finalize, const unsigned char* __restrict__ chunk_ident,
                                                          const double* __restrict__ ident_src) {
    if (MODE != JB_DOT_NONE) {
        if (sc[KS_DONE] != 0.0) return;
    }
    __shared__ int32_t s_rp[JB_CHUNK_ROWS + 1];
    extern __shared__ double s_prod[];
    double d0 = 0.0, d1 = 0.0;
    // nchunks = number of work items; chunk_list (optional) selects a subset of the chunks (interior / boundary rows of a rank)
    for (int k = blockIdx.x; k < nchunks; k += gridDim.x) {
        const int c = chunk_list ? __ldg(chunk_list + k) : k;
        const int t0 = __ldg(chunk_ptr + c), nr = __ldg(chunk_ptr + c + 1) - t0;
        if (chunk_ident && __ldg(chunk_ident + c)) {
            // x = N^{-1} w with a two-colour ILU(0): rows of the first colour satisfy (A x)_i = w_i identically
            // (x_i = D_i^{-1}(w_i - sum_j A_ij x_j), D_i = A_ii), so the product is read off w = ident_src (krylov.cu)
            if ((int)threadIdx.x < nr) {
                const size_t row = (size_t)t0 + threadIdx.x;
#pragma unroll
                for (int e = 0; e < BS; e++) {
                    const double v = ident_src[row * BS + e];
                    y[row * BS + e] = v;
                    if (MODE != JB_DOT_NONE && (i64)row < n_dot) {
                        if (MODE == JB_DOT_CV) d0 = fma(__ldg(u + row * BS + e), v, d0);
                        if (MODE == JB_DOT_TS_TT) { d0 = fma(v, __ldg(u + row * BS + e), d0); d1 = fma(v, v, d1); }
                    }
                }
            }
            continue;   // CTA-uniform branch, no barrier pending
        }
        stream_chunk_products<BS, JB_STREAM_U>(t0, nr, rowptr, colidx, val, 0, x, s_rp, s_prod);
        if ((int)threadIdx.x < nr) {
            const size_t row = (size_t)t0 + threadIdx.x;
            double acc[BS];
            stream_row_sum<BS>(threadIdx.x, s_rp, s_prod, acc);
#pragma unroll
            for (int e = 0; e < BS; e++) {
                double v = alpha * acc[e];
                if (beta != 0.0) v += beta * y[row * BS + e];
                y[row * BS + e] = v;
                if (MODE != JB_DOT_NONE && (i64)row < n_dot) {     // inner products run over owned rows only
                    if (MODE == JB_DOT_CV) d0 = fma(__ldg(u + row * BS + e), v, d0);
                    if (MODE == JB_DOT_TS_TT) { d0 = fma(v, __ldg(u + row * BS + e), d0); d1 = fma(v, v, d1); }
                }
            }
        }
        __syncthreads();
    }
    if (MODE == JB_DOT_CV) {
        double r[1] = {d0};
        grid_reduce_ex<1, OpSum>(r, partials, counter, slot_offset, total_slots, finalize != 0, [=](double(&t)[1]) {
            sc[KS_SUM0] = t[0];
            if (sc[KS_DIST] == 0.0) ks_fin_alpha(sc);
        });
    } else if (MODE == JB_DOT_TS_TT) {
        double r[2] = {d0, d1};
        grid_reduce_ex<2, OpSum>(r, partials, counter, slot_offset, total_slots, finalize != 0, [=](double(&t)[2]) {
            sc[KS_SUM0] = t[0]; sc[KS_SUM1] = t[1];
            if (sc[KS_DIST] == 0.0) ks_fin_omega(sc);
        });
    }
}

// ---- TMA-staged lane-pair form for 2x2 blocks (default when bs = 2): see jb_stream2.cuh -----------------------------
template <int MODE>
__global__ void __launch_bounds__(JB_S2_THREADS, 6) spmv_s2_kernel(int nchunks, const S2Chunk* __restrict__ table, const int32_t* __restrict__ rowptr,
                                                                   const int32_t* __restrict__ colidx, const double* __restrict__ val,
                                                                   const double* __restrict__ x, double* __restrict__ y, double alpha, double beta,
                                                                   const double* __restrict__ u, double* sc, double* partials, unsigned int* counter,
                                                                   i64 n_dot, const double* __restrict__ ident_src) {
    if (MODE != JB_DOT_NONE) {
        if (sc[KS_DONE] != 0.0) return;
    }
    extern __shared__ __align__(128) unsigned char s2_raw[];
    S2Smem& sm = *reinterpret_cast<S2Smem*>(s2_raw);
    s2_prologue(sm, table, 0, nchunks, rowptr, colidx, val, 0, nullptr);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, jl = lane >> 1, e = lane & 1;
    double d0 = 0.0, d1 = 0.0;
    int it = 0;
    for (int k = blockIdx.x; k < nchunks; k += gridDim.x, it++) {
        const int s = it & 1;
        jb_mbar_wait(&sm.bar[s], (uint32_t)(it >> 1) & 1u);
        const S2Stage& S = sm.stage[s];
        const int t0 = sm.meta[s].t0, nr = sm.meta[s].nr, e0 = sm.meta[s].e0, flags = sm.meta[s].flags;
        if (flags & 1) {
            // x = N^{-1} w with a two-colour ILU(0): (A x)_i = w_i on first-colour rows (krylov.cu) — read it off w.
            // An identity entry covers up to JB_S2_IDENT_ROWS consecutive rows; the CTA streams them as plain vectors.
            const size_t base = (size_t)t0 * 2;
            const int nd = nr * 2;
            for (int q0 = threadIdx.x; q0 < nd; q0 += JB_S2_THREADS * 4) {
                double v[4];
#pragma unroll
                for (int m = 0; m < 4; m++) { const int q = q0 + m * JB_S2_THREADS; if (q < nd) v[m] = ident_src[base + q]; }
#pragma unroll
                for (int m = 0; m < 4; m++) {
                    const int q = q0 + m * JB_S2_THREADS;
                    if (q < nd) {
                        y[base + q] = v[m];
                        if (MODE != JB_DOT_NONE && (i64)((base + q) >> 1) < n_dot) {
                            if (MODE == JB_DOT_CV) d0 = fma(__ldg(u + base + q), v[m], d0);
                            if (MODE == JB_DOT_TS_TT) { d0 = fma(v[m], __ldg(u + base + q), d0); d1 = fma(v[m], v[m], d1); }
                        }
                    }
                }
            }
            s2_release(sm, s, table, k, nchunks, rowptr, colidx, val, 0, nullptr);
            continue;
        }
        const int j = warp * 16 + jl;
        if (j < nr) {
            const size_t row = (size_t)t0 + j;
            double v;
            {
                const int lrp = jb_span_lead<int32_t>((size_t)t0), lcol = jb_span_lead<int32_t>((size_t)e0);
                const double acc = s2_row_sum<8>(S, lcol, S.rp[lrp + j] - e0, S.rp[lrp + j + 1] - e0, e, x);
                v = alpha * acc;
                if (beta != 0.0) v += beta * y[row * 2 + e];
            }
            y[row * 2 + e] = v;
            if (MODE != JB_DOT_NONE && (i64)row < n_dot) {     // inner products run over owned rows only
                if (MODE == JB_DOT_CV) d0 = fma(__ldg(u + row * 2 + e), v, d0);
                if (MODE == JB_DOT_TS_TT) { d0 = fma(v, __ldg(u + row * 2 + e), d0); d1 = fma(v, v, d1); }
            }
        }
        s2_release(sm, s, table, k, nchunks, rowptr, colidx, val, 0, nullptr);
    }
    if (MODE == JB_DOT_CV) {
        double r[1] = {d0};
        grid_reduce_ex<1, OpSum>(r, partials, counter, 0, gridDim.x, true, [=](double(&t)[1]) {
            sc[KS_SUM0] = t[0];
            if (sc[KS_DIST] == 0.0) ks_fin_alpha(sc);
        });
    } else if (MODE == JB_DOT_TS_TT) {
        double r[2] = {d0, d1};
        grid_reduce_ex<2, OpSum>(r, partials, counter, 0, gridDim.x, true, [=](double(&t)[2]) {
            sc[KS_SUM0] = t[0]; sc[KS_SUM1] = t[1];
            if (sc[KS_DIST] == 0.0) ks_fin_omega(sc);
        });
    }
}

static bool s2_enabled() {   // JB_STREAM_VARIANT=1 selects the first-generation register-stream kernels (A/B measurements)
    const char* e = getenv("JB_STREAM_VARIANT");
    return !(e && e[0] == '1');
}

template <int MODE>
static int launch_spmv_s2(jb_csr* A, double alpha, const double* x, double beta, double* y, const double* u, double* sc, i64 n_dot) {
    jb_ctx* ctx = A->ctx;
    ProfScope _ps(ctx, JB_PROF_SPMV);
    const bool ident = A->ident_src != nullptr && A->d_s2_ident.p != nullptr && A->n_s2_ident > 0;
    const int nchunks = ident ? A->n_s2_ident : (int)A->h_s2.size();
    const size_t smem = sizeof(S2Smem);
    static int per_sm = 0;
    if (per_sm == 0) {
        cudaFuncSetAttribute(spmv_s2_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, spmv_s2_kernel<MODE>, JB_S2_THREADS, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    }
    int cap = ctx->sm_count * per_sm;
    if (MODE != JB_DOT_NONE && cap > JB_MAX_PARTIALS) cap = JB_MAX_PARTIALS;
    const int grid = std::max(1, std::min(nchunks, cap));
    spmv_s2_kernel<MODE><<<grid, JB_S2_THREADS, smem, ctx->stream>>>(nchunks, ident ? A->d_s2_ident.p : A->d_s2.p, A->d_rowptr.p, A->d_colidx.p,
                                                                    A->d_val.p, x, y, alpha, beta, u, sc, ctx->d_partials, ctx->d_counters,
                                                                    n_dot < 0 ? A->n : n_dot, ident ? A->ident_src : nullptr);
    JB_CHECK_LAUNCH(ctx);
    return JB_OK;
}

template <int BS, int MODE>
static int launch_spmv_stream(jb_csr* A, double alpha, const double* x, double beta, double* y, const double* u, double* sc, i64 n_dot) {
    jb_ctx* ctx = A->ctx;
    ProfScope _ps(ctx, JB_PROF_SPMV);
    const int nchunks = (int)A->h_chunks.size() - 1;
    const size_t smem = (size_t)JB_CHUNK_CAP * BS * sizeof(double);
    static int per_sm = 0;   // resident CTAs per SM: the persistent grid is sized to exactly one wave
    if (per_sm == 0) {
        cudaFuncSetAttribute(spmv_stream_kernel<BS, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, spmv_stream_kernel<BS, MODE>, 256, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    }
    int cap = ctx->sm_count * per_sm;
    if (MODE != JB_DOT_NONE && cap > JB_MAX_PARTIALS) cap = JB_MAX_PARTIALS;
    if (A->split_phase != 0) {
        // distributed two-launch form: phase 1 = interior chunks (no ghost column; runs while the halo is in flight, deposits its
        // partial sums), phase 2 = boundary chunks (after the halo pull; finalises over both launches' partials)
        const bool interior = A->split_phase == 1;
        const int nwork = interior ? A->n_chunks_int : A->n_chunks_bnd;
        const int cap2 = std::min(cap, JB_MAX_PARTIALS / 2);
        const int g_int = std::max(1, std::min(A->n_chunks_int, cap2));
        const int grid2 = interior ? g_int : std::max(1, std::min(std::max(nwork, 1), cap2));
        spmv_stream_kernel<BS, MODE><<<grid2, 256, smem, ctx->stream>>>(nwork, A->d_chunks.p, A->d_rowptr.p, A->d_colidx.p, A->d_val.p, x, y, alpha, beta,
                                                                       u, sc, ctx->d_partials, ctx->d_counters, n_dot < 0 ? A->n : n_dot,
                                                                       interior ? A->d_chunks_int.p : A->d_chunks_bnd.p, interior ? 0 : g_int,
                                                                       interior ? g_int : g_int + grid2, interior ? 0 : 1,
                                                                       A->ident_src ? A->d_ident.p : nullptr, A->ident_src);
        JB_CHECK_LAUNCH(ctx);
        return JB_OK;
    }
    const int grid = std::max(1, std::min(nchunks, cap));
    spmv_stream_kernel<BS, MODE><<<grid, 256, smem, ctx->stream>>>(nchunks, A->d_chunks.p, A->d_rowptr.p, A->d_colidx.p, A->d_val.p, x, y, alpha,
                                                                  beta, u, sc, ctx->d_partials, ctx->d_counters, n_dot < 0 ? A->n : n_dot, nullptr, 0,
                                                                  grid, 1, A->ident_src ? A->d_ident.p : nullptr, A->ident_src);
    JB_CHECK_LAUNCH(ctx);
    return JB_OK;
}

template <int BS, int LPR, int MODE>
static int launch_spmv_t(jb_csr* A, double alpha, const double* x, double beta, double* y, const double* u, double* sc) {
    jb_ctx* ctx = A->ctx;
    ProfScope _ps(ctx, JB_PROF_SPMV);
    const int threads = 256;
    const i64 rows_per_cta = threads / LPR;
    i64 want = (A->n + rows_per_cta - 1) / rows_per_cta;
    i64 cap = (MODE == JB_DOT_NONE) ? want : (i64)ctx->sm_count * 8;   // fused dots: bounded grid for the partials buffer
    if (cap > JB_MAX_PARTIALS && MODE != JB_DOT_NONE) cap = JB_MAX_PARTIALS;
    int grid = (int)std::max<i64>(1, std::min(want, cap));
    spmv_kernel<BS, LPR, MODE><<<grid, threads, 0, ctx->stream>>>(A->n, A->d_rowptr.p, A->d_colidx.p, A->d_val.p, x, y, alpha, beta, u,
                                                                  sc, ctx->d_partials, ctx->d_counters);
    JB_CHECK_LAUNCH(ctx);
    return JB_OK;
}

template <int MODE>
static int launch_spmv_mode(jb_csr* A, double alpha, const double* x, double beta, double* y, const double* u, double* sc, i64 n_dot = -1) {
    if (A->bs == 2 && A->s2_ok && A->split_phase == 0 && s2_enabled() && x != y) return launch_spmv_s2<MODE>(A, alpha, x, beta, y, u, sc, n_dot);
    if (!A->h_chunks.empty() || n_dot >= 0) {
        if (A->h_chunks.empty()) return JB_ERR_UNSUPPORTED;   // distributed dots need the stream form
        switch (A->bs) {
            case 1: return launch_spmv_stream<1, MODE>(A, alpha, x, beta, y, u, sc, n_dot);
            case 2: return launch_spmv_stream<2, MODE>(A, alpha, x, beta, y, u, sc, n_dot);
            case 3: return launch_spmv_stream<3, MODE>(A, alpha, x, beta, y, u, sc, n_dot);
            case 4: return launch_spmv_stream<4, MODE>(A, alpha, x, beta, y, u, sc, n_dot);
        }
    }
    // fallback (a row longer than the shared-memory tile): lanes-per-row kernel,
    // lanes per row from the average row length (7 for a hex-like TPFA stencil)
    const double avg = (double)A->nnzb / (double)A->n;
    const int lpr = avg <= 2.5 ? 2 : (avg <= 5.0 ? 4 : (avg <= 12.0 ? 8 : 16));
#define JB_SPMV_CASE(BS)                                                                   \
    case BS:                                                                               \
        if (lpr == 2) return launch_spmv_t<BS, 2, MODE>(A, alpha, x, beta, y, u, sc);      \
        if (lpr == 4) return launch_spmv_t<BS, 4, MODE>(A, alpha, x, beta, y, u, sc);      \
        if (lpr == 8) return launch_spmv_t<BS, 8, MODE>(A, alpha, x, beta, y, u, sc);      \
        return launch_spmv_t<BS, 16, MODE>(A, alpha, x, beta, y, u, sc);
    switch (A->bs) {
        JB_SPMV_CASE(1)
        JB_SPMV_CASE(2)
        JB_SPMV_CASE(3)
        JB_SPMV_CASE(4)
    }
#undef JB_SPMV_CASE
    return JB_ERR_UNSUPPORTED;
}

int jb_launch_spmv(jb_csr* A, double alpha, const double* d_x, double beta, double* d_y) {
    return launch_spmv_mode<JB_DOT_NONE>(A, alpha, d_x, beta, d_y, nullptr, nullptr);
}
int jb_launch_spmv_dots(jb_csr* A, const double* d_x, double* d_y, int mode, const double* d_u, double* d_sc, i64 n_dot) {
    if (mode == JB_DOT_CV) return launch_spmv_mode<JB_DOT_CV>(A, 1.0, d_x, 0.0, d_y, d_u, d_sc, n_dot);
    if (mode == JB_DOT_TS_TT) return launch_spmv_mode<JB_DOT_TS_TT>(A, 1.0, d_x, 0.0, d_y, d_u, d_sc, n_dot);
    return launch_spmv_mode<JB_DOT_NONE>(A, 1.0, d_x, 0.0, d_y, nullptr, nullptr);
}

extern "C" int32_t jb_spmv(jb_csr* A, double alpha, const double* d_x, double beta, double* d_y) {
    if (!A || !d_x || !d_y) return JB_ERR_ARG;
    int rc = jb_launch_spmv(A, alpha, d_x, beta, d_y);
    if (rc != JB_OK) return rc;
    JB_CUDA(A->ctx, cudaStreamSynchronize(A->ctx->stream));
    return JB_OK;
}
