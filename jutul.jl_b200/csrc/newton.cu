// Newton glue kernels (streaming, one pass each) and the two in-tree test systems.
//   update_jutul_variable_internal!/choose_increment  src/variables/utils.jl:117-174
//   unit_update_pairs!                                 src/variables/utils.jl:471-482
//   increment_norm                                     src/models.jl:955-965
//   convergence_criterion                              src/equations.jl:619-629
//   apply_scaling_to_linearized_system!                src/linsolve/default.jl:325-385
//   unit_diagonalize!                                  ext/JutulPartitionedArraysExt/linalg.jl:1-35
//   SimpleHeatEquation                                 src/applications/test_systems/heat_2d/heat_2d.jl:7-49
//   VariablePoissonEquation                            src/applications/test_systems/variable_poisson/variable_poisson.jl:90-133
#include "jb_internal.cuh"
#include "jb_reduce.cuh"

__device__ __forceinline__ double sgn(double x) { return (double)((x > 0.0) - (x < 0.0)); }
__device__ __forceinline__ double choose_increment(double v, double dv, double abs_max, double rel_max, double minv, double maxv, double scale) {
    if (!isnan(scale)) dv = dv * scale;
    if (!isnan(abs_max)) dv = sgn(dv) * fmin(fabs(dv), abs_max);
    if (!isnan(rel_max)) dv = sgn(dv) * fmin(fabs(dv), rel_max * fabs(v));
    if (!isnan(minv)) dv = fmax(dv, minv - v);
    if (!isnan(maxv)) dv = fmin(dv, maxv - v);
    return dv;
}

__global__ void __launch_bounds__(256) update_scalar_kernel(i64 n, double* __restrict__ v, const double* __restrict__ dx, i64 stride, double w,
                                                            double abs_max, double rel_max, double minv, double maxv, double scale) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        const double vi = v[i];
        v[i] = vi + choose_increment(vi, w * __ldg(dx + i * stride), abs_max, rel_max, minv, maxv, scale);
    }
}
__global__ void __launch_bounds__(256) update_pair_kernel(i64 n, double* __restrict__ s, const double* __restrict__ dx, i64 stride, double w,
                                                          double abs_max, double minval, double maxval) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        double2 sv = reinterpret_cast<double2*>(s)[i];
        const double dv = w * choose_increment(sv.x, __ldg(dx + i * stride), abs_max, NAN, minval, maxval, NAN);
        sv.x += dv; sv.y -= dv;
        reinterpret_cast<double2*>(s)[i] = sv;
    }
}
// unit_sum_update! for more than two fractions (src/variables/utils.jl:393-521): unit_update_direction_local! with its
// fall-back unit_update_magnitude_local!, restated operation by operation (pick_relaxation included as written, :519-526).
// s is nf x n (fraction fastest), dx holds the nf - 1 increments of a cell at dx[i + stride * cell].
#define JB_MAX_FRACTIONS 8
#define JB_MINIMUM_SAT_RELAX 1e-3
__device__ __forceinline__ double pick_relaxation(double w, double dv, double dv0) {
    const double r = dv / dv0;
    if (dv0 != 0.0) w = fmin(w, r);
    return fmin(w, JB_MINIMUM_SAT_RELAX);
}
__device__ __forceinline__ void unit_update_magnitude_local(double* s, const double* dx, int nf, double minval, double maxval, double abs_max) {
    double dlast0 = 0.0;
    for (int i = 0; i < nf - 1; i++) {
        const double dv = choose_increment(s[i], dx[i], abs_max, NAN, minval, maxval, NAN);
        s[i] += dv;
        dlast0 -= dv;
    }
    const double dlast = choose_increment(s[nf - 1], dlast0, abs_max, NAN, minval, maxval, NAN);
    s[nf - 1] += dlast;
    if (dlast != dlast0) {      // the last value was not within bounds: renormalise
        double t = 0.0;
        for (int i = 0; i < nf; i++) t += s[i];
        for (int i = 0; i < nf; i++) s[i] = fmin(fmax(s[i], minval), maxval) / t;
    }
}
__global__ void __launch_bounds__(256) update_fractions_kernel(i64 n, int nf, double* __restrict__ sg, const double* __restrict__ dxg, i64 stride, double w0,
                                                               double abs_max, double minval, double maxval, int preserve_direction) {
    for (i64 c = (i64)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += (i64)gridDim.x * blockDim.x) {
        double s[JB_MAX_FRACTIONS], dx[JB_MAX_FRACTIONS];
        for (int i = 0; i < nf; i++) s[i] = sg[c * nf + i];
        for (int i = 0; i < nf - 1; i++) dx[i] = __ldg(dxg + c * stride + i);
        if (!preserve_direction) unit_update_magnitude_local(s, dx, nf, minval, maxval, abs_max);
        else {
            double w = 1.0, dlast0 = 0.0;
            for (int i = 0; i < nf - 1; i++) {
                const double dv = choose_increment(s[i], dx[i], abs_max, NAN, minval, maxval, NAN);
                dlast0 -= dx[i];
                w = pick_relaxation(w, dv, dx[i]);
            }
            const double dlast = choose_increment(s[nf - 1], dlast0, abs_max, NAN, minval, maxval, NAN);
            w = w0 * pick_relaxation(w, dlast, dlast0);
            if (w <= JB_MINIMUM_SAT_RELAX) unit_update_magnitude_local(s, dx, nf, minval, maxval, abs_max);
            else {
                for (int i = 0; i < nf - 1; i++) s[i] += w * dx[i];
                s[nf - 1] += w * dlast0;
            }
        }
        for (int i = 0; i < nf; i++) sg[c * nf + i] = s[i];
    }
}
__global__ void __launch_bounds__(256) copy_strided_kernel(i64 n, double* __restrict__ dst, i64 ds, const double* __restrict__ src, i64 ss) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) dst[i * ds] = __ldg(src + i * ss);
}
__global__ void __launch_bounds__(256) increment_norm_kernel(i64 n, const double* __restrict__ dx, i64 stride, double* out, double* partials,
                                                             unsigned int* counter) {
    double a[1] = {0.0}, b[1] = {0.0};
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        const double x = fabs(__ldg(dx + i * stride));
        a[0] += x; b[0] = fmax(b[0], x);
    }
    // two reductions with different operators share the workspace through disjoint halves
    grid_reduce<1, OpSum>(a, partials, counter, [=](double(&t)[1]) { out[0] = t[0]; });
    grid_reduce<1, OpMax>(b, partials + JB_MAX_PARTIALS, counter + 1, [=](double(&t)[1]) { out[1] = t[0]; });
}
template <int BS>
__global__ void __launch_bounds__(256) maxabs_rows_kernel(i64 n, const double* __restrict__ r, double* out, double* partials, unsigned int* counter) {
    double m[BS];
#pragma unroll
    for (int e = 0; e < BS; e++) m[e] = 0.0;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
#pragma unroll
        for (int e = 0; e < BS; e++) {
            const double x = fabs(__ldg(r + i * BS + e));
            // NaN must surface as an error, not vanish in fmax
            m[e] = (x != x) ? x : fmax(m[e], x);
        }
    }
    grid_reduce<BS, OpMax>(m, partials, counter, [=](double(&t)[BS]) {
#pragma unroll
        for (int e = 0; e < BS; e++) out[e] = t[e];
    });
}

__global__ void __launch_bounds__(256) scale_kernel(i64 n, double a, double* __restrict__ x) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) x[i] *= a;
}
template <int BS>
__global__ void __launch_bounds__(256) scale_diag_kernel(i64 n, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ diag,
                                                         double* __restrict__ val, double* __restrict__ r) {
    constexpr int B2 = BS * BS;
    for (i64 row = (i64)blockIdx.x * blockDim.x + threadIdx.x; row < n; row += (i64)gridDim.x * blockDim.x) {
        double D[B2], Di[B2], a[B2], o[B2], rv[BS], ro[BS];
        const int32_t dp = __ldg(diag + row);
#pragma unroll
        for (int q = 0; q < B2; q++) D[q] = val[(size_t)dp * B2 + q];
        blk_inv<BS>(D, Di);
        for (int32_t k = __ldg(rowptr + row); k < __ldg(rowptr + row + 1); k++) {
#pragma unroll
            for (int q = 0; q < B2; q++) a[q] = val[(size_t)k * B2 + q];
            blk_mul<BS>(Di, a, o);
#pragma unroll
            for (int q = 0; q < B2; q++) val[(size_t)k * B2 + q] = o[q];
        }
#pragma unroll
        for (int e = 0; e < BS; e++) rv[e] = r[row * BS + e];
        blk_mulvec<BS>(Di, rv, ro);
#pragma unroll
        for (int e = 0; e < BS; e++) r[row * BS + e] = ro[e];
    }
}
template <int BS>
__global__ void __launch_bounds__(256) unit_diag_kernel(i64 n_owned, i64 n, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                                                        double* __restrict__ val, double* __restrict__ r) {
    constexpr int B2 = BS * BS;
    for (i64 row = n_owned + (i64)blockIdx.x * blockDim.x + threadIdx.x; row < n; row += (i64)gridDim.x * blockDim.x) {
        for (int32_t k = __ldg(rowptr + row); k < __ldg(rowptr + row + 1); k++) {
            const bool dg = __ldg(colidx + k) == row;
#pragma unroll
            for (int q = 0; q < B2; q++) val[(size_t)k * B2 + q] = (dg && (q % (BS + 1) == 0)) ? -1.0 : 0.0;
        }
#pragma unroll
        for (int e = 0; e < BS; e++) r[row * BS + e] = 0.0;
    }
}

// ---- heat: one thread per cell writes its 5-entry row and residual ----
__global__ void __launch_bounds__(256) heat_assemble_kernel(i64 nx, i64 ny, double hx, double hy, double dt, const double* __restrict__ T,
                                                            const double* __restrict__ T0, const int32_t* __restrict__ rowptr,
                                                            const int32_t* __restrict__ colidx, double* __restrict__ nz, double* __restrict__ r) {
    const i64 nc = nx * ny;
    for (i64 c = (i64)blockIdx.x * blockDim.x + threadIdx.x; c < nc; c += (i64)gridDim.x * blockDim.x) {
        const i64 i = c % nx, j = c / nx;
        const i64 L = j * nx + (i == 0 ? nx - 1 : i - 1), R = j * nx + (i == nx - 1 ? 0 : i + 1);
        const i64 U = (j == 0 ? ny - 1 : j - 1) * nx + i, D = (j == ny - 1 ? 0 : j + 1) * nx + i;
        const double Tc = __ldg(T + c);
        const double dtt = (Tc - __ldg(T0 + c)) / dt;
        const double d2x = (__ldg(T + L) - 2 * Tc + __ldg(T + R)) / (hx * hx);
        const double d2y = (__ldg(T + U) - 2 * Tc + __ldg(T + D)) / (hy * hy);
        r[c] = dtt - (d2x + d2y);
        const double cx = -(1.0 / (hx * hx)), cy = -(1.0 / (hy * hy));
        const double cd = 1.0 / dt - ((-2.0) / (hx * hx) + (-2.0) / (hy * hy));
        for (int32_t k = __ldg(rowptr + c); k < __ldg(rowptr + c + 1); k++) {
            const i64 col = __ldg(colidx + k);
            double v = 0.0;   // aliasing neighbours on tiny periodic grids accumulate
            if (col == c) v += cd;
            if (col == L) v += cx;
            if (col == R) v += cx;
            if (col == U) v += cy;
            if (col == D) v += cy;
            nz[k] = v;
        }
    }
}

// ---- Poisson: row-owner over the half-face map ----
__global__ void __launch_bounds__(256) poisson_assemble_kernel(i64 nc, const int32_t* __restrict__ hf_pos, const int32_t* __restrict__ hf_other,
                                                               const int32_t* __restrict__ hf_face, const int32_t* __restrict__ hf_rowpos,
                                                               const int32_t* __restrict__ diag, const double* __restrict__ K,
                                                               const double* __restrict__ U, const double* __restrict__ U0, int time_dependent,
                                                               double dt, const double* __restrict__ src, double* __restrict__ nz,
                                                               double* __restrict__ r) {
    for (i64 c = (i64)blockIdx.x * blockDim.x + threadIdx.x; c < nc; c += (i64)gridDim.x * blockDim.x) {
        const double Uc = __ldg(U + c);
        double val = 0.0, dself = 0.0, div = 0.0, ddiv = 0.0;
        if (time_dependent) { val = (Uc - __ldg(U0 + c)) / dt; dself = 1.0 / dt; }
        // zero the row first: several half-faces may connect the same pair of cells
        for (int32_t i = __ldg(hf_pos + c); i < __ldg(hf_pos + c + 1); i++) nz[__ldg(hf_rowpos + i)] = 0.0;
        for (int32_t i = __ldg(hf_pos + c); i < __ldg(hf_pos + c + 1); i++) {
            const double Kf = __ldg(K + __ldg(hf_face + i));
            const double q = -Kf * (__ldg(U + __ldg(hf_other + i)) - Uc);
            div = div + q; ddiv = ddiv + Kf;
            nz[__ldg(hf_rowpos + i)] += -Kf;
        }
        val = time_dependent ? val + div : div;
        dself += ddiv;
        if (!time_dependent && c == 0) { val = val + 1e-10 * Uc; dself += 1e-10; }
        if (src) val += __ldg(src + c);
        r[c] = val;
        nz[__ldg(diag + c)] = dself;
    }
}
__global__ void scatter_scalar_sources_kernel(i64 nsrc, const int32_t* __restrict__ cells, const double* __restrict__ vals, double* __restrict__ src) {
    if (blockIdx.x == 0 && threadIdx.x == 0) for (i64 k = 0; k < nsrc; k++) src[cells[k]] += vals[k];
}

static int sgrid(jb_ctx* ctx, i64 n) {
    i64 want = (n + 255) / 256;
    i64 cap = std::min<i64>((i64)ctx->sm_count * 8, JB_MAX_PARTIALS);
    return (int)std::max<i64>(1, std::min(want, cap));
}

int jb_launch_update_scalar(jb_ctx* ctx, double* d_v, const double* d_dx, i64 stride, i64 n, double w, double abs_max, double rel_max,
                            double minv, double maxv, double scale) {
    ProfScope _ps(ctx, JB_PROF_NEWTON);
    update_scalar_kernel<<<sgrid(ctx, n), 256, 0, ctx->stream>>>(n, d_v, d_dx, stride, w, abs_max, rel_max, minv, maxv, scale);
    JB_CHECK_LAUNCH(ctx);
    return JB_OK;
}
int jb_launch_update_pair(jb_ctx* ctx, double* d_s, const double* d_dx, i64 stride, i64 n, double w, double abs_max, double minval, double maxval) {
    // unit_sum_update! (src/variables/utils.jl:393-413): maxval - nf*minval, then unit_update_pairs! clamps
    maxval = maxval - 2 * minval;
    maxval = fmin(1 - minval, maxval);
    minval = fmax(minval, maxval - 1);
    ProfScope _ps(ctx, JB_PROF_NEWTON);
    update_pair_kernel<<<sgrid(ctx, n), 256, 0, ctx->stream>>>(n, d_s, d_dx, stride, w, abs_max, minval, maxval);
    JB_CHECK_LAUNCH(ctx);
    return JB_OK;
}
int jb_launch_maxabs_rows(jb_ctx* ctx, const double* d_r, int bs, i64 n, double* d_out) {
    const int g = sgrid(ctx, n);
    ProfScope _ps(ctx, JB_PROF_NEWTON);
    switch (bs) {
        case 1: maxabs_rows_kernel<1><<<g, 256, 0, ctx->stream>>>(n, d_r, d_out, ctx->d_partials, ctx->d_counters); break;
        case 2: maxabs_rows_kernel<2><<<g, 256, 0, ctx->stream>>>(n, d_r, d_out, ctx->d_partials, ctx->d_counters); break;
        case 3: maxabs_rows_kernel<3><<<g, 256, 0, ctx->stream>>>(n, d_r, d_out, ctx->d_partials, ctx->d_counters); break;
        case 4: maxabs_rows_kernel<4><<<g, 256, 0, ctx->stream>>>(n, d_r, d_out, ctx->d_partials, ctx->d_counters); break;
        default: return JB_ERR_UNSUPPORTED;
    }
    JB_CHECK_LAUNCH(ctx);
    return JB_OK;
}

extern "C" {

int32_t jb_update_scalar(jb_ctx* ctx, double* d_v, const double* d_dx, int64_t dx_stride, int64_t n, double w, double abs_max, double rel_max,
                         double minv, double maxv, double scale) {
    if (!ctx || !d_v || !d_dx || n < 0 || dx_stride < 1) return JB_ERR_ARG;
    if (n == 0) return JB_OK;
    int rc = jb_launch_update_scalar(ctx, d_v, d_dx, dx_stride, n, w, abs_max, rel_max, minv, maxv, scale);
    if (rc != JB_OK) return rc;
    JB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return JB_OK;
}
int32_t jb_update_fraction_pair(jb_ctx* ctx, double* d_s, const double* d_dx, int64_t dx_stride, int64_t n, double w, double abs_max,
                                double minval, double maxval) {
    if (!ctx || !d_s || !d_dx || n < 0 || dx_stride < 1) return JB_ERR_ARG;
    if (n == 0) return JB_OK;
    int rc = jb_launch_update_pair(ctx, d_s, d_dx, dx_stride, n, w, abs_max, minval, maxval);
    if (rc != JB_OK) return rc;
    JB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return JB_OK;
}
// dst[i * dst_stride] = src[i * src_stride]: rows of a k x n state array (Saturations is 2 x n) as the contiguous per-cell
// vectors the secondary-variable graph reads, and back (total masses of two phases -> 2 x n)
int32_t jb_copy_strided(jb_ctx* ctx, double* d_dst, int64_t dst_stride, const double* d_src, int64_t src_stride, int64_t n) {
    if (!ctx || !d_dst || !d_src || n < 0 || dst_stride < 1 || src_stride < 1) return JB_ERR_ARG;
    if (n == 0) return JB_OK;
    {
        ProfScope _ps(ctx, JB_PROF_NEWTON);
        copy_strided_kernel<<<sgrid(ctx, n), 256, 0, ctx->stream>>>(n, d_dst, dst_stride, d_src, src_stride);
        JB_CHECK_LAUNCH(ctx);
    }
    JB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return JB_OK;
}
int32_t jb_update_fractions(jb_ctx* ctx, double* d_s, const double* d_dx, int64_t dx_stride, int32_t nf, int64_t n, double w, double abs_max,
                            double minval, double maxval, int32_t preserve_direction) {
    if (!ctx || !d_s || !d_dx || n < 0 || nf < 2 || nf > JB_MAX_FRACTIONS || dx_stride < nf - 1) return JB_ERR_ARG;
    if (n == 0) return JB_OK;
    if (nf == 2) return jb_update_fraction_pair(ctx, d_s, d_dx, dx_stride, n, w, abs_max, minval, maxval);
    {
        ProfScope _ps(ctx, JB_PROF_NEWTON);
        update_fractions_kernel<<<sgrid(ctx, n), 256, 0, ctx->stream>>>(n, nf, d_s, d_dx, dx_stride, w, abs_max, minval, maxval - nf * minval,
                                                                       preserve_direction);
        JB_CHECK_LAUNCH(ctx);
    }
    JB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return JB_OK;
}
int32_t jb_increment_norm(jb_ctx* ctx, const double* d_dx, int64_t stride, int64_t n, double* sum_abs, double* max_abs) {
    if (!ctx || !d_dx || n < 0 || stride < 1) return JB_ERR_ARG;
    if (n == 0) { if (sum_abs) *sum_abs = 0; if (max_abs) *max_abs = 0; return JB_OK; }
    increment_norm_kernel<<<sgrid(ctx, n), 256, 0, ctx->stream>>>(n, d_dx, stride, ctx->d_scalars, ctx->d_partials, ctx->d_counters);
    JB_CHECK_LAUNCH(ctx);
    JB_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned, ctx->d_scalars, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    JB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (sum_abs) *sum_abs = ctx->h_pinned[0];
    if (max_abs) *max_abs = ctx->h_pinned[1];
    return JB_OK;
}
int32_t jb_maxabs_rows(jb_ctx* ctx, const double* d_r, int32_t bs, int64_t n, double* out) {
    if (!ctx || !d_r || !out || n < 0 || bs < 1 || bs > 4) return JB_ERR_ARG;
    if (n == 0) { for (int e = 0; e < bs; e++) out[e] = 0; return JB_OK; }
    int rc = jb_launch_maxabs_rows(ctx, d_r, bs, n, ctx->d_scalars);
    if (rc != JB_OK) return rc;
    JB_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned, ctx->d_scalars, bs * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    JB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int e = 0; e < bs; e++) out[e] = ctx->h_pinned[e];
    return JB_OK;
}

int32_t jb_scale_system(jb_csr* A, double* d_r, int32_t kind, double dt) {
    if (!A || !d_r) return JB_ERR_ARG;
    jb_ctx* ctx = A->ctx;
    if (kind == 0) return JB_OK;
    jb_csr_touch(A);
    if (kind == 2) {
        const i64 nv = A->nnzb * A->bs * A->bs, nr = A->n * A->bs;
        scale_kernel<<<sgrid(ctx, nv), 256, 0, ctx->stream>>>(nv, dt, A->d_val.p); JB_CHECK_LAUNCH(ctx);
        scale_kernel<<<sgrid(ctx, nr), 256, 0, ctx->stream>>>(nr, dt, d_r); JB_CHECK_LAUNCH(ctx);
    } else if (kind == 1) {
        const int g = sgrid(ctx, A->n);
        switch (A->bs) {
            case 1: scale_diag_kernel<1><<<g, 256, 0, ctx->stream>>>(A->n, A->d_rowptr.p, A->d_diag.p, A->d_val.p, d_r); break;
            case 2: scale_diag_kernel<2><<<g, 256, 0, ctx->stream>>>(A->n, A->d_rowptr.p, A->d_diag.p, A->d_val.p, d_r); break;
            case 3: scale_diag_kernel<3><<<g, 256, 0, ctx->stream>>>(A->n, A->d_rowptr.p, A->d_diag.p, A->d_val.p, d_r); break;
            case 4: scale_diag_kernel<4><<<g, 256, 0, ctx->stream>>>(A->n, A->d_rowptr.p, A->d_diag.p, A->d_val.p, d_r); break;
        }
        JB_CHECK_LAUNCH(ctx);
    } else return JB_ERR_ARG;
    JB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return JB_OK;
}

int32_t jb_unit_diagonalize_ghosts(jb_csr* A, double* d_r, int64_t n_owned) {
    if (!A || !d_r || n_owned < 0 || n_owned > A->n) return JB_ERR_ARG;
    jb_ctx* ctx = A->ctx;
    jb_csr_touch(A);
    const i64 ng = A->n - n_owned;
    if (ng == 0) return JB_OK;
    const int g = sgrid(ctx, ng);
    switch (A->bs) {
        case 1: unit_diag_kernel<1><<<g, 256, 0, ctx->stream>>>(n_owned, A->n, A->d_rowptr.p, A->d_colidx.p, A->d_val.p, d_r); break;
        case 2: unit_diag_kernel<2><<<g, 256, 0, ctx->stream>>>(n_owned, A->n, A->d_rowptr.p, A->d_colidx.p, A->d_val.p, d_r); break;
        case 3: unit_diag_kernel<3><<<g, 256, 0, ctx->stream>>>(n_owned, A->n, A->d_rowptr.p, A->d_colidx.p, A->d_val.p, d_r); break;
        case 4: unit_diag_kernel<4><<<g, 256, 0, ctx->stream>>>(n_owned, A->n, A->d_rowptr.p, A->d_colidx.p, A->d_val.p, d_r); break;
    }
    JB_CHECK_LAUNCH(ctx);
    JB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return JB_OK;
}

int32_t jb_heat_pattern(jb_ctx* ctx, int64_t nx, int64_t ny, jb_csr** out) {
    if (!ctx || !out || nx < 1 || ny < 1) return JB_ERR_ARG;
    const i64 nc = nx * ny;
    std::vector<int64_t> I(5 * nc), J(5 * nc);
    i64 k = 0;
    for (i64 c = 0; c < nc; c++) {
        const i64 i = c % nx, j = c / nx;
        const i64 nb[5] = {c, j * nx + (i == 0 ? nx - 1 : i - 1), j * nx + (i == nx - 1 ? 0 : i + 1), (j == 0 ? ny - 1 : j - 1) * nx + i,
                           (j == ny - 1 ? 0 : j + 1) * nx + i};
        for (int w = 0; w < 5; w++) { I[k] = c + 1; J[k++] = nb[w] + 1; }
    }
    return jb_csr_create_from_coo(ctx, I.data(), J.data(), 5 * nc, nc, 1, out);
}
int32_t jb_heat_assemble(jb_csr* A, int64_t nx, int64_t ny, double hx, double hy, double dt, const double* d_T, const double* d_T0, double* d_r) {
    if (!A || !d_T || !d_T0 || !d_r || A->bs != 1 || A->n != nx * ny || !(dt > 0)) return JB_ERR_ARG;
    jb_ctx* ctx = A->ctx;
    jb_csr_touch(A);
    heat_assemble_kernel<<<sgrid(ctx, A->n), 256, 0, ctx->stream>>>(nx, ny, hx, hy, dt, d_T, d_T0, A->d_rowptr.p, A->d_colidx.p, A->d_val.p, d_r);
    JB_CHECK_LAUNCH(ctx);
    JB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return JB_OK;
}
int32_t jb_poisson_assemble(jb_tpfa* t, const double* d_K, const double* d_U, const double* d_U0, int32_t time_dependent, double dt,
                            int64_t nsrc, const int64_t* src_cells, const double* src_vals, double* d_r) {
    if (!t || !d_K || !d_U || !d_r || t->csr->bs != 1 || (time_dependent && (!d_U0 || !(dt > 0)))) return JB_ERR_ARG;
    jb_ctx* ctx = t->mesh->ctx;
    const i64 nc = t->mesh->nc;
    jb_csr_touch(t->csr);
    DBuf<double> src, sv;
    DBuf<int32_t> scell;
    if (nsrc > 0) {
        std::vector<int32_t> hc(nsrc);
        for (i64 k = 0; k < nsrc; k++) {
            if (src_cells[k] < 1 || src_cells[k] > nc) JB_FAIL(ctx, JB_ERR_ARG, "jb_poisson_assemble: source cell out of range");
            hc[k] = (int32_t)(src_cells[k] - 1);
        }
        std::vector<double> hv(src_vals, src_vals + nsrc);
        if (src.alloc(nc) != cudaSuccess || scell.upload(hc, ctx->stream) != cudaSuccess || sv.upload(hv, ctx->stream) != cudaSuccess)
            JB_FAIL(ctx, JB_ERR_ALLOC, "jb_poisson_assemble: allocation failed");
        JB_CUDA(ctx, cudaMemsetAsync(src.p, 0, nc * sizeof(double), ctx->stream));
        scatter_scalar_sources_kernel<<<1, 32, 0, ctx->stream>>>(nsrc, scell.p, sv.p, src.p);
        JB_CHECK_LAUNCH(ctx);
    }
    poisson_assemble_kernel<<<sgrid(ctx, nc), 256, 0, ctx->stream>>>(nc, t->mesh->d_hf_pos.p, t->mesh->d_hf_other.p, t->mesh->d_hf_face.p,
                                                                     t->d_hf_rowpos.p, t->csr->d_diag.p, d_K, d_U, d_U0, time_dependent, dt,
                                                                     nsrc > 0 ? src.p : nullptr, t->csr->d_val.p, d_r);
    JB_CHECK_LAUNCH(ctx);
    JB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return JB_OK;
}

}  // extern "C"

// ---- fill_equation_entries! for a GenericAutoDiffCache (src/ad/generic.jl:53-96) ---------------------------------------
// One thread per (slot, equation): the entry's value goes to the residual when the slot is the residual slot of its
// entity (diagonal_positions, else the first slot), each partial is SET at its aligned position of the flat nzval.
struct jb_generic {
    jb_csr* csr;
    int ne, np;
    i64 nu, nslots;
    DBuf<int32_t> d_slot_entity;   // entity of slot j, or -1 when the slot does not feed the residual ... stored as (entity << 1) | is_residual
    DBuf<int64_t> d_pos;           // 0-based flat nzval index, -1 = not aligned
    DBuf<double> d_entries;
    double* h_stage = nullptr;     // pinned staging of the entries
};
__global__ void __launch_bounds__(256) generic_fill_kernel(i64 nslots, int ne, int np, const int32_t* __restrict__ slot_entity,
                                                           const int64_t* __restrict__ pos, const double* __restrict__ entries,
                                                           double* __restrict__ nz, double* __restrict__ r) {
    const i64 total = nslots * ne;
    for (i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (i64)gridDim.x * blockDim.x) {
        const i64 j = t / ne;
        const int e = (int)(t % ne);
        const double* a = entries + (size_t)t * (1 + np);       // (j*ne + e) * (1 + np)
        const int32_t se = __ldg(slot_entity + j);
        if (se & 1) r[(size_t)(se >> 1) * ne + e] = a[0];
        for (int d = 0; d < np; d++) {
            const int64_t q = __ldg(pos + (size_t)j * ne * np + (size_t)e * np + d);
            if (q >= 0) nz[q] = a[1 + d];
        }
    }
}

extern "C" {
int32_t jb_generic_create(jb_csr* A, int32_t ne, int32_t np, int64_t nu, const int64_t* vpos, const int64_t* dpos, const int64_t* positions,
                          jb_generic** out) {
    if (!A || !vpos || !positions || !out || ne < 1 || np < 1 || nu < 0) return JB_ERR_ARG;
    jb_ctx* ctx = A->ctx;
    if (vpos[0] != 1) JB_FAIL(ctx, JB_ERR_ARG, "jb_generic_create: vpos must start at 1");
    const i64 nslots = vpos[nu] - 1;
    const i64 nzlen = A->nnzb * A->bs * A->bs;
    std::vector<int32_t> se((size_t)nslots, 0);
    for (i64 i = 0; i < nu; i++) {
        if (vpos[i + 1] < vpos[i]) JB_FAIL(ctx, JB_ERR_ARG, "jb_generic_create: vpos must be non-decreasing");
        i64 rs = -1;
        if (dpos) {
            rs = dpos[i] - 1;
            if (rs < vpos[i] - 1 || rs >= vpos[i + 1] - 1) JB_FAIL(ctx, JB_ERR_ARG, "jb_generic_create: diagonal position outside the entity's slots");
        } else if (vpos[i + 1] > vpos[i]) rs = vpos[i] - 1;
        for (i64 j = vpos[i] - 1; j < vpos[i + 1] - 1; j++) se[j] = (int32_t)((i << 1) | (j == rs ? 1 : 0));
    }
    std::vector<int64_t> hp((size_t)nslots * ne * np);
    for (size_t q = 0; q < hp.size(); q++) {
        if (positions[q] < 0 || positions[q] > nzlen) JB_FAIL(ctx, JB_ERR_ARG, "jb_generic_create: Jacobian position out of range");
        hp[q] = positions[q] - 1;
    }
    jb_generic* g = new jb_generic();
    g->csr = A; g->ne = ne; g->np = np; g->nu = nu; g->nslots = nslots;
    const size_t nent = (size_t)nslots * ne * (1 + np);
    bool ok = g->d_slot_entity.upload(se, ctx->stream) == cudaSuccess && g->d_pos.upload(hp, ctx->stream) == cudaSuccess &&
              g->d_entries.alloc(std::max<size_t>(nent, 1)) == cudaSuccess &&
              cudaMallocHost((void**)&g->h_stage, std::max<size_t>(nent, 1) * sizeof(double)) == cudaSuccess;
    if (!ok) { jb_generic_destroy(g); JB_FAIL(ctx, JB_ERR_ALLOC, "jb_generic_create: allocation failed"); }
    *out = g;
    return JB_OK;
}
int32_t jb_generic_destroy(jb_generic* g) {
    if (g && g->h_stage) cudaFreeHost(g->h_stage);
    delete g;
    return JB_OK;
}
int32_t jb_generic_fill(jb_generic* g, const double* entries_host, double* d_r, int64_t r_offset) {
    if (!g || !entries_host || !d_r || r_offset < 0) return JB_ERR_ARG;
    jb_ctx* ctx = g->csr->ctx;
    const size_t nent = (size_t)g->nslots * g->ne * (1 + g->np);
    if (nent == 0) return JB_OK;
    jb_csr_touch(g->csr);
    memcpy(g->h_stage, entries_host, nent * sizeof(double));
    JB_CUDA(ctx, cudaMemcpyAsync(g->d_entries.p, g->h_stage, nent * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    {
        ProfScope _ps(ctx, JB_PROF_ASSEMBLY);
        generic_fill_kernel<<<sgrid(ctx, g->nslots * g->ne), 256, 0, ctx->stream>>>(g->nslots, g->ne, g->np, g->d_slot_entity.p, g->d_pos.p,
                                                                                  g->d_entries.p, g->csr->d_val.p, d_r + r_offset);
        JB_CHECK_LAUNCH(ctx);
    }
    JB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return JB_OK;
}
}  // extern "C"

// ---- cell renumbering between the caller's numbering and the device numbering (jb_order_multicolor) ----
template <int BS>
__global__ void __launch_bounds__(256) permute_kernel(i64 n, const int32_t* __restrict__ perm, const double* __restrict__ src,
                                                      double* __restrict__ dst, int to_caller) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        const size_t j = (size_t)__ldg(perm + i);
#pragma unroll
        for (int e = 0; e < BS; e++) {
            if (to_caller) dst[(size_t)i * BS + e] = __ldg(src + j * BS + e);   // caller[i] = device[perm[i]]
            else dst[j * BS + e] = __ldg(src + (size_t)i * BS + e);             // device[perm[i]] = caller[i]
        }
    }
}

int jb_launch_permute(jb_ctx* ctx, const int32_t* d_perm, i64 n, int bs, const double* d_src, double* d_dst, int to_caller) {
    ProfScope _ps(ctx, JB_PROF_NEWTON);
    const int g = sgrid(ctx, n);
    switch (bs) {
        case 1: permute_kernel<1><<<g, 256, 0, ctx->stream>>>(n, d_perm, d_src, d_dst, to_caller); break;
        case 2: permute_kernel<2><<<g, 256, 0, ctx->stream>>>(n, d_perm, d_src, d_dst, to_caller); break;
        case 3: permute_kernel<3><<<g, 256, 0, ctx->stream>>>(n, d_perm, d_src, d_dst, to_caller); break;
        case 4: permute_kernel<4><<<g, 256, 0, ctx->stream>>>(n, d_perm, d_src, d_dst, to_caller); break;
        default: return JB_ERR_UNSUPPORTED;
    }
    JB_CHECK_LAUNCH(ctx);
    return JB_OK;
}

extern "C" {

int32_t jb_perm_create(jb_ctx* ctx, const int64_t* perm, int64_t n, jb_perm** out) {
    if (!ctx || !perm || !out || n < 1) return JB_ERR_ARG;
    jb_perm* P = new jb_perm();
    P->ctx = ctx; P->n = n;
    std::vector<int32_t> h(n);
    std::vector<char> seen(n, 0);
    for (i64 i = 0; i < n; i++) {
        if (perm[i] < 1 || perm[i] > n || seen[perm[i] - 1]) { delete P; JB_FAIL(ctx, JB_ERR_ARG, "jb_perm_create: not a permutation of 1..n"); }
        seen[perm[i] - 1] = 1;
        h[i] = (int32_t)(perm[i] - 1);
    }
    if (P->d_perm.upload(h, ctx->stream) != cudaSuccess) { delete P; JB_FAIL(ctx, JB_ERR_ALLOC, "jb_perm_create: allocation failed"); }
    *out = P;
    return JB_OK;
}
int32_t jb_perm_destroy(jb_perm* P) { delete P; return JB_OK; }
int32_t jb_perm_apply(jb_perm* P, const double* d_src, double* d_dst, int32_t bs, int32_t to_caller) {
    if (!P || !d_src || !d_dst || d_src == d_dst) return JB_ERR_ARG;
    int rc = jb_launch_permute(P->ctx, P->d_perm.p, P->n, bs, d_src, d_dst, to_caller);
    if (rc != JB_OK) return rc;
    JB_CUDA(P->ctx, cudaStreamSynchronize(P->ctx->stream));
    return JB_OK;
}

}  // extern "C"

// ---- NFVM run-time flux: evaluate_flux / ntpfa_half_flux / compute_r / tpfa_flux (src/NFVM/evaluation.jl:1-88) ----
// One lane per face; the MPFA remainder is a short CSR row of (cell, T) pairs per face.
struct jb_nfvm {
    jb_ctx* ctx;
    i64 nf;
    int scheme;   // 0 linear, 1 :ntpfa, 2 :nmpfa
    DBuf<int32_t> left, right, L_ptr, L_cell, R_ptr, R_cell;
    DBuf<double> L_Tl, L_Tr, L_T, R_Tl, R_Tr, R_T;
    // conservation law on this flux with the :fvm assembly (below): host copies of the discretisation, the stencil of every
    // face (discretization_stencil), per stencil slot the coefficient of p[cell] in (q_l, r_l, q_R, r_R) and the aligned
    // Jacobian positions of (left, cell) / (right, cell)
    i64 nc = 0;
    std::vector<int32_t> h_left, h_right, h_Lptr, h_Lcell, h_Rptr, h_Rcell;
    std::vector<double> h_LTl, h_LTr, h_LT, h_RTl, h_RTr, h_RT;
    std::vector<int32_t> h_vpos, h_vars, h_lpos, h_rpos;
    DBuf<int32_t> d_vpos, d_vars, d_lpos, d_rpos;
    DBuf<double4> d_coef;
    jb_csr* csr = nullptr;
};

__device__ __forceinline__ double nfvm_compute_r(const int32_t* __restrict__ ptr, const int32_t* __restrict__ cell, const double* __restrict__ T,
                                                 const double* __restrict__ p, i64 nph, i64 ph0, i64 f) {
    double r = 0.0;
    for (int32_t k = __ldg(ptr + f); k < __ldg(ptr + f + 1); k++) r += __ldg(p + (size_t)__ldg(cell + k) * nph + ph0) * __ldg(T + k);
    return r;
}
__global__ void __launch_bounds__(256) nfvm_flux_kernel(i64 nf, int scheme, const int32_t* __restrict__ left, const int32_t* __restrict__ right,
                                                        const double* __restrict__ L_Tl, const double* __restrict__ L_Tr,
                                                        const int32_t* __restrict__ L_ptr, const int32_t* __restrict__ L_cell, const double* __restrict__ L_T,
                                                        const double* __restrict__ R_Tl, const double* __restrict__ R_Tr,
                                                        const int32_t* __restrict__ R_ptr, const int32_t* __restrict__ R_cell, const double* __restrict__ R_T,
                                                        const double* __restrict__ p, i64 nph, i64 ph0, double* __restrict__ q) {
    for (i64 f = (i64)blockIdx.x * blockDim.x + threadIdx.x; f < nf; f += (i64)gridDim.x * blockDim.x) {
        const double p_l = __ldg(p + (size_t)__ldg(left + f) * nph + ph0), p_r = __ldg(p + (size_t)__ldg(right + f) * nph + ph0);
        const double r_l = nfvm_compute_r(L_ptr, L_cell, L_T, p, nph, ph0, f);
        double q_l = __dadd_rn(__dmul_rn(__ldg(L_Tl + f), p_l), __dmul_rn(__ldg(L_Tr + f), p_r));
        q_l = q_l + r_l;
        if (scheme == 0) { q[f] = q_l; continue; }
        double r_r = nfvm_compute_r(R_ptr, R_cell, R_T, p, nph, ph0, f);
        double q_r = __dadd_rn(__dmul_rn(__ldg(R_Tl + f), p_l), __dmul_rn(__ldg(R_Tr + f), p_r));
        q_r = -(q_r + r_r); r_r = -r_r;
        double r_lw = r_l, r_rw = r_r;
        if (scheme == 2) { r_lw = fabs(r_l); r_rw = fabs(r_r); }
        const double r_total = r_lw + r_rw;
        double mu_l = 0.5, mu_r = 0.5;
        if (!(fabs(r_total) < 1e-10)) { mu_l = r_rw / r_total; mu_r = r_lw / r_total; }
        q[f] = __dsub_rn(__dmul_rn(mu_l, q_l), __dmul_rn(mu_r, q_r));
    }
}

extern "C" {

static int nfvm_half(jb_ctx* ctx, i64 nf, i64 nc, const double* Tl, const double* Tr, const int64_t* ptr, const int64_t* cell, const double* T,
                     DBuf<double>& dTl, DBuf<double>& dTr, DBuf<int32_t>& dptr, DBuf<int32_t>& dcell, DBuf<double>& dT,
                     std::vector<int32_t>& hp, std::vector<int32_t>& hc, std::vector<double>& a, std::vector<double>& b, std::vector<double>& c) {
    hp.assign(nf + 1, 0);
    for (i64 f = 0; f <= nf; f++) { if (ptr[f] < 1) return JB_ERR_ARG; hp[f] = (int32_t)(ptr[f] - 1); }
    const i64 nnz = hp[nf];
    hc.assign(nnz, 0);
    for (i64 k = 0; k < nnz; k++) { if (cell[k] < 1 || cell[k] > nc) return JB_ERR_ARG; hc[k] = (int32_t)(cell[k] - 1); }
    a.assign(Tl, Tl + nf); b.assign(Tr, Tr + nf); c.assign(T, T + nnz);
    cudaStream_t s = ctx->stream;
    bool ok = dTl.upload(a, s) == cudaSuccess && dTr.upload(b, s) == cudaSuccess && dptr.upload(hp, s) == cudaSuccess &&
              dcell.upload(hc, s) == cudaSuccess && dT.upload(c, s) == cudaSuccess;
    return ok ? JB_OK : JB_ERR_ALLOC;
}

int32_t jb_nfvm_create(jb_ctx* ctx, int64_t nf, int64_t nc, int32_t scheme, const int64_t* left, const int64_t* right, const double* L_Tl,
                       const double* L_Tr, const int64_t* L_ptr, const int64_t* L_cell, const double* L_T, const double* R_Tl, const double* R_Tr,
                       const int64_t* R_ptr, const int64_t* R_cell, const double* R_T, jb_nfvm** out) {
    if (!ctx || !out || nf < 1 || scheme < 0 || scheme > 2 || !left || !right || !L_Tl || !L_Tr || !L_ptr) return JB_ERR_ARG;
    if (scheme > 0 && (!R_Tl || !R_Tr || !R_ptr)) return JB_ERR_ARG;
    jb_nfvm* d = new jb_nfvm();
    d->ctx = ctx; d->nf = nf; d->scheme = scheme; d->nc = nc;
    std::vector<int32_t>&hl = d->h_left, &hr = d->h_right;
    hl.assign(nf, 0); hr.assign(nf, 0);
    for (i64 f = 0; f < nf; f++) {
        if (left[f] < 1 || left[f] > nc || right[f] < 1 || right[f] > nc) { delete d; JB_FAIL(ctx, JB_ERR_ARG, "jb_nfvm_create: cell out of range"); }
        hl[f] = (int32_t)(left[f] - 1); hr[f] = (int32_t)(right[f] - 1);
    }
    int rc = (d->left.upload(hl, ctx->stream) == cudaSuccess && d->right.upload(hr, ctx->stream) == cudaSuccess) ? JB_OK : JB_ERR_ALLOC;
    if (rc == JB_OK) rc = nfvm_half(ctx, nf, nc, L_Tl, L_Tr, L_ptr, L_cell, L_T, d->L_Tl, d->L_Tr, d->L_ptr, d->L_cell, d->L_T,
                                    d->h_Lptr, d->h_Lcell, d->h_LTl, d->h_LTr, d->h_LT);
    if (rc == JB_OK && scheme > 0) rc = nfvm_half(ctx, nf, nc, R_Tl, R_Tr, R_ptr, R_cell, R_T, d->R_Tl, d->R_Tr, d->R_ptr, d->R_cell, d->R_T,
                                                  d->h_Rptr, d->h_Rcell, d->h_RTl, d->h_RTr, d->h_RT);
    if (rc != JB_OK) { delete d; JB_FAIL(ctx, rc, "jb_nfvm_create: bad stencil arrays or allocation failure"); }
    *out = d;
    return JB_OK;
}
int32_t jb_nfvm_destroy(jb_nfvm* d) { delete d; return JB_OK; }
int32_t jb_nfvm_evaluate_flux(jb_nfvm* d, const double* d_p, int64_t nph, int64_t ph, double* d_q) {
    if (!d || !d_p || !d_q || nph < 1 || ph < 1 || ph > nph) return JB_ERR_ARG;
    jb_ctx* ctx = d->ctx;
    ProfScope _ps(ctx, JB_PROF_ASSEMBLY);
    nfvm_flux_kernel<<<sgrid(ctx, d->nf), 256, 0, ctx->stream>>>(d->nf, d->scheme, d->left.p, d->right.p, d->L_Tl.p, d->L_Tr.p, d->L_ptr.p, d->L_cell.p,
                                                                d->L_T.p, d->R_Tl.p, d->R_Tr.p, d->R_ptr.p, d->R_cell.p, d->R_T.p, d_p, nph, ph - 1, d_q);
    JB_CHECK_LAUNCH(ctx);
    JB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return JB_OK;
}

}  // extern "C"

// ---- conservation law on the NFVM flux with the face-based assembly (PotentialFlow{:fvm}) -------------------------------
// Reference: ConservationLawFiniteVolumeStorage (src/conservation/fvm_assembly.jl:1-53), declare_pattern (:55-89),
// align_to_jacobian! (:98-165), fvm_update_face_fluxes_inner! (:194-208: the face flux is evaluated once per stencil cell with
// that cell carrying the partial), update_linearized_system_equation! / fvm_face_assembly! (:216-283): zero, r = accumulation,
// J[c,c] = d(accumulation), then per face r[l] += q, r[r] -= q, J[l,cell] += dq/dp_cell, J[r,cell] -= dq/dp_cell.
//
// evaluate_flux (src/NFVM/evaluation.jl:1-88) is differentiated in closed form instead of once per stencil cell: with
// q_l = <cLq, p>, r_l = <cLr, p>, q_R = <cRq, p>, r_R = <cRr, p> (q_r = -q_R, r_r = -r_R) and q = mu_l q_l - mu_r q_r,
//   dq/dp_s = mu_l cLq_s + mu_r cRq_s + (q_l + q_r) dmu_l/dp_s,   dmu_l = (r_lw d r_rw - r_rw d r_lw) / r_total^2,
// d r_lw = sigma_l cLr_s, d r_rw = -sigma_r cRr_s (sigma = 1 for :ntpfa, the sign ForwardDiff gives abs for :nmpfa), and
// dmu = 0 inside the |r_total| < 1e-10 guard. One lane per face, one pass over the face's stencil slots; the scatter uses the
// warp-aggregated atomics of jb_reduce.cuh (faces of one cell sit in neighbouring lanes).
__device__ __forceinline__ double4 ld_coef(const double4* __restrict__ p) {
    const double2* q = reinterpret_cast<const double2*>(p);
    const double2 a = __ldg(q), b = __ldg(q + 1);
    return make_double4(a.x, a.y, b.x, b.y);
}
// zero the touched Jacobian entries (here: all of them, the law is the only equation of the system) and start the residual
// from the accumulation term (fvm_assembly.jl:216-236)
__global__ void __launch_bounds__(256) nfvm_law_init_kernel(i64 nc, i64 nnz, const double* __restrict__ acc, double* __restrict__ nz,
                                                            double* __restrict__ r) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += (i64)gridDim.x * blockDim.x) nz[i] = 0.0;
    for (i64 c = (i64)blockIdx.x * blockDim.x + threadIdx.x; c < nc; c += (i64)gridDim.x * blockDim.x) r[c] = acc ? __ldg(acc + c) : 0.0;
}
__global__ void __launch_bounds__(256) nfvm_law_diag_kernel(i64 nc, const int32_t* __restrict__ diag, const double* __restrict__ dacc, double* __restrict__ nz) {
    for (i64 c = (i64)blockIdx.x * blockDim.x + threadIdx.x; c < nc; c += (i64)gridDim.x * blockDim.x) nz[__ldg(diag + c)] = __ldg(dacc + c);
}
__global__ void __launch_bounds__(256) nfvm_law_faces_kernel(i64 nf, int scheme, const int32_t* __restrict__ left, const int32_t* __restrict__ right,
                                                             const int32_t* __restrict__ vpos, const int32_t* __restrict__ vars,
                                                             const double4* __restrict__ coef, const int32_t* __restrict__ lpos,
                                                             const int32_t* __restrict__ rpos, const double* __restrict__ p, i64 nph, i64 ph0,
                                                             double* __restrict__ nz, double* __restrict__ r, double* __restrict__ qout) {
    for (i64 f = (i64)blockIdx.x * blockDim.x + threadIdx.x; f < nf; f += (i64)gridDim.x * blockDim.x) {
        const int32_t s0 = __ldg(vpos + f), s1 = __ldg(vpos + f + 1);
        double q_l = 0.0, r_l = 0.0, q_R = 0.0, r_R = 0.0;
        for (int32_t s = s0; s < s1; s++) {
            const double ps = __ldg(p + (size_t)__ldg(vars + s) * nph + ph0);
            const double4 c = ld_coef(coef + s);
            q_l = fma(c.x, ps, q_l); r_l = fma(c.y, ps, r_l); q_R = fma(c.z, ps, q_R); r_R = fma(c.w, ps, r_R);
        }
        double a = 1.0, cc = 0.0, b = 0.0, d = 0.0, q = q_l;
        if (scheme != 0) {
            const double q_r = -q_R, r_r = -r_R;
            double r_lw = r_l, r_rw = r_r, sg_l = 1.0, sg_r = 1.0;
            if (scheme == 2) { r_lw = fabs(r_l); r_rw = fabs(r_r); sg_l = signbit(r_l) ? -1.0 : 1.0; sg_r = signbit(r_r) ? -1.0 : 1.0; }
            const double r_total = r_lw + r_rw;
            double mu_l = 0.5, mu_r = 0.5;
            if (!(fabs(r_total) < 1e-10)) {
                mu_l = r_rw / r_total; mu_r = r_lw / r_total;
                const double w = (q_l + q_r) / (r_total * r_total);
                b = -w * r_rw * sg_l;      // multiplies cLr_s
                d = -w * r_lw * sg_r;      // multiplies cRr_s
            }
            a = mu_l; cc = mu_r;
            q = mu_l * q_l - mu_r * q_r;
        }
        if (qout) qout[f] = q;
        const int32_t l = __ldg(left + f), rr = __ldg(right + f);
        warp_agg_atomic_add(r + l, q);
        warp_agg_atomic_add(r + rr, -q);
        for (int32_t s = s0; s < s1; s++) {
            const double4 c = ld_coef(coef + s);
            const double dq = a * c.x + b * c.y + cc * c.z + d * c.w;
            atomicAdd(nz + __ldg(lpos + s), dq);
            atomicAdd(nz + __ldg(rpos + s), -dq);
        }
    }
}

// discretization_stencil of every face: unique!([left, right, mpfa cells of ft_left..., (left, right,) mpfa cells of ft_right...])
static int nfvm_build_stencil(jb_nfvm* d) {
    if (!d->h_vpos.empty()) return JB_OK;
    const i64 nf = d->nf;
    d->h_vpos.assign(nf + 1, 0);
    std::vector<double4> coef;
    std::vector<int32_t> cells;
    auto slot_of = [&](int32_t first, int32_t c) -> int32_t {
        for (int32_t s = first; s < (int32_t)d->h_vars.size(); s++) if (d->h_vars[s] == c) return s;
        d->h_vars.push_back(c); coef.push_back(make_double4(0.0, 0.0, 0.0, 0.0));
        return (int32_t)d->h_vars.size() - 1;
    };
    for (i64 f = 0; f < nf; f++) {
        const int32_t first = (int32_t)d->h_vars.size();
        const int32_t sl = slot_of(first, d->h_left[f]), sr = slot_of(first, d->h_right[f]);
        coef[sl].x += d->h_LTl[f]; coef[sr].x += d->h_LTr[f];
        for (int32_t k = d->h_Lptr[f]; k < d->h_Lptr[f + 1]; k++) { const int32_t s = slot_of(first, d->h_Lcell[k]); coef[s].x += d->h_LT[k]; coef[s].y += d->h_LT[k]; }
        if (d->scheme > 0) {
            coef[sl].z += d->h_RTl[f]; coef[sr].z += d->h_RTr[f];
            for (int32_t k = d->h_Rptr[f]; k < d->h_Rptr[f + 1]; k++) { const int32_t s = slot_of(first, d->h_Rcell[k]); coef[s].z += d->h_RT[k]; coef[s].w += d->h_RT[k]; }
        }
        d->h_vpos[f + 1] = (int32_t)d->h_vars.size();
    }
    if (d->d_coef.upload(coef, d->ctx->stream) != cudaSuccess || d->d_vpos.upload(d->h_vpos, d->ctx->stream) != cudaSuccess ||
        d->d_vars.upload(d->h_vars, d->ctx->stream) != cudaSuccess) {
        d->h_vpos.clear(); d->h_vars.clear();
        JB_FAIL(d->ctx, JB_ERR_ALLOC, "jb_nfvm: allocation of the face stencils failed");
    }
    return JB_OK;
}

extern "C" {

// parity dump of the face stencils (face_cache.vpos / .variables): vpos[nf+1] and vars[vpos[nf]-1], 1-based
int32_t jb_nfvm_stencil(jb_nfvm* d, int64_t* vpos, int64_t* vars, int64_t cap, int64_t* n_out) {
    if (!d) return JB_ERR_ARG;
    { const int rc_ = nfvm_build_stencil(d); if (rc_ != JB_OK) return rc_; }
    if (n_out) *n_out = (int64_t)d->h_vars.size();
    if (vpos) for (i64 f = 0; f <= d->nf; f++) vpos[f] = d->h_vpos[f] + 1;
    if (vars) {
        if (cap < (int64_t)d->h_vars.size()) return JB_ERR_ARG;
        for (size_t s = 0; s < d->h_vars.size(); s++) vars[s] = d->h_vars[s] + 1;
    }
    return JB_OK;
}
// declare_pattern (fvm_assembly.jl:55-89): diagonal, and for every face and stencil cell c: (l,c), (r,c), (c,l), (c,r)
int32_t jb_nfvm_pattern(jb_nfvm* d, jb_csr** out) {
    if (!d || !out) return JB_ERR_ARG;
    { const int rc_ = nfvm_build_stencil(d); if (rc_ != JB_OK) return rc_; }
    std::vector<int64_t> I, J;
    I.reserve(d->nc + 4 * d->h_vars.size()); J.reserve(d->nc + 4 * d->h_vars.size());
    for (i64 c = 1; c <= d->nc; c++) { I.push_back(c); J.push_back(c); }
    for (i64 f = 0; f < d->nf; f++) {
        const int64_t l = d->h_left[f] + 1, r = d->h_right[f] + 1;
        for (int32_t s = d->h_vpos[f]; s < d->h_vpos[f + 1]; s++) {
            const int64_t c = d->h_vars[s] + 1;
            I.push_back(l); J.push_back(c); I.push_back(r); J.push_back(c);
            I.push_back(c); J.push_back(l); I.push_back(c); J.push_back(r);
        }
    }
    return jb_csr_create_from_coo(d->ctx, I.data(), J.data(), (int64_t)I.size(), d->nc, 1, out);
}
// align_to_jacobian! (fvm_assembly.jl:98-165): left / right positions of every stencil slot
int32_t jb_nfvm_align(jb_nfvm* d, jb_csr* A) {
    if (!d || !A || A->bs != 1 || A->n != d->nc) return JB_ERR_ARG;
    { const int rc_ = nfvm_build_stencil(d); if (rc_ != JB_OK) return rc_; }
    jb_ctx* ctx = d->ctx;
    auto find = [&](int32_t row, int32_t col) -> int32_t {
        const int32_t* b = A->h_colidx.data() + A->h_rowptr[row];
        const int32_t* e = A->h_colidx.data() + A->h_rowptr[row + 1];
        const int32_t* it = std::lower_bound(b, e, col);
        return (it != e && *it == col) ? (int32_t)(it - A->h_colidx.data()) : -1;
    };
    d->h_lpos.assign(d->h_vars.size(), -1); d->h_rpos.assign(d->h_vars.size(), -1);
    for (i64 f = 0; f < d->nf; f++)
        for (int32_t s = d->h_vpos[f]; s < d->h_vpos[f + 1]; s++) {
            d->h_lpos[s] = find(d->h_left[f], d->h_vars[s]); d->h_rpos[s] = find(d->h_right[f], d->h_vars[s]);
            if (d->h_lpos[s] < 0 || d->h_rpos[s] < 0) JB_FAIL(ctx, JB_ERR_ARG, "jb_nfvm_align: Jacobian alignment failed (entry not allocated)");
        }
    for (i64 c = 0; c < d->nc; c++) if (A->h_diag[c] < 0) JB_FAIL(ctx, JB_ERR_ARG, "jb_nfvm_align: diagonal entry not allocated");
    if (d->d_lpos.upload(d->h_lpos, ctx->stream) != cudaSuccess || d->d_rpos.upload(d->h_rpos, ctx->stream) != cudaSuccess)
        JB_FAIL(ctx, JB_ERR_ALLOC, "jb_nfvm_align: allocation failed");
    d->csr = A;
    return JB_OK;
}
int32_t jb_nfvm_positions(jb_nfvm* d, int64_t* left_pos, int64_t* right_pos) {
    if (!d || !d->csr) return JB_ERR_ARG;
    for (size_t s = 0; s < d->h_vars.size(); s++) { if (left_pos) left_pos[s] = d->h_lpos[s] + 1; if (right_pos) right_pos[s] = d->h_rpos[s] + 1; }
    return JB_OK;
}
// update_equation! + update_linearized_system_equation! for the law  acc_c + sum_faces (+/-) q_f(p) = 0:
// d_acc / d_dacc (nc, may be NULL = 0): value and d/dp of the accumulation term; writes nonzeros(jac), d_r (nc) and,
// if d_q != NULL, the face fluxes.
int32_t jb_nfvm_assemble(jb_nfvm* d, const double* d_p, int64_t nph, int64_t ph, const double* d_acc, const double* d_dacc, double* d_r,
                         double* d_q) {
    if (!d || !d->csr || !d_p || !d_r || nph < 1 || ph < 1 || ph > nph) return JB_ERR_ARG;
    jb_ctx* ctx = d->ctx;
    jb_csr* A = d->csr;
    {
        ProfScope _ps(ctx, JB_PROF_ASSEMBLY);
        nfvm_law_init_kernel<<<sgrid(ctx, std::max(A->nnzb, d->nc)), 256, 0, ctx->stream>>>(d->nc, A->nnzb, d_acc, A->d_val.p, d_r);
        JB_CHECK_LAUNCH(ctx);
        if (d_dacc) { nfvm_law_diag_kernel<<<sgrid(ctx, d->nc), 256, 0, ctx->stream>>>(d->nc, A->d_diag.p, d_dacc, A->d_val.p); JB_CHECK_LAUNCH(ctx); }
        nfvm_law_faces_kernel<<<sgrid(ctx, d->nf), 256, 0, ctx->stream>>>(d->nf, d->scheme, d->left.p, d->right.p, d->d_vpos.p, d->d_vars.p, d->d_coef.p,
                                                                         d->d_lpos.p, d->d_rpos.p, d_p, nph, ph - 1, A->d_val.p, d_r, d_q);
        JB_CHECK_LAUNCH(ctx);
    }
    jb_csr_touch(A);
    JB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return JB_OK;
}

}  // extern "C"
