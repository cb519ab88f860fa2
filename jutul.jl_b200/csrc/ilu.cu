// ILU(0) numeric factorisation and triangular solves, level-scheduled for the GPU.
//
// Semantics are those of ilu0_factor! / ilu_solve! (src/StaticCSR/ilu0.jl:108-193)
// and of ParallelILUFactorCSR over a partition (src/StaticCSR/par_ilu0.jl:47-87):
// row-by-row IKJ elimination in ascending row order inside each block, L holding
// the multipliers L_ik * inv(D_k), D stored inverted, couplings between blocks
// dropped. The sequential row loop is replaced by dependency levels computed at
// setup (host_setup.cu): rows of a level are independent, one launch per level,
// one thread per row. Within a row the L entries are visited in ascending column
// order and every update L_ij / D_i / U_ij -= m * U_kj is applied from a
// precomputed (target, source) list, so the arithmetic per row equals the
// reference's. Storage is level-major (rows of one level are contiguous in the
// L / U value arrays) so each level streams its part of the factor once.
#include "jb_internal.cuh"
#include "jb_krylov_scalars.cuh"
#include "jb_stream.cuh"
#include "jb_stream2.cuh"

template <int BS>
__global__ void __launch_bounds__(256) ilu_gather_kernel(i64 nL, i64 n, i64 nU, const int32_t* __restrict__ Lmap,
                                                         const int32_t* __restrict__ Dmap, const int32_t* __restrict__ Umap,
                                                         const double* __restrict__ A, double* __restrict__ fv, i64 first_block) {
    // update_values! (src/StaticCSR/ilu0.jl:83-98): refresh L, D, U from the Jacobian through the stored maps.
    // first_block = nL skips the L part (the two-colour stream factorisation reads L straight from the Jacobian).
    constexpr int B2 = BS * BS;
    const i64 total = (nL + n + nU) * B2;
    for (i64 idx = first_block * B2 + (i64)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (i64)gridDim.x * blockDim.x) {
        const i64 blk = idx / B2;
        const int q = (int)(idx % B2);
        int32_t src;
        if (blk < nL) src = __ldg(Lmap + blk);
        else if (blk < nL + n) src = __ldg(Dmap + (blk - nL));
        else src = __ldg(Umap + (blk - nL - n));
        fv[idx] = __ldg(A + (size_t)src * B2 + q);
    }
}

template <int BS>
__global__ void __launch_bounds__(128) ilu_factor_level_kernel(int32_t t0, int32_t t1, i64 nL, const int32_t* __restrict__ forder,
                                                               const int32_t* __restrict__ Lstart, const int32_t* __restrict__ Lend,
                                                               const int32_t* __restrict__ Lcol, const int32_t* __restrict__ upd_ptr,
                                                               const int32_t* __restrict__ upd_tgt, const int32_t* __restrict__ upd_src,
                                                               double* __restrict__ fv, double* __restrict__ dinv, int32_t* status) {
    constexpr int B2 = BS * BS;
    const int32_t t = t0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= t1) return;
    const int32_t i = __ldg(forder + t);
    const int32_t l0 = __ldg(Lstart + t), l1 = __ldg(Lend + t);   // level-ordered copies: contiguous in t
    for (int32_t li = l0; li < l1; li++) {
        const int32_t k = __ldg(Lcol + li);
        double Lik[B2], Dk[B2], m[B2];
#pragma unroll
        for (int q = 0; q < B2; q++) { Lik[q] = fv[(size_t)li * B2 + q]; Dk[q] = dinv[(size_t)k * B2 + q]; }
        blk_mul<BS>(Lik, Dk, m);   // A_ik = L_ik * inv(D_kk)
#pragma unroll
        for (int q = 0; q < B2; q++) fv[(size_t)li * B2 + q] = m[q];
        const int32_t u0 = __ldg(upd_ptr + li), u1 = __ldg(upd_ptr + li + 1);
        for (int32_t u = u0; u < u1; u++) {
            const size_t tg = (size_t)__ldg(upd_tgt + u) * B2, sr = (size_t)__ldg(upd_src + u) * B2;
            double Ukj[B2], prod[B2];
#pragma unroll
            for (int q = 0; q < B2; q++) Ukj[q] = fv[sr + q];
            blk_mul<BS>(m, Ukj, prod);
#pragma unroll
            for (int q = 0; q < B2; q++) fv[tg + q] -= prod[q];
        }
    }
    double D[B2], Di[B2];
#pragma unroll
    for (int q = 0; q < B2; q++) D[q] = fv[(size_t)(nL + i) * B2 + q];
    blk_inv<BS>(D, Di);
    bool bad = false;
#pragma unroll
    for (int q = 0; q < B2; q++) { dinv[(size_t)i * B2 + q] = Di[q]; bad |= !isfinite(Di[q]); }
    if (bad) *status = JB_BAD_PIVOT;
}

// Two-colour numeric factorisation of the second colour in row-chunk stream form (jb_stream.cuh skeleton). Rows of the
// second colour have L entries only, towards rows k of the first colour whose D_k^{-1} is final, and the only update of
// the IKJ elimination is D_i -= m_ik U_ki. All threads stream the chunk's L entries: m_ik = A_ik D_k^{-1} is read from
// the Jacobian through Lmap (no intermediate copy of L), written to the factor, multiplied with U_ki and parked in shared
// memory; one thread per row then subtracts the products from A_ii in ascending k (the order of ilu0_factor!,
// src/StaticCSR/ilu0.jl:108-144) and inverts the block.
template <int BS>
__global__ void __launch_bounds__(256) ilu_factor_rb_kernel(int c0, int c1, const int32_t* __restrict__ chunk_ptr, const int32_t* __restrict__ ptrT,
                                                            const int32_t* __restrict__ Lcol, const int32_t* __restrict__ Lmap,
                                                            const int32_t* __restrict__ usrc, const int32_t* __restrict__ forder,
                                                            const int32_t* __restrict__ Dmap, i64 nL, const double* __restrict__ A,
                                                            double* fv, double* dinv, int32_t* status) {
    constexpr int B2 = BS * BS;
    __shared__ int32_t s_rp[JB_CHUNK_ROWS + 1];
    extern __shared__ double s_prod[];   // B2 doubles per L entry of the chunk
    for (int c = c0 + blockIdx.x; c < c1; c += gridDim.x) {
        const int t0 = __ldg(chunk_ptr + c), nr = __ldg(chunk_ptr + c + 1) - t0;
        for (int j = threadIdx.x; j <= nr; j += blockDim.x) s_rp[j] = __ldg(ptrT + t0 + j);
        __syncthreads();
        const int base = s_rp[0], cnt = s_rp[nr] - base;
        // two entries per thread in flight: indices first, then the three gathers of both entries, then the arithmetic
        constexpr int UF = 2;
        for (int eb = threadIdx.x; eb < cnt; eb += blockDim.x * UF) {
            int32_t kk[UF], src[UF], us[UF];
            double Lik[UF][B2], Dk[UF][B2], Ukj[UF][B2];
#pragma unroll
            for (int u = 0; u < UF; u++) {
                const int e = eb + u * blockDim.x;
                if (e < cnt) { const size_t li = (size_t)base + e; kk[u] = __ldcs(Lcol + li); src[u] = __ldcs(Lmap + li); us[u] = __ldcs(usrc + li); }
            }
#pragma unroll
            for (int u = 0; u < UF; u++) {
                const int e = eb + u * blockDim.x;
                if (e < cnt) {
#pragma unroll
                    for (int q = 0; q < B2; q++) {
                        Lik[u][q] = __ldcs(A + (size_t)src[u] * B2 + q);
                        Dk[u][q] = dinv[(size_t)kk[u] * B2 + q];
                        Ukj[u][q] = us[u] >= 0 ? fv[(size_t)us[u] * B2 + q] : 0.0;
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < UF; u++) {
                const int e = eb + u * blockDim.x;
                if (e < cnt) {
                    const size_t li = (size_t)base + e;
                    double m[B2], prod[B2];
                    blk_mul<BS>(Lik[u], Dk[u], m);
                    blk_mul<BS>(m, Ukj[u], prod);
#pragma unroll
                    for (int q = 0; q < B2; q++) { __stcs(fv + li * B2 + q, m[q]); s_prod[(size_t)e * B2 + q] = us[u] >= 0 ? prod[q] : 0.0; }
                }
            }
        }
        __syncthreads();
        if ((int)threadIdx.x < nr) {
            const size_t i = (size_t)__ldg(forder + t0 + threadIdx.x);
            const size_t dsrc = (size_t)__ldg(Dmap + i);
            double D[B2], Di[B2];
#pragma unroll
            for (int q = 0; q < B2; q++) D[q] = __ldg(A + dsrc * B2 + q);
            for (int e = s_rp[threadIdx.x] - base; e < s_rp[threadIdx.x + 1] - base; e++) {
#pragma unroll
                for (int q = 0; q < B2; q++) D[q] -= s_prod[(size_t)e * B2 + q];
            }
            blk_inv<BS>(D, Di);
            bool bad = false;
#pragma unroll
            for (int q = 0; q < B2; q++) { fv[((size_t)nL + i) * B2 + q] = D[q]; dinv[i * B2 + q] = Di[q]; bad |= !isfinite(Di[q]); }
            if (bad) *status = JB_BAD_PIVOT;
        }
        __syncthreads();
    }
}


// Single-pass two-colour numeric refactorisation (2x2 blocks): ONE sweep over the Jacobian in row chunks produces L, D^-1 and
// U. Everything a row needs is in the Jacobian itself: first-colour rows k have no L entries, so D_k = A_kk and U_k* = A_k*
// (copied to their slots); a second-colour row i needs, per L entry, m_ik = A_ik inv(A_kk) and D_i -= m_ik A_ki — A_kk and
// A_ki are gathered from the Jacobian (rows k are spatial neighbours: the chunk order interleaves the two colours by
// position so those rows were streamed moments ago and sit in L2). No intermediate copy of D / U, no second pass over L,
// and no dependency between rows, hence no levels. Arithmetic per entry is that of ilu_factor_level_kernel /
// ilu_factor_rb_kernel (blk_inv, blk_mul, products subtracted in ascending column order; src/StaticCSR/ilu0.jl:108-144).
// Thread per block of the chunk; products parked in shared memory; one thread per row finishes D_i.
#define JB_RB2_ROWS 128
#define JB_RB2_CAP 1024
__global__ void __launch_bounds__(256, 4) ilu_factor_rb2_kernel(int nchunks, const int32_t* __restrict__ chunks, const int32_t* __restrict__ rowptr,
                                                                const int32_t* __restrict__ diag, const int32_t* __restrict__ fdst,
                                                                const int2* __restrict__ gL, int32_t kB0, int32_t kB1, const unsigned char* __restrict__ rtype, i64 nL,
                                                                const double* __restrict__ A, double* __restrict__ fv, double* __restrict__ dinv,
                                                                int32_t* status) {
    __shared__ int32_t s_rp[JB_RB2_ROWS + 1];
    __shared__ __align__(16) double s_prod[JB_RB2_CAP * 4];   // 4 doubles per block of the chunk
    constexpr int UF = 2;     // blocks in flight per thread: slot, gather indices and the block itself are loaded independently, then A_kk / A_ki;
                              // all loads of both blocks are issued before any arithmetic
    for (int c = blockIdx.x; c < nchunks; c += gridDim.x) {
        const int r0 = __ldg(chunks + 2 * c), nr = __ldg(chunks + 2 * c + 1) - r0;
        for (int j = threadIdx.x; j <= nr; j += blockDim.x) s_rp[j] = __ldg(rowptr + r0 + j);
        __syncthreads();
        const int base = s_rp[0], cnt = s_rp[nr] - base;
        for (int eb = threadIdx.x; eb < cnt; eb += blockDim.x * UF) {
            int32_t dst[UF];
            double2 a0[UF], a1[UF], k0[UF], k1[UF], u0[UF], u1[UF];
            int2 g[UF];
#pragma unroll
            for (int u = 0; u < UF; u++) {
                const int e = eb + u * blockDim.x;
                dst[u] = -1;
                g[u] = make_int2(-1, -1);
                if (e < cnt) {
                    const size_t kA = (size_t)base + e;
                    dst[u] = __ldg(fdst + kA);
                    if ((int32_t)kA >= kB0 && (int32_t)kA < kB1) g[u] = __ldg(gL + (kA - (size_t)kB0));
                    const double2* ap = reinterpret_cast<const double2*>(A + kA * 4);
                    a0[u] = ap[0]; a1[u] = ap[1];
                }
            }
#pragma unroll
            for (int u = 0; u < UF; u++) {
                k0[u] = k1[u] = u0[u] = u1[u] = make_double2(0.0, 0.0);
                if (g[u].x >= 0) { const double2* kp = reinterpret_cast<const double2*>(A + (size_t)g[u].x * 4); k0[u] = kp[0]; k1[u] = kp[1]; }
                if (g[u].y >= 0) { const double2* up = reinterpret_cast<const double2*>(A + (size_t)g[u].y * 4); u0[u] = up[0]; u1[u] = up[1]; }
            }
#pragma unroll
            for (int u = 0; u < UF; u++) {
                const int e = eb + u * blockDim.x;
                if (e >= cnt) continue;
                double prod[4] = {0.0, 0.0, 0.0, 0.0};
                if (dst[u] == -2) { prod[0] = a0[u].x; prod[1] = a0[u].y; prod[2] = a1[u].x; prod[3] = a1[u].y; }   // diagonal block: parked for the row thread
                else if (dst[u] >= 0) {
                    double2* out = reinterpret_cast<double2*>(fv + (size_t)dst[u] * 4);
                    if ((i64)dst[u] >= nL) { __stcs(out, a0[u]); __stcs(out + 1, a1[u]); }            // U entry of a first-colour row: copy
                    else {                                                                             // L entry of a second-colour row
                        const double Lik[4] = {a0[u].x, a0[u].y, a1[u].x, a1[u].y}, Akk[4] = {k0[u].x, k0[u].y, k1[u].x, k1[u].y};
                        const double Uki[4] = {u0[u].x, u0[u].y, u1[u].x, u1[u].y};
                        double Dk[4], m[4];
                        blk_inv<2>(Akk, Dk);
                        blk_mul<2>(Lik, Dk, m);
                        __stcs(out, make_double2(m[0], m[1])); __stcs(out + 1, make_double2(m[2], m[3]));
                        if (g[u].y >= 0) blk_mul<2>(m, Uki, prod);
                    }
                }
                double2* sp = reinterpret_cast<double2*>(s_prod + (size_t)e * 4);
                sp[0] = make_double2(prod[0], prod[1]); sp[1] = make_double2(prod[2], prod[3]);
            }
        }
        __syncthreads();
        if ((int)threadIdx.x < nr) {
            const size_t i = (size_t)r0 + threadIdx.x;
            const int ed = __ldg(diag + i) - base;
            double D[4], Di[4];
#pragma unroll
            for (int q = 0; q < 4; q++) D[q] = s_prod[(size_t)ed * 4 + q];
            if (__ldg(rtype + i)) {
                for (int e = s_rp[threadIdx.x] - base; e < s_rp[threadIdx.x + 1] - base; e++) {
                    if (e == ed) continue;
#pragma unroll
                    for (int q = 0; q < 4; q++) D[q] -= s_prod[(size_t)e * 4 + q];
                }
            }
            blk_inv<2>(D, Di);
            bool bad = false;
            double2* fd = reinterpret_cast<double2*>(fv + ((size_t)nL + i) * 4);
            double2* dd = reinterpret_cast<double2*>(dinv + i * 4);
            fd[0] = make_double2(D[0], D[1]); fd[1] = make_double2(D[2], D[3]);
            dd[0] = make_double2(Di[0], Di[1]); dd[1] = make_double2(Di[2], Di[3]);
#pragma unroll
            for (int q = 0; q < 4; q++) bad |= !isfinite(Di[q]);
            if (bad) *status = JB_BAD_PIVOT;
        }
        __syncthreads();
    }
}

// forward sweep: x_i = b_i - sum_j L_ij x_j  (unit diagonal), rows of one level
template <int BS>
__global__ void __launch_bounds__(256) ilu_forward_level_kernel(int32_t t0, int32_t t1, const int32_t* __restrict__ forder,
                                                                const int32_t* __restrict__ Lstart, const int32_t* __restrict__ Lend,
                                                                const int32_t* __restrict__ Lcol, const double* __restrict__ fv,
                                                                const double* __restrict__ b, double* x, const double* sc) {
    constexpr int B2 = BS * BS;
    if (sc && sc[KS_DONE] != 0.0) return;
    const int32_t t = t0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= t1) return;
    const int32_t i = __ldg(forder + t);
    double v[BS];
#pragma unroll
    for (int e = 0; e < BS; e++) v[e] = b[(size_t)i * BS + e];
    const int32_t l0 = __ldg(Lstart + t), l1 = __ldg(Lend + t);   // level-ordered copies: contiguous in t
    for (int32_t li = l0; li < l1; li++) {
        const int32_t j = __ldg(Lcol + li);
        double a[B2], xj[BS];
#pragma unroll
        for (int q = 0; q < B2; q++) a[q] = __ldg(fv + (size_t)li * B2 + q);
#pragma unroll
        for (int e = 0; e < BS; e++) xj[e] = x[(size_t)j * BS + e];
        blk_submulvec<BS>(a, xj, v);
    }
#pragma unroll
    for (int e = 0; e < BS; e++) x[(size_t)i * BS + e] = v[e];
}

// backward sweep: x_i = Dinv_i (x_i - sum_j U_ij x_j)
template <int BS>
__global__ void __launch_bounds__(256) ilu_backward_level_kernel(int32_t t0, int32_t t1, i64 baseU, const int32_t* __restrict__ border,
                                                                 const int32_t* __restrict__ Ustart, const int32_t* __restrict__ Uend,
                                                                 const int32_t* __restrict__ Ucol, const double* __restrict__ fv,
                                                                 const double* __restrict__ dinv, double* x, const double* sc) {
    constexpr int B2 = BS * BS;
    if (sc && sc[KS_DONE] != 0.0) return;
    const int32_t t = t0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= t1) return;
    const int32_t i = __ldg(border + t);
    double v[BS], out[BS], d[B2];
#pragma unroll
    for (int e = 0; e < BS; e++) v[e] = x[(size_t)i * BS + e];
    const int32_t u0 = __ldg(Ustart + t), u1 = __ldg(Uend + t);
    for (int32_t ui = u0; ui < u1; ui++) {
        const int32_t j = __ldg(Ucol + ui);
        double a[B2], xj[BS];
#pragma unroll
        for (int q = 0; q < B2; q++) a[q] = __ldg(fv + (size_t)(baseU + ui) * B2 + q);
#pragma unroll
        for (int e = 0; e < BS; e++) xj[e] = x[(size_t)j * BS + e];
        blk_submulvec<BS>(a, xj, v);
    }
#pragma unroll
    for (int q = 0; q < B2; q++) d[q] = __ldg(dinv + (size_t)i * B2 + q);
    blk_mulvec<BS>(d, v, out);
#pragma unroll
    for (int e = 0; e < BS; e++) x[(size_t)i * BS + e] = out[e];
}

// ---- row-chunk stream form of the sweeps (default; see jb_stream.cuh). One launch per level; a CTA streams the
//      L (or U) entries of its chunk with unit stride and gathers x of earlier levels. ----
template <int BS>
__global__ void __launch_bounds__(256) ilu_sweep_stream_kernel(int c0, int c1, const int32_t* __restrict__ chunk_ptr,
                                                               const int32_t* __restrict__ ptrT, const int32_t* __restrict__ col,
                                                               const double* __restrict__ fv, size_t val_block_offset,
                                                               const int32_t* __restrict__ order, const double* __restrict__ dinv,
                                                               const double* rhs, const double* gsrc, int apply_dinv, double* x, const double* sc) {
    // x_i = [D_i^{-1}] (rhs_i - sum_j M_ij gsrc_j).  forward: rhs = b, gsrc = x;  backward: rhs = gsrc = x, with D^{-1};
    // two-colour fused sweeps: rhs = b, gsrc = b (first) / x (second), both with D^{-1}.
    if (sc && sc[KS_DONE] != 0.0) return;
    __shared__ int32_t s_rp[JB_CHUNK_ROWS + 1];
    extern __shared__ double s_prod[];
    for (int c = c0 + blockIdx.x; c < c1; c += gridDim.x) {
        const int t0 = __ldg(chunk_ptr + c), nr = __ldg(chunk_ptr + c + 1) - t0;
        // own row ids and right-hand sides are independent of the products: issue them first
        int i = 0;
        double v[BS];
        if ((int)threadIdx.x < nr) {
            i = __ldg(order + t0 + threadIdx.x);
#pragma unroll
            for (int e = 0; e < BS; e++) v[e] = rhs[(size_t)i * BS + e];
        }
        stream_chunk_products<BS, JB_STREAM_U>(t0, nr, ptrT, col, fv, val_block_offset, gsrc, s_rp, s_prod);
        if ((int)threadIdx.x < nr) {
            double acc[BS];
            stream_row_sum<BS>(threadIdx.x, s_rp, s_prod, acc);
#pragma unroll
            for (int e = 0; e < BS; e++) v[e] -= acc[e];
            if (apply_dinv) {
                double d[BS * BS], out[BS];
#pragma unroll
                for (int q = 0; q < BS * BS; q++) d[q] = __ldg(dinv + (size_t)i * BS * BS + q);
                blk_mulvec<BS>(d, v, out);
#pragma unroll
                for (int e = 0; e < BS; e++) x[(size_t)i * BS + e] = out[e];
            } else {
#pragma unroll
                for (int e = 0; e < BS; e++) x[(size_t)i * BS + e] = v[e];
            }
        }
        __syncthreads();
    }
}

// ---- TMA-staged lane-pair form of the sweeps for 2x2 blocks (default when bs = 2): see jb_stream2.cuh ----
// x_i = [D_i^{-1}] (rhs_i - sum_j M_ij gsrc_j) over the table entries [c0, c1) of one level.
__global__ void __launch_bounds__(JB_S2_THREADS, 6) ilu_sweep_s2_kernel(int c0, int c1, const S2Chunk* __restrict__ table,
                                                                        const int32_t* __restrict__ ptrT, const int32_t* __restrict__ col,
                                                                        const double* __restrict__ fv, size_t val_block_offset,
                                                                        const int32_t* __restrict__ order, const double* __restrict__ dinv,
                                                                        const double* rhs, const double* gsrc, int apply_dinv, double* x,
                                                                        const double* sc) {
    if (sc && sc[KS_DONE] != 0.0) return;
    extern __shared__ __align__(128) unsigned char s2_raw[];
    S2Smem& sm = *reinterpret_cast<S2Smem*>(s2_raw);
    s2_prologue(sm, table, c0, c1, ptrT, col, fv, val_block_offset, order);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, jl = lane >> 1, e = lane & 1;
    const unsigned pairmask = 3u << (lane & ~1);
    int it = 0;
    for (int k = c0 + blockIdx.x; k < c1; k += gridDim.x, it++) {
        const int s = it & 1;
        jb_mbar_wait(&sm.bar[s], (uint32_t)(it >> 1) & 1u);
        const S2Stage& S = sm.stage[s];
        const int t0 = sm.meta[s].t0, nr = sm.meta[s].nr, e0 = sm.meta[s].e0;
        const int j = warp * 16 + jl;
        if (j < nr) {
            const int lrp = jb_span_lead<int32_t>((size_t)t0), lcol = jb_span_lead<int32_t>((size_t)e0);
            const size_t i = (size_t)S.ord[lrp + j];
            double v = rhs[i * 2 + e];                 // independent of the products: requested first
            double da = 0.0, db = 0.0;
            if (apply_dinv) { da = __ldg(dinv + i * 4 + e); db = __ldg(dinv + i * 4 + 2 + e); }
            v -= s2_row_sum<6>(S, lcol, S.rp[lrp + j] - e0, S.rp[lrp + j + 1] - e0, e, gsrc);
            if (apply_dinv) {
                const double other = __shfl_xor_sync(pairmask, v, 1);
                const double v0 = e ? other : v, v1 = e ? v : other;
                v = fma(db, v1, da * v0);               // (D^{-1} v)_e, column-major 2x2
            }
            x[i * 2 + e] = v;
        }
        s2_release(sm, s, table, k, c1, ptrT, col, fv, val_block_offset, order);
    }
}
static bool ilu_s2_enabled() {
    const char* e = getenv("JB_STREAM_VARIANT");
    return !(e && e[0] == '1');
}
static int ilu_s2_grid(jb_ctx* ctx, int nchunks) {
    static int per_sm = 0;
    if (per_sm == 0) {
        cudaFuncSetAttribute(ilu_sweep_s2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(S2Smem));
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ilu_sweep_s2_kernel, JB_S2_THREADS, sizeof(S2Smem)) != cudaSuccess || per_sm < 1)
            per_sm = 1;
    }
    return std::max(1, std::min(nchunks, ctx->sm_count * per_sm));
}

// Levels whose rows have no off-diagonal entry inside the block (e.g. the first colour of a multicolour ordering) need no
// gather at all: x_i = b_i (forward) or x_i = D_i^{-1} x_i (backward). One thread per row, full occupancy.
template <int BS, bool BACKWARD>
__global__ void __launch_bounds__(256) ilu_light_level_kernel(int32_t t0, int32_t t1, const int32_t* __restrict__ order, const double* __restrict__ dinv,
                                                              const double* b, double* x, const double* sc) {
    if (sc && sc[KS_DONE] != 0.0) return;
    const int32_t t = t0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= t1) return;
    const size_t i = (size_t)__ldg(order + t);
    double v[BS];
#pragma unroll
    for (int e = 0; e < BS; e++) v[e] = BACKWARD ? x[i * BS + e] : b[i * BS + e];
    if (BACKWARD) {
        double d[BS * BS], out[BS];
#pragma unroll
        for (int q = 0; q < BS * BS; q++) d[q] = __ldg(dinv + i * BS * BS + q);
        blk_mulvec<BS>(d, v, out);
#pragma unroll
        for (int e = 0; e < BS; e++) x[i * BS + e] = out[e];
    } else {
#pragma unroll
        for (int e = 0; e < BS; e++) x[i * BS + e] = v[e];
    }
}

template <int BS>
__global__ void __launch_bounds__(256) ilu_iso_kernel(int32_t n, const int32_t* __restrict__ rows, const double* __restrict__ dinv, const double* b,
                                                      double* x, const double* sc) {
    if (sc && sc[KS_DONE] != 0.0) return;
    const int32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const size_t i = (size_t)__ldg(rows + t);
    double v[BS], d[BS * BS], out[BS];
#pragma unroll
    for (int e = 0; e < BS; e++) v[e] = b[i * BS + e];
#pragma unroll
    for (int q = 0; q < BS * BS; q++) d[q] = __ldg(dinv + i * BS * BS + q);
    blk_mulvec<BS>(d, v, out);
#pragma unroll
    for (int e = 0; e < BS; e++) x[i * BS + e] = out[e];
}

template <int BS>
static int ilu_apply_stream_t(jb_ilu* F, const double* b, double* x, const double* sc) {
    jb_ctx* ctx = F->csr->ctx;
    ProfScope _ps(ctx, JB_PROF_ILU_APPLY);
    cudaStream_t s = ctx->stream;
    const size_t smem = (size_t)JB_CHUNK_CAP * BS * sizeof(double);
    static int per_sm = 0;
    if (per_sm == 0) {
        cudaFuncSetAttribute(ilu_sweep_stream_kernel<BS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ilu_sweep_stream_kernel<BS>, 256, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    }
    const int cap = ctx->sm_count * per_sm;
    const bool s2 = BS == 2 && F->s2_ok && ilu_s2_enabled();
    if (F->two_colour && s2) {
        {
            const int c0 = F->h_levF_s2[1], c1 = F->h_levF_s2[2];
            if (c1 > c0) {
                ilu_sweep_s2_kernel<<<ilu_s2_grid(ctx, c1 - c0), JB_S2_THREADS, sizeof(S2Smem), s>>>(c0, c1, F->d_s2F.p, F->d_LptrT.p, F->d_Lcol.p, F->d_fv.p, 0,
                                                                                                   F->d_forder.p, F->d_dinv.p, b, b, 1, x, sc);
                JB_CHECK_LAUNCH(ctx);
            }
        }
        {
            const int c0 = F->h_levB_s2[1], c1 = F->h_levB_s2[2];
            if (c1 > c0) {
                ilu_sweep_s2_kernel<<<ilu_s2_grid(ctx, c1 - c0), JB_S2_THREADS, sizeof(S2Smem), s>>>(c0, c1, F->d_s2B.p, F->d_UptrT.p, F->d_Ucol.p, F->d_fv.p,
                                                                                                   (size_t)(F->nL + F->n), F->d_border.p, F->d_dinv.p, b, x, 1, x, sc);
                JB_CHECK_LAUNCH(ctx);
            }
        }
        if (!F->h_iso.empty()) {   // isolated rows: x_i = D_i^{-1} b_i
            const int32_t ni = (int32_t)F->h_iso.size();
            ilu_iso_kernel<BS><<<(ni + 255) / 256, 256, 0, s>>>(ni, F->d_iso.p, F->d_dinv.p, b, x, sc);
            JB_CHECK_LAUNCH(ctx);
        }
        return JB_OK;
    }
    if (F->two_colour) {
        // Two-colour fast path (rows of the second forward level have no U entries, rows of the first have no L entries):
        //   second colour:  x_B = D_B^{-1} (b_B - L_BR b_R)        (its forward and backward steps fused; y_R = b_R is read in place)
        //   first colour:   x_R = D_R^{-1} (b_R - U_RB x_B)
        // two launches instead of four, no copy of b_R, no second pass over x_B; arithmetic identical to the general path.
        {
            const int c0 = F->h_levF_chunk[1], c1 = F->h_levF_chunk[2];
            ilu_sweep_stream_kernel<BS><<<std::max(1, std::min(c1 - c0, cap)), 256, smem, s>>>(c0, c1, F->d_chunksF.p, F->d_LptrT.p, F->d_Lcol.p, F->d_fv.p, 0,
                                                                                              F->d_forder.p, F->d_dinv.p, b, b, 1, x, sc);
            JB_CHECK_LAUNCH(ctx);
        }
        {
            const int c0 = F->h_levB_chunk[1], c1 = F->h_levB_chunk[2];
            ilu_sweep_stream_kernel<BS><<<std::max(1, std::min(c1 - c0, cap)), 256, smem, s>>>(c0, c1, F->d_chunksB.p, F->d_UptrT.p, F->d_Ucol.p, F->d_fv.p,
                                                                                              (size_t)(F->nL + F->n), F->d_border.p, F->d_dinv.p, b, x, 1, x, sc);
            JB_CHECK_LAUNCH(ctx);
        }
        if (!F->h_iso.empty()) {   // isolated rows: x_i = D_i^{-1} b_i
            const int32_t ni = (int32_t)F->h_iso.size();
            ilu_iso_kernel<BS><<<(ni + 255) / 256, 256, 0, s>>>(ni, F->d_iso.p, F->d_dinv.p, b, x, sc);
            JB_CHECK_LAUNCH(ctx);
        }
        return JB_OK;
    }
    for (int l = 0; l < F->nlevF; l++) {
        const int c0 = F->h_levF_chunk[l], c1 = F->h_levF_chunk[l + 1];
        if (c1 <= c0) continue;
        const int32_t t0 = F->h_levF_ptr[l], t1 = F->h_levF_ptr[l + 1];
        if (F->h_LptrT[t1] == F->h_LptrT[t0]) {
            ilu_light_level_kernel<BS, false><<<(t1 - t0 + 255) / 256, 256, 0, s>>>(t0, t1, F->d_forder.p, F->d_dinv.p, b, x, sc);
            JB_CHECK_LAUNCH(ctx);
            continue;
        }
        if (s2) {
            const int a0 = F->h_levF_s2[l], a1 = F->h_levF_s2[l + 1];
            ilu_sweep_s2_kernel<<<ilu_s2_grid(ctx, a1 - a0), JB_S2_THREADS, sizeof(S2Smem), s>>>(a0, a1, F->d_s2F.p, F->d_LptrT.p, F->d_Lcol.p, F->d_fv.p, 0,
                                                                                               F->d_forder.p, F->d_dinv.p, b, x, 0, x, sc);
            JB_CHECK_LAUNCH(ctx);
            continue;
        }
        ilu_sweep_stream_kernel<BS><<<std::min(c1 - c0, cap), 256, smem, s>>>(c0, c1, F->d_chunksF.p, F->d_LptrT.p, F->d_Lcol.p, F->d_fv.p, 0,
                                                                              F->d_forder.p, F->d_dinv.p, b, x, 0, x, sc);
        JB_CHECK_LAUNCH(ctx);
    }
    for (int l = 0; l < F->nlevB; l++) {
        const int c0 = F->h_levB_chunk[l], c1 = F->h_levB_chunk[l + 1];
        if (c1 <= c0) continue;
        const int32_t t0 = F->h_levB_ptr[l], t1 = F->h_levB_ptr[l + 1];
        if (F->h_UptrT[t1] == F->h_UptrT[t0]) {
            ilu_light_level_kernel<BS, true><<<(t1 - t0 + 255) / 256, 256, 0, s>>>(t0, t1, F->d_border.p, F->d_dinv.p, b, x, sc);
            JB_CHECK_LAUNCH(ctx);
            continue;
        }
        if (s2) {
            const int a0 = F->h_levB_s2[l], a1 = F->h_levB_s2[l + 1];
            ilu_sweep_s2_kernel<<<ilu_s2_grid(ctx, a1 - a0), JB_S2_THREADS, sizeof(S2Smem), s>>>(a0, a1, F->d_s2B.p, F->d_UptrT.p, F->d_Ucol.p, F->d_fv.p,
                                                                                               (size_t)(F->nL + F->n), F->d_border.p, F->d_dinv.p, x, x, 1, x, sc);
            JB_CHECK_LAUNCH(ctx);
            continue;
        }
        ilu_sweep_stream_kernel<BS><<<std::min(c1 - c0, cap), 256, smem, s>>>(c0, c1, F->d_chunksB.p, F->d_UptrT.p, F->d_Ucol.p, F->d_fv.p,
                                                                              (size_t)(F->nL + F->n), F->d_border.p, F->d_dinv.p, x, x, 1, x, sc);
        JB_CHECK_LAUNCH(ctx);
    }
    return JB_OK;
}

template <int BS>
static int ilu_factor_t(jb_ilu* F) {
    jb_ctx* ctx = F->csr->ctx;
    ProfScope _ps(ctx, JB_PROF_ILU_FACTOR);
    cudaStream_t s = ctx->stream;
    JB_CUDA(ctx, cudaMemsetAsync(F->d_status.p, 0, sizeof(int32_t), s));
    const i64 total = (F->nL + F->n + F->nU) * BS * BS;
    int grid = (int)std::min<i64>((total + 255) / 256, (i64)ctx->sm_count * 16);
    if (BS == 2 && F->rb2_ok && F->two_colour && !getenv("JB_ILU_FACTOR_GENERIC") && !getenv("JB_ILU_FACTOR_RB1")) {
        // single pass over the Jacobian (ilu_factor_rb2_kernel)
        static int per_sm2 = 0;
        if (per_sm2 == 0 && (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm2, ilu_factor_rb2_kernel, 256, 0) != cudaSuccess || per_sm2 < 1)) per_sm2 = 1;
        ilu_factor_rb2_kernel<<<std::max(1, std::min(F->n_rb2_chunks, ctx->sm_count * per_sm2)), 256, 0, s>>>(
            F->n_rb2_chunks, F->d_rb2_chunks.p, F->csr->d_rowptr.p, F->csr->d_diag.p, F->d_fdst.p, F->d_gL.p, F->rb2_kB0, F->rb2_kB0 + (int32_t)F->d_gL.n,
            F->d_rtype.p, F->nL, F->csr->d_val.p,
            F->d_fv.p, F->d_dinv.p, F->d_status.p);
        JB_CHECK_LAUNCH(ctx);
        return JB_OK;
    }
    if (BS <= 2 && F->rb_factor && F->two_colour && F->stream_ok && !getenv("JB_ILU_FACTOR_GENERIC")) {
        // two-colour stream factorisation: D and U from the Jacobian, first colour inverted, second colour in one stream pass
        constexpr int BSS = BS <= 2 ? BS : 1;
        ilu_gather_kernel<BS><<<std::max(grid, 1), 256, 0, s>>>(F->nL, F->n, F->nU, F->d_Lmap.p, F->d_Dmap.p, F->d_Umap.p, F->csr->d_val.p, F->d_fv.p,
                                                                F->nL);
        JB_CHECK_LAUNCH(ctx);
        {
            const int32_t t0 = F->h_levF_ptr[0], t1 = F->h_levF_ptr[1];
            if (t1 > t0) {
                ilu_factor_level_kernel<BS><<<(t1 - t0 + 127) / 128, 128, 0, s>>>(t0, t1, F->nL, F->d_forder.p, F->d_Lstart.p, F->d_Lend.p, F->d_Lcol.p,
                                                                                  F->d_upd_ptr.p, F->d_upd_tgt.p, F->d_upd_src.p, F->d_fv.p, F->d_dinv.p,
                                                                                  F->d_status.p);
                JB_CHECK_LAUNCH(ctx);
            }
        }
        const int c0 = F->h_levF_chunk[1], c1 = F->h_levF_chunk[2];
        if (c1 > c0) {
            const size_t smem = (size_t)JB_CHUNK_CAP * BSS * BSS * sizeof(double);
            static int per_sm = 0;
            if (per_sm == 0) {
                cudaFuncSetAttribute(ilu_factor_rb_kernel<BSS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ilu_factor_rb_kernel<BSS>, 256, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
            }
            ilu_factor_rb_kernel<BSS><<<std::max(1, std::min(c1 - c0, ctx->sm_count * per_sm)), 256, smem, s>>>(
                c0, c1, F->d_chunksF.p, F->d_LptrT.p, F->d_Lcol.p, F->d_Lmap.p, F->d_usrc.p, F->d_forder.p, F->d_Dmap.p, F->nL, F->csr->d_val.p,
                F->d_fv.p, F->d_dinv.p, F->d_status.p);
            JB_CHECK_LAUNCH(ctx);
        }
        return JB_OK;
    }
    ilu_gather_kernel<BS><<<std::max(grid, 1), 256, 0, s>>>(F->nL, F->n, F->nU, F->d_Lmap.p, F->d_Dmap.p, F->d_Umap.p, F->csr->d_val.p, F->d_fv.p, 0);
    JB_CHECK_LAUNCH(ctx);
    for (int l = 0; l < F->nlevF; l++) {
        const int32_t t0 = F->h_levF_ptr[l], t1 = F->h_levF_ptr[l + 1];
        if (t1 <= t0) continue;
        ilu_factor_level_kernel<BS><<<(t1 - t0 + 127) / 128, 128, 0, s>>>(t0, t1, F->nL, F->d_forder.p, F->d_Lstart.p, F->d_Lend.p, F->d_Lcol.p,
                                                                          F->d_upd_ptr.p, F->d_upd_tgt.p, F->d_upd_src.p, F->d_fv.p, F->d_dinv.p,
                                                                          F->d_status.p);
        JB_CHECK_LAUNCH(ctx);
    }
    return JB_OK;
}

template <int BS>
static int ilu_apply_t(jb_ilu* F, const double* b, double* x, const double* sc) {
    jb_ctx* ctx = F->csr->ctx;
    ProfScope _ps(ctx, JB_PROF_ILU_APPLY);
    cudaStream_t s = ctx->stream;
    for (int l = 0; l < F->nlevF; l++) {
        const int32_t t0 = F->h_levF_ptr[l], t1 = F->h_levF_ptr[l + 1];
        if (t1 <= t0) continue;
        ilu_forward_level_kernel<BS><<<(t1 - t0 + 255) / 256, 256, 0, s>>>(t0, t1, F->d_forder.p, F->d_Lstart.p, F->d_Lend.p, F->d_Lcol.p,
                                                                           F->d_fv.p, b, x, sc);
        JB_CHECK_LAUNCH(ctx);
    }
    for (int l = 0; l < F->nlevB; l++) {
        const int32_t t0 = F->h_levB_ptr[l], t1 = F->h_levB_ptr[l + 1];
        if (t1 <= t0) continue;
        ilu_backward_level_kernel<BS><<<(t1 - t0 + 255) / 256, 256, 0, s>>>(t0, t1, F->nL + F->n, F->d_border.p, F->d_Ustart.p, F->d_Uend.p,
                                                                            F->d_Ucol.p, F->d_fv.p, F->d_dinv.p, x, sc);
        JB_CHECK_LAUNCH(ctx);
    }
    return JB_OK;
}

int jb_launch_ilu_factor(jb_ilu* F) {
    F->factored_gen = F->csr->val_gen;   // the factors now belong to this generation of Jacobian values
    if (F->diag_kind != 0) return jb_launch_diag_factor(F);
    switch (F->bs) {
        case 1: return ilu_factor_t<1>(F);
        case 2: return ilu_factor_t<2>(F);
        case 3: return ilu_factor_t<3>(F);
        case 4: return ilu_factor_t<4>(F);
    }
    return JB_ERR_UNSUPPORTED;
}
int jb_launch_ilu_apply_sc(jb_ilu* F, const double* d_b, double* d_x, const double* d_sc) {
    if (F->diag_kind != 0) return jb_launch_diag_apply(F, d_b, d_x, d_sc);   // Jacobi / SPAI(0) behind the same handle
    if (F->stream_ok) {
        switch (F->bs) {
            case 1: return ilu_apply_stream_t<1>(F, d_b, d_x, d_sc);
            case 2: return ilu_apply_stream_t<2>(F, d_b, d_x, d_sc);
            case 3: return ilu_apply_stream_t<3>(F, d_b, d_x, d_sc);
            case 4: return ilu_apply_stream_t<4>(F, d_b, d_x, d_sc);
        }
    }
    switch (F->bs) {   // fallback: one thread per row (a row longer than the shared-memory tile)
        case 1: return ilu_apply_t<1>(F, d_b, d_x, d_sc);
        case 2: return ilu_apply_t<2>(F, d_b, d_x, d_sc);
        case 3: return ilu_apply_t<3>(F, d_b, d_x, d_sc);
        case 4: return ilu_apply_t<4>(F, d_b, d_x, d_sc);
    }
    return JB_ERR_UNSUPPORTED;
}
int jb_launch_ilu_apply(jb_ilu* F, const double* d_b, double* d_x) { return jb_launch_ilu_apply_sc(F, d_b, d_x, nullptr); }

extern "C" {

int32_t jb_ilu0_create(jb_csr* A, const int64_t* partition, jb_ilu** out) {
    if (!A || !out) return JB_ERR_ARG;
    jb_ilu* F = new jb_ilu();
    F->csr = A; F->n = A->n; F->bs = A->bs;
    int rc = jb_ilu_symbolic(F, partition);
    if (rc == JB_OK) rc = jb_ilu_upload(F);
    if (rc != JB_OK) { delete F; JB_FAIL(A->ctx, rc, "jb_ilu0_create: symbolic phase failed (missing diagonal, bad partition or allocation)"); }
    A->n_ident_chunks = -1;   // identity-chunk flags of the Krylov driver are rebuilt for the new factor
    *out = F;
    return JB_OK;
}
int32_t jb_ilu0_destroy(jb_ilu* F) {
    if (F && F->apply_graph) cudaGraphExecDestroy(F->apply_graph);
    delete F;
    return JB_OK;
}
int32_t jb_ilu0_update(jb_ilu* F) {
    if (!F) return JB_ERR_ARG;
    jb_ctx* ctx = F->csr->ctx;
    int rc = jb_launch_ilu_factor(F);
    if (rc != JB_OK) return rc;
    int32_t st = 0;
    JB_CUDA(ctx, cudaMemcpyAsync(&st, F->d_status.p, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    JB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return st;
}
int32_t jb_ilu0_apply(jb_ilu* F, const double* d_b, double* d_x) {
    if (!F || !d_b || !d_x) return JB_ERR_ARG;
    int rc = jb_launch_ilu_apply(F, d_b, d_x);
    if (rc != JB_OK) return rc;
    JB_CUDA(F->csr->ctx, cudaStreamSynchronize(F->csr->ctx->stream));
    return JB_OK;
}
int32_t jb_ilu0_info(jb_ilu* F, int64_t* info) {
    if (!F || !info) return JB_ERR_ARG;
    info[0] = F->nlevF; info[1] = F->nlevB; info[2] = F->nL; info[3] = F->nU;
    return JB_OK;
}
// Dump in the reference's layout: L and U as CSR over all rows (ascending row), 1-based.
int32_t jb_ilu0_get(jb_ilu* F, int64_t* Lptr, int64_t* Lcol, double* L, int64_t* Uptr, int64_t* Ucol, double* U, double* Dinv) {
    if (!F) return JB_ERR_ARG;
    jb_ctx* ctx = F->csr->ctx;
    const int b2 = F->bs * F->bs;
    if (F->diag_kind != 0) {   // diagonal preconditioner: only the blocks D_i exist
        if (Lptr || Lcol || L || Uptr || Ucol || U) JB_FAIL(ctx, JB_ERR_ARG, "jb_ilu0_get: a diagonal preconditioner has no L / U factors");
        if (Dinv) JB_CUDA(ctx, cudaMemcpy(Dinv, F->d_dinv.p, (size_t)F->n * b2 * sizeof(double), cudaMemcpyDeviceToHost));
        return JB_OK;
    }
    std::vector<double> fv((size_t)(F->nL + F->n + F->nU) * b2);
    JB_CUDA(ctx, cudaMemcpyAsync(fv.data(), F->d_fv.p, fv.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (Dinv) JB_CUDA(ctx, cudaMemcpyAsync(Dinv, F->d_dinv.p, (size_t)F->n * b2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    JB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    i64 ol = 0, ou = 0;
    for (i64 r = 0; r < F->n; r++) {
        if (Lptr) Lptr[r] = ol + 1;
        if (Uptr) Uptr[r] = ou + 1;
        for (int32_t k = F->h_Lstart[r]; k < F->h_Lend[r]; k++, ol++) {
            if (Lcol) Lcol[ol] = F->h_Lcol[k] + 1;
            if (L) memcpy(L + ol * b2, &fv[(size_t)k * b2], sizeof(double) * b2);
        }
        for (int32_t k = F->h_Ustart[r]; k < F->h_Uend[r]; k++, ou++) {
            if (Ucol) Ucol[ou] = F->h_Ucol[k] + 1;
            if (U) memcpy(U + ou * b2, &fv[(size_t)(F->nL + F->n + k) * b2], sizeof(double) * b2);
        }
    }
    if (Lptr) Lptr[F->n] = ol + 1;
    if (Uptr) Uptr[F->n] = ou + 1;
    return JB_OK;
}

}  // extern "C"
