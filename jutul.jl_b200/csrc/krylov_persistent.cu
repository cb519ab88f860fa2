// Whole-solve persistent BiCGStab for sm_100a: ONE cooperative kernel runs every iteration of the ILU(0)-preconditioned
// BiCGStab of Jutul's linear_solve! (src/linsolve/krylov.jl:71-182; Krylov.jl 0.9 `bicgstab!` operation order, x0 = 0,
// c = b, right preconditioning) — SpMV, triangular sweeps, vector updates, inner products, scalar recurrence, termination
// test and, in a distributed run, the halo exchange and the all-reduce over peer memory (consistent!/dot of
// ext/JutulPartitionedArraysExt/krylov.jl:54-124, linalg.jl:37-55).
//
// Why: the multi-kernel driver (krylov.cu) needs ~12 launches and, distributed, 5 collective launches per iteration; at
// 1.26M cells per GPU the kernels are 50-80 us and launch gaps, prologues, tails and the one-warp collective kernels are a
// quarter of the iteration. Here the phases of an iteration are separated by grid barriers (one atomic + one spin per CTA),
// the TMA pipeline of the next phase is primed before the barrier (the matrix streams are static during a solve), the
// vector updates are folded into the sweeps' right-hand-side reads where the dependency allows, and the collectives are
// executed by the last CTA to reach the barrier (all-reduce) or by the producers themselves (halo: boundary values stored
// straight into the neighbour's vector over NVLink; the consumer waits for the epoch flag only in front of its first chunk
// with a ghost column, so interior rows overlap the transfer).
//
// Scope: 2x2 blocks, two-colour ILU(0) (colour R = rows without L entries, colour B = rows without U entries), right
// preconditioning, factors built from the current Jacobian values. Everything else takes the multi-kernel driver.
//
// Arithmetic per vector element is the multi-kernel driver's (same fma forms); only the summation order of the inner
// products differs. With x = N^-1 w, rows i < n_id ("identity rows", krylov.cu) satisfy (A x)_i = w_i, so v_i = p_i and
// t_i = s_i there and neither vector is stored on those rows.
//
// One iteration (R = rows outside [b0, b1), B = rows [b0, b1); "barrier" = grid barrier):
//   A1  B rows, L stream : [p_B = r_B + beta (p_B - omega v_B)]  y_B = D^-1 (p_B - L p_R)              barrier
//   A2  R rows, U stream : y_R = D^-1 (p_R - U y_B); <c,v> over identity rows                          barrier [halo y]
//   A3  rows >= n_id, A  : v = A y; <c,v>                                                              barrier+reduce -> alpha
//   V1  R rows           : s_R = r_R - alpha v_R                                                       barrier
//   A4  B rows, L stream : s_B = r_B - alpha v_B;  z_B = D^-1 (s_B - L s_R)                            barrier
//   A5  R rows, U stream : z_R = D^-1 (s_R - U z_B); <t,s>, <t,t> over identity rows                   barrier [halo z]
//   A6  rows >= n_id, A  : t = A z; <t,s>, <t,t>                                                       barrier+reduce -> omega
//   V2  all rows         : x += alpha y + omega z; r = s - omega t; <c,r>, <r,r>                       barrier+reduce -> beta, stop?
//   V3  R rows           : p_R = r_R + beta (p_R - omega v_R)                                          barrier
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <thread>

#include "jb_internal.cuh"
#include "jb_krylov_scalars.cuh"
#include "jb_reduce.cuh"
#include "jb_stream2.cuh"
#include "jb_persistent.cuh"
#ifndef PK_MINB
#define PK_MINB 6
#endif

// ---- device-side synchronisation -------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned pk_ld_acquire_gpu(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void pk_st_release_gpu(unsigned* p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void pk_st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long pk_ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long pk_globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// Bounded spins: a lost CTA or peer raises the abort word instead of hanging the GPU.
#define PK_SPIN_LIMIT_NS 40000000000ULL   // 40 s (a peer may itself be waiting ~30 s for a third rank)
// Wait until the phase bit of the barrier word differs from the one in `old` (volatile polls, as cooperative groups does:
// measured 2.0 us per barrier at 888 CTAs against 3.3 us for an atomic counter + release store + acquire polls,
// profiles/r02_barrier_microbench.txt).
__device__ __forceinline__ bool pk_spin_flip(const PKSync* S, unsigned old) {
    const unsigned long long t0 = pk_globaltimer();
    unsigned n = 0;
    while (((old ^ *reinterpret_cast<const volatile unsigned*>(&S->bar)) & 0x80000000u) == 0u) {
        if ((++n & 1023u) == 0) {
            if (pk_ld_acquire_gpu(&S->abort) != 0) return false;
            if (pk_globaltimer() - t0 > PK_SPIN_LIMIT_NS) { atomicExch(const_cast<unsigned*>(&S->abort), 1u); return false; }
        }
    }
    return true;
}
__device__ __forceinline__ bool pk_spin_peer(const unsigned long long* flag, unsigned long long epoch, PKSync* S) {
    const unsigned long long t0 = pk_globaltimer();
    unsigned n = 0;
    while (pk_ld_acquire_sys(flag) < epoch) {
        if ((++n & 255u) == 0) {
            if (pk_ld_acquire_gpu(&S->abort) != 0) return false;
            if (pk_globaltimer() - t0 > PK_SPIN_LIMIT_NS * 3 / 4) { atomicExch(&S->abort, 2u); return false; }
        }
        __nanosleep(20);
    }
    return true;
}

struct PKState {
    unsigned long long ar_epoch;     // all-reduce epoch (continues jb_dist::ar_epoch)
    unsigned long long halo_epoch;   // epoch of the in-kernel halo exchanges
    unsigned long long t_prev;       // phase timer (CTA 0, thread 0)
};

__device__ __forceinline__ void pk_phase_time(const PKArgs& a, PKState& st, int ph) {
    if (a.phase_ns && blockIdx.x == 0 && threadIdx.x == 0) {
        const unsigned long long t = pk_globaltimer();
        a.phase_ns[ph] += t - st.t_prev;
        st.t_prev = t;
    }
}

// Grid barrier: ONE atomic per CTA on one word. CTA 0 adds 2^31 - (grid - 1), every other CTA adds 1, so the word's phase bit
// flips exactly when the last CTA arrives and its low bits are back where they were — arrival and release are the same
// operation (the protocol of cooperative groups' grid.sync()). Returns false when the solve was aborted (timeout).
__device__ __forceinline__ bool pk_sync(const PKArgs& a, PKState& st, int* flag_s) {
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned nb = blockIdx.x == 0 ? 0x80000000u - (gridDim.x - 1u) : 1u;
        __threadfence();
        const unsigned old = atomicAdd(&a.sync->bar, nb);
        const bool ok = pk_spin_flip(a.sync, old);
        __threadfence();
        *flag_s = ok ? 1 : 0;
    }
    __syncthreads();
    return *flag_s != 0;
}

// All-reduce (sum) of vals[0..n) over the ranks through the peer slot tables (same protocol and slots as
// p2p_allreduce_kernel, dist.cu): executed by warp 0 of ONE CTA. Result in every lane's vals.
__device__ __forceinline__ bool pk_allreduce_warp(const PKArgs& a, PKState& st, double (&vals)[2]) {
    const PKDist& d = a.dist;
    const int lane = threadIdx.x & 31;
    const unsigned long long epoch = st.ar_epoch;      // already advanced by pk_sync_reduce (every thread of every CTA counts)
    const int parity = (int)(epoch & 1ULL);
    if (lane < d.world) {
        double* dst = d.peers[lane];
        dst[pk_ar_slot(parity, d.rank) + 0] = vals[0];
        dst[pk_ar_slot(parity, d.rank) + 1] = vals[1];
        __threadfence_system();
        pk_st_release_sys(reinterpret_cast<unsigned long long*>(dst) + pk_ar_flag(parity, d.rank), epoch);
    }
    bool ok = true;
    if (lane < d.world) ok = pk_spin_peer(reinterpret_cast<const unsigned long long*>(d.peers[d.rank]) + pk_ar_flag(parity, lane), epoch, a.sync);
    ok = __all_sync(0xffffffffu, ok);
    __threadfence_system();
    const double* mine = d.peers[d.rank];
    double s0 = __ldcv(mine + pk_ar_slot(parity, 0)), s1 = __ldcv(mine + pk_ar_slot(parity, 0) + 1);
    for (int q = 1; q < d.world; q++) {      // rank order: bitwise the same totals on every rank
        s0 += __ldcv(mine + pk_ar_slot(parity, q));
        s1 += __ldcv(mine + pk_ar_slot(parity, q) + 1);
    }
    vals[0] = s0; vals[1] = s1;
    return ok;
}

// Grid barrier + sum of two per-thread values over the grid (and the ranks) + scalar recurrence `fin(sc, hist)` run by ONE
// thread before the barrier opens: every CTA then reads the new scalars from the scalar block.
template <class Fin>
__device__ __forceinline__ bool pk_sync_reduce(const PKArgs& a, PKState& st, double d0, double d1, double* red_s, int* flag_s, Fin fin) {
    double v[2] = {d0, d1};
    if (a.dist.world > 1) st.ar_epoch++;               // one exchange per reduction, on every rank alike
    block_reduce<2, OpSum>(v, red_s);
    unsigned old = 0u;       // barrier word before this barrier: it cannot change until this CTA has arrived on `count`
    if (threadIdx.x == 0) {
        a.partials[2 * blockIdx.x] = v[0];
        a.partials[2 * blockIdx.x + 1] = v[1];
        old = *reinterpret_cast<const volatile unsigned*>(&a.sync->bar);
        __threadfence();
        flag_s[1] = (atomicAdd(&a.sync->count, 1u) == gridDim.x - 1) ? 1 : 0;
    }
    __syncthreads();
    if (flag_s[1]) {
        __threadfence();
        double acc[2] = {0.0, 0.0};
        for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
            acc[0] += __ldcg(a.partials + 2 * b);
            acc[1] += __ldcg(a.partials + 2 * b + 1);
        }
        block_reduce<2, OpSum>(acc, red_s);
        if (threadIdx.x < 32) {
            acc[0] = __shfl_sync(0xffffffffu, acc[0], 0);
            acc[1] = __shfl_sync(0xffffffffu, acc[1], 0);
            bool ok = true;
            if (a.dist.world > 1) ok = pk_allreduce_warp(a, st, acc);
            if (threadIdx.x == 0) {
                if (!ok) atomicExch(&a.sync->abort, 2u);
                a.sc[KS_SUM0] = acc[0]; a.sc[KS_SUM1] = acc[1];
                fin(a.sc, a.hist);
                a.sync->count = 0;
                __threadfence();
                atomicAdd(&a.sync->bar, 0x80000000u);      // flip the phase bit: the barrier opens
                flag_s[0] = ok ? 1 : 0;
            }
        }
    } else if (threadIdx.x == 0) {
        const bool ok = pk_spin_flip(a.sync, old);
        __threadfence();
        flag_s[0] = ok ? 1 : 0;
    }
    __syncthreads();
    return flag_s[0] != 0;
}

// ---- TMA-staged phases --------------------------------------------------------------------------------------------------------
struct PKStream {
    const S2Chunk* table; int c0, c1;
    const int32_t* ptr; const int32_t* col; const double* val; size_t voff; const int32_t* order;
};

// Prime the two stages with the first chunks of a phase (thread 0, stages idle). The matrix streams are static during a
// solve, so this is issued BEFORE the grid barrier that precedes the phase.
__device__ __forceinline__ void pk_prime(S2Smem& sm, const PKStream& m) {
    if (threadIdx.x == 0) {
        const int k0 = m.c0 + (int)blockIdx.x, k1 = k0 + (int)gridDim.x;
        if (k0 < m.c1) s2_issue(sm, 0, m.table, k0, m.ptr, m.col, m.val, m.voff, m.order);
        if (k1 < m.c1) s2_issue(sm, 1, m.table, k1, m.ptr, m.col, m.val, m.voff, m.order);
    }
}

// (D^-1 w)_e for the lane pair of a row (column-major 2x2 block: da = D^-1[e,0], db = D^-1[e,1])
__device__ __forceinline__ double pk_apply_dinv(double w, double da, double db, int e, unsigned pairmask) {
    const double other = __shfl_xor_sync(pairmask, w, 1);
    const double w0 = e ? other : w, w1 = e ? w : other;
    return fma(db, w1, da * w0);
}

// Halo exchange inside the kernel: every owned boundary value is stored straight into the ghost section of the
// neighbour's vector (peer memory); the last CTA to finish publishes the epoch to all neighbours.
__device__ __forceinline__ void pk_halo_push(const PKArgs& a, PKState& st, const double* vec, int which, int* flag_s) {
    const PKDist& d = a.dist;
    st.halo_epoch++;
    if (d.nneigh == 0) return;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
    for (long long i = tid; i < d.nsend; i += nth) {
        int k = 0;
        while (k + 1 < d.nneigh && i >= d.send_ptr[k + 1]) k++;
        double* dst = d.peers[d.neigh[k]] + (which == 0 ? d.remote_y[k] : d.remote_z[k]) + 2 * (i - d.send_ptr[k]);
        const size_t c = (size_t)__ldg(d.send_idx + i);
        const double2 val = *reinterpret_cast<const double2*>(vec + 2 * c);
        *reinterpret_cast<double2*>(dst) = val;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) flag_s[1] = (atomicAdd(&a.sync->push_count, 1u) == gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (flag_s[1]) {
        if (threadIdx.x == 0) a.sync->push_count = 0;
        if ((int)threadIdx.x < d.nneigh) {
            __threadfence_system();
            pk_st_release_sys(reinterpret_cast<unsigned long long*>(d.peers[d.neigh[threadIdx.x]]) + d.kflag_off + d.rank, st.halo_epoch);
        }
    }
}
__device__ __forceinline__ void pk_halo_wait(const PKArgs& a, const PKState& st) {
    const PKDist& d = a.dist;
    const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(d.peers[d.rank]) + d.kflag_off;
    for (int k = 0; k < d.nneigh; k++) pk_spin_peer(mine + d.neigh[k], st.halo_epoch, a.sync);
}

// ---- the phases: functions of their own with by-value arguments, so that the register allocator sees one phase at a time
//      and nothing in a hot loop is read through a pointer to the argument block --------------------------------------------------
enum { PH_INIT = 0, PH_A1, PH_A2, PH_A3, PH_V1, PH_A4, PH_A5, PH_A6, PH_V2, PH_V3, PH_HALO, PH_FINAL, PH_COUNT };

struct PKSums { double d0, d1; };
struct PKPar { uint32_t p0, p1; };

// The chunk loop of a stream phase, written out (no closures: every accumulator is a plain local and stays in a register).
#define PK_CHUNK_LOOP_BEGIN(C0, C1)                                                                         \
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, jl = lane >> 1, e = lane & 1;               \
    uint32_t par0 = pp.p0, par1 = pp.p1;                                                                    \
    int it = 0;                                                                                             \
    for (int k = (C0) + (int)blockIdx.x; k < (C1); k += (int)gridDim.x, it++) {                             \
        const int s = it & 1;                                                                               \
        jb_mbar_wait(&sm.bar[s], s ? par1 : par0);                                                          \
        if (s) par1 ^= 1u; else par0 ^= 1u;                                                                 \
        const S2Stage& S = sm.stage[s];                                                                     \
        const int t0 = sm.meta[s].t0, nr = sm.meta[s].nr, e0 = sm.meta[s].e0;                               \
        const int j = warp * 16 + jl;
#define PK_CHUNK_LOOP_END(TABLE, C1, PTR, COL, VAL, VOFF, ORDER)                                            \
        s2_release(sm, s, TABLE, k, C1, PTR, COL, VAL, VOFF, ORDER);                                        \
    }                                                                                                       \
    __syncthreads(); /* every warp is done with the stages: the next phase may be primed */                 \
    const PKPar qq = {par0, par1};

// B rows, L stream. mode 0: rhs = p_B; mode 1: p_B = r_B + beta (p_B - omega v_B) first; mode 2: s_B = r_B - alpha v_B first
// (then w = s). out_B = D^-1 (rhs - L w_R).
template <int MODE>
__device__ __forceinline__ PKPar pk_phase_L(S2Smem* smp, PKPar pp, const S2Chunk* __restrict__ table, int c0, int c1, const int32_t* __restrict__ ptr,
                                         const int32_t* __restrict__ col, const double* __restrict__ fv, const int32_t* __restrict__ order,
                                         const double* __restrict__ dinv, double* w, const double* r, const double* v, double* out, double c_a,
                                         double c_b) {
    S2Smem& sm = *smp;
    const unsigned pairmask = 3u << (threadIdx.x & 30);
    PK_CHUNK_LOOP_BEGIN(c0, c1)
        if (j < nr) {
            const int lrp = jb_span_lead<int32_t>((size_t)t0), lcol = jb_span_lead<int32_t>((size_t)e0);
            const size_t i = (size_t)S.ord[lrp + j];
            const double da = __ldg(dinv + i * 4 + e), db = __ldg(dinv + i * 4 + 2 + e);
            double wi;
            if (MODE == 0) wi = w[i * 2 + e];
            else if (MODE == 1) {                  // c_a = beta, c_b = omega
                const double pa = fma(-c_b, v[i * 2 + e], w[i * 2 + e]);
                wi = fma(c_a, pa, r[i * 2 + e]);
                w[i * 2 + e] = wi;
            } else {                               // c_a = alpha
                wi = fma(-c_a, v[i * 2 + e], r[i * 2 + e]);
                w[i * 2 + e] = wi;
            }
            const double sum = s2_row_sum<6>(S, lcol, S.rp[lrp + j] - e0, S.rp[lrp + j + 1] - e0, e, w);
            out[i * 2 + e] = pk_apply_dinv(wi - sum, da, db, e, pairmask);
        }
    PK_CHUNK_LOOP_END(table, c1, ptr, col, fv, 0, order)
    return qq;
}

// R rows, U stream: out_R = D^-1 (w_R - U out_B) with (w, out) = (p, y) [WHICH 0] or (s, z) [WHICH 1]; inner products over
// the identity rows: <c,v> with v_i = p_i, or <t,s> = <t,t> = <s,s> with t_i = s_i.
template <int WHICH>
__device__ __forceinline__ PKPar pk_phase_U(S2Smem* smp, PKPar pp, const S2Chunk* __restrict__ table, int c0, int c1, const int32_t* __restrict__ ptr,
                                         const int32_t* __restrict__ col, const double* __restrict__ fv, size_t voff, const int32_t* __restrict__ order,
                                         const double* __restrict__ dinv, const double* w, double* out, const double* __restrict__ c, int n_id,
                                         PKSums* sums) {
    S2Smem& sm = *smp;
    const unsigned pairmask = 3u << (threadIdx.x & 30);
    double d0 = 0.0, d1 = 0.0;
    PK_CHUNK_LOOP_BEGIN(c0, c1)
        if (j < nr) {
            const int lrp = jb_span_lead<int32_t>((size_t)t0), lcol = jb_span_lead<int32_t>((size_t)e0);
            const size_t i = (size_t)S.ord[lrp + j];
            const double da = __ldg(dinv + i * 4 + e), db = __ldg(dinv + i * 4 + 2 + e);
            const double wi = w[i * 2 + e];
            if ((int)i < n_id) {
                if (WHICH == 0) d0 = fma(__ldg(c + i * 2 + e), wi, d0);
                else { d0 = fma(wi, wi, d0); d1 = fma(wi, wi, d1); }
            }
            const double sum = s2_row_sum<6>(S, lcol, S.rp[lrp + j] - e0, S.rp[lrp + j + 1] - e0, e, out);
            out[i * 2 + e] = pk_apply_dinv(wi - sum, da, db, e, pairmask);
        }
    PK_CHUNK_LOOP_END(table, c1, ptr, col, fv, voff, order)
    sums->d0 = d0; sums->d1 = d1;      // also for a CTA without a chunk of this phase: the sums of the previous reduction must not leak
    return qq;
}
// rows with neither L nor U entries: out_i = D_i^-1 w_i
template <int WHICH>
__device__ __forceinline__ void pk_phase_iso(const int32_t* __restrict__ iso, int n_iso, const double* __restrict__ dinv, const double* w, double* out,
                                          const double* __restrict__ c, int n_id, PKSums* sums) {
    double d0 = sums->d0, d1 = sums->d1;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
    for (long long q = tid; q < n_iso; q += nth) {
        const size_t i = (size_t)__ldg(iso + q);
        const double2 wi = reinterpret_cast<const double2*>(w)[i];
        const double4 di = *reinterpret_cast<const double4*>(dinv + i * 4);
        reinterpret_cast<double2*>(out)[i] = make_double2(fma(di.z, wi.y, di.x * wi.x), fma(di.w, wi.y, di.y * wi.x));
        if ((int)i < n_id) {
            if (WHICH == 0) { d0 = fma(__ldg(c + i * 2), wi.x, d0); d0 = fma(__ldg(c + i * 2 + 1), wi.y, d0); }
            else { d0 = fma(wi.x, wi.x, d0); d0 = fma(wi.y, wi.y, d0); d1 = fma(wi.x, wi.x, d1); d1 = fma(wi.y, wi.y, d1); }
        }
    }
    sums->d0 = d0; sums->d1 = d1;
}

// Rows that are not identity rows: o = A g with (g, o) = (y, v) [WHICH 0] or (z, t) [WHICH 1]; <u,o> (u = c) or <o,u>, <o,o>
// (u = s) added to *sums. Chunks with a ghost column wait for the neighbours' halo epoch first (halo_args != nullptr).
template <int WHICH>
__device__ __forceinline__ PKPar pk_phase_A(S2Smem* smp, PKPar pp, const S2Chunk* __restrict__ table, int nA, const int32_t* __restrict__ ptr,
                                         const int32_t* __restrict__ col, const double* __restrict__ val, const double* g, double* o, const double* u,
                                         PKSums* sums, const PKArgs* halo_args, unsigned long long halo_epoch) {
    S2Smem& sm = *smp;
    double d0 = sums->d0, d1 = sums->d1;
    bool halo_ok = halo_args == nullptr;
    PK_CHUNK_LOOP_BEGIN(0, nA)
        if ((sm.meta[s].flags & 2) && !halo_ok) {
            PKState hst; hst.halo_epoch = halo_epoch;
            pk_halo_wait(*halo_args, hst);
            halo_ok = true;
        }
        if (j < nr) {
            const int lrp = jb_span_lead<int32_t>((size_t)t0), lcol = jb_span_lead<int32_t>((size_t)e0);
            const size_t i = (size_t)(t0 + j);
            const double ui = u[i * 2 + e];
            const double sum = s2_row_sum<8>(S, lcol, S.rp[lrp + j] - e0, S.rp[lrp + j + 1] - e0, e, g);
            o[i * 2 + e] = sum;
            if (WHICH == 0) d0 = fma(ui, sum, d0);
            else { d0 = fma(sum, ui, d0); d1 = fma(sum, sum, d1); }
        }
    PK_CHUNK_LOOP_END(table, nA, ptr, col, val, 0, nullptr)
    sums->d0 = d0; sums->d1 = d1;
    return qq;
}

// rows outside [b0, b1): V1 s = r - alpha v, V3 p = r + beta (p - omega v); v_i = p_i on the identity rows. Two rows per
// thread and trip, loads first.
template <int WHICH>
__device__ __forceinline__ void pk_phase_VR(int n_own, int n_id, int b0, int b1, const double* r, double* p, const double* v, double* s, double alpha,
                                         double beta, double omega) {
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
    const int nb = b1 - b0, n_R = n_own - nb;
    for (long long q0 = tid; q0 < n_R; q0 += 2 * nth) {
        const long long q1 = q0 + nth;
        const bool two = q1 < n_R;
        const long long ia = q0 < b0 ? q0 : q0 + nb;
        const long long ib = two ? (q1 < b0 ? q1 : q1 + nb) : ia;
        const double2 ra = reinterpret_cast<const double2*>(r)[ia], rb = reinterpret_cast<const double2*>(r)[ib];
        const double2 pa = reinterpret_cast<const double2*>(p)[ia], pb = reinterpret_cast<const double2*>(p)[ib];
        const double2 va = ia < n_id ? pa : reinterpret_cast<const double2*>(v)[ia];
        const double2 vb = ib < n_id ? pb : reinterpret_cast<const double2*>(v)[ib];
        if (WHICH == 0) {
            reinterpret_cast<double2*>(s)[ia] = make_double2(fma(-alpha, va.x, ra.x), fma(-alpha, va.y, ra.y));
            if (two) reinterpret_cast<double2*>(s)[ib] = make_double2(fma(-alpha, vb.x, rb.x), fma(-alpha, vb.y, rb.y));
        } else {
            const double pax = fma(-omega, va.x, pa.x), pay = fma(-omega, va.y, pa.y);
            reinterpret_cast<double2*>(p)[ia] = make_double2(fma(beta, pax, ra.x), fma(beta, pay, ra.y));
            if (two) {
                const double pbx = fma(-omega, vb.x, pb.x), pby = fma(-omega, vb.y, pb.y);
                reinterpret_cast<double2*>(p)[ib] = make_double2(fma(beta, pbx, rb.x), fma(beta, pby, rb.y));
            }
        }
    }
}

// all owned rows: x += alpha y + omega z; r = s - omega t (t_i = s_i on the identity rows); <c,r>, <r,r>
// Two rows per thread and trip: the 14 loads of both rows are issued before the first fma (the phase has no TMA pipeline;
// its memory-level parallelism is what the threads themselves keep in flight).
__device__ __forceinline__ void pk_phase_V2(int n_own, int n_id, const double* y, const double* z, const double* s, const double* t,
                                         const double* __restrict__ c, double* x, double* r, double alpha, double omega, PKSums* sums) {
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
    double d0 = 0.0, d1 = 0.0;
    for (long long i0 = tid; i0 < n_own; i0 += 2 * nth) {
        const long long i1 = i0 + nth;
        const bool two = i1 < n_own;
        const long long j1 = two ? i1 : i0;
        const double2 ya = reinterpret_cast<const double2*>(y)[i0], yb = reinterpret_cast<const double2*>(y)[j1];
        const double2 za = reinterpret_cast<const double2*>(z)[i0], zb = reinterpret_cast<const double2*>(z)[j1];
        const double2 sa = reinterpret_cast<const double2*>(s)[i0], sb = reinterpret_cast<const double2*>(s)[j1];
        const double2 ta = i0 < n_id ? sa : reinterpret_cast<const double2*>(t)[i0];
        const double2 tb = j1 < n_id ? sb : reinterpret_cast<const double2*>(t)[j1];
        const double2 ca = __ldg(reinterpret_cast<const double2*>(c) + i0), cb = __ldg(reinterpret_cast<const double2*>(c) + j1);
        double2 xa = reinterpret_cast<const double2*>(x)[i0], xb = reinterpret_cast<const double2*>(x)[j1];
        xa.x = fma(omega, za.x, fma(alpha, ya.x, xa.x));
        xa.y = fma(omega, za.y, fma(alpha, ya.y, xa.y));
        reinterpret_cast<double2*>(x)[i0] = xa;
        const double rax = fma(-omega, ta.x, sa.x), ray = fma(-omega, ta.y, sa.y);
        reinterpret_cast<double2*>(r)[i0] = make_double2(rax, ray);
        d0 = fma(ca.x, rax, d0); d0 = fma(ca.y, ray, d0);
        d1 = fma(rax, rax, d1); d1 = fma(ray, ray, d1);
        if (two) {
            xb.x = fma(omega, zb.x, fma(alpha, yb.x, xb.x));
            xb.y = fma(omega, zb.y, fma(alpha, yb.y, xb.y));
            reinterpret_cast<double2*>(x)[i1] = xb;
            const double rbx = fma(-omega, tb.x, sb.x), rby = fma(-omega, tb.y, sb.y);
            reinterpret_cast<double2*>(r)[i1] = make_double2(rbx, rby);
            d0 = fma(cb.x, rbx, d0); d0 = fma(cb.y, rby, d0);
            d1 = fma(rbx, rbx, d1); d1 = fma(rby, rby, d1);
        }
    }
    sums->d0 = d0; sums->d1 = d1;
}

// r = b, p = b, x = 0; <r,r> (= <c,r>: c = b)
__device__ __forceinline__ void pk_phase_init(int n_own, const double* __restrict__ b, double* r, double* p, double* x, PKSums* sums) {
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
    double d0 = 0.0;
    for (long long q = tid; q < n_own; q += nth) {
        const double2 bi = reinterpret_cast<const double2*>(b)[q];
        reinterpret_cast<double2*>(r)[q] = bi;
        reinterpret_cast<double2*>(p)[q] = bi;
        reinterpret_cast<double2*>(x)[q] = make_double2(0.0, 0.0);
        d0 = fma(bi.x, bi.x, d0); d0 = fma(bi.y, bi.y, d0);
    }
    sums->d0 = d0; sums->d1 = d0;
}
// dx = -x (update_dx_from_vector!, src/linsolve/default.jl:444-446)
__device__ __forceinline__ void pk_phase_final(int n_own, const double* x, double* dx) {
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
    for (long long i = tid; i < n_own; i += nth) {
        const double2 xi = reinterpret_cast<const double2*>(x)[i];
        reinterpret_cast<double2*>(dx)[i] = make_double2(-xi.x, -xi.y);
    }
}
__device__ __forceinline__ void pk_prime_any(S2Smem* sm, const S2Chunk* table, int c0, int c1, const int32_t* ptr, const int32_t* col, const double* val,
                                          size_t voff, const int32_t* order) {
    const PKStream m = {table, c0, c1, ptr, col, val, voff, order};
    pk_prime(*sm, m);
}
__device__ __forceinline__ bool pk_barrier(const PKArgs& a, PKState& st, int* flag_s, int ph) {
    const bool ok = pk_sync(a, st, flag_s);
    pk_phase_time(a, st, ph);
    return ok;
}
// which: 0 init, 1 alpha, 2 omega, 3 update2
__device__ __forceinline__ bool pk_barrier_reduce(const PKArgs& a, PKState& st, const PKSums* sums, double* red_s, int* flag_s, int which, int ph) {
    const int hist_cap = a.hist_cap;
    const bool ok = pk_sync_reduce(a, st, sums->d0, sums->d1, red_s, flag_s, [which, hist_cap](double* sc, double* hist) {
        if (which == 0) ks_fin_init(sc, hist);
        else if (which == 1) ks_fin_alpha(sc);
        else if (which == 2) ks_fin_omega(sc);
        else ks_fin_update2(sc, hist, hist_cap);
    });
    pk_phase_time(a, st, ph);
    return ok;
}
__device__ __forceinline__ void pk_halo(const PKArgs& a, PKState& st, int which, int* flag_s) {
    pk_halo_push(a, st, which ? a.z : a.y, which, flag_s);
    pk_phase_time(a, st, PH_HALO);
}

// ---- the kernel: ONE BiCGStab iteration per launch --------------------------------------------------------------------------------
// (The iteration loop is on the host, one launch ahead of the device: with the loop inside the kernel, ptxas 12.9 spills the
// gather registers of every phase — 1.8 KB of spill code in the hot loops, measured 2x slower — while the loop-free body
// compiles to 80 registers without a single local-memory access.)
// flags: bit 0 = first launch of a solve (initialisation phase, then iteration 1).
// The termination test of launch `it` reaches the host through a pinned word the kernel writes itself ((it << 1) | done): no
// copy sits on the stream between two launches (a D2H copy + event there cost ~15 us per iteration: profiles/README.md).
__device__ __forceinline__ void pk_report(const PKArgs& a, int flags) {
    if (a.host_prog && blockIdx.x == 0 && threadIdx.x == 0) {
        const unsigned long long done = __ldcg(a.sc + KS_DONE) != 0.0 ? 1ULL : 0ULL;
        *reinterpret_cast<volatile unsigned long long*>(a.host_prog) = ((unsigned long long)(flags >> 1) << 1) | done;
        __threadfence_system();
    }
}
__global__ void __launch_bounds__(JB_S2_THREADS, PK_MINB) bicgstab_rb_iteration_kernel(const __grid_constant__ PKArgs a, const int flags) {
    extern __shared__ __align__(128) unsigned char s2_raw[];
    S2Smem* sm = reinterpret_cast<S2Smem*>(s2_raw);
    __shared__ double red_s[2 * 32];
    __shared__ int flag_s[2];
    const bool first = (flags & 1) != 0;
    if (!first && __ldcg(a.sc + KS_DONE) != 0.0) return;      // a launch enqueued ahead of the termination test: nothing to do
    PKPar par = {0u, 0u};
    PKState st;
    PKSums sums = {0.0, 0.0};
    st.ar_epoch = __ldcg(a.state);
    st.halo_epoch = __ldcg(a.state + 1);
    st.t_prev = a.phase_ns ? pk_globaltimer() : 0ULL;
    if (threadIdx.x == 0) {
        jb_mbar_init(&sm->bar[0], 1);
        jb_mbar_init(&sm->bar[1], 1);
        sm->cnt[0] = 0; sm->cnt[1] = 0;
        jb_mbar_fence_init();
    }
    __syncthreads();
    const bool dist = a.dist.world > 1;
#define PK_PRIME_L() pk_prime_any(sm, a.tabL, a.cL0, a.cL1, a.LptrT, a.Lcol, a.fv, 0, a.forder)
#define PK_PRIME_U() pk_prime_any(sm, a.tabU, a.cU0, a.cU1, a.UptrT, a.Ucol, a.fv, a.Uoff, a.border)
#define PK_PRIME_A() pk_prime_any(sm, a.tabA, 0, a.nA, a.rowptr, a.colidx, a.val, 0, nullptr)
#define PK_EXIT() { if (blockIdx.x == 0 && threadIdx.x == 0) { a.state[0] = st.ar_epoch; a.state[1] = st.halo_epoch; } return; }

    if (first) {
        pk_phase_init(a.n_own, a.b, a.r, a.p, a.x, &sums);
        if (!pk_barrier_reduce(a, st, &sums, red_s, flag_s, 0, PH_INIT)) return;
        if (__ldcg(a.sc + KS_DONE) != 0.0) {            // x = 0 solves the system (or breakdown): dx = -x = 0
            pk_report(a, flags);
            pk_phase_final(a.n_own, a.x, a.dx);
            PK_EXIT()
        }
    }
    PK_PRIME_L();
    // A1
    if (first) par = pk_phase_L<0>(sm, par, a.tabL, a.cL0, a.cL1, a.LptrT, a.Lcol, a.fv, a.forder, a.dinv, a.p, a.r, a.v, a.y, 0.0, 0.0);
    else par = pk_phase_L<1>(sm, par, a.tabL, a.cL0, a.cL1, a.LptrT, a.Lcol, a.fv, a.forder, a.dinv, a.p, a.r, a.v, a.y, __ldcg(a.sc + KS_BETA),
                             __ldcg(a.sc + KS_OMEGA));
    PK_PRIME_U();
    if (!pk_barrier(a, st, flag_s, PH_A1)) return;
    // A2
    par = pk_phase_U<0>(sm, par, a.tabU, a.cU0, a.cU1, a.UptrT, a.Ucol, a.fv, a.Uoff, a.border, a.dinv, a.p, a.y, a.b, a.n_id, &sums);
    if (a.n_iso > 0) pk_phase_iso<0>(a.iso, a.n_iso, a.dinv, a.p, a.y, a.b, a.n_id, &sums);
    PK_PRIME_A();
    if (!pk_barrier(a, st, flag_s, PH_A2)) return;
    if (dist) pk_halo(a, st, 0, flag_s);
    // A3
    par = pk_phase_A<0>(sm, par, a.tabA, a.nA, a.rowptr, a.colidx, a.val, a.y, a.v, a.b, &sums, dist ? &a : nullptr, st.halo_epoch);
    if (!pk_barrier_reduce(a, st, &sums, red_s, flag_s, 1, PH_A3)) return;
    // V1
    const double alpha = __ldcg(a.sc + KS_ALPHA);
    pk_phase_VR<0>(a.n_own, a.n_id, a.b0, a.b1, a.r, a.p, a.v, a.s, alpha, 0.0, 0.0);
    PK_PRIME_L();
    if (!pk_barrier(a, st, flag_s, PH_V1)) return;
    // A4
    par = pk_phase_L<2>(sm, par, a.tabL, a.cL0, a.cL1, a.LptrT, a.Lcol, a.fv, a.forder, a.dinv, a.s, a.r, a.v, a.z, alpha, 0.0);
    PK_PRIME_U();
    if (!pk_barrier(a, st, flag_s, PH_A4)) return;
    // A5
    par = pk_phase_U<1>(sm, par, a.tabU, a.cU0, a.cU1, a.UptrT, a.Ucol, a.fv, a.Uoff, a.border, a.dinv, a.s, a.z, a.b, a.n_id, &sums);
    if (a.n_iso > 0) pk_phase_iso<1>(a.iso, a.n_iso, a.dinv, a.s, a.z, a.b, a.n_id, &sums);
    PK_PRIME_A();
    if (!pk_barrier(a, st, flag_s, PH_A5)) return;
    if (dist) pk_halo(a, st, 1, flag_s);
    // A6
    par = pk_phase_A<1>(sm, par, a.tabA, a.nA, a.rowptr, a.colidx, a.val, a.z, a.t, a.s, &sums, dist ? &a : nullptr, st.halo_epoch);
    if (!pk_barrier_reduce(a, st, &sums, red_s, flag_s, 2, PH_A6)) return;
    // V2
    const double omega = __ldcg(a.sc + KS_OMEGA);
    pk_phase_V2(a.n_own, a.n_id, a.y, a.z, a.s, a.t, a.b, a.x, a.r, alpha, omega, &sums);
    if (!pk_barrier_reduce(a, st, &sums, red_s, flag_s, 3, PH_V2)) return;
    pk_report(a, flags);
    if (__ldcg(a.sc + KS_DONE) != 0.0) {
        pk_phase_final(a.n_own, a.x, a.dx);
        pk_phase_time(a, st, PH_FINAL);
        PK_EXIT()
    }
    // V3 (the B rows' part is folded into A1 of the next launch)
    pk_phase_VR<1>(a.n_own, a.n_id, a.b0, a.b1, a.r, a.p, a.v, a.s, 0.0, __ldcg(a.sc + KS_BETA), omega);
    pk_phase_time(a, st, PH_V3);
    PK_EXIT()
#undef PK_PRIME_L
#undef PK_PRIME_U
#undef PK_PRIME_A
#undef PK_EXIT
}

// ---- host ---------------------------------------------------------------------------------------------------------------------------
static bool pk_enabled() {
    const char* e = getenv("JB_PERSISTENT");
    return !(e && e[0] == '0');
}

// Build (once per Jacobian / factor / distribution) the tables of the persistent solve. Returns false when the structure
// does not qualify; the caller then takes the multi-kernel driver.
static bool pk_prepare(jb_krylov* K, i64 n_own) {
    jb_csr* A = K->csr;
    jb_ilu* F = K->ilu;
    PKHost& H = K->pk;
    if (H.built && H.for_ilu == (const void*)F && H.n_own == n_own) return H.ok;
    H.built = true; H.ok = false; H.for_ilu = (const void*)F; H.n_own = n_own;
    if (!F || F->diag_kind != 0 || A->bs != 2 || !A->s2_ok || !F->two_colour || !F->s2_ok || F->nlevF != 2 || F->nlevB != 2) return false;
    if (n_own > A->n || n_own < 1 || A->n >= (i64)1 << 30) return false;
    // B rows = second forward level: must be the contiguous ascending row range [b0, b1) inside the owned rows
    const int32_t tB0 = F->h_levF_ptr[1], tB1 = F->h_levF_ptr[2];
    if (tB1 <= tB0) return false;
    const int32_t b0 = F->h_forder[tB0], b1 = b0 + (tB1 - tB0);
    for (int32_t t = tB0; t < tB1; t++)
        if (F->h_forder[t] != b0 + (t - tB0)) return false;
    if (b1 > n_own) return false;
    // R rows = second backward level: owned rows only (rows of the level beyond n_own would be swept without being owned)
    for (int32_t t = F->h_levB_ptr[1]; t < F->h_levB_ptr[2]; t++)
        if (F->h_border[t] >= n_own) return false;
    // isolated owned rows
    std::vector<int32_t> iso;
    for (int32_t r : F->h_iso)
        if (r < n_own) iso.push_back(r);
    // identity rows: the longest prefix of rows without L entries whose couplings are all kept in U (krylov.cu)
    int32_t n_id = 0;
    while (n_id < n_own && n_id < b0) {
        const int32_t rowlen = A->h_rowptr[n_id + 1] - A->h_rowptr[n_id];
        if (F->h_Lend[n_id] != F->h_Lstart[n_id] || F->h_Uend[n_id] - F->h_Ustart[n_id] != rowlen - 1) break;
        n_id++;
    }
    const char* eid = getenv("JB_RB_IDENTITY");
    if (eid && eid[0] == '0') n_id = 0;
    // SpMV table over the rows [n_id, n_own); bit 1 of the flags = the chunk reads a ghost column
    std::vector<S2Chunk> tab;
    if (n_id < n_own && !jb_s2_cut(A->h_rowptr, n_id, (int32_t)n_own, tab, A->ctx->sm_count * JB_S2_CTAS_PER_SM)) return false;
    for (S2Chunk& c : tab) {
        bool ghost = false;
        for (int32_t k = c.e0; !ghost && k < c.e0 + c.cnt; k++) ghost = A->h_colidx[k] >= n_own;
        if (ghost) c.flags |= 2;
    }
    cudaStream_t s = A->ctx->stream;
    if (H.d_tabA.upload(tab, s) != cudaSuccess) return false;
    if (!iso.empty() && H.d_iso.upload(iso, s) != cudaSuccess) return false;
    H.nA = (int)tab.size(); H.n_iso = (int)iso.size(); H.n_id = n_id; H.b0 = b0; H.b1 = b1;
    H.n_id_blocks = A->h_rowptr[n_id];
    if (!H.d_sync.p) {
        std::vector<unsigned> z(8, 0u);
        if (H.d_sync.upload(z, s) != cudaSuccess) return false;
    }
    if (!H.d_partials.p && H.d_partials.alloc(2 * 4096) != cudaSuccess) return false;
    if (!H.d_state.p) {
        std::vector<unsigned long long> z(PH_COUNT + 4, 0ULL);
        if (H.d_state.upload(z, s) != cudaSuccess) return false;
    }
    // co-resident grid of the cooperative launch
    int per_sm = 0;
    cudaFuncSetAttribute(bicgstab_rb_iteration_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(S2Smem));
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, bicgstab_rb_iteration_kernel, JB_S2_THREADS, sizeof(S2Smem)) != cudaSuccess || per_sm < 1)
        return false;
    H.grid = std::min(A->ctx->sm_count * per_sm, 4096);
    const char* eg = getenv("JB_PK_CTAS_PER_SM");
    if (eg && atoi(eg) >= 1) H.grid = std::min(H.grid, A->ctx->sm_count * atoi(eg));
    int coop = 0;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, A->ctx->device);
    if (!coop) return false;
    H.ok = true;
    return true;
}

bool jb_krylov_persistent_ok(jb_krylov* K, int side, i64 n_own) {
    if (!pk_enabled() || K->kind != 0 || side != 0 || !K->ilu || K->schur) return false;
    jb_csr* A = K->csr;
    if (K->ilu->factored_gen != A->val_gen) return false;   // the identity rows need factors of the current values
    if (K->dist && !jb_dist_is_p2p(K->dist)) return false;
    return pk_prepare(K, n_own);
}
// per-phase device times (ms) of the fused iteration kernel accumulated while profiling is enabled:
// out[0..11] = init, A1, A2, A3, V1, A4, A5, A6, V2, V3, halo pushes, final; out[12] = iterations, out[13] = solves. Clears.
extern "C" int32_t jb_krylov_phase_times(jb_krylov* K, double* out) {
    if (!K || !out) return JB_ERR_ARG;
    for (int q = 0; q < 12; q++) { out[q] = K->pk.phase_ms[q]; K->pk.phase_ms[q] = 0.0; }
    out[12] = (double)K->pk.phase_iters; out[13] = (double)K->pk.phase_solves;
    K->pk.phase_iters = 0; K->pk.phase_solves = 0;
    return JB_OK;
}
bool jb_krylov_persistent_aligned(const double* d_b, const double* d_dx) {
    return (((uintptr_t)d_b | (uintptr_t)d_dx) & 15) == 0;   // vectors are streamed as double2
}

int jb_dist_pk_fill(jb_dist* D, PKDist* out, unsigned long long* ar_epoch0, unsigned long long* halo_epoch0);
void jb_dist_pk_done(jb_dist* D, unsigned long long ar_epoch, unsigned long long halo_epoch);

// Same contract as jb_krylov_solve_impl (krylov.cu).
int jb_krylov_solve_persistent(jb_krylov* K, const double* d_b, double* d_dx, double rtol, double atol, int itmax, int min_it, int* iters,
                               double* hist, int hist_cap, int* status_out) {
    jb_csr* A = K->csr;
    jb_ctx* ctx = A->ctx;
    jb_ilu* F = K->ilu;
    cudaStream_t st = ctx->stream;
    PKHost& H = K->pk;
    if (itmax > K->hist_cap - 2) itmax = K->hist_cap - 2;
    double h_sc[KS_SIZE];
    memset(h_sc, 0, sizeof(h_sc));
    const bool manual = min_it > 1;
    h_sc[KS_MANUAL] = manual ? 1.0 : 0.0;
    h_sc[KS_ABS_TOL] = atol; h_sc[KS_REL_TOL] = rtol; h_sc[KS_MIN_IT] = (double)min_it;
    h_sc[KS_ATOL] = manual ? 1e-20 : atol;
    h_sc[KS_RTOL] = manual ? 1e-20 : rtol;
    h_sc[KS_ITMAX] = (double)itmax;
    h_sc[KS_DIST] = K->dist ? 1.0 : 0.0;
    memcpy(K->h_flags + 2 * KS_SIZE, h_sc, sizeof(h_sc));
    JB_CUDA(ctx, cudaMemcpyAsync(K->d_sc.p, K->h_flags + 2 * KS_SIZE, sizeof(h_sc), cudaMemcpyHostToDevice, st));

    PKArgs a;
    memset(&a, 0, sizeof(a));
    a.n_own = (int)H.n_own; a.n_id = H.n_id; a.b0 = H.b0; a.b1 = H.b1;
    a.tabA = H.d_tabA.p; a.nA = H.nA; a.rowptr = A->d_rowptr.p; a.colidx = A->d_colidx.p; a.val = A->d_val.p;
    a.tabL = F->d_s2F.p; a.cL0 = F->h_levF_s2[1]; a.cL1 = F->h_levF_s2[2]; a.LptrT = F->d_LptrT.p; a.Lcol = F->d_Lcol.p; a.forder = F->d_forder.p;
    a.tabU = F->d_s2B.p; a.cU0 = F->h_levB_s2[1]; a.cU1 = F->h_levB_s2[2]; a.UptrT = F->d_UptrT.p; a.Ucol = F->d_Ucol.p; a.border = F->d_border.p;
    a.Uoff = (size_t)(F->nL + F->n);
    a.fv = F->d_fv.p; a.dinv = F->d_dinv.p;
    a.iso = H.d_iso.p; a.n_iso = H.n_iso;
    a.b = d_b; a.r = K->r.p; a.p = K->p.p; a.v = K->v.p; a.s = K->s.p; a.t = K->t.p; a.y = K->yv; a.z = K->zv; a.x = K->x.p; a.dx = d_dx;
    a.sc = K->d_sc.p; a.hist = K->d_hist.p; a.hist_cap = K->hist_cap;
    a.sync = reinterpret_cast<PKSync*>(H.d_sync.p); a.partials = H.d_partials.p; a.state = H.d_state.p;
    a.phase_ns = ctx->prof_on ? H.d_state.p + 4 : nullptr;
    a.dist.world = 1;
    if (K->dist) {
        int rc = jb_dist_pk_fill(K->dist, &a.dist, &a.ar_epoch0, &a.halo_epoch0);
        if (rc != JB_OK) return rc;
    }
    {   // state words: epochs as they stand before the solve (each launch reads them and writes them back), phase timers zeroed
        unsigned long long* hs0 = reinterpret_cast<unsigned long long*>(K->h_flags + 3 * KS_SIZE);   // pinned scratch, read back below
        for (int q = 0; q < PH_COUNT + 4; q++) hs0[q] = 0ULL;
        hs0[0] = a.ar_epoch0; hs0[1] = a.halo_epoch0;
        JB_CUDA(ctx, cudaMemcpyAsync(H.d_state.p, hs0, (PH_COUNT + 4) * sizeof(unsigned long long), cudaMemcpyHostToDevice, st));
    }
    // The host stays one launch ahead of the device: before enqueuing iteration `it` it waits for the flags written after
    // iteration it-2 (same pinned slot this iteration reuses); a launch that finds the solve finished returns at once.
    int kflags = 1;
    void* kargs[] = {(void*)&a, (void*)&kflags};
    const char* ehp = getenv("JB_PK_HOST_FLAGS");
    const bool mapped = K->h_prog != nullptr && !(ehp && ehp[0] == '0');
    a.host_prog = mapped ? K->d_prog : nullptr;
    volatile unsigned long long* prog = K->h_prog;
    if (mapped) *prog = 0ULL;      // the stream is idle here: the previous solve ended with a synchronisation
    if (ctx->prof_on) jb_prof_begin(ctx, JB_PROF_FUSED_SOLVE);
    for (int it = 1; it <= std::max(itmax, 1); it++) {
        if (it >= 3) {     // stay two launches ahead of the device: wait for the termination test of launch it - 2
            if (mapped) {
                const auto t0 = std::chrono::steady_clock::now();
                unsigned long long w;
                unsigned spins = 0;
                while ((((w = *prog) >> 1) < (unsigned long long)(it - 2)) && !(w & 1ULL)) {
                    if ((++spins & 0xfffu) == 0) {
                        if (cudaStreamQuery(st) != cudaErrorNotReady) break;                 // everything enqueued has finished (abort / error path)
                        if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(90)) break;
                        std::this_thread::yield();
                    }
                }
                if (*prog & 1ULL) break;
            } else {
                const int slot_prev = it & 1;
                JB_CUDA(ctx, cudaEventSynchronize(K->ev[slot_prev]));
                if (K->h_flags[slot_prev * KS_SIZE + KS_DONE] != 0.0) break;
            }
        }
        kflags = (it << 1) | (it == 1 ? 1 : 0);
        JB_CUDA(ctx, cudaLaunchCooperativeKernel((const void*)bicgstab_rb_iteration_kernel, dim3(H.grid), dim3(JB_S2_THREADS), kargs, sizeof(S2Smem), st));
        ctx->launches++;
        if (!mapped) {
            const int slot = it & 1;
            JB_CUDA(ctx, cudaMemcpyAsync(K->h_flags + slot * KS_SIZE, K->d_sc.p, KS_SIZE * sizeof(double), cudaMemcpyDeviceToHost, st));
            JB_CUDA(ctx, cudaEventRecord(K->ev[slot], st));
        }
    }
    if (ctx->prof_on) jb_prof_end(ctx);
    JB_CUDA(ctx, cudaMemcpyAsync(K->h_flags + 3 * KS_SIZE, K->d_sc.p, KS_SIZE * sizeof(double), cudaMemcpyDeviceToHost, st));
    unsigned long long h_state[PH_COUNT + 4];
    unsigned h_sync[8];
    JB_CUDA(ctx, cudaMemcpyAsync(h_state, H.d_state.p, sizeof(h_state), cudaMemcpyDeviceToHost, st));
    JB_CUDA(ctx, cudaMemcpyAsync(h_sync, H.d_sync.p, sizeof(h_sync), cudaMemcpyDeviceToHost, st));
    JB_CUDA(ctx, cudaStreamSynchronize(st));
    const PKSync* hs = reinterpret_cast<const PKSync*>(h_sync);
    if (hs->abort != 0) {
        cudaMemsetAsync(H.d_sync.p, 0, sizeof(h_sync), st);
        cudaStreamSynchronize(st);
        JB_FAIL(ctx, JB_ERR_CUDA, hs->abort == 2 ? "persistent BiCGStab: a peer did not answer (collective timed out)"
                                                 : "persistent BiCGStab: grid barrier timed out");
    }
    if (K->dist) jb_dist_pk_done(K->dist, h_state[0], h_state[1]);
    const double* f = K->h_flags + 3 * KS_SIZE;
    const int niter = (int)f[KS_ITER];
    if (iters) *iters = niter;
    if (hist && hist_cap > 0) {
        const int nh = std::min(hist_cap, niter + 1);
        JB_CUDA(ctx, cudaMemcpy(hist, K->d_hist.p, nh * sizeof(double), cudaMemcpyDeviceToHost));
    }
    if (ctx->prof_on) {
        const unsigned long long* ns = h_state + 4;
        const double ms = 1e-6;
        for (int q = 0; q < PH_COUNT; q++) H.phase_ms[q] += ms * (double)ns[q];
        H.phase_iters += niter; H.phase_solves += 1;
        ctx->prof_extra_ms[JB_PROF_ILU_APPLY] += ms * (double)(ns[PH_A1] + ns[PH_A2] + ns[PH_A4] + ns[PH_A5]);
        ctx->prof_extra_cnt[JB_PROF_ILU_APPLY] += 2 * niter;
        ctx->prof_extra_ms[JB_PROF_SPMV] += ms * (double)(ns[PH_A3] + ns[PH_A6]);
        ctx->prof_extra_cnt[JB_PROF_SPMV] += 2 * niter;
        ctx->prof_extra_ms[JB_PROF_VECTOR] += ms * (double)(ns[PH_INIT] + ns[PH_V1] + ns[PH_V2] + ns[PH_V3] + ns[PH_FINAL]);
        ctx->prof_extra_cnt[JB_PROF_VECTOR] += 1;
        ctx->prof_extra_ms[JB_PROF_OTHER] += ms * (double)ns[PH_HALO];
        ctx->prof_extra_cnt[JB_PROF_OTHER] += 2 * niter;
    }
    int status = (int)f[KS_STATUS];
    if (f[KS_DONE] == 0.0) status = JB_NOT_CONVERGED;
    *status_out = status;
    return JB_OK;
}
