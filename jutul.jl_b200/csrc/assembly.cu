// Finite-volume residual/Jacobian assembly for sm_100a.
//
// Reference path replaced (variant A, hard-coded TPFA,
// src/conservation/conservation.jl): update_accumulation! (:558-568),
// update_half_face_flux_tpfa! (:606-626, AD w.r.t. the self cell of each
// half-face = "local perspective", src/ad/local_ad.jl:54-79), apply_forces!
// (src/models.jl:889-901) and fill_conservation_eq! (:373-430). The reference
// stores ne x nhf Duals (285 MB at 1M cells) between evaluate and fill; here the
// three stages are fused so a flux never leaves registers:
//
//   twophase_state_kernel   — "secondary variables": one 32-byte record per cell
//                             {p, Sw, rho_w, rho_o}; a neighbour gather is then a
//                             single sector.
//   twophase_assemble_kernel — row-owner form: LPC lanes per cell, one half-face
//                             per lane. A lane evaluates the two-point flux from
//                             the self perspective (-> r and the diagonal block,
//                             combined over the lanes by shuffles) and from the
//                             neighbour's perspective (-> J[self, other] =
//                             -dF_{other->self}/dx_other, exactly the entry the
//                             reference writes from the neighbour's loop), and
//                             writes its 2x2 block into the row it owns. Every
//                             Jacobian entry is written exactly once: no zeroing,
//                             no atomics, deterministic.
//   twophase_faces_kernel   — variant B of the reference (fvm_face_assembly!,
//                             src/conservation/fvm_assembly.jl:253-283): one lane
//                             per face, off-diagonal blocks written directly,
//                             r and diagonal blocks scattered with warp-aggregated
//                             FP64 atomics (jb_reduce.cuh; order-dependent rounding).
//
// Physics: builder-defined two-phase immiscible plug-in for the face_flux! slot
// (src/conservation/conservation.jl:642-649), see SURVEY.md §8(d) / DESIGN.md.
#include "jb_internal.cuh"
#include "jb_reduce.cuh"
#include "jb_tma.cuh"
#include <algorithm>
#include <cstddef>
#include <map>

// inv_mu: the kernels multiply by 1/mu instead of dividing (FP64 division is a ~30-instruction sequence)
struct TPParams { double rho0[2], c[2], mu[2], p0, inv_mu[2]; };

__global__ void __launch_bounds__(256) twophase_state_kernel(i64 nc, TPParams P, const double* __restrict__ p, const double* __restrict__ s,
                                                             double4* __restrict__ rec) {
    for (i64 c = (i64)blockIdx.x * blockDim.x + threadIdx.x; c < nc; c += (i64)gridDim.x * blockDim.x) {
        const double pc = __ldg(p + c);
        const double sw = __ldg(s + 2 * c);
        double4 o;
        o.x = pc; o.y = sw;
        o.z = P.rho0[0] * exp(P.c[0] * (pc - P.p0));
        o.w = P.rho0[1] * exp(P.c[1] * (pc - P.p0));
        rec[c] = o;
    }
}

__device__ __forceinline__ double4 ld_rec(const double4* __restrict__ p) {
    const double2* q = reinterpret_cast<const double2*>(p);
    const double2 a = __ldg(q), b = __ldg(q + 1);
    return make_double4(a.x, a.y, b.x, b.y);
}

struct CellP {            // properties of a cell needed by the flux, with d/dp and d/dSw of the mobility
    double p, rho[2], mob[2], dmob_dp[2], dmob_ds[2];
};
__device__ __forceinline__ CellP cell_props(const TPParams& P, const double4 r) {
    CellP o;
    o.p = r.x;
    const double S[2] = {r.y, 1.0 - r.y};
    const double dS[2] = {1.0, -1.0};
    o.rho[0] = r.z; o.rho[1] = r.w;
#pragma unroll
    for (int a = 0; a < 2; a++) {
        const double kr = S[a] * S[a];
        o.mob[a] = (o.rho[a] * kr) * P.inv_mu[a];
        o.dmob_dp[a] = P.c[a] * o.mob[a];
        o.dmob_ds[a] = (o.rho[a] * (2.0 * S[a] * dS[a])) * P.inv_mu[a];
    }
    return o;
}
// Flux of phase a out of `self` towards `other`, value and d/d(p_self), d/d(Sw_self).
__device__ __forceinline__ void flux_phase(const TPParams& P, const CellP& self, const CellP& other, int a, double T, double sgdz,
                                           double& F, double& dFdp, double& dFds) {
    const double rho_avg = 0.5 * (self.rho[a] + other.rho[a]);
    const double theta = self.p - other.p + sgdz * rho_avg;
    const double dtheta = 1.0 + sgdz * (0.5 * (P.c[a] * self.rho[a]));
    const double q = T * theta;
    const double dq = T * dtheta;
    if (q > 0) {  // upstream = self
        F = self.mob[a] * q;
        dFdp = self.mob[a] * dq + q * self.dmob_dp[a];
        dFds = q * self.dmob_ds[a];
    } else {      // upstream = other: constant w.r.t. self
        F = other.mob[a] * q;
        dFdp = other.mob[a] * dq;
        dFds = 0.0;
    }
}

template <int LPC, bool JAC>
__global__ void __launch_bounds__(256, 4) twophase_assemble_kernel(i64 nc, TPParams P, const int32_t* __restrict__ hf_pos,
                                                                const int32_t* __restrict__ hf_other, const int32_t* __restrict__ hf_rowpos,
                                                                const double* __restrict__ hf_T, const double* __restrict__ hf_sgdz,
                                                                const int32_t* __restrict__ diag_pos, const double4* __restrict__ rec,
                                                                const double* __restrict__ pv, const double* __restrict__ M0,
                                                                const double* __restrict__ src, double inv_dt, double dt,
                                                                double* __restrict__ nz, double* __restrict__ r) {
    constexpr int CELLS_PER_WARP = 32 / LPC;
    const int lane = threadIdx.x % LPC;
    const i64 warps_per_cta = blockDim.x >> 5;
    const i64 stride = (i64)gridDim.x * warps_per_cta * CELLS_PER_WARP;
    for (i64 base = ((i64)blockIdx.x * warps_per_cta + (threadIdx.x >> 5)) * CELLS_PER_WARP; base < nc; base += stride) {
        const i64 c = base + (threadIdx.x & 31) / LPC;
        const bool live = c < nc;
        // accumulators: residual (2) and diagonal block (column-major: [e + 2*d])
        double racc[2] = {0.0, 0.0};
        double dacc[4] = {0.0, 0.0, 0.0, 0.0};
        if (live) {
            const double4 rs = ld_rec(rec + c);
            const CellP self = cell_props(P, rs);
            const int32_t h0 = __ldg(hf_pos + c), h1 = __ldg(hf_pos + c + 1);
            for (int32_t i = h0 + lane; i < h1; i += LPC) {
                const int32_t o = __ldg(hf_other + i);
                const double T = __ldg(hf_T + i);
                const double sg = __ldg(hf_sgdz + i);
                const CellP other = cell_props(P, ld_rec(rec + o));
                double blk[4];
#pragma unroll
                for (int a = 0; a < 2; a++) {
                    double F, dFdp, dFds;
                    flux_phase(P, self, other, a, T, sg, F, dFdp, dFds);
                    racc[a] += F;
                    dacc[a] += dFdp;
                    dacc[2 + a] += dFds;
                    if (JAC) {
                        // the neighbour's half-face towards us: J[self, other] = -dF_{o->c}/dx_o
                        double Fo, dFo_dp, dFo_ds;
                        flux_phase(P, other, self, a, T, -sg, Fo, dFo_dp, dFo_ds);
                        blk[a] = -dFo_dp;
                        blk[2 + a] = -dFo_ds;
                    }
                }
                if (JAC) {
                    double2* dst = reinterpret_cast<double2*>(nz + (size_t)__ldg(hf_rowpos + i) * 4);
                    dst[0] = make_double2(blk[0], blk[1]);
                    dst[1] = make_double2(blk[2], blk[3]);
                }
            }
            if (lane == 0) {
                // accumulation term (M - M0)/dt with its partials, plus sources on the diagonal entries
                const double S[2] = {rs.y, 1.0 - rs.y};
                const double dS[2] = {1.0, -1.0};
                const double pvc = __ldg(pv + c);
#pragma unroll
                for (int a = 0; a < 2; a++) {
                    const double mass = pvc * (self.rho[a] * S[a]);
                    double acc = (mass - __ldg(M0 + 2 * c + a)) * inv_dt;
                    if (src) acc += __ldg(src + 2 * c + a);
                    racc[a] += acc;
                    dacc[a] += (pvc * (P.c[a] * self.rho[a] * S[a])) * inv_dt;
                    dacc[2 + a] += (pvc * (self.rho[a] * dS[a])) * inv_dt;
                }
            }
        }
#pragma unroll
        for (int o = LPC / 2; o > 0; o >>= 1) {
#pragma unroll
            for (int e = 0; e < 2; e++) racc[e] += __shfl_xor_sync(0xffffffffu, racc[e], o);
            if (JAC) {
#pragma unroll
                for (int q = 0; q < 4; q++) dacc[q] += __shfl_xor_sync(0xffffffffu, dacc[q], o);
            }
        }
        if (live && lane == 0) {
            reinterpret_cast<double2*>(r)[c] = make_double2(racc[0], racc[1]);
            if (JAC) {
                double2* dst = reinterpret_cast<double2*>(nz + (size_t)__ldg(diag_pos + c) * 4);
                dst[0] = make_double2(dacc[0], dacc[1]);
                dst[1] = make_double2(dacc[2], dacc[3]);
            }
        }
    }
}


// ---- row-chunk stream form of the fused assembly (default) --------------------------------------------------
// A CTA owns a chunk of <= 128 consecutive cells / <= 1024 half-faces. All threads stream the chunk's half-face
// arrays (other, T, sign*gdz, Jacobian position) with unit stride, find the owning cell by bisection of the chunk's
// offsets in shared memory, gather the two 32-byte cell records, evaluate both perspectives of the two-point flux,
// write the off-diagonal block straight to its slot in the row and park the (residual, diagonal) contribution of the
// half-face in shared memory. One thread per cell then adds accumulation + half-faces in conn_pos order (the order of
// fill_conservation_eq!, src/conservation/conservation.jl:373-430) and writes r and the diagonal block.
#define JB_ASM_CELLS 64
#define JB_ASM_HF 512
#define JB_ASM_THREADS 128
template <bool JAC>
__global__ void __launch_bounds__(JB_ASM_THREADS) twophase_assemble_stream_kernel(int nchunks, const int32_t* __restrict__ chunk_ptr, TPParams P,
                                                                       const int32_t* __restrict__ hf_pos, const int32_t* __restrict__ hf_other,
                                                                       const int32_t* __restrict__ hf_rowpos, const double* __restrict__ hf_T,
                                                                       const double* __restrict__ hf_sgdz, const int32_t* __restrict__ diag_pos,
                                                                       const double4* __restrict__ rec, const double* __restrict__ pv,
                                                                       const double* __restrict__ M0, const double* __restrict__ src, double inv_dt,
                                                                       double* __restrict__ nz, double* __restrict__ r) {
    __shared__ int32_t s_pos[JB_ASM_CELLS + 1];
    __shared__ unsigned char s_owner[JB_ASM_HF];   // chunk-local cell of each half-face
    extern __shared__ double s_part[];     // [6][JB_ASM_HF + pad]: r_w, r_o, d(w)/dp, d(o)/dp, d(w)/dS, d(o)/dS per half-face
    constexpr int NP = JAC ? 6 : 2;
    constexpr int LD = JB_ASM_HF + 1;
    for (int ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
        const int c0 = __ldg(chunk_ptr + ch), nr = __ldg(chunk_ptr + ch + 1) - c0;
        for (int j = threadIdx.x; j <= nr; j += blockDim.x) s_pos[j] = __ldg(hf_pos + c0 + j);
        __syncthreads();
        const int base = s_pos[0], cnt = s_pos[nr] - base;
        for (int j = threadIdx.x; j < nr; j += blockDim.x)
            for (int e = s_pos[j] - base; e < s_pos[j + 1] - base; e++) s_owner[e] = (unsigned char)j;
        __syncthreads();
        for (int e = threadIdx.x; e < cnt; e += blockDim.x) {
            const int i = base + e;
            const int32_t o = __ldcs(hf_other + i);
            const double T = __ldcs(hf_T + i), sg = __ldcs(hf_sgdz + i);
            const CellP self = cell_props(P, ld_rec(rec + c0 + s_owner[e]));
            const CellP other = cell_props(P, ld_rec(rec + o));
            double part[6], blk[4];
#pragma unroll
            for (int a = 0; a < 2; a++) {
                double F, dFdp, dFds;
                flux_phase(P, self, other, a, T, sg, F, dFdp, dFds);
                part[a] = F; part[2 + a] = dFdp; part[4 + a] = dFds;
                if (JAC) {
                    double Fo, dFo_dp, dFo_ds;
                    flux_phase(P, other, self, a, T, -sg, Fo, dFo_dp, dFo_ds);   // the neighbour's half-face towards us
                    blk[a] = -dFo_dp; blk[2 + a] = -dFo_ds;
                }
            }
            if (JAC) {
                double2* dst = reinterpret_cast<double2*>(nz + (size_t)__ldcs(hf_rowpos + i) * 4);
                dst[0] = make_double2(blk[0], blk[1]);
                dst[1] = make_double2(blk[2], blk[3]);
            }
#pragma unroll
            for (int q = 0; q < NP; q++) s_part[q * LD + e] = part[q];
        }
        __syncthreads();
        // two threads per cell: thread (j, a) owns equation a -> r[a], d/dp, d/dS
        for (int w = threadIdx.x; w < 2 * nr; w += blockDim.x) {
            const int j = w >> 1, a = w & 1;
            const size_t c = (size_t)c0 + j;
            const double4 rs = ld_rec(rec + c);
            const double S = a ? 1.0 - rs.y : rs.y;
            const double dS = a ? -1.0 : 1.0;
            const double rho = a ? rs.w : rs.z;
            const double pvc = __ldg(pv + c);
            const double mass = pvc * (rho * S);
            double ar = (mass - __ldg(M0 + 2 * c + a)) * inv_dt;
            if (src) ar += __ldg(src + 2 * c + a);
            double adp = (pvc * ((a ? P.c[1] : P.c[0]) * rho * S)) * inv_dt;
            double ads = (pvc * (rho * dS)) * inv_dt;
            for (int e = s_pos[j] - base; e < s_pos[j + 1] - base; e++) {
                ar += s_part[a * LD + e];
                if (JAC) { adp += s_part[(2 + a) * LD + e]; ads += s_part[(4 + a) * LD + e]; }
            }
            r[2 * c + a] = ar;
            if (JAC) {
                double* dst = nz + (size_t)__ldg(diag_pos + c) * 4;
                dst[a] = adp; dst[2 + a] = ads;
            }
        }
        __syncthreads();
    }
}

// ---- TMA-staged persistent form (default) ---------------------------------------------------------------------
// Same row-owner arithmetic and the same summation order as the stream kernel above; what changes is how the bytes
// move. A persistent CTA walks a host-built chunk table (<= 64 cells / <= 448 half-faces per chunk). For every chunk
// ONE thread issues 1-D bulk copies (cp.async.bulk, jb_tma.cuh) of all contiguous inputs of the chunk — half-face
// neighbour list, T, sign*gdz, 16-bit Jacobian slots, half-face offsets, diagonal slots, pore volumes, M0 and the
// chunk's own 32-byte cell records — into one of two shared-memory stages, completing on an mbarrier; the copies of
// chunk i+1 fly while chunk i is evaluated. The only load a thread still waits for is the 32-byte neighbour-record
// gather (one 256-bit LDG), issued for all of a thread's half-faces before the first flux is evaluated.
// Byte diet relative to the stream kernel: Jacobian slot as uint16 offset from the chunk's first row (2 B instead of
// 4 B per half-face); the two perspectives of a half-face share rho_avg / theta / q (flux_pair); the dense source
// buffer is read only by chunks that contain a source cell; Jacobian / residual stores are streaming (st.cs) so the
// cell records stay in L2; the chunk table is ordered so that chunks of different colours that are each other's
// neighbours are processed at the same time (their records are then fetched from HBM once, not twice).
#define JB_ASM2_CELLS 64
#define JB_ASM2_HF 448
#define JB_ASM2_THREADS 128
#define JB_ASM2_U 4
static_assert(JB_ASM2_U * JB_ASM2_THREADS >= JB_ASM2_HF, "one unrolled pass must cover a chunk");
static_assert(2 * JB_ASM2_CELLS <= JB_ASM2_THREADS, "two threads per cell in the row phase");

struct __align__(16) Asm2Stage {     // every member starts on a 16-byte boundary and can hold the widened window
    double T[JB_ASM2_HF + 2];
    double sg[JB_ASM2_HF + 2];
    double rec[JB_ASM2_CELLS * 4];
    double M0[JB_ASM2_CELLS * 2];
    double pv[JB_ASM2_CELLS + 2];
    int32_t other[JB_ASM2_HF + 4];
    int32_t hfpos[JB_ASM2_CELLS + 8];
    int32_t diag[JB_ASM2_CELLS + 4];
    uint16_t lp[JB_ASM2_HF + 8];
};
static_assert(offsetof(Asm2Stage, sg) % 16 == 0 && offsetof(Asm2Stage, rec) % 16 == 0 && offsetof(Asm2Stage, M0) % 16 == 0 &&
              offsetof(Asm2Stage, pv) % 16 == 0 && offsetof(Asm2Stage, other) % 16 == 0 && offsetof(Asm2Stage, hfpos) % 16 == 0 &&
              offsetof(Asm2Stage, diag) % 16 == 0 && offsetof(Asm2Stage, lp) % 16 == 0 && sizeof(Asm2Stage) % 16 == 0, "stage alignment");

__device__ __forceinline__ void ld_rec256(const double* __restrict__ rec, int32_t c, double (&o)[4]) {
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(o[0]), "=d"(o[1]), "=d"(o[2]), "=d"(o[3]) : "l"(rec + 4 * (size_t)c));
}

// Both perspectives of the two-point flux of phase a across one half-face self -> other:
//   F, dFdp, dFds : flux out of self and its partials w.r.t. (p, Sw) of self  (residual + diagonal block of row self)
//   Bp, Bs        : -d F_{other->self} / d (p, Sw) of other                      (the entry J[self, other])
// q_{other->self} = -q exactly (same rho_avg, negated terms), so one theta / q / upwind decision serves both.
// ca = compressibility, imu = 1/viscosity of the phase, dS = dS_a/dSw (+1 water, -1 oil).
__device__ __forceinline__ void flux_pair(double ca, double imu, double dS, double ps, double Ss, double rs, double po, double So, double ro,
                                          double T, double sg, double& F, double& dFdp, double& dFds, double& Bp, double& Bs) {
    const double mob_s = (rs * (Ss * Ss)) * imu;
    const double mob_o = (ro * (So * So)) * imu;
    const double rho_avg = 0.5 * (rs + ro);
    const double theta = ps - po + sg * rho_avg;
    const double q = T * theta;
    const double dq_s = T * (1.0 + sg * (0.5 * (ca * rs)));
    const double dq_o = T * (1.0 + (-sg) * (0.5 * (ca * ro)));
    const bool ups = q > 0, upo = q < 0;
    const double mF = ups ? mob_s : mob_o;   // upstream mobility seen from self (q == 0 / NaN: other, as upw_flux)
    const double mN = upo ? mob_o : mob_s;   // upstream mobility seen from the neighbour
    const double u_rho = ups ? rs : ro, u_S = ups ? Ss : So;
    const double dm_dp = ca * mF;
    const double dm_ds = (u_rho * (2.0 * u_S * dS)) * imu;
    F = mF * q;
    dFdp = mF * dq_s;
    dFds = 0.0;
    double np = mN * dq_o, ns = 0.0;
    if (ups) { dFdp += q * dm_dp; dFds = q * dm_ds; }
    if (upo) { const double qo = -q; np += qo * dm_dp; ns = qo * dm_ds; }
    Bp = -np;
    Bs = -ns;
}

template <bool JAC>
__global__ void __launch_bounds__(JB_ASM2_THREADS, 4) twophase_assemble_tma_kernel(
    int nchunks, const Asm2Chunk* __restrict__ table, TPParams P, const int32_t* __restrict__ hf_pos, const int32_t* __restrict__ hf_other,
    const uint16_t* __restrict__ hf_lp, const double* __restrict__ hf_T, const double* __restrict__ hf_sgdz, const int32_t* __restrict__ diag_pos,
    const double* __restrict__ rec, const double* __restrict__ pv, const double* __restrict__ M0, const double* __restrict__ src, double inv_dt,
    double* __restrict__ nz, double* __restrict__ r) {
    constexpr int NP = JAC ? 6 : 2;
    constexpr int LD = JB_ASM2_HF + 1;
    extern __shared__ __align__(128) unsigned char asm2_smem[];
    Asm2Stage* stage = reinterpret_cast<Asm2Stage*>(asm2_smem);
    double* s_part = reinterpret_cast<double*>(asm2_smem + 2 * sizeof(Asm2Stage));   // [NP][LD]
    Asm2Chunk* s_meta = reinterpret_cast<Asm2Chunk*>(s_part + NP * LD + (NP * LD & 1));
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_meta + 2);
    unsigned char* s_owner = reinterpret_cast<unsigned char*>(s_bar + 2);
    const int tid = threadIdx.x;

    auto issue = [&](int k, int s) {   // thread 0 only
        const int4* tp = reinterpret_cast<const int4*>(table + k);
        const int4 t0 = __ldg(tp), t1 = __ldg(tp + 1);
        Asm2Chunk ci;
        ci.c0 = t0.x; ci.nr = t0.y; ci.hf0 = t0.z; ci.cnt = t0.w; ci.row0 = t1.x; ci.flags = t1.y; ci.pad0 = 0; ci.pad1 = 0;
        s_meta[s] = ci;
        Asm2Stage& S = stage[s];
        const JbSpan<double> sT(hf_T, ci.hf0, ci.cnt), sG(hf_sgdz, ci.hf0, ci.cnt), sPv(pv, ci.c0, ci.nr);
        const JbSpan<int32_t> sO(hf_other, ci.hf0, ci.cnt), sHp(hf_pos, ci.c0, ci.nr + 1), sD(diag_pos, ci.c0, ci.nr);
        const JbSpan<uint16_t> sL(hf_lp, ci.hf0, ci.cnt);
        const uint32_t bRec = (uint32_t)ci.nr * 32u, bM0 = (uint32_t)ci.nr * 16u;
        uint32_t total = sT.bytes + sG.bytes + sPv.bytes + sO.bytes + sHp.bytes + bRec + bM0;
        if (JAC) total += sD.bytes + sL.bytes;
        jb_mbar_expect_tx(&s_bar[s], total);
        const uint64_t pol = jb_policy_evict_first();
        if (sT.bytes) jb_bulk_g2s_hint(S.T, sT.src, sT.bytes, &s_bar[s], pol);
        if (sG.bytes) jb_bulk_g2s_hint(S.sg, sG.src, sG.bytes, &s_bar[s], pol);
        if (sO.bytes) jb_bulk_g2s_hint(S.other, sO.src, sO.bytes, &s_bar[s], pol);
        if (JAC && sL.bytes) jb_bulk_g2s_hint(S.lp, sL.src, sL.bytes, &s_bar[s], pol);
        if (sHp.bytes) jb_bulk_g2s_hint(S.hfpos, sHp.src, sHp.bytes, &s_bar[s], pol);
        if (JAC && sD.bytes) jb_bulk_g2s_hint(S.diag, sD.src, sD.bytes, &s_bar[s], pol);
        if (sPv.bytes) jb_bulk_g2s_hint(S.pv, sPv.src, sPv.bytes, &s_bar[s], pol);
        if (bM0) jb_bulk_g2s_hint(S.M0, M0 + 2 * (size_t)ci.c0, bM0, &s_bar[s], pol);
        if (bRec) jb_bulk_g2s(S.rec, rec + 4 * (size_t)ci.c0, bRec, &s_bar[s]);
    };

    if (tid == 0) {
        jb_mbar_init(&s_bar[0], 1);
        jb_mbar_init(&s_bar[1], 1);
        jb_mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) {
        if ((int)blockIdx.x < nchunks) issue(blockIdx.x, 0);
        if ((int)(blockIdx.x + gridDim.x) < nchunks) issue(blockIdx.x + gridDim.x, 1);
    }
    __syncthreads();

    int it = 0;
    for (int k = blockIdx.x; k < nchunks; k += gridDim.x, it++) {
        const int s = it & 1;
        jb_mbar_wait(&s_bar[s], (uint32_t)(it >> 1) & 1u);
        const Asm2Stage& S = stage[s];
        const int c0 = s_meta[s].c0, nr = s_meta[s].nr, hf0 = s_meta[s].hf0, cnt = s_meta[s].cnt, row0 = s_meta[s].row0;
        const int flags = s_meta[s].flags;
        const int l8 = jb_span_lead<double>((size_t)hf0), l4 = jb_span_lead<int32_t>((size_t)hf0), l2 = jb_span_lead<uint16_t>((size_t)hf0);
        const int lc4 = jb_span_lead<int32_t>((size_t)c0), lc8 = jb_span_lead<double>((size_t)c0);
        for (int j = tid; j < nr; j += JB_ASM2_THREADS) {
            const int e1 = S.hfpos[lc4 + j + 1] - hf0;
            for (int e = S.hfpos[lc4 + j] - hf0; e < e1; e++) s_owner[e] = (unsigned char)j;
        }
        __syncthreads();
        {
            double ro[JB_ASM2_U][4];
            double Tv[JB_ASM2_U], sgv[JB_ASM2_U];
            int own[JB_ASM2_U];
            // all neighbour gathers of this thread go out before the first flux is evaluated
#pragma unroll
            for (int u = 0; u < JB_ASM2_U; u++) {
                const int e = tid + u * JB_ASM2_THREADS;
                if (e < cnt) ld_rec256(rec, S.other[l4 + e], ro[u]);
            }
#pragma unroll
            for (int u = 0; u < JB_ASM2_U; u++) {
                const int e = tid + u * JB_ASM2_THREADS;
                if (e < cnt) { Tv[u] = S.T[l8 + e]; sgv[u] = S.sg[l8 + e]; own[u] = s_owner[e]; }
            }
#pragma unroll
            for (int u = 0; u < JB_ASM2_U; u++) {
                const int e = tid + u * JB_ASM2_THREADS;
                if (e < cnt) {
                    const double2 sa = *reinterpret_cast<const double2*>(S.rec + 4 * own[u]);
                    const double2 sb = *reinterpret_cast<const double2*>(S.rec + 4 * own[u] + 2);
                    double part[6], blk[4];
#pragma unroll
                    for (int a = 0; a < 2; a++) {
                        const double Ss = a ? 1.0 - sa.y : sa.y, So = a ? 1.0 - ro[u][1] : ro[u][1];
                        const double rs = a ? sb.y : sb.x, rn = a ? ro[u][3] : ro[u][2];
                        flux_pair(P.c[a], P.inv_mu[a], a ? -1.0 : 1.0, sa.x, Ss, rs, ro[u][0], So, rn, Tv[u], sgv[u], part[a], part[2 + a], part[4 + a],
                                  blk[a], blk[2 + a]);
                    }
                    if (JAC) {
                        double2* dst = reinterpret_cast<double2*>(nz + ((size_t)row0 + S.lp[l2 + e]) * 4);
                        __stcs(dst, make_double2(blk[0], blk[1]));
                        __stcs(dst + 1, make_double2(blk[2], blk[3]));
                    }
#pragma unroll
                    for (int q = 0; q < NP; q++) s_part[q * LD + e] = part[q];
                }
            }
        }
        __syncthreads();
        // two threads per cell: thread (j, a) owns equation a -> r[a], d/dp, d/dS; accumulation first, then the half-faces
        // in conn_pos order (fill_conservation_eq!, src/conservation/conservation.jl:373-430)
        if (tid < 2 * nr) {
            const int j = tid >> 1, a = tid & 1;
            const size_t c = (size_t)c0 + j;
            const double sw = S.rec[4 * j + 1];
            const double Sa = a ? 1.0 - sw : sw;
            const double dS = a ? -1.0 : 1.0;
            const double rho = S.rec[4 * j + 2 + a];
            const double pvc = S.pv[lc8 + j];
            const double mass = pvc * (rho * Sa);
            double ar = (mass - S.M0[2 * j + a]) * inv_dt;
            if (flags & 1) ar += __ldg(src + 2 * c + a);
            double adp = (pvc * ((a ? P.c[1] : P.c[0]) * rho * Sa)) * inv_dt;
            double ads = (pvc * (rho * dS)) * inv_dt;
            const int e1 = S.hfpos[lc4 + j + 1] - hf0;
            for (int e = S.hfpos[lc4 + j] - hf0; e < e1; e++) {
                ar += s_part[a * LD + e];
                if (JAC) { adp += s_part[(2 + a) * LD + e]; ads += s_part[(4 + a) * LD + e]; }
            }
            r[2 * c + a] = ar;
            if (JAC) {
                double* dst = nz + (size_t)S.diag[lc4 + j] * 4;
                __stcs(dst + a, adp);
                __stcs(dst + 2 + a, ads);
            }
        }
        __syncthreads();   // stage s and s_part are free again
        if (tid == 0) {
            const int k2 = k + 2 * (int)gridDim.x;
            if (k2 < nchunks) issue(k2, s);
        }
    }
}

// ---- TMA-staged, lane-pair-per-cell form (default) --------------------------------------------------------------
// Same staging as above, but no shared-memory partials and no CTA barrier in steady state: lane pair (2j, 2j+1) of a warp
// owns chunk cell j — lane a evaluates equation (phase) a of every half-face of the cell in conn_pos order and keeps
// r_a, dr_a/dp, dr_a/dSw in registers, so the sums are formed exactly in the order of fill_conservation_eq!
// (accumulation term first, then the half-faces, src/conservation/conservation.jl:373-430). The pair writes the 2x2
// off-diagonal block of each half-face with two 16-byte streaming stores (one value exchanged by shuffle). A warp's 16
// cells run independently of the other warps; a stage is handed back for the next bulk copy by the LAST warp to finish
// with it (shared-memory counter), so nothing ever waits on a __syncthreads and the phases of different warps overlap.
// NG = neighbour gathers in flight per lane (the records of NG half-faces are requested before the first flux is
// evaluated).
template <bool JAC, int NG, int MINB>
__global__ void __launch_bounds__(JB_ASM2_THREADS, MINB) twophase_assemble_pair_kernel(
    int nchunks, const Asm2Chunk* __restrict__ table, TPParams P, const int32_t* __restrict__ hf_pos, const int32_t* __restrict__ hf_other,
    const uint16_t* __restrict__ hf_lp, const double* __restrict__ hf_T, const double* __restrict__ hf_sgdz, const int32_t* __restrict__ diag_pos,
    const double* __restrict__ rec, const double* __restrict__ pv, const double* __restrict__ M0, const double* __restrict__ src, double inv_dt,
    double* __restrict__ nz, double* __restrict__ r) {
    extern __shared__ __align__(128) unsigned char asm2_smem[];
    Asm2Stage* stage = reinterpret_cast<Asm2Stage*>(asm2_smem);
    Asm2Chunk* s_meta = reinterpret_cast<Asm2Chunk*>(asm2_smem + 2 * sizeof(Asm2Stage));
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_meta + 2);
    int* s_cnt = reinterpret_cast<int*>(s_bar + 2);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NWARPS = JB_ASM2_THREADS / 32;
    constexpr int CPW = JB_ASM2_CELLS / NWARPS;   // cells per warp
    static_assert(CPW * 2 == 32, "one lane pair per cell");

    auto issue = [&](int k, int s) {   // one lane
        const int4* tp = reinterpret_cast<const int4*>(table + k);
        const int4 t0 = __ldg(tp), t1 = __ldg(tp + 1);
        Asm2Chunk ci;
        ci.c0 = t0.x; ci.nr = t0.y; ci.hf0 = t0.z; ci.cnt = t0.w; ci.row0 = t1.x; ci.flags = t1.y; ci.pad0 = 0; ci.pad1 = 0;
        s_meta[s] = ci;
        Asm2Stage& S = stage[s];
        const JbSpan<double> sT(hf_T, ci.hf0, ci.cnt), sG(hf_sgdz, ci.hf0, ci.cnt), sPv(pv, ci.c0, ci.nr);
        const JbSpan<int32_t> sO(hf_other, ci.hf0, ci.cnt), sHp(hf_pos, ci.c0, ci.nr + 1), sD(diag_pos, ci.c0, ci.nr);
        const JbSpan<uint16_t> sL(hf_lp, ci.hf0, ci.cnt);
        const uint32_t bRec = (uint32_t)ci.nr * 32u, bM0 = (uint32_t)ci.nr * 16u;
        uint32_t total = sT.bytes + sG.bytes + sPv.bytes + sO.bytes + sHp.bytes + bRec + bM0;
        if (JAC) total += sD.bytes + sL.bytes;
        jb_mbar_expect_tx(&s_bar[s], total);
        const uint64_t pol = jb_policy_evict_first();
        if (sO.bytes) jb_bulk_g2s_hint(S.other, sO.src, sO.bytes, &s_bar[s], pol);
        if (sHp.bytes) jb_bulk_g2s_hint(S.hfpos, sHp.src, sHp.bytes, &s_bar[s], pol);
        if (bRec) jb_bulk_g2s(S.rec, rec + 4 * (size_t)ci.c0, bRec, &s_bar[s]);
        if (sT.bytes) jb_bulk_g2s_hint(S.T, sT.src, sT.bytes, &s_bar[s], pol);
        if (sG.bytes) jb_bulk_g2s_hint(S.sg, sG.src, sG.bytes, &s_bar[s], pol);
        if (JAC && sL.bytes) jb_bulk_g2s_hint(S.lp, sL.src, sL.bytes, &s_bar[s], pol);
        if (JAC && sD.bytes) jb_bulk_g2s_hint(S.diag, sD.src, sD.bytes, &s_bar[s], pol);
        if (sPv.bytes) jb_bulk_g2s_hint(S.pv, sPv.src, sPv.bytes, &s_bar[s], pol);
        if (bM0) jb_bulk_g2s_hint(S.M0, M0 + 2 * (size_t)ci.c0, bM0, &s_bar[s], pol);
    };

    if (tid == 0) {
        jb_mbar_init(&s_bar[0], 1);
        jb_mbar_init(&s_bar[1], 1);
        s_cnt[0] = 0; s_cnt[1] = 0;
        jb_mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) {
        if ((int)blockIdx.x < nchunks) issue(blockIdx.x, 0);
        if ((int)(blockIdx.x + gridDim.x) < nchunks) issue(blockIdx.x + gridDim.x, 1);
    }
    __syncthreads();

    const int jl = lane >> 1, a = lane & 1;
    const unsigned pairmask = 3u << (lane & ~1);
    const double dS = a ? -1.0 : 1.0;
    const double ca = a ? P.c[1] : P.c[0];
    const double imu = a ? P.inv_mu[1] : P.inv_mu[0];
    int it = 0;
    for (int k = blockIdx.x; k < nchunks; k += gridDim.x, it++) {
        const int s = it & 1;
        jb_mbar_wait(&s_bar[s], (uint32_t)(it >> 1) & 1u);
        const Asm2Stage& S = stage[s];
        const int c0 = s_meta[s].c0, nr = s_meta[s].nr, hf0 = s_meta[s].hf0, row0 = s_meta[s].row0, flags = s_meta[s].flags;
        const int l8 = jb_span_lead<double>((size_t)hf0), l4 = jb_span_lead<int32_t>((size_t)hf0), l2 = jb_span_lead<uint16_t>((size_t)hf0);
        const int lc4 = jb_span_lead<int32_t>((size_t)c0), lc8 = jb_span_lead<double>((size_t)c0);
        const int j = warp * CPW + jl;
        if (j < nr) {
            const size_t c = (size_t)c0 + j;
            const int e0 = S.hfpos[lc4 + j] - hf0, e1 = S.hfpos[lc4 + j + 1] - hf0;
            const double2 sa = *reinterpret_cast<const double2*>(S.rec + 4 * j);
            const double ps = sa.x;
            const double Ss = a ? 1.0 - sa.y : sa.y;
            const double rs = S.rec[4 * j + 2 + a];
            const double pvc = S.pv[lc8 + j];
            // accumulation term (M - M0)/dt with its partials, sources on the diagonal entries (apply_forces!)
            double ar = (pvc * (rs * Ss) - S.M0[2 * j + a]) * inv_dt;
            if (flags & 1) ar += __ldg(src + 2 * c + a);
            double adp = (pvc * (ca * rs * Ss)) * inv_dt;
            double ads = (pvc * (rs * dS)) * inv_dt;
            for (int eb = e0; eb < e1; eb += NG) {
                double ro[NG][4];
#pragma unroll
                for (int u = 0; u < NG; u++)
                    if (eb + u < e1) ld_rec256(rec, S.other[l4 + eb + u], ro[u]);
#pragma unroll
                for (int u = 0; u < NG; u++) {
                    const int e = eb + u;
                    if (e < e1) {
                        const double So = a ? 1.0 - ro[u][1] : ro[u][1];
                        const double rn = a ? ro[u][3] : ro[u][2];
                        double F, dFdp, dFds, Bp, Bs;
                        flux_pair(ca, imu, dS, ps, Ss, rs, ro[u][0], So, rn, S.T[l8 + e], S.sg[l8 + e], F, dFdp, dFds, Bp, Bs);
                        ar += F;
                        if (JAC) {
                            adp += dFdp; ads += dFds;
                            // block (column-major): [Bp_w, Bp_o, Bs_w, Bs_o]; lane 0 stores the first pair, lane 1 the second
                            const double got = __shfl_xor_sync(pairmask, a ? Bp : Bs, 1);
                            double2* dst = reinterpret_cast<double2*>(nz + ((size_t)row0 + S.lp[l2 + e]) * 4) + a;
                            __stcs(dst, a ? make_double2(got, Bs) : make_double2(Bp, got));
                        }
                    }
                }
            }
            r[2 * c + a] = ar;
            if (JAC) {
                double* dst = nz + (size_t)S.diag[lc4 + j] * 4;
                __stcs(dst + a, adp);
                __stcs(dst + 2 + a, ads);
            }
        }
        __syncwarp();
        if (lane == 0) {
            __threadfence_block();
            const int old = atomicAdd(&s_cnt[s], 1);
            if (old == NWARPS - 1) {     // last warp done with stage s: refill it
                s_cnt[s] = 0;
                const int k2 = k + 2 * (int)gridDim.x;
                if (k2 < nchunks) issue(k2, s);
            }
        }
    }
}

// ---- variant B: face-parallel with atomics -----------------------------------------------
__global__ void __launch_bounds__(256) twophase_acc_kernel(i64 nc, TPParams P, const double4* __restrict__ rec, const double* __restrict__ pv,
                                                           const double* __restrict__ M0, const double* __restrict__ src, double dt,
                                                           const int32_t* __restrict__ diag_pos, double* __restrict__ nz,
                                                           double* __restrict__ r) {
    for (i64 c = (i64)blockIdx.x * blockDim.x + threadIdx.x; c < nc; c += (i64)gridDim.x * blockDim.x) {
        const double4 rs = ld_rec(rec + c);
        const double S[2] = {rs.y, 1.0 - rs.y};
        const double dS[2] = {1.0, -1.0};
        const double rho[2] = {rs.z, rs.w};
        const double pvc = __ldg(pv + c);
        double ra[2], da[4];
#pragma unroll
        for (int a = 0; a < 2; a++) {
            const double mass = pvc * (rho[a] * S[a]);
            ra[a] = (mass - __ldg(M0 + 2 * c + a)) / dt;
            if (src) ra[a] += __ldg(src + 2 * c + a);
            da[a] = (pvc * (P.c[a] * rho[a] * S[a])) / dt;
            da[2 + a] = (pvc * (rho[a] * dS[a])) / dt;
        }
        reinterpret_cast<double2*>(r)[c] = make_double2(ra[0], ra[1]);
        double2* dst = reinterpret_cast<double2*>(nz + (size_t)__ldg(diag_pos + c) * 4);
        dst[0] = make_double2(da[0], da[1]);
        dst[1] = make_double2(da[2], da[3]);
    }
}
__global__ void __launch_bounds__(256) twophase_faces_kernel(i64 nf, TPParams P, const int32_t* __restrict__ left, const int32_t* __restrict__ right,
                                                             const double* __restrict__ fT, const double* __restrict__ fgdz,
                                                             const int32_t* __restrict__ pos_lr, const int32_t* __restrict__ pos_rl,
                                                             const int32_t* __restrict__ diag_pos, const double4* __restrict__ rec,
                                                             double* __restrict__ nz, double* __restrict__ r) {
    for (i64 f = (i64)blockIdx.x * blockDim.x + threadIdx.x; f < nf; f += (i64)gridDim.x * blockDim.x) {
        const int32_t l = __ldg(left + f), rr = __ldg(right + f);
        const double T = __ldg(fT + f), g = __ldg(fgdz + f);
        const CellP L = cell_props(P, ld_rec(rec + l)), R = cell_props(P, ld_rec(rec + rr));
        double bl[4], br[4];
#pragma unroll
        for (int a = 0; a < 2; a++) {
            double Fl, dl_dp, dl_ds, Fr, dr_dp, dr_ds;
            flux_phase(P, L, R, a, T, g, Fl, dl_dp, dl_ds);     // out of left  (sign +1)
            flux_phase(P, R, L, a, T, -g, Fr, dr_dp, dr_ds);    // out of right (sign -1)
            warp_agg_atomic_add(r + 2 * (size_t)l + a, Fl);
            warp_agg_atomic_add(r + 2 * (size_t)rr + a, Fr);
            const size_t dl = (size_t)__ldg(diag_pos + l) * 4, dr = (size_t)__ldg(diag_pos + rr) * 4;
            warp_agg_atomic_add(nz + dl + a, dl_dp); warp_agg_atomic_add(nz + dl + 2 + a, dl_ds);
            warp_agg_atomic_add(nz + dr + a, dr_dp); warp_agg_atomic_add(nz + dr + 2 + a, dr_ds);
            bl[a] = -dr_dp; bl[2 + a] = -dr_ds;   // J[l, r] = -dF_{r->l}/dx_r
            br[a] = -dl_dp; br[2 + a] = -dl_ds;   // J[r, l] = -dF_{l->r}/dx_l
        }
        double2* d1 = reinterpret_cast<double2*>(nz + (size_t)__ldg(pos_lr + f) * 4);
        d1[0] = make_double2(bl[0], bl[1]); d1[1] = make_double2(bl[2], bl[3]);
        double2* d2 = reinterpret_cast<double2*>(nz + (size_t)__ldg(pos_rl + f) * 4);
        d2[0] = make_double2(br[0], br[1]); d2[1] = make_double2(br[2], br[3]);
    }
}

__global__ void __launch_bounds__(256) twophase_mass_kernel(i64 nc, TPParams P, const double* __restrict__ p, const double* __restrict__ s,
                                                            const double* __restrict__ pv, double* __restrict__ M) {
    for (i64 c = (i64)blockIdx.x * blockDim.x + threadIdx.x; c < nc; c += (i64)gridDim.x * blockDim.x) {
        const double pc = __ldg(p + c), sw = __ldg(s + 2 * c), pvc = __ldg(pv + c);
        const double rw = P.rho0[0] * exp(P.c[0] * (pc - P.p0));
        const double ro = P.rho0[1] * exp(P.c[1] * (pc - P.p0));
        reinterpret_cast<double2*>(M)[c] = make_double2(pvc * (rw * sw), pvc * (ro * (1.0 - sw)));
    }
}

__global__ void scatter_sources_kernel(i64 nsrc, const int32_t* __restrict__ cells, const double* __restrict__ vals, double* __restrict__ src) {
    // few entries; duplicates accumulate (serial per thread 0 keeps the order deterministic)
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (i64 k = 0; k < nsrc; k++) { src[2 * (size_t)cells[k]] += vals[2 * k]; src[2 * (size_t)cells[k] + 1] += vals[2 * k + 1]; }
}

// Chunk table of the TMA-staged kernel: cut [0, n_assemble) into chunks, record the 16-bit Jacobian slots, flag chunks
// holding a source cell, and order the table by min(own position, mean neighbour position) so that chunks that are each
// other's neighbours (e.g. the two colours of a multicolour numbering) are in flight together. Leaves asm2_ok = false
// when a cell has more half-faces than a stage holds (the lane-per-half-face kernel takes over).
static int build_asm2(jb_twophase* m) {
    jb_tpfa* t = m->t;
    jb_mesh* mesh = t->mesh;
    jb_ctx* ctx = mesh->ctx;
    m->asm2_ok = false;
    m->h_asm2.clear();
    const int32_t n_asm = (int32_t)(m->n_assemble >= 0 ? m->n_assemble : mesh->nc);
    std::vector<uint16_t> lp((size_t)mesh->nhf, 0);
    std::vector<char> has_src((size_t)mesh->nc, 0);
    for (int32_t c : m->h_src_cells) has_src[c] = 1;
    std::vector<double> key;
    int32_t start = 0;
    while (start < n_asm) {
        int32_t end = start;
        while (end < n_asm && end - start < JB_ASM2_CELLS && mesh->h_hf_pos[end + 1] - mesh->h_hf_pos[start] <= JB_ASM2_HF) end++;
        if (end == start) return JB_OK;
        Asm2Chunk ci;
        ci.c0 = start; ci.nr = end - start; ci.hf0 = mesh->h_hf_pos[start]; ci.cnt = mesh->h_hf_pos[end] - ci.hf0;
        ci.row0 = t->csr->h_rowptr[start]; ci.flags = 0; ci.pad0 = ci.pad1 = 0;
        double so = 0.0;
        for (int32_t i = ci.hf0; i < ci.hf0 + ci.cnt; i++) {
            const int64_t d = (int64_t)t->h_hf_rowpos[i] - ci.row0;
            if (d < 0 || d > 65535) return JB_OK;
            lp[i] = (uint16_t)d;
            so += mesh->h_hf_other[i];
        }
        for (int32_t c = start; c < end; c++) if (has_src[c]) ci.flags |= 1;
        const double mid = 0.5 * ((double)start + end);
        key.push_back(ci.cnt > 0 ? std::min(mid, so / ci.cnt) : mid);
        m->h_asm2.push_back(ci);
        start = end;
    }
    std::vector<int32_t> ord(m->h_asm2.size());
    for (size_t i = 0; i < ord.size(); i++) ord[i] = (int32_t)i;
    std::stable_sort(ord.begin(), ord.end(), [&](int32_t a, int32_t b) { return key[a] < key[b]; });
    std::vector<Asm2Chunk> tab(ord.size());
    for (size_t i = 0; i < ord.size(); i++) tab[i] = m->h_asm2[ord[i]];
    m->h_asm2.swap(tab);
    if (m->h_asm2.empty()) return JB_OK;
    if (m->d_asm2.upload(m->h_asm2, ctx->stream) != cudaSuccess || m->d_hf_lp.upload(lp, ctx->stream) != cudaSuccess)
        JB_FAIL(ctx, JB_ERR_ALLOC, "jb_twophase: device allocation failed (assembly chunk table)");
    m->asm2_ok = true;
    return JB_OK;
}

static int asm_variant() {   // JB_ASM_VARIANT: 0/unset = TMA-staged lane pairs, 3 = TMA-staged CTA form, 1 = stream, 2 = lane-per-half-face (read per call)
    const char* e = getenv("JB_ASM_VARIANT");
    return e ? atoi(e) : 0;
}

static TPParams make_params(const jb_twophase* m) {
    TPParams P;
    P.rho0[0] = m->params[0]; P.rho0[1] = m->params[1]; P.c[0] = m->params[2]; P.c[1] = m->params[3];
    P.mu[0] = m->params[4]; P.mu[1] = m->params[5]; P.p0 = m->params[6];
    P.inv_mu[0] = 1.0 / P.mu[0]; P.inv_mu[1] = 1.0 / P.mu[1];
    return P;
}
static int grid_for(jb_ctx* ctx, i64 n, int threads, int per_sm) {
    i64 want = (n + threads - 1) / threads;
    return (int)std::max<i64>(1, std::min<i64>(want, (i64)ctx->sm_count * per_sm));
}

int jb_launch_twophase_state(jb_twophase* m, const double* d_p, const double* d_s) {
    jb_ctx* ctx = m->t->mesh->ctx;
    const i64 nc = m->t->mesh->nc;
    ProfScope _ps(ctx, JB_PROF_STATE);
    twophase_state_kernel<<<grid_for(ctx, nc, 256, 16), 256, 0, ctx->stream>>>(nc, make_params(m), d_p, d_s, reinterpret_cast<double4*>(m->d_rec.p));
    JB_CHECK_LAUNCH(ctx);
    return JB_OK;
}

int jb_launch_twophase_assemble(jb_twophase* m, const double* d_M0, double dt, double* d_r, bool jac) {
    jb_tpfa* t = m->t;
    jb_ctx* ctx = t->mesh->ctx;
    const i64 nc = m->n_assemble >= 0 ? m->n_assemble : t->mesh->nc;   // owned rows only in a distributed run
    constexpr int LPC = 8;
    ProfScope _ps(ctx, JB_PROF_ASSEMBLY);
    const i64 cells_per_cta = 256 / LPC;
    const int grid = (int)std::max<i64>(1, (nc + cells_per_cta - 1) / cells_per_cta);
    const double* src = m->nsrc > 0 ? m->d_src.p : nullptr;
    if (jac) jb_csr_touch(t->csr);
    if (asm_variant() == 0 && m->asm2_ok && ((uintptr_t)d_M0 & 15) == 0) {
        // lane-pair-per-cell kernel; JB_ASM_NG picks (neighbour gathers in flight per lane, min CTAs per SM) for experiments
        const char* eng = getenv("JB_ASM_NG");
        const int ng = eng ? atoi(eng) : 36;
        const size_t smem = 2 * sizeof(Asm2Stage) + 2 * sizeof(Asm2Chunk) + 2 * sizeof(uint64_t) + 4 * sizeof(int);
        const int nchunks = (int)m->h_asm2.size();
        auto go = [&](auto kern) -> int {
            static std::map<const void*, int> occ;     // resident CTAs per SM of each instantiation (queried once)
            int& per_sm = occ[(const void*)kern];
            if (per_sm == 0) {
                cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, JB_ASM2_THREADS, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
            }
            const int g = std::max(1, std::min(nchunks, ctx->sm_count * per_sm));
            kern<<<g, JB_ASM2_THREADS, smem, ctx->stream>>>(nchunks, m->d_asm2.p, make_params(m), t->mesh->d_hf_pos.p, t->mesh->d_hf_other.p,
                                                            m->d_hf_lp.p, m->d_hf_T.p, m->d_hf_sgdz.p, t->csr->d_diag.p, m->d_rec.p, m->d_pv.p, d_M0, src,
                                                            1.0 / dt, t->csr->d_val.p, d_r);
            JB_CHECK_LAUNCH(ctx);
            return JB_OK;
        };
        if (!jac) return go(twophase_assemble_pair_kernel<false, 6, 4>);
        if (ng == 6) return go(twophase_assemble_pair_kernel<true, 6, 4>);
        if (ng == 35) return go(twophase_assemble_pair_kernel<true, 3, 5>);
        if (ng == 4) return go(twophase_assemble_pair_kernel<true, 4, 5>);
        if (ng == 27) return go(twophase_assemble_pair_kernel<true, 2, 7>);
        if (ng == 37) return go(twophase_assemble_pair_kernel<true, 3, 7>);
        return go(twophase_assemble_pair_kernel<true, 3, 6>);   // measured best at 10M cells (879 us; 6/4: 917, 4/5: 887, 3/5: 909)
    }
    if (asm_variant() == 3 && m->asm2_ok && ((uintptr_t)d_M0 & 15) == 0) {
        constexpr int LD = JB_ASM2_HF + 1;
        const int np = jac ? 6 : 2;
        const size_t smem = 2 * sizeof(Asm2Stage) + (size_t)(np * LD + (np * LD & 1)) * sizeof(double) + 2 * sizeof(Asm2Chunk) + 2 * sizeof(uint64_t) +
                            JB_ASM2_HF;
        static int per_sm2[2] = {0, 0};
        if (per_sm2[jac] == 0) {
            if (jac) {
                cudaFuncSetAttribute(twophase_assemble_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm2[1], twophase_assemble_tma_kernel<true>, JB_ASM2_THREADS, smem);
            } else {
                cudaFuncSetAttribute(twophase_assemble_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm2[0], twophase_assemble_tma_kernel<false>, JB_ASM2_THREADS, smem);
            }
            if (per_sm2[jac] < 1) per_sm2[jac] = 1;
        }
        const int nchunks = (int)m->h_asm2.size();
        const int g = std::max(1, std::min(nchunks, ctx->sm_count * per_sm2[jac]));
        if (jac)
            twophase_assemble_tma_kernel<true><<<g, JB_ASM2_THREADS, smem, ctx->stream>>>(
                nchunks, m->d_asm2.p, make_params(m), t->mesh->d_hf_pos.p, t->mesh->d_hf_other.p, m->d_hf_lp.p, m->d_hf_T.p, m->d_hf_sgdz.p,
                t->csr->d_diag.p, m->d_rec.p, m->d_pv.p, d_M0, src, 1.0 / dt, t->csr->d_val.p, d_r);
        else
            twophase_assemble_tma_kernel<false><<<g, JB_ASM2_THREADS, smem, ctx->stream>>>(
                nchunks, m->d_asm2.p, make_params(m), t->mesh->d_hf_pos.p, t->mesh->d_hf_other.p, m->d_hf_lp.p, m->d_hf_T.p, m->d_hf_sgdz.p,
                t->csr->d_diag.p, m->d_rec.p, m->d_pv.p, d_M0, src, 1.0 / dt, t->csr->d_val.p, d_r);
        JB_CHECK_LAUNCH(ctx);
        return JB_OK;
    }
    if (!m->h_chunks.empty() && asm_variant() != 2) {
        // chunks were cut over all local cells; a distributed run assembles the owned prefix only
        int nchunks = (int)m->h_chunks.size() - 1;
        if (m->n_assemble >= 0) nchunks = m->n_chunks_owned;
        const size_t smem = (size_t)(jac ? 6 : 2) * (JB_ASM_HF + 1) * sizeof(double);
        static int per_sm[2] = {0, 0};
        if (per_sm[jac] == 0) {
            if (jac) {
                cudaFuncSetAttribute(twophase_assemble_stream_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm[1], twophase_assemble_stream_kernel<true>, JB_ASM_THREADS, smem);
            } else {
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm[0], twophase_assemble_stream_kernel<false>, JB_ASM_THREADS, smem);
            }
            if (per_sm[jac] < 1) per_sm[jac] = 1;
        }
        const int g = std::max(1, std::min(nchunks, ctx->sm_count * per_sm[jac]));
        if (jac)
            twophase_assemble_stream_kernel<true><<<g, JB_ASM_THREADS, smem, ctx->stream>>>(nchunks, m->d_chunks.p, make_params(m), t->mesh->d_hf_pos.p,
                                                                                 t->mesh->d_hf_other.p, t->d_hf_rowpos.p, m->d_hf_T.p, m->d_hf_sgdz.p,
                                                                                 t->csr->d_diag.p, reinterpret_cast<const double4*>(m->d_rec.p),
                                                                                 m->d_pv.p, d_M0, src, 1.0 / dt, t->csr->d_val.p, d_r);
        else
            twophase_assemble_stream_kernel<false><<<g, JB_ASM_THREADS, smem, ctx->stream>>>(nchunks, m->d_chunks.p, make_params(m), t->mesh->d_hf_pos.p,
                                                                                  t->mesh->d_hf_other.p, t->d_hf_rowpos.p, m->d_hf_T.p, m->d_hf_sgdz.p,
                                                                                  t->csr->d_diag.p, reinterpret_cast<const double4*>(m->d_rec.p),
                                                                                  m->d_pv.p, d_M0, src, 1.0 / dt, t->csr->d_val.p, d_r);
        JB_CHECK_LAUNCH(ctx);
        return JB_OK;
    }
    if (jac)
        twophase_assemble_kernel<LPC, true><<<grid, 256, 0, ctx->stream>>>(nc, make_params(m), t->mesh->d_hf_pos.p, t->mesh->d_hf_other.p,
                                                                           t->d_hf_rowpos.p, m->d_hf_T.p, m->d_hf_sgdz.p, t->csr->d_diag.p,
                                                                           reinterpret_cast<const double4*>(m->d_rec.p), m->d_pv.p, d_M0, src,
                                                                           1.0 / dt, dt, t->csr->d_val.p, d_r);
    else
        twophase_assemble_kernel<LPC, false><<<grid, 256, 0, ctx->stream>>>(nc, make_params(m), t->mesh->d_hf_pos.p, t->mesh->d_hf_other.p,
                                                                            t->d_hf_rowpos.p, m->d_hf_T.p, m->d_hf_sgdz.p, t->csr->d_diag.p,
                                                                            reinterpret_cast<const double4*>(m->d_rec.p), m->d_pv.p, d_M0, src,
                                                                            1.0 / dt, dt, t->csr->d_val.p, d_r);
    JB_CHECK_LAUNCH(ctx);
    return JB_OK;
}

int jb_launch_twophase_assemble_faces(jb_twophase* m, const double* d_M0, double dt, double* d_r) {
    jb_tpfa* t = m->t;
    jb_ctx* ctx = t->mesh->ctx;
    const i64 nc = t->mesh->nc, nf = t->mesh->nf;
    const double* src = m->nsrc > 0 ? m->d_src.p : nullptr;
    jb_csr_touch(t->csr);
    ProfScope _ps(ctx, JB_PROF_ASSEMBLY);
    twophase_acc_kernel<<<grid_for(ctx, nc, 256, 16), 256, 0, ctx->stream>>>(nc, make_params(m), reinterpret_cast<const double4*>(m->d_rec.p),
                                                                             m->d_pv.p, d_M0, src, dt, t->csr->d_diag.p, t->csr->d_val.p, d_r);
    JB_CHECK_LAUNCH(ctx);
    twophase_faces_kernel<<<grid_for(ctx, nf, 256, 16), 256, 0, ctx->stream>>>(nf, make_params(m), t->mesh->d_left.p, t->mesh->d_right.p,
                                                                               m->d_face_T.p, m->d_face_gdz.p, m->d_pos_lr.p, m->d_pos_rl.p,
                                                                               t->csr->d_diag.p, reinterpret_cast<const double4*>(m->d_rec.p),
                                                                               t->csr->d_val.p, d_r);
    JB_CHECK_LAUNCH(ctx);
    return JB_OK;
}

int jb_launch_twophase_mass(jb_twophase* m, const double* d_p, const double* d_s, double* d_M) {
    jb_ctx* ctx = m->t->mesh->ctx;
    const i64 nc = m->t->mesh->nc;
    twophase_mass_kernel<<<grid_for(ctx, nc, 256, 16), 256, 0, ctx->stream>>>(nc, make_params(m), d_p, d_s, m->d_pv.p, d_M);
    JB_CHECK_LAUNCH(ctx);
    return JB_OK;
}

extern "C" {

int32_t jb_twophase_create(jb_tpfa* t, const double* Tf, const double* gdz, const double* pv, const double* params, jb_twophase** out) {
    if (!t || !Tf || !gdz || !pv || !params || !out) return JB_ERR_ARG;
    jb_mesh* mesh = t->mesh;
    jb_ctx* ctx = mesh->ctx;
    if (t->csr->bs != 2) JB_FAIL(ctx, JB_ERR_ARG, "jb_twophase_create: the two-phase law needs a Jacobian with 2x2 blocks");
    jb_twophase* m = new jb_twophase();
    m->t = t;
    for (int i = 0; i < 7; i++) m->params[i] = params[i];
    std::vector<double> hT(mesh->nhf), hG(mesh->nhf), fT(Tf, Tf + mesh->nf), fG(gdz, gdz + mesh->nf), hpv(pv, pv + mesh->nc);
    for (i64 i = 0; i < mesh->nhf; i++) {
        const int32_t f = mesh->h_hf_face[i];
        hT[i] = Tf[f];
        hG[i] = (double)mesh->h_hf_sign[i] * gdz[f];
    }
    // per-face Jacobian positions for the face-parallel variant
    std::vector<int32_t> plr(mesh->nf), prl(mesh->nf);
    for (i64 c = 0; c < mesh->nc; c++)
        for (int32_t i = mesh->h_hf_pos[c]; i < mesh->h_hf_pos[c + 1]; i++) {
            const int32_t f = mesh->h_hf_face[i];
            if (mesh->h_hf_sign[i] > 0) plr[f] = t->h_hf_rowpos[i]; else prl[f] = t->h_hf_rowpos[i];
        }
    // row chunks for the stream assembly kernel (<= JB_ASM_CELLS cells, <= JB_ASM_HF half-faces)
    {
        int32_t start = 0;
        bool fits = true;
        const int32_t ncell = (int32_t)mesh->nc;
        while (start < ncell) {
            int32_t end = start;
            while (end < ncell && end - start < JB_ASM_CELLS && mesh->h_hf_pos[end + 1] - mesh->h_hf_pos[start] <= JB_ASM_HF) end++;
            if (end == start) { fits = false; break; }
            m->h_chunks.push_back(start);
            start = end;
        }
        if (fits) m->h_chunks.push_back(ncell); else m->h_chunks.clear();
    }
    cudaStream_t s = ctx->stream;
    if (!m->h_chunks.empty() && m->d_chunks.upload(m->h_chunks, s) != cudaSuccess) { delete m; JB_FAIL(ctx, JB_ERR_ALLOC, "jb_twophase_create: device allocation failed"); }
    bool ok = m->d_hf_T.upload(hT, s) == cudaSuccess && m->d_hf_sgdz.upload(hG, s) == cudaSuccess && m->d_face_T.upload(fT, s) == cudaSuccess &&
              m->d_face_gdz.upload(fG, s) == cudaSuccess && m->d_pv.upload(hpv, s) == cudaSuccess && m->d_pos_lr.upload(plr, s) == cudaSuccess &&
              m->d_pos_rl.upload(prl, s) == cudaSuccess && m->d_rec.alloc((size_t)mesh->nc * 4) == cudaSuccess;
    if (!ok) { delete m; JB_FAIL(ctx, JB_ERR_ALLOC, "jb_twophase_create: device allocation failed"); }
    { const int rc = build_asm2(m); if (rc != JB_OK) { delete m; return rc; } }
    *out = m;
    return JB_OK;
}
int32_t jb_twophase_destroy(jb_twophase* m) { delete m; return JB_OK; }
int32_t jb_twophase_set_owned(jb_twophase* m, int64_t n_owned) {
    if (!m || n_owned > m->t->mesh->nc) return JB_ERR_ARG;
    m->n_assemble = n_owned;
    if (n_owned >= 0 && !m->h_chunks.empty()) {
        // re-cut so that a chunk boundary falls exactly on n_owned
        jb_mesh* mesh = m->t->mesh;
        m->h_chunks.clear();
        const int32_t bounds[2] = {(int32_t)n_owned, (int32_t)mesh->nc};
        int32_t start = 0;
        bool fits = true;
        m->n_chunks_owned = 0;
        for (int part = 0; part < 2 && fits; part++) {
            while (start < bounds[part]) {
                int32_t end = start;
                while (end < bounds[part] && end - start < JB_ASM_CELLS && mesh->h_hf_pos[end + 1] - mesh->h_hf_pos[start] <= JB_ASM_HF) end++;
                if (end == start) { fits = false; break; }
                m->h_chunks.push_back(start);
                start = end;
            }
            if (part == 0) m->n_chunks_owned = (int)m->h_chunks.size();
        }
        if (fits) m->h_chunks.push_back((int32_t)mesh->nc); else m->h_chunks.clear();
        if (!m->h_chunks.empty() && m->d_chunks.upload(m->h_chunks, mesh->ctx->stream) != cudaSuccess) return JB_ERR_ALLOC;
    }
    return build_asm2(m);
}

int32_t jb_twophase_set_sources(jb_twophase* m, int64_t nsrc, const int64_t* cells, const double* vals) {
    if (!m || nsrc < 0 || (nsrc > 0 && (!cells || !vals))) return JB_ERR_ARG;
    jb_ctx* ctx = m->t->mesh->ctx;
    const i64 nc = m->t->mesh->nc;
    m->nsrc = nsrc;
    if (nsrc == 0) { m->h_src_cells.clear(); return build_asm2(m); }
    std::vector<int32_t> hc(nsrc);
    for (i64 k = 0; k < nsrc; k++) {
        if (cells[k] < 1 || cells[k] > nc) JB_FAIL(ctx, JB_ERR_ARG, "jb_twophase_set_sources: cell out of range");
        hc[k] = (int32_t)(cells[k] - 1);
    }
    std::vector<double> hv(vals, vals + 2 * nsrc);
    if (m->d_src_cells.upload(hc, ctx->stream) != cudaSuccess || m->d_src_vals.upload(hv, ctx->stream) != cudaSuccess ||
        (m->d_src.n != (size_t)2 * nc && m->d_src.alloc((size_t)2 * nc) != cudaSuccess))
        JB_FAIL(ctx, JB_ERR_ALLOC, "jb_twophase_set_sources: device allocation failed");
    JB_CUDA(ctx, cudaMemsetAsync(m->d_src.p, 0, (size_t)2 * nc * sizeof(double), ctx->stream));
    scatter_sources_kernel<<<1, 32, 0, ctx->stream>>>(nsrc, m->d_src_cells.p, m->d_src_vals.p, m->d_src.p);
    JB_CHECK_LAUNCH(ctx);
    JB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    m->h_src_cells = hc;
    return build_asm2(m);
}

int32_t jb_twophase_update_state(jb_twophase* m, const double* d_p, const double* d_s) {
    if (!m || !d_p || !d_s) return JB_ERR_ARG;
    int rc = jb_launch_twophase_state(m, d_p, d_s);
    if (rc != JB_OK) return rc;
    JB_CUDA(m->t->mesh->ctx, cudaStreamSynchronize(m->t->mesh->ctx->stream));
    return JB_OK;
}
int32_t jb_twophase_mass(jb_twophase* m, const double* d_p, const double* d_s, double* d_M) {
    if (!m || !d_p || !d_s || !d_M) return JB_ERR_ARG;
    int rc = jb_launch_twophase_mass(m, d_p, d_s, d_M);
    if (rc != JB_OK) return rc;
    JB_CUDA(m->t->mesh->ctx, cudaStreamSynchronize(m->t->mesh->ctx->stream));
    return JB_OK;
}
static int32_t assemble_common(jb_twophase* m, const double* d_p, const double* d_s, const double* d_M0, double dt, double* d_r, int variant) {
    if (!m || !d_p || !d_s || !d_M0 || !d_r || !(dt > 0)) return JB_ERR_ARG;
    jb_ctx* ctx = m->t->mesh->ctx;
    int rc = jb_launch_twophase_state(m, d_p, d_s);
    if (rc != JB_OK) return rc;
    if (variant == 0) rc = jb_launch_twophase_assemble(m, d_M0, dt, d_r, true);
    else if (variant == 1) rc = jb_launch_twophase_assemble_faces(m, d_M0, dt, d_r);
    else rc = jb_launch_twophase_assemble(m, d_M0, dt, d_r, false);
    if (rc != JB_OK) return rc;
    JB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return JB_OK;
}
int32_t jb_twophase_assemble(jb_twophase* m, const double* d_p, const double* d_s, const double* d_M0, double dt, double* d_r) {
    return assemble_common(m, d_p, d_s, d_M0, dt, d_r, 0);
}
int32_t jb_twophase_assemble_faces(jb_twophase* m, const double* d_p, const double* d_s, const double* d_M0, double dt, double* d_r) {
    return assemble_common(m, d_p, d_s, d_M0, dt, d_r, 1);
}
int32_t jb_twophase_residual(jb_twophase* m, const double* d_p, const double* d_s, const double* d_M0, double dt, double* d_r) {
    return assemble_common(m, d_p, d_s, d_M0, dt, d_r, 2);
}

}  // extern "C"
