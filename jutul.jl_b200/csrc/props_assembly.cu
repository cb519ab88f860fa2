// Two-phase conservation law with property values supplied by the secondary-variable graph (variables.cu) instead of the
// closed forms built into assembly.cu: the step JutulDarcy-style models take between update_secondary_variables!
// (src/variable_evaluation.jl:87-148) and update_equation! / fill_conservation_eq! (src/conservation/conservation.jl:558-626,
// 373-430). Per cell and phase a the caller hands over, each as value plane + d/dp plane + d/dSw plane (the layout
// jb_varprog_evaluate writes): the conserved mass M_a, the mass density rho_a (gravity term) and the mobility
// lambda_a = rho_a k_ra / mu_a. The law itself is unchanged (two_point_potential_drop, upw_flux; src/conservation/flux.jl:335-435):
//   theta = p_self - p_other + sign*g*dz * (rho_self + rho_other)/2,  q = T theta,  F_a = lambda_a|upstream * q,
//   r_a,c = (M_a - M0_a)/dt + sum_hf F_a,  J[c,c] = d(acc) + sum dF/dx_c,  J[c,o] = -dF_{o->c}/dx_o,
// with every property partial taken from the planes (chain rule through whatever tables / power laws produced them).
// Schedule: a pack kernel writes one 128-byte record per cell, a 64-byte half per phase {p, rho, lambda with partials}, so that a
// neighbour gather is two 256-bit loads of one aligned line; lane pair (2j, 2j+1) owns cell j, lane a evaluates phase a of every half-face in conn_pos
// order with register sums (the summation order of fill_conservation_eq!), the pair writes each 2x2 off-diagonal block with two
// 16-byte stores. Row-owner form: no atomics, every entry written once.
#include "jb_internal.cuh"

__global__ void __launch_bounds__(256) props_pack_kernel(i64 nc, const double* __restrict__ p, const double* __restrict__ rw, const double* __restrict__ ro,
                                                         const double* __restrict__ mw, const double* __restrict__ mo, double* __restrict__ rec) {
    for (i64 c = (i64)blockIdx.x * blockDim.x + threadIdx.x; c < nc; c += (i64)gridDim.x * blockDim.x) {
        // one 64-byte half per phase: {p, rho, rho_p, rho_s | lambda, lambda_p, lambda_s, -}: a lane reads its phase with two 256-bit loads
        double* o = rec + 16 * (size_t)c;
        const double pc = __ldg(p + c);
        reinterpret_cast<double4*>(o)[0] = make_double4(pc, __ldg(rw + c), __ldg(rw + nc + c), __ldg(rw + 2 * nc + c));
        reinterpret_cast<double4*>(o)[1] = make_double4(__ldg(mw + c), __ldg(mw + nc + c), __ldg(mw + 2 * nc + c), 0.0);
        reinterpret_cast<double4*>(o)[2] = make_double4(pc, __ldg(ro + c), __ldg(ro + nc + c), __ldg(ro + 2 * nc + c));
        reinterpret_cast<double4*>(o)[3] = make_double4(__ldg(mo + c), __ldg(mo + nc + c), __ldg(mo + 2 * nc + c), 0.0);
    }
}

struct PhaseRec { double p, rho, rho_p, rho_s, mob, mob_p, mob_s; };
__device__ __forceinline__ PhaseRec load_phase(const double* __restrict__ rec, size_t c, int a) {
    const double* r = rec + 16 * c + 8 * a;
    double u[4], v[4];
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(u[0]), "=d"(u[1]), "=d"(u[2]), "=d"(u[3]) : "l"(r));
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(r + 4));
    PhaseRec o;
    o.p = u[0]; o.rho = u[1]; o.rho_p = u[2]; o.rho_s = u[3];
    o.mob = v[0]; o.mob_p = v[1]; o.mob_s = v[2];
    return o;
}

template <bool JAC>
__global__ void __launch_bounds__(256) twophase_props_kernel(i64 nc, const int32_t* __restrict__ hf_pos, const int32_t* __restrict__ hf_other,
                                                             const int32_t* __restrict__ hf_rowpos, const double* __restrict__ hf_T,
                                                             const double* __restrict__ hf_sgdz, const int32_t* __restrict__ diag_pos,
                                                             const double* __restrict__ rec, const double* __restrict__ massw,
                                                             const double* __restrict__ masso, i64 nc_planes, const double* __restrict__ M0,
                                                             const double* __restrict__ src, double inv_dt, double* __restrict__ nz,
                                                             double* __restrict__ r) {
    const int a = threadIdx.x & 1;
    const unsigned pairmask = 3u << (threadIdx.x & 30);
    for (i64 c = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 1; c < nc; c += ((i64)gridDim.x * blockDim.x) >> 1) {
        const PhaseRec S = load_phase(rec, (size_t)c, a);
        const double* mass = a ? masso : massw;
        // accumulation term (M - M0)/dt and its partials; sources on the diagonal entries (apply_forces!)
        double ar = (__ldg(mass + c) - __ldg(M0 + 2 * c + a)) * inv_dt;
        if (src) ar += __ldg(src + 2 * c + a);
        double adp = __ldg(mass + nc_planes + c) * inv_dt;
        double ads = __ldg(mass + 2 * nc_planes + c) * inv_dt;
        for (int32_t i = __ldg(hf_pos + c); i < __ldg(hf_pos + c + 1); i++) {
            const int32_t o = __ldg(hf_other + i);
            const PhaseRec O = load_phase(rec, (size_t)o, a);
            const double T = __ldg(hf_T + i), sg = __ldg(hf_sgdz + i);
            const double rho_avg = 0.5 * (S.rho + O.rho);
            const double theta = S.p - O.p + sg * rho_avg;
            const double q = T * theta;
            const bool ups = q > 0, upo = q < 0;
            const double mF = ups ? S.mob : O.mob;        // upstream mobility seen from self (q == 0 / NaN: other, as upw_flux)
            const double mN = upo ? O.mob : S.mob;        // upstream mobility seen from the neighbour
            ar += mF * q;
            if (JAC) {
                const double dq_ps = T * (1.0 + sg * (0.5 * S.rho_p)), dq_ss = T * (sg * (0.5 * S.rho_s));
                double dFdp = mF * dq_ps, dFds = mF * dq_ss;
                if (ups) { dFdp += q * S.mob_p; dFds += q * S.mob_s; }
                adp += dFdp; ads += dFds;
                // the neighbour's half-face towards us: q' = -q, J[self, other] = -dF_{o->c}/dx_o
                const double dqn_po = T * (1.0 + (-sg) * (0.5 * O.rho_p)), dqn_so = T * ((-sg) * (0.5 * O.rho_s));
                double np = mN * dqn_po, ns = mN * dqn_so;
                if (upo) { const double qo = -q; np += qo * O.mob_p; ns += qo * O.mob_s; }
                const double Bp = -np, Bs = -ns;
                const double got = __shfl_xor_sync(pairmask, a ? Bp : Bs, 1);
                double2* dst = reinterpret_cast<double2*>(nz + (size_t)__ldg(hf_rowpos + i) * 4) + a;
                *dst = a ? make_double2(got, Bs) : make_double2(Bp, got);
            }
        }
        r[2 * c + a] = ar;
        if (JAC) {
            double* dst = nz + (size_t)__ldg(diag_pos + c) * 4;
            dst[a] = adp; dst[2 + a] = ads;
        }
    }
}

extern "C" int32_t jb_twophase_assemble_props(jb_twophase* m, const double* d_p, const double* const* d_props, const double* d_M0, double dt,
                                              double* d_r) {
    if (!m || !d_p || !d_props || !d_M0 || !d_r || !(dt > 0.0)) return JB_ERR_ARG;
    for (int q = 0; q < 6; q++) if (!d_props[q]) return JB_ERR_ARG;
    jb_tpfa* t = m->t;
    jb_ctx* ctx = t->mesh->ctx;
    const i64 nc_all = t->mesh->nc;
    const i64 nc = m->n_assemble >= 0 ? m->n_assemble : nc_all;   // owned rows only in a distributed run
    if (m->d_rec16.n != (size_t)16 * nc_all && m->d_rec16.alloc((size_t)16 * nc_all) != cudaSuccess) JB_FAIL(ctx, JB_ERR_ALLOC, "jb_twophase_assemble_props: allocation failed");
    const double* src = m->nsrc > 0 ? m->d_src.p : nullptr;
    const int g1 = (int)std::max<i64>(1, std::min<i64>((nc_all + 255) / 256, (i64)ctx->sm_count * 16));
    const int g2 = (int)std::max<i64>(1, std::min<i64>((2 * nc + 255) / 256, (i64)ctx->sm_count * 16));
    {
        ProfScope _ps(ctx, JB_PROF_STATE);
        props_pack_kernel<<<g1, 256, 0, ctx->stream>>>(nc_all, d_p, d_props[2], d_props[3], d_props[4], d_props[5], m->d_rec16.p);
        JB_CHECK_LAUNCH(ctx);
    }
    {
        ProfScope _ps(ctx, JB_PROF_ASSEMBLY);
        twophase_props_kernel<true><<<g2, 256, 0, ctx->stream>>>(nc, t->mesh->d_hf_pos.p, t->mesh->d_hf_other.p, t->d_hf_rowpos.p, m->d_hf_T.p,
                                                                m->d_hf_sgdz.p, t->csr->d_diag.p, m->d_rec16.p, d_props[0], d_props[1], nc_all, d_M0,
                                                                src, 1.0 / dt, t->csr->d_val.p, d_r);
        JB_CHECK_LAUNCH(ctx);
    }
    jb_csr_touch(t->csr);
    JB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return JB_OK;
}
