// Device-resident scalar block of the Krylov solvers. Kernels read and update it
// in place so that an iteration needs no host round trip.
#pragma once
enum {
    KS_RHO = 0,       // rho_k = <c, r_{k-1}>
    KS_ALPHA = 1,
    KS_OMEGA = 2,
    KS_BETA = 3,
    KS_RNORM = 4,
    KS_EPS = 5,       // atol + rtol*|r0|
    KS_R0 = 6,
    KS_DONE = 7,      // != 0: all later kernels of this solve are no-ops
    KS_STATUS = 8,    // 0 solved, 1 running/itmax, 2 breakdown
    KS_ITER = 9,
    KS_ABS_TOL = 10,  // manual (min_it > 1) criterion
    KS_REL_TOL = 11,
    KS_MIN_IT = 12,
    KS_MANUAL = 13,
    KS_ITMAX = 14,
    KS_RTOL = 15,
    KS_ATOL = 16,
    KS_SUM0 = 17,     // raw inner products of the last reducing kernel (local, then all-reduced when distributed)
    KS_SUM1 = 18,
    KS_DIST = 19,     // != 0: reductions are finalised by krylov_finalize_kernel after the NCCL all-reduce
    KS_SIZE = 32
};

#ifdef __CUDACC__
// Scalar recurrences of Krylov.jl's bicgstab!, evaluated by one thread once the inner products are complete.
__device__ __forceinline__ void ks_fin_alpha(double* sc) { sc[KS_ALPHA] = sc[KS_RHO] / sc[KS_SUM0]; }      // alpha = rho / <c, v>
__device__ __forceinline__ void ks_fin_omega(double* sc) { sc[KS_OMEGA] = sc[KS_SUM0] / sc[KS_SUM1]; }    // omega = <t,s> / <t,t>
__device__ __forceinline__ void ks_fin_init(double* sc, double* hist) {     // SUM0 = <r,r>, SUM1 = <c,r>
    const double rnorm = sqrt(sc[KS_SUM0]);
    sc[KS_RNORM] = rnorm; sc[KS_R0] = rnorm;
    sc[KS_EPS] = sc[KS_ATOL] + sc[KS_RTOL] * rnorm;
    sc[KS_RHO] = sc[KS_SUM1];
    sc[KS_ALPHA] = 1.0; sc[KS_OMEGA] = 1.0; sc[KS_BETA] = 0.0;
    sc[KS_ITER] = 0.0;
    hist[0] = rnorm;
    double done = 0.0, status = 1.0;
    if (rnorm == 0.0) { done = 1.0; status = 0.0; }                  // x = 0 is the solution
    else if (sc[KS_SUM1] == 0.0) { done = 1.0; status = 2.0; }       // "Breakdown b'c = 0"
    else if (rnorm <= sc[KS_EPS]) { done = 1.0; status = 0.0; }
    else if (sc[KS_ITMAX] <= 0.0) { done = 1.0; status = 1.0; }
    sc[KS_DONE] = done; sc[KS_STATUS] = status;
}
__device__ __forceinline__ void ks_fin_update2(double* sc, double* hist, int hist_cap) {   // SUM0 = <c,r>, SUM1 = <r,r>
    const double alpha = sc[KS_ALPHA], om = sc[KS_OMEGA], rho = sc[KS_RHO];
    const double next_rho = sc[KS_SUM0];
    sc[KS_BETA] = (next_rho / rho) * (alpha / om);
    sc[KS_RHO] = next_rho;
    const double rnorm = sqrt(sc[KS_SUM1]);
    sc[KS_RNORM] = rnorm;
    const double iter = sc[KS_ITER] + 1.0;
    sc[KS_ITER] = iter;
    if ((int)iter < hist_cap) hist[(int)iter] = rnorm;
    const bool mach = (rnorm + 1.0 <= 1.0);
    const bool solved = (rnorm <= sc[KS_EPS]) || mach;
    bool user_exit = false;
    if (sc[KS_MANUAL] != 0.0)   // krylov_termination_criterion (src/linsolve/krylov.jl:198-205)
        user_exit = (rnorm <= sc[KS_ABS_TOL] + sc[KS_REL_TOL] * sc[KS_R0]) && (iter + 1.0 > sc[KS_MIN_IT]);
    const bool tired = iter >= sc[KS_ITMAX];
    const bool breakdown = (alpha == 0.0) || isnan(alpha);
    if (solved || user_exit) { sc[KS_DONE] = 1.0; sc[KS_STATUS] = 0.0; }
    else if (breakdown) { sc[KS_DONE] = 1.0; sc[KS_STATUS] = 2.0; }
    else if (tired) { sc[KS_DONE] = 1.0; sc[KS_STATUS] = 1.0; }
}
#endif
